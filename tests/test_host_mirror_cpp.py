"""The C++ host mirror (include/infur_b200_processors.hpp: trait Processor, Scale, Model, ColorCode, GpuPipeline with the
reference's names and error behaviour) and the reference's own unit tests ported onto it (infur_b200/host/host_check.cpp).
CPU: the program builds, links against the C ABI and reports the missing device loudly.  GPU: every ported test passes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "host_check")


def _build():
    from infur_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "infur_b200", "host")], check=True, capture_output=True)


def test_host_check_builds_and_fails_loudly_without_gpu():
    import torch

    _build()
    p = subprocess.run([EXE, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    if not torch.cuda.is_available():
        assert "code 10" in p.stdout and "no CPU fallback" in p.stdout   # INFUR_E_NO_DEVICE


@pytest.mark.gpu
def test_reference_unit_tests_on_cpp_mirror(tiny):
    _build()
    path, _ = tiny
    p = subprocess.run([EXE, path], capture_output=True, text=True, timeout=600)
    print(p.stdout)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "all reference tests passed" in p.stdout and "FAIL" not in p.stdout
