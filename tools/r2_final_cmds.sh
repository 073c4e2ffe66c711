# Final verification pass of round 2 (gpurun -- 'bash tools/r2_final_cmds.sh'): tests, smoke, sanitizers, bench, per-kernel tables.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/r2_compute_sanitizer_memcheck.txt 2>&1
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_fused.py > gpurun_out/r2_compute_sanitizer_racecheck.txt 2>&1
timeout 400 compute-sanitizer --tool synccheck python tools/sanitize_fused.py fcn50 > gpurun_out/r2_compute_sanitizer_synccheck.txt 2>&1
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
python tools/profile_step.py --iters 5 > gpurun_out/r2_per_kernel_events_f16_b8_1080p.txt 2>&1
python tools/profile_step.py --kind fcn50_int8 --iters 5 > gpurun_out/r2_per_kernel_events_int8_b8_1080p.txt 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> /dev/null
tail -n 2 gpurun_out/final_tests.txt gpurun_out/final_smoke.txt
tail -n 2 gpurun_out/r2_compute_sanitizer_memcheck.txt gpurun_out/r2_compute_sanitizer_racecheck.txt gpurun_out/r2_compute_sanitizer_synccheck.txt
tail -n 2 gpurun_out/r2_per_kernel_events_f16_b8_1080p.txt gpurun_out/r2_per_kernel_events_int8_b8_1080p.txt
