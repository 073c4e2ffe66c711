"""Multi-GPU check (torchrun, one rank per GPU): the weight arena broadcast from rank 0 makes every rank compute the same
result as rank 0 computes from the file, for the fp16 and for the quantised network.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from infur_b200 import processors as P
    from infur_b200 import quantize, sharding, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        synth.ensure_fixture("fcn_tiny")
        quantize.ensure_fixture("fcn_tiny_int8")
    dist.barrier()
    frames = np.stack([synth.synth_frame(320, 240, i) for i in range(2 * world)])
    ok = True
    for kind in ("fcn_tiny", "fcn_tiny_int8"):
        path = synth.fixture_path(kind)
        with P.Handle(device=local, max_batch=2) as h:
            sharding.load_model_sharded(h, path, rank, world, dev)
            ids = sharding.shard(range(1, 2 * world + 1), rank, world)
            res = h.advance_batch(np.ascontiguousarray(frames[[i - 1 for i in ids]]), ids=ids, want=("class_map", "decoded_rgba"))
            mine = {i: (zlib.crc32(r["class_map"].tobytes()), zlib.crc32(r["decoded_rgba"].tobytes())) for i, r in zip(ids, res)}
        allr = sharding.gather_ordered(mine)
        if rank == 0:
            with P.Handle(device=local, max_batch=2 * world) as h:   # the whole stream on one GPU, weights straight from the file
                h.model_load(path)
                ref = h.advance_batch(frames, ids=list(range(1, 2 * world + 1)), want=("class_map", "decoded_rgba"))
            want = [(zlib.crc32(r["class_map"].tobytes()), zlib.crc32(r["decoded_rgba"].tobytes())) for r in ref]
            same = allr == want
            ok &= same
            print(f"{kind}: {2 * world} frames over {world} ranks {'identical to' if same else 'DIFFER from'} the single-GPU result", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
