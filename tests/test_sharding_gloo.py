"""CPU, world_size 2, gloo: the N > 1 host logic (frame sharding by id, the weight-blob broadcast, ordered re-assembly).
The per-frame work here is the oracle's ColorCode on a tiny synthetic logit map, so the test also proves that sharded
results are identical to the single-process ones."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from infur_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _frame_result(frame_id: int) -> int:
    rng = np.random.default_rng(frame_id)
    hm = rng.standard_normal((21, 6, 8)).astype(np.float32)
    k, rgba = oracle.color_code_image(hm)
    return int(k.sum()) * 1000003 + int(rgba.astype(np.int64).sum())


def _worker(rank, world, port, n_frames, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # weight blob: only rank 0 has the real bytes
        blob = torch.arange(4099, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.zeros(4099, dtype=torch.uint8)
        sharding.broadcast_blob(blob, 0)
        assert bool((blob == torch.arange(4099, dtype=torch.int64).to(torch.uint8)).all())
        mine = sharding.shard(range(1, n_frames + 1), rank, world)
        assert all(sharding.owner_rank(i, world) == rank for i in mine)
        for b in sharding.batches(mine, 8):
            assert 1 <= len(b) <= 8
        local = {i: _frame_result(i) for i in mine}
        ordered = sharding.gather_ordered(local)
        if rank == 0:
            out.put(ordered)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    n_frames, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    ordered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ordered == [_frame_result(i) for i in range(1, n_frames + 1)]


def test_shard_partition_properties():
    ids = list(range(1, 101))
    for world in (1, 2, 4, 8):
        parts = [sharding.shard(ids, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == ids                      # a partition: every frame exactly once
        assert max(map(len, parts)) - min(map(len, parts)) <= 1   # balanced
        assert all(p == sorted(p) for p in parts)                 # stream order kept inside a rank
    assert sharding.gather_ordered({3: "c", 1: "a", 2: "b"}) == ["a", "b", "c"]
