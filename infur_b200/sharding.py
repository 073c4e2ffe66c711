"""Host-side multi-GPU plumbing: one process per GPU, frames sharded by id, no collective on the frame path.

Frames are independent units of work (``Scale`` / ``Model`` / ``ColorCode`` keep no temporal state;
infur/src/app.rs:107-153), ids are 1-based and consecutive (ff-video/src/decoder.rs:163-164), so frame ``id``
belongs to rank ``(id - 1) % world``.  The only collective is the broadcast of the packed weight arena from
rank 0 when a model is loaded; results are re-ordered by id on the host.

Works with any ``torch.distributed`` backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence


def owner_rank(frame_id: int, world: int) -> int:
    """Rank that processes frame ``frame_id`` (1-based ids, round-robin)."""
    if frame_id < 1:
        raise ValueError("frame ids are 1-based (ff-video/src/decoder.rs:163-164)")
    return (frame_id - 1) % world


def shard(frame_ids: Iterable[int], rank: int, world: int) -> List[int]:
    """The ids of ``frame_ids`` this rank owns, in stream order."""
    return [i for i in frame_ids if owner_rank(i, world) == rank]


def batches(ids: Sequence[int], batch: int) -> List[List[int]]:
    """Consecutive groups of at most ``batch`` ids: one pinned ring slot each."""
    return [list(ids[i:i + batch]) for i in range(0, len(ids), batch)]


def broadcast_blob(blob, src: int = 0):
    """Broadcast a byte tensor (the packed weight arena) from ``src`` to every rank, in place."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src)
    return blob


def load_model_sharded(handle, path: str, rank: int, world: int, device):
    """Model::control(Load) on every rank with ONE upload: rank 0 parses, packs and uploads the weights; the other
    ranks parse the graph only (``INFUR_LOAD_SKIP_WEIGHTS``) and receive the packed arena over the collective."""
    import torch

    handle.model_load(path, skip_weights=(rank != 0))
    if world > 1:
        nbytes = handle.weights_size()
        blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        if rank == 0:
            handle.weights_export(blob.data_ptr(), nbytes)
        torch.cuda.synchronize()
        broadcast_blob(blob, 0)
        torch.cuda.synchronize()
        if rank != 0:
            handle.weights_import(blob.data_ptr(), nbytes)
        del blob


def gather_ordered(local: Dict[int, object]) -> List[object]:
    """All ranks' ``{frame id: result}`` maps merged and returned in id order (on every rank)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, local)
    else:
        parts = [local]
    merged: Dict[int, object] = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise ValueError(f"frame {k} was processed by two ranks")
            merged[k] = v
    return [merged[k] for k in sorted(merged)]
