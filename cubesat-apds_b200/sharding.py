"""Host-side logic of the sharded matcher: contiguous row-range shards, the all-gather
of per-shard top-2 records, and which rank owns which frame.  Pure bookkeeping (no numerics): the
local top-2 and the merge are CUDA kernels (`dunk_db_knn2_dev`, `dunk_top2_merge_dev`)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from ._lib import TOP2_DTYPE


def shard_ranges(n_rows: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) global row range of every rank: contiguous, sizes differ by at most one."""
    cuts = [n_rows * r // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def frame_partition(n_frames: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) frame range of every rank (extraction and RANSAC partition by frame batch)."""
    return shard_ranges(n_frames, world)


def all_gather_top2(local: np.ndarray, group=None) -> np.ndarray:
    """All-gather of per-shard top-2 records through torch.distributed (gloo on CPU, NCCL on GPU
    tensors in bench.py).  local: (nq,) TOP2_DTYPE -> (world, nq) TOP2_DTYPE, part-major — the
    layout `dunk_top2_merge_dev` expects."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).reshape(-1).copy())
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.stack([o.numpy().view(TOP2_DTYPE) for o in out])


def merge_top2_host(parts: np.ndarray) -> np.ndarray:
    """Reference merge for tests of the plumbing: lexicographic (distance, index) over all parts."""
    p = np.asarray(parts)
    d = np.concatenate([p["d1"], p["d2"]], axis=0).astype(np.uint64)      # (2*world, nq)
    i = np.concatenate([p["i1"], p["i2"]], axis=0).astype(np.uint64)
    key = (d << np.uint64(32)) | i
    order = np.argsort(key, axis=0, kind="stable")[:2]
    k = np.take_along_axis(key, order, axis=0)
    out = np.empty(p.shape[1], dtype=TOP2_DTYPE)
    out["d1"], out["i1"] = (k[0] >> np.uint64(32)).astype(np.uint32), (k[0] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    out["d2"], out["i2"] = (k[1] >> np.uint64(32)).astype(np.uint32), (k[1] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return out
