"""CPU, world_size 2 over gloo: the host plumbing of the sharded matcher (row-range shards, global
index bases, all-gather layout, lexicographic merge) reproduces the unsharded cv2 golden result.
The per-shard top-2 is computed by the oracle here (no GPU); on a GPU box the same plumbing feeds
`dunk_db_knn2_dev` / `dunk_top2_merge_dev` (tests/test_match_gpu.py, bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cubesat_apds_b200 as dunk
    from cubesat_apds_b200 import sharding
    from oracle import match_oracle as mo
    g = np.load(os.path.join(ROOT, "tests", "golden", "match_golden.npz"))
    ok = True
    for name in ("rand", "ties", "dups"):
        qd, t = g[f"{name}_q"], g[f"{name}_t"]
        a, b = sharding.shard_ranges(t.shape[0], world)[rank]
        idx, dd = mo.knn2(qd, t[a:b], index_base=a)           # local top-2 with global indices
        local = np.empty(qd.shape[0], dtype=dunk.TOP2_DTYPE)
        local["d1"] = np.where(idx[:, 0] < 0, 0xFFFFFFFF, dd[:, 0]).astype(np.uint32)
        local["i1"] = np.where(idx[:, 0] < 0, 0xFFFFFFFF, idx[:, 0]).astype(np.uint32)
        local["d2"] = np.where(idx[:, 1] < 0, 0xFFFFFFFF, dd[:, 1]).astype(np.uint32)
        local["i2"] = np.where(idx[:, 1] < 0, 0xFFFFFFFF, idx[:, 1]).astype(np.uint32)
        parts = sharding.all_gather_top2(local)
        m = sharding.merge_top2_host(parts)
        ok &= np.array_equal(m["i1"].astype(np.int64), g[f"{name}_idx"][:, 0])
        ok &= np.array_equal(m["i2"].astype(np.int64), g[f"{name}_idx"][:, 1])
        ok &= np.array_equal(m["d1"].astype(np.int64), g[f"{name}_dist"][:, 0])
        ok &= np.array_equal(m["d2"].astype(np.int64), g[f"{name}_dist"][:, 1])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_merge_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_ranges_cover_rows():
    import cubesat_apds_b200 as dunk
    from cubesat_apds_b200 import sharding
    for n, w in ((50_000_000, 8), (7, 3), (5, 8), (0, 2)):
        r = sharding.shard_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1
