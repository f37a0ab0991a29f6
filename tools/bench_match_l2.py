"""Float-descriptor matcher micro-benchmark: nq x nt f32 descriptors of `dim` floats, device-resident.
Usage: python tools/bench_match_l2.py [nq=3163] [nt=2000000] [dim=64]"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cubesat_apds_b200 as dunk
from cubesat_apds_b200._lib import check, load

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 3163
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 64
lib = load()
ctx = dunk.Context(0, 4)
slot = ctx.reserve_slot()
g = torch.Generator(device="cuda").manual_seed(0)
t = torch.randn(nt, dim, device="cuda", generator=g)
t = t / t.norm(dim=1, keepdim=True)
sel = torch.randint(0, nt, (nq,), device="cuda", generator=g)
q = t[sel] + 0.08 * torch.randn(nq, dim, device="cuda", generator=g)
q = (q / q.norm(dim=1, keepdim=True)).contiguous()
idx = torch.empty(nq, 2, dtype=torch.int32, device="cuda")
dist = torch.empty(nq, 2, dtype=torch.float32, device="cuda")
stats = (C.c_int * 2)()
torch.cuda.synchronize()
for _ in range(2):
    check(lib.dunk_knn2_l2_dev(ctx.handle, slot, q.data_ptr(), nq, t.data_ptr(), nt, dim, idx.data_ptr(), dist.data_ptr(), stats))
ctx.sync(slot)
reps = 5
check(lib.dunk_profile_begin(ctx.handle))
ctx.timer_begin(slot)
for _ in range(reps):
    check(lib.dunk_knn2_l2_dev(ctx.handle, slot, q.data_ptr(), nq, t.data_ptr(), nt, dim, idx.data_ptr(), dist.data_ptr(), stats))
ms = ctx.timer_end(slot) / reps
names = (C.c_char * 4096)()
tm = (C.c_double * 64)()
cnt = (C.c_int * 64)()
alg = (C.c_double * 64)()
k = lib.dunk_profile_end(ctx.handle, names, 4096, tm, cnt, alg, 64)
labels = names.value.decode().split(";")[:k]
stages = {lab: tm[i] / reps for i, lab in enumerate(labels)}
found = float((idx[:, 0].long() == sel).float().mean())
tc_ms = stages.get("match.l2_tcgen05", ms)
print(json.dumps({"nq": nq, "nt": nt, "dim": dim, "ms_total": ms, "stages_ms": stages, "fallback_queries": stats[0], "slabs": stats[1],
                  "planted_top1_rate": found, "gpairs_per_s_total": nq * nt / ms / 1e6,
                  "tcgen05_stage_gpairs_per_s": nq * nt / tc_ms / 1e6,
                  "tcgen05_stage_tflops_tf32": 2.0 * dim * nq * nt / tc_ms / 1e9,
                  "train_bytes_GBps_if_read_once": nt * dim * 4 / tc_ms / 1e6}))
os._exit(0)
