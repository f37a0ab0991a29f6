"""Matcher-only micro-benchmark (config 3 shape): nq query descriptors vs an HBM-resident random DB.
Usage: python tools/bench_match.py [nq] [db_rows]   (DUNK_B200_LIB selects a kernel-variant build)"""
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
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
lib = load()
ctx = dunk.Context(0, 4)
slot = ctx.reserve_slot()
db = dunk.feature_database.DescriptorDatabase(ctx, capacity=nt)
db.append_random(nt, 7)
rng = np.random.default_rng(0)
q = rng.integers(0, 256, (nq, 64), dtype=np.uint8)
q[:, 60] &= 0x3F
q[:, 61:] = 0
qd = torch.from_numpy(q).cuda()
top2 = torch.empty(nq * 16, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    check(lib.dunk_db_knn2_dev(db.handle, slot, qd.data_ptr(), nq, 0, top2.data_ptr()))
ctx.sync(slot)
reps = 5
ctx.timer_begin(slot)
for _ in range(reps):
    check(lib.dunk_db_knn2_dev(db.handle, slot, qd.data_ptr(), nq, 0, top2.data_ptr()))
ms = ctx.timer_end(slot) / reps
popc = ctx.microbench_popc()
print(json.dumps({"lib": os.path.basename(dunk._lib.LIB_PATH), "nq": nq, "nt": nt, "ms": ms,
                  "gpairs_per_s": nq * nt / ms / 1e6, "popc_tpopc_s": popc,
                  "frac_of_16popc_roofline": nq * nt / ms / 1e6 / (popc * 1e3 / 16)}))
os._exit(0)
