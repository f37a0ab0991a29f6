"""Extraction-only micro-benchmark (BASELINE config 2 shape): a batch of 1024 x 1024 u8 frames through
AKAZE detect + MLDB describe on one GPU, frames resident in HBM; prints frames/s and the per-kernel
device times (CUDA events inside the library) with achieved algorithmic GB/s.
Usage: python tools/bench_extract.py [frames=256] [sub_batch=64] [reps=3]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cubesat_apds_b200 as dunk
import synthdata
from cubesat_apds_b200._lib import PipelineView, check, load

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sub = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N = 1024
lib = load()
ctx = dunk.Context(0, 4)
slot = ctx.reserve_slot()
# frames synth(1024, seed = 100 + i); 16 distinct images cycled to bound the host time
base = np.stack([synthdata.synth_image(N, N, 100 + i) for i in range(16)])
imgs = torch.from_numpy(base[np.arange(frames) % 16]).cuda()
ws_bytes = int(lib.dunk_pipeline_workspace_bytes(ctx.handle, sub, N, N))
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
view = PipelineView()
torch.cuda.synchronize()


def run_all():
    total = 0
    for f0 in range(0, frames, sub):
        nf = min(sub, frames - f0)
        check(lib.dunk_pipeline_extract_dev(ctx.handle, slot, imgs.data_ptr() + f0 * N * N, nf, N, N, 1, N, N * N, 0,
                                            ws.data_ptr(), ws_bytes, C.byref(view)))
        total += view.total_queries
    return total


for _ in range(2):
    kps = run_all()
ctx.sync(slot)
ctx.timer_begin(slot)
for _ in range(reps):
    run_all()
ms = ctx.timer_end(slot) / reps
check(lib.dunk_profile_begin(ctx.handle))
for _ in range(reps):
    run_all()
ctx.sync(slot)
names = (C.c_char * 4096)()
t = (C.c_double * 64)()
cnt = (C.c_int * 64)()
alg = (C.c_double * 64)()
k = lib.dunk_profile_end(ctx.handle, names, 4096, t, cnt, alg, 64)
labels = names.value.decode().split(";")[:k]
stages = {lab: {"ms": t[i] / reps, "launches": cnt[i] // reps, "alg_GBps": (alg[i] / t[i] / 1e6 if t[i] > 0 else 0)}
          for i, lab in enumerate(labels)}
print(json.dumps({"frames": frames, "sub_batch": sub, "ms": ms, "frames_per_s": frames / ms * 1e3,
                  "keypoints_per_frame": kps / frames, "stages": stages}))
for lab, v in sorted(stages.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {lab:24s} {v['ms']:8.3f} ms  {v['launches']:4d} launches  {v['alg_GBps']:8.1f} GB/s (algorithmic)", file=sys.stderr)
os._exit(0)
