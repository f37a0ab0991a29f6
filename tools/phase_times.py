#!/usr/bin/env python
"""Phase times of the two latency-bound tail kernels (block 0, clock64 stamps) from the `make timing` variant:
    make -C cubesat-apds_b200/csrc timing && DUNK_B200_LIB=cubesat-apds_b200/libdunk_b200_timing.so python tools/phase_times.py
Workload = the bench's shape: 64 problems of ~310 correspondences with 1 % outliers."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DUNK_B200_LIB", os.path.join(ROOT, "cubesat-apds_b200", "libdunk_b200_timing.so"))
import cubesat_apds_b200 as dunk  # noqa: E402
import synthdata  # noqa: E402

MHZ = 1965.0


def deltas(stamps, names):
    out = []
    prev = None
    for k, nm in names:
        t = stamps[k]
        if t and prev is not None:
            out.append((nm, (t - prev) / MHZ))
        if t:
            prev = t
    return out


def main():
    lib = C.CDLL(os.environ["DUNK_B200_LIB"])
    ctx = dunk.Context(0, 2)
    rng = np.random.default_rng(0)
    B, n = 64, 310
    hg = dunk.homographier
    # homography problems
    H = synthdata.H_CONFIG1
    srcs, dsts = [], []
    for b in range(B):
        s = rng.uniform(0, 1024, (n, 2)).astype(np.float32)
        p = np.c_[s, np.ones(n)] @ H.T
        d = (p[:, :2] / p[:, 2:] + rng.normal(0, 0.3, (n, 2))).astype(np.float32)
        d[:3] = rng.uniform(0, 1024, (3, 2))
        srcs.append(s); dsts.append(d)
    for _ in range(3):
        hg.find_homography_batch(srcs, dsts, 3.0, ctx=ctx)
    st = (C.c_longlong * 64)()
    lib.dunk_debug_phases_homography(st, 64)
    print("find_homography_kernel (block 0), microseconds:")
    for nm, us in deltas(list(st), [(32, "start"), (33, "RANSAC loop"), (34, "dlt_refit"), (35, "lm_refine"), (36, "mask + finish")]):
        print(f"  {nm:28s} {us:9.1f}")
    # PnP problems: the bench geometry
    S = 10980
    origin = synthdata.scene_origin(S)
    objs, imgs = [], []
    Hs, Rs, ts, _ = synthdata.config5_views(B, S, 77)
    for b in range(B):
        Hi = np.linalg.inv(Hs[b])
        uv = rng.uniform(20, 1000, (n, 2))
        q = np.c_[uv, np.ones(n)] @ Hi.T
        px = q[:, :2] / q[:, 2:]
        X = synthdata.scene_points_ecef(px[:, 0], px[:, 1], S) - origin
        P = X @ Rs[b].T + ts[b]
        im = np.c_[synthdata.CAMERA_F * P[:, 0] / P[:, 2] + 512, synthdata.CAMERA_F * P[:, 1] / P[:, 2] + 512] + rng.normal(0, 0.3, (n, 2))
        im[:3] = rng.uniform(0, 1024, (3, 2))
        objs.append(X); imgs.append(im)
    for _ in range(3):
        rv, tv, masks, info = hg.pnp_solver_ransac_batch(objs, imgs, synthdata.CAMERA_K, 1000, 3.0, 0.99, ctx)
    lib.dunk_debug_phases_pnp(st, 64)
    print("pnp_ransac_kernel (block 0), microseconds:  [info of problem 0: found, inliers, iterations, hypotheses =", info[0].tolist(), "]")
    names = [(0, "start"), (1, "sample stream (thread 0)"), (2, "minimal solves (1 thread each)"), (3, "scoring (warp per hyp.)"),
             (4, "accept rule ... loop end"), (5, "inlier mask"), (6, "final: control points"), (7, "final: alphas + MtM sums"),
             (8, "final: 12x12 Jacobi SVD"), (9, "final: L6x10, rho"), (10, "final: 3 beta approximations"),
             (11, "final: Gauss-Newton + ccs"), (12, "final: sign, R and t (Procrustes)"), (20, "final: reprojection errors"),
             (21, "Rodrigues + finish")]
    for nm, us in deltas(list(st), names):
        print(f"  {nm:36s} {us:9.1f}")
    ctx.close()


if __name__ == "__main__":
    main()
