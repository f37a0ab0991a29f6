"""Micro-benchmark of the two latency-bound tail kernels at cv2's probe shapes:
  * findHomography(RANSAC, 3.0): N = 1000 pairs, 40 % outliers;
  * solvePnPRansac(EPNP, 1000 iterations, thr 2.0, conf 0.99): N = 1000 points, 40 % outliers.
Batches of B independent problems go through the host-buffer batch calls (dunk_find_homography_batch /
dunk_pnp_ransac_batch: H2D + kernel + D2H inside the time); cv2 runs the same problems one after another on the
box's host cores.  Prints one JSON line per (kernel, B): problems/s, hypotheses/s, point evaluations/s and the
share of the FP32 FMA peak those evaluations amount to (the kernels are latency-bound: serial f64 solves).

    python tools/bench_tail.py > profiles/r2_tail_microbench.jsonl
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cubesat_apds_b200 as dunk  # noqa: E402

H_TRUE = np.array([[0.98, -0.12, 60], [0.10, 1.03, -40], [1e-5, -2e-5, 1]])
K = np.array([[800., 0, 512], [0, 820., 500], [0, 0, 1]])
FP32_PEAK = 148 * 128 * 2 * 1.965e9          # FMA lanes x 2 flop x SM clock (B200)


def h_case(n, out_frac, sigma, seed):
    r = np.random.default_rng(seed)
    src = r.uniform(0, 1024, (n, 2)).astype(np.float32)
    p = np.c_[src, np.ones(n)] @ H_TRUE.T
    dst = (p[:, :2] / p[:, 2:]) + r.normal(0, sigma, (n, 2))
    k = r.permutation(n)[:int(n * out_frac)]
    dst[k] = r.uniform(0, 1024, (len(k), 2))
    return src, dst.astype(np.float32)


def pnp_case(n, out_frac, noise, seed):
    r = np.random.default_rng(seed)
    obj = r.uniform(-2, 2, (n, 3))
    rv, tv = r.normal(0, 0.3, 3), np.array([0.2, -0.1, 8.0]) + r.normal(0, 0.3, 3)
    import cv2
    R, _ = cv2.Rodrigues(rv)
    P = obj @ R.T + tv
    img = np.stack([K[0, 0] * P[:, 0] / P[:, 2] + K[0, 2], K[1, 1] * P[:, 1] / P[:, 2] + K[1, 2]], 1) + r.normal(0, noise, (n, 2))
    k = r.permutation(n)[:int(n * out_frac)]
    img[k] = r.uniform(0, 1024, (len(k), 2))
    return obj, img


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def main():
    import cv2
    ctx = dunk.Context(0, 2)
    hg = dunk.homographier
    n = 1000
    for B in (1, 64, 148, 296, 1024):
        cases = [h_case(n, 0.4, 0.5, s) for s in range(B)]
        dt, (H, masks, info) = timed(lambda: hg.find_homography_batch([c[0] for c in cases], [c[1] for c in cases], 3.0, ctx=ctx), 5)
        hyp = int(info[:, 3].sum())
        rec = {"kernel": "find_homography_kernel (RANSAC, 3.0, 2000 it, 0.995)", "problems": B, "pairs_per_problem": n, "outliers": 0.4,
               "ms_per_batch": dt * 1e3, "problems_per_s": B / dt, "hypotheses_per_s": hyp / dt, "point_evals_per_s": hyp * n / dt,
               "frac_fp32_peak": hyp * n * 30 / dt / FP32_PEAK, "flop_per_eval": 30, "found": int(info[:, 0].sum()),
               "hypotheses_per_problem_mean": hyp / B, "timing": "host-buffer batch call (H2D + kernel + D2H)"}
        if B <= 64:
            t0 = time.perf_counter()
            for s, d in cases:
                cv2.findHomography(s, d, cv2.RANSAC, 3.0)
            rec["cv2_ms_per_problem"] = (time.perf_counter() - t0) / B * 1e3
            rec["cv2_threads"] = cv2.getNumThreads()
        print(json.dumps(rec), flush=True)
    for B in (1, 64, 148, 296, 1024):
        cases = [pnp_case(n, 0.4, 0.5, 100 + s) for s in range(B)]
        dt, (rv, tv, masks, info) = timed(lambda: hg.pnp_solver_ransac_batch([c[0] for c in cases], [c[1] for c in cases], K, 1000, 2.0,
                                                                              0.99, ctx), 5)
        hyp = int(info[:, 3].sum())
        rec = {"kernel": "pnp_ransac_kernel (EPNP, 1000 it, thr 2.0, 0.99)", "problems": B, "points_per_problem": n, "outliers": 0.4,
               "ms_per_batch": dt * 1e3, "problems_per_s": B / dt, "hypotheses_per_s": hyp / dt, "point_evals_per_s": hyp * n / dt,
               "frac_fp32_peak": hyp * n * 40 / dt / FP32_PEAK, "flop_per_eval": 40, "found": int(info[:, 0].sum()),
               "hypotheses_per_problem_mean": hyp / B, "timing": "host-buffer batch call (H2D + kernel + D2H)"}
        if B <= 64:
            t0 = time.perf_counter()
            for o, i in cases:
                cv2.solvePnPRansac(o, i, K, np.zeros((4, 1)), None, None, False, 1000, 2.0, 0.99, None, cv2.SOLVEPNP_EPNP)
            rec["cv2_ms_per_problem"] = (time.perf_counter() - t0) / B * 1e3
            rec["cv2_threads"] = cv2.getNumThreads()
        print(json.dumps(rec), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
