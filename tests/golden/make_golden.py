"""Generates the golden vectors under tests/golden/ by running the reference's OpenCV entry
points (through cv2, the same C++ library the `opencv` crate binds) with the reference's exact
arguments.  Run in the build container:  python tests/golden/make_golden.py [match|ransac|akaze|all]
The OpenCV version is recorded in every file (parity is defined against that version)."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def knn_cv(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(q, t, 2)      # lib.rs:101-103
    idx = np.array([[a.trainIdx for a in r] + [-1] * (2 - len(r)) for r in m], dtype=np.int32).reshape(-1, 2)
    dist = np.array([[a.distance for a in r] + [-1] * (2 - len(r)) for r in m], dtype=np.int32).reshape(-1, 2)
    return idx, dist


def cross_cv(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(q, t)             # lib.rs:121-123
    return np.array([(a.queryIdx, a.trainIdx, a.distance) for a in m], dtype=np.int32).reshape(-1, 3)


def make_match():
    rng = np.random.default_rng(20240321)
    cases = {}
    # uniform random 61-byte rows (last byte 6 bits) — ragged sizes, not multiples of any tile
    q = rng.integers(0, 256, (333, 61), dtype=np.uint8); q[:, 60] &= 0x3F
    t = rng.integers(0, 256, (1777, 61), dtype=np.uint8); t[:, 60] &= 0x3F
    cases["rand"] = (q, t)
    # tie-heavy: few distinct bits -> many equal distances (tie-break = lowest train index)
    q = np.zeros((257, 61), np.uint8); t = np.zeros((901, 61), np.uint8)
    q[:, :2] = rng.integers(0, 256, (257, 2)); t[:, :2] = rng.integers(0, 256, (901, 2))
    cases["ties"] = (q, t)
    # duplicates planted: exact zero-distance matches and duplicate train rows
    t = rng.integers(0, 256, (640, 61), dtype=np.uint8); t[:, 60] &= 0x3F
    q = t[rng.integers(0, 640, 200)].copy()
    t[100:110] = t[5]                                                  # duplicate rows
    cases["dups"] = (q, t)
    # minimum train size for the ratio test, and a single query
    cases["tiny"] = (rng.integers(0, 256, (1, 61), dtype=np.uint8), rng.integers(0, 256, (2, 61), dtype=np.uint8))
    out = {"opencv_version": np.array(cv2.__version__)}
    for name, (q, t) in cases.items():
        idx, dist = knn_cv(q, t)
        out[f"{name}_q"], out[f"{name}_t"] = q, t
        out[f"{name}_idx"], out[f"{name}_dist"] = idx, dist
        out[f"{name}_cross"] = cross_cv(q, t)
    np.savez_compressed(os.path.join(HERE, "match_golden.npz"), **out)
    print("match_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("match", "all"):
        make_match()
    if what in ("ransac", "all") and "make_ransac" in globals():
        make_ransac()
    if what in ("akaze", "all") and "make_akaze" in globals():
        make_akaze()
