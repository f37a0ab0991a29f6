"""Host-side mirror of the reference crate `feature_extraction` (feature_extraction/src/lib.rs).

Same function names, argument order, meaning and error behaviour as the Rust items; the
bodies only marshal numpy buffers into libdunk_b200.so (the Rust shim in INTEGRATION.md does
exactly the same through `extern "C"`).  `Mat` -> numpy array, `Vector<KeyPoint>` ->
structured array `KEYPOINT_DTYPE` (cv::KeyPoint layout), `Vector<DMatch>` -> `DMATCH_DTYPE`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import DMATCH_DTYPE, KEYPOINT_DTYPE, DunkError, check, default_context, ptr

# lib.rs:12-13
MAX_POINTS_SHIFT = 18
MAX_POINTS = (1 << MAX_POINTS_SHIFT) - 1


@dataclass
class DbKeypoints:
    """lib.rs:20-31"""
    x_coord: float
    y_coord: float
    size: float
    angle: float
    response: float
    octave: int
    class_id: int
    descriptor: bytes
    image_id: int


class ExtractedKeyPoint:
    """lib.rs:15-18 — keypoints + descriptors of one image."""

    def __init__(self, keypoints: np.ndarray, descriptors: np.ndarray):
        self.keypoints = keypoints      # (N,) KEYPOINT_DTYPE
        self.descriptors = descriptors  # (N, 61) u8

    def to_db_type(self, image_id: int) -> List[DbKeypoints]:
        """lib.rs:33-59"""
        out = []
        for i, kp in enumerate(self.keypoints):
            out.append(DbKeypoints(float(kp["x"]), float(kp["y"]), float(kp["size"]),
                                   float(kp["angle"]), float(kp["response"]), int(kp["octave"]),
                                   int(kp["class_id"]), self.descriptors[i].tobytes(), image_id))
        return out


def akaze_keypoint_descriptor_extraction_def(img: np.ndarray, max_points: Optional[int] = None,
                                             ctx: Optional[_lib.Context] = None) -> ExtractedKeyPoint:
    """lib.rs:61-92 — AKAZE(MLDB, size 0, 3 channels, thr 1e-3, 4 octaves x 4 sublevels, PM_G2,
    max_points or 2^18-1).detectAndCompute(img).  img: HxW (gray) or HxWx3 (BGR) / HxWx4 (BGRA) u8."""
    from . import _extract
    return _extract.extract(img, MAX_POINTS if max_points is None else int(max_points), ctx)


def get_knn_matches(origin_desc: np.ndarray, target_desc: np.ndarray, k: int, filter_strength: float,
                    ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """lib.rs:94-114 — Hamming brute-force k-NN (query=origin, train=target) + Lowe ratio filter
    `m0.distance < m1.distance * filter_strength` (f32, strict).  Returns DMATCH_DTYPE in query order.
    Raises DunkError(-211) when a neighbour list has < 2 entries (the reference's `i.get(1)?`)."""
    ctx = ctx or default_context()
    q = _lib.as_desc(origin_desc, "origin_desc")
    t = _lib.as_desc(target_desc, "target_desc")
    if q.shape[0] and t.shape[0] and q.shape[1] != t.shape[1]:
        raise DunkError(_lib.ERR_ASSERT, f"descriptor widths differ: {q.shape[1]} vs {t.shape[1]}")
    width = q.shape[1] if q.shape[0] else (t.shape[1] if t.shape[0] else _lib.DESC_BYTES)
    out = np.empty(max(q.shape[0], 1), dtype=DMATCH_DTYPE)
    n = C.c_int(0)
    check(_lib.load().dunk_knn_match_hamming(ctx.handle, ptr(q), q.shape[0], ptr(t), t.shape[0], width,
                                             int(k), float(filter_strength), ptr(out), out.shape[0],
                                             C.byref(n)))
    return out[: n.value].copy()


def knn2(origin_desc: np.ndarray, target_desc: np.ndarray,
         ctx: Optional[_lib.Context] = None) -> Tuple[np.ndarray, np.ndarray]:
    """The unfiltered 2-NN lists `knn_train_match_def` produces (lib.rs:103): (idx, dist) nq x 2 i32;
    -1 marks a missing neighbour (train set shorter than 2)."""
    ctx = ctx or default_context()
    q = _lib.as_desc(origin_desc, "origin_desc")
    t = _lib.as_desc(target_desc, "target_desc")
    width = q.shape[1] if q.shape[0] else (t.shape[1] if t.shape[0] else _lib.DESC_BYTES)
    idx = np.empty((q.shape[0], 2), dtype=np.int32)
    dist = np.empty((q.shape[0], 2), dtype=np.int32)
    check(_lib.load().dunk_knn2_hamming(ctx.handle, ptr(q), q.shape[0], ptr(t), t.shape[0], width,
                                        ptr(idx), ptr(dist)))
    return idx, dist


def knn2_l2(origin_desc: np.ndarray, target_desc: np.ndarray, ctx: Optional[_lib.Context] = None):
    """Float-descriptor 2-NN (north_star extension; the reference only builds NORM_HAMMING matchers,
    lib.rs:101,121): BFMatcher(NORM_L2).knnMatch(origin, target, 2) for N x 64 or N x 128 f32 descriptors,
    dot products on tcgen05 tensor cores, exact f32 re-rank.  Returns (idx nq x 2 i32, dist nq x 2 f32,
    stats = (queries re-done exactly, slabs))."""
    ctx = ctx or default_context()
    q = np.ascontiguousarray(origin_desc, dtype=np.float32)
    t = np.ascontiguousarray(target_desc, dtype=np.float32)
    if q.ndim != 2 or t.ndim != 2 or (q.shape[0] and t.shape[0] and q.shape[1] != t.shape[1]):
        raise DunkError(_lib.ERR_ASSERT, f"float descriptors must be N x D with equal D: {q.shape} vs {t.shape}")
    dim = q.shape[1] if q.shape[0] else t.shape[1]
    idx = np.empty((q.shape[0], 2), dtype=np.int32)
    dist = np.empty((q.shape[0], 2), dtype=np.float32)
    stats = np.zeros(2, dtype=np.int32)
    check(_lib.load().dunk_knn2_l2(ctx.handle, ptr(q), q.shape[0], ptr(t), t.shape[0], int(dim), ptr(idx), ptr(dist), ptr(stats)))
    return idx, dist, (int(stats[0]), int(stats[1]))


def get_knn_matches_l2(origin_desc: np.ndarray, target_desc: np.ndarray, k: int, filter_strength: float,
                       ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """`get_knn_matches` (lib.rs:94-114) for float descriptors: the same Lowe ratio rule
    `m0.distance < m1.distance * filter_strength` (f32, strict) on the L2 neighbours.  DMATCH_DTYPE in query order."""
    if k < 2:
        raise DunkError(_lib.ERR_OUT_OF_RANGE, "k < 2: the ratio test needs two neighbours (lib.rs:108)")
    idx, dist, _ = knn2_l2(origin_desc, target_desc, ctx)
    keep = dist[:, 0] < dist[:, 1] * np.float32(filter_strength)
    out = np.zeros(int(keep.sum()), dtype=DMATCH_DTYPE)
    out["query_idx"] = np.nonzero(keep)[0]
    out["train_idx"] = idx[keep, 0]
    out["distance"] = dist[keep, 0]
    return out


def get_bruteforce_matches(origin_desc: np.ndarray, target_desc: np.ndarray,
                           ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """lib.rs:116-126 — BFMatcher(NORM_HAMMING, crossCheck=true).match."""
    ctx = ctx or default_context()
    q = _lib.as_desc(origin_desc, "origin_desc")
    t = _lib.as_desc(target_desc, "target_desc")
    width = q.shape[1] if q.shape[0] else (t.shape[1] if t.shape[0] else _lib.DESC_BYTES)
    out = np.empty(max(q.shape[0], 1), dtype=DMATCH_DTYPE)
    n = C.c_int(0)
    check(_lib.load().dunk_match_crosscheck_hamming(ctx.handle, ptr(q), q.shape[0], ptr(t), t.shape[0],
                                                    width, ptr(out), out.shape[0], C.byref(n)))
    return out[: n.value].copy()


def get_points_from_matches(img1_keypoints: np.ndarray, img2_keypoints: np.ndarray,
                            matches: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """lib.rs:161-180 — matched keypoints -> two (M, 2) f32 point arrays.

    Deviation: the reference indexes image 1 by `m.img_idx`
    (always 0, lib.rs:169) and converts image-1 keypoints into BOTH outputs (lib.rs:176-177).
    We implement the evidently intended semantics: `query_idx` -> image 1, `train_idx` -> image 2.
    Out-of-range indices raise DunkError(-211) like `Vector::get`."""
    m = np.asarray(matches, dtype=DMATCH_DTYPE)
    k1 = np.asarray(img1_keypoints, dtype=KEYPOINT_DTYPE)
    k2 = np.asarray(img2_keypoints, dtype=KEYPOINT_DTYPE)
    qi, ti = m["query_idx"], m["train_idx"]
    if m.size and (qi.min() < 0 or qi.max() >= k1.shape[0] or ti.min() < 0 or ti.max() >= k2.shape[0]):
        raise DunkError(_lib.ERR_OUT_OF_RANGE, "match index outside the keypoint vectors")
    p1 = np.stack([k1["x"][qi], k1["y"][qi]], axis=1).astype(np.float32).reshape(-1, 2)
    p2 = np.stack([k2["x"][ti], k2["y"][ti]], axis=1).astype(np.float32).reshape(-1, 2)
    return p1, p2
