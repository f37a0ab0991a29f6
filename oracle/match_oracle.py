"""Oracle for stage 2 (Hamming brute-force matching) — TEST INFRASTRUCTURE ONLY.

Restates what the reference obtains from OpenCV at feature_extraction/src/lib.rs:94-126:
  * BFMatcher(NORM_HAMMING, crossCheck=false).knnMatch(query, train, k=2)   (lib.rs:101-103)
  * the Lowe ratio filter `m0.distance < m1.distance * r` in f32, strict     (lib.rs:107-111)
  * BFMatcher(NORM_HAMMING, crossCheck=true).match(query, train)            (lib.rs:121-123)
OpenCV (`core/src/batch_distance.cpp`, not vendored in the reference; opencv crate 0.88.8 binds
the system libopencv 4.x) computes the full integer distance row per query and keeps the K
smallest in a stable order, so ties resolve to the LOWER train index; with crossCheck (cv2 4.13.0)
a pair survives iff each row is the other's nearest with lowest-index tie-breaks.  Pinned against cv2 4.13.0 outputs in tests/golden/match_*.npz
(tests/golden/make_golden.py) — parity pinned.
"""
from __future__ import annotations

import numpy as np


def hamming_matrix(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """(nq, nt) int32 Hamming distances between u8 rows."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    out = np.empty((q.shape[0], t.shape[0]), dtype=np.int32)
    step = max(1, (1 << 24) // max(1, t.shape[0] * q.shape[1]))
    for a in range(0, q.shape[0], step):
        x = q[a:a + step, None, :] ^ t[None, :, :]
        out[a:a + step] = np.bitwise_count(x).sum(axis=2, dtype=np.int32)
    return out


def knn2(q: np.ndarray, t: np.ndarray, index_base: int = 0):
    """(idx, dist) nq x 2 int64/int32: the two nearest train rows in stable (distance, index)
    order; -1 where the train set has fewer rows."""
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -1, dtype=np.int64)
    dist = np.full((nq, 2), -1, dtype=np.int32)
    if nq == 0 or nt == 0:
        return idx, dist
    d = hamming_matrix(q, t)
    order = np.argsort(d, axis=1, kind="stable")[:, :2]
    k = order.shape[1]
    idx[:, :k] = order + index_base
    dist[:, :k] = np.take_along_axis(d, order, axis=1)
    return idx, dist


def ratio_filter(idx: np.ndarray, dist: np.ndarray, ratio: float):
    """lib.rs:107-111: keep m0 iff f32(d0) < f32(d1) * f32(ratio).  Returns rows
    (query_idx, train_idx, distance)."""
    d0 = dist[:, 0].astype(np.float32)
    d1 = dist[:, 1].astype(np.float32)
    keep = d0 < d1 * np.float32(ratio)
    qi = np.nonzero(keep)[0]
    return qi.astype(np.int32), idx[qi, 0].astype(np.int64), d0[qi]


def knn_match(q, t, ratio):
    if t.shape[0] < 2 and q.shape[0] > 0:
        raise IndexError("neighbour list shorter than 2 (lib.rs:108 `i.get(1)?`)")
    idx, dist = knn2(q, t)
    return ratio_filter(idx, dist, ratio)


def crosscheck_match(q, t):
    """BFMatcher(crossCheck=true).match: rows (query_idx, train_idx, distance) in query order.
    cv2 4.13.0 keeps (q, i) iff i is q's nearest train row AND q is i's nearest query row, both
    with lowest-index tie-breaks (verified against cv2 incl. tie-heavy inputs)."""
    nq, nt = q.shape[0], t.shape[0]
    if nq == 0 or nt == 0:
        return (np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.float32))
    d = hamming_matrix(q, t)
    sidx = np.argmin(d, axis=1)                   # first minimum = lowest train index
    tidx = np.argmin(d, axis=0)                   # first minimum = lowest query index
    qs = np.nonzero(tidx[sidx] == np.arange(nq))[0]
    return qs.astype(np.int32), sidx[qs].astype(np.int64), d[qs, sidx[qs]].astype(np.float32)


def merge_top2(parts):
    """Lexicographic (distance, index) merge of per-shard top-2 lists.
    parts: list of (idx nq x 2, dist nq x 2) with -1 for empty."""
    idx = np.concatenate([p[0] for p in parts], axis=1).astype(np.int64)
    dist = np.concatenate([p[1] for p in parts], axis=1).astype(np.int64)
    big = np.int64(1) << 40
    key = np.where(idx < 0, big * 1024, dist * big + idx)
    order = np.argsort(key, axis=1, kind="stable")[:, :2]
    return np.take_along_axis(idx, order, axis=1), np.take_along_axis(dist, order, axis=1).astype(np.int32)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def random_db_rows(n: int, seed: int, row_offset: int = 0, desc_bytes: int = 61) -> np.ndarray:
    """The synthetic DB rows `dunk_db_append_random` writes (bench config 3): u64 word j of row r
    = splitmix64(seed + r*8 + j); byte 60 keeps 6 bits, bytes 61..63 zero."""
    r = (np.arange(n, dtype=np.uint64) + np.uint64(row_offset))[:, None]
    j = np.arange(8, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        v = _splitmix64(np.uint64(seed) + r * np.uint64(8) + j)
    v[:, 7] &= np.uint64(0x0000003FFFFFFFFF)
    rows = v.astype("<u8").view(np.uint8).reshape(n, 64)
    return np.ascontiguousarray(rows[:, :desc_bytes])


def knn2_l2(q, t, chunk=256):
    """cv::BFMatcher(NORM_L2).knnMatch(q, t, 2) for f32 descriptors: squared differences summed in f32
    (sequential order), sqrt, two smallest by (distance, index).  Pinned against cv2 4.13.0 in
    tests/golden/l2_golden.npz: indices identical, distances within 1e-6 relative (OpenCV's SIMD sum
    associates differently)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.float32)
    idx = np.empty((q.shape[0], 2), np.int32)
    dist = np.empty((q.shape[0], 2), np.float32)
    for a in range(0, q.shape[0], chunk):
        d = q[a:a + chunk, None, :] - t[None, :, :]
        d2 = np.zeros(d.shape[:2], np.float32)
        for k in range(d.shape[2]):
            d2 = d2 + d[:, :, k] * d[:, :, k]
        order = np.argsort(d2, axis=1, kind="stable")[:, :2]
        idx[a:a + chunk] = order
        dist[a:a + chunk] = np.sqrt(np.take_along_axis(d2, order, axis=1))
    return idx, dist
