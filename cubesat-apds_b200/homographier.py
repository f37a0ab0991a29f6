"""Host-side mirror of the reference crate `homographier` (homographier/src/homographier/mod.rs).

Same names, argument order and error behaviour as the Rust items on the hot path:
`HomographyMethod`, `MatError`, `Cmat`, `raster_to_mat`, `find_homography_mat`.  `Cmat<T>` wraps a
numpy array (the checked-Mat idea: never empty, element type checked)."""
from __future__ import annotations

import ctypes as C
import enum
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import DunkError, check, default_context, ptr


class HomographyMethod(enum.IntEnum):
    """mod.rs:25-31"""
    Default = 0
    LMEDS = 4
    RANSAC = 8
    RHO = 16


class MatError(Exception):
    """mod.rs:33-44 — kinds: 'Opencv' (carries a DunkError), 'Empty', 'Jagged', 'Unknown'."""

    def __init__(self, kind: str, inner: Optional[DunkError] = None):
        super().__init__(kind if inner is None else f"{kind}: {inner}")
        self.kind = kind
        self.inner = inner

    @property
    def code(self):
        return None if self.inner is None else self.inner.code


class Cmat:
    """mod.rs:71-146 — checked matrix: guaranteed non-empty, element type fixed at construction."""

    def __init__(self, mat: np.ndarray, dtype=None):
        mat = np.asarray(mat)
        if mat.size == 0 or mat.ndim < 2:
            raise MatError("Empty")          # check_owned: dims() == 0 (mod.rs:86-92)
        if dtype is not None and mat.dtype != np.dtype(dtype):
            raise MatError("Empty")          # Cmat::new type mismatch (mod.rs:116-121)
        self.mat = mat

    @classmethod
    def from_2d_slice(cls, rows: Sequence[Sequence], dtype=None) -> "Cmat":
        """mod.rs:95-101 — jagged input is an OpenCV error in Mat::from_slice_2d."""
        lens = {len(r) for r in rows}
        if len(lens) > 1:
            raise MatError("Opencv", DunkError(_lib.ERR_BAD_ARG, "jagged 2-d slice"))
        if not rows or lens == {0}:
            raise MatError("Empty")
        return cls(np.array(rows, dtype=dtype), dtype)

    @classmethod
    def zeros(cls, rows: int, cols: int, dtype=np.float64) -> "Cmat":
        """mod.rs:139-145"""
        return cls(np.zeros((rows, cols), dtype=dtype), dtype)

    def at_2d(self, row: int, col: int):
        """mod.rs:129-137 — note the reference's transposed bound check (`row > width || col >
        height`, strict '>'), kept as-is; then Mat::at_2d's own range check."""
        h, w = self.mat.shape[:2]
        if row > w or col > h:
            raise MatError("Opencv", DunkError(_lib.ERR_OUT_OF_RANGE, ""))
        if not (0 <= row < h and 0 <= col < w):
            raise MatError("Opencv", DunkError(_lib.ERR_OUT_OF_RANGE, "Mat::at_2d index out of range"))
        return self.mat[row, col]


def raster_to_mat(pixels: np.ndarray, w: int, h: int) -> Cmat:
    """mod.rs:183-197 — `&[RGBA8]` (len w*h, row-major) -> BGRA `Cmat<Vec4b>` of shape (h, w, 4).
    len != w*h -> MatError::Unknown."""
    px = np.asarray(pixels, dtype=np.uint8).reshape(-1, 4)
    if px.shape[0] != w * h:
        raise MatError("Unknown")
    if w * h == 0:
        raise MatError("Empty")
    return Cmat(np.ascontiguousarray(px[:, [2, 1, 0, 3]].reshape(h, w, 4)), np.uint8)


def find_homography_mat(input: np.ndarray, reference: np.ndarray,
                        method: Optional[HomographyMethod] = None,
                        reproj_threshold: Optional[float] = None,
                        ctx: Optional[_lib.Context] = None) -> Tuple[Cmat, Optional[Cmat]]:
    """mod.rs:231-259 — findHomography(input, reference, mask, method or Default, thr or 3.0).
    Returns (Cmat<f64> 3x3, Some(Cmat<u8> N x 1 mask)) — the mask only for RANSAC / LMEDS.
    Errors: < 4 pairs -> MatError('Opencv', code -28); no model -> MatError('Empty')."""
    ctx = ctx or default_context()
    src = np.ascontiguousarray(input, dtype=np.float32).reshape(-1, 2)
    dst = np.ascontiguousarray(reference, dtype=np.float32).reshape(-1, 2)
    if src.shape[0] != dst.shape[0]:
        raise MatError("Opencv", DunkError(_lib.ERR_ASSERT, "src and dst point counts differ"))
    m = int(HomographyMethod.Default if method is None else method)
    thr = 3.0 if reproj_threshold is None else float(reproj_threshold)
    H = np.zeros(9, dtype=np.float64)
    mask = np.zeros(max(src.shape[0], 1), dtype=np.uint8)
    found = C.c_int(0)
    try:
        check(_lib.load().dunk_find_homography(ctx.handle, ptr(src), ptr(dst), src.shape[0], m, thr,
                                               ptr(H), ptr(mask), C.byref(found)))
    except DunkError as e:
        raise MatError("Opencv", e) from None
    if not found.value:
        raise MatError("Empty")
    out_mask = None
    if method in (HomographyMethod.RANSAC, HomographyMethod.LMEDS):
        out_mask = Cmat(mask[: src.shape[0]].reshape(-1, 1), np.uint8)
    return Cmat(H.reshape(3, 3), np.float64), out_mask


def find_homography_batch(src_list, dst_list, reproj_threshold: float = 3.0,
                          method: HomographyMethod = HomographyMethod.RANSAC,
                          ctx: Optional[_lib.Context] = None):
    """Frame-batched form of find_homography_mat (one CTA per frame; SURVEY 8e: partitioned by
    frame, no collective).  Returns (H [B,3,3], masks list, info [B,4] = found/inliers/iters/hyps)."""
    ctx = ctx or default_context()
    B = len(src_list)
    lens = [len(s) for s in src_list]
    offsets = np.zeros(B + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(lens)
    src = np.ascontiguousarray(np.concatenate([np.asarray(s, np.float32).reshape(-1, 2) for s in src_list]))
    dst = np.ascontiguousarray(np.concatenate([np.asarray(d, np.float32).reshape(-1, 2) for d in dst_list]))
    H = np.zeros((B, 9), dtype=np.float64)
    mask = np.zeros(max(int(offsets[-1]), 1), dtype=np.uint8)
    info = np.zeros((B, 4), dtype=np.int32)
    check(_lib.load().dunk_find_homography_batch(ctx.handle, ptr(src), ptr(dst), ptr(offsets), B, int(method),
                                                 float(reproj_threshold), ptr(H), ptr(mask), ptr(info)))
    masks = [mask[offsets[i]:offsets[i + 1]].copy() for i in range(B)]
    return H.reshape(B, 3, 3), masks, info


def score_hypotheses(src, dst, samples, reproj_threshold: float = 3.0, ctx: Optional[_lib.Context] = None):
    """Per-hypothesis inlier counts / minimal-sample H for explicit 4-index samples (parity hook)."""
    ctx = ctx or default_context()
    src = np.ascontiguousarray(src, dtype=np.float32).reshape(-1, 2)
    dst = np.ascontiguousarray(dst, dtype=np.float32).reshape(-1, 2)
    smp = np.ascontiguousarray(samples, dtype=np.int32).reshape(-1, 4)
    counts = np.zeros(max(smp.shape[0], 1), dtype=np.int32)
    Hs = np.zeros((max(smp.shape[0], 1), 9), dtype=np.float64)
    check(_lib.load().dunk_ransac_score_hypotheses(ctx.handle, ptr(src), ptr(dst), src.shape[0], ptr(smp),
                                                   smp.shape[0], float(reproj_threshold), ptr(counts), ptr(Hs)))
    return counts[: smp.shape[0]], Hs[: smp.shape[0]]
