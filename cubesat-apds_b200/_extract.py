"""Marshalling for stage 1 (AKAZE extraction) — feature_extraction/src/lib.rs:61-92."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import KEYPOINT_DTYPE, DunkError, check, default_context, ptr


def _as_image(img: np.ndarray) -> np.ndarray:
    a = np.asarray(img)
    if a.dtype != np.uint8:
        raise DunkError(_lib.ERR_ASSERT, f"image depth {a.dtype} unsupported (CV_8U expected)")
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3 or a.shape[2] not in (1, 3, 4) or a.size == 0:
        raise DunkError(_lib.ERR_ASSERT, f"image shape {a.shape} unsupported (HxW, HxWx3 BGR, HxWx4 BGRA)")
    return np.ascontiguousarray(a)


def default_capacity(rows: int, cols: int) -> int:
    return int(min(max(rows * cols // 32, 2048), 1 << 20))


def extract(img: np.ndarray, max_points: int, ctx: Optional[_lib.Context] = None):
    from .feature_extraction import ExtractedKeyPoint
    ctx = ctx or default_context()
    a = _as_image(img)
    rows, cols, ch = a.shape
    cap = default_capacity(rows, cols)
    kps = np.empty(cap, dtype=KEYPOINT_DTYPE)
    desc = np.empty((cap, 61), dtype=np.uint8)
    n = C.c_int(0)
    check(_lib.load().dunk_akaze_extract(ctx.handle, ptr(a), rows, cols, ch, cols * ch, int(max_points), ptr(kps),
                                         ptr(desc), cap, C.byref(n)))
    return ExtractedKeyPoint(kps[: n.value].copy(), desc[: n.value].copy())


def extract_batch(frames: np.ndarray, max_points: int = _lib.MAX_POINTS, ctx: Optional[_lib.Context] = None) -> List:
    """Batch of same-shape frames [B, H, W] or [B, H, W, C] -> list of ExtractedKeyPoint (frame-batched
    kernels; SURVEY 8e: extraction partitions by frame with no collective)."""
    from .feature_extraction import ExtractedKeyPoint
    ctx = ctx or default_context()
    a = np.asarray(frames)
    if a.ndim == 3:
        a = a[..., None]
    if a.dtype != np.uint8 or a.ndim != 4 or a.shape[3] not in (1, 3, 4):
        raise DunkError(_lib.ERR_ASSERT, f"frame batch shape {a.shape} / dtype {a.dtype} unsupported")
    a = np.ascontiguousarray(a)
    B, rows, cols, ch = a.shape
    cap = default_capacity(rows, cols)
    kps = np.empty((B, cap), dtype=KEYPOINT_DTYPE)
    desc = np.empty((B, cap, 61), dtype=np.uint8)
    counts = np.zeros(B, dtype=np.int32)
    check(_lib.load().dunk_akaze_extract_batch(ctx.handle, ptr(a), B, rows, cols, ch, cols * ch, rows * cols * ch,
                                               int(max_points), ptr(kps), ptr(desc), cap, ptr(counts)))
    return [ExtractedKeyPoint(kps[i, : counts[i]].copy(), desc[i, : counts[i]].copy()) for i in range(B)]


def debug_level(img: np.ndarray, level: int, ctx: Optional[_lib.Context] = None):
    """Per-stage parity hook: (Lt, Lx, Ly, Ldet, kcontrast) of one evolution level."""
    ctx = ctx or default_context()
    a = _as_image(img)
    rows, cols, ch = a.shape
    bufs = [np.zeros((rows, cols), np.float32) for _ in range(4)]
    k = np.zeros(1, np.float32)
    w, h, nl = C.c_int(0), C.c_int(0), C.c_int(0)
    check(_lib.load().dunk_akaze_debug_level(ctx.handle, ptr(a), rows, cols, ch, cols * ch, level, ptr(bufs[0]),
                                             ptr(bufs[1]), ptr(bufs[2]), ptr(bufs[3]), ptr(k), C.byref(w), C.byref(h),
                                             C.byref(nl)))
    out = [b.ravel()[: w.value * h.value].reshape(h.value, w.value).copy() for b in bufs]
    return out[0], out[1], out[2], out[3], float(k[0]), nl.value
