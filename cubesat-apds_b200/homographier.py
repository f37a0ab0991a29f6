"""Host-side mirror of the reference crate `homographier` (homographier/src/homographier/mod.rs).

Same names, argument order and error behaviour as the Rust items on the hot path:
`HomographyMethod`, `MatError`, `Cmat`, `raster_to_mat`, `find_homography_mat`, `SolvePnPMethod`,
`ImgObjCorrespondence`, `PNPRANSACSolution`, `pnp_solver_ransac`, `warp_image_perspective`.  `Cmat<T>` wraps a
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
    Returns (Cmat<f64> 3x3, Some(Cmat<u8> N x 1 mask)) — the mask only for RANSAC / LMEDS, as the reference's
    `match method` (mod.rs:252-256); RHO is accepted and served by the RANSAC estimator (tolerance parity with cv2's
    PROSAC-based RHO), its mask is available through find_homography_batch.
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


def warp_image_perspective(src: Cmat, m, size: Optional[Tuple[int, int]] = None, ctx: Optional[_lib.Context] = None) -> Cmat:
    """mod.rs:271-300 — warpPerspective(src, M, size or src.size(), INTER_LINEAR, BORDER_CONSTANT,
    Scalar(1,1,1,1)).  `size` is (width, height) like Size2i.  8-bit images with 1..4 channels.
    m must be 3x3 (anything else is an OpenCV assertion, -215)."""
    ctx = ctx or default_context()
    a = src.mat if isinstance(src, Cmat) else np.asarray(src)
    M = m.mat if isinstance(m, Cmat) else np.asarray(m)
    if M.shape != (3, 3):
        raise MatError("Opencv", DunkError(_lib.ERR_ASSERT, "warpPerspective: M must be 3x3"))
    if a.dtype != np.uint8 or a.ndim not in (2, 3):
        raise MatError("Opencv", DunkError(_lib.ERR_ASSERT, f"warp_image_perspective: {a.dtype} / {a.ndim}-d images unsupported"))
    a = np.ascontiguousarray(a)
    rows, cols = a.shape[:2]
    ch = 1 if a.ndim == 2 else a.shape[2]
    w, h = (cols, rows) if size is None else (int(size[0]), int(size[1]))
    out = np.empty((h, w) if a.ndim == 2 else (h, w, ch), dtype=np.uint8)
    Md = np.ascontiguousarray(M, dtype=np.float64)
    try:
        check(_lib.load().dunk_warp_perspective(ctx.handle, ptr(a), rows, cols, ch, cols * ch, ptr(Md), h, w, None, ptr(out)))
    except DunkError as e:
        raise MatError("Opencv", e) from None
    return Cmat(out, np.uint8)


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


# ------------------------------------------------------------------------------------------ PnP
class SolvePnPMethod(enum.IntEnum):
    """opencv::calib3d::SolvePnPMethod values reachable from mod.rs:320-361"""
    SOLVEPNP_ITERATIVE = 0
    SOLVEPNP_EPNP = 1
    SOLVEPNP_P3P = 2


class ImgObjCorrespondence:
    """mod.rs:53-65 — a 3-D object point and its 2-D image point"""
    __slots__ = ("obj_point", "img_point")

    def __init__(self, obj_point, img_point):
        self.obj_point = tuple(float(v) for v in obj_point)
        self.img_point = tuple(float(v) for v in img_point)
        if len(self.obj_point) != 3 or len(self.img_point) != 2:
            raise MatError("Unknown")


class PNPRANSACSolution:
    """mod.rs:46-51 — rvec, tvec: Cmat<f64> 3x1; inliers: Cmat<i32> M x 1"""
    __slots__ = ("rvec", "tvec", "inliers")

    def __init__(self, rvec: Cmat, tvec: Cmat, inliers: Cmat):
        self.rvec, self.tvec, self.inliers = rvec, tvec, inliers


def _split_correspondences(point_correspondences):
    if isinstance(point_correspondences, tuple) and len(point_correspondences) == 2:
        obj, img = point_correspondences                      # (N x 3, N x 2) arrays
    else:
        obj = [c.obj_point for c in point_correspondences]
        img = [c.img_point for c in point_correspondences]
    obj = np.ascontiguousarray(obj, dtype=np.float64).reshape(-1, 3)
    img = np.ascontiguousarray(img, dtype=np.float64).reshape(-1, 2)
    if obj.shape[0] != img.shape[0]:
        raise MatError("Opencv", DunkError(_lib.ERR_ASSERT, "object and image point counts differ"))
    return obj, img


def pnp_solver_ransac(point_correspondences, camera_intrinsic, iter_count: int, reproj_thres: float, confidence: float,
                      dist_coeffs=None, method: Optional[SolvePnPMethod] = None,
                      ctx: Optional[_lib.Context] = None) -> Optional[PNPRANSACSolution]:
    """mod.rs:320-369 — solvePnPRansac(obj, img, K, zeros(4,1), rvec, tvec, false, iter_count,
    reproj_thres, confidence, inliers, method or SOLVEPNP_EPNP).  `dist_coeffs` is accepted and ignored
    exactly as in the reference (it shadows the argument with zeros, :344).  Returns the solution or
    None when no pose was found (`Ok(None)`); fewer than 4 correspondences -> MatError('Opencv', -215)
    (reference test :627-638).  `point_correspondences`: a sequence of ImgObjCorrespondence, or an
    (obj N x 3, img N x 2) pair of arrays."""
    ctx = ctx or default_context()
    obj, img = _split_correspondences(point_correspondences)
    K = camera_intrinsic.mat if isinstance(camera_intrinsic, Cmat) else np.asarray(camera_intrinsic)
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(3, 3)
    m = int(SolvePnPMethod.SOLVEPNP_EPNP if method is None else method)
    n = obj.shape[0]
    rvec, tvec = np.zeros(3), np.zeros(3)
    inliers = np.zeros(max(n, 1), dtype=np.int32)
    n_inl, found = C.c_int(0), C.c_int(0)
    try:
        check(_lib.load().dunk_pnp_ransac(ctx.handle, ptr(obj), ptr(img), n, ptr(K), int(iter_count), float(reproj_thres),
                                          float(confidence), m, ptr(rvec), ptr(tvec), ptr(inliers), n, C.byref(n_inl),
                                          C.byref(found)))
    except DunkError as e:
        raise MatError("Opencv", e) from None
    if not found.value:
        return None
    return PNPRANSACSolution(Cmat(rvec.reshape(3, 1), np.float64), Cmat(tvec.reshape(3, 1), np.float64),
                             Cmat(inliers[: n_inl.value].reshape(-1, 1).copy(), np.int32))


def pnp_solver_ransac_batch(obj_list, img_list, camera_intrinsics, iter_count: int, reproj_thres: float, confidence: float,
                            ctx: Optional[_lib.Context] = None, method: Optional[SolvePnPMethod] = None):
    """Frame-batched pnp_solver_ransac (one CTA per frame; partitioned by frame, no collective).
    camera_intrinsics: one 3x3 for all frames or one per frame.
    Returns (rvecs [B,3], tvecs [B,3], inlier masks list, info [B,4] = found/inliers/iters/hypotheses)."""
    ctx = ctx or default_context()
    B = len(obj_list)
    lens = [len(o) for o in obj_list]
    offsets = np.zeros(B + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(lens)
    obj = np.ascontiguousarray(np.concatenate([np.asarray(o, np.float64).reshape(-1, 3) for o in obj_list]))
    img = np.ascontiguousarray(np.concatenate([np.asarray(i, np.float64).reshape(-1, 2) for i in img_list]))
    K = np.asarray(camera_intrinsics, dtype=np.float64)
    K = np.ascontiguousarray(np.broadcast_to(K.reshape(-1, 3, 3), (B, 3, 3)))
    rvecs, tvecs = np.zeros((B, 3)), np.zeros((B, 3))
    mask = np.zeros(max(int(offsets[-1]), 1), dtype=np.uint8)
    info = np.zeros((B, 4), dtype=np.int32)
    check(_lib.load().dunk_pnp_ransac_batch(ctx.handle, ptr(obj), ptr(img), ptr(offsets), B, ptr(K), int(iter_count),
                                            float(reproj_thres), float(confidence),
                                            int(SolvePnPMethod.SOLVEPNP_EPNP if method is None else method),
                                            ptr(rvecs), ptr(tvecs), ptr(mask), ptr(info)))
    return rvecs, tvecs, [mask[offsets[i]:offsets[i + 1]].astype(bool) for i in range(B)], info


def pnp_score_hypotheses(obj, img, camera_intrinsic, samples, reproj_thres: float, ctx: Optional[_lib.Context] = None):
    """Per-hypothesis inlier counts and (rvec, tvec) for explicit 5-index samples (parity hook)."""
    ctx = ctx or default_context()
    obj = np.ascontiguousarray(obj, dtype=np.float64).reshape(-1, 3)
    img = np.ascontiguousarray(img, dtype=np.float64).reshape(-1, 2)
    K = np.ascontiguousarray(camera_intrinsic, dtype=np.float64).reshape(3, 3)
    smp = np.ascontiguousarray(samples, dtype=np.int32).reshape(-1, 5)
    counts = np.zeros(max(smp.shape[0], 1), dtype=np.int32)
    rt = np.zeros((max(smp.shape[0], 1), 6), dtype=np.float64)
    check(_lib.load().dunk_pnp_score_hypotheses(ctx.handle, ptr(obj), ptr(img), obj.shape[0], ptr(K), ptr(smp), smp.shape[0],
                                                float(reproj_thres), ptr(counts), ptr(rt)))
    return counts[: smp.shape[0]], rt[: smp.shape[0]]
