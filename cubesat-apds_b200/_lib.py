"""ctypes binding of libdunk_b200.so (include/dunk_b200.h).

The library is the product; this module only marshals numpy buffers across the C ABI.
There is no CPU fallback: if the shared object is missing or no B200 is present the
calls raise (`DunkError`), they never route anywhere else.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DUNK_B200_LIB") or os.path.join(_HERE, "libdunk_b200.so")   # env: kernel-variant A/B runs

# status codes (include/dunk_b200.h)
OK = 0
ERR_NO_MEM = -4
ERR_BAD_ARG = -5
ERR_VEC_LENGTH = -28
ERR_OUT_OF_RANGE = -211
ERR_ASSERT = -215
ERR_CUDA = -217

MAX_POINTS_SHIFT = 18
MAX_POINTS = (1 << MAX_POINTS_SHIFT) - 1
DESC_BYTES = 61

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
DMATCH_DTYPE = np.dtype(
    [("query_idx", "<i4"), ("train_idx", "<i4"), ("img_idx", "<i4"), ("distance", "<f4")]
)
REGISTRATION_DTYPE = np.dtype(
    [("H", "<f8", (9,)), ("found", "<i4"), ("inliers", "<i4"), ("matches", "<i4"), ("keypoints", "<i4"),
     ("ransac_iters", "<i4"), ("hypotheses", "<i4")]
)
IMAGE_DTYPE = np.dtype([("id", "<i4"), ("x_start", "<i4"), ("y_start", "<i4"), ("x_end", "<i4"), ("y_end", "<i4"),
                        ("level_of_detail", "<i4")])


class RowFilter(C.Structure):
    """DunkRowFilter (include/dunk_b200.h)"""
    _fields_ = [("image_id", C.c_int32), ("level_of_detail", C.c_int32), ("use_box", C.c_int32), ("x_start", C.c_float),
                ("y_start", C.c_float), ("x_end", C.c_float), ("y_end", C.c_float)]


TOP2_DTYPE = np.dtype([("d1", "<u4"), ("i1", "<u4"), ("d2", "<u4"), ("i2", "<u4")])
POSE_DTYPE = np.dtype([("rvec", "<f8", (3,)), ("tvec", "<f8", (3,)), ("found", "<i4"), ("inliers", "<i4"),
                       ("ransac_iters", "<i4"), ("hypotheses", "<i4")])
assert POSE_DTYPE.itemsize == 64
SHARD_ID_BYTES = 128


class PoseConfig(C.Structure):
    """DunkPoseConfig (include/dunk_b200.h): the pose stage behind the homography — get_world_coordinates
    (feature_database/src/elevationdb.rs:64-104) + pnp_solver_ransac (homographier/src/homographier/mod.rs:320-369)"""
    _fields_ = [("elevation", C.c_void_p), ("K", C.c_double * 9), ("origin", C.c_double * 3), ("method", C.c_int32),
                ("iters", C.c_int32), ("thr", C.c_float), ("confidence", C.c_double)]


assert REGISTRATION_DTYPE.itemsize == 96 and KEYPOINT_DTYPE.itemsize == 28 and DMATCH_DTYPE.itemsize == 16 and TOP2_DTYPE.itemsize == 16


class PipelineView(C.Structure):
    """DunkPipelineView (include/dunk_b200.h)"""
    _fields_ = [("query64_dev", C.c_void_p), ("query_offsets_dev", C.c_void_p), ("keypoints_dev", C.c_void_p),
                ("keypoint_counts_dev", C.c_void_p), ("top2_dev", C.c_void_p), ("total_queries", C.c_int32),
                ("keypoint_capacity", C.c_int32), ("query_capacity", C.c_int64)]


class DunkError(RuntimeError):
    """Mirror of `opencv::Error { code, message }` (feature_extraction/src/lib.rs:61)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"dunk_b200 error {code}: {message}")
        self.code = code
        self.message = message


_lib = None
_lib_lock = threading.Lock()

_vp, _i, _i64, _u32, _u64, _f, _d = (C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64,
                                      C.c_float, C.c_double)
_pi = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol include/dunk_b200.h declares
SIGNATURES = {
    "dunk_ctx_create": (_i, [_i, _i, C.POINTER(_vp)]),
    "dunk_ctx_destroy": (None, [_vp]),
    "dunk_last_error": (C.c_char_p, []),
    "dunk_version": (C.c_char_p, []),
    "dunk_ctx_stream": (_vp, [_vp, _i]),
    "dunk_ctx_device": (_i, [_vp]),
    "dunk_ctx_sm_count": (_i, [_vp]),
    "dunk_ctx_launch_count": (_u64, [_vp]),
    "dunk_timer_begin": (_i, [_vp, _i]),
    "dunk_timer_end": (_i, [_vp, _i, C.POINTER(_f)]),
    "dunk_sync": (_i, [_vp, _i]),
    "dunk_ctx_reserve_slot": (_i, [_vp]),
    "dunk_ctx_release_slot": (_i, [_vp, _i]),
    "dunk_knn_match_hamming": (_i, [_vp, _vp, _i, _vp, _i64, _i, _i, _f, _vp, _i, _pi]),
    "dunk_knn2_hamming": (_i, [_vp, _vp, _i, _vp, _i64, _i, _vp, _vp]),
    "dunk_match_crosscheck_hamming": (_i, [_vp, _vp, _i, _vp, _i64, _i, _vp, _i, _pi]),
    "dunk_knn2_l2": (_i, [_vp, _vp, _i, _vp, _i64, _i, _vp, _vp, _vp]),
    "dunk_knn2_l2_dev": (_i, [_vp, _i, _vp, _i, _vp, _i64, _i, _vp, _vp, _vp]),
    "dunk_db_create": (_i, [_vp, _i64, _i, C.POINTER(_vp)]),
    "dunk_db_destroy": (None, [_vp]),
    "dunk_db_append": (_i, [_vp, _vp, _vp, _vp, _i64]),
    "dunk_db_append_random": (_i, [_vp, _i64, _u64]),
    "dunk_db_append_random_at": (_i, [_vp, _i64, _u64, _u64]),
    "dunk_db_size": (_i64, [_vp]),
    "dunk_db_clear": (_i, [_vp]),
    "dunk_selftest_gamma_lut": (_i, [_vp, _vp]),
    "dunk_db_read": (_i, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "dunk_db_create_image": (_i, [_vp, _i, _i, _i, _i, _i, _pi]),
    "dunk_db_read_image": (_i, [_vp, _i, _vp]),
    "dunk_db_find_images": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _pi]),
    "dunk_db_image_count": (_i, [_vp]),
    "dunk_db_select": (_i, [_vp, _vp, _i64, C.POINTER(_vp)]),
    "dunk_db_read_ids": (_i, [_vp, _i64, _i64, _vp]),
    "dunk_db_save": (_i, [_vp, C.c_char_p]),
    "dunk_db_load": (_i, [_vp, C.c_char_p, _i64, C.POINTER(_vp)]),
    "dunk_db_match": (_i, [_vp, _vp, _i, _f, _vp, _i, _pi]),
    "dunk_db_knn2": (_i, [_vp, _vp, _i, _u32, _vp]),
    "dunk_db_knn2_dev": (_i, [_vp, _i, _vp, _i, _u32, _vp]),
    "dunk_top2_merge_dev": (_i, [_vp, _i, _vp, _i, _i, _vp]),
    "dunk_top2_ratio_dev": (_i, [_vp, _i, _vp, _i, _f, _vp, _vp]),
    "dunk_pad_desc_dev": (_i, [_vp, _i, _vp, _i64, _i, _vp]),
    "dunk_profile_begin": (_i, [_vp]),
    "dunk_profile_end": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i]),
    "dunk_microbench_popc": (_i, [_vp, _i, C.POINTER(_d)]),
    "dunk_akaze_extract": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _pi]),
    "dunk_akaze_extract_batch": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.c_size_t, _i, _vp, _vp, _i, _vp]),
    "dunk_akaze_debug_level": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _pi, _pi, _pi]),
    "dunk_register_frames": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.c_size_t, _f, _d, _i, _vp]),
    "dunk_register_workspace_bytes": (C.c_size_t, [_vp, _i, _i, _i]),
    "dunk_register_frames_dev": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, C.c_size_t, _f, _d, _i, _vp, C.c_size_t, _vp]),
    "dunk_pipeline_workspace_bytes": (C.c_size_t, [_vp, _i, _i, _i]),
    "dunk_pipeline_extract_dev": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, C.c_size_t, _i, _vp, C.c_size_t, _vp]),
    "dunk_pipeline_finish_dev": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i64, _i, _vp, _u32, _f, _d, _vp, C.c_size_t, _vp]),
    "dunk_db_append_dev": (_i, [_vp, _i, _vp, _vp, _vp, _i64]),
    "dunk_memcpy_dev": (_i, [_vp, _i, _vp, _vp, C.c_size_t]),
    "dunk_db_keypoints_dev": (_vp, [_vp]),
    "dunk_db_descriptors_dev": (_vp, [_vp]),
    "dunk_db_append_tiles": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.c_size_t, _vp, _vp, _vp, _vp, _i, _vp]),
    "dunk_db_build_from_bands": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _pi, _pi, _pi]),
    "dunk_db_build_from_bands_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _pi, _pi, _pi]),
    "dunk_db_build_from_bands_part_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _pi, _pi, _pi]),
    "dunk_find_homography": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp, _vp, _pi]),
    "dunk_find_homography_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _vp, _vp, _vp]),
    "dunk_ransac_score_hypotheses": (_i, [_vp, _vp, _vp, _i, _vp, _i, _d, _vp, _vp]),
    "dunk_warp_perspective": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "dunk_warp_perspective_batch_dev": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "dunk_band_merger": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _i, _vp]),
    "dunk_band_merger_dev": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _vp, _i, _vp]),
    "dunk_raster_to_mat": (_i, [_vp, _vp, _i, _i, _vp]),
    "dunk_elevation_create": (_i, [_vp, _vp, _vp, _vp, _i, _i, C.POINTER(_vp)]),
    "dunk_elevation_destroy": (None, [_vp]),
    "dunk_world_coordinates": (_i, [_vp, _vp, _vp, _i64, _vp, _pi]),
    "dunk_pnp_ransac": (_i, [_vp, _vp, _vp, _i, _vp, _i, _f, _d, _i, _vp, _vp, _vp, _i, _pi, _pi]),
    "dunk_pnp_ransac_batch": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _f, _d, _i, _vp, _vp, _vp, _vp]),
    "dunk_pnp_score_hypotheses": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _d, _vp, _vp]),
    "dunk_register_frames_pose": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.c_size_t, _f, _d, _i, _vp, _vp, _vp]),
    "dunk_register_frames_pose_dev": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, C.c_size_t, _f, _d, _i, _vp, _vp, C.c_size_t, _vp, _vp]),
    "dunk_shard_unique_id": (_i, [_vp]),
    "dunk_shard_group_create": (_i, [_vp, _i, _i, _vp, C.POINTER(_vp)]),
    "dunk_shard_group_destroy": (None, [_vp]),
    "dunk_shard_group_rank": (_i, [_vp]),
    "dunk_shard_group_world": (_i, [_vp]),
    "dunk_nccl_version": (_i, []),
    "dunk_shard_group_balance": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "dunk_shard_group_total_rows": (_i64, [_vp]),
    "dunk_shard_group_base": (_i64, [_vp, _i]),
    "dunk_db_match_sharded": (_i, [_vp, _vp, _vp, _i, _u32, _f, _vp, _i, _pi]),
    "dunk_db_match_sharded_dev": (_i, [_vp, _vp, _i, _vp, _i, _u32, _f, _vp, _vp, _vp]),
    "dunk_register_sharded_workspace_bytes": (C.c_size_t, [_vp, _i, _i, _i]),
    "dunk_register_frames_sharded_dev": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, C.c_size_t, _f, _d, _i, _vp, _vp, C.c_size_t,
                                              _vp, _vp]),
    "dunk_memcpy_h2d": (_i, [_vp, _i, _vp, _vp, C.c_size_t]),
    "dunk_memcpy_d2h": (_i, [_vp, _i, _vp, _vp, C.c_size_t]),
    "dunk_dev_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "dunk_dev_free": (_i, [_vp, _vp]),
    "dunk_host_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "dunk_host_free": (_i, [_vp, _vp]),
    "dunk_db_desc_bytes": (_i, [_vp]),
}


def load():
    """dlopen the in-tree library; raises if it has not been built (no fallback)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DunkError(ERR_CUDA, f"{LIB_PATH} not built; run `python -c 'import "
                            "__graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int):
    if rc != 0:
        raise DunkError(rc, load().dunk_last_error().decode("utf-8", "replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """`dunk_ctx`: one per process and GPU; owns the stream/workspace slots."""

    def __init__(self, device: int = 0, n_slots: int = 4):
        lib = load()
        h = C.c_void_p()
        check(lib.dunk_ctx_create(device, n_slots, C.byref(h)))
        self._h = h
        self.device = device
        self.n_slots = n_slots

    @property
    def handle(self):
        if self._h is None:
            raise DunkError(ERR_BAD_ARG, "context destroyed")
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            load().dunk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stream(self, slot: int) -> int:
        return int(load().dunk_ctx_stream(self.handle, slot) or 0)

    @property
    def sm_count(self) -> int:
        return load().dunk_ctx_sm_count(self.handle)

    @property
    def launch_count(self) -> int:
        return int(load().dunk_ctx_launch_count(self.handle))

    def reserve_slot(self) -> int:
        s = load().dunk_ctx_reserve_slot(self.handle)
        if s < 0:
            check(s)
        return s

    def release_slot(self, slot: int):
        check(load().dunk_ctx_release_slot(self.handle, slot))

    def sync(self, slot: int):
        check(load().dunk_sync(self.handle, slot))

    def microbench_popc(self, iters: int = 4096) -> float:
        """Measured POPC-pipe peak in Tpopc/s (roofline denominator of the Hamming matcher)."""
        v = C.c_double()
        check(load().dunk_microbench_popc(self.handle, iters, C.byref(v)))
        return float(v.value)

    def timer_begin(self, slot: int):
        check(load().dunk_timer_begin(self.handle, slot))

    def timer_end(self, slot: int) -> float:
        ms = C.c_float()
        check(load().dunk_timer_end(self.handle, slot, C.byref(ms)))
        return float(ms.value)


class DeviceBuffer:
    """`nbytes` of HBM owned by the caller (dunk_dev_alloc): what the `_dev` entry points take."""

    def __init__(self, ctx: Context, nbytes: int):
        p = C.c_void_p()
        check(load().dunk_dev_alloc(ctx.handle, int(nbytes), C.byref(p)))
        self.ctx, self.ptr, self.nbytes = ctx, int(p.value), int(nbytes)

    def free(self):
        if getattr(self, "ptr", 0) and self.ctx._h is not None:
            load().dunk_dev_free(self.ctx.handle, C.c_void_p(self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def upload(self, slot: int, host: np.ndarray, offset: int = 0):
        """async H2D on the slot's stream (truly asynchronous only from a PinnedBuffer's array)"""
        check(load().dunk_memcpy_h2d(self.ctx.handle, slot, C.c_void_p(self.ptr + offset), host.ctypes.data_as(C.c_void_p),
                                     host.nbytes))

    def download(self, slot: int, host: np.ndarray, offset: int = 0):
        check(load().dunk_memcpy_d2h(self.ctx.handle, slot, host.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr + offset),
                                     host.nbytes))


class PinnedBuffer:
    """page-locked host memory (dunk_host_alloc) exposed as a numpy array"""

    def __init__(self, ctx: Context, shape, dtype=np.uint8):
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        check(load().dunk_host_alloc(ctx.handle, n, C.byref(p)))
        self.ctx, self.ptr = ctx, int(p.value)
        self.array = np.frombuffer((C.c_uint8 * max(n, 1)).from_address(self.ptr), dtype=np.uint8, count=n).view(dt).reshape(shape)

    def free(self):
        if getattr(self, "ptr", 0) and self.ctx._h is not None:
            self.array = None
            load().dunk_host_free(self.ctx.handle, C.c_void_p(self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


_default_ctx = None
_default_lock = threading.Lock()


def default_context() -> Context:
    """Process-wide context on cuda:LOCAL_RANK (what the Rust shim's lazy static would hold)."""
    global _default_ctx
    with _default_lock:
        if _default_ctx is None:
            _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")), 4)
        return _default_ctx


def as_desc(a, name="descriptors") -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim != 2:
        raise DunkError(ERR_BAD_ARG, f"{name}: expected an N x desc_bytes u8 matrix, got shape {a.shape}")
    return a
