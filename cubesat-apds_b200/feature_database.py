"""Host-side mirror of the reference crate `feature_database`, READ / LOAD side only
(feature_database/src/keypointdb.rs:38-109, src/models.rs:30-55, src/schema.rs:27-40).

The reference keeps keypoints as AoS Postgres rows (`descriptor bytea`); here the same rows
live in HBM as SoA arrays (x, y, size, angle, response, octave, class_id, image_id and an
N x 64-B descriptor array), one `DescriptorDatabase` per GPU shard.  Postgres / diesel I/O is
out of scope: `create_keypoint` takes the row fields the reference's
`InsertKeypoint` carries and appends them to the shard.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import DMATCH_DTYPE, KEYPOINT_DTYPE, TOP2_DTYPE, DunkError, check, ptr

# keypointdb.rs:12
OPENCV_KEYPOINT_LIMIT = 2 ** 18 - 1

# models::Keypoint (models.rs:30-41) as a record array: what the read_keypoints_* calls return
DB_KEYPOINT_DTYPE = np.dtype([("id", "<i4"), ("x_coord", "<f4"), ("y_coord", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                              ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4"), ("descriptor", "u1", (_lib.DESC_BYTES,)),
                              ("image_id", "<i4")])


class NotFound(KeyError):
    """diesel::result::Error::NotFound"""


class DescriptorDatabase:
    """One HBM-resident shard of the `keypoint` table (schema.rs:27-40)."""

    def __init__(self, ctx: Optional[_lib.Context] = None, capacity: int = 1 << 20,
                 desc_bytes: int = _lib.DESC_BYTES, _handle=None):
        self.ctx = ctx or _lib.default_context()
        self.desc_bytes = desc_bytes
        if _handle is not None:
            self._h = _handle
            return
        h = C.c_void_p()
        check(_lib.load().dunk_db_create(self.ctx.handle, int(capacity), int(desc_bytes), C.byref(h)))
        self._h = h

    # -- ImageDatabase (feature_database/src/imagedb.rs:14-77): the ref_image table
    def create_image(self, x_start: int, y_start: int, x_end: int, y_end: int, level_of_detail: int) -> int:
        """ImageDatabase::create_image with Image::One(InsertImage{..}) -> new id (1-based)"""
        out = C.c_int(0)
        check(_lib.load().dunk_db_create_image(self.handle, int(x_start), int(y_start), int(x_end), int(y_end),
                                               int(level_of_detail), C.byref(out)))
        return out.value

    def read_image_from_id(self, id: int) -> np.void:
        out = np.zeros(1, dtype=_lib.IMAGE_DTYPE)
        try:
            check(_lib.load().dunk_db_read_image(self.handle, int(id), ptr(out)))
        except DunkError as e:
            if e.code == _lib.ERR_OUT_OF_RANGE:
                raise NotFound(id) from None
            raise
        return out[0]

    def _find_images(self, use_box, x_start, y_start, x_end, y_end, lod) -> List[int]:
        cap = max(1, int(_lib.load().dunk_db_image_count(self.handle)))
        ids = np.zeros(cap, dtype=np.int32)
        n = C.c_int(0)
        check(_lib.load().dunk_db_find_images(self.handle, int(use_box), int(x_start), int(y_start), int(x_end), int(y_end),
                                              int(lod), ptr(ids), cap, C.byref(n)))
        return ids[: n.value].tolist()

    def find_images_from_dimensions(self, x_start: int, y_start: int, x_end: int, y_end: int, level_of_detail: int) -> List[int]:
        """imagedb.rs:39-56"""
        return self._find_images(1, x_start, y_start, x_end, y_end, level_of_detail)

    def find_images_from_lod(self, level_of_detail: int) -> List[int]:
        """imagedb.rs:58-66"""
        return self._find_images(0, 0, 0, 0, 0, level_of_detail)

    # -- KeypointDatabase keyed reads (keypointdb.rs:38-90): filter, ORDER BY response DESC, LIMIT 2^18-1
    def select(self, image_id: int = -1, level_of_detail: int = -1, box=None,
               limit: int = OPENCV_KEYPOINT_LIMIT) -> "DescriptorDatabase":
        """The rows of a keyed read as a new HBM-resident shard (to match against, or to read back)."""
        f = _lib.RowFilter(int(image_id), int(level_of_detail), 0 if box is None else 1, *(box or (0.0, 0.0, 0.0, 0.0)))
        h = C.c_void_p()
        check(_lib.load().dunk_db_select(self.handle, C.byref(f), int(limit), C.byref(h)))
        return DescriptorDatabase(self.ctx, desc_bytes=self.desc_bytes, _handle=h)

    def rows(self) -> np.ndarray:
        """every row as models::Keypoint records"""
        n = len(self)
        out = np.zeros(n, dtype=DB_KEYPOINT_DTYPE)
        if n == 0:
            return out
        d, k, im = self.read_rows(0, n)
        ids = np.zeros(n, dtype=np.int32)
        check(_lib.load().dunk_db_read_ids(self.handle, 0, n, ptr(ids)))
        out["id"], out["descriptor"], out["image_id"] = ids, d, im
        out["x_coord"], out["y_coord"] = k["x"], k["y"]
        for name in ("size", "angle", "response", "octave", "class_id"):
            out[name] = k[name]
        return out

    def _read(self, **kw) -> np.ndarray:
        sub = self.select(**kw)
        try:
            return sub.rows()
        finally:
            sub.close()

    def read_keypoints_from_image_id(self, image_id: int) -> np.ndarray:
        """keypointdb.rs:38-49"""
        return self._read(image_id=image_id)

    def read_keypoints_from_lod(self, level_of_detail: int) -> np.ndarray:
        """keypointdb.rs:51-66"""
        return self._read(level_of_detail=level_of_detail)

    def read_keypoints_from_coordinates(self, x_start: float, y_start: float, x_end: float, y_end: float,
                                        level_of_detail: int) -> np.ndarray:
        """keypointdb.rs:68-90 (bounds are floor()/ceil()-ed, inclusive)"""
        return self._read(level_of_detail=level_of_detail, box=(float(x_start), float(y_start), float(x_end), float(y_end)))

    def read_keypoint_from_id(self, id: int) -> np.void:
        """keypointdb.rs:28-36 — the `id` column: 1 + row index in a base DB, the source row's id in a
        `select()` result (looked up through dunk_db_read_ids)"""
        n = len(self)
        ids = np.zeros(n, dtype=np.int32)
        if n:
            check(_lib.load().dunk_db_read_ids(self.handle, 0, n, ptr(ids)))
        hit = np.nonzero(ids == id)[0]
        if hit.size == 0:
            raise NotFound(id)
        d, k, im = self.read_rows(int(hit[0]), 1)
        out = np.zeros(1, dtype=DB_KEYPOINT_DTYPE)
        out["id"], out["descriptor"], out["image_id"] = id, d, im
        out["x_coord"], out["y_coord"] = k["x"], k["y"]
        for name in ("size", "angle", "response", "octave", "class_id"):
            out[name] = k[name]
        return out[0]

    # -- flat dump / load
    def save(self, path: str):
        check(_lib.load().dunk_db_save(self.handle, str(path).encode()))

    @classmethod
    def load(cls, path: str, ctx: Optional[_lib.Context] = None, min_capacity: int = 0) -> "DescriptorDatabase":
        ctx = ctx or _lib.default_context()
        h = C.c_void_p()
        check(_lib.load().dunk_db_load(ctx.handle, str(path).encode(), int(min_capacity), C.byref(h)))
        return cls(ctx, desc_bytes=int(_lib.load().dunk_db_desc_bytes(h)), _handle=h)

    @property
    def handle(self):
        if self._h is None:
            raise DunkError(_lib.ERR_BAD_ARG, "database closed")
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().dunk_db_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        """drop every keypoint and ref_image row, keep the HBM buffers (the migrations' down.sql + up.sql)"""
        check(_lib.load().dunk_db_clear(self.handle))

    def __len__(self) -> int:
        return int(_lib.load().dunk_db_size(self.handle))

    # -- write side: what `create_keypoint(Keypoint::Multiple(..))` inserts (keypointdb.rs:100-109)
    def append(self, descriptors: np.ndarray, keypoints: Optional[np.ndarray] = None,
               image_ids: Optional[np.ndarray] = None):
        d = _lib.as_desc(descriptors)
        if d.shape[0] and d.shape[1] != self.desc_bytes:
            raise DunkError(_lib.ERR_ASSERT, f"descriptor width {d.shape[1]} != {self.desc_bytes}")
        k = None if keypoints is None else np.ascontiguousarray(keypoints, dtype=KEYPOINT_DTYPE)
        if image_ids is not None and np.isscalar(image_ids):
            image_ids = np.full(d.shape[0], image_ids, dtype=np.int32)
        im = None if image_ids is None else np.ascontiguousarray(image_ids, dtype=np.int32)
        for a in (k, im):
            if a is not None and a.shape[0] != d.shape[0]:
                raise DunkError(_lib.ERR_VEC_LENGTH, "row-count mismatch between columns")
        check(_lib.load().dunk_db_append(self.handle, ptr(d), ptr(k), ptr(im), d.shape[0]))

    def append_tiles(self, tiles: np.ndarray, x_off=None, y_off=None, scale=None, image_ids=None,
                     max_points: int = _lib.MAX_POINTS) -> np.ndarray:
        """Reference-DB build for a tile batch [T, H, W(, C)] u8: AKAZE on every tile, rows appended
        with keypoint coordinates mapped to scene pixels `x * scale + x_off` — what
        preprocessor::feature_extraction_to_database does per tile (preprocessor/src/main.rs:248-327)
        minus GDAL / Postgres.  Returns the rows added per tile."""
        a = np.asarray(tiles)
        if a.ndim == 3:
            a = a[..., None]
        if a.dtype != np.uint8 or a.ndim != 4:
            raise DunkError(_lib.ERR_ASSERT, f"tile batch shape {a.shape} / dtype {a.dtype} unsupported")
        a = np.ascontiguousarray(a)
        T, rows, cols, ch = a.shape

        def col(v, dt):
            return None if v is None else np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=dt), (T,)))
        xo, yo, sc, ids = col(x_off, np.float32), col(y_off, np.float32), col(scale, np.float32), col(image_ids, np.int32)
        counts = np.zeros(T, dtype=np.int32)
        check(_lib.load().dunk_db_append_tiles(self.handle, ptr(a), T, rows, cols, ch, cols * ch, rows * cols * ch,
                                               ptr(xo), ptr(yo), ptr(sc), ptr(ids), int(max_points), ptr(counts)))
        return counts

    def build_from_bands(self, red: np.ndarray, green: np.ndarray, blue: np.ndarray, min_max, lods: int,
                         resample: str = "area", max_points: int = _lib.MAX_POINTS):
        """The preprocessor's DB build (preprocessor/src/main.rs:160-327) for a scene given as three f32
        bands [H, W]: LoD windows resampled to tile size, band_merger, AKAZE, rows + ref_image rows
        inserted — all on the device.  Returns (tiles processed, (tile_w, tile_h))."""
        r, g, b = (np.ascontiguousarray(x, dtype=np.float32) for x in (red, green, blue))
        if not (r.ndim == 2 and r.shape == g.shape == b.shape):
            raise DunkError(_lib.ERR_ASSERT, "build_from_bands: three equally shaped 2-d f32 bands expected")
        mm = np.ascontiguousarray(min_max, dtype=np.float64)
        n, tw, th = C.c_int(0), C.c_int(0), C.c_int(0)
        check(_lib.load().dunk_db_build_from_bands(self.handle, ptr(r), ptr(g), ptr(b), r.shape[1], r.shape[0], ptr(mm), int(lods),
                                                   {"area": 0, "lanczos": 1}[resample], int(max_points), C.byref(n), C.byref(tw),
                                                   C.byref(th)))
        return n.value, (tw.value, th.value)

    def register_frames(self, frames: np.ndarray, ratio: float = 0.8, reproj_threshold: float = 3.0,
                        max_points: int = _lib.MAX_POINTS, pose: Optional["PoseStage"] = None):
        """The whole hot path for a frame batch [B, H, W(, C)] u8 against this shard: extract ->
        2-NN + Lowe ratio -> RANSAC homography.  Returns REGISTRATION_DTYPE records (H maps frame
        pixels to scene pixels); with `pose` (a PoseStage) also POSE_DTYPE records — the attitude from
        get_world_coordinates + pnp_solver_ransac over the correspondences the homography kept."""
        a = np.asarray(frames)
        if a.ndim == 3:
            a = a[..., None]
        if a.dtype != np.uint8 or a.ndim != 4:
            raise DunkError(_lib.ERR_ASSERT, f"frame batch shape {a.shape} / dtype {a.dtype} unsupported")
        a = np.ascontiguousarray(a)
        B, rows, cols, ch = a.shape
        out = np.zeros(B, dtype=_lib.REGISTRATION_DTYPE)
        if pose is None:
            check(_lib.load().dunk_register_frames(self.handle, ptr(a), B, rows, cols, ch, cols * ch, rows * cols * ch,
                                                   float(ratio), float(reproj_threshold), int(max_points), ptr(out)))
            return out
        poses = np.zeros(B, dtype=_lib.POSE_DTYPE)
        check(_lib.load().dunk_register_frames_pose(self.handle, ptr(a), B, rows, cols, ch, cols * ch, rows * cols * ch,
                                                    float(ratio), float(reproj_threshold), int(max_points), C.byref(pose.config),
                                                    ptr(out), ptr(poses)))
        return out, poses

    def append_random(self, n: int, seed: int, global_row_offset: Optional[int] = None):
        if global_row_offset is None:
            check(_lib.load().dunk_db_append_random(self.handle, int(n), int(seed)))
        else:
            check(_lib.load().dunk_db_append_random_at(self.handle, int(n), int(seed), int(global_row_offset)))

    # -- read side
    def read_descriptors(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.desc_bytes), dtype=np.uint8)
        check(_lib.load().dunk_db_read(self.handle, first, n, ptr(out), None, None))
        return out

    def read_rows(self, first: int, n: int):
        d = np.empty((n, self.desc_bytes), dtype=np.uint8)
        k = np.empty(n, dtype=KEYPOINT_DTYPE)
        im = np.empty(n, dtype=np.int32)
        check(_lib.load().dunk_db_read(self.handle, first, n, ptr(d), ptr(k), ptr(im)))
        return d, k, im

    # -- matching against the shard
    def match(self, query_desc: np.ndarray, ratio: float) -> np.ndarray:
        q = _lib.as_desc(query_desc)
        out = np.empty(max(q.shape[0], 1), dtype=DMATCH_DTYPE)
        n = C.c_int(0)
        check(_lib.load().dunk_db_match(self.handle, ptr(q), q.shape[0], float(ratio), ptr(out),
                                        out.shape[0], C.byref(n)))
        return out[: n.value].copy()

    def knn2(self, query_desc: np.ndarray, index_base: int = 0) -> np.ndarray:
        q = _lib.as_desc(query_desc)
        out = np.empty(q.shape[0], dtype=TOP2_DTYPE)
        check(_lib.load().dunk_db_knn2(self.handle, ptr(q), q.shape[0], int(index_base), ptr(out)))
        return out


class Geotransform:
    """`feature_database::elevationdb::geotransform` + `elevation` (elevationdb.rs:12-104, 182-244): the
    "dataset" and "elevation" GDAL geotransforms and the elevation raster, resident in HBM, so that
    `get_world_coordinates` (pixel -> ECEF object point for pnp_solver_ransac) runs batched on the GPU."""

    def __init__(self, dataset_transform, elevation_transform=None, heights: Optional[np.ndarray] = None,
                 ctx: Optional[_lib.Context] = None):
        self.ctx = ctx or _lib.default_context()
        gd = np.ascontiguousarray(dataset_transform, dtype=np.float64).reshape(6)
        ge = None if elevation_transform is None else np.ascontiguousarray(elevation_transform, dtype=np.float64).reshape(6)
        hh, xs, ys = None, 0, 0
        if heights is not None:
            hh = np.ascontiguousarray(heights, dtype=np.float64)
            ys, xs = hh.shape
        h = C.c_void_p()
        check(_lib.load().dunk_elevation_create(self.ctx.handle, ptr(gd), ptr(ge), ptr(hh), xs, ys, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().dunk_elevation_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def world_coordinates(self, x, y):
        """batched get_world_coordinates: returns (xyz [n, 3] f64, number of points without an elevation sample)"""
        px = np.ascontiguousarray(x, dtype=np.float64).ravel()
        py = np.ascontiguousarray(y, dtype=np.float64).ravel()
        out = np.empty((px.shape[0], 3), dtype=np.float64)
        miss = C.c_int(0)
        check(_lib.load().dunk_world_coordinates(self._h, ptr(px), ptr(py), px.shape[0], ptr(out), C.byref(miss)))
        return out, miss.value

    def get_world_coordinates(self, x: float, y: float):
        """elevationdb.rs:64-90 — one point; a missing elevation sample is the reference's diesel NotFound"""
        xyz, miss = self.world_coordinates([x], [y])
        if miss:
            raise NotFound((x, y))
        return tuple(float(v) for v in xyz[0])


class PoseStage:
    """Arguments of the pose stage (DunkPoseConfig): the scene's Geotransform (pixel -> ECEF), the camera matrix and the
    pnp_solver_ransac parameters (homographier mod.rs:320-328: iter_count, reproj_thres, confidence, method)."""

    def __init__(self, geotransform: Geotransform, camera_matrix, origin=(0.0, 0.0, 0.0), iter_count: int = 1000,
                 reproj_thres: float = 3.0, confidence: float = 0.99, method: int = 1):
        self.geotransform = geotransform            # keeps the handle alive
        K = np.ascontiguousarray(camera_matrix, dtype=np.float64).reshape(9)
        o = np.ascontiguousarray(origin, dtype=np.float64).reshape(3)
        self.config = _lib.PoseConfig(geotransform._h, (C.c_double * 9)(*K), (C.c_double * 3)(*o), int(method), int(iter_count),
                                      float(reproj_thres), float(confidence))


class ShardGroup:
    """`dunk_shard_group`: the ranks holding the row-range shards of the reference DB, one process per GPU, NCCL
    inside the library.  Rank 0 creates the id (`ShardGroup.unique_id()`), the host application
    hands it to the other ranks, every rank constructs the group (collective)."""

    def __init__(self, ctx: _lib.Context, rank: int, world: int, unique_id: Optional[bytes] = None):
        self.ctx = ctx
        h = C.c_void_p()
        idbuf = None
        if world > 1:
            if unique_id is None or len(unique_id) != _lib.SHARD_ID_BYTES:
                raise DunkError(_lib.ERR_BAD_ARG, "ShardGroup: world > 1 needs rank 0's 128-byte unique id")
            idbuf = (C.c_uint8 * _lib.SHARD_ID_BYTES).from_buffer_copy(unique_id)
        check(_lib.load().dunk_shard_group_create(ctx.handle, int(rank), int(world), idbuf, C.byref(h)))
        self._h, self.rank, self.world = h, int(rank), int(world)

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * _lib.SHARD_ID_BYTES)()
        check(_lib.load().dunk_shard_unique_id(buf))
        return bytes(buf)

    @property
    def handle(self):
        if self._h is None:
            raise DunkError(_lib.ERR_BAD_ARG, "shard group destroyed")
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.load().dunk_shard_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def balance(self, built: "DescriptorDatabase") -> "DescriptorDatabase":
        """collective: re-cut the ranks' locally built rows into equal contiguous row ranges; returns this rank's shard"""
        h = C.c_void_p()
        check(_lib.load().dunk_shard_group_balance(self.handle, built.handle, C.byref(h)))
        return DescriptorDatabase(self.ctx, desc_bytes=built.desc_bytes, _handle=h)

    @property
    def total_rows(self) -> int:
        return int(_lib.load().dunk_shard_group_total_rows(self.handle))

    def base(self, rank: int) -> int:
        return int(_lib.load().dunk_shard_group_base(self.handle, int(rank)))

    def match(self, shard: "DescriptorDatabase", query_desc: np.ndarray, ratio: float, index_base: int) -> np.ndarray:
        """collective get_knn_matches against the sharded DB (host buffers; every rank passes the same queries)"""
        q = _lib.as_desc(query_desc)
        out = np.empty(max(q.shape[0], 1), dtype=DMATCH_DTYPE)
        n = C.c_int(0)
        check(_lib.load().dunk_db_match_sharded(self.handle, shard.handle, ptr(q), q.shape[0], int(index_base), float(ratio),
                                                ptr(out), out.shape[0], C.byref(n)))
        return out[: n.value].copy()

    def register_frames(self, shard: "DescriptorDatabase", frames: np.ndarray, ratio: float = 0.8,
                        reproj_threshold: float = 3.0, max_points: int = _lib.MAX_POINTS, pose: Optional[PoseStage] = None):
        """collective: this rank's frame batch [B, H, W(, C)] u8 through extract -> sharded match -> RANSAC (-> PnP)"""
        lib = _lib.load()
        a = np.asarray(frames)
        if a.ndim == 3:
            a = a[..., None]
        a = np.ascontiguousarray(a)
        B, rows, cols, ch = a.shape
        ctx = self.ctx
        f_dev = _lib.DeviceBuffer(ctx, a.nbytes)
        ws_bytes = int(lib.dunk_register_sharded_workspace_bytes(self.handle, B, rows, cols))
        ws = _lib.DeviceBuffer(ctx, ws_bytes)
        res_dev = _lib.DeviceBuffer(ctx, B * _lib.REGISTRATION_DTYPE.itemsize)
        pose_dev = _lib.DeviceBuffer(ctx, B * _lib.POSE_DTYPE.itemsize)
        res = np.zeros(B, dtype=_lib.REGISTRATION_DTYPE)
        poses = np.zeros(B, dtype=_lib.POSE_DTYPE)
        slot = None
        try:
            slot = ctx.reserve_slot()
            f_dev.upload(slot, a)
            check(lib.dunk_register_frames_sharded_dev(self.handle, shard.handle, slot, C.c_void_p(f_dev.ptr), B, rows, cols, ch,
                                                       cols * ch, rows * cols * ch, float(ratio), float(reproj_threshold),
                                                       int(max_points), C.byref(pose.config) if pose else None,
                                                       C.c_void_p(ws.ptr), ws_bytes, C.c_void_p(res_dev.ptr),
                                                       C.c_void_p(pose_dev.ptr) if pose else None))
            res_dev.download(slot, res)
            if pose:
                pose_dev.download(slot, poses)
            ctx.sync(slot)
        finally:
            if slot is not None:
                ctx.sync(slot)
                ctx.release_slot(slot)
            for b in (f_dev, ws, res_dev, pose_dev):
                b.free()
        return (res, poses) if pose else res


def merge_top2(ctx: _lib.Context, parts: Sequence[np.ndarray]) -> np.ndarray:
    """(distance, index)-lexicographic merge of per-shard top-2 records on the GPU — the step
    that follows the allgather in the sharded matcher.  Host arrays in/out."""
    import torch  # device memory plumbing only
    nq = parts[0].shape[0]
    stacked = np.ascontiguousarray(np.stack([np.asarray(p, dtype=TOP2_DTYPE) for p in parts]))
    dev = torch.device("cuda", ctx.device)
    t = torch.from_numpy(stacked.view(np.uint8).reshape(-1)).to(dev)
    out = torch.empty(nq * 16, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    slot = ctx.reserve_slot()
    try:
        check(_lib.load().dunk_top2_merge_dev(ctx.handle, slot, t.data_ptr(), len(parts), nq, out.data_ptr()))
        ctx.sync(slot)
    finally:
        ctx.release_slot(slot)
    return out.cpu().numpy().view(TOP2_DTYPE).copy()


def register_frames_sharded_local(ctx: _lib.Context, shards: Sequence["DescriptorDatabase"], frames: np.ndarray,
                                  ratio: float = 0.8, reproj_threshold: float = 3.0,
                                  max_points: int = _lib.MAX_POINTS) -> np.ndarray:
    """The sharded pipeline's three phases with every shard on THIS GPU: phase 1 extract,
    phase 2 local top-2 against each shard (global row indices = shard bases), phase 3 merge + ratio +
    RANSAC.  bench.py runs the same phases with one shard per rank and NCCL between them; this
    single-process form is what the parity test compares with the unsharded `register_frames`."""
    import torch  # device memory plumbing only
    lib = _lib.load()
    a = np.asarray(frames)
    if a.ndim == 3:
        a = a[..., None]
    a = np.ascontiguousarray(a)
    B, rows, cols, ch = a.shape
    dev = torch.device("cuda", ctx.device)
    f_dev = torch.from_numpy(a.reshape(-1)).to(dev)
    ws_bytes = int(lib.dunk_pipeline_workspace_bytes(ctx.handle, B, rows, cols))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    res = torch.zeros(B * _lib.REGISTRATION_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    # replicated keypoint column of all shards (global row index -> keypoint)
    sizes = [len(s) for s in shards]
    bases = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    kps_all = torch.empty(int(bases[-1]) * 28, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    slot = None
    try:
        slot = ctx.reserve_slot()
        for s, b0, n in zip(shards, bases, sizes):
            if n:
                check(lib.dunk_memcpy_dev(ctx.handle, slot, kps_all.data_ptr() + int(b0) * 28,
                                          lib.dunk_db_keypoints_dev(s.handle), n * 28))
        view = _lib.PipelineView()
        check(lib.dunk_pipeline_extract_dev(ctx.handle, slot, f_dev.data_ptr(), B, rows, cols, ch, cols * ch,
                                            rows * cols * ch, int(max_points), ws.data_ptr(), ws_bytes, C.byref(view)))
        nq = view.total_queries
        parts = torch.empty(len(shards) * max(nq, 1) * 16, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        for i, s in enumerate(shards):
            check(lib.dunk_db_knn2_dev(s.handle, slot, view.query64_dev, nq, int(bases[i]),
                                       parts.data_ptr() + i * max(nq, 1) * 16))
        check(lib.dunk_pipeline_finish_dev(ctx.handle, slot, B, rows, cols, parts.data_ptr(), len(shards), max(nq, 1), nq,
                                           kps_all.data_ptr(), 0, float(ratio), float(reproj_threshold), ws.data_ptr(),
                                           ws_bytes, res.data_ptr()))
        ctx.sync(slot)
    finally:
        if slot is not None:
            ctx.release_slot(slot)
    return res.cpu().numpy().view(_lib.REGISTRATION_DTYPE).copy()
