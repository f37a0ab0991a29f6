//! Raw FFI of `libdunk_b200.so` — GENERATED from include/dunk_b200.h by tools/gen_rust_sys.py; do not edit.
//! One item per declaration of the header: every `extern "C"` entry point, `#[repr(C)]` struct, enum value and
//! constant.  The safe wrappers live in the `feature_extraction`, `homographier` and `feature_database` shim crates.
#![allow(non_camel_case_types, non_upper_case_globals, non_snake_case, clippy::too_many_arguments)]
use std::os::raw::{c_char, c_int, c_void};

pub const DUNK_OK: c_int = 0;
pub const DUNK_ERR_NO_MEM: c_int = -4;
pub const DUNK_ERR_BAD_ARG: c_int = -5;
pub const DUNK_ERR_VEC_LENGTH: c_int = -28;
pub const DUNK_ERR_OUT_OF_RANGE: c_int = -211;
pub const DUNK_ERR_ASSERT: c_int = -215;
pub const DUNK_ERR_CUDA: c_int = -217;
pub const DUNK_MAX_POINTS_SHIFT: c_int = 18;
pub const DUNK_MAX_POINTS: c_int = (1 << DUNK_MAX_POINTS_SHIFT) - 1;
pub const DUNK_DESC_BYTES: c_int = 61;
pub const DUNK_DESC_STRIDE: c_int = 64;
pub const DUNK_SHARD_ID_BYTES: c_int = 128;

/// `enum DunkHomographyMethod`
pub const DUNK_H_DEFAULT: c_int = 0;
pub const DUNK_H_LMEDS: c_int = 4;
pub const DUNK_H_RANSAC: c_int = 8;
pub const DUNK_H_RHO: c_int = 16;

/// `enum DunkPnPMethod`
pub const DUNK_PNP_ITERATIVE: c_int = 0;
pub const DUNK_PNP_EPNP: c_int = 1;
pub const DUNK_PNP_P3P: c_int = 2;

#[repr(C)] pub struct DunkCtx { _private: [u8; 0] }
#[repr(C)] pub struct DunkDb { _private: [u8; 0] }
#[repr(C)] pub struct DunkElevation { _private: [u8; 0] }
#[repr(C)] pub struct DunkShardGroup { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct DunkKeyPoint {
    pub x: f32,
    pub y: f32,
    pub size: f32,
    pub angle: f32,
    pub response: f32,
    pub octave: i32,
    pub class_id: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct DunkDMatch {
    pub query_idx: i32,
    pub train_idx: i32,
    pub img_idx: i32,
    pub distance: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct DunkTop2 {
    pub d1: u32,
    pub i1: u32,
    pub d2: u32,
    pub i2: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct DunkImage {
    pub id: i32,
    pub x_start: i32,
    pub y_start: i32,
    pub x_end: i32,
    pub y_end: i32,
    pub level_of_detail: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct DunkRowFilter {
    pub image_id: i32,
    pub level_of_detail: i32,
    pub use_box: i32,
    pub x_start: f32,
    pub y_start: f32,
    pub x_end: f32,
    pub y_end: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct DunkRegistration {
    pub H: [f64; 9],
    pub found: i32,
    pub inliers: i32,
    pub matches: i32,
    pub keypoints: i32,
    pub ransac_iters: i32,
    pub hypotheses: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct DunkPoseConfig {
    pub elevation: *const DunkElevation,
    pub K: [f64; 9],
    pub origin: [f64; 3],
    pub method: i32,
    pub iters: i32,
    pub thr: f32,
    pub confidence: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct DunkPose {
    pub rvec: [f64; 3],
    pub tvec: [f64; 3],
    pub found: i32,
    pub inliers: i32,
    pub ransac_iters: i32,
    pub hypotheses: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct DunkPipelineView {
    pub query64_dev: *mut c_void,
    pub query_offsets_dev: *mut c_void,
    pub keypoints_dev: *mut c_void,
    pub keypoint_counts_dev: *mut c_void,
    pub top2_dev: *mut c_void,
    pub total_queries: i32,
    pub keypoint_capacity: i32,
    pub query_capacity: i64,
}

#[link(name = "dunk_b200")]
extern "C" {
    pub fn dunk_ctx_create(device: c_int, n_slots: c_int, out: *mut *mut DunkCtx) -> c_int;
    pub fn dunk_ctx_destroy(ctx: *mut DunkCtx);
    pub fn dunk_last_error() -> *const c_char;
    pub fn dunk_version() -> *const c_char;
    pub fn dunk_ctx_stream(ctx: *mut DunkCtx, slot: c_int) -> *mut c_void;
    pub fn dunk_ctx_device(ctx: *mut DunkCtx) -> c_int;
    pub fn dunk_ctx_sm_count(ctx: *mut DunkCtx) -> c_int;
    pub fn dunk_ctx_launch_count(ctx: *mut DunkCtx) -> u64;
    pub fn dunk_timer_begin(ctx: *mut DunkCtx, slot: c_int) -> c_int;
    pub fn dunk_timer_end(ctx: *mut DunkCtx, slot: c_int, ms: *mut f32) -> c_int;
    pub fn dunk_sync(ctx: *mut DunkCtx, slot: c_int) -> c_int;
    pub fn dunk_ctx_reserve_slot(ctx: *mut DunkCtx) -> c_int;
    pub fn dunk_ctx_release_slot(ctx: *mut DunkCtx, slot: c_int) -> c_int;
    pub fn dunk_knn_match_hamming(ctx: *mut DunkCtx, query: *const u8, nq: c_int, train: *const u8, nt: i64, desc_bytes: c_int, k: c_int, ratio: f32, out: *mut DunkDMatch, out_cap: c_int, n_out: *mut c_int) -> c_int;
    pub fn dunk_knn2_hamming(ctx: *mut DunkCtx, query: *const u8, nq: c_int, train: *const u8, nt: i64, desc_bytes: c_int, idx: *mut i32, dist: *mut i32) -> c_int;
    pub fn dunk_match_crosscheck_hamming(ctx: *mut DunkCtx, query: *const u8, nq: c_int, train: *const u8, nt: i64, desc_bytes: c_int, out: *mut DunkDMatch, out_cap: c_int, n_out: *mut c_int) -> c_int;
    pub fn dunk_knn2_l2(ctx: *mut DunkCtx, query: *const f32, nq: c_int, train: *const f32, nt: i64, dim: c_int, idx: *mut i32, dist: *mut f32, stats: *mut c_int) -> c_int;
    pub fn dunk_knn2_l2_dev(ctx: *mut DunkCtx, slot: c_int, q_dev: *const c_void, nq: c_int, t_dev: *const c_void, nt: i64, dim: c_int, idx_dev: *mut c_void, dist_dev: *mut c_void, stats: *mut c_int) -> c_int;
    pub fn dunk_db_create(ctx: *mut DunkCtx, capacity_rows: i64, desc_bytes: c_int, out: *mut *mut DunkDb) -> c_int;
    pub fn dunk_db_destroy(db: *mut DunkDb);
    pub fn dunk_db_append(db: *mut DunkDb, desc: *const u8, kps: *const DunkKeyPoint, image_ids: *const i32, n: i64) -> c_int;
    pub fn dunk_db_append_random(db: *mut DunkDb, n: i64, seed: u64) -> c_int;
    pub fn dunk_db_append_random_at(db: *mut DunkDb, n: i64, seed: u64, global_row_offset: u64) -> c_int;
    pub fn dunk_db_size(db: *mut DunkDb) -> i64;
    pub fn dunk_db_desc_bytes(db: *mut DunkDb) -> c_int;
    pub fn dunk_db_clear(db: *mut DunkDb) -> c_int;
    pub fn dunk_db_read(db: *mut DunkDb, first: i64, n: i64, desc: *mut u8, kps: *mut DunkKeyPoint, image_ids: *mut i32) -> c_int;
    pub fn dunk_db_match(db: *mut DunkDb, query: *const u8, nq: c_int, ratio: f32, out: *mut DunkDMatch, out_cap: c_int, n_out: *mut c_int) -> c_int;
    pub fn dunk_db_knn2(db: *mut DunkDb, query: *const u8, nq: c_int, index_base: u32, out: *mut DunkTop2) -> c_int;
    pub fn dunk_db_create_image(db: *mut DunkDb, x_start: i32, y_start: i32, x_end: i32, y_end: i32, level_of_detail: i32, id_out: *mut i32) -> c_int;
    pub fn dunk_db_read_image(db: *mut DunkDb, id: i32, out: *mut DunkImage) -> c_int;
    pub fn dunk_db_find_images(db: *mut DunkDb, use_box: c_int, x_start: i32, y_start: i32, x_end: i32, y_end: i32, level_of_detail: i32, ids: *mut i32, cap: c_int, n_out: *mut c_int) -> c_int;
    pub fn dunk_db_image_count(db: *mut DunkDb) -> c_int;
    pub fn dunk_db_select(db: *mut DunkDb, filter: *const DunkRowFilter, limit: i64, out: *mut *mut DunkDb) -> c_int;
    pub fn dunk_db_read_ids(db: *mut DunkDb, first: i64, n: i64, ids: *mut i32) -> c_int;
    pub fn dunk_db_save(db: *mut DunkDb, path: *const c_char) -> c_int;
    pub fn dunk_db_load(ctx: *mut DunkCtx, path: *const c_char, min_capacity_rows: i64, out: *mut *mut DunkDb) -> c_int;
    pub fn dunk_db_knn2_dev(db: *mut DunkDb, slot: c_int, query64_dev: *const c_void, nq: c_int, index_base: u32, top2_dev: *mut c_void) -> c_int;
    pub fn dunk_top2_merge_dev(ctx: *mut DunkCtx, slot: c_int, parts_dev: *const c_void, n_parts: c_int, nq: c_int, merged_dev: *mut c_void) -> c_int;
    pub fn dunk_top2_ratio_dev(ctx: *mut DunkCtx, slot: c_int, merged_dev: *const c_void, nq: c_int, ratio: f32, matches_dev: *mut c_void, count_dev: *mut c_void) -> c_int;
    pub fn dunk_pad_desc_dev(ctx: *mut DunkCtx, slot: c_int, src_dev: *const c_void, n: i64, desc_bytes: c_int, dst64_dev: *mut c_void) -> c_int;
    pub fn dunk_akaze_extract(ctx: *mut DunkCtx, image: *const u8, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, max_points: c_int, kps: *mut DunkKeyPoint, desc: *mut u8, cap: c_int, n_out: *mut c_int) -> c_int;
    pub fn dunk_akaze_extract_batch(ctx: *mut DunkCtx, images: *const u8, n_frames: c_int, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, frame_stride_bytes: usize, max_points: c_int, kps: *mut DunkKeyPoint, desc: *mut u8, cap_per_frame: c_int, counts: *mut c_int) -> c_int;
    pub fn dunk_akaze_debug_level(ctx: *mut DunkCtx, image: *const u8, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, level: c_int, Lt: *mut f32, Lx: *mut f32, Ly: *mut f32, Ldet: *mut f32, kcontrast: *mut f32, level_w: *mut c_int, level_h: *mut c_int, n_levels: *mut c_int) -> c_int;
    pub fn dunk_find_homography(ctx: *mut DunkCtx, src: *const f32, dst: *const f32, n: c_int, method: c_int, thr: f64, H: *mut f64, mask: *mut u8, found: *mut c_int) -> c_int;
    pub fn dunk_find_homography_batch(ctx: *mut DunkCtx, src: *const f32, dst: *const f32, offsets: *const c_int, n_problems: c_int, method: c_int, thr: f64, H: *mut f64, mask: *mut u8, info: *mut c_int) -> c_int;
    pub fn dunk_ransac_score_hypotheses(ctx: *mut DunkCtx, src: *const f32, dst: *const f32, n: c_int, samples: *const c_int, n_hyp: c_int, thr: f64, counts: *mut c_int, Hs: *mut f64) -> c_int;
    pub fn dunk_pnp_ransac(ctx: *mut DunkCtx, obj: *const f64, img: *const f64, n: c_int, K: *const f64, iters: c_int, thr: f32, confidence: f64, method: c_int, rvec: *mut f64, tvec: *mut f64, inliers: *mut i32, inliers_cap: c_int, n_inliers: *mut c_int, found: *mut c_int) -> c_int;
    pub fn dunk_pnp_ransac_batch(ctx: *mut DunkCtx, obj: *const f64, img: *const f64, offsets: *const c_int, n_problems: c_int, K: *const f64, iters: c_int, thr: f32, confidence: f64, method: c_int, rvecs: *mut f64, tvecs: *mut f64, inlier_mask: *mut u8, info: *mut c_int) -> c_int;
    pub fn dunk_pnp_score_hypotheses(ctx: *mut DunkCtx, obj: *const f64, img: *const f64, n: c_int, K: *const f64, samples: *const c_int, n_hyp: c_int, thr: f64, counts: *mut c_int, rt: *mut f64) -> c_int;
    pub fn dunk_warp_perspective(ctx: *mut DunkCtx, src: *const u8, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, M: *const f64, out_rows: c_int, out_cols: c_int, border_value: *const f64, dst: *mut u8) -> c_int;
    pub fn dunk_warp_perspective_batch_dev(ctx: *mut DunkCtx, slot: c_int, src_dev: *const c_void, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, M: *const f64, n: c_int, out_rows: c_int, out_cols: c_int, border_value: *const f64, dst_dev: *mut c_void) -> c_int;
    pub fn dunk_band_merger(ctx: *mut DunkCtx, red: *const f32, green: *const f32, blue: *const f32, n: i64, min_max: *const f64, bgra: c_int, out_rgba: *mut u8) -> c_int;
    pub fn dunk_band_merger_dev(ctx: *mut DunkCtx, slot: c_int, red_dev: *const c_void, green_dev: *const c_void, blue_dev: *const c_void, n: i64, min_max: *const f64, bgra: c_int, out_dev: *mut c_void) -> c_int;
    pub fn dunk_raster_to_mat(ctx: *mut DunkCtx, rgba: *const u8, w: c_int, h: c_int, bgra: *mut u8) -> c_int;
    pub fn dunk_elevation_create(ctx: *mut DunkCtx, gt_dataset: *const f64, gt_elevation: *const f64, heights: *const f64, x_size: c_int, y_size: c_int, out: *mut *mut DunkElevation) -> c_int;
    pub fn dunk_elevation_destroy(e: *mut DunkElevation);
    pub fn dunk_world_coordinates(e: *mut DunkElevation, px: *const f64, py: *const f64, n: i64, xyz: *mut f64, n_missing: *mut c_int) -> c_int;
    pub fn dunk_register_frames(db: *mut DunkDb, images: *const u8, n_frames: c_int, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, frame_stride_bytes: usize, ratio: f32, thr: f64, max_points: c_int, results: *mut DunkRegistration) -> c_int;
    pub fn dunk_register_workspace_bytes(db: *mut DunkDb, n_frames: c_int, rows: c_int, cols: c_int) -> usize;
    pub fn dunk_register_frames_dev(db: *mut DunkDb, slot: c_int, images_dev: *const c_void, n_frames: c_int, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, frame_stride_bytes: usize, ratio: f32, thr: f64, max_points: c_int, workspace_dev: *mut c_void, workspace_bytes: usize, results_dev: *mut c_void) -> c_int;
    pub fn dunk_register_frames_pose(db: *mut DunkDb, images: *const u8, n_frames: c_int, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, frame_stride_bytes: usize, ratio: f32, thr: f64, max_points: c_int, pose: *const DunkPoseConfig, results: *mut DunkRegistration, poses: *mut DunkPose) -> c_int;
    pub fn dunk_register_frames_pose_dev(db: *mut DunkDb, slot: c_int, images_dev: *const c_void, n_frames: c_int, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, frame_stride_bytes: usize, ratio: f32, thr: f64, max_points: c_int, pose: *const DunkPoseConfig, workspace_dev: *mut c_void, workspace_bytes: usize, results_dev: *mut c_void, poses_dev: *mut c_void) -> c_int;
    pub fn dunk_pipeline_workspace_bytes(ctx: *mut DunkCtx, n_frames: c_int, rows: c_int, cols: c_int) -> usize;
    pub fn dunk_pipeline_extract_dev(ctx: *mut DunkCtx, slot: c_int, images_dev: *const c_void, n_frames: c_int, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, frame_stride_bytes: usize, max_points: c_int, workspace_dev: *mut c_void, workspace_bytes: usize, view: *mut DunkPipelineView) -> c_int;
    pub fn dunk_pipeline_finish_dev(ctx: *mut DunkCtx, slot: c_int, n_frames: c_int, rows: c_int, cols: c_int, parts_dev: *const c_void, n_parts: c_int, part_stride_records: i64, total_queries: c_int, db_keypoints_dev: *const c_void, index_base: u32, ratio: f32, thr: f64, workspace_dev: *mut c_void, workspace_bytes: usize, results_dev: *mut c_void) -> c_int;
    pub fn dunk_shard_unique_id(id128: *mut u8) -> c_int;
    pub fn dunk_shard_group_create(ctx: *mut DunkCtx, rank: c_int, world: c_int, id128: *const u8, out: *mut *mut DunkShardGroup) -> c_int;
    pub fn dunk_shard_group_destroy(g: *mut DunkShardGroup);
    pub fn dunk_shard_group_rank(g: *mut DunkShardGroup) -> c_int;
    pub fn dunk_shard_group_world(g: *mut DunkShardGroup) -> c_int;
    pub fn dunk_nccl_version() -> c_int;
    pub fn dunk_shard_group_balance(g: *mut DunkShardGroup, built: *mut DunkDb, shard_out: *mut *mut DunkDb) -> c_int;
    pub fn dunk_shard_group_total_rows(g: *mut DunkShardGroup) -> i64;
    pub fn dunk_shard_group_base(g: *mut DunkShardGroup, rank: c_int) -> i64;
    pub fn dunk_db_match_sharded(g: *mut DunkShardGroup, shard: *mut DunkDb, query: *const u8, nq: c_int, index_base: u32, ratio: f32, out: *mut DunkDMatch, out_cap: c_int, n_out: *mut c_int) -> c_int;
    pub fn dunk_db_match_sharded_dev(g: *mut DunkShardGroup, shard: *mut DunkDb, slot: c_int, query64_dev: *const c_void, nq: c_int, index_base: u32, ratio: f32, top2_merged_dev: *mut c_void, matches_dev: *mut c_void, count_dev: *mut c_void) -> c_int;
    pub fn dunk_register_sharded_workspace_bytes(g: *mut DunkShardGroup, n_frames: c_int, rows: c_int, cols: c_int) -> usize;
    pub fn dunk_register_frames_sharded_dev(g: *mut DunkShardGroup, shard: *mut DunkDb, slot: c_int, images_dev: *const c_void, n_frames: c_int, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, frame_stride_bytes: usize, ratio: f32, thr: f64, max_points: c_int, pose: *const DunkPoseConfig, workspace_dev: *mut c_void, workspace_bytes: usize, results_dev: *mut c_void, poses_dev: *mut c_void) -> c_int;
    pub fn dunk_memcpy_h2d(ctx: *mut DunkCtx, slot: c_int, dst_dev: *mut c_void, src_host: *const c_void, nbytes: usize) -> c_int;
    pub fn dunk_memcpy_d2h(ctx: *mut DunkCtx, slot: c_int, dst_host: *mut c_void, src_dev: *const c_void, nbytes: usize) -> c_int;
    pub fn dunk_dev_alloc(ctx: *mut DunkCtx, nbytes: usize, out_dev: *mut *mut c_void) -> c_int;
    pub fn dunk_dev_free(ctx: *mut DunkCtx, dev: *mut c_void) -> c_int;
    pub fn dunk_host_alloc(ctx: *mut DunkCtx, nbytes: usize, out_host: *mut *mut c_void) -> c_int;
    pub fn dunk_host_free(ctx: *mut DunkCtx, host: *mut c_void) -> c_int;
    pub fn dunk_db_append_dev(db: *mut DunkDb, slot: c_int, desc64_dev: *const c_void, kps_dev: *const c_void, image_ids_dev: *const c_void, n: i64) -> c_int;
    pub fn dunk_db_keypoints_dev(db: *mut DunkDb) -> *const c_void;
    pub fn dunk_db_descriptors_dev(db: *mut DunkDb) -> *const c_void;
    pub fn dunk_memcpy_dev(ctx: *mut DunkCtx, slot: c_int, dst_dev: *mut c_void, src_dev: *const c_void, nbytes: usize) -> c_int;
    pub fn dunk_db_append_tiles(db: *mut DunkDb, images: *const u8, n_tiles: c_int, rows: c_int, cols: c_int, channels: c_int, row_stride_bytes: c_int, frame_stride_bytes: usize, x_off: *const f32, y_off: *const f32, scale: *const f32, image_ids: *const i32, max_points: c_int, counts: *mut c_int) -> c_int;
    pub fn dunk_db_build_from_bands(db: *mut DunkDb, red: *const f32, green: *const f32, blue: *const f32, width: c_int, height: c_int, min_max: *const f64, lods: c_int, resample: c_int, max_points: c_int, n_tiles_out: *mut c_int, tile_w_out: *mut c_int, tile_h_out: *mut c_int) -> c_int;
    pub fn dunk_db_build_from_bands_dev(db: *mut DunkDb, red_dev: *const c_void, green_dev: *const c_void, blue_dev: *const c_void, width: c_int, height: c_int, min_max: *const f64, lods: c_int, resample: c_int, max_points: c_int, n_tiles_out: *mut c_int, tile_w_out: *mut c_int, tile_h_out: *mut c_int) -> c_int;
    pub fn dunk_db_build_from_bands_part_dev(db: *mut DunkDb, red_dev: *const c_void, green_dev: *const c_void, blue_dev: *const c_void, width: c_int, height: c_int, min_max: *const f64, lods: c_int, resample: c_int, max_points: c_int, part: c_int, n_parts: c_int, n_tiles_out: *mut c_int, tile_w_out: *mut c_int, tile_h_out: *mut c_int) -> c_int;
    pub fn dunk_profile_begin(ctx: *mut DunkCtx) -> c_int;
    pub fn dunk_profile_end(ctx: *mut DunkCtx, names: *mut c_char, names_cap: c_int, ms: *mut f64, launches: *mut c_int, alg: *mut f64, cap: c_int) -> c_int;
    pub fn dunk_selftest_gamma_lut(ctx: *mut DunkCtx, mismatches: *mut u64) -> c_int;
    pub fn dunk_microbench_popc(ctx: *mut DunkCtx, iters: c_int, tpopc_per_s: *mut f64) -> c_int;
}

/// message of the calling thread's last failed call (`dunk_last_error`, thread-local in the library)
pub fn last_error() -> String {
    // SAFETY: the library returns a NUL-terminated string that stays valid until the thread's next failing call
    unsafe { std::ffi::CStr::from_ptr(dunk_last_error()) }.to_string_lossy().into_owned()
}

/// Process-wide context on the GPU named by `DUNK_DEVICE` (default 0) with 8 stream / workspace slots: the reference's
/// callers are rayon workers (preprocessor/src/main.rs:233-243), concurrent calls take distinct slots.
/// There is no CPU fallback: without an sm_100 device this panics with the library's message.
pub fn ctx() -> *mut DunkCtx {
    use std::sync::OnceLock;
    struct P(*mut DunkCtx);
    // SAFETY: the context is internally synchronised (slot pool under a mutex); the pointer itself is immutable
    unsafe impl Send for P {}
    unsafe impl Sync for P {}
    static CTX: OnceLock<P> = OnceLock::new();
    CTX.get_or_init(|| {
        let device = std::env::var("DUNK_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut c = std::ptr::null_mut();
        // SAFETY: `c` is a valid out pointer
        let rc = unsafe { dunk_ctx_create(device, 8, &mut c) };
        assert_eq!(rc, 0, "dunk_ctx_create failed ({rc}): {}", last_error());
        P(c)
    })
    .0
}
