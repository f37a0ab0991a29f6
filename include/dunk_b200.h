/*
 * dunk_b200.h — C ABI of libdunk_b200.so, the sm_100a implementation of DUNK's
 * image-to-reference registration hot path (extract -> 2-NN match -> RANSAC).
 *
 * This is the drop-in boundary: every entry point below is what the reference's Rust
 * crates would bind through `extern "C"` in place of the `opencv` crate call they make
 * today.  Citations are file:line under the reference tree (Murmeldyret/cubesat-APDS).
 *
 * Conventions
 *   - All pointers are HOST pointers unless the name ends in `_dev`.
 *   - Outputs are caller-allocated with an explicit capacity; the count comes back
 *     through an out parameter.  Nothing allocated by the library crosses the ABI
 *     except the opaque handles (dunk_ctx, dunk_db).
 *   - Return value: 0 = ok, negative = OpenCV-compatible status code (the reference
 *     surfaces `opencv::Error{code,message}`): DUNK_ERR_ASSERT (-215),
 *     DUNK_ERR_OUT_OF_RANGE (-211), DUNK_ERR_VEC_LENGTH (-28), DUNK_ERR_BAD_ARG (-5),
 *     DUNK_ERR_NO_MEM (-4), DUNK_ERR_CUDA (-217, "GpuApiCallError").
 *     The message is available from dunk_last_error() (thread-local).
 *   - Calls are synchronous and re-entrant: a context owns a pool of CUDA streams +
 *     workspaces; concurrent callers (the reference calls from rayon workers,
 *     preprocessor/src/main.rs:233-243) each take one slot and synchronise only it.
 *   - There is no CPU fallback: without a CUDA device dunk_ctx_create fails.
 */
#ifndef DUNK_B200_H
#define DUNK_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUNK_OK 0
#define DUNK_ERR_NO_MEM (-4)
#define DUNK_ERR_BAD_ARG (-5)
#define DUNK_ERR_VEC_LENGTH (-28)
#define DUNK_ERR_OUT_OF_RANGE (-211)
#define DUNK_ERR_ASSERT (-215)
#define DUNK_ERR_CUDA (-217)

/* reference: feature_extraction/src/lib.rs:12-13 */
#define DUNK_MAX_POINTS_SHIFT 18
#define DUNK_MAX_POINTS ((1 << DUNK_MAX_POINTS_SHIFT) - 1)
#define DUNK_DESC_BYTES 61 /* MLDB-486 -> 61 bytes (lib.rs:64-73 descriptor_size=0, channels=3) */
#define DUNK_DESC_STRIDE 64 /* HBM row stride (16 x u32) */

/* cv::KeyPoint layout (28 B) — the element type of Vector<KeyPoint>, lib.rs:15-18, :33-59 */
typedef struct DunkKeyPoint {
    float x, y;      /* pt */
    float size;
    float angle;     /* degrees */
    float response;
    int32_t octave;
    int32_t class_id;
} DunkKeyPoint;

/* cv::DMatch layout (16 B) — element of Vector<DMatch>, lib.rs:94-126 */
typedef struct DunkDMatch {
    int32_t query_idx;
    int32_t train_idx;
    int32_t img_idx;
    float distance;
} DunkDMatch;

/* device-side local/global top-2 record: one per query (16 B) */
typedef struct DunkTop2 {
    uint32_t d1, i1, d2, i2; /* distance, row index; empty slot = 0xFFFFFFFF */
} DunkTop2;

/* homographier/src/homographier/mod.rs:25-31 */
enum DunkHomographyMethod { DUNK_H_DEFAULT = 0, DUNK_H_LMEDS = 4, DUNK_H_RANSAC = 8, DUNK_H_RHO = 16 };
/* opencv SolvePnPMethod values used by mod.rs:320-361 */
enum DunkPnPMethod { DUNK_PNP_ITERATIVE = 0, DUNK_PNP_EPNP = 1, DUNK_PNP_P3P = 2 };

typedef struct dunk_ctx dunk_ctx;
typedef struct dunk_db dunk_db;
typedef struct dunk_elevation dunk_elevation;

/* ---- context ------------------------------------------------------------------ */
int dunk_ctx_create(int device, int n_slots, dunk_ctx** out);
void dunk_ctx_destroy(dunk_ctx* ctx);
const char* dunk_last_error(void);
const char* dunk_version(void);
/* cudaStream_t of slot i (for callers that time with their own CUDA events) */
void* dunk_ctx_stream(dunk_ctx* ctx, int slot);
int dunk_ctx_device(dunk_ctx* ctx);
int dunk_ctx_sm_count(dunk_ctx* ctx);
/* number of kernels this context has launched so far (bench.py `gpu_launches`) */
uint64_t dunk_ctx_launch_count(dunk_ctx* ctx);
/* CUDA-event timing on the library's own stream: begin/end bracket, returns ms */
int dunk_timer_begin(dunk_ctx* ctx, int slot);
int dunk_timer_end(dunk_ctx* ctx, int slot, float* ms);
int dunk_sync(dunk_ctx* ctx, int slot);
/* take a slot out of the pool for the caller's exclusive use with the `_dev` entry points
 * (returns the slot index, or a negative status); give it back with dunk_ctx_release_slot */
int dunk_ctx_reserve_slot(dunk_ctx* ctx);
int dunk_ctx_release_slot(dunk_ctx* ctx, int slot);

/* ---- stage 2: Hamming brute-force matching -------------------------------------- */
/* replaces BFMatcher(NORM_HAMMING,false).knnMatch + Lowe ratio filter,
 * feature_extraction/src/lib.rs:94-114.  query/train: n x desc_bytes u8 row-major.
 * Keeps m[0] iff (float)d0 < (float)d1 * ratio (f32, strict).  k < 2 or nt < 2 ->
 * DUNK_ERR_OUT_OF_RANGE (the reference's `i.get(1)?`, lib.rs:108); nt >= 2^18 is
 * accepted (OpenCV asserts; the HBM path uses 32-bit row indices). */
int dunk_knn_match_hamming(dunk_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train,
                           int64_t nt, int desc_bytes, int k, float ratio, DunkDMatch* out,
                           int out_cap, int* n_out);
/* unfiltered 2-NN lists (what knn_train_match_def returns, lib.rs:103): idx/dist nq x 2 */
int dunk_knn2_hamming(dunk_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train,
                      int64_t nt, int desc_bytes, int32_t* idx, int32_t* dist);
/* replaces BFMatcher(NORM_HAMMING,true).match, lib.rs:116-126 */
int dunk_match_crosscheck_hamming(dunk_ctx* ctx, const uint8_t* query, int nq,
                                  const uint8_t* train, int64_t nt, int desc_bytes,
                                  DunkDMatch* out, int out_cap, int* n_out);

/* ---- stage 2, float descriptors (north_star only; the reference builds no NORM_L2 matcher) ------
 * replaces cv::BFMatcher(NORM_L2).knnMatch(query, train, 2) for f32 descriptors of dim = 64 or 128
 * floats.  ||q-t||^2 = ||q||^2 + ||t||^2 - 2 q.t: the dot products run on tcgen05 tensor cores
 * (kind::tf32, TMA-fed, TMEM accumulators) and only NOMINATE 4 candidates per (query, slab); the
 * returned neighbours are re-ranked with exact f32 distances and proven complete against the tf32
 * error bound, queries that cannot be proven fall back to an exact scan.  idx: nq x 2 int32
 * (ties -> lower train index), dist: nq x 2 f32 (sqrt of the exact squared distance).
 * stats (may be NULL): {queries re-done by the exact fallback, slabs used}.  nt < 2 -> -211. */
int dunk_knn2_l2(dunk_ctx* ctx, const float* query, int nq, const float* train, int64_t nt, int dim,
                 int32_t* idx, float* dist, int* stats);
/* device-resident variant (16-byte aligned inputs; async on the slot's stream except for two
 * 4-byte read-backs) */
int dunk_knn2_l2_dev(dunk_ctx* ctx, int slot, const void* q_dev, int nq, const void* t_dev, int64_t nt,
                     int dim, void* idx_dev, void* dist_dev, int* stats);

/* ---- HBM-resident reference descriptor database (feature_database read side) ----
 * rows = models::Keypoint (feature_database/src/models.rs:30-41): SoA arrays in HBM,
 * descriptors padded to 64-B rows.  One dunk_db is one shard (one GPU). */
int dunk_db_create(dunk_ctx* ctx, int64_t capacity_rows, int desc_bytes, dunk_db** out);
void dunk_db_destroy(dunk_db* db);
/* append n rows; kps / image_ids may be NULL (descriptor-only DB) */
int dunk_db_append(dunk_db* db, const uint8_t* desc, const DunkKeyPoint* kps,
                   const int32_t* image_ids, int64_t n);
/* fill n rows with device-generated uniform random descriptors (seeded; bench config 3) */
int dunk_db_append_random(dunk_db* db, int64_t n, uint64_t seed);
/* same, for a shard: row r of the call is row (global_row_offset + r) of the seeded global sequence,
 * so the union of the shards equals the unsharded DB whatever the shard count */
int dunk_db_append_random_at(dunk_db* db, int64_t n, uint64_t seed, uint64_t global_row_offset);
int64_t dunk_db_size(dunk_db* db);
int dunk_db_desc_bytes(dunk_db* db);   /* descriptor width of the shard (a loaded dump carries its own) */
/* forget every keypoint and ref_image row (capacity and device buffers are kept): TRUNCATE keypoint, ref_image */
int dunk_db_clear(dunk_db* db);
/* read back rows [first, first+n) (any of the outputs may be NULL) */
int dunk_db_read(dunk_db* db, int64_t first, int64_t n, uint8_t* desc, DunkKeyPoint* kps,
                 int32_t* image_ids);
/* 2-NN + ratio of host queries against the shard; train_idx = index_base + local row */
int dunk_db_match(dunk_db* db, const uint8_t* query, int nq, float ratio, DunkDMatch* out,
                  int out_cap, int* n_out);
/* local top-2 of host queries against the shard (for an external merge), out: nq records */
int dunk_db_knn2(dunk_db* db, const uint8_t* query, int nq, uint32_t index_base, DunkTop2* out);

/* ---- feature_database: ref_image table, keyed reads, flat dump / load ------
 * models::Image (feature_database/src/models.rs:5-15): the tile a keypoint row belongs to. */
typedef struct DunkImage {
    int32_t id;                 /* 1-based, assigned by dunk_db_create_image (Postgres SERIAL) */
    int32_t x_start, y_start, x_end, y_end;
    int32_t level_of_detail;
} DunkImage;
/* ImageDatabase::create_image (imagedb.rs:14-30): returns the new id through *id_out */
int dunk_db_create_image(dunk_db* db, int32_t x_start, int32_t y_start, int32_t x_end, int32_t y_end,
                         int32_t level_of_detail, int32_t* id_out);
/* ImageDatabase::read_image_from_id (imagedb.rs:32-37); unknown id -> DUNK_ERR_OUT_OF_RANGE */
int dunk_db_read_image(dunk_db* db, int32_t id, DunkImage* out);
/* ImageDatabase::find_images_from_dimensions / find_images_from_lod (imagedb.rs:39-66): ids of the
 * images of `level_of_detail` whose extent intersects the box (use_box = 0: every image of the LoD) */
int dunk_db_find_images(dunk_db* db, int use_box, int32_t x_start, int32_t y_start, int32_t x_end,
                        int32_t y_end, int32_t level_of_detail, int32_t* ids, int cap, int* n_out);
int dunk_db_image_count(dunk_db* db);
/* KeypointDatabase::read_keypoints_from_{image_id, lod, coordinates} (keypointdb.rs:38-90):
 * SELECT rows WHERE [image_id = f.image_id] [AND ref_image.level_of_detail = f.level_of_detail]
 * [AND x_coord >= floor(x_start) AND x_coord <= ceil(x_end) AND y likewise]
 * ORDER BY response DESC LIMIT `limit` (the reference passes 2^18 - 1).  Runs on the device
 * (predicate, stable descending radix sort on the response, gather) and returns the result as a new
 * HBM-resident dunk_db (read it with dunk_db_read / dunk_db_read_ids, or match against it).
 * A negative image_id / level_of_detail disables that predicate. */
typedef struct DunkRowFilter {
    int32_t image_id;
    int32_t level_of_detail;
    int32_t use_box;
    float x_start, y_start, x_end, y_end;
} DunkRowFilter;
int dunk_db_select(dunk_db* db, const DunkRowFilter* filter, int64_t limit, dunk_db** out);
/* the `id` column of rows [first, first+n): 1 + row index in the DB the rows were inserted into */
int dunk_db_read_ids(dunk_db* db, int64_t first, int64_t n, int32_t* ids);
/* flat binary dump / load of a shard (header, ref_image table, SoA columns as they lie in HBM), so a
 * multi-GB reference DB is built once (dunk_db_append_tiles) and re-loaded at H2D speed.
 * load: capacity = max(rows in file, min_capacity_rows). */
int dunk_db_save(dunk_db* db, const char* path);
int dunk_db_load(dunk_ctx* ctx, const char* path, int64_t min_capacity_rows, dunk_db** out);

/* ---- device-pointer variants (async on slot's stream; no sync, no host copies) ---- */
/* queries: nq x 64-B padded rows in device memory; top2_dev: nq DunkTop2 records */
int dunk_db_knn2_dev(dunk_db* db, int slot, const void* query64_dev, int nq, uint32_t index_base,
                     void* top2_dev);
/* merge `n_parts` gathered DunkTop2 arrays (part-major, nq each) lexicographically by
 * (distance,index) -> merged_dev (nq records).  SURVEY 8(e): the step after ncclAllGather */
int dunk_top2_merge_dev(dunk_ctx* ctx, int slot, const void* parts_dev, int n_parts, int nq,
                        void* merged_dev);
/* ratio filter + ordered compaction of merged top-2 -> DunkDMatch list on device;
 * count_dev: one int32 */
int dunk_top2_ratio_dev(dunk_ctx* ctx, int slot, const void* merged_dev, int nq, float ratio,
                        void* matches_dev, void* count_dev);
/* pad n x desc_bytes rows to n x 64-B rows on device */
int dunk_pad_desc_dev(dunk_ctx* ctx, int slot, const void* src_dev, int64_t n, int desc_bytes,
                      void* dst64_dev);

/* ---- stage 1: AKAZE keypoints + MLDB-486 descriptors -------------------------------------
 * replaces AKAZE::create(DESCRIPTOR_MLDB, 0, 3, 0.001, 4, 4, DIFF_PM_G2, max_points)
 * .detectAndCompute(img, no mask) — akaze_keypoint_descriptor_extraction_def,
 * feature_extraction/src/lib.rs:61-92.  image: rows x cols, 1 (gray) / 3 (BGR) / 4 (BGRA)
 * interleaved u8 channels.  Outputs: kps (cv::KeyPoint layout, OpenCV's order: evolution level
 * ascending, row-major inside a level) and desc (n x 61 u8); *n_out keypoints written.
 * max_points <= 0: unlimited; otherwise the max_points strongest responses are kept.
 * Errors: empty / unsupported image -> DUNK_ERR_ASSERT (-215); capacity exceeded -> DUNK_ERR_NO_MEM. */
int dunk_akaze_extract(dunk_ctx* ctx, const uint8_t* image, int rows, int cols, int channels,
                       int row_stride_bytes, int max_points, DunkKeyPoint* kps, uint8_t* desc,
                       int cap, int* n_out);
/* frame batch (same shape): frame f at images + f*frame_stride_bytes (0 = tightly packed);
 * outputs strided by cap_per_frame; counts: n_frames int32 */
int dunk_akaze_extract_batch(dunk_ctx* ctx, const uint8_t* images, int n_frames, int rows, int cols,
                             int channels, int row_stride_bytes, size_t frame_stride_bytes,
                             int max_points, DunkKeyPoint* kps, uint8_t* desc, int cap_per_frame,
                             int* counts);
/* per-stage parity hook: f32 planes (level_w x level_h) of one evolution level + k-contrast */
int dunk_akaze_debug_level(dunk_ctx* ctx, const uint8_t* image, int rows, int cols, int channels,
                           int row_stride_bytes, int level, float* Lt, float* Lx, float* Ly,
                           float* Ldet, float* kcontrast, int* level_w, int* level_h, int* n_levels);

/* ---- stage 3: RANSAC homography ---------------------------------------------------------
 * replaces cv::findHomography(src, dst, mask, method, thr) as called by find_homography_mat,
 * homographier/src/homographier/mod.rs:231-259 (5-arg overload: maxIters 2000, conf 0.995).
 * src/dst: n x 2 f32 (Point2f).  H: 9 f64 row-major, H[8] = 1.  mask: n u8 (may be NULL).
 * method: DUNK_H_RANSAC (8), DUNK_H_LMEDS (4: 55 iterations of the same sample stream, least median of the
 * f32 errors, sigma-inliers refitted; thr only shapes the returned mask) or DUNK_H_DEFAULT (0, least squares on
 * all pairs); DUNK_H_RHO (16) is served by the RANSAC estimator (OpenCV's PROSAC + SPRT schedule is not restated:
 * same contract, tolerance parity only).  n < 4 -> DUNK_ERR_VEC_LENGTH (-28, OpenCV's StsVecLengthErr).
 * *found = 0 when no model was found (OpenCV returns an empty Mat -> MatError::Empty). */
int dunk_find_homography(dunk_ctx* ctx, const float* src, const float* dst, int n, int method,
                         double thr, double* H, uint8_t* mask, int* found);
/* batch of independent problems (one CTA each): points concatenated, offsets[n_problems+1];
 * H: n_problems x 9; mask: offsets[n_problems] bytes; info: n_problems x 4 int32 =
 * {found, inliers, RANSAC iterations run, hypotheses scored} */
int dunk_find_homography_batch(dunk_ctx* ctx, const float* src, const float* dst,
                               const int* offsets, int n_problems, int method, double thr,
                               double* H, uint8_t* mask, int* info);
/* parity hook (north_star: "identical seeded hypothesis sets giving identical inlier counts"):
 * scores explicit 4-index samples; counts: n_hyp int32 (-1 = degenerate), Hs: n_hyp x 9 f64 */
int dunk_ransac_score_hypotheses(dunk_ctx* ctx, const float* src, const float* dst, int n,
                                 const int* samples, int n_hyp, double thr, int* counts,
                                 double* Hs);

/* ---- stage 3b: PnP-RANSAC pose ------------------------------------------------------------
 * replaces cv::solvePnPRansac(obj, img, K, dist = zeros(4,1), rvec, tvec, false, iters, thr, conf,
 * inliers, method) as called by pnp_solver_ransac, homographier/src/homographier/mod.rs:320-369
 * (distortion is forced to zero at :344; method defaults to SOLVEPNP_EPNP at :359).
 * obj: n x 3 f64 (Point3d), img: n x 2 f64 (Point2d), K: 9 f64 row-major camera matrix.  As in
 * OpenCV the f64 points are rounded to f32 first.  rvec/tvec: 3 f64 each.  inliers: indices of the
 * inliers of the best minimal model in increasing order (capacity inliers_cap; may be NULL).
 * *found = 0 when no pose was found (the reference returns Ok(None), mod.rs:367).
 * n < 4 -> DUNK_ERR_ASSERT (-215, reference test mod.rs:627-638).  Methods: DUNK_PNP_EPNP (5-point
 * samples, EPnP kernel) and DUNK_PNP_P3P (4-point samples, P3P kernel; also used, as in OpenCV,
 * whenever n == 4); the final pose is EPnP over the inliers in both cases.  DUNK_PNP_ITERATIVE: the EPnP
 * RANSAC stage, then the Levenberg-Marquardt minimum of the reprojection error over the inliers (cv2 agrees
 * with the minimiser to ~1e-8).  Other methods -> DUNK_ERR_BAD_ARG. */
int dunk_pnp_ransac(dunk_ctx* ctx, const double* obj, const double* img, int n, const double* K,
                    int iters, float thr, double confidence, int method, double* rvec, double* tvec,
                    int32_t* inliers, int inliers_cap, int* n_inliers, int* found);
/* batch of independent problems (one CTA each; frames partition with no collective): points
 * concatenated, offsets[n_problems+1], K: n_problems x 9, rvecs/tvecs: n_problems x 3,
 * inlier_mask: offsets[n_problems] bytes (may be NULL), info: n_problems x 4 int32 =
 * {found, inliers, RANSAC iterations run, hypotheses scored} */
int dunk_pnp_ransac_batch(dunk_ctx* ctx, const double* obj, const double* img, const int* offsets,
                          int n_problems, const double* K, int iters, float thr, double confidence,
                          int method, double* rvecs, double* tvecs, uint8_t* inlier_mask, int* info);
/* parity hook ("identical seeded hypothesis sets giving identical inlier counts"): scores explicit
 * 5-index samples with the EPnP kernel; counts: n_hyp int32 (-1 = no finite pose), rt: n_hyp x 6
 * f64 (rvec, tvec of every hypothesis) */
int dunk_pnp_score_hypotheses(dunk_ctx* ctx, const double* obj, const double* img, int n,
                              const double* K, const int* samples, int n_hyp, double thr,
                              int* counts, double* rt);

/* ---- warp_image_perspective -----------------------------------------------
 * replaces cv::warpPerspective(src, dst, M, size, INTER_LINEAR, BORDER_CONSTANT, Scalar(1,1,1,1)) as
 * called by warp_image_perspective, homographier/src/homographier/mod.rs:271-300, for 8-bit images
 * with 1..4 interleaved channels; bit-exact with OpenCV 4.13 (1/32-pixel grid, 15-bit weights).
 * M: 9 f64 row-major, src -> dst (inverted internally like OpenCV without WARP_INVERSE_MAP).
 * border_value: 4 f64 (NULL = the reference's 1,1,1,1).  dst: out_rows x out_cols x channels. */
int dunk_warp_perspective(dunk_ctx* ctx, const uint8_t* src, int rows, int cols, int channels,
                          int row_stride_bytes, const double* M, int out_rows, int out_cols,
                          const double* border_value, uint8_t* dst);
/* device-resident batch: n outputs of the same source, one matrix each (M: n x 9, host), written
 * back to back at dst_dev; async on the slot's stream */
int dunk_warp_perspective_batch_dev(dunk_ctx* ctx, int slot, const void* src_dev, int rows, int cols,
                                    int channels, int row_stride_bytes, const double* M, int n,
                                    int out_rows, int out_cols, const double* border_value,
                                    void* dst_dev);

/* ---- the steps either side of the path ------------------------------------
 * band_merger + f32_to_u8 + gamma_correction, geotiff_extractor/src/image_extractor/mod.rs:346-378,
 * 402-422: three f32 bands (n samples each) -> n RGBA8 pixels.  Per channel: NaN -> 0;
 * f = (v - min) / (max - min) in f32; f outside 0..=1 -> 0; else round(powf(f, 1/2.2) * 255).
 * Alpha = 0 where all three bands are NaN, else 255.  min_max: {red_min, red_max, green_min,
 * green_max, blue_min, blue_max} (f64, cast to f32 as the reference does).  bgra != 0 writes the
 * channels in the BGRA order raster_to_mat produces (homographier mod.rs:183-220), so the output
 * feeds dunk_akaze_extract (channels = 4) directly. */
int dunk_band_merger(dunk_ctx* ctx, const float* red, const float* green, const float* blue, int64_t n,
                     const double* min_max, int bgra, uint8_t* out_rgba);
int dunk_band_merger_dev(dunk_ctx* ctx, int slot, const void* red_dev, const void* green_dev,
                         const void* blue_dev, int64_t n, const double* min_max, int bgra, void* out_dev);
/* raster_to_mat, homographier/src/homographier/mod.rs:183-220: w*h RGBA8 -> BGRA8 (Cmat<Vec4b>) */
int dunk_raster_to_mat(dunk_ctx* ctx, const uint8_t* rgba, int w, int h, uint8_t* bgra);
/* geotransform::get_world_coordinates, feature_database/src/elevationdb.rs:64-104: reference-image
 * pixel (x, y) -> GDAL geotransform "dataset" -> (lon, lat) -> nearest sample of the elevation raster
 * through the inverted "elevation" geotransform (row id = round(y) * x_size + round(x) + 1) ->
 * EPSG:4326 -> EPSG:4978 (WGS-84 geodetic -> ECEF metres): the object points of pnp_solver_ransac.
 * gt_elevation / heights may be NULL: height 0, as the reference falls back to (:76-79).
 * heights: y_size x x_size f64 (the `elevation` table in row-id order), kept in HBM by the handle. */
int dunk_elevation_create(dunk_ctx* ctx, const double* gt_dataset, const double* gt_elevation,
                          const double* heights, int x_size, int y_size, dunk_elevation** out);
void dunk_elevation_destroy(dunk_elevation* e);
/* xyz: n x 3 f64.  Points whose elevation sample does not exist (diesel NotFound in the reference)
 * get NaN coordinates and are counted in *n_missing (may be NULL). */
int dunk_world_coordinates(dunk_elevation* e, const double* px, const double* py, int64_t n,
                           double* xyz, int* n_missing);

/* ---- the whole path: frame batch -> extract -> 2-NN + ratio vs the shard -> RANSAC homography
 * The composition the reference performs in feature_extraction/src/lib.rs:196-249 followed by
 * find_homography_mat (mod.rs:231-259), for a batch of same-shape frames, without leaving the
 * device between the stages.  H maps query-frame pixels to reference (scene) pixels. */
typedef struct DunkRegistration {
    double H[9];        /* row-major, H[8] = 1; zeros when !found */
    int32_t found;      /* 1 = a homography was found */
    int32_t inliers;    /* inliers of the final H */
    int32_t matches;    /* ratio-test survivors fed to RANSAC */
    int32_t keypoints;  /* keypoints extracted from the frame */
    int32_t ransac_iters, hypotheses;
} DunkRegistration;
int dunk_register_frames(dunk_db* db, const uint8_t* images, int n_frames, int rows, int cols,
                         int channels, int row_stride_bytes, size_t frame_stride_bytes, float ratio,
                         double thr, int max_points, DunkRegistration* results);
/* device-resident variant (async on the slot's stream; results_dev: n_frames records) */
size_t dunk_register_workspace_bytes(dunk_db* db, int n_frames, int rows, int cols);
int dunk_register_frames_dev(dunk_db* db, int slot, const void* images_dev, int n_frames, int rows,
                             int cols, int channels, int row_stride_bytes, size_t frame_stride_bytes,
                             float ratio, double thr, int max_points, void* workspace_dev,
                             size_t workspace_bytes, void* results_dev);
/* ---- ... continued to the attitude (north_star stage 3: "RANSAC homography/PnP scoring that yields the attitude
 * estimate").  The correspondences the homography kept go through geotransform::get_world_coordinates
 * (feature_database/src/elevationdb.rs:64-104: reference pixel -> geotransform -> elevation sample -> ECEF) and
 * pnp_solver_ransac (homographier/src/homographier/mod.rs:320-369), all on the device.  The caller's `origin` is
 * subtracted from the ECEF coordinates in f64 before the f32 rounding OpenCV applies to solvePnPRansac inputs
 * (|ECEF| = 6.4e6 m would otherwise quantise object points to 0.5 m); tvec is relative to that origin. */
typedef struct DunkPoseConfig {
    const dunk_elevation* elevation; /* geotransform + elevation table of the reference scene (dunk_elevation_create) */
    double K[9];                     /* camera matrix, row-major */
    double origin[3];                /* ECEF metres subtracted from every object point (zeros: raw ECEF as the reference) */
    int32_t method;                  /* DunkPnPMethod */
    int32_t iters;                   /* iter_count (mod.rs:323) */
    float thr;                       /* reproj_thres */
    double confidence;
} DunkPoseConfig;
typedef struct DunkPose {
    double rvec[3], tvec[3];         /* zeros when !found */
    int32_t found, inliers, ransac_iters, hypotheses;
} DunkPose;
/* pose == NULL: identical to dunk_register_frames(_dev); otherwise poses (n_frames records) is filled too */
int dunk_register_frames_pose(dunk_db* db, const uint8_t* images, int n_frames, int rows, int cols, int channels,
                              int row_stride_bytes, size_t frame_stride_bytes, float ratio, double thr, int max_points,
                              const DunkPoseConfig* pose, DunkRegistration* results, DunkPose* poses);
int dunk_register_frames_pose_dev(dunk_db* db, int slot, const void* images_dev, int n_frames, int rows, int cols,
                                  int channels, int row_stride_bytes, size_t frame_stride_bytes, float ratio,
                                  double thr, int max_points, const DunkPoseConfig* pose, void* workspace_dev,
                                  size_t workspace_bytes, void* results_dev, void* poses_dev);

/* ---- the same path with the DB sharded over GPUs (one process per GPU) ----------
 * phase 1 (frame owner): extract + pack the batch's descriptors as 64-B query rows;
 * phase 2 (every shard): dunk_db_knn2_dev of all ranks' query rows against the local shard;
 * phase 3 (frame owner): (distance, index)-merge of the shards' top-2, ratio, RANSAC.
 * The two exchanges between the phases (queries out, top-2 back) are the caller's collectives. */
typedef struct DunkPipelineView {
    void* query64_dev;          /* total_queries x 64-B rows (capacity: query_capacity rows) */
    void* query_offsets_dev;    /* n_frames + 1 int32: first query of every frame */
    void* keypoints_dev;        /* n_frames x keypoint_capacity DunkKeyPoint */
    void* keypoint_counts_dev;  /* n_frames int32 */
    void* top2_dev;             /* scratch: query_capacity DunkTop2 records */
    int32_t total_queries;
    int32_t keypoint_capacity;
    int64_t query_capacity;
} DunkPipelineView;
size_t dunk_pipeline_workspace_bytes(dunk_ctx* ctx, int n_frames, int rows, int cols);
int dunk_pipeline_extract_dev(dunk_ctx* ctx, int slot, const void* images_dev, int n_frames, int rows,
                              int cols, int channels, int row_stride_bytes, size_t frame_stride_bytes,
                              int max_points, void* workspace_dev, size_t workspace_bytes,
                              DunkPipelineView* view);
/* parts_dev: n_parts arrays of DunkTop2 (part p at parts_dev + p*part_stride_records), global row
 * indices; db_keypoints_dev: keypoints of ALL shards addressed by (index - index_base) */
int dunk_pipeline_finish_dev(dunk_ctx* ctx, int slot, int n_frames, int rows, int cols,
                             const void* parts_dev, int n_parts, int64_t part_stride_records,
                             int total_queries, const void* db_keypoints_dev, uint32_t index_base,
                             float ratio, double thr, void* workspace_dev, size_t workspace_bytes,
                             void* results_dev);
/* ---- the shard group: the exchange step inside the library (NCCL over NVLink / NVSwitch) ---------------
 * One process per GPU; rank 0 calls dunk_shard_unique_id and hands the 128 bytes to the other ranks out of band
 * (MPI / a file / any rendezvous the host application has); every rank then calls dunk_shard_group_create, which is
 * ncclCommInitRank.  world == 1 needs no id and no NCCL.  All calls below are COLLECTIVE: every rank of the group
 * makes the same call with the same sizes. */
#define DUNK_SHARD_ID_BYTES 128
typedef struct dunk_shard_group dunk_shard_group;
int dunk_shard_unique_id(uint8_t* id128);
int dunk_shard_group_create(dunk_ctx* ctx, int rank, int world, const uint8_t* id128, dunk_shard_group** out);
void dunk_shard_group_destroy(dunk_shard_group* g);
int dunk_shard_group_rank(dunk_shard_group* g);
int dunk_shard_group_world(dunk_shard_group* g);
int dunk_nccl_version(void);   /* e.g. 22809; 0 when libnccl.so.2 cannot be loaded */
/* Re-cut the rows the ranks built locally (dunk_db_append_tiles / dunk_db_build_from_bands of each rank's share of the
 * tiles; global order = rank-major) into `world` equal contiguous row ranges and replicate the 28-byte keypoint column
 * on every rank (global row -> reference point).  *shard_out: this rank's shard, rows [base(rank), base(rank + 1)). */
int dunk_shard_group_balance(dunk_shard_group* g, dunk_db* built, dunk_db** shard_out);
int64_t dunk_shard_group_total_rows(dunk_shard_group* g);
int64_t dunk_shard_group_base(dunk_shard_group* g, int rank);   /* rank in 0..world (world: total rows) */
/* get_knn_matches (feature_extraction/src/lib.rs:94-114) against the sharded DB (BASELINE config 3): local top-2 of
 * the replicated queries on every shard, one ncclAllGather of the 16-byte records, (distance, index) merge, ratio test
 * after the merge -> identical to the unsharded result incl. ties.  index_base = global row of the shard's row 0.
 * _dev: nq x 64-B query rows in device memory; outputs (any may be NULL): merged top-2 records, DunkDMatch list + count. */
int dunk_db_match_sharded(dunk_shard_group* g, dunk_db* shard, const uint8_t* query, int nq, uint32_t index_base,
                          float ratio, DunkDMatch* out, int out_cap, int* n_out);
int dunk_db_match_sharded_dev(dunk_shard_group* g, dunk_db* shard, int slot, const void* query64_dev, int nq,
                              uint32_t index_base, float ratio, void* top2_merged_dev, void* matches_dev,
                              void* count_dev);
/* The whole registration step with the DB sharded (BASELINE config 5), ONE call per step and rank: every rank extracts
 * its own frame batch; the ranks exchange their query counts (one 16-byte ncclAllGather + the step's only host sync,
 * which the single-GPU path has too), all-gather the query rows padded to the largest count of THIS step, match all
 * ranks' queries against the local shard in one launch, return the 16-byte top-2 records to the frame owners (grouped
 * ncclSend / ncclRecv), merge by (distance, index), ratio test, RANSAC homography and, with `pose`, PnP — the last
 * three partitioned by frame with no collective.  The group must hold the keypoint column (dunk_shard_group_balance). */
size_t dunk_register_sharded_workspace_bytes(dunk_shard_group* g, int n_frames, int rows, int cols);
int dunk_register_frames_sharded_dev(dunk_shard_group* g, dunk_db* shard, int slot, const void* images_dev, int n_frames,
                                     int rows, int cols, int channels, int row_stride_bytes, size_t frame_stride_bytes,
                                     float ratio, double thr, int max_points, const DunkPoseConfig* pose,
                                     void* workspace_dev, size_t workspace_bytes, void* results_dev, void* poses_dev);
/* async copies on a slot's stream for callers that drive the _dev entry points from pinned host buffers */
int dunk_memcpy_h2d(dunk_ctx* ctx, int slot, void* dst_dev, const void* src_host, size_t nbytes);
int dunk_memcpy_d2h(dunk_ctx* ctx, int slot, void* dst_host, const void* src_dev, size_t nbytes);
/* device / pinned-host buffers owned by the caller (so a host application needs no CUDA binding of its own) */
int dunk_dev_alloc(dunk_ctx* ctx, size_t nbytes, void** out_dev);
int dunk_dev_free(dunk_ctx* ctx, void* dev);
int dunk_host_alloc(dunk_ctx* ctx, size_t nbytes, void** out_host);
int dunk_host_free(dunk_ctx* ctx, void* host);

/* append n rows whose columns already live on the device (64-B descriptor rows, DunkKeyPoint,
 * int32 image ids; the last two may be NULL) — used to re-cut shards into equal row ranges */
int dunk_db_append_dev(dunk_db* db, int slot, const void* desc64_dev, const void* kps_dev,
                       const void* image_ids_dev, int64_t n);
const void* dunk_db_keypoints_dev(dunk_db* db);   /* shard columns, for replication / exchange */
const void* dunk_db_descriptors_dev(dunk_db* db);
/* device-to-device copy, async on the slot's stream (keeps exchanges ordered with the kernels) */
int dunk_memcpy_dev(dunk_ctx* ctx, int slot, void* dst_dev, const void* src_dev, size_t nbytes);

/* reference-DB build (preprocessor/src/main.rs:248-327 minus GDAL/Postgres): extract a tile
 * batch and append rows to the shard; keypoint coordinates become x*scale[t] + x_off[t]
 * (main.rs:296-304); image_ids[t] -> image_id column.  counts (may be NULL): rows per tile. */
int dunk_db_append_tiles(dunk_db* db, const uint8_t* images, int n_tiles, int rows, int cols,
                         int channels, int row_stride_bytes, size_t frame_stride_bytes,
                         const float* x_off, const float* y_off, const float* scale,
                         const int32_t* image_ids, int max_points, int* counts);

/* reference-DB build straight from the scene's three f32 bands (SURVEY 8f rank 2;
 * preprocessor/src/main.rs:160-327 minus GDAL / Postgres): tile = scene >> (lods - 1); for every lod in
 * 0..lods the (tile << lod)-sized windows are resampled to tile size on the device (resample 0 = box
 * mean, 1 = Lanczos-3 stretched by the decimation factor, the kernel GDAL's RasterIO uses; the
 * reference's exact pixels additionally depend on GDAL's overview selection and cannot be pinned
 * here), converted by band_merger (min_max as in dunk_band_merger), extracted (BGRA), and appended
 * with x * 2^lod + window offset (main.rs:300-301); one ref_image row per tile (main.rs:283-289).
 * Outputs (may be NULL): number of tiles processed and the tile size. */
int dunk_db_build_from_bands(dunk_db* db, const float* red, const float* green, const float* blue,
                             int width, int height, const double* min_max, int lods, int resample,
                             int max_points, int* n_tiles_out, int* tile_w_out, int* tile_h_out);
/* same with the three bands already resident in HBM (width*height f32 each, device pointers) */
int dunk_db_build_from_bands_dev(dunk_db* db, const void* red_dev, const void* green_dev, const void* blue_dev,
                                 int width, int height, const double* min_max, int lods, int resample,
                                 int max_points, int* n_tiles_out, int* tile_w_out, int* tile_h_out);

/* one part of a build partitioned over ranks (tiles are independent, preprocessor/src/main.rs:258-277: no halo, no
 * collective): part p of n_parts extracts the contiguous share [jobs*p/n, jobs*(p+1)/n) of the lod-major tile walk and
 * inserts ALL ref_image rows, so image ids agree across ranks; follow with dunk_shard_group_balance */
int dunk_db_build_from_bands_part_dev(dunk_db* db, const void* red_dev, const void* green_dev, const void* blue_dev,
                                      int width, int height, const double* min_max, int lods, int resample,
                                      int max_points, int part, int n_parts, int* n_tiles_out, int* tile_w_out,
                                      int* tile_h_out);

/* ---- roofline denominators measured on the box ---------------------------------------- */
/* per-kernel-class device times: between begin and end every launch site brackets its kernels
 * with CUDA events on the launching stream.  end() returns the number of classes n and fills
 * names ("a;b;...;"), ms / launches / alg (algorithmic bytes, or pairs for the matcher) [n]. */
int dunk_profile_begin(dunk_ctx* ctx);
int dunk_profile_end(dunk_ctx* ctx, char* names, int names_cap, double* ms, int* launches,
                     double* alg, int cap);
/* f32_to_u8's gamma step is evaluated through a 255-entry threshold table built from the direct formula;
 * this compares the two on every f32 in [0, 1] (2^30 values, ~1 s): *mismatches must come back 0 */
int dunk_selftest_gamma_lut(dunk_ctx* ctx, uint64_t* mismatches);
/* POPC-pipe peak in 1e12 popc/s (best of 4 timed launches of `iters` x 32 popc per thread) */
int dunk_microbench_popc(dunk_ctx* ctx, int iters, double* tpopc_per_s);

#ifdef __cplusplus
}
#endif
#endif /* DUNK_B200_H */
