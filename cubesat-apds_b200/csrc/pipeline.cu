// End-to-end registration of a frame batch against the HBM-resident reference database:
// extract (stage 1) -> 2-NN + ratio (stage 2) -> RANSAC homography (stage 3), everything on the
// device, one stream, no host round trips between the stages.  This is the composition the
// reference only performs in a test (feature_extraction/src/lib.rs:196-249 extract -> knn ->
// points) followed by find_homography_mat (homographier/src/homographier/mod.rs:231-259).
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include "akaze.h"
#include "gamma_lut.cuh"
#include "geo.cuh"
#include "match.h"
#include "pipeline.h"
#include "shard.h"

namespace dunk {

// from akaze_api.cu
int akaze_run(dunk_ctx* ctx, cudaStream_t st, const LevelTable& lt, const AkazeWorkspace& ws,
              const unsigned char* images_dev, size_t image_stride, int row_stride, int channels, int frames,
              int max_points);

namespace {

// exclusive scan of per-frame keypoint counts -> query offsets (frames <= 1024), one CTA
// offsets[frames + 1] = the largest raw-extrema count of the batch: k_extrema keeps counting past the
// candidate capacity and drops the entry, so the caller compares it with cand_cap in the read-back it
// already makes and fails loudly instead of returning an atomics-order-dependent keypoint subset
__global__ void __launch_bounds__(1024) k_frame_offsets(const int* __restrict__ counts, int frames, int* __restrict__ offsets,
                                                        const int* __restrict__ cand_count) {
    __shared__ int sh[1024];
    __shared__ int cmax;
    const int t = threadIdx.x;
    if (t == 0) cmax = 0;
    sh[t] = t < frames ? counts[t] : 0;
    __syncthreads();
    if (t < frames) atomicMax(&cmax, cand_count[t]);
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = t >= o ? sh[t - o] : 0;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    if (t < frames) offsets[t + 1] = sh[t];
    if (t == 0) {
        offsets[0] = 0;
        offsets[frames + 1] = cmax;
    }
}

// gather every frame's descriptor rows into one contiguous query array
__global__ void __launch_bounds__(256)
k_pack_queries(const uint4* __restrict__ desc64, int kp_cap, const int* __restrict__ counts, const int* __restrict__ offsets,
               uint4* __restrict__ q64) {
    const int f = blockIdx.y;
    const int n = counts[f];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // uint4 index inside the frame
    if (i >= n * 4) return;
    q64[(size_t)offsets[f] * 4 + i] = desc64[(size_t)f * kp_cap * 4 + i];
}

// per frame: Lowe ratio test + ordered compaction of (query point, reference point) pairs
// (get_knn_matches lib.rs:107-111 + get_points_from_matches lib.rs:161-180, intended semantics)
__global__ void __launch_bounds__(1024)
k_frame_pairs(const uint4* __restrict__ top2, const int* __restrict__ counts, const int* __restrict__ offsets, float ratio,
              const DunkKeyPoint* __restrict__ kps, int kp_cap, const DunkKeyPoint* __restrict__ db_kps, uint32_t index_base,
              float2* __restrict__ src, float2* __restrict__ dst, DunkDMatch* __restrict__ matches, int* __restrict__ n_pairs) {
    __shared__ int wsum[32];
    __shared__ int running, chunk_total;
    const int f = blockIdx.x;
    const int n = counts[f], off = offsets[f];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int qi = base + tid;
        uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (qi < n) v = top2[off + qi];
        const bool keep = qi < n && v.z != ~0u && ((float)v.x < __fmul_rn((float)v.z, ratio));
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            const int x = wsum[lane];
            int incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            wsum[lane] = incl - x;
            if (lane == 31) chunk_total = incl;
        }
        __syncthreads();
        if (keep) {
            const int pos = off + running + wsum[warp] + __popc(bal & ((1u << lane) - 1u));
            const DunkKeyPoint q = kps[(size_t)f * kp_cap + qi];
            const DunkKeyPoint r = db_kps[v.y - index_base];
            src[pos] = make_float2(q.x, q.y);
            dst[pos] = make_float2(r.x, r.y);
            matches[pos] = DunkDMatch{qi, (int)v.y, 0, (float)v.x};
        }
        __syncthreads();
        if (tid == 0) running += chunk_total;
        __syncthreads();
    }
    if (tid == 0) n_pairs[f] = running;
}

struct RegResult {   // DunkRegistration in the header
    double H[9];
    int found, inliers, matches, keypoints, ransac_iters, hypotheses;
};

__global__ void k_pack_results(const double* __restrict__ H, const int* __restrict__ info, const int* __restrict__ n_pairs,
                               const int* __restrict__ counts, int frames, DunkRegistration* __restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    DunkRegistration r;
    for (int i = 0; i < 9; ++i) r.H[i] = H[f * 9 + i];
    r.found = info[f * 4 + 0];
    r.inliers = info[f * 4 + 1];
    r.ransac_iters = info[f * 4 + 2];
    r.hypotheses = info[f * 4 + 3];
    r.matches = n_pairs[f];
    r.keypoints = counts[f];
    out[f] = r;
}

// ---- pose stage (stage 3b): homography inliers -> (object point, image point) pairs -> PnP-RANSAC -------------
// What a caller of the reference composes by hand: get_world_coordinates (feature_database/src/elevationdb.rs:64-104)
// for every matched reference keypoint, then pnp_solver_ransac (homographier/src/homographier/mod.rs:320-369).
struct PoseParams {
    GeoParams geo;
    const double* heights;
    double K[9];
    double origin[3];
};

// per frame: ordered compaction of the pairs the homography kept (mask != 0) into the f32 point lists
// solvePnPRansac works on (OpenCV rounds its f64 inputs to f32 first; the caller's origin is subtracted in f64
// before that rounding so ECEF magnitudes of 6.4e6 m do not eat the mantissa).  Pairs whose elevation sample is
// missing are dropped.  Block 0 also places the camera matrix in device memory for the PnP kernel.
__global__ void __launch_bounds__(1024)
k_pose_inputs(const float2* __restrict__ src, const float2* __restrict__ dst, const uint8_t* __restrict__ mask,
              const int* __restrict__ offsets, const int* __restrict__ n_pairs, const int* __restrict__ h_info, PoseParams pp,
              float* __restrict__ obj, float* __restrict__ img, int* __restrict__ pose_count, double* __restrict__ K_dev) {
    __shared__ int wsum[32];
    __shared__ int running, chunk_total;
    const int f = blockIdx.x;
    const int off = offsets[f];
    const int n = h_info[f * 4 + 0] ? n_pairs[f] : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (f == 0 && tid < 9) K_dev[tid] = pp.K[tid];
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        bool keep = i < n && mask[off + i] != 0;
        double w[3] = {0, 0, 0};
        float2 q = make_float2(0.f, 0.f);
        if (keep) {
            const float2 r = dst[off + i];
            q = src[off + i];
            keep = world_point(pp.geo, pp.heights, (double)r.x, (double)r.y, w);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            const int x = wsum[lane];
            int incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            wsum[lane] = incl - x;
            if (lane == 31) chunk_total = incl;
        }
        __syncthreads();
        if (keep) {
            const size_t pos = (size_t)off + running + wsum[warp] + __popc(bal & ((1u << lane) - 1u));
            obj[pos * 3 + 0] = (float)(w[0] - pp.origin[0]);
            obj[pos * 3 + 1] = (float)(w[1] - pp.origin[1]);
            obj[pos * 3 + 2] = (float)(w[2] - pp.origin[2]);
            img[pos * 2 + 0] = q.x;
            img[pos * 2 + 1] = q.y;
        }
        __syncthreads();
        if (tid == 0) running += chunk_total;
        __syncthreads();
    }
    if (tid == 0) pose_count[f] = running;
}

__global__ void k_pack_poses(const double* __restrict__ rt, const int* __restrict__ info, int frames, DunkPose* __restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    DunkPose p;
    for (int i = 0; i < 3; ++i) {
        p.rvec[i] = rt[f * 6 + i];
        p.tvec[i] = rt[f * 6 + 3 + i];
    }
    p.found = info[f * 4 + 0];
    p.inliers = info[f * 4 + 1];
    p.ransac_iters = info[f * 4 + 2];
    p.hypotheses = info[f * 4 + 3];
    out[f] = p;
}

size_t al(size_t b) { return (b + 255) & ~size_t(255); }

struct PipelineBuffers {
    AkazeWorkspace ws;
    int* q_off;          // [frames + 1]
    uint4* q64;          // [frames * kp_cap] rows
    uint4* top2;         // [frames * kp_cap]
    uint4* partial;      // knn slab partials
    float2 *src, *dst;   // [frames * kp_cap]
    DunkDMatch* matches; // [frames * kp_cap]
    int* n_pairs;        // [frames]
    double* H;           // [frames * 9]
    uint8_t* mask;       // [frames * kp_cap]
    int* info;           // [frames * 4]
    DunkRegistration* results;  // [frames]
    // pose stage
    float* p_obj;        // [frames * kp_cap][3]
    float* p_img;        // [frames * kp_cap][2]
    int* p_count;        // [frames]
    double* p_K;         // [9]
    double* p_rt;        // [frames][6]
    uint8_t* p_mask;     // [frames * kp_cap]
    int* p_info;         // [frames][4]
    DunkPose* poses;     // [frames]
};

}  // namespace

struct PipelinePlan {
    LevelTable lt;
    int frames, cand_cap, kp_cap;
    size_t ws_bytes, total_bytes, partial_bytes;
};

static PipelinePlan plan_pipeline(dunk_ctx* ctx, int rows, int cols, int frames, int64_t db_rows) {
    PipelinePlan p;
    p.lt = make_level_table(cols, rows);
    p.frames = frames;
    long long c = (long long)cols * rows / 32;
    c = std::max<long long>(c, 2048);
    c = std::min<long long>(c, 1 << 20);
    p.cand_cap = p.kp_cap = (int)c;
    p.ws_bytes = akaze_workspace_bytes(p.lt, frames, p.cand_cap, p.kp_cap);
    const size_t nq_max = (size_t)frames * p.kp_cap;
    // worst-case slab count for the matcher: plan with the largest query count
    (void)db_rows;
    p.partial_bytes = knn2_partial_bound(ctx, nq_max);
    p.total_bytes = al(p.ws_bytes) + al((frames + 2) * 4) + al(nq_max * 64) + al(nq_max * 16) + al(p.partial_bytes) +
                    2 * al(nq_max * 8) + al(nq_max * 16) + al(frames * 4) + al((size_t)frames * 72) + al(nq_max) +
                    al((size_t)frames * 16) + al((size_t)frames * sizeof(DunkRegistration)) +
                    al(nq_max * 12) + al(nq_max * 8) + al(frames * 4) + al(72) + al((size_t)frames * 48) + al(nq_max) +
                    al((size_t)frames * 16) + al((size_t)frames * sizeof(DunkPose));
    return p;
}

static void carve_pipeline(void* base, const PipelinePlan& p, PipelineBuffers* b) {
    char* ptr = (char*)base;
    auto take = [&](size_t bytes) { void* r = ptr; ptr += al(bytes); return r; };
    akaze_carve_workspace(take(p.ws_bytes), p.lt, p.frames, p.cand_cap, p.kp_cap, &b->ws);
    const size_t nq_max = (size_t)p.frames * p.kp_cap;
    b->q_off = (int*)take((p.frames + 2) * 4);
    b->q64 = (uint4*)take(nq_max * 64);
    b->top2 = (uint4*)take(nq_max * 16);
    b->partial = (uint4*)take(p.partial_bytes);
    b->src = (float2*)take(nq_max * 8);
    b->dst = (float2*)take(nq_max * 8);
    b->matches = (DunkDMatch*)take(nq_max * 16);
    b->n_pairs = (int*)take(p.frames * 4);
    b->H = (double*)take((size_t)p.frames * 72);
    b->mask = (uint8_t*)take(nq_max);
    b->info = (int*)take((size_t)p.frames * 16);
    b->results = (DunkRegistration*)take((size_t)p.frames * sizeof(DunkRegistration));
    b->p_obj = (float*)take(nq_max * 12);
    b->p_img = (float*)take(nq_max * 8);
    b->p_count = (int*)take(p.frames * 4);
    b->p_K = (double*)take(72);
    b->p_rt = (double*)take((size_t)p.frames * 48);
    b->p_mask = (uint8_t*)take(nq_max);
    b->p_info = (int*)take((size_t)p.frames * 16);
    b->poses = (DunkPose*)take((size_t)p.frames * sizeof(DunkPose));
}

// stage 1 for `frames` device-resident images + packing of every frame's descriptors into one
// query array; returns the total query count (one 4-byte read back: the matcher grid depends on it)
static int pipeline_extract_async(dunk_ctx* ctx, cudaStream_t st, const PipelinePlan& p, const PipelineBuffers& b,
                                  const unsigned char* images_dev, size_t frame_stride, int row_stride, int channels, int frames,
                                  int max_points) {
    int rc = akaze_run(ctx, st, p.lt, b.ws, images_dev, frame_stride, row_stride, channels, frames, max_points);
    if (rc) return rc;
    k_frame_offsets<<<1, 1024, 0, st>>>(b.ws.kp_count, frames, b.q_off, b.ws.cand_count);
    DUNK_KERNEL_CHECK(ctx);
    k_pack_queries<<<dim3(div_up((long long)p.kp_cap * 4, 256), frames), 256, 0, st>>>(b.ws.desc64, p.kp_cap, b.ws.kp_count,
                                                                                      b.q_off, b.q64);
    DUNK_KERNEL_CHECK(ctx);
    return DUNK_OK;
}

static int pipeline_extract(dunk_ctx* ctx, cudaStream_t st, const PipelinePlan& p, const PipelineBuffers& b,
                            const unsigned char* images_dev, size_t frame_stride, int row_stride, int channels, int frames,
                            int max_points, int* total_q) {
    int rc = pipeline_extract_async(ctx, st, p, b, images_dev, frame_stride, row_stride, channels, frames, max_points);
    if (rc) return rc;
    int h_tail[2] = {0, 0};     // {total queries, largest raw-extrema count}
    DUNK_CUDA(cudaMemcpyAsync(h_tail, b.q_off + frames, 8, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    *total_q = h_tail[0];
    DUNK_REQUIRE(h_tail[1] <= p.cand_cap, DUNK_ERR_NO_MEM,
                 "pipeline: a frame produced %d raw extrema, candidate capacity %d (w*h/32)", h_tail[1], p.cand_cap);
    return DUNK_OK;
}

// stage 2 tail + stage 3: per-frame ratio test -> point pairs -> RANSAC homography -> result records.
// top2: merged top-2 per query; db_kps: keypoints addressed by (train index - index_base)
static int pipeline_finish(dunk_ctx* ctx, cudaStream_t st, const PipelinePlan& p, const PipelineBuffers& b,
                           const uint4* top2, const DunkKeyPoint* db_kps, uint32_t index_base, int frames, float ratio,
                           float thr, const DunkPoseConfig* pose = nullptr) {
    {
        ProfScope ps(ctx, st, "pipe.frame_pairs", 0.0);
        k_frame_pairs<<<frames, 1024, 0, st>>>(top2, b.ws.kp_count, b.q_off, ratio, b.ws.kps, p.kp_cap, db_kps, index_base, b.src,
                                               b.dst, b.matches, b.n_pairs);
        DUNK_KERNEL_CHECK(ctx);
    }
    int rc;
    {
        ProfScope ps(ctx, st, "ransac.find_homography", 0.0);
        if ((rc = launch_find_homography(ctx, st, b.src, b.dst, b.q_off, b.n_pairs, frames, thr, b.H, b.mask, b.info))) return rc;
    }
    {
        ProfScope ps(ctx, st, "pipe.pack_results", 0.0);
        k_pack_results<<<div_up(frames, 128), 128, 0, st>>>(b.H, b.info, b.n_pairs, b.ws.kp_count, frames, b.results);
        DUNK_KERNEL_CHECK(ctx);
    }
    if (!pose) return DUNK_OK;
    // stage 3b: the correspondences the homography kept -> ECEF object points -> PnP-RANSAC pose per frame
    PoseParams pp;
    pp.geo = make_geo_params(pose->elevation);
    pp.heights = pose->elevation->heights;
    for (int i = 0; i < 9; ++i) pp.K[i] = pose->K[i];
    for (int i = 0; i < 3; ++i) pp.origin[i] = pose->origin[i];
    {
        ProfScope ps(ctx, st, "pose.inputs", 0.0);
        k_pose_inputs<<<frames, 1024, 0, st>>>(b.src, b.dst, b.mask, b.q_off, b.n_pairs, b.info, pp, b.p_obj, b.p_img, b.p_count, b.p_K);
        DUNK_KERNEL_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, st, "ransac.pnp", 0.0);
        if ((rc = launch_pnp_ransac(ctx, st, b.p_obj, b.p_img, b.q_off, b.p_count, frames, b.p_K, 0, pose->method, pose->iters,
                                    pose->thr, pose->confidence, b.p_rt, b.p_mask, b.p_info)))
            return rc;
    }
    {
        ProfScope ps(ctx, st, "pipe.pack_results", 0.0);
        k_pack_poses<<<div_up(frames, 128), 128, 0, st>>>(b.p_rt, b.p_info, frames, b.poses);
        DUNK_KERNEL_CHECK(ctx);
    }
    return DUNK_OK;
}

static int check_pose(const char* fn, const DunkPoseConfig* pose) {
    if (!pose) return DUNK_OK;
    DUNK_REQUIRE(pose->elevation, DUNK_ERR_BAD_ARG, "%s: pose config without an elevation / geotransform handle", fn);
    DUNK_REQUIRE(pose->method == DUNK_PNP_EPNP || pose->method == DUNK_PNP_P3P || pose->method == DUNK_PNP_ITERATIVE, DUNK_ERR_BAD_ARG,
                 "%s: PnP method %d not implemented", fn, pose->method);
    return DUNK_OK;
}

// all three stages for `frames` device-resident images against one shard; results (device) in b.results
static int run_pipeline(dunk_ctx* ctx, cudaStream_t st, dunk_db* db, const PipelinePlan& p, const PipelineBuffers& b,
                        const unsigned char* images_dev, size_t frame_stride, int row_stride, int channels, int frames,
                        float ratio, float thr, int max_points, const DunkPoseConfig* pose) {
    int total_q = 0;
    int rc = pipeline_extract(ctx, st, p, b, images_dev, frame_stride, row_stride, channels, frames, max_points, &total_q);
    if (rc) return rc;
    if (total_q > 0 && db->size >= 2) {
        const KnnPlan kp = plan_knn2(ctx, total_q, (uint32_t)db->size);
        if ((size_t)kp.gx * total_q * 16 > p.partial_bytes) {
            set_error("pipeline: matcher partial buffer too small");
            return DUNK_ERR_NO_MEM;
        }
        if ((rc = launch_knn2(ctx, st, db->desc64, (uint32_t)db->size, b.q64, total_q, 0, b.partial, b.top2, kp))) return rc;
    } else if (total_q > 0) {
        DUNK_CUDA(cudaMemsetAsync(b.top2, 0xFF, (size_t)total_q * 16, st));
    }
    return pipeline_finish(ctx, st, p, b, b.top2, db->kps, 0, frames, ratio, thr, pose);
}

}  // namespace dunk

using namespace dunk;

extern "C" {

static int check_frames(const char* fn, int n_frames, int rows, int cols, int channels, int row_stride) {
    DUNK_REQUIRE(n_frames >= 0 && n_frames <= 1024, DUNK_ERR_BAD_ARG, "%s: n_frames=%d (0..1024 per call)", fn, n_frames);
    DUNK_REQUIRE(rows >= 16 && cols >= 16, DUNK_ERR_ASSERT, "%s: image %dx%d too small", fn, cols, rows);
    DUNK_REQUIRE(channels == 1 || channels == 3 || channels == 4, DUNK_ERR_ASSERT, "%s: %d channels", fn, channels);
    DUNK_REQUIRE(row_stride >= cols * channels, DUNK_ERR_BAD_ARG, "%s: row stride < row bytes", fn);
    return DUNK_OK;
}

size_t dunk_register_workspace_bytes(dunk_db* db, int n_frames, int rows, int cols) {
    if (!db || n_frames <= 0) return 0;
    return plan_pipeline(db->ctx, rows, cols, n_frames, db->size).total_bytes;
}

int dunk_register_frames_pose_dev(dunk_db* db, int slot, const void* images_dev, int n_frames, int rows, int cols, int channels,
                                  int row_stride_bytes, size_t frame_stride_bytes, float ratio, double thr, int max_points,
                                  const DunkPoseConfig* pose, void* workspace_dev, size_t workspace_bytes, void* results_dev,
                                  void* poses_dev) {
    DUNK_REQUIRE(db && images_dev && workspace_dev && results_dev, DUNK_ERR_BAD_ARG, "dunk_register_frames_dev: NULL argument");
    DUNK_REQUIRE(!pose || poses_dev, DUNK_ERR_BAD_ARG, "dunk_register_frames_pose_dev: pose config without a pose output");
    dunk_ctx* ctx = db->ctx;
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_register_frames_dev: bad slot");
    int rc = check_frames("dunk_register_frames_dev", n_frames, rows, cols, channels, row_stride_bytes);
    if (rc || (rc = check_pose("dunk_register_frames_pose_dev", pose))) return rc;
    if (n_frames == 0) return DUNK_OK;
    if (frame_stride_bytes == 0) frame_stride_bytes = (size_t)rows * row_stride_bytes;
    DUNK_CUDA(cudaSetDevice(ctx->device));
    const PipelinePlan p = plan_pipeline(ctx, rows, cols, n_frames, db->size);
    DUNK_REQUIRE(workspace_bytes >= p.total_bytes, DUNK_ERR_NO_MEM, "dunk_register_frames_dev: workspace %zu < %zu bytes",
                 workspace_bytes, p.total_bytes);
    PipelineBuffers b;
    carve_pipeline(workspace_dev, p, &b);
    cudaStream_t st = ctx->slots[slot].stream;
    if ((rc = run_pipeline(ctx, st, db, p, b, (const unsigned char*)images_dev, frame_stride_bytes, row_stride_bytes, channels,
                           n_frames, ratio, (float)thr, max_points <= 0 ? 0 : max_points, pose)))
        return rc;
    DUNK_CUDA(cudaMemcpyAsync(results_dev, b.results, (size_t)n_frames * sizeof(DunkRegistration), cudaMemcpyDeviceToDevice, st));
    if (pose) DUNK_CUDA(cudaMemcpyAsync(poses_dev, b.poses, (size_t)n_frames * sizeof(DunkPose), cudaMemcpyDeviceToDevice, st));
    return DUNK_OK;
}

int dunk_register_frames_dev(dunk_db* db, int slot, const void* images_dev, int n_frames, int rows, int cols, int channels,
                             int row_stride_bytes, size_t frame_stride_bytes, float ratio, double thr, int max_points,
                             void* workspace_dev, size_t workspace_bytes, void* results_dev) {
    return dunk_register_frames_pose_dev(db, slot, images_dev, n_frames, rows, cols, channels, row_stride_bytes, frame_stride_bytes,
                                         ratio, thr, max_points, nullptr, workspace_dev, workspace_bytes, results_dev, nullptr);
}

/* Host buffers in, host records out.  The frames move through two pinned staging halves filled by the calling
 * thread while the previous sub-batch is on the device, so pageable callers (a Rust Vec<u8>) get an async
 * H2D copy at pinned speed that overlaps the kernels instead of the driver's staged synchronous copy. */
int dunk_register_frames_pose(dunk_db* db, const uint8_t* images, int n_frames, int rows, int cols, int channels,
                              int row_stride_bytes, size_t frame_stride_bytes, float ratio, double thr, int max_points,
                              const DunkPoseConfig* pose, DunkRegistration* results, DunkPose* poses) {
    DUNK_REQUIRE(db && results, DUNK_ERR_BAD_ARG, "dunk_register_frames: NULL argument");
    DUNK_REQUIRE(!pose || poses, DUNK_ERR_BAD_ARG, "dunk_register_frames_pose: pose config without a pose output");
    dunk_ctx* ctx = db->ctx;
    int rc = check_frames("dunk_register_frames", n_frames, rows, cols, channels, row_stride_bytes);
    if (rc || (rc = check_pose("dunk_register_frames_pose", pose))) return rc;
    if (n_frames == 0) return DUNK_OK;
    DUNK_REQUIRE(images, DUNK_ERR_ASSERT, "dunk_register_frames: empty image");
    if (frame_stride_bytes == 0) frame_stride_bytes = (size_t)rows * row_stride_bytes;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    // sub-batches of <= 64 frames keep the workspace bounded
    const int sub = std::min(n_frames, 64);
    const PipelinePlan p = plan_pipeline(ctx, rows, cols, sub, db->size);
    const size_t img_bytes = al((size_t)sub * frame_stride_bytes);
    void* scratch = ctx->dev_scratch(g.s, al(p.total_bytes) + 2 * img_bytes);
    if (!scratch) return DUNK_ERR_NO_MEM;
    const size_t res_bytes = al((size_t)n_frames * sizeof(DunkRegistration));
    unsigned char* pin = (unsigned char*)ctx->pin_scratch(g.s, 2 * img_bytes + res_bytes + al((size_t)n_frames * sizeof(DunkPose)));
    if (!pin) return DUNK_ERR_NO_MEM;
    DunkRegistration* pin_res = (DunkRegistration*)(pin + 2 * img_bytes);     // pageable outputs would make the D2H copy block
    DunkPose* pin_pose = (DunkPose*)(pin + 2 * img_bytes + res_bytes);
    PipelineBuffers b;
    carve_pipeline(scratch, p, &b);
    unsigned char* d_img = (unsigned char*)scratch + al(p.total_bytes);
    cudaEvent_t staged[2] = {g.slot().ev0, g.slot().ev1};   // H2D of half h has left the pinned half
    int half = 0;
    for (int f0 = 0; f0 < n_frames; f0 += sub, half ^= 1) {
        const int nf = std::min(sub, n_frames - f0);
        if (f0 >= 2 * sub) DUNK_CUDA(cudaEventSynchronize(staged[half]));
        memcpy(pin + half * img_bytes, images + (size_t)f0 * frame_stride_bytes, (size_t)nf * frame_stride_bytes);
        DUNK_CUDA(cudaMemcpyAsync(d_img + half * img_bytes, pin + half * img_bytes, (size_t)nf * frame_stride_bytes,
                                  cudaMemcpyHostToDevice, st));
        DUNK_CUDA(cudaEventRecord(staged[half], st));
        if ((rc = run_pipeline(ctx, st, db, p, b, d_img + half * img_bytes, frame_stride_bytes, row_stride_bytes, channels, nf, ratio,
                               (float)thr, max_points <= 0 ? 0 : max_points, pose)))
            return rc;
        DUNK_CUDA(cudaMemcpyAsync(pin_res + f0, b.results, (size_t)nf * sizeof(DunkRegistration), cudaMemcpyDeviceToHost, st));
        if (pose) DUNK_CUDA(cudaMemcpyAsync(pin_pose + f0, b.poses, (size_t)nf * sizeof(DunkPose), cudaMemcpyDeviceToHost, st));
    }
    DUNK_CUDA(cudaStreamSynchronize(st));
    memcpy(results, pin_res, (size_t)n_frames * sizeof(DunkRegistration));
    if (pose) memcpy(poses, pin_pose, (size_t)n_frames * sizeof(DunkPose));
    return DUNK_OK;
}

int dunk_register_frames(dunk_db* db, const uint8_t* images, int n_frames, int rows, int cols, int channels,
                         int row_stride_bytes, size_t frame_stride_bytes, float ratio, double thr, int max_points,
                         DunkRegistration* results) {
    return dunk_register_frames_pose(db, images, n_frames, rows, cols, channels, row_stride_bytes, frame_stride_bytes, ratio, thr,
                                     max_points, nullptr, results, nullptr);
}

/* ---- sharded pipeline phases: extract on the frame owner, match on every shard,
 * merge + RANSAC on the frame owner.  The exchange between the phases is the caller's (NCCL). ---- */
size_t dunk_pipeline_workspace_bytes(dunk_ctx* ctx, int n_frames, int rows, int cols) {
    if (!ctx || n_frames <= 0) return 0;
    return plan_pipeline(ctx, rows, cols, n_frames, 1).total_bytes;
}

int dunk_pipeline_extract_dev(dunk_ctx* ctx, int slot, const void* images_dev, int n_frames, int rows, int cols, int channels,
                              int row_stride_bytes, size_t frame_stride_bytes, int max_points, void* workspace_dev,
                              size_t workspace_bytes, DunkPipelineView* view) {
    DUNK_REQUIRE(ctx && images_dev && workspace_dev && view, DUNK_ERR_BAD_ARG, "dunk_pipeline_extract_dev: NULL argument");
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_pipeline_extract_dev: bad slot");
    int rc = check_frames("dunk_pipeline_extract_dev", n_frames, rows, cols, channels, row_stride_bytes);
    if (rc) return rc;
    DUNK_REQUIRE(n_frames > 0, DUNK_ERR_BAD_ARG, "dunk_pipeline_extract_dev: no frames");
    if (frame_stride_bytes == 0) frame_stride_bytes = (size_t)rows * row_stride_bytes;
    DUNK_CUDA(cudaSetDevice(ctx->device));
    const PipelinePlan p = plan_pipeline(ctx, rows, cols, n_frames, 1);
    DUNK_REQUIRE(workspace_bytes >= p.total_bytes, DUNK_ERR_NO_MEM, "dunk_pipeline_extract_dev: workspace %zu < %zu bytes",
                 workspace_bytes, p.total_bytes);
    PipelineBuffers b;
    carve_pipeline(workspace_dev, p, &b);
    int total_q = 0;
    if ((rc = pipeline_extract(ctx, ctx->slots[slot].stream, p, b, (const unsigned char*)images_dev, frame_stride_bytes,
                               row_stride_bytes, channels, n_frames, max_points <= 0 ? 0 : max_points, &total_q)))
        return rc;
    view->query64_dev = b.q64;
    view->query_offsets_dev = b.q_off;
    view->keypoints_dev = b.ws.kps;
    view->keypoint_counts_dev = b.ws.kp_count;
    view->top2_dev = b.top2;
    view->total_queries = total_q;
    view->keypoint_capacity = p.kp_cap;
    view->query_capacity = (int64_t)n_frames * p.kp_cap;
    return DUNK_OK;
}

int dunk_pipeline_finish_dev(dunk_ctx* ctx, int slot, int n_frames, int rows, int cols, const void* parts_dev, int n_parts,
                             int64_t part_stride_records, int total_queries, const void* db_keypoints_dev,
                             uint32_t index_base, float ratio, double thr, void* workspace_dev, size_t workspace_bytes,
                             void* results_dev) {
    DUNK_REQUIRE(ctx && parts_dev && db_keypoints_dev && workspace_dev && results_dev && n_parts >= 1, DUNK_ERR_BAD_ARG,
                 "dunk_pipeline_finish_dev: bad argument");
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_pipeline_finish_dev: bad slot");
    DUNK_REQUIRE(n_frames > 0 && n_frames <= 1024 && total_queries >= 0, DUNK_ERR_BAD_ARG, "dunk_pipeline_finish_dev: bad sizes");
    DUNK_CUDA(cudaSetDevice(ctx->device));
    const PipelinePlan p = plan_pipeline(ctx, rows, cols, n_frames, 1);
    DUNK_REQUIRE(workspace_bytes >= p.total_bytes, DUNK_ERR_NO_MEM, "dunk_pipeline_finish_dev: workspace too small");
    PipelineBuffers b;
    carve_pipeline(workspace_dev, p, &b);
    cudaStream_t st = ctx->slots[slot].stream;
    const uint4* top2 = (const uint4*)parts_dev;
    int rc;
    if (total_queries > 0 && (n_parts > 1 || part_stride_records != total_queries)) {
        // lexicographic (distance, index) merge of the shards' top-2 records; parts are part-major with
        // a stride, so merge each part row range through the strided view
        {
            ProfScope ps(ctx, st, "match.top2_merge", (double)total_queries * n_parts * 16);
            if ((rc = launch_top2_merge_strided(ctx, st, (const uint4*)parts_dev, n_parts, part_stride_records, total_queries, b.top2)))
                return rc;
        }
        top2 = b.top2;
    }
    if ((rc = pipeline_finish(ctx, st, p, b, top2, (const DunkKeyPoint*)db_keypoints_dev, index_base, n_frames, ratio, (float)thr)))
        return rc;
    DUNK_CUDA(cudaMemcpyAsync(results_dev, b.results, (size_t)n_frames * sizeof(DunkRegistration), cudaMemcpyDeviceToDevice, st));
    return DUNK_OK;
}

/* ---- the sharded registration step, one call per step and rank (BASELINE config 5) ---------------- */
size_t dunk_register_sharded_workspace_bytes(dunk_shard_group* g, int n_frames, int rows, int cols) {
    if (!g || n_frames <= 0) return 0;
    return plan_pipeline(g->ctx, rows, cols, n_frames, 1).total_bytes;
}

int dunk_register_frames_sharded_dev(dunk_shard_group* g, dunk_db* shard, int slot, const void* images_dev, int n_frames, int rows,
                                     int cols, int channels, int row_stride_bytes, size_t frame_stride_bytes, float ratio, double thr,
                                     int max_points, const DunkPoseConfig* pose, void* workspace_dev, size_t workspace_bytes,
                                     void* results_dev, void* poses_dev) {
    DUNK_REQUIRE(g && shard && images_dev && workspace_dev && results_dev, DUNK_ERR_BAD_ARG, "dunk_register_frames_sharded_dev: NULL argument");
    DUNK_REQUIRE(!pose || poses_dev, DUNK_ERR_BAD_ARG, "dunk_register_frames_sharded_dev: pose config without a pose output");
    dunk_ctx* ctx = g->ctx;
    DUNK_REQUIRE(shard->ctx == ctx, DUNK_ERR_BAD_ARG, "dunk_register_frames_sharded_dev: the shard lives on another context");
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_register_frames_sharded_dev: bad slot");
    DUNK_REQUIRE(g->kps_all && g->total_rows == g->bases[g->world], DUNK_ERR_BAD_ARG,
                 "dunk_register_frames_sharded_dev: the group holds no keypoint column (dunk_shard_group_balance first)");
    int rc = check_frames("dunk_register_frames_sharded_dev", n_frames, rows, cols, channels, row_stride_bytes);
    if (rc || (rc = check_pose("dunk_register_frames_sharded_dev", pose))) return rc;
    DUNK_REQUIRE(n_frames > 0, DUNK_ERR_BAD_ARG, "dunk_register_frames_sharded_dev: no frames (the call is collective)");
    if (frame_stride_bytes == 0) frame_stride_bytes = (size_t)rows * row_stride_bytes;
    DUNK_CUDA(cudaSetDevice(ctx->device));
    const PipelinePlan p = plan_pipeline(ctx, rows, cols, n_frames, 1);
    DUNK_REQUIRE(workspace_bytes >= p.total_bytes, DUNK_ERR_NO_MEM, "dunk_register_frames_sharded_dev: workspace %zu < %zu bytes",
                 workspace_bytes, p.total_bytes);
    PipelineBuffers b;
    carve_pipeline(workspace_dev, p, &b);
    cudaStream_t st = ctx->slots[slot].stream;
    const int W = g->world, me = g->rank;
    // phase 1 (frame owner): extract + pack; then the ranks learn each other's {query count, raw-extrema maximum}
    if ((rc = pipeline_extract_async(ctx, st, p, b, (const unsigned char*)images_dev, frame_stride_bytes, row_stride_bytes, channels,
                                     n_frames, max_points <= 0 ? 0 : max_points)))
        return rc;
    int* d_cnt = (int*)g->d_counts;          // [W][2] int32
    int* h_cnt = (int*)g->h_counts;
    {
        ProfScope ps(ctx, st, "shard.all_gather_counts", 8.0 * W);
        if ((rc = shard_all_gather(g, b.q_off + n_frames, d_cnt, 8, st))) return rc;
    }
    DUNK_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));     // the step's one host sync (grids and message sizes depend on the counts)
    int qmax = 0, cand_max = 0;
    for (int r = 0; r < W; ++r) {
        qmax = std::max(qmax, h_cnt[2 * r]);
        cand_max = std::max(cand_max, h_cnt[2 * r + 1]);
    }
    DUNK_REQUIRE(cand_max <= p.cand_cap, DUNK_ERR_NO_MEM, "pipeline: a frame produced %d raw extrema, candidate capacity %d (w*h/32)",
                 cand_max, p.cand_cap);
    const int my_q = h_cnt[2 * me];
    const uint4* top2 = b.top2;
    if (qmax > 0) {
        // phase 2 (every shard): all ranks' query rows, padded to the largest count of this step (rows past a rank's
        // count are stale rows of its workspace: matched and ignored), against the local shard in ONE launch
        const size_t all_q = (size_t)W * qmax;
        uint4* q_all = W == 1 ? b.q64 : (uint4*)g->ensure(0, all_q * 64, st);
        uint4* t_out = W == 1 ? b.top2 : (uint4*)g->ensure(1, all_q * 16, st);
        uint4* parts = W == 1 ? b.top2 : (uint4*)g->ensure(2, all_q * 16, st);
        if (!q_all || !t_out || !parts) return DUNK_ERR_NO_MEM;
        if (W > 1) {
            ProfScope ps(ctx, st, "shard.all_gather_queries", (double)all_q * 64);
            if ((rc = shard_all_gather(g, b.q64, q_all, (size_t)qmax * 64, st))) return rc;
        }
        if (shard->size > 0) {
            const KnnPlan kp = plan_knn2(ctx, (int)all_q, (uint32_t)shard->size);
            uint4* partial = W == 1 ? b.partial : (uint4*)g->ensure(3, knn2_partial_bytes(kp, (int)all_q), st);
            if (!partial) return DUNK_ERR_NO_MEM;
            if (W == 1 && knn2_partial_bytes(kp, (int)all_q) > p.partial_bytes) {
                set_error("pipeline: matcher partial buffer too small");
                return DUNK_ERR_NO_MEM;
            }
            if ((rc = launch_knn2(ctx, st, shard->desc64, (uint32_t)shard->size, q_all, (int)all_q, (uint32_t)g->bases[me], partial, t_out, kp,
                                  W > 1 ? d_cnt : nullptr, qmax)))
                return rc;
        } else {
            DUNK_CUDA(cudaMemsetAsync(t_out, 0xFF, all_q * 16, st));
        }
        if (W > 1) {
            // phase 3 (frame owner): every shard's top-2 of MY queries, merged by (distance, index)
            {
                ProfScope ps(ctx, st, "shard.all_to_all_top2", (double)all_q * 16);
                if ((rc = shard_all_to_all(g, t_out, parts, (size_t)qmax * 16, st))) return rc;
            }
            if (my_q > 0) {
                ProfScope ps(ctx, st, "match.top2_merge", (double)my_q * W * 16);
                if ((rc = launch_top2_merge_strided(ctx, st, parts, W, qmax, my_q, b.top2))) return rc;
            }
        }
    }
    if ((rc = pipeline_finish(ctx, st, p, b, top2, g->kps_all, 0, n_frames, ratio, (float)thr, pose))) return rc;
    DUNK_CUDA(cudaMemcpyAsync(results_dev, b.results, (size_t)n_frames * sizeof(DunkRegistration), cudaMemcpyDeviceToDevice, st));
    if (pose) DUNK_CUDA(cudaMemcpyAsync(poses_dev, b.poses, (size_t)n_frames * sizeof(DunkPose), cudaMemcpyDeviceToDevice, st));
    return DUNK_OK;
}

int dunk_memcpy_dev(dunk_ctx* ctx, int slot, void* dst_dev, const void* src_dev, size_t nbytes) {
    DUNK_REQUIRE(ctx && slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_memcpy_dev: bad ctx / slot");
    if (nbytes == 0) return DUNK_OK;
    DUNK_REQUIRE(dst_dev && src_dev, DUNK_ERR_BAD_ARG, "dunk_memcpy_dev: NULL pointer");
    DUNK_CUDA(cudaSetDevice(ctx->device));
    DUNK_CUDA(cudaMemcpyAsync(dst_dev, src_dev, nbytes, cudaMemcpyDeviceToDevice, ctx->slots[slot].stream));
    return DUNK_OK;
}

int dunk_db_append_dev(dunk_db* db, int slot, const void* desc64_dev, const void* kps_dev, const void* image_ids_dev,
                       int64_t n) {
    DUNK_REQUIRE(db && n >= 0, DUNK_ERR_BAD_ARG, "dunk_db_append_dev: bad argument");
    if (n == 0) return DUNK_OK;
    DUNK_REQUIRE(desc64_dev, DUNK_ERR_BAD_ARG, "dunk_db_append_dev: NULL descriptors");
    dunk_ctx* ctx = db->ctx;
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_db_append_dev: bad slot");
    std::lock_guard<std::mutex> lk(db->mu);
    DUNK_REQUIRE(db->size + n <= db->capacity, DUNK_ERR_NO_MEM, "dunk_db_append_dev: exceeds capacity %lld",
                 (long long)db->capacity);
    DUNK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[slot].stream;
    DUNK_CUDA(cudaMemcpyAsync(db->desc64 + db->size * 4, desc64_dev, (size_t)n * 64, cudaMemcpyDeviceToDevice, st));
    if (kps_dev) DUNK_CUDA(cudaMemcpyAsync(db->kps + db->size, kps_dev, (size_t)n * sizeof(DunkKeyPoint), cudaMemcpyDeviceToDevice, st));
    else DUNK_CUDA(cudaMemsetAsync(db->kps + db->size, 0, (size_t)n * sizeof(DunkKeyPoint), st));
    if (image_ids_dev) DUNK_CUDA(cudaMemcpyAsync(db->image_id + db->size, image_ids_dev, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    else DUNK_CUDA(cudaMemsetAsync(db->image_id + db->size, 0, (size_t)n * 4, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    db->size += n;
    return DUNK_OK;
}

const void* dunk_db_keypoints_dev(dunk_db* db) { return db ? db->kps : nullptr; }
const void* dunk_db_descriptors_dev(dunk_db* db) { return db ? db->desc64 : nullptr; }

static __global__ void __launch_bounds__(256)
k_append_rows(const uint4* __restrict__ desc64, const DunkKeyPoint* __restrict__ kps, int kp_cap, const int* __restrict__ counts,
              const int* __restrict__ offsets, const float* __restrict__ x_off, const float* __restrict__ y_off,
              const float* __restrict__ scale, const int32_t* __restrict__ image_ids, int tile0, uint4* __restrict__ db_desc,
              DunkKeyPoint* __restrict__ db_kps, int32_t* __restrict__ db_ids, long long db_base) {
    const int f = blockIdx.y;
    const int n = counts[f];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long row = db_base + offsets[f] + i;
#pragma unroll
    for (int k = 0; k < 4; ++k) db_desc[row * 4 + k] = desc64[((size_t)f * kp_cap + i) * 4 + k];
    DunkKeyPoint kp = kps[(size_t)f * kp_cap + i];
    const float s = scale ? scale[tile0 + f] : 1.f;
    // preprocessor/src/main.rs:300-301: x * 2^lod + column offset
    kp.x = kp.x * s + (x_off ? x_off[tile0 + f] : 0.f);
    kp.y = kp.y * s + (y_off ? y_off[tile0 + f] : 0.f);
    db_kps[row] = kp;
    db_ids[row] = image_ids ? image_ids[tile0 + f] : tile0 + f;
}

/* extract a tile batch and append keypoints + descriptors to the shard, mapping keypoint
 * coordinates into scene pixels: x*scale + x_off (preprocessor/src/main.rs:296-304) */
int dunk_db_append_tiles(dunk_db* db, const uint8_t* images, int n_tiles, int rows, int cols, int channels,
                         int row_stride_bytes, size_t frame_stride_bytes, const float* x_off, const float* y_off,
                         const float* scale, const int32_t* image_ids, int max_points, int* counts) {
    DUNK_REQUIRE(db, DUNK_ERR_BAD_ARG, "dunk_db_append_tiles: db is NULL");
    dunk_ctx* ctx = db->ctx;
    DUNK_REQUIRE(n_tiles >= 0, DUNK_ERR_BAD_ARG, "dunk_db_append_tiles: n_tiles < 0");
    if (n_tiles == 0) return DUNK_OK;
    int rc = check_frames("dunk_db_append_tiles", std::min(n_tiles, 1024), rows, cols, channels, row_stride_bytes);
    if (rc) return rc;
    DUNK_REQUIRE(images, DUNK_ERR_ASSERT, "dunk_db_append_tiles: empty image");
    if (frame_stride_bytes == 0) frame_stride_bytes = (size_t)rows * row_stride_bytes;
    std::lock_guard<std::mutex> lk(db->mu);
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const int sub = std::min(n_tiles, 32);
    const LevelTable lt = make_level_table(cols, rows);
    long long c = (long long)cols * rows / 32;
    c = std::min<long long>(std::max<long long>(c, 2048), 1 << 20);
    const int cap = (int)c;
    const size_t ws_bytes = akaze_workspace_bytes(lt, sub, cap, cap);
    const size_t meta = al((size_t)n_tiles * 4);
    const size_t need = al(ws_bytes) + al((size_t)sub * frame_stride_bytes) + al((sub + 2) * 4) + 4 * meta;
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    char* ptr = (char*)scratch;
    AkazeWorkspace ws;
    akaze_carve_workspace(ptr, lt, sub, cap, cap, &ws);
    ptr += al(ws_bytes);
    unsigned char* d_img = (unsigned char*)ptr; ptr += al((size_t)sub * frame_stride_bytes);
    int* d_off = (int*)ptr; ptr += al((sub + 2) * 4);
    float* d_xo = (float*)ptr; ptr += meta;
    float* d_yo = (float*)ptr; ptr += meta;
    float* d_sc = (float*)ptr; ptr += meta;
    int32_t* d_id = (int32_t*)ptr;
    if (x_off) DUNK_CUDA(cudaMemcpyAsync(d_xo, x_off, (size_t)n_tiles * 4, cudaMemcpyHostToDevice, st));
    if (y_off) DUNK_CUDA(cudaMemcpyAsync(d_yo, y_off, (size_t)n_tiles * 4, cudaMemcpyHostToDevice, st));
    if (scale) DUNK_CUDA(cudaMemcpyAsync(d_sc, scale, (size_t)n_tiles * 4, cudaMemcpyHostToDevice, st));
    if (image_ids) DUNK_CUDA(cudaMemcpyAsync(d_id, image_ids, (size_t)n_tiles * 4, cudaMemcpyHostToDevice, st));
    std::vector<int> h_off(sub + 2), h_cnt(sub);
    for (int t0 = 0; t0 < n_tiles; t0 += sub) {
        const int nf = std::min(sub, n_tiles - t0);
        DUNK_CUDA(cudaMemcpyAsync(d_img, images + (size_t)t0 * frame_stride_bytes, (size_t)nf * frame_stride_bytes,
                                  cudaMemcpyHostToDevice, st));
        if ((rc = akaze_run(ctx, st, lt, ws, d_img, frame_stride_bytes, row_stride_bytes, channels, nf, max_points <= 0 ? 0 : max_points)))
            return rc;
        k_frame_offsets<<<1, 1024, 0, st>>>(ws.kp_count, nf, d_off, ws.cand_count);
        DUNK_KERNEL_CHECK(ctx);
        DUNK_CUDA(cudaMemcpyAsync(h_off.data(), d_off, (size_t)(nf + 2) * 4, cudaMemcpyDeviceToHost, st));
        DUNK_CUDA(cudaMemcpyAsync(h_cnt.data(), ws.kp_count, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
        DUNK_CUDA(cudaStreamSynchronize(st));
        DUNK_REQUIRE(h_off[nf + 1] <= cap, DUNK_ERR_NO_MEM, "dunk_db_append_tiles: a tile produced %d raw extrema, candidate capacity %d",
                     h_off[nf + 1], cap);
        const int total = h_off[nf];
        DUNK_REQUIRE(db->size + total <= db->capacity, DUNK_ERR_NO_MEM,
                     "dunk_db_append_tiles: %lld + %d rows exceed capacity %lld", (long long)db->size, total,
                     (long long)db->capacity);
        int maxc = 0;
        for (int f = 0; f < nf; ++f) {
            maxc = std::max(maxc, h_cnt[f]);
            if (counts) counts[t0 + f] = h_cnt[f];
        }
        if (maxc > 0) {
            k_append_rows<<<dim3(div_up(maxc, 256), nf), 256, 0, st>>>(ws.desc64, ws.kps, cap, ws.kp_count, d_off,
                                                                      x_off ? d_xo : nullptr, y_off ? d_yo : nullptr,
                                                                      scale ? d_sc : nullptr, image_ids ? d_id : nullptr, t0,
                                                                      db->desc64, db->kps, db->image_id, db->size);
            DUNK_KERNEL_CHECK(ctx);
            DUNK_CUDA(cudaStreamSynchronize(st));
        }
        db->size += total;
    }
    return DUNK_OK;
}

}  // extern "C"

/* ---- reference-DB build straight from the scene's bands ----------------------
 * preprocessor/src/main.rs:160-327: for every level of detail lod in 0..lods the scene is cut into
 * (tile_w * 2^lod) x (tile_h * 2^lod) windows, tile = scene / 2^(lods-1) (integer division, remainder
 * rows / columns dropped), each window is read at tile resolution (GDAL Lanczos in the reference),
 * converted by band_merger, extracted, and its rows inserted with x * 2^lod + column offset.  Here the
 * whole chain runs on the device per tile batch: resample + radiometric conversion fused into one
 * kernel that writes BGRA tiles, batched AKAZE, row append, ref_image rows. */
namespace dunk {
namespace {

constexpr int kMaxTaps = 6 * 16;   // Lanczos-3 support at decimation 16

struct ResampleTaps {
    int n;                 // taps per axis (1 = copy)
    int first;             // offset of tap 0 relative to floor(centre)
    double w[kMaxTaps];
};

// One thread per output pixel: 3 bands resampled (f64 accumulation in a fixed tap order, weights
// renormalised where the support leaves the scene), then f32_to_u8 per channel -> BGRA.
__global__ void __launch_bounds__(256)
k_lod_tiles(const float* __restrict__ red, const float* __restrict__ green, const float* __restrict__ blue, int W, int H,
            int tile_w, int tile_h, int scale, int tiles_x, int tile0, ResampleTaps taps, float rmin, float rmax, float gmin,
            float gmax, float bmin, float bmax, const float* __restrict__ thr, uchar4* __restrict__ out) {
    __shared__ float s_thr[256];
    s_thr[threadIdx.x] = thr[threadIdx.x];
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, t = tile0 + blockIdx.z;
    if (x >= tile_w) return;
    const int tcol = t % tiles_x, trow = t / tiles_x;
    const long long sx0 = ((long long)tcol * tile_w + x) * scale, sy0 = ((long long)trow * tile_h + y) * scale;   // window of this pixel
    const float* bands[3] = {red, green, blue};
    float v[3];
    if (taps.n == 1) {
#pragma unroll
        for (int b = 0; b < 3; ++b) v[b] = bands[b][(size_t)sy0 * W + sx0];
    } else {
        // centre of the output pixel in source coordinates is (s0 + scale / 2); taps start at first
        const long long cx = sx0 + scale / 2, cy = sy0 + scale / 2;
        double acc[3] = {0, 0, 0}, wsum_y = 0;
        for (int ky = 0; ky < taps.n; ++ky) {
            const long long yy = cy + taps.first + ky;
            if (yy < 0 || yy >= H) continue;
            double row[3] = {0, 0, 0}, wsum_x = 0;
            for (int kx = 0; kx < taps.n; ++kx) {
                const long long xx = cx + taps.first + kx;
                if (xx < 0 || xx >= W) continue;
                const double w = taps.w[kx];
                wsum_x = __dadd_rn(wsum_x, w);
#pragma unroll
                for (int b = 0; b < 3; ++b) row[b] = __dadd_rn(row[b], __dmul_rn(w, (double)bands[b][(size_t)yy * W + xx]));
            }
            const double wy = taps.w[ky];
            wsum_y = __dadd_rn(wsum_y, wy);
#pragma unroll
            for (int b = 0; b < 3; ++b) acc[b] = __dadd_rn(acc[b], __dmul_rn(wy, __ddiv_rn(row[b], wsum_x)));
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) v[b] = (float)__ddiv_rn(acc[b], wsum_y);
    }
    auto to_u8 = [&](float val, float lo, float hi) -> unsigned char { return f32_to_u8_lut(s_thr, val, lo, hi); };
    uchar4 o;
    o.z = to_u8(v[0], rmin, rmax);   // BGRA (raster_to_mat)
    o.y = to_u8(v[1], gmin, gmax);
    o.x = to_u8(v[2], bmin, bmax);
    o.w = (isnan(v[0]) && isnan(v[1]) && isnan(v[2])) ? 0 : 255;
    out[((size_t)blockIdx.z * tile_h + y) * tile_w + x] = o;
}

double lanczos3(double x) {
    if (x == 0.0) return 1.0;
    if (std::fabs(x) >= 3.0) return 0.0;
    const double px = M_PI * x;
    return 3.0 * std::sin(px) * std::sin(px / 3.0) / (px * px);
}

// taps for decimation by `scale` (a power of two): area = box mean, lanczos = GDAL-style convolution
// kernel stretched by the decimation factor (support 3 * scale either side of the pixel centre)
bool make_taps(int scale, int resample, ResampleTaps* t) {
    if (scale == 1) { t->n = 1; t->first = 0; t->w[0] = 1.0; return true; }
    if (resample == 0) {
        t->n = scale; t->first = -(scale / 2);
        for (int k = 0; k < scale; ++k) t->w[k] = 1.0;
        return true;
    }
    const int n = 6 * scale;
    if (n > kMaxTaps) return false;
    t->n = n; t->first = -3 * scale;
    // source pixel j covers [j, j+1); its centre sits (k + first + 0.5) away from the output centre
    for (int k = 0; k < n; ++k) t->w[k] = lanczos3(((double)(k + t->first) + 0.5) / scale);
    return true;
}

}  // namespace
}  // namespace dunk

namespace dunk {
namespace {
// Helper thread of the host-band build: copies the three planes stripe by stripe (pageable -> pinned ring half ->
// device, on `stream`) and records one event per finished stripe.  The owner waits host-side until the event of the
// stripe it needs has been RECORDED (cudaStreamWaitEvent on a never-recorded event would be a no-op).
struct StripeUpload {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    int recorded = 0, rc = DUNK_OK;
    std::vector<cudaEvent_t> ev;
    cudaEvent_t ring_ev[2] = {nullptr, nullptr};

    void start(int device, cudaStream_t stream, char* ring, const float* const (&host)[3], float* const (&dev)[3], int width,
               int height, int stripe_rows) {
        const float* h0 = host[0]; const float* h1 = host[1]; const float* h2 = host[2];
        float* d0 = dev[0]; float* d1 = dev[1]; float* d2 = dev[2];
        th = std::thread([=] {
            const float* hb[3] = {h0, h1, h2};
            float* db[3] = {d0, d1, d2};
            const size_t half = (size_t)64 << 20;
            int h = 0, err = DUNK_OK;
            auto ok = [&](cudaError_t e) {
                if (e != cudaSuccess && err == DUNK_OK) err = DUNK_ERR_CUDA;
                return e == cudaSuccess;
            };
            ok(cudaSetDevice(device));
            for (size_t s = 0; s < ev.size(); ++s) {
                const size_t r0 = s * (size_t)stripe_rows, r1 = std::min<size_t>(height, r0 + stripe_rows);
                const size_t bytes = (r1 - r0) * (size_t)width * 4;
                for (int b = 0; b < 3 && err == DUNK_OK; ++b) {
                    const char* src = (const char*)(hb[b] + r0 * width);
                    char* dst = (char*)(db[b] + r0 * width);
                    for (size_t off = 0; off < bytes && err == DUNK_OK; off += half, h ^= 1) {
                        const size_t n = std::min(half, bytes - off);
                        if (!ok(cudaEventSynchronize(ring_ev[h]))) break;      // the half's previous copy has left it
                        par_memcpy(ring + h * half, src + off, n);
                        ok(cudaMemcpyAsync(dst + off, ring + h * half, n, cudaMemcpyHostToDevice, stream));
                        ok(cudaEventRecord(ring_ev[h], stream));
                    }
                }
                if (err == DUNK_OK) ok(cudaEventRecord(ev[s], stream));
                {
                    std::lock_guard<std::mutex> lk(mu);
                    rc = err;
                    if (err == DUNK_OK) recorded = (int)s + 1;
                }
                cv.notify_all();
                if (err != DUNK_OK) return;
            }
        });
    }
    // blocks until stripe `s` has been enqueued + its event recorded (or the upload failed)
    int wait_recorded(int s) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return rc != DUNK_OK || recorded > s; });
        return rc;
    }
    ~StripeUpload() {
        if (th.joinable()) th.join();
        for (cudaEvent_t e : ev)
            if (e) { cudaEventSynchronize(e); cudaEventDestroy(e); }
        for (cudaEvent_t e : ring_ev)
            if (e) { cudaEventSynchronize(e); cudaEventDestroy(e); }
    }
};

// bands_on_device: red/green/blue are device pointers (the scene already in HBM); otherwise host pointers copied in first
int build_from_bands(dunk_db* db, const float* red, const float* green, const float* blue, bool bands_on_device, int width, int height,
                     const double* min_max, int lods, int resample, int max_points, int* n_tiles_out, int* tile_w_out,
                     int* tile_h_out, int part = 0, int n_parts = 1) {
    DUNK_REQUIRE(n_parts >= 1 && part >= 0 && part < n_parts, DUNK_ERR_BAD_ARG, "dunk_db_build_from_bands: part %d of %d", part, n_parts);
    DUNK_REQUIRE(db && red && green && blue && min_max, DUNK_ERR_BAD_ARG, "dunk_db_build_from_bands: NULL argument");
    DUNK_REQUIRE(lods >= 1 && lods <= 5, DUNK_ERR_BAD_ARG, "dunk_db_build_from_bands: lods=%d (1..5)", lods);
    DUNK_REQUIRE(resample == 0 || resample == 1, DUNK_ERR_BAD_ARG, "dunk_db_build_from_bands: resample %d (0 = area, 1 = Lanczos-3)", resample);
    const int tile_w = width >> (lods - 1), tile_h = height >> (lods - 1);      // main.rs:212
    DUNK_REQUIRE(tile_w >= 16 && tile_h >= 16, DUNK_ERR_ASSERT, "dunk_db_build_from_bands: tiles of %dx%d are too small", tile_w, tile_h);
    if (tile_w_out) *tile_w_out = tile_w;
    if (tile_h_out) *tile_h_out = tile_h;
    dunk_ctx* ctx = db->ctx;
    std::lock_guard<std::mutex> lk(db->mu);
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const size_t plane = (size_t)width * height;
    // tiles per extraction batch: about 64 Mpx of tile area (the batch the stencil kernels fill the GPU with)
    const int sub = (int)std::min<long long>(64, std::max<long long>(1, (64ll << 20) / ((long long)tile_w * tile_h)));
    const LevelTable lt = make_level_table(tile_w, tile_h);
    long long c = (long long)tile_w * tile_h / 32;
    c = std::min<long long>(std::max<long long>(c, 2048), 1 << 20);
    const int cap = (int)c;
    const size_t ws_bytes = akaze_workspace_bytes(lt, sub, cap, cap);
    const size_t tile_bytes = (size_t)tile_w * tile_h * 4;
    const size_t need = (bands_on_device ? 0 : 3 * al(plane * 4)) + al(ws_bytes) + al(sub * tile_bytes) + al((sub + 2) * 4) + 4 * al(sub * 4);
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    char* ptr = (char*)scratch;
    const float* d_band[3] = {red, green, blue};
    // Host bands (3 x 482 MB pageable planes at config 4) are uploaded by a helper thread in STRIPES of one LoD-0 tile
    // row, in the order the tile walk needs them, through the slot's pinned ring on the second stream; the tile
    // batches below wait (stream-side) for the last stripe they read, so extraction of row r overlaps the upload of
    // rows r+1...  (The plain sequence upload -> build spent 45 of its 66 ms in the copy.)
    StripeUpload up;
    const int n_stripes = bands_on_device ? 0 : div_up(height, tile_h);
    if (!bands_on_device) {
        if (!ctx->upload_ring(g.s)) return DUNK_ERR_NO_MEM;
        float* dev[3];
        for (int b = 0; b < 3; ++b) {
            dev[b] = (float*)ptr;
            d_band[b] = dev[b];
            ptr += al(plane * 4);
        }
        up.ev.resize(n_stripes, nullptr);
        for (auto& e : up.ev) DUNK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        DUNK_CUDA(cudaEventCreateWithFlags(&up.ring_ev[0], cudaEventDisableTiming));
        DUNK_CUDA(cudaEventCreateWithFlags(&up.ring_ev[1], cudaEventDisableTiming));
        const float* h_band[3] = {red, green, blue};
        try {
            up.start(ctx->device, g.slot().stream2, (char*)g.slot().ring, h_band, dev, width, height, tile_h);
        } catch (...) {      // std::thread could not be created: no exception may cross the C ABI
            set_error("dunk_db_build_from_bands: cannot start the upload thread");
            return DUNK_ERR_NO_MEM;
        }
    }
    AkazeWorkspace ws;
    akaze_carve_workspace(ptr, lt, sub, cap, cap, &ws);
    ptr += al(ws_bytes);
    unsigned char* d_tiles = (unsigned char*)ptr; ptr += al(sub * tile_bytes);
    int* d_off = (int*)ptr; ptr += al((sub + 2) * 4);
    float* d_xo = (float*)ptr; ptr += al(sub * 4);
    float* d_yo = (float*)ptr; ptr += al(sub * 4);
    float* d_sc = (float*)ptr; ptr += al(sub * 4);
    int32_t* d_id = (int32_t*)ptr;
    // the reference walks lod-major, tile-minor (main.rs:212-271); a batch may span several LoDs
    struct TileJob { int lod, t, tiles_x; };
    std::vector<TileJob> jobs;
    for (int lod = 0; lod < lods; ++lod) {
        const int scale = 1 << lod;
        const int tiles_x = width / (tile_w * scale), tiles_y = height / (tile_h * scale);    // main.rs:215-216
        ResampleTaps taps;
        DUNK_REQUIRE(make_taps(scale, resample, &taps), DUNK_ERR_BAD_ARG, "dunk_db_build_from_bands: decimation %d too large", scale);
        for (int t = 0; t < tiles_x * tiles_y; ++t) jobs.push_back({lod, t, tiles_x});
    }
    // every part inserts ALL ref_image rows (InsertImage, main.rs:283-289; ids follow the walk order, so they agree
    // across the ranks of a partitioned build) and extracts only its contiguous share of the tiles
    const int32_t id_base = (int32_t)db->images.size();
    for (size_t j = 0; j < jobs.size(); ++j) {
        const TileJob& jb = jobs[j];
        const int scale = 1 << jb.lod, col = jb.t % jb.tiles_x, row = jb.t / jb.tiles_x;
        const int xs = col * tile_w * scale, ys = row * tile_h * scale;
        db->images.push_back(DunkImage{id_base + (int32_t)j + 1, xs, ys, xs + tile_w * scale - 1, ys + tile_h * scale - 1, jb.lod});
    }
    db->image_lod_dirty = true;
    const size_t j_begin = jobs.size() * (size_t)part / n_parts, j_end = jobs.size() * (size_t)(part + 1) / n_parts;
    const float* thr = gamma_table(ctx, st);
    DUNK_REQUIRE(thr, DUNK_ERR_CUDA, "dunk_db_build_from_bands: gamma table");
    int n_tiles = 0;
    std::vector<int> h_off(sub + 2), h_cnt(sub);
    for (size_t j0 = j_begin, j_next; j0 < j_end; j0 = j_next) {
        int nf = (int)std::min<size_t>(sub, j_end - j0);
        if (n_stripes) {
            // host bands: LoD-0 batches are one tile row (= one upload stripe) so they start as soon as it has landed
            if (jobs[j0].lod == 0) {
                const int row = jobs[j0].t / jobs[j0].tiles_x;
                int k = 1;
                while (k < nf && jobs[j0 + k].lod == 0 && jobs[j0 + k].t / jobs[j0 + k].tiles_x == row) ++k;
                nf = k;
            }
            int need = 0;
            for (int f = 0; f < nf; ++f) {
                const TileJob& jb = jobs[j0 + f];
                const int scale = 1 << jb.lod, row = jb.t / jb.tiles_x;
                // rows [row * tile_h * scale, (row + 1) * tile_h * scale) plus the Lanczos support of 3 * scale rows
                const int last_row = std::min(height - 1, (row + 1) * tile_h * scale - 1 + (resample ? 3 * scale : 0));
                need = std::max(need, last_row / tile_h);
            }
            const int rcu = up.wait_recorded(need);
            DUNK_REQUIRE(rcu == DUNK_OK, rcu, "dunk_db_build_from_bands: band upload failed");
            DUNK_CUDA(cudaStreamWaitEvent(st, up.ev[need], 0));
        }
        j_next = j0 + nf;
        for (int f0 = 0; f0 < nf;) {            // one resample launch per run of equal LoD
            int f1 = f0;
            while (f1 < nf && jobs[j0 + f1].lod == jobs[j0 + f0].lod) ++f1;
            const TileJob& jb = jobs[j0 + f0];
            const int scale = 1 << jb.lod, run = f1 - f0;
            ResampleTaps taps;
            make_taps(scale, resample, &taps);
            ProfScope ps(ctx, st, "lod.resample_merge", (double)run * tile_w * tile_h * (12.0 * scale * scale + 4.0));
            k_lod_tiles<<<dim3(div_up(tile_w, 256), tile_h, run), 256, 0, st>>>(
                d_band[0], d_band[1], d_band[2], width, height, tile_w, tile_h, scale, jb.tiles_x, jb.t, taps, (float)min_max[0],
                (float)min_max[1], (float)min_max[2], (float)min_max[3], (float)min_max[4], (float)min_max[5], thr,
                (uchar4*)(d_tiles + (size_t)f0 * tile_bytes));
            DUNK_KERNEL_CHECK(ctx);
            f0 = f1;
        }
        int rc = akaze_run(ctx, st, lt, ws, d_tiles, tile_bytes, tile_w * 4, 4, nf, max_points <= 0 ? 0 : max_points);
        if (rc) return rc;
        k_frame_offsets<<<1, 1024, 0, st>>>(ws.kp_count, nf, d_off, ws.cand_count);
        DUNK_KERNEL_CHECK(ctx);
        DUNK_CUDA(cudaMemcpyAsync(h_off.data(), d_off, (size_t)(nf + 2) * 4, cudaMemcpyDeviceToHost, st));
        DUNK_CUDA(cudaMemcpyAsync(h_cnt.data(), ws.kp_count, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
        DUNK_CUDA(cudaStreamSynchronize(st));
        DUNK_REQUIRE(h_off[nf + 1] <= cap, DUNK_ERR_NO_MEM, "dunk_db_build_from_bands: a tile produced %d raw extrema, candidate capacity %d",
                     h_off[nf + 1], cap);
        const int rows_new = h_off[nf];
        DUNK_REQUIRE(db->size + rows_new <= db->capacity, DUNK_ERR_NO_MEM, "dunk_db_build_from_bands: %lld + %d rows exceed capacity %lld",
                     (long long)db->size, rows_new, (long long)db->capacity);
        std::vector<float> xo(nf), yo(nf), sc(nf);
        std::vector<int32_t> ids(nf);
        int maxc = 0;
        for (int f = 0; f < nf; ++f) {
            const TileJob& jb = jobs[j0 + f];
            const int scale = 1 << jb.lod, col = jb.t % jb.tiles_x, row = jb.t / jb.tiles_x;
            const int xs = col * tile_w * scale, ys = row * tile_h * scale;
            ids[f] = id_base + (int32_t)(j0 + f) + 1;
            xo[f] = (float)xs; yo[f] = (float)ys; sc[f] = (float)scale;
            maxc = std::max(maxc, h_cnt[f]);
        }
        DUNK_CUDA(cudaMemcpyAsync(d_xo, xo.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
        DUNK_CUDA(cudaMemcpyAsync(d_yo, yo.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
        DUNK_CUDA(cudaMemcpyAsync(d_sc, sc.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
        DUNK_CUDA(cudaMemcpyAsync(d_id, ids.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
        if (maxc > 0) {
            k_append_rows<<<dim3(div_up(maxc, 256), nf), 256, 0, st>>>(ws.desc64, ws.kps, cap, ws.kp_count, d_off, d_xo, d_yo, d_sc, d_id,
                                                                      0, db->desc64, db->kps, db->image_id, db->size);
            DUNK_KERNEL_CHECK(ctx);
        }
        DUNK_CUDA(cudaStreamSynchronize(st));
        db->size += rows_new;
        n_tiles += nf;
    }
    if (n_tiles_out) *n_tiles_out = n_tiles;
    return DUNK_OK;
}
}  // namespace
}  // namespace dunk

extern "C" int dunk_db_build_from_bands(dunk_db* db, const float* red, const float* green, const float* blue, int width, int height,
                                        const double* min_max, int lods, int resample, int max_points, int* n_tiles_out,
                                        int* tile_w_out, int* tile_h_out) {
    return dunk::build_from_bands(db, red, green, blue, false, width, height, min_max, lods, resample, max_points, n_tiles_out,
                                  tile_w_out, tile_h_out);
}

extern "C" int dunk_db_build_from_bands_part_dev(dunk_db* db, const void* red_dev, const void* green_dev, const void* blue_dev,
                                                 int width, int height, const double* min_max, int lods, int resample,
                                                 int max_points, int part, int n_parts, int* n_tiles_out, int* tile_w_out,
                                                 int* tile_h_out) {
    return dunk::build_from_bands(db, (const float*)red_dev, (const float*)green_dev, (const float*)blue_dev, true, width, height,
                                  min_max, lods, resample, max_points, n_tiles_out, tile_w_out, tile_h_out, part, n_parts);
}

extern "C" int dunk_db_build_from_bands_dev(dunk_db* db, const void* red_dev, const void* green_dev, const void* blue_dev, int width,
                                            int height, const double* min_max, int lods, int resample, int max_points,
                                            int* n_tiles_out, int* tile_w_out, int* tile_h_out) {
    return dunk::build_from_bands(db, (const float*)red_dev, (const float*)green_dev, (const float*)blue_dev, true, width, height,
                                  min_max, lods, resample, max_points, n_tiles_out, tile_w_out, tile_h_out);
}
