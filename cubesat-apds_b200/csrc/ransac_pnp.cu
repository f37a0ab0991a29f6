// Stage 3b of the hot path: batched PnP-RANSAC pose (replaces cv::solvePnPRansac as called by
// pnp_solver_ransac, homographier/src/homographier/mod.rs:320-369: zero distortion, EPnP kernel,
// caller-chosen iteration count / reprojection threshold / confidence).
//
// One CTA per problem (= one frame's 3-D/2-D correspondences), so a frame batch is one launch and
// the stage partitions by frame with no collective.  Inside a CTA:
//   * thread 0 replays OpenCV's fixed-seed MWC sample stream (5 distinct indices per hypothesis, no
//     degeneracy test — PnPRansacCallback does not override checkSubset);
//   * one THREAD per hypothesis solves the 5-point EPnP (pnp_math.cuh: f64, local memory), goes
//     through the rvec round trip the reference model takes (Rodrigues there and back);
//   * one WARP per hypothesis scores it: f64 projection rounded to f32, f32 squared error, inlier
//     iff err <= (float)thr^2, warp-reduced count;
//   * thread 0 applies the sequential accept / adaptive-iteration rule in stream order, so the
//     result is what the sequential CPU loop produces;
//   * the final pose is EPnP over all inliers of the best hypothesis, CTA-wide: every thread runs
//     the scalar part redundantly, the sums over points are block reductions.
#include <algorithm>
#include <vector>
#include "ctx.h"
#include "pnp_math.cuh"

namespace dunk {
namespace {

using namespace pnp;

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxRound = 128;       // hypotheses solved + scored per round
constexpr int kFirstRound = 32;      // the adaptive stop usually fires inside the first round
constexpr unsigned long long kRngCoeff = 4164903690ull;
constexpr int kMaxModelPoints = 5;   // EPnP kernel: 5-point samples; P3P kernel: 4-point samples

struct Rng {
    unsigned long long state;
    __device__ unsigned next() {
        state = (unsigned long long)(unsigned)state * kRngCoeff + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    __device__ int uniform(int n) { return (int)(next() % (unsigned)n); }
};

__device__ int update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmin(fmax(p, 0.), 1.);
    ep = fmin(fmax(ep, 0.), 1.);
    double num = fmax(1. - p, DBL_MIN);
    double denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : __double2int_rn(num / denom);
}

struct Shared {
    unsigned long long rng_state;
    int niters, iter, best_count, best_h, done, nh;
    int hyp[kMaxRound][kMaxModelPoints];
    int counts[kMaxRound];
    double model[kMaxRound][12];     // Rodrigues(rvec(R)) (9) + t (3) of every hypothesis of the round
    double best[12];
    double red_buf[kWarps * 78];
    double red[78];
    int first_inlier, n_inliers;
    int wsum[kWarps];
};

// CTA-wide executor for the final solve: the used points are those with mask[i] != 0
struct BlockExec {
    static constexpr bool kBlock = true;
    const float* obj;
    const float* img;
    const uint8_t* mask;
    int n_total, n_used, first_idx;
    Camera cam;
    Shared* sh;
    __device__ double count() const { return (double)n_used; }
    __device__ Point first() const { return load_point(obj, img, first_idx, cam, false); }
    template <int K, class F>
    __device__ void sum(F f, double (&out)[K]) const {
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0;
        for (int i = threadIdx.x; i < n_total; i += kThreads)
            if (mask[i]) f(load_point(obj, img, i, cam, false), acc);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __syncthreads();   // previous users of red[] are done
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = acc[k];
#pragma unroll
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) sh->red_buf[warp * K + k] = x;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += kThreads) {
            double s = 0;
            for (int w = 0; w < kWarps; ++w) s += sh->red_buf[w * K + k];
            sh->red[k] = s;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = sh->red[k];
    }
};

__device__ int warp_count_inliers(const double* model, const Camera& cam, const float* __restrict__ obj,
                                  const float* __restrict__ img, int n, float thr2, int lane) {
    double R2[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R2[i] = model[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = model[9 + i];
    int c = 0;
    for (int i = lane; i < n; i += 32) c += (reproj_err_f32(R2, t, cam, obj, img, i) <= thr2);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}

// minimal solve (5-point EPnP or 4-point P3P) + the model's rvec round trip; false when there is no finite pose
__device__ bool solve_minimal(const float* obj, const float* img, const int* idx, int n_pts, const Camera& cam, double* model) {
    double R[9], t[3], r[3];
    if (n_pts == 4) {
        if (!p3p_solve4(obj, img, idx, cam, true, R, t)) return false;
    } else {
        SerialExec ex{obj, img, idx, n_pts, cam, true};
        epnp_solve(ex, cam, R, t);
    }
    rodrigues_to_vector(R, r);
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) ok = ok && isfinite(r[i]) && isfinite(t[i]);
    rodrigues_to_matrix(r, model);
#pragma unroll
    for (int i = 0; i < 3; ++i) model[9 + i] = t[i];
    return ok;
}

// obj: [total][3] f32, img: [total][2] f32 (already rounded from the caller's f64, as OpenCV does)
__global__ void __launch_bounds__(kThreads)
pnp_ransac_kernel(const float* __restrict__ obj_all, const float* __restrict__ img_all, const int* __restrict__ starts,
                  const int* __restrict__ counts, const double* __restrict__ K_all, int k_stride, int method, int max_iters, float thr,
                  double confidence, double* __restrict__ rt_out /* [B][6] rvec, tvec */, uint8_t* __restrict__ mask_out,
                  int* __restrict__ info_out /* [B][4]: found, inliers, iterations, hypotheses */) {
    __shared__ Shared sh;
    const int b = blockIdx.x;
    const int off = starts[b], n = counts[b];
    const float* obj = obj_all + (size_t)off * 3;
    const float* img = img_all + (size_t)off * 2;
    uint8_t* mask = mask_out + off;
    const double* Kp = K_all + (size_t)b * k_stride;
    const Camera cam{Kp[0], Kp[4], Kp[2], Kp[5]};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float thr2 = (float)((double)thr * (double)thr);

    auto finish = [&](int found, int inl, int iters, int nhyp, const double* rvec, const double* tvec) {
        if (tid == 0) {
            info_out[b * 4 + 0] = found; info_out[b * 4 + 1] = inl;
            info_out[b * 4 + 2] = iters; info_out[b * 4 + 3] = nhyp;
            for (int i = 0; i < 3; ++i) {
                rt_out[b * 6 + i] = found ? rvec[i] : 0.0;
                rt_out[b * 6 + 3 + i] = found ? tvec[i] : 0.0;
            }
        }
    };

    // solvepnp.cpp: 5-point samples + EPnP kernel, except 4-point samples + P3P kernel when P3P is requested or
    // when there are exactly 4 correspondences
    const int mp = (method == DUNK_PNP_P3P || n == 4) ? 4 : 5;
    if (n < 4) {   // the host API rejects this (-215)
        for (int i = tid; i < n; i += kThreads) mask[i] = 0;
        finish(0, 0, 0, 0, nullptr, nullptr);
        return;
    }
    if (n == mp) {
        // model_points == npoints -> one solvePnP on all (f32) points, every point an inlier
        double R[9], t[3], r[3];
        bool ok = true;
        if (mp == 4) ok = p3p_solve4(obj, img, nullptr, cam, true, R, t);
        else {
            SerialExec ex{obj, img, nullptr, n, cam, true};
            epnp_solve(ex, cam, R, t);
        }
        if (!ok) {
            for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0);
            t[0] = t[1] = t[2] = 0;
        }
        rodrigues_to_vector(R, r);
        for (int i = 0; i < 3; ++i) ok = ok && isfinite(r[i]) && isfinite(t[i]);
        for (int i = tid; i < n; i += kThreads) mask[i] = ok;
        finish(ok, ok ? n : 0, 0, 0, r, t);
        return;
    }

    if (tid == 0) {
        sh.rng_state = ~0ull;
        sh.niters = max_iters; sh.iter = 0; sh.best_count = 0; sh.best_h = -1; sh.done = max_iters <= 0; sh.nh = 0;
    }
    __syncthreads();
    int total_hyp = 0, round = 0;
    DUNK_PHASE(0);

    while (!sh.done) {
        // ---- (1) sample stream, thread 0 ----------------------------------------------------
        if (tid == 0) {
            const int want = min(round == 0 ? kFirstRound : kMaxRound, sh.niters - sh.iter);
            Rng rng{sh.rng_state};
            for (int h = 0; h < want; ++h) {
                int idx[kMaxModelPoints];
                for (int i = 0; i < mp; ++i) {
                    int v;
                    bool dup;
                    do {
                        v = rng.uniform(n);
                        dup = false;
                        for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
                    } while (dup);
                    idx[i] = v;
                }
                for (int i = 0; i < mp; ++i) sh.hyp[h][i] = idx[i];
            }
            sh.rng_state = rng.state;
            sh.nh = want;
        }
        __syncthreads();
        const int nh = sh.nh;
        if (round == 0) DUNK_PHASE(1);
        // ---- (2) solve: one thread per hypothesis -------------------------------------------
        if (tid < nh) {
            double model[12];
            const bool ok = solve_minimal(obj, img, sh.hyp[tid], mp, cam, model);
            for (int i = 0; i < 12; ++i) sh.model[tid][i] = model[i];
            sh.counts[tid] = ok ? 0 : -1;
        }
        __syncthreads();
        if (round == 0) DUNK_PHASE(2);
        // ---- (3) score: one warp per hypothesis ---------------------------------------------
        for (int h = warp; h < nh; h += kWarps) {
            if (sh.counts[h] < 0) continue;
            const int c = warp_count_inliers(sh.model[h], cam, obj, img, n, thr2, lane);
            if (lane == 0) sh.counts[h] = c;
        }
        __syncthreads();
        if (round == 0) DUNK_PHASE(3);
        // ---- (4) sequential accept rule in stream order (ptsetreg.cpp run()) -----------------
        if (tid == 0) {
            for (int h = 0; h < nh && sh.iter < sh.niters; ++h) {
                const int good = sh.counts[h];
                if (good > max(sh.best_count, mp - 1)) {
                    sh.best_count = good;
                    for (int i = 0; i < 12; ++i) sh.best[i] = sh.model[h][i];
                    sh.niters = update_num_iters(confidence, (double)(n - good) / n, mp, sh.niters);
                }
                ++sh.iter;
            }
            sh.done = sh.iter >= sh.niters;
        }
        total_hyp += nh;
        ++round;
        __syncthreads();
    }

    if (sh.best_count == 0) {
        for (int i = tid; i < n; i += kThreads) mask[i] = 0;
        finish(0, 0, sh.iter, total_hyp, nullptr, nullptr);
        return;
    }
    DUNK_PHASE(4);
    // ---- (5) inlier mask of the best minimal model, then EPnP over those inliers ------------
    if (tid == 0) { sh.first_inlier = n; sh.n_inliers = 0; }
    __syncthreads();
    {
        double R2[9], t[3];
        for (int i = 0; i < 9; ++i) R2[i] = sh.best[i];
        for (int i = 0; i < 3; ++i) t[i] = sh.best[9 + i];
        int c = 0, first = n;
        for (int i = tid; i < n; i += kThreads) {
            const bool in = reproj_err_f32(R2, t, cam, obj, img, i) <= thr2;
            mask[i] = in;
            c += in;
            if (in && i < first) first = i;
        }
        atomicAdd(&sh.n_inliers, c);
        atomicMin(&sh.first_inlier, first);
    }
    __syncthreads();   // also orders the global mask writes before the block-wide reads below
    BlockExec ex{obj, img, mask, n, sh.n_inliers, sh.first_inlier, cam, &sh};
    double R[9], t[3], r[3];
    DUNK_PHASE(5);
    epnp_solve(ex, cam, R, t);
    DUNK_PHASE(20);
    // SOLVEPNP_ITERATIVE: same RANSAC stage, then the Levenberg-Marquardt minimum over the inliers (solvepnp.cpp calls
    // solvePnP(inliers, flags); with exactly model_points points OpenCV returns the kernel's pose unrefined)
    if (method == DUNK_PNP_ITERATIVE && n > mp) pnp_refine(ex, cam, R, t);
    rodrigues_to_vector(R, r);
    bool ok = true;
    for (int i = 0; i < 3; ++i) ok = ok && isfinite(r[i]) && isfinite(t[i]);
    finish(ok ? 1 : 0, sh.n_inliers, sh.iter, total_hyp, r, t);
    DUNK_PHASE(21);
}

// parity hook: explicit 5-index samples -> per-hypothesis inlier count and pose (one thread solves,
// one warp scores; 8 hypotheses per CTA)
__global__ void __launch_bounds__(kThreads)
pnp_score_kernel(const float* __restrict__ obj, const float* __restrict__ img, int n, const double* __restrict__ Kp,
                 const int* __restrict__ samples, int n_hyp, float thr, int* __restrict__ counts, double* __restrict__ rt) {
    __shared__ double model[kWarps][12];
    __shared__ int okf[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x * kWarps + warp;
    const Camera cam{Kp[0], Kp[4], Kp[2], Kp[5]};
    const float thr2 = (float)((double)thr * (double)thr);
    if (h < n_hyp && lane == 0) {
        SerialExec ex{obj, img, samples + h * 5, 5, cam, true};
        double R[9], t[3], r[3];
        epnp_solve(ex, cam, R, t);
        rodrigues_to_vector(R, r);
        bool ok = true;
        for (int i = 0; i < 3; ++i) ok = ok && isfinite(r[i]) && isfinite(t[i]);
        rodrigues_to_matrix(r, model[warp]);
        for (int i = 0; i < 3; ++i) {
            model[warp][9 + i] = t[i];
            rt[h * 6 + i] = r[i];
            rt[h * 6 + 3 + i] = t[i];
        }
        okf[warp] = ok;
    }
    __syncwarp();
    if (h >= n_hyp) return;
    const int c = okf[warp] ? warp_count_inliers(model[warp], cam, obj, img, n, thr2, lane) : -1;
    if (lane == 0) counts[h] = c;
}

}  // namespace

// device-resident launcher (pipeline.cu composes it behind the homography stage): obj [total][3] f32, img [total][2] f32,
// problem b = points [starts[b], starts[b] + counts[b]), camera matrix at K + b * k_stride (k_stride 0: one K for all)
int launch_pnp_ransac(dunk_ctx* ctx, cudaStream_t st, const float* obj, const float* img, const int* starts, const int* counts,
                      int n_problems, const double* K, int k_stride, int method, int iters, float thr, double confidence,
                      double* rt, uint8_t* mask, int* info) {
    pnp_ransac_kernel<<<n_problems, kThreads, 0, st>>>(obj, img, starts, counts, K, k_stride, method, iters, thr, confidence, rt, mask,
                                                       info);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    return DUNK_OK;
}

}  // namespace dunk

using namespace dunk;

extern "C" {

int dunk_pnp_ransac_batch(dunk_ctx* ctx, const double* obj, const double* img, const int* offsets, int n_problems,
                          const double* K, int iters, float thr, double confidence, int method, double* rvecs, double* tvecs,
                          uint8_t* inlier_mask, int* info) {
    DUNK_REQUIRE(ctx && offsets && K && rvecs && tvecs && info && n_problems >= 0, DUNK_ERR_BAD_ARG,
                 "dunk_pnp_ransac_batch: NULL argument");
    if (n_problems == 0) return DUNK_OK;
    DUNK_REQUIRE(method == DUNK_PNP_EPNP || method == DUNK_PNP_P3P || method == DUNK_PNP_ITERATIVE, DUNK_ERR_BAD_ARG,
                 "dunk_pnp_ransac: method %d not implemented (SOLVEPNP_ITERATIVE = 0, SOLVEPNP_EPNP = 1, the reference's default, "
                 "and SOLVEPNP_P3P = 2 are)",
                 method);
    const int total = offsets[n_problems];
    for (int b = 0; b < n_problems; ++b) {
        const int n = offsets[b + 1] - offsets[b];
        // cv::solvePnPRansac: CV_Assert(npoints >= 4 && ...) -> StsAssert (-215); reference test
        // pnp_solver_ransac_no_work_lthan_3_points, homographier/src/homographier/mod.rs:627-638
        DUNK_REQUIRE(n >= 4, DUNK_ERR_ASSERT, "dunk_pnp_ransac: problem %d has %d correspondences, at least 4 are needed", b, n);
    }
    DUNK_REQUIRE(obj && img, DUNK_ERR_BAD_ARG, "dunk_pnp_ransac_batch: NULL points");
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    // solvepnp.cpp: CV_64F inputs are converted to CV_32F before anything else
    std::vector<float> h_obj((size_t)total * 3), h_img((size_t)total * 2);
    for (size_t i = 0; i < h_obj.size(); ++i) h_obj[i] = (float)obj[i];
    for (size_t i = 0; i < h_img.size(); ++i) h_img[i] = (float)img[i];
    std::vector<int> h_cnt(n_problems);
    for (int b = 0; b < n_problems; ++b) h_cnt[b] = offsets[b + 1] - offsets[b];
    size_t need = Carver::need((size_t)total * 12) + Carver::need((size_t)total * 8) + 2 * Carver::need((size_t)(n_problems + 1) * 4) +
                  Carver::need((size_t)n_problems * 72) + Carver::need((size_t)n_problems * 48) + Carver::need((size_t)total) +
                  Carver::need((size_t)n_problems * 16);
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    float* d_obj = cv.take<float>((size_t)total * 3);
    float* d_img = cv.take<float>((size_t)total * 2);
    int* d_off = cv.take<int>(n_problems + 1);
    int* d_cnt = cv.take<int>(n_problems + 1);
    double* d_K = cv.take<double>((size_t)n_problems * 9);
    double* d_rt = cv.take<double>((size_t)n_problems * 6);
    uint8_t* d_mask = cv.take<uint8_t>(total);
    int* d_info = cv.take<int>((size_t)n_problems * 4);
    DUNK_CUDA(cudaMemcpyAsync(d_obj, h_obj.data(), h_obj.size() * 4, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_img, h_img.data(), h_img.size() * 4, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_off, offsets, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_cnt, h_cnt.data(), (size_t)n_problems * 4, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_K, K, (size_t)n_problems * 72, cudaMemcpyHostToDevice, st));
    {
        ProfScope ps(ctx, st, "ransac.pnp", 0.0);
        const int rc = launch_pnp_ransac(ctx, st, d_obj, d_img, d_off, d_cnt, n_problems, d_K, 9, method, iters, thr, confidence, d_rt,
                                         d_mask, d_info);
        if (rc) return rc;
    }
    std::vector<double> h_rt((size_t)n_problems * 6);
    DUNK_CUDA(cudaMemcpyAsync(h_rt.data(), d_rt, h_rt.size() * 8, cudaMemcpyDeviceToHost, st));
    if (inlier_mask) DUNK_CUDA(cudaMemcpyAsync(inlier_mask, d_mask, (size_t)total, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaMemcpyAsync(info, d_info, (size_t)n_problems * 16, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < n_problems; ++b)
        for (int i = 0; i < 3; ++i) {
            rvecs[b * 3 + i] = h_rt[b * 6 + i];
            tvecs[b * 3 + i] = h_rt[b * 6 + 3 + i];
        }
    return DUNK_OK;
}

int dunk_pnp_ransac(dunk_ctx* ctx, const double* obj, const double* img, int n, const double* K, int iters, float thr,
                    double confidence, int method, double* rvec, double* tvec, int32_t* inliers, int inliers_cap,
                    int* n_inliers, int* found) {
    DUNK_REQUIRE(found && n_inliers, DUNK_ERR_BAD_ARG, "dunk_pnp_ransac: found / n_inliers is NULL");
    *found = 0;
    *n_inliers = 0;
    DUNK_REQUIRE(n >= 0, DUNK_ERR_BAD_ARG, "dunk_pnp_ransac: n < 0");
    const int offsets[2] = {0, n};
    int info[4] = {0, 0, 0, 0};
    std::vector<uint8_t> mask((size_t)n);
    const int rc = dunk_pnp_ransac_batch(ctx, obj, img, offsets, 1, K, iters, thr, confidence, method, rvec, tvec, mask.data(), info);
    if (rc) return rc;
    *found = info[0];
    if (!info[0]) return DUNK_OK;     // OpenCV releases the inlier list when no pose is found
    int k = 0;
    for (int i = 0; i < n; ++i)
        if (mask[i]) {
            DUNK_REQUIRE(!inliers || k < inliers_cap, DUNK_ERR_NO_MEM, "dunk_pnp_ransac: inlier capacity %d too small", inliers_cap);
            if (inliers) inliers[k] = i;
            ++k;
        }
    *n_inliers = k;
    return DUNK_OK;
}

int dunk_pnp_score_hypotheses(dunk_ctx* ctx, const double* obj, const double* img, int n, const double* K, const int* samples,
                              int n_hyp, double thr, int* counts, double* rt) {
    DUNK_REQUIRE(ctx && obj && img && K && samples && counts && rt && n >= 5 && n_hyp >= 0, DUNK_ERR_BAD_ARG,
                 "dunk_pnp_score_hypotheses: bad argument");
    if (n_hyp == 0) return DUNK_OK;
    for (int i = 0; i < n_hyp * 5; ++i)
        DUNK_REQUIRE(samples[i] >= 0 && samples[i] < n, DUNK_ERR_OUT_OF_RANGE,
                     "dunk_pnp_score_hypotheses: sample index %d outside 0..%d", samples[i], n - 1);
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    std::vector<float> h_obj((size_t)n * 3), h_img((size_t)n * 2);
    for (size_t i = 0; i < h_obj.size(); ++i) h_obj[i] = (float)obj[i];
    for (size_t i = 0; i < h_img.size(); ++i) h_img[i] = (float)img[i];
    size_t need = Carver::need((size_t)n * 12) + Carver::need((size_t)n * 8) + Carver::need(72) +
                  Carver::need((size_t)n_hyp * 20) + Carver::need((size_t)n_hyp * 4) + Carver::need((size_t)n_hyp * 48);
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    float* d_obj = cv.take<float>((size_t)n * 3);
    float* d_img = cv.take<float>((size_t)n * 2);
    double* d_K = cv.take<double>(9);
    int* d_s = cv.take<int>((size_t)n_hyp * 5);
    int* d_c = cv.take<int>(n_hyp);
    double* d_rt = cv.take<double>((size_t)n_hyp * 6);
    DUNK_CUDA(cudaMemcpyAsync(d_obj, h_obj.data(), h_obj.size() * 4, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_img, h_img.data(), h_img.size() * 4, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_K, K, 72, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_s, samples, (size_t)n_hyp * 5 * 4, cudaMemcpyHostToDevice, st));
    pnp_score_kernel<<<div_up(n_hyp, kWarps), kThreads, 0, st>>>(d_obj, d_img, n, d_K, d_s, n_hyp, (float)thr, d_c, d_rt);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    DUNK_CUDA(cudaMemcpyAsync(counts, d_c, (size_t)n_hyp * 4, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaMemcpyAsync(rt, d_rt, (size_t)n_hyp * 48, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    return DUNK_OK;
}

#ifdef DUNK_PHASE_TIMING
/* timing variant only: clock64() stamps of block 0 (PnP: 0-21, homography: 32-47) */
int dunk_debug_phases_pnp(long long* out, int n) {
    DUNK_CUDA(cudaDeviceSynchronize());
    DUNK_CUDA(cudaMemcpyFromSymbol(out, g_phase, (size_t)std::min(n, 64) * 8));
    return DUNK_OK;
}
#endif

}  // extern "C"
