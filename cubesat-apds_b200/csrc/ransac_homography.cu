// Stage 3 of the hot path: batched RANSAC homography (replaces cv::findHomography as called at
// homographier/src/homographier/mod.rs:243-250: method RANSAC, maxIters 2000, confidence 0.995).
//
// One CTA per problem (= one query frame's correspondences), so a frame batch is one launch and
// the stage partitions by frame with no collective.  Inside a CTA:
//   * warp 0 replays OpenCV's fixed-seed MWC sample stream (sequential draws by lane 0, the
//     degeneracy / orientation test `checkSubset` for 32 candidate subsets in parallel);
//   * every warp scores hypotheses: all lanes solve the 4-point homography redundantly in f64
//     registers, then stride over the N correspondences with the f32, FMA-free reprojection error
//     OpenCV uses, and warp-reduce the inlier count;
//   * thread 0 applies the sequential accept / adaptive-iteration rule in stream order, so the
//     result is the one the sequential CPU loop produces;
//   * refit on the inliers (normalised DLT, 9x9 symmetric Jacobi eigen, f64), <= 10
//     Levenberg-Marquardt iterations with block-wide f64 reductions, final mask from the final H.
#include "ctx.h"
#include "pipeline.h"
#include <cfloat>
#include <cmath>

namespace dunk {
namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxHyp = 64;          // hypothesis queue per round
constexpr int kRoundTarget = 2 * kWarps;
constexpr unsigned long long kRngCoeff = 4164903690ull;
constexpr int kMaxAttempts = 10000;        // RANSACPointSetRegistrator::run passes 10000 to getSubset
constexpr int kMaxAttemptsLmeds = 1000;    // LMeDSPointSetRegistrator::run uses getSubset's default (ptsetreg.cpp)

struct Rng {
    unsigned long long state;
    __device__ unsigned next() {
        state = (unsigned long long)(unsigned)state * kRngCoeff + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    __device__ int uniform(int n) { return (int)(next() % (unsigned)n); }
};

__device__ __forceinline__ double det3(double a00, double a01, double a10, double a11, double a20,
                                       double a21) {
    // | a00 a01 1 ; a10 a11 1 ; a20 a21 1 |
    return a00 * (a11 - a21) - a01 * (a10 - a20) + (a10 * a21 - a20 * a11);
}

// precomp.hpp haveCollinearPoints: only the last point is tested against earlier pairs
__device__ bool have_collinear(const float2* p) {
    const int i = 3;
    for (int j = 0; j < i; ++j) {
        const double dx1 = (double)p[j].x - (double)p[i].x, dy1 = (double)p[j].y - (double)p[i].y;
        for (int k = 0; k < j; ++k) {
            const double dx2 = (double)p[k].x - (double)p[i].x, dy2 = (double)p[k].y - (double)p[i].y;
            if (fabs(dx2 * dy1 - dy2 * dx1) <=
                (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
                return true;
        }
    }
    return false;
}

__device__ bool check_subset(const float2* s, const float2* d) {
    if (have_collinear(s) || have_collinear(d)) return false;
    const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
    int negative = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = tt[i][0], b = tt[i][1], c = tt[i][2];
        const double A = det3(s[a].x, s[a].y, s[b].x, s[b].y, s[c].x, s[c].y);
        const double B = det3(d[a].x, d[a].y, d[b].x, d[b].y, d[c].x, d[c].y);
        negative += (A * B < 0);
    }
    return negative == 0 || negative == 4;
}

// unit square -> quad (Heckbert); q[k] = image of (0,0),(1,0),(1,1),(0,1)
__device__ void square_to_quad(const double (&qx)[4], const double (&qy)[4], double (&m)[9]) {
    const double dx1 = qx[1] - qx[2], dx2 = qx[3] - qx[2], sx = qx[0] - qx[1] + qx[2] - qx[3];
    const double dy1 = qy[1] - qy[2], dy2 = qy[3] - qy[2], sy = qy[0] - qy[1] + qy[2] - qy[3];
    const double den = dx1 * dy2 - dy1 * dx2;
    const double g = (sx * dy2 - dx2 * sy) / den;
    const double h = (dx1 * sy - sx * dy1) / den;
    m[0] = qx[1] - qx[0] + g * qx[1]; m[1] = qx[3] - qx[0] + h * qx[3]; m[2] = qx[0];
    m[3] = qy[1] - qy[0] + g * qy[1]; m[4] = qy[3] - qy[0] + h * qy[3]; m[5] = qy[0];
    m[6] = g; m[7] = h; m[8] = 1.0;
}

__device__ void mat3_mul(const double (&a)[9], const double (&b)[9], double (&c)[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
__device__ void mat3_adj(const double (&m)[9], double (&a)[9]) {
    a[0] = m[4] * m[8] - m[5] * m[7]; a[1] = m[2] * m[7] - m[1] * m[8]; a[2] = m[1] * m[5] - m[2] * m[4];
    a[3] = m[5] * m[6] - m[3] * m[8]; a[4] = m[0] * m[8] - m[2] * m[6]; a[5] = m[2] * m[3] - m[0] * m[5];
    a[6] = m[3] * m[7] - m[4] * m[6]; a[7] = m[1] * m[6] - m[0] * m[7]; a[8] = m[0] * m[4] - m[1] * m[3];
}

// exact 4-point homography src -> dst in normalised coordinates (same normalisation as
// HomographyEstimatorCallback::runKernel), H[8] = 1.  false if the scale test fails.
__device__ bool homography_4pt(const float2* s, const float2* d, double (&H)[9]) {
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { cMx += s[i].x; cMy += s[i].y; cmx += d[i].x; cmy += d[i].y; }
    cMx *= 0.25; cMy *= 0.25; cmx *= 0.25; cmy *= 0.25;
    double sMx = 0, sMy = 0, smx = 0, smy = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sMx += fabs(s[i].x - cMx); sMy += fabs(s[i].y - cMy);
        smx += fabs(d[i].x - cmx); smy += fabs(d[i].y - cmy);
    }
    if (fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON || fabs(smx) < DBL_EPSILON ||
        fabs(smy) < DBL_EPSILON)
        return false;
    sMx = 4.0 / sMx; sMy = 4.0 / sMy; smx = 4.0 / smx; smy = 4.0 / smy;
    double X[4], Y[4], x[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        X[i] = (s[i].x - cMx) * sMx; Y[i] = (s[i].y - cMy) * sMy;
        x[i] = (d[i].x - cmx) * smx; y[i] = (d[i].y - cmy) * smy;
    }
    double A[9], B[9], Aadj[9], H0[9];
    square_to_quad(X, Y, A);
    square_to_quad(x, y, B);
    mat3_adj(A, Aadj);
    mat3_mul(B, Aadj, H0);
    // H = invHnorm * H0 * Hnorm2
    const double inv[9] = {1.0 / smx, 0, cmx, 0, 1.0 / smy, cmy, 0, 0, 1};
    const double nrm[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
    double T[9];
    mat3_mul(inv, H0, T);
    mat3_mul(T, nrm, H);
    const double s22 = 1.0 / H[8];
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] *= s22;
    H[8] = 1.0;
    return true;
}

// HomographyEstimatorCallback::computeError: f32, no FMA contraction
__device__ __forceinline__ float reproj_err(const float (&Hf)[8], float2 M, float2 m) {
    const float ww = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(Hf[6], M.x), __fmul_rn(Hf[7], M.y)), 1.f));
    const float dx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[0], M.x), __fmul_rn(Hf[1], M.y)), Hf[2]), ww), m.x);
    const float dy = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[3], M.x), __fmul_rn(Hf[4], M.y)), Hf[5]), ww), m.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__device__ int warp_count_inliers(const double* H, const float2* __restrict__ src,
                                  const float2* __restrict__ dst, int n, float thr2, int lane) {
    float Hf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) Hf[i] = (float)H[i];
    int c = 0;
    for (int i = lane; i < n; i += 32) c += (reproj_err(Hf, src[i], dst[i]) <= thr2);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}

// LMEDS score: bit pattern of the (n / 2)-th smallest f32 reprojection error (what std::nth_element at count / 2
// leaves there in ptsetreg.cpp).  Non-negative floats order like their bit patterns, so the answer is the
// smallest v with #{err bits <= v} >= n / 2 + 1: 32 bisection steps, each a pass over the points (errors are
// recomputed instead of stored: 55 hypotheses per problem, the path is not throughput critical).
__device__ unsigned warp_median_error_bits(const double* H, const float2* __restrict__ src,
                                           const float2* __restrict__ dst, int n, int lane) {
    float Hf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) Hf[i] = (float)H[i];
    const int need = n / 2 + 1;
    unsigned lo = 0u, hi = 0xffffffffu;
    while (lo < hi) {
        const unsigned mid = lo + (hi - lo) / 2u;
        int c = 0;
        for (int i = lane; i < n; i += 32) c += (__float_as_uint(reproj_err(Hf, src[i], dst[i])) <= mid);
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= need) hi = mid; else lo = mid + 1u;
    }
    return lo;
}

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

// cyclic Jacobi for a symmetric n x n matrix (single thread; A destroyed, rows of V = vectors)
__device__ void jacobi_eigen(double* A, double* V, double* w, int n) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;
        for (int p = 0; p < n; ++p) {
            diag += A[p * n + p] * A[p * n + p];
            for (int q = p + 1; q < n; ++q) off += A[p * n + q] * A[p * n + q];
        }
        if (off <= 1e-32 * diag || off == 0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {  // columns p,q
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {  // rows p,q
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vpk = V[p * n + k], vqk = V[q * n + k];
                    V[p * n + k] = c * vpk - s * vqk;
                    V[q * n + k] = s * vpk + c * vqk;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

// The same cyclic Jacobi run by ONE WARP on matrices in shared memory: lane k owns element k of the two rows /
// columns a rotation touches, so a rotation is three warp-wide steps instead of 3 n serial round trips through
// shared memory (the single-thread version was the critical path of the refit + LM tail: ~100 cycles per element).
// Every element goes through exactly the arithmetic of jacobi_eigen, in the same order -> identical bits.
// All 32 lanes must call it; n <= 32.
__device__ void jacobi_eigen_warp(double* A, double* V, double* w, int n) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < n * n; i += 32) V[i] = (i / n == i % n);
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;                      // every lane sums in the serial order (broadcast reads)
        for (int p = 0; p < n; ++p) {
            diag += A[p * n + p] * A[p * n + p];
            for (int q = p + 1; q < n; ++q) off += A[p * n + q] * A[p * n + q];
        }
        if (off <= 1e-32 * diag || off == 0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0) continue;                // uniform: every lane read the same value
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                __syncwarp();                          // all lanes hold c, s before anything is overwritten
                if (lane < n) {                        // columns p, q
                    const double akp = A[lane * n + p], akq = A[lane * n + q];
                    A[lane * n + p] = c * akp - s * akq;
                    A[lane * n + q] = s * akp + c * akq;
                }
                __syncwarp();
                if (lane < n) {                        // rows p, q of A and of V
                    const double apk = A[p * n + lane], aqk = A[q * n + lane];
                    A[p * n + lane] = c * apk - s * aqk;
                    A[q * n + lane] = s * apk + c * aqk;
                    const double vpk = V[p * n + lane], vqk = V[q * n + lane];
                    V[p * n + lane] = c * vpk - s * vqk;
                    V[q * n + lane] = s * vpk + c * vqk;
                }
                __syncwarp();
            }
    }
    for (int i = lane; i < n; i += 32) w[i] = A[i * n + i];
    __syncwarp();
}

// x = V^T diag(1/w) V b with OpenCV's SVBkSb threshold (2*eps*sum|w|): cv::solve(DECOMP_EIG)
__device__ void eig_backsolve(const double* V, const double* w, const double* b, double* x, int n) {
    double thr = 0;
    for (int i = 0; i < n; ++i) thr += fabs(w[i]);
    thr *= 2 * DBL_EPSILON;
    for (int i = 0; i < n; ++i) x[i] = 0;
    for (int k = 0; k < n; ++k) {
        if (fabs(w[k]) <= thr) continue;
        double s = 0;
        for (int i = 0; i < n; ++i) s += V[k * n + i] * b[i];
        s /= w[k];
        for (int i = 0; i < n; ++i) x[i] += s * V[k * n + i];
    }
}

template <int K>
__device__ void block_reduce(double (&v)[K], double* warp_buf /*[kWarps][K]*/, double* out /*[K]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) warp_buf[warp * K + k] = x;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double s = 0;
        for (int w2 = 0; w2 < kWarps; ++w2) s += warp_buf[w2 * K + k];
        out[k] = s;
    }
    __syncthreads();
}

struct Shared {
    // RANSAC state
    unsigned long long rng_state;
    int niters, iter, best_count, rejects, exhausted, done, nh;
    int cand[32][4];
    unsigned long long cand_state[32];  // RNG state after each candidate (exact stream rewind)
    int hyp[kMaxHyp][4];
    int counts[kMaxHyp];
    double Hs[kMaxHyp][9];
    double bestH[9];
    double min_median;   // LMEDS
    float fit_thr2;      // squared inlier threshold of the refit (RANSAC: thr^2, LMEDS: sigma^2)
    int have_best;
    // refit / LM
    double red_buf[kWarps * 46];
    double red[46];
    double A[81], V[81], w[9];
    double x[8], xd[8], d[8], v[8], JtJ[64], Ap[64], D[8];
    double norm[8];  // cM.x cM.y cm.x cm.y sM.x sM.y sm.x sm.y
    int n_in, ok;
};

// normalised DLT over the points selected by `use` (all lanes of the CTA); result H (9) in sh.bestH
template <class Use>
__device__ bool dlt_refit(Shared& sh, const float2* __restrict__ src, const float2* __restrict__ dst,
                          int n, Use use) {
    const int tid = threadIdx.x;
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0;
    for (int i = tid; i < n; i += kThreads)
        if (use(i)) {
            acc[0] += src[i].x; acc[1] += src[i].y; acc[2] += dst[i].x; acc[3] += dst[i].y; acc[4] += 1.0;
        }
    block_reduce<9>(acc, sh.red_buf, sh.red);
    const double cnt = sh.red[4];
    if (cnt < 4) return false;
    const double cMx = sh.red[0] / cnt, cMy = sh.red[1] / cnt, cmx = sh.red[2] / cnt, cmy = sh.red[3] / cnt;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0;
    for (int i = tid; i < n; i += kThreads)
        if (use(i)) {
            acc[0] += fabs(src[i].x - cMx); acc[1] += fabs(src[i].y - cMy);
            acc[2] += fabs(dst[i].x - cmx); acc[3] += fabs(dst[i].y - cmy);
        }
    block_reduce<9>(acc, sh.red_buf, sh.red);
    double sMx = sh.red[0], sMy = sh.red[1], smx = sh.red[2], smy = sh.red[3];
    __syncthreads();
    if (fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON || fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON)
        return false;
    sMx = cnt / sMx; sMy = cnt / sMy; smx = cnt / smx; smy = cnt / smy;
    double L[45];
#pragma unroll
    for (int k = 0; k < 45; ++k) L[k] = 0;
    for (int i = tid; i < n; i += kThreads)
        if (use(i)) {
            const double x = (dst[i].x - cmx) * smx, y = (dst[i].y - cmy) * smy;
            const double X = (src[i].x - cMx) * sMx, Y = (src[i].y - cMy) * sMy;
            const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
            const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
            int k = 0;
#pragma unroll
            for (int a = 0; a < 9; ++a)
#pragma unroll
                for (int b = a; b < 9; ++b) L[k++] += Lx[a] * Lx[b] + Ly[a] * Ly[b];
        }
    block_reduce<45>(L, sh.red_buf, sh.red);
    if (tid < 32) {
        if (tid == 0) {
            int k = 0;
            for (int a = 0; a < 9; ++a)
                for (int b = a; b < 9; ++b) {
                    sh.A[a * 9 + b] = sh.red[k];
                    sh.A[b * 9 + a] = sh.red[k];
                    ++k;
                }
        }
        __syncwarp();
        jacobi_eigen_warp(sh.A, sh.V, sh.w, 9);
    }
    if (tid == 0) {
        int mi = 0;
        for (int i = 1; i < 9; ++i)
            if (sh.w[i] < sh.w[mi]) mi = i;
        double H0[9], T[9], H[9];
        for (int i = 0; i < 9; ++i) H0[i] = sh.V[mi * 9 + i];
        const double inv[9] = {1.0 / smx, 0, cmx, 0, 1.0 / smy, cmy, 0, 0, 1};
        const double nrm[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
        mat3_mul(inv, H0, T);
        mat3_mul(T, nrm, H);
        const double s22 = 1.0 / H[8];
        for (int i = 0; i < 9; ++i) sh.bestH[i] = H[i] * s22;
        sh.bestH[8] = 1.0;
    }
    __syncthreads();
    return true;
}

// HomographyRefineCallback::compute at parameters h8 over the used points.
// want_jac: accumulates JtJ (36 upper) + Jtr (8) + S (1) = 45, else only S.
template <class Use>
__device__ void lm_accumulate(Shared& sh, const double* h8, const float2* __restrict__ src,
                              const float2* __restrict__ dst, int n, Use use, bool want_jac) {
    const int tid = threadIdx.x;
    double h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = h8[k];
    if (want_jac) {
        double acc[45];
#pragma unroll
        for (int k = 0; k < 45; ++k) acc[k] = 0;
        for (int i = tid; i < n; i += kThreads)
            if (use(i)) {
                const double Mx = src[i].x, My = src[i].y;
                double ww = h[6] * Mx + h[7] * My + 1.0;
                ww = fabs(ww) > DBL_EPSILON ? 1.0 / ww : 0.0;
                const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
                const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
                const double r0 = xi - dst[i].x, r1 = yi - dst[i].y;
                const double a = Mx * ww, b = My * ww, c = ww;
                const double J0[8] = {a, b, c, 0, 0, 0, -a * xi, -b * xi};
                const double J1[8] = {0, 0, 0, a, b, c, -a * yi, -b * yi};
                int k = 0;
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int q = p; q < 8; ++q) acc[k++] += J0[p] * J0[q] + J1[p] * J1[q];
#pragma unroll
                for (int p = 0; p < 8; ++p) acc[36 + p] += J0[p] * r0 + J1[p] * r1;
                acc[44] += r0 * r0 + r1 * r1;
            }
        block_reduce<45>(acc, sh.red_buf, sh.red);
    } else {
        double acc[1] = {0};
        for (int i = tid; i < n; i += kThreads)
            if (use(i)) {
                const double Mx = src[i].x, My = src[i].y;
                double ww = h[6] * Mx + h[7] * My + 1.0;
                ww = fabs(ww) > DBL_EPSILON ? 1.0 / ww : 0.0;
                const double r0 = (h[0] * Mx + h[1] * My + h[2]) * ww - dst[i].x;
                const double r1 = (h[3] * Mx + h[4] * My + h[5]) * ww - dst[i].y;
                acc[0] += r0 * r0 + r1 * r1;
            }
        block_reduce<1>(acc, sh.red_buf, sh.red);
    }
}

// max |r| over used points at h8 (LM stopping rule norm(r, INF) >= epsf)
template <class Use>
__device__ double lm_max_residual(Shared& sh, const double* h8, const float2* __restrict__ src,
                                  const float2* __restrict__ dst, int n, Use use) {
    double h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = h8[k];
    double m = 0;
    for (int i = threadIdx.x; i < n; i += kThreads)
        if (use(i)) {
            const double Mx = src[i].x, My = src[i].y;
            double ww = h[6] * Mx + h[7] * My + 1.0;
            ww = fabs(ww) > DBL_EPSILON ? 1.0 / ww : 0.0;
            m = fmax(m, fabs((h[0] * Mx + h[1] * My + h[2]) * ww - dst[i].x));
            m = fmax(m, fabs((h[3] * Mx + h[4] * My + h[5]) * ww - dst[i].y));
        }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh.red_buf[threadIdx.x >> 5] = m;
    __syncthreads();
    double r = 0;
    for (int w2 = 0; w2 < kWarps; ++w2) r = fmax(r, sh.red_buf[w2]);
    __syncthreads();
    return r;
}

// LMSolverImpl::run (levmarq.cpp, OpenCV 4.x), maxIters 10, eps = FLT_EPSILON, on sh.bestH[0..7]
template <class Use>
__device__ void lm_refine(Shared& sh, const float2* __restrict__ src, const float2* __restrict__ dst,
                          int n, Use use) {
    const int tid = threadIdx.x;
    __shared__ double S, Sd, lambda, lc, maxd, nu_s;
    __shared__ int proceed, need_inv;
    if (tid == 0) need_inv = 0;
    auto load_normal_eq = [&]() {  // red -> JtJ, v, S   (thread 0)
        int k = 0;
        for (int p = 0; p < 8; ++p)
            for (int q = p; q < 8; ++q) {
                sh.JtJ[p * 8 + q] = sh.red[k];
                sh.JtJ[q * 8 + p] = sh.red[k];
                ++k;
            }
        for (int p = 0; p < 8; ++p) sh.v[p] = sh.red[36 + p];
    };
    if (tid < 8) sh.x[tid] = sh.bestH[tid];
    __syncthreads();
    lm_accumulate(sh, sh.x, src, dst, n, use, true);
    if (tid == 0) {
        load_normal_eq();
        S = sh.red[44];
        for (int p = 0; p < 8; ++p) sh.D[p] = sh.JtJ[p * 8 + p];
        lambda = 1.0;
        lc = 0.75;
    }
    __syncthreads();
    for (int iter = 0; iter < 10;) {
        if (tid < 32) {
            for (int i = tid; i < 64; i += 32) sh.Ap[i] = sh.JtJ[i];
            __syncwarp();
            if (tid < 8) sh.Ap[tid * 8 + tid] += lambda * sh.D[tid];
            __syncwarp();
            jacobi_eigen_warp(sh.Ap, sh.V, sh.w, 8);
        }
        if (tid == 0) {
            eig_backsolve(sh.V, sh.w, sh.v, sh.d, 8);
            double md = 0;
            for (int p = 0; p < 8; ++p) {
                sh.xd[p] = sh.x[p] - sh.d[p];
                md = fmax(md, fabs(sh.d[p]));
            }
            maxd = md;
        }
        __syncthreads();
        lm_accumulate(sh, sh.xd, src, dst, n, use, false);
        if (tid == 0) {
            Sd = sh.red[0];
            double dS = 0, t = 0;
            for (int p = 0; p < 8; ++p) {
                double Ad = 0;
                for (int q = 0; q < 8; ++q) Ad += sh.JtJ[p * 8 + q] * sh.d[q];
                dS += sh.d[p] * (2.0 * sh.v[p] - Ad);
                t += sh.d[p] * sh.v[p];
            }
            const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1.0);
            if (R > 0.75) {
                lambda *= 0.5;
                if (lambda < lc) lambda = 0;
            } else if (R < 0.25) {
                double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1.0) + 2.0;
                nu = fmin(fmax(nu, 2.0), 10.0);
                nu_s = nu;
                if (lambda == 0) need_inv = 1;
                else lambda *= nu;
            }
        }
        __syncthreads();
        if (need_inv) {                                // block-uniform (shared flag)
            if (tid < 32) {
                // invert(A, DECOMP_EIG): diag of V^T diag(1/w) V
                for (int i = tid; i < 64; i += 32) sh.Ap[i] = sh.JtJ[i];
                __syncwarp();
                jacobi_eigen_warp(sh.Ap, sh.V, sh.w, 8);
            }
            if (tid == 0) {
                {
                    double nu = nu_s;
                    double thr = 0;
                    for (int i = 0; i < 8; ++i) thr += fabs(sh.w[i]);
                    thr *= 2 * DBL_EPSILON;
                    double maxval = DBL_EPSILON;
                    for (int i = 0; i < 8; ++i) {
                        double s = 0;
                        for (int k = 0; k < 8; ++k)
                            if (fabs(sh.w[k]) > thr) s += sh.V[k * 8 + i] * sh.V[k * 8 + i] / sh.w[k];
                        maxval = fmax(maxval, fabs(s));
                    }
                    lambda = lc = 1.0 / maxval;
                    nu *= 0.5;
                    lambda *= nu;
                }
                need_inv = 0;
            }
            __syncthreads();
        }
        const bool improved = Sd < S;
        if (improved) {
            if (tid < 8) sh.x[tid] = sh.xd[tid];
            __syncthreads();
            lm_accumulate(sh, sh.x, src, dst, n, use, true);
            if (tid == 0) {
                load_normal_eq();
                S = sh.red[44];
            }
            __syncthreads();
        }
        ++iter;
        const double rmax = lm_max_residual(sh, sh.x, src, dst, n, use);
        if (tid == 0) proceed = iter < 10 && maxd >= (double)FLT_EPSILON && rmax >= (double)FLT_EPSILON;
        __syncthreads();
        if (!proceed) break;
    }
    if (tid < 8) sh.bestH[tid] = sh.x[tid];
    if (tid == 0) sh.bestH[8] = 1.0;
    __syncthreads();
}

// method: 8 = RANSAC, 4 = LMEDS, 0 = all points least squares (+LM)
__global__ void __launch_bounds__(kThreads)
find_homography_kernel(const float2* __restrict__ src_all, const float2* __restrict__ dst_all,
                       const int* __restrict__ starts, const int* __restrict__ counts, int method, float thr, int max_iters,
                       double confidence, double* __restrict__ H_out, uint8_t* __restrict__ mask_out,
                       int* __restrict__ info_out /* [B][4]: found, inliers, iterations, hypotheses */) {
    __shared__ Shared sh;
    const int b = blockIdx.x;
    const int off = starts[b], n = counts[b];
    const float2* src = src_all + off;
    const float2* dst = dst_all + off;
    uint8_t* mask = mask_out ? mask_out + off : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float thr2 = (float)((double)thr * (double)thr);

    auto finish = [&](int found, int inl, int iters, int nhyp) {
        if (tid == 0) {
            info_out[b * 4 + 0] = found; info_out[b * 4 + 1] = inl;
            info_out[b * 4 + 2] = iters; info_out[b * 4 + 3] = nhyp;
        }
        if (tid < 9) H_out[b * 9 + tid] = found ? sh.bestH[tid] : 0.0;
    };

    if (n < 4) {  // the host API rejects this earlier (-28); the pipeline reports "not found"
        for (int i = tid; i < n && mask; i += kThreads) mask[i] = 0;
        finish(0, 0, 0, 0);
        return;
    }
    auto all_points = [](int) { return true; };

    const bool lmeds = method == 4;
    if ((method != 8 && !lmeds) || n == 4) {
        // Default(0): least-squares DLT on all points (+ LM if n > 4); RANSAC with exactly 4 points:
        // single model, mask all ones (ptsetreg.cpp count == modelPoints)
        const bool ok = dlt_refit(sh, src, dst, n, all_points);
        if (ok && n > 4) lm_refine(sh, src, dst, n, all_points);
        for (int i = tid; i < n && mask; i += kThreads) mask[i] = ok ? 1 : 0;
        __syncthreads();
        finish(ok ? 1 : 0, ok ? n : 0, 0, 0);
        return;
    }

    if (tid == 0) {
        sh.rng_state = ~0ull;
        // LMeDSPointSetRegistrator: a fixed number of iterations from an assumed outlier ratio of 0.45, at least 3
        sh.niters = lmeds ? max(update_num_iters(confidence, 0.45, 4, max_iters), 3) : max_iters;
        sh.iter = 0; sh.best_count = 0; sh.rejects = 0;
        sh.exhausted = 0; sh.done = 0; sh.nh = 0;
        sh.min_median = DBL_MAX; sh.have_best = 0; sh.fit_thr2 = thr2;
    }
    __syncthreads();
    int total_hyp = 0;
    DUNK_PHASE(32);

    while (true) {
        // ---- (1) candidate generation + checkSubset, warp 0 ---------------------------------
        if (warp == 0) {
            while (sh.nh < kRoundTarget && !sh.exhausted) {
                if (lane == 0) {
                    Rng rng{sh.rng_state};
                    for (int c = 0; c < 32; ++c) {
                        int idx[4];
                        for (int i = 0; i < 4; ++i) {
                            int v;
                            bool dup;
                            do {
                                v = rng.uniform(n);
                                dup = false;
                                for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
                            } while (dup);
                            idx[i] = v;
                        }
                        sh.cand[c][0] = idx[0]; sh.cand[c][1] = idx[1];
                        sh.cand[c][2] = idx[2]; sh.cand[c][3] = idx[3];
                        sh.cand_state[c] = rng.state;
                    }
                }
                __syncwarp();
                float2 s[4], d[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { s[i] = src[sh.cand[lane][i]]; d[i] = dst[sh.cand[lane][i]]; }
                const bool ok = check_subset(s, d);
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) {
                    int c = 0;
                    for (; c < 32; ++c) {
                        if (bal >> c & 1u) {
                            sh.rejects = 0;
                            const int h = sh.nh++;
                            sh.hyp[h][0] = sh.cand[c][0]; sh.hyp[h][1] = sh.cand[c][1];
                            sh.hyp[h][2] = sh.cand[c][2]; sh.hyp[h][3] = sh.cand[c][3];
                            if (sh.nh == kMaxHyp) { ++c; break; }
                        } else if (++sh.rejects >= (lmeds ? kMaxAttemptsLmeds : kMaxAttempts)) {
                            sh.exhausted = 1; ++c;
                            break;
                        }
                    }
                    // continue the stream right after the last candidate consumed
                    sh.rng_state = sh.cand_state[c - 1];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        const int nh = sh.nh;
        // ---- (2) score: one warp per hypothesis -------------------------------------------
        for (int h = warp; h < nh; h += kWarps) {
            float2 s[4], d[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { s[i] = src[sh.hyp[h][i]]; d[i] = dst[sh.hyp[h][i]]; }
            double H[9];
            const bool ok = homography_4pt(s, d, H);
            const int c = !ok ? -1 : lmeds ? (int)warp_median_error_bits(H, src, dst, n, lane)
                                           : warp_count_inliers(H, src, dst, n, thr2, lane);
            if (lane == 0) sh.counts[h] = c;
            if (lane < 9) sh.Hs[h][lane] = H[lane];
        }
        __syncthreads();
        // ---- (3) sequential accept rule in stream order ------------------------------------
        if (tid == 0) {
            for (int h = 0; h < nh && sh.iter < sh.niters; ++h) {
                const int good = sh.counts[h];
                if (lmeds) {
                    // median < minMedian, strict, in stream order; a failed minimal solve is skipped
                    if (good != -1 && (double)__int_as_float(good) < sh.min_median) {
                        sh.min_median = (double)__int_as_float(good);
                        sh.have_best = 1;
                        for (int i = 0; i < 9; ++i) sh.bestH[i] = sh.Hs[h][i];
                    }
                } else if (good > max(sh.best_count, 3)) {
                    sh.best_count = good;
                    for (int i = 0; i < 9; ++i) sh.bestH[i] = sh.Hs[h][i];
                    sh.niters = update_num_iters(confidence, (double)(n - good) / n, 4, sh.niters);
                }
                ++sh.iter;
            }
            sh.done = (sh.iter >= sh.niters) || sh.exhausted;
            sh.nh = 0;
        }
        total_hyp += nh;
        __syncthreads();
        if (sh.done) break;
    }

    if (lmeds) {
        // sigma = 2.5 * 1.4826 * (1 + 5 / (count - modelPoints)) * sqrt(minMedian), at least 0.001; the model stands
        // if at least 4 points lie within sigma (findInliers: err <= (float)(sigma * sigma))
        if (tid == 0 && sh.have_best) {
            double sigma = 2.5 * 1.4826 * (1.0 + 5.0 / (double)(n - 4)) * sqrt(sh.min_median);
            sigma = fmax(sigma, 0.001);
            sh.fit_thr2 = (float)(sigma * sigma);
        }
        __syncthreads();
        double cnt[1] = {0};
        if (sh.have_best) {
            float Hb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) Hb[i] = (float)sh.bestH[i];
            const float t2 = sh.fit_thr2;
            for (int i = tid; i < n; i += kThreads) cnt[0] += (reproj_err(Hb, src[i], dst[i]) <= t2);
        }
        block_reduce<1>(cnt, sh.red_buf, sh.red);
        if (tid == 0) sh.best_count = sh.have_best && sh.red[0] >= 4.0 ? (int)sh.red[0] : 0;
        __syncthreads();
    }
    if (sh.best_count == 0) {
        for (int i = tid; i < n && mask; i += kThreads) mask[i] = 0;
        __syncthreads();
        finish(0, 0, sh.iter, total_hyp);
        return;
    }
    DUNK_PHASE(33);
    // ---- (4) refit on the inliers of the best minimal model + LM (fundam.cpp) ----------------
    float Hf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) Hf[i] = (float)sh.bestH[i];
    __syncthreads();
    const float fit2 = sh.fit_thr2;
    auto inlier = [&](int i) { return reproj_err(Hf, src[i], dst[i]) <= fit2; };
    dlt_refit(sh, src, dst, n, inlier);   // on failure keeps the minimal-sample model
    DUNK_PHASE(34);
    lm_refine(sh, src, dst, n, inlier);
    DUNK_PHASE(35);
    // ---- (5) returned mask = inliers of the final H (cv2 4.13) ------
    float Hf2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) Hf2[i] = (float)sh.bestH[i];
    double cnt[1] = {0};
    for (int i = tid; i < n; i += kThreads) {
        const bool in = reproj_err(Hf2, src[i], dst[i]) <= thr2;
        if (mask) mask[i] = in;
        cnt[0] += in;
    }
    block_reduce<1>(cnt, sh.red_buf, sh.red);
    finish(1, (int)sh.red[0], sh.iter, total_hyp);
    DUNK_PHASE(36);
}

// parity / diagnostics: score explicit hypothesis sets (one warp each)
__global__ void __launch_bounds__(kThreads)
score_hypotheses_kernel(const float2* __restrict__ src, const float2* __restrict__ dst, int n,
                        const int* __restrict__ samples, int n_hyp, float thr, int* __restrict__ counts,
                        double* __restrict__ Hs) {
    const int h = blockIdx.x * kWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (h >= n_hyp) return;
    const float thr2 = (float)((double)thr * (double)thr);
    float2 s[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[i] = src[samples[h * 4 + i]]; d[i] = dst[samples[h * 4 + i]]; }
    double H[9];
    const bool ok = homography_4pt(s, d, H);
    const int c = ok ? warp_count_inliers(H, src, dst, n, thr2, lane) : -1;
    if (lane == 0) counts[h] = c;
    if (lane < 9) Hs[h * 9 + lane] = ok ? H[lane] : 0.0;
}

}  // namespace
}  // namespace dunk

namespace dunk {
int launch_find_homography(dunk_ctx* ctx, cudaStream_t st, const float2* src, const float2* dst, const int* starts,
                           const int* counts, int n_problems, float thr, double* H, uint8_t* mask, int* info) {
    if (n_problems <= 0) return DUNK_OK;
    find_homography_kernel<<<n_problems, kThreads, 0, st>>>(src, dst, starts, counts, DUNK_H_RANSAC, thr, 2000, 0.995, H,
                                                            mask, info);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    return DUNK_OK;
}
}  // namespace dunk

using namespace dunk;

extern "C" {

int dunk_find_homography_batch(dunk_ctx* ctx, const float* src, const float* dst, const int* offsets,
                               int n_problems, int method, double thr, double* H, uint8_t* mask,
                               int* info) {
    DUNK_REQUIRE(ctx && offsets && H && info && n_problems >= 0, DUNK_ERR_BAD_ARG,
                 "dunk_find_homography_batch: NULL argument");
    if (n_problems == 0) return DUNK_OK;
    DUNK_REQUIRE(method == DUNK_H_RANSAC || method == DUNK_H_LMEDS || method == DUNK_H_DEFAULT || method == DUNK_H_RHO, DUNK_ERR_BAD_ARG,
                 "dunk_find_homography: unknown method %d (Default=0, LMEDS=4, RANSAC=8, RHO=16)", method);
    // HomographyMethod::RHO (mod.rs:25-31): OpenCV's rho.cpp is PROSAC sampling + SPRT verification with its own
    // generator; its result depends on the ORDER of the pairs and is not restated here.  A RHO request is served by the
    // RANSAC estimator with RHO's contract (same threshold, 2000 iterations, confidence 0.995, robust H + inlier mask);
    // parity for this method is by tolerance only (tests/test_ransac_gpu.py::test_rho_vs_cv2), see DESIGN.md section 1.
    if (method == DUNK_H_RHO) method = DUNK_H_RANSAC;
    const int total = offsets[n_problems];
    for (int b = 0; b < n_problems; ++b) {
        const int n = offsets[b + 1] - offsets[b];
        // cv::findHomography: fewer than 4 pairs -> StsVecLengthErr (-28)
        DUNK_REQUIRE(n >= 4, DUNK_ERR_VEC_LENGTH,
                     "dunk_find_homography: problem %d has %d point pairs, at least 4 are needed", b, n);
    }
    DUNK_REQUIRE(src && dst, DUNK_ERR_BAD_ARG, "dunk_find_homography_batch: NULL points");
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    size_t need = 2 * Carver::need((size_t)total * 8) + 2 * Carver::need((size_t)(n_problems + 1) * 4) +
                  Carver::need((size_t)n_problems * 72) + Carver::need((size_t)total) +
                  Carver::need((size_t)n_problems * 16);
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    float2* d_src = cv.take<float2>(total);
    float2* d_dst = cv.take<float2>(total);
    int* d_off = cv.take<int>(n_problems + 1);
    int* d_cnt = cv.take<int>(n_problems + 1);
    double* d_H = cv.take<double>((size_t)n_problems * 9);
    uint8_t* d_mask = cv.take<uint8_t>(total);
    int* d_info = cv.take<int>((size_t)n_problems * 4);
    DUNK_CUDA(cudaMemcpyAsync(d_src, src, (size_t)total * 8, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_dst, dst, (size_t)total * 8, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_off, offsets, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, st));
    std::vector<int> h_cnt(n_problems);
    for (int b = 0; b < n_problems; ++b) h_cnt[b] = offsets[b + 1] - offsets[b];
    DUNK_CUDA(cudaMemcpyAsync(d_cnt, h_cnt.data(), (size_t)n_problems * 4, cudaMemcpyHostToDevice, st));
    find_homography_kernel<<<n_problems, kThreads, 0, st>>>(d_src, d_dst, d_off, d_cnt, method, (float)thr, 2000,
                                                            0.995, d_H, d_mask, d_info);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    DUNK_CUDA(cudaMemcpyAsync(H, d_H, (size_t)n_problems * 72, cudaMemcpyDeviceToHost, st));
    if (mask) DUNK_CUDA(cudaMemcpyAsync(mask, d_mask, (size_t)total, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaMemcpyAsync(info, d_info, (size_t)n_problems * 16, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    return DUNK_OK;
}

int dunk_find_homography(dunk_ctx* ctx, const float* src, const float* dst, int n, int method,
                         double thr, double* H, uint8_t* mask, int* found) {
    DUNK_REQUIRE(found, DUNK_ERR_BAD_ARG, "dunk_find_homography: found is NULL");
    *found = 0;
    DUNK_REQUIRE(n >= 0, DUNK_ERR_BAD_ARG, "dunk_find_homography: n < 0");
    const int offsets[2] = {0, n};
    int info[4] = {0, 0, 0, 0};
    const int rc = dunk_find_homography_batch(ctx, src, dst, offsets, 1, method, thr, H, mask, info);
    if (rc) return rc;
    *found = info[0];
    return DUNK_OK;
}

int dunk_ransac_score_hypotheses(dunk_ctx* ctx, const float* src, const float* dst, int n,
                                 const int* samples, int n_hyp, double thr, int* counts, double* Hs) {
    DUNK_REQUIRE(ctx && src && dst && samples && counts && Hs && n >= 4 && n_hyp >= 0, DUNK_ERR_BAD_ARG,
                 "dunk_ransac_score_hypotheses: bad argument");
    if (n_hyp == 0) return DUNK_OK;
    for (int i = 0; i < n_hyp * 4; ++i)
        DUNK_REQUIRE(samples[i] >= 0 && samples[i] < n, DUNK_ERR_OUT_OF_RANGE,
                     "dunk_ransac_score_hypotheses: sample index %d outside 0..%d", samples[i], n - 1);
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    size_t need = 2 * Carver::need((size_t)n * 8) + Carver::need((size_t)n_hyp * 16) +
                  Carver::need((size_t)n_hyp * 4) + Carver::need((size_t)n_hyp * 72);
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    float2* d_src = cv.take<float2>(n);
    float2* d_dst = cv.take<float2>(n);
    int* d_s = cv.take<int>((size_t)n_hyp * 4);
    int* d_c = cv.take<int>(n_hyp);
    double* d_H = cv.take<double>((size_t)n_hyp * 9);
    DUNK_CUDA(cudaMemcpyAsync(d_src, src, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_dst, dst, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_s, samples, (size_t)n_hyp * 16, cudaMemcpyHostToDevice, st));
    score_hypotheses_kernel<<<div_up(n_hyp, kWarps), kThreads, 0, st>>>(d_src, d_dst, n, d_s, n_hyp, (float)thr,
                                                                        d_c, d_H);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    DUNK_CUDA(cudaMemcpyAsync(counts, d_c, (size_t)n_hyp * 4, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaMemcpyAsync(Hs, d_H, (size_t)n_hyp * 72, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    return DUNK_OK;
}

#ifdef DUNK_PHASE_TIMING
/* timing variant only: clock64() stamps of block 0 (32-36) */
int dunk_debug_phases_homography(long long* out, int n) {
    DUNK_CUDA(cudaDeviceSynchronize());
    DUNK_CUDA(cudaMemcpyFromSymbol(out, g_phase, (size_t)(n < 64 ? n : 64) * 8));
    return DUNK_OK;
}
#endif

}  // extern "C"
