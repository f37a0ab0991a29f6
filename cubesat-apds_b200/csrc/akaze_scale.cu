// AKAZE stage 1a: nonlinear scale space + determinant-of-Hessian response, batched over frames.
// Every kernel is a shared-memory tiled stencil over f32 planes (HBM-bound):
//   gray+Gauss9 | Gauss5+Scharr+|grad| (k-contrast) | 2x area decimation | Gauss5+Scharr+PM-G2 |
//   FED diffusion, K explicit steps per launch temporally blocked in shared memory |
//   fused Hessian (Lx, Ly, Lxx, Lxy, Lyy -> Ldet).
// Follows OpenCV's AKAZEFeatures.cpp / nldiffusion_functions.cpp / fed.cpp as restated in
// oracle/akaze_oracle.py; the FED step keeps OpenCV's f32 operation order
// (no FMA contraction) because 166 dependent steps amplify reassociation noise.
#include "akaze.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace dunk {

namespace {

constexpr int kBX = 32, kBY = 8;           // thread block 32 x 8
constexpr int kTW = 64, kTH = 32;          // generic stencil tile
constexpr int kFedMaxK = 8;                // FED steps fused per launch
constexpr int kNBins = 300;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int reflect101(int v, int n) {
    if (n == 1) return 0;
    while (v < 0 || v >= n) v = v < 0 ? -v : 2 * (n - 1) - v;
    return v;
}

// one reflection is enough when the overshoot is smaller than the extent (callers guarantee it)
__device__ __forceinline__ int reflect101_once(int v, int n) {
    v = v < 0 ? -v : v;
    return v >= n ? 2 * (n - 1) - v : v;
}

struct Gauss9 { float k[5]; };   // k[0] = centre
struct Gauss5 { float k[3]; };

__device__ __forceinline__ float load_gray(const unsigned char* __restrict__ img, int row_stride, int channels,
                                           int x, int y) {
    const unsigned char* p = img + (size_t)y * row_stride + (size_t)x * channels;
    int v;
    if (channels == 1) v = p[0];
    else v = (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15;   // cvtColor BGR(A)2GRAY, 8U
    return __fmul_rn((float)v, (float)(1.0 / 255.0));                    // convertTo(CV_32F, 1/255)
}

// ---- level 0: gray -> GaussianBlur 9x9 sigma 1.6 (BORDER_REPLICATE) -> Lt0 (= Lsmooth0) ----------
__global__ void __launch_bounds__(kBX* kBY)
k_gray_gauss9(const unsigned char* __restrict__ images, size_t image_stride, int row_stride, int channels,
              int W, int H, Gauss9 g, float* __restrict__ Lt0, size_t pyr_stride) {
    __shared__ float A[kTH + 8][kTW + 8];
    __shared__ float T[kTH + 8][kTW];
    const int f = blockIdx.z;
    const unsigned char* img = images + (size_t)f * image_stride;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int ly = ty; ly < kTH + 8; ly += kBY)
        for (int lx = tx; lx < kTW + 8; lx += kBX)
            A[ly][lx] = load_gray(img, row_stride, channels, clampi(x0 + lx - 4, 0, W - 1),
                                  clampi(y0 + ly - 4, 0, H - 1));
    __syncthreads();
    for (int ly = ty; ly < kTH + 8; ly += kBY)
        for (int lx = tx; lx < kTW; lx += kBX) {
            const float* a = &A[ly][lx + 4];
            T[ly][lx] = g.k[0] * a[0] + g.k[1] * (a[-1] + a[1]) + g.k[2] * (a[-2] + a[2]) +
                        g.k[3] * (a[-3] + a[3]) + g.k[4] * (a[-4] + a[4]);
        }
    __syncthreads();
    float* out = Lt0 + (size_t)f * pyr_stride;
    for (int ly = ty; ly < kTH; ly += kBY)
        for (int lx = tx; lx < kTW; lx += kBX) {
            const int gx = x0 + lx, gy = y0 + ly;
            if (gx < W && gy < H) {
                const int c = ly + 4;
                out[(size_t)gy * W + gx] = g.k[0] * T[c][lx] + g.k[1] * (T[c - 1][lx] + T[c + 1][lx]) +
                                           g.k[2] * (T[c - 2][lx] + T[c + 2][lx]) + g.k[3] * (T[c - 3][lx] + T[c + 3][lx]) +
                                           g.k[4] * (T[c - 4][lx] + T[c + 4][lx]);
            }
        }
}

// Gauss5 (replicate) of a clamped-loaded tile A (tile + halo 3) -> B (tile + halo 1), then the
// out-of-image ring of B is filled by reflect-101 (what Scharr's BORDER_DEFAULT sees)
template <int TW, int TH>
__device__ __forceinline__ void gauss5_tile(float (&A)[TH + 6][TW + 6], float (&T)[TH + 6][TW + 2],
                                            float (&B)[TH + 2][TW + 2], const Gauss5& g, int x0, int y0, int W, int H) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int ly = ty; ly < TH + 6; ly += kBY)
        for (int lx = tx; lx < TW + 2; lx += kBX) {
            const float* a = &A[ly][lx + 2];
            T[ly][lx] = g.k[0] * a[0] + g.k[1] * (a[-1] + a[1]) + g.k[2] * (a[-2] + a[2]);
        }
    __syncthreads();
    for (int ly = ty; ly < TH + 2; ly += kBY)
        for (int lx = tx; lx < TW + 2; lx += kBX) {
            const int c = ly + 2;
            B[ly][lx] = g.k[0] * T[c][lx] + g.k[1] * (T[c - 1][lx] + T[c + 1][lx]) + g.k[2] * (T[c - 2][lx] + T[c + 2][lx]);
        }
    __syncthreads();
    // B(lx,ly) sits at global (x0-1+lx, y0-1+ly); cells outside the image take the reflected value
    for (int ly = ty; ly < TH + 2; ly += kBY)
        for (int lx = tx; lx < TW + 2; lx += kBX) {
            const int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
            if (gx < 0 || gx >= W || gy < 0 || gy >= H) {
                const int rx = reflect101(gx, W), ry = reflect101(gy, H);
                const int sx = rx - (x0 - 1), sy = ry - (y0 - 1);
                if (sx >= 0 && sx < TW + 2 && sy >= 0 && sy < TH + 2 && rx >= 0 && ry >= 0) B[ly][lx] = B[sy][sx];
            }
        }
    __syncthreads();
}

__device__ __forceinline__ void scharr_at(const float* up, const float* mid, const float* dn, float& lx, float& ly) {
    // cv::Scharr, scale 1: [-1 0 1] x [3 10 3] (un-normalised)
    lx = 3.f * (up[1] - up[-1]) + 10.f * (mid[1] - mid[-1]) + 3.f * (dn[1] - dn[-1]);
    ly = 3.f * (dn[-1] - up[-1]) + 10.f * (dn[0] - up[0]) + 3.f * (dn[1] - up[1]);
}

// ---- k-contrast pass 1: Gauss5(sigma 1) -> Scharr -> |grad| plane + per-frame max ---------------
__global__ void __launch_bounds__(kBX* kBY)
k_contrast_modg(const unsigned char* __restrict__ images, size_t image_stride, int row_stride, int channels,
                int W, int H, Gauss5 g, float* __restrict__ modg, size_t plane_stride, float* __restrict__ hmax) {
    __shared__ float A[kTH + 6][kTW + 6];
    __shared__ float T[kTH + 6][kTW + 2];
    __shared__ float B[kTH + 2][kTW + 2];
    __shared__ float red[kBX * kBY / 32];
    const int f = blockIdx.z;
    const unsigned char* img = images + (size_t)f * image_stride;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int ly = ty; ly < kTH + 6; ly += kBY)
        for (int lx = tx; lx < kTW + 6; lx += kBX)
            A[ly][lx] = load_gray(img, row_stride, channels, clampi(x0 + lx - 3, 0, W - 1),
                                  clampi(y0 + ly - 3, 0, H - 1));
    __syncthreads();
    gauss5_tile<kTW, kTH>(A, T, B, g, x0, y0, W, H);
    float m = 0.f;
    float* out = modg + (size_t)f * plane_stride;
    for (int ly = ty; ly < kTH; ly += kBY)
        for (int lx = tx; lx < kTW; lx += kBX) {
            const int gx = x0 + lx, gy = y0 + ly;
            if (gx < W && gy < H) {
                float dx, dy;
                scharr_at(&B[ly][lx + 1], &B[ly + 1][lx + 1], &B[ly + 2][lx + 1], dx, dy);
                const float v = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                out[(size_t)gy * W + gx] = v;
                if (gx >= 1 && gx < W - 1 && gy >= 1 && gy < H - 1) m = fmaxf(m, v);   // interior only
            }
        }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const int tid = ty * kBX + tx;
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < kBX * kBY / 32; ++i) m = fmaxf(m, red[i]);
        atomicMax((int*)&hmax[f], __float_as_int(m));   // non-negative floats order like ints
    }
}

// ---- k-contrast pass 2: 300-bin histogram of the interior gradient magnitudes ---------------------
__global__ void __launch_bounds__(256)
k_contrast_hist(const float* __restrict__ modg, size_t plane_stride, int W, int H, const float* __restrict__ hmax,
                int* __restrict__ hist) {
    __shared__ int sh[kNBins];
    const int f = blockIdx.y;
    for (int i = threadIdx.x; i < kNBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const float hm = hmax[f];
    if (hm > 0.f) {
        const float scale = __fdiv_rn((float)(kNBins - 1), hm);
        const float* src = modg + (size_t)f * plane_stride;
        // interior rows 1 .. H-2 are dealt to the blocks round-robin, two rows per pass so that a thread has
        // 2 x ceil((W-2)/256) independent loads in flight (no per-element division)
        for (int y = 1 + 2 * blockIdx.x; y < H - 1; y += 2 * gridDim.x) {
            const float* r0 = src + (size_t)y * W;
            const bool two = y + 1 < H - 1;
            const float* r1 = r0 + (two ? W : 0);
            // the shared atomics order memory, so loads are not hoisted across them: fetch 4 columns x 2 rows first
            // (8 independent loads in flight per thread), then count
            for (int xb = 1 + threadIdx.x; xb < W - 1; xb += 4 * blockDim.x) {
                float v0[4], v1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = xb + u * blockDim.x;
                    const bool in = x < W - 1;
                    v0[u] = in ? r0[x] : -1.f;
                    v1[u] = in ? r1[x] : -1.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (xb + u * (int)blockDim.x >= W - 1) break;
                    int b0 = (int)__fmul_rn(v0[u], scale), b1 = (int)__fmul_rn(v1[u], scale);
                    b0 = min(max(b0, 0), kNBins - 1);
                    b1 = min(max(b1, 0), kNBins - 1);
                    atomicAdd(&sh[b0], 1);
                    if (two) atomicAdd(&sh[b1], 1);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kNBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[f * kNBins + i], sh[i]);
}

// compute_kcontrast tail: 70th percentile of the non-background histogram
__global__ void k_contrast_final(const int* __restrict__ hist, const float* __restrict__ hmax, int W, int H,
                                 float perc, float* __restrict__ kcontrast, int frames) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    const float hm = hmax[f];
    float k = 0.03f;
    if (hm > 0.f) {
        const int total = (W - 2) * (H - 2);
        const int* h = hist + f * kNBins;
        const int nthreshold = (int)((float)(total - h[0]) * perc);
        int nelements = 0;
        for (int b = 1; b < kNBins; ++b) {
            if (nelements >= nthreshold) {
                k = __fdiv_rn(__fmul_rn(hm, (float)b), (float)kNBins);
                break;
            }
            nelements += h[b];
        }
    }
    kcontrast[f] = k;
}

// ---- octave change: cv::resize(INTER_AREA) by ~2 ------------------------------------------------
__global__ void __launch_bounds__(256)
k_halfsample(const float* __restrict__ src, size_t src_stride, int sw, int sh, float* __restrict__ dst,
             size_t dst_stride, int dw, int dh) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    const float* s = src + (size_t)f * src_stride;
    float v;
    if (sw == 2 * dw && sh == 2 * dh) {
        const float* r0 = s + (size_t)(2 * y) * sw + 2 * x;
        const float* r1 = r0 + sw;
        v = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(r0[0], r0[1]), r1[0]), r1[1]), 0.25f);
    } else {
        // general area resize: fractional coverage weights (computeResizeAreaTab)
        const double scx = (double)sw / dw, scy = (double)sh / dh;
        const double fx0 = x * scx, fx1 = fx0 + scx, fy0 = y * scy, fy1 = fy0 + scy;
        const double cellx = fmin(scx, sw - fx0), celly = fmin(scy, sh - fy0);
        const int sx0 = (int)ceil(fx0), sx1 = min((int)floor(fx1), sw);
        const int sy0 = (int)ceil(fy0), sy1 = min((int)floor(fy1), sh);
        float acc = 0.f;
        for (int yy = sy0 - 1; yy <= sy1; ++yy) {
            float wy;
            if (yy == sy0 - 1) wy = (sy0 - fy0 > 1e-3) ? (float)((sy0 - fy0) / celly) : 0.f;
            else if (yy == sy1) wy = (fy1 - sy1 > 1e-3 && sy1 < sh) ? (float)(fmin(fmin(fy1 - sy1, 1.0), celly) / celly) : 0.f;
            else wy = (float)(1.0 / celly);
            if (wy == 0.f || yy < 0 || yy >= sh) continue;
            float racc = 0.f;
            for (int xx = sx0 - 1; xx <= sx1; ++xx) {
                float wx;
                if (xx == sx0 - 1) wx = (sx0 - fx0 > 1e-3) ? (float)((sx0 - fx0) / cellx) : 0.f;
                else if (xx == sx1) wx = (fx1 - sx1 > 1e-3 && sx1 < sw) ? (float)(fmin(fmin(fx1 - sx1, 1.0), cellx) / cellx) : 0.f;
                else wx = (float)(1.0 / cellx);
                if (wx == 0.f || xx < 0 || xx >= sw) continue;
                racc += s[(size_t)yy * sw + xx] * wx;
            }
            acc += racc * wy;
        }
        v = acc;
    }
    dst[(size_t)f * dst_stride + (size_t)y * dw + x] = v;
}

// exact 2:1 decimation, two outputs per thread: two 16-byte loads and one 8-byte store (the one-output kernel above
// kept a single 4-load dependency per thread in flight: long-scoreboard stall ratio 16); same f32 operation order
__global__ void __launch_bounds__(256)
k_halfsample_x2(const float* __restrict__ src, size_t src_stride, int sw, float* __restrict__ dst, size_t dst_stride, int dw, int dh) {
    const int f = blockIdx.z;
    const int xp = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;      // xp = output pair index
    if (2 * xp >= dw || y >= dh) return;
    const float* r0 = src + (size_t)f * src_stride + (size_t)(2 * y) * sw + 4 * xp;
    const float4 a = *reinterpret_cast<const float4*>(r0);
    const float4 b = *reinterpret_cast<const float4*>(r0 + sw);
    const float v0 = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.x, a.y), b.x), b.y), 0.25f);
    const float v1 = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.z, a.w), b.z), b.w), 0.25f);
    *reinterpret_cast<float2*>(dst + (size_t)f * dst_stride + (size_t)y * dw + 2 * xp) = make_float2(v0, v1);
}

// ---- per level: Lsmooth = Gauss5(Lt_init), Lflow = PM-G2(Scharr(Lsmooth), k) --------------------
// No shared memory: a warp owns 32 adjacent columns and walks down the rows.
// Horizontal neighbours come from warp shuffles, vertical neighbours from rolling register windows (5 rows of
// the row-filtered values, 3 rows of the Gauss5 output), so a pixel costs ~6 SHFL instead of ~25 shared-memory
// accesses.  Lanes 3..28 produce outputs (halo 2 for the Gaussian + 1 for Scharr on each side): 26 columns per
// warp, kPrepRows rows per warp (+6 halo rows).  Same arithmetic expressions as gauss5_tile + scharr_at above
// (the keypoint sets stay identical to OpenCV's).
constexpr int kPrepCols = 26;
constexpr int kPrepRows = 64;

__global__ void __launch_bounds__(256)
k_prep_level(const float* __restrict__ Lt, size_t lt_stride, int W, int H, Gauss5 g,
                  const float* __restrict__ kcontrast, float kscale, float* __restrict__ Lsmooth,
                  float* __restrict__ Lflow, size_t plane_stride) {
    const int f = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xs0 = (blockIdx.x * 8 + warp) * kPrepCols;      // first output column of this warp
    if (xs0 >= W) return;
    const int x = xs0 - 3 + lane;                             // this lane's column (may be outside the image)
    const int xc = clampi(x, 0, W - 1);
    const int y0 = blockIdx.y * kPrepRows;
    const float* src = Lt + (size_t)f * lt_stride;
    float* sm = Lsmooth + (size_t)f * plane_stride;
    float* fl = Lflow + (size_t)f * plane_stride;
    const float k = __fmul_rn(kcontrast[f], kscale);
    const float inv_k = __fdiv_rn(1.f, __fmul_rn(k, k));
    const bool out_lane = lane >= 3 && lane < 3 + kPrepCols && x < W;
    const bool left_edge = x - 1 < 0, right_edge = x + 1 >= W;
    const unsigned full = 0xffffffffu;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;          // row-filtered rows yy-4 .. yy
    float ul = 0.f, uc = 0.f, ur = 0.f, ml = 0.f, mc = 0.f, mr = 0.f;   // Gauss5 rows yb-2 (u) and yb-1 (m), with x neighbours
    const int y_end = min(y0 + kPrepRows, H);
    // software pipeline: the loads of the next 4 rows are in flight while the current 4 rows are processed
    // (one load per lane per row would leave ~16 KB in flight per SM; Little's law then caps the kernel at ~2 TB/s)
    constexpr int kAhead = 4;
    float nxt[kAhead];
#pragma unroll
    for (int i = 0; i < kAhead; ++i) nxt[i] = src[(size_t)clampi(y0 - 3 + i, 0, H - 1) * W + xc];
    for (int base = y0 - 3; base <= y_end + 2; base += kAhead) {
        float cur[kAhead];
#pragma unroll
        for (int i = 0; i < kAhead; ++i) cur[i] = nxt[i];
#pragma unroll
        for (int i = 0; i < kAhead; ++i) nxt[i] = src[(size_t)clampi(base + kAhead + i, 0, H - 1) * W + xc];
#pragma unroll
        for (int i = 0; i < kAhead; ++i) {
        const int yy = base + i;
        if (yy > y_end + 2) break;
        const float a = cur[i];
        const float am1 = __shfl_up_sync(full, a, 1), ap1 = __shfl_down_sync(full, a, 1);
        const float am2 = __shfl_up_sync(full, a, 2), ap2 = __shfl_down_sync(full, a, 2);
        t0 = t1; t1 = t2; t2 = t3; t3 = t4;
        t4 = g.k[0] * a + g.k[1] * (am1 + ap1) + g.k[2] * (am2 + ap2);
        const int yb = yy - 2;                                  // Gauss5 row that is complete now
        if (yb < y0 - 1) continue;
        const float b = g.k[0] * t2 + g.k[1] * (t1 + t3) + g.k[2] * (t0 + t4);
        float bl = __shfl_up_sync(full, b, 1), br = __shfl_down_sync(full, b, 1);
        if (left_edge) bl = br;                                 // reflect-101 of the smoothed image (Scharr's border)
        if (right_edge) br = bl;
        const int y = yb - 1;                                   // output row: needs Gauss5 rows y-1 (u), y (m), y+1 (b)
        if (y >= y0 && y < y_end && out_lane) {
            // rows outside the image are replaced by their reflection: row -1 -> row 1, row H -> row H-2
            const float ul_ = y - 1 < 0 ? bl : ul, uc_ = y - 1 < 0 ? b : uc, ur_ = y - 1 < 0 ? br : ur;
            const float dl_ = y + 1 >= H ? ul : bl, dc_ = y + 1 >= H ? uc : b, dr_ = y + 1 >= H ? ur : br;
            const float lx = 3.f * (ur_ - ul_) + 10.f * (mr - ml) + 3.f * (dr_ - dl_);
            const float ly = 3.f * (dl_ - ul_) + 10.f * (dc_ - uc_) + 3.f * (dr_ - ur_);
            const size_t o = (size_t)y * W + x;
            sm[o] = mc;
            fl[o] = __fdiv_rn(1.f, __fadd_rn(1.f, __fmul_rn(inv_k, __fadd_rn(__fmul_rn(lx, lx), __fmul_rn(ly, ly)))));
        }
        ul = ml; uc = mc; ur = mr;
        ml = bl; mc = b; mr = br;
        }
    }
}

// Register-window version of k_prep_level for even W >= 128, H >= 16, in the style of k_hessian_reg: one warp per
// block owns a span of 64 adjacent columns (two per lane, interleaved) and walks down R output rows (+6 halo rows).
// The row loop is unrolled by the window period 5, so the Gauss5 row-filter window (5 rows) and the window of
// smoothed rows with their outer neighbours (3 live rows, 5 slots) are statically indexed registers; a lane's own
// two columns are each other's +-1 neighbours, so a row costs 8 shuffles for 2 pixels.  Outputs: span columns
// 4 .. 59 (halo 2 for the Gaussian + 1 for Scharr, rounded up to keep the float2 accesses aligned).
// Expressions are those of k_prep_level above, term for term.
struct PrepWin {
    float t[2][5];                    // row-filtered input rows
    float b[2][5], bl0[5], br1[5];    // Gauss5 rows: both columns, left neighbour of column 0, right neighbour of column 1
};

// the 5 input rows of one round (walk indices base .. base + 4), rows replicated beyond the image
__device__ __forceinline__ void prep_round_load(float2 (&dst)[5], const float* __restrict__ src, int W, int H, int base,
                                                int r_start, int n_rows, int c0, int ca, int cb, bool interior) {
#pragma unroll
    for (int ph = 0; ph < 5; ++ph) {
        const float* row = src + (size_t)clampi(r_start + min(base + ph, n_rows - 1), 0, H - 1) * W;   // BORDER_REPLICATE
        if (interior) dst[ph] = __ldg((const float2*)(row + c0));
        else dst[ph] = make_float2(row[ca], row[cb]);
    }
}

template <bool FAST>
__device__ __forceinline__ void prep_round(PrepWin& w, const float2 (&cur)[5], float* __restrict__ sm,
                                           float* __restrict__ fl, int W, int H, int base, int r_start, int n_rows,
                                           int y0, int c0, bool col_ok, bool le0, bool re1, const Gauss5& g, float inv_k) {
    const unsigned full = 0xffffffffu;
    // the operation sequences k_prep_level compiles to (checked in its SASS), pinned here with intrinsics
    auto gauss = [&](float c, float s1, float s2) { return __fmaf_rn(g.k[2], s2, __fmaf_rn(g.k[0], c, __fmul_rn(g.k[1], s1))); };
    auto scharr = [](float a, float b, float c) { return __fmaf_rn(3.f, c, __fmaf_rn(10.f, b, __fmul_rn(3.f, a))); };
    const long long oy = (long long)(r_start + base - 3) * W + c0;      // output row y = r - 3 at ph = 0
#pragma unroll
    for (int ph = 0; ph < 5; ++ph) {
        const int p1 = (ph + 4) % 5, p2 = (ph + 3) % 5, p3 = (ph + 2) % 5, p4 = (ph + 1) % 5;   // 1 .. 4 rows earlier
        const float a0 = cur[ph].x, a1 = cur[ph].y;
        {   // Gauss5 along the row: neighbours at +-1 and +-2 columns
            const float a0m1 = __shfl_up_sync(full, a1, 1), a1p1 = __shfl_down_sync(full, a0, 1);
            const float a0m2 = __shfl_up_sync(full, a0, 1), a0p2 = __shfl_down_sync(full, a0, 1);
            const float a1m2 = __shfl_up_sync(full, a1, 1), a1p2 = __shfl_down_sync(full, a1, 1);
            w.t[0][ph] = gauss(a0, __fadd_rn(a0m1, a1), __fadd_rn(a0m2, a0p2));
            w.t[1][ph] = gauss(a1, __fadd_rn(a0, a1p1), __fadd_rn(a1m2, a1p2));
        }
        // Gauss5 row yb = r - 2 is complete: rows yb-2 .. yb+2 sit in slots p4, p3, p2, p1, ph
        const float b0 = gauss(w.t[0][p2], __fadd_rn(w.t[0][p3], w.t[0][p1]), __fadd_rn(w.t[0][p4], w.t[0][ph]));
        const float b1 = gauss(w.t[1][p2], __fadd_rn(w.t[1][p3], w.t[1][p1]), __fadd_rn(w.t[1][p4], w.t[1][ph]));
        float l0 = __shfl_up_sync(full, b1, 1), r1 = __shfl_down_sync(full, b0, 1);
        if (!FAST) {                  // reflect-101 of the smoothed image (Scharr's border)
            if (le0) l0 = b1;
            if (re1) r1 = b0;
        }
        w.b[0][ph] = b0; w.b[1][ph] = b1; w.bl0[ph] = l0; w.br1[ph] = r1;
        // output row y = yb - 1: Gauss5 rows u = y-1 (slot p2), m = y (slot p1), d = y+1 (slot ph)
        const int i = base + ph, y = r_start + i - 3;
        int su = p2, sd = ph;
        float ul0, uc0, ur0, ul1, uc1, ur1, dl0, dc0, dr0, dl1, dc1, dr1;
        ul0 = w.bl0[su]; uc0 = w.b[0][su]; ur0 = w.b[1][su]; ul1 = w.b[0][su]; uc1 = w.b[1][su]; ur1 = w.br1[su];
        dl0 = w.bl0[sd]; dc0 = w.b[0][sd]; dr0 = w.b[1][sd]; dl1 = w.b[0][sd]; dc1 = w.b[1][sd]; dr1 = w.br1[sd];
        if (!FAST) {
            // rows outside the image are replaced by their reflection: row -1 -> row 1, row H -> row H-2
            if (y - 1 < 0) { ul0 = dl0; uc0 = dc0; ur0 = dr0; ul1 = dl1; uc1 = dc1; ur1 = dr1; }
            else if (y + 1 >= H) { dl0 = ul0; dc0 = uc0; dr0 = ur0; dl1 = ul1; dc1 = uc1; dr1 = ur1; }
        }
        const float ml0 = w.bl0[p1], mc0 = w.b[0][p1], mr0 = w.b[1][p1], ml1 = w.b[0][p1], mc1 = w.b[1][p1], mr1 = w.br1[p1];
        const float lx0 = scharr(__fsub_rn(ur0, ul0), __fsub_rn(mr0, ml0), __fsub_rn(dr0, dl0));
        const float ly0 = scharr(__fsub_rn(dl0, ul0), __fsub_rn(dc0, uc0), __fsub_rn(dr0, ur0));
        const float lx1 = scharr(__fsub_rn(ur1, ul1), __fsub_rn(mr1, ml1), __fsub_rn(dr1, dl1));
        const float ly1 = scharr(__fsub_rn(dl1, ul1), __fsub_rn(dc1, uc1), __fsub_rn(dr1, ur1));
        const float f0 = __fdiv_rn(1.f, __fadd_rn(1.f, __fmul_rn(inv_k, __fadd_rn(__fmul_rn(lx0, lx0), __fmul_rn(ly0, ly0)))));
        const float f1 = __fdiv_rn(1.f, __fadd_rn(1.f, __fmul_rn(inv_k, __fadd_rn(__fmul_rn(lx1, lx1), __fmul_rn(ly1, ly1)))));
        const bool ok = FAST ? col_ok : (col_ok && i >= 6 && i < n_rows);
        const long long o = oy + (long long)ph * W;
        if (ok) *(float2*)(sm + o) = make_float2(mc0, mc1);
        if (ok) *(float2*)(fl + o) = make_float2(f0, f1);
    }
}

__global__ void __launch_bounds__(32, 20)
k_prep_level_reg(const float* __restrict__ Lt, size_t lt_stride, int W, int H, int R, Gauss5 g,
                 const float* __restrict__ kcontrast, float kscale, float* __restrict__ Lsmooth,
                 float* __restrict__ Lflow, size_t plane_stride) {
    const int f = blockIdx.z, lane = threadIdx.x;
    const int xs = blockIdx.x * 56 - 4, y0 = blockIdx.y * R;     // span start (even), first output row
    const int c0 = xs + 2 * lane;
    const float* src = Lt + (size_t)f * lt_stride;
    float* sm = Lsmooth + (size_t)f * plane_stride;
    float* fl = Lflow + (size_t)f * plane_stride;
    const float k = __fmul_rn(kcontrast[f], kscale);
    const float inv_k = __fdiv_rn(1.f, __fmul_rn(k, k));
    const int y_end = min(y0 + R, H);
    const int r_start = y0 - 3, n_rows = y_end - y0 + 6;
    const bool interior = xs >= 0 && xs + 64 <= W;
    const bool col_ok = lane >= 2 && lane < 30 && c0 < W;
    const bool le0 = c0 == 0, re1 = c0 + 1 == W - 1;
    const int ca = clampi(c0, 0, W - 1), cb = clampi(c0 + 1, 0, W - 1);
    PrepWin w;
#pragma unroll
    for (int k5 = 0; k5 < 5; ++k5) {
        w.t[0][k5] = w.t[1][k5] = w.b[0][k5] = w.b[1][k5] = w.bl0[k5] = w.br1[k5] = 0.f;
    }
    // software pipeline: the next round's 5 rows are in flight while this round is processed
    float2 nxt[5];
    prep_round_load(nxt, src, W, H, 0, r_start, n_rows, c0, ca, cb, interior);
    for (int base = 0; base < n_rows; base += 5) {
        float2 cur[5];
#pragma unroll
        for (int k5 = 0; k5 < 5; ++k5) cur[k5] = nxt[k5];
        if (base + 5 < n_rows) prep_round_load(nxt, src, W, H, base + 5, r_start, n_rows, c0, ca, cb, interior);
        // outputs exist for walk indices [6, n_rows); FAST additionally needs the rows above / below every output row
        // inside the image and an interior span
        const bool fast = interior && base >= 6 && base + 5 <= n_rows && r_start + base >= 4 && r_start + base + 5 <= H;
        if (fast) prep_round<true>(w, cur, sm, fl, W, H, base, r_start, n_rows, y0, c0, col_ok, le0, re1, g, inv_k);
        else prep_round<false>(w, cur, sm, fl, W, H, base, r_start, n_rows, y0, c0, col_ok, le0, re1, g, inv_k);
    }
}

// ---- FED: K explicit diffusion steps per launch, temporally blocked, register-resident patches ----
// The smem tile is always kFedS x kFedS cells; a launch of K steps produces the inner
// (kFedS - 2K)^2 cells.  Each of the 16 x 16 threads owns a 5 x 5 patch in registers together
// with the conductivity sums on its 60 cell boundaries (constant over the K steps), so a step only
// exchanges patch perimeters through shared memory (ping-pong, one barrier per step).
// Arithmetic is OpenCV's nld_step_scalar, f32 without FMA contraction:
//   Lstep = 0.5 tau (((xpos + xneg) + ypos) + yneg),  xpos = (c + c_E)(L_E - L), ...
// Every boundary flux f = (c_a + c_b)(L_b - L_a) is computed once and used by both cells:
// xneg of the right-hand cell is exactly -f (IEEE negation is exact), so the result is bit-identical
// to the per-pixel formula while doing half the multiplies.
struct FedSteps {
    int k;
    float step[kFedMaxK];   // tau * 0.5
};

constexpr int kFedS = 80;     // smem tile edge = 16 threads x 5 cells
constexpr int kFedP = 5;      // patch edge

__global__ void __launch_bounds__(256, 2)
k_fed(const float* __restrict__ Lin, size_t in_stride, float* __restrict__ Lout, size_t out_stride,
      const float* __restrict__ Lflow, size_t flow_stride, int W, int H, FedSteps fs) {
    extern __shared__ float smem[];
    float* bufA = smem;
    float* bufB = smem + kFedS * kFedS;
    const int K = fs.k;
    const int T = kFedS - 2 * K;
    const int f = blockIdx.z;
    const float* lin = Lin + (size_t)f * in_stride;
    const float* lfl = Lflow + (size_t)f * flow_stride;
    const int x0 = blockIdx.x * T - K, y0 = blockIdx.y * T - K;
    const int tx = threadIdx.x, ty = threadIdx.y;
    // tile load: thread (tx, ty) copies elements (ty + 16k, tx + 16m), k, m < 5: 50 independent 4-byte async
    // copies per thread (the whole 51 KB tile in flight, no registers staged), global and shared addresses are
    // a per-thread base plus compile-time offsets; a warp writes 2 rows x 16 columns = 32 distinct banks
    const bool tile_inside = x0 >= 0 && y0 >= 0 && x0 + kFedS <= W && y0 + kFedS <= H;
    {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(bufA + ty * kFedS + tx);
        const unsigned sb = (unsigned)__cvta_generic_to_shared(bufB + ty * kFedS + tx);
        if (tile_inside) {
            const float* pa = lin + (size_t)(y0 + ty) * W + (x0 + tx);
            const float* pc = lfl + (size_t)(y0 + ty) * W + (x0 + tx);
#pragma unroll
            for (int k = 0; k < kFedP; ++k) {
                const float* ra = pa + (size_t)(16 * k) * W;
                const float* rc = pc + (size_t)(16 * k) * W;
#pragma unroll
                for (int m = 0; m < kFedP; ++m) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa + (16 * k * kFedS + 16 * m) * 4), "l"(ra + 16 * m) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sb + (16 * k * kFedS + 16 * m) * 4), "l"(rc + 16 * m) : "memory");
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < kFedP; ++k)
#pragma unroll
                for (int m = 0; m < kFedP; ++m) {
                    const int gx = x0 + tx + 16 * m, gy = y0 + ty + 16 * k;
                    const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
                    const size_t o = in ? (size_t)gy * W + gx : 0;
                    const unsigned nbytes = in ? 4u : 0u;     // zero-fill outside the image
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa + (16 * k * kFedS + 16 * m) * 4), "l"(lin + o), "r"(nbytes) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sb + (16 * k * kFedS + 16 * m) * 4), "l"(lfl + o), "r"(nbytes) : "memory");
                }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const int px = tx * kFedP, py = ty * kFedP;
    float L[kFedP][kFedP];
    float CE[kFedP][kFedP + 1];   // CE[r][j]: boundary between columns px+j-1 | px+j of row py+r
    float CS[kFedP + 1][kFedP];   // CS[i][c]: boundary between rows py+i-1 | py+i of column px+c
    unsigned corner = 0;          // cells of this patch that are image corners (Lstep forced to 0)
#pragma unroll
    for (int r = 0; r < kFedP; ++r)
#pragma unroll
        for (int c = 0; c < kFedP; ++c) L[r][c] = bufA[(py + r) * kFedS + px + c];
    // block-uniform fast path: the whole tile lies inside the image.  Ring reads are clamped into the tile; a
    // tile-edge boundary then gets some finite conductivity instead of 0, which only touches cells of the outer
    // rings — those are invalid after the first step anyway (the valid region shrinks by one ring per step)
    if (tile_inside) {
        int cx[kFedP + 2], cy[kFedP + 2];
#pragma unroll
        for (int j = 0; j < kFedP + 2; ++j) {
            cx[j] = min(max(px - 1 + j, 0), kFedS - 1);
            cy[j] = min(max(py - 1 + j, 0), kFedS - 1);
        }
#pragma unroll
        for (int r = 0; r < kFedP; ++r) {
            const float* row = bufB + (py + r) * kFedS;
#pragma unroll
            for (int j = 0; j <= kFedP; ++j) CE[r][j] = __fadd_rn(row[cx[j]], row[cx[j + 1]]);
        }
#pragma unroll
        for (int i = 0; i <= kFedP; ++i)
#pragma unroll
            for (int c = 0; c < kFedP; ++c) CS[i][c] = __fadd_rn(bufB[cy[i] * kFedS + px + c], bufB[cy[i + 1] * kFedS + px + c]);
    } else {
        auto in_img = [&](int lx, int ly) {
            const int gx = x0 + lx, gy = y0 + ly;
            return lx >= 0 && lx < kFedS && ly >= 0 && ly < kFedS && gx >= 0 && gx < W && gy >= 0 && gy < H;
        };
#pragma unroll
        for (int r = 0; r < kFedP; ++r)
#pragma unroll
            for (int c = 0; c < kFedP; ++c) {
                const int gx = x0 + px + c, gy = y0 + py + r;
                if ((gy == 0 || gy == H - 1) && (gx == 0 || gx == W - 1)) corner |= 1u << (r * kFedP + c);
            }
#pragma unroll
        for (int r = 0; r < kFedP; ++r)
#pragma unroll
            for (int j = 0; j <= kFedP; ++j) {
                const int cl = px + j - 1, cr = px + j, ly = py + r;
                float v = 0.f;
                if (in_img(cl, ly) && in_img(cr, ly)) v = __fadd_rn(bufB[ly * kFedS + cl], bufB[ly * kFedS + cr]);
                CE[r][j] = v;
            }
#pragma unroll
        for (int i = 0; i <= kFedP; ++i)
#pragma unroll
            for (int c = 0; c < kFedP; ++c) {
                const int lu = py + i - 1, ld = py + i, lx = px + c;
                float v = 0.f;
                if (in_img(lx, lu) && in_img(lx, ld)) v = __fadd_rn(bufB[lu * kFedS + lx], bufB[ld * kFedS + lx]);
                CS[i][c] = v;
            }
    }
    __syncthreads();   // bufB (conductivities) becomes the second ping-pong buffer
    float* cur = bufA;
    float* nxt = bufB;
    // halo indices clamped into the tile: an out-of-tile neighbour always meets a zero conductivity sum
    const int xl = px > 0 ? px - 1 : 0, xr = px + kFedP < kFedS ? px + kFedP : kFedS - 1;
    const int yu = py > 0 ? py - 1 : 0, yd = py + kFedP < kFedS ? py + kFedP : kFedS - 1;
    for (int s = 0; s < K; ++s) {
        const float step = fs.step[s];
        float fE[kFedP][kFedP + 1], fS[kFedP + 1][kFedP];
#pragma unroll
        for (int r = 0; r < kFedP; ++r) {
            const float hl = cur[(py + r) * kFedS + xl], hr = cur[(py + r) * kFedS + xr];
            fE[r][0] = __fmul_rn(CE[r][0], __fsub_rn(L[r][0], hl));
#pragma unroll
            for (int j = 1; j < kFedP; ++j) fE[r][j] = __fmul_rn(CE[r][j], __fsub_rn(L[r][j], L[r][j - 1]));
            fE[r][kFedP] = __fmul_rn(CE[r][kFedP], __fsub_rn(hr, L[r][kFedP - 1]));
        }
#pragma unroll
        for (int c = 0; c < kFedP; ++c) {
            const float hu = cur[yu * kFedS + px + c], hd = cur[yd * kFedS + px + c];
            fS[0][c] = __fmul_rn(CS[0][c], __fsub_rn(L[0][c], hu));
#pragma unroll
            for (int i = 1; i < kFedP; ++i) fS[i][c] = __fmul_rn(CS[i][c], __fsub_rn(L[i][c], L[i - 1][c]));
            fS[kFedP][c] = __fmul_rn(CS[kFedP][c], __fsub_rn(hd, L[kFedP - 1][c]));
        }
        if (corner == 0) {
#pragma unroll
            for (int r = 0; r < kFedP; ++r)
#pragma unroll
                for (int c = 0; c < kFedP; ++c) {
                    const float sum = __fsub_rn(__fadd_rn(__fsub_rn(fE[r][c + 1], fE[r][c]), fS[r + 1][c]), fS[r][c]);
                    L[r][c] = __fadd_rn(L[r][c], __fmul_rn(sum, step));
                }
        } else {
#pragma unroll
            for (int r = 0; r < kFedP; ++r)
#pragma unroll
                for (int c = 0; c < kFedP; ++c) {
                    float sum = __fsub_rn(__fadd_rn(__fsub_rn(fE[r][c + 1], fE[r][c]), fS[r + 1][c]), fS[r][c]);
                    sum = __fmul_rn(sum, step);
                    if (corner >> (r * kFedP + c) & 1u) sum = 0.f;   // the four image corners are written 0
                    L[r][c] = __fadd_rn(L[r][c], sum);
                }
        }
        // publish the patch perimeter for the neighbours' next step
#pragma unroll
        for (int c = 0; c < kFedP; ++c) {
            nxt[py * kFedS + px + c] = L[0][c];
            nxt[(py + kFedP - 1) * kFedS + px + c] = L[kFedP - 1][c];
        }
#pragma unroll
        for (int r = 1; r < kFedP - 1; ++r) {
            nxt[(py + r) * kFedS + px] = L[r][0];
            nxt[(py + r) * kFedS + px + kFedP - 1] = L[r][kFedP - 1];
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
    }
    // full patch -> smem -> coalesced store of the valid inner window
#pragma unroll
    for (int r = 0; r < kFedP; ++r)
#pragma unroll
        for (int c = 0; c < kFedP; ++c) nxt[(py + r) * kFedS + px + c] = L[r][c];
    __syncthreads();
    float* lout = Lout + (size_t)f * out_stride;
    // valid window [K, kFedS - K)^2 -> global, same (ty + 16k, tx + 16m) mapping as the load
#pragma unroll
    for (int k = 0; k < kFedP; ++k) {
        const int ly = ty + 16 * k, gy = y0 + ly;
        if (ly < K || ly >= kFedS - K || gy >= H) continue;
#pragma unroll
        for (int m = 0; m < kFedP; ++m) {
            const int lx = tx + 16 * m, gx = x0 + lx;
            if (lx >= K && lx < kFedS - K && gx < W) lout[(size_t)gy * W + gx] = nxt[ly * kFedS + lx];
        }
    }
}

// ---- level 0 in the register-window style (one warp per block, 64-column span, two columns per lane) --------
// Both kernels read the u8 frame directly (gray conversion + 1/255 in registers) and keep the separable filter's
// row-filtered rows in statically indexed register windows; FP sequences are the ones k_gray_gauss9 /
// k_contrast_modg compile to (checked in their SASS), pinned with intrinsics.
__device__ __forceinline__ float gray_value(int b, int g, int r) {
    return __fmul_rn((float)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15), (float)(1.0 / 255.0));
}

// Raw pixel pairs: the loads of a round are issued back to back and only converted when the round is processed
// (a conversion right behind its load would serialise the load latencies).  MODE 0: gray, 2-byte aligned pair in
// .x; MODE 1: BGRA, 8-byte aligned pair in .x / .y; MODE 2: anything else - the two 8-bit gray values in .x.
template <int MODE>
__device__ __forceinline__ uint2 raw_pair(const unsigned char* __restrict__ img, int row_stride, int channels, int y, int c0) {
    const unsigned char* p = img + (size_t)y * row_stride + (size_t)c0 * channels;
    if (MODE == 0) return make_uint2(__ldg((const unsigned short*)p), 0u);
    if (MODE == 1) return __ldg((const uint2*)p);
    if (channels == 1) return make_uint2((unsigned)p[0] | (unsigned)p[1] << 8, 0u);
    const unsigned v0 = (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15;
    const unsigned v1 = (p[channels] * 3735 + p[channels + 1] * 19235 + p[channels + 2] * 9798 + 16384) >> 15;
    return make_uint2(v0 | v1 << 8, 0u);
}

// border spans: the two columns are clamped independently
__device__ __forceinline__ uint2 raw_pair_clamped(const unsigned char* __restrict__ img, int row_stride, int channels, int y,
                                                  int ca, int cb) {
    const unsigned char* p = img + (size_t)y * row_stride;
    unsigned v0, v1;
    if (channels == 1) {
        v0 = p[ca]; v1 = p[cb];
    } else {
        const unsigned char* a = p + (size_t)ca * channels;
        const unsigned char* b = p + (size_t)cb * channels;
        v0 = (a[0] * 3735 + a[1] * 19235 + a[2] * 9798 + 16384) >> 15;
        v1 = (b[0] * 3735 + b[1] * 19235 + b[2] * 9798 + 16384) >> 15;
    }
    return make_uint2(v0 | v1 << 8, 0u);
}

template <int MODE>
__device__ __forceinline__ float2 raw_to_gray(uint2 q, bool interior) {
    if (MODE == 1 && interior)
        return make_float2(gray_value(q.x & 255, q.x >> 8 & 255, q.x >> 16 & 255), gray_value(q.y & 255, q.y >> 8 & 255, q.y >> 16 & 255));
    return make_float2(__fmul_rn((float)(q.x & 255), (float)(1.0 / 255.0)), __fmul_rn((float)(q.x >> 8 & 255), (float)(1.0 / 255.0)));
}

// the P rows of one round (walk indices base .. base + P - 1), rows replicated beyond the image
template <int P, int MODE>
__device__ __forceinline__ void gray_round_load(uint2 (&dst)[P], const unsigned char* __restrict__ img, int row_stride,
                                                int channels, int H, int base, int r_start, int n_rows, int c0, int ca,
                                                int cb, bool interior) {
#pragma unroll
    for (int ph = 0; ph < P; ++ph) {
        const int y = clampi(r_start + min(base + ph, n_rows - 1), 0, H - 1);
        dst[ph] = interior ? raw_pair<MODE>(img, row_stride, channels, y, c0) : raw_pair_clamped(img, row_stride, channels, y, ca, cb);
    }
}

template <bool FAST, int MODE>
__device__ __forceinline__ void gauss9_round(float (&t)[2][9], const uint2 (&raw)[9], bool interior, float* __restrict__ out,
                                             int W, int base, int r_start, int n_rows, int c0, bool col_ok, const Gauss9& g) {
    const unsigned full = 0xffffffffu;
    auto gauss = [&](float c, float s1, float s2, float s3, float s4) {
        return __fmaf_rn(g.k[4], s4, __fmaf_rn(g.k[3], s3, __fmaf_rn(g.k[2], s2, __fmaf_rn(g.k[0], c, __fmul_rn(g.k[1], s1)))));
    };
    const long long oy = (long long)(r_start + base - 4) * W + c0;      // output row r - 4 at ph = 0
#pragma unroll
    for (int ph = 0; ph < 9; ++ph) {
        const float2 px = raw_to_gray<MODE>(raw[ph], interior);
        const float a0 = px.x, a1 = px.y;
        const float u1a0 = __shfl_up_sync(full, a0, 1), u1a1 = __shfl_up_sync(full, a1, 1);
        const float d1a0 = __shfl_down_sync(full, a0, 1), d1a1 = __shfl_down_sync(full, a1, 1);
        const float u2a0 = __shfl_up_sync(full, a0, 2), u2a1 = __shfl_up_sync(full, a1, 2);
        const float d2a0 = __shfl_down_sync(full, a0, 2), d2a1 = __shfl_down_sync(full, a1, 2);
        // column 0 (span index 2t): -1 = u1a1, +1 = a1, -2 = u1a0, +2 = d1a0, -3 = u2a1, +3 = d1a1, -4 = u2a0, +4 = d2a0
        t[0][ph] = gauss(a0, __fadd_rn(u1a1, a1), __fadd_rn(u1a0, d1a0), __fadd_rn(u2a1, d1a1), __fadd_rn(u2a0, d2a0));
        // column 1 (2t + 1): -1 = a0, +1 = d1a0, -2 = u1a1, +2 = d1a1, -3 = u1a0, +3 = d2a0, -4 = u2a1, +4 = d2a1
        t[1][ph] = gauss(a1, __fadd_rn(a0, d1a0), __fadd_rn(u1a1, d1a1), __fadd_rn(u1a0, d2a0), __fadd_rn(u2a1, d2a1));
        // output row y = r - 4: row-filtered rows y-4 .. y+4 sit in slots ph+1 .. ph (mod 9)
        const int m4 = (ph + 1) % 9, m3 = (ph + 2) % 9, m2 = (ph + 3) % 9, m1 = (ph + 4) % 9, c = (ph + 5) % 9, p1 = (ph + 6) % 9,
                  p2 = (ph + 7) % 9, p3 = (ph + 8) % 9;
        const float o0 = gauss(t[0][c], __fadd_rn(t[0][m1], t[0][p1]), __fadd_rn(t[0][m2], t[0][p2]), __fadd_rn(t[0][m3], t[0][p3]),
                               __fadd_rn(t[0][m4], t[0][ph]));
        const float o1 = gauss(t[1][c], __fadd_rn(t[1][m1], t[1][p1]), __fadd_rn(t[1][m2], t[1][p2]), __fadd_rn(t[1][m3], t[1][p3]),
                               __fadd_rn(t[1][m4], t[1][ph]));
        const int i = base + ph;
        if (FAST ? col_ok : (col_ok && i >= 8 && i < n_rows)) *(float2*)(out + oy + (long long)ph * W) = make_float2(o0, o1);
    }
}

template <int MODE>
__global__ void __launch_bounds__(32, 24)
k_gray_gauss9_reg(const unsigned char* __restrict__ images, size_t image_stride, int row_stride, int channels,
                  int W, int H, int R, Gauss9 g, float* __restrict__ Lt0, size_t pyr_stride) {
    const int f = blockIdx.z, lane = threadIdx.x;
    const int xs = blockIdx.x * 56 - 4, y0 = blockIdx.y * R;     // span start (even), first output row
    const int c0 = xs + 2 * lane;
    const unsigned char* img = images + (size_t)f * image_stride;
    float* out = Lt0 + (size_t)f * pyr_stride;
    const int y_end = min(y0 + R, H);
    const int r_start = y0 - 4, n_rows = y_end - y0 + 8;
    const bool interior = xs >= 0 && xs + 64 <= W;
    const bool col_ok = lane >= 2 && lane < 30 && c0 < W;
    const int ca = clampi(c0, 0, W - 1), cb = clampi(c0 + 1, 0, W - 1);
    float t[2][9];
#pragma unroll
    for (int k = 0; k < 9; ++k) t[0][k] = t[1][k] = 0.f;
    // software pipeline: the next round's 9 rows are in flight while this round is filtered
    uint2 nxt[9];
    gray_round_load<9, MODE>(nxt, img, row_stride, channels, H, 0, r_start, n_rows, c0, ca, cb, interior);
    for (int base = 0; base < n_rows; base += 9) {
        uint2 cur[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) cur[k] = nxt[k];
        if (base + 9 < n_rows) gray_round_load<9, MODE>(nxt, img, row_stride, channels, H, base + 9, r_start, n_rows, c0, ca, cb, interior);
        const bool fast = interior && base >= 8 && base + 9 <= n_rows;
        if (fast) gauss9_round<true, MODE>(t, cur, interior, out, W, base, r_start, n_rows, c0, col_ok, g);
        else gauss9_round<false, MODE>(t, cur, interior, out, W, base, r_start, n_rows, c0, col_ok, g);
    }
}

// k-contrast pass 1: Gauss5(sigma 1) -> Scharr -> |grad| plane + per-frame max; structure of k_prep_level_reg
template <bool FAST, int MODE>
__device__ __forceinline__ void modg_round(PrepWin& w, float& mx, const uint2 (&raw)[5], bool interior, float* __restrict__ out,
                                           int W, int H, int base, int r_start, int n_rows, int c0, bool col_ok, bool le0,
                                           bool re1, const Gauss5& g) {
    const unsigned full = 0xffffffffu;
    auto gauss = [&](float c, float s1, float s2) { return __fmaf_rn(g.k[2], s2, __fmaf_rn(g.k[0], c, __fmul_rn(g.k[1], s1))); };
    // k_contrast_modg contracts Scharr differently from k_prep_level: the multiply sits on the middle term
    auto scharr = [](float a, float b, float c) { return __fmaf_rn(3.f, c, __fmaf_rn(3.f, a, __fmul_rn(10.f, b))); };
    const long long oy = (long long)(r_start + base - 3) * W + c0;      // output row y = r - 3 at ph = 0
#pragma unroll
    for (int ph = 0; ph < 5; ++ph) {
        const int p1 = (ph + 4) % 5, p2 = (ph + 3) % 5, p3 = (ph + 2) % 5, p4 = (ph + 1) % 5;   // 1 .. 4 rows earlier
        const float2 px = raw_to_gray<MODE>(raw[ph], interior);
        const float a0 = px.x, a1 = px.y;
        {
            const float a0m1 = __shfl_up_sync(full, a1, 1), a1p1 = __shfl_down_sync(full, a0, 1);
            const float a0m2 = __shfl_up_sync(full, a0, 1), a1p2 = __shfl_down_sync(full, a1, 1);
            w.t[0][ph] = gauss(a0, __fadd_rn(a0m1, a1), __fadd_rn(a0m2, a1p1));
            w.t[1][ph] = gauss(a1, __fadd_rn(a0, a1p1), __fadd_rn(a0m1, a1p2));
        }
        const float b0 = gauss(w.t[0][p2], __fadd_rn(w.t[0][p3], w.t[0][p1]), __fadd_rn(w.t[0][p4], w.t[0][ph]));
        const float b1 = gauss(w.t[1][p2], __fadd_rn(w.t[1][p3], w.t[1][p1]), __fadd_rn(w.t[1][p4], w.t[1][ph]));
        float l0 = __shfl_up_sync(full, b1, 1), r1 = __shfl_down_sync(full, b0, 1);
        if (!FAST) {                  // reflect-101 of the smoothed image (Scharr's border)
            if (le0) l0 = b1;
            if (re1) r1 = b0;
        }
        w.b[0][ph] = b0; w.b[1][ph] = b1; w.bl0[ph] = l0; w.br1[ph] = r1;
        const int i = base + ph, y = r_start + i - 3;
        const int su = p2, sd = ph;
        float ul0 = w.bl0[su], uc0 = w.b[0][su], ur0 = w.b[1][su], ul1 = w.b[0][su], uc1 = w.b[1][su], ur1 = w.br1[su];
        float dl0 = w.bl0[sd], dc0 = w.b[0][sd], dr0 = w.b[1][sd], dl1 = w.b[0][sd], dc1 = w.b[1][sd], dr1 = w.br1[sd];
        if (!FAST) {
            if (y - 1 < 0) { ul0 = dl0; uc0 = dc0; ur0 = dr0; ul1 = dl1; uc1 = dc1; ur1 = dr1; }
            else if (y + 1 >= H) { dl0 = ul0; dc0 = uc0; dr0 = ur0; dl1 = ul1; dc1 = uc1; dr1 = ur1; }
        }
        const float ml0 = w.bl0[p1], mr0 = w.b[1][p1], ml1 = w.b[0][p1], mr1 = w.br1[p1];
        const float lx0 = scharr(__fsub_rn(ur0, ul0), __fsub_rn(mr0, ml0), __fsub_rn(dr0, dl0));
        const float ly0 = scharr(__fsub_rn(dl0, ul0), __fsub_rn(dc0, uc0), __fsub_rn(dr0, ur0));
        const float lx1 = scharr(__fsub_rn(ur1, ul1), __fsub_rn(mr1, ml1), __fsub_rn(dr1, dl1));
        const float ly1 = scharr(__fsub_rn(dl1, ul1), __fsub_rn(dc1, uc1), __fsub_rn(dr1, ur1));
        const float v0 = sqrtf(__fadd_rn(__fmul_rn(lx0, lx0), __fmul_rn(ly0, ly0)));
        const float v1 = sqrtf(__fadd_rn(__fmul_rn(lx1, lx1), __fmul_rn(ly1, ly1)));
        const bool ok = FAST ? col_ok : (col_ok && i >= 6 && i < n_rows);
        if (ok) *(float2*)(out + oy + (long long)ph * W) = make_float2(v0, v1);
        // maximum over the interior pixels only (1 <= x < W-1, 1 <= y < H-1)
        if (FAST) {
            if (ok) mx = fmaxf(mx, fmaxf(v0, v1));
        } else {
            const bool yin = y >= 1 && y < H - 1;
            if (ok && yin && c0 >= 1 && c0 < W - 1) mx = fmaxf(mx, v0);
            if (ok && yin && c0 + 1 >= 1 && c0 + 1 < W - 1) mx = fmaxf(mx, v1);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(32, 20)
k_contrast_modg_reg(const unsigned char* __restrict__ images, size_t image_stride, int row_stride, int channels,
                    int W, int H, int R, Gauss5 g, float* __restrict__ modg, size_t plane_stride, float* __restrict__ hmax) {
    const int f = blockIdx.z, lane = threadIdx.x;
    const int xs = blockIdx.x * 56 - 4, y0 = blockIdx.y * R;
    const int c0 = xs + 2 * lane;
    const unsigned char* img = images + (size_t)f * image_stride;
    float* out = modg + (size_t)f * plane_stride;
    const int y_end = min(y0 + R, H);
    const int r_start = y0 - 3, n_rows = y_end - y0 + 6;
    const bool interior = xs >= 0 && xs + 64 <= W;
    const bool col_ok = lane >= 2 && lane < 30 && c0 < W;
    const bool le0 = c0 == 0, re1 = c0 + 1 == W - 1;
    const int ca = clampi(c0, 0, W - 1), cb = clampi(c0 + 1, 0, W - 1);
    PrepWin w;
#pragma unroll
    for (int k5 = 0; k5 < 5; ++k5) w.t[0][k5] = w.t[1][k5] = w.b[0][k5] = w.b[1][k5] = w.bl0[k5] = w.br1[k5] = 0.f;
    float mx = 0.f;
    // software pipeline: the next round's 5 rows are in flight while this round is processed
    uint2 nxt[5];
    gray_round_load<5, MODE>(nxt, img, row_stride, channels, H, 0, r_start, n_rows, c0, ca, cb, interior);
    for (int base = 0; base < n_rows; base += 5) {
        uint2 cur[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) cur[k] = nxt[k];
        if (base + 5 < n_rows) gray_round_load<5, MODE>(nxt, img, row_stride, channels, H, base + 5, r_start, n_rows, c0, ca, cb, interior);
        // FAST: every output of the round is stored, the rows around every output row lie strictly inside the image
        // (so the outputs are interior pixels of the maximum too: y >= 1, y + 1 < H) and the span is interior
        const bool fast = interior && base >= 6 && base + 5 <= n_rows && r_start + base >= 4 && r_start + base + 5 <= H;
        if (fast) modg_round<true, MODE>(w, mx, cur, interior, out, W, H, base, r_start, n_rows, c0, col_ok, le0, re1, g);
        else modg_round<false, MODE>(w, mx, cur, interior, out, W, H, base, r_start, n_rows, c0, col_ok, le0, re1, g);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx > 0.f) atomicMax((int*)&hmax[f], __float_as_int(mx));   // non-negative floats order like ints
}

// ---- FED without shared memory: a cascade of K stages in registers ---------------------------------
// One warp per block owns a span of 64 adjacent columns (two per lane, interleaved) and walks down the rows.
// Stage s (1..K) turns rows of the field after s-1 steps into rows after s steps: when the stage receives row
// rho it forms the horizontal fluxes of that row, the vertical flux between rho-1 and rho, and emits row rho-1,
// which is the next stage's input in the same walk step - so stage s trails the loads by s rows and keeps only
// its previous row (value, horizontal flux difference, vertical flux above): 6 registers per stage.  The
// conductivity sums of a row are formed once when the Lflow row is loaded and kept in a K-deep delay line
// (statically indexed: the walk loop is unrolled by K), since stage s needs the sums of the row s-1 steps back.
// Horizontal neighbours: one shuffle for the left value, one for the flux through the right boundary (the next
// lane's left flux), i.e. every boundary flux is computed exactly once.  The span loses Kh = K rounded up to even
// columns either side (garbage creeps in one column per step from the span edges), the walk K rows above and
// below.  Same operations as k_fed: CE = c_left + c_right, CS = c_up + c_down (0 across the image border),
// flux = C * (L_b - L_a), L += (((fE_r - fE_l) + fS_d) - fS_u) * step, image corners unchanged.
template <int K>
struct FedWin {
    float L[K][2], hs[K][2], fv[K][2];     // per stage: previous input row, its fE_r - fE_l, the flux above it
    float ce0[K], ce1[K], cs0[K], cs1[K];  // delay lines: CE(left|c0), CE(c0|c1), CS(row-1|row) for both columns
    float cp0, cp1;                        // previous conductivity row
};

template <int K, int ROUND, bool FAST>
__device__ __forceinline__ void fed_round(FedWin<K>& w, const float* __restrict__ lin, const float* __restrict__ lfl,
                                          float* __restrict__ lout, int W, int H, int base, int r_start, int n_rows,
                                          int y0, int y_end, int c0, int ca, int cb, bool interior, bool col_ok,
                                          const FedSteps& fs) {
    const unsigned full = 0xffffffffu;
    float2 cl[ROUND], cc[ROUND];
    if (FAST) {
        const size_t o = (size_t)(r_start + base) * W + c0;
#pragma unroll
        for (int ph = 0; ph < ROUND; ++ph) {
            cl[ph] = __ldg((const float2*)(lin + o + (size_t)ph * W));
            cc[ph] = __ldg((const float2*)(lfl + o + (size_t)ph * W));
        }
    } else {
#pragma unroll
        for (int ph = 0; ph < ROUND; ++ph) {
            const size_t ro = (size_t)clampi(r_start + min(base + ph, n_rows - 1), 0, H - 1) * W;
            if (interior) {
                cl[ph] = __ldg((const float2*)(lin + ro + c0));
                cc[ph] = __ldg((const float2*)(lfl + ro + c0));
            } else {
                cl[ph] = make_float2(lin[ro + ca], lin[ro + cb]);
                cc[ph] = make_float2(lfl[ro + ca], lfl[ro + cb]);
            }
        }
    }
    const long long oy = (long long)(r_start + base - K) * W + c0;      // output row r - K at ph = 0
#pragma unroll
    for (int ph = 0; ph < ROUND; ++ph) {
        const int slot = ph % K;                                        // base is a multiple of K
        const int i = base + ph, r = r_start + i;
        {   // conductivity sums of row r
            const float c0v = cc[ph].x, c1v = cc[ph].y;
            const float cleft = __shfl_up_sync(full, c1v, 1);
            float e0 = __fadd_rn(cleft, c0v), e1 = __fadd_rn(c0v, c1v);
            float s0 = __fadd_rn(w.cp0, c0v), s1 = __fadd_rn(w.cp1, c1v);
            if (!FAST) {
                const bool row_in = r >= 0 && r < H, up_in = r - 1 >= 0 && r < H;
                const bool in0 = c0 >= 0 && c0 < W, in1 = c0 + 1 >= 0 && c0 + 1 < W;
                if (!(row_in && in0 && c0 - 1 >= 0)) e0 = 0.f;
                if (!(row_in && in0 && in1)) e1 = 0.f;
                if (!(up_in && in0)) s0 = 0.f;
                if (!(up_in && in1)) s1 = 0.f;
            }
            w.ce0[slot] = e0; w.ce1[slot] = e1; w.cs0[slot] = s0; w.cs1[slot] = s1;
            w.cp0 = c0v; w.cp1 = c1v;
        }
        float in0 = cl[ph].x, in1 = cl[ph].y;
#pragma unroll
        for (int s = 0; s < K; ++s) {                                   // stage s + 1: input row rho = r - s
            const int sl = (slot + K - s) % K;
            const float left = __shfl_up_sync(full, in1, 1);
            const float f01 = __fmul_rn(w.ce1[sl], __fsub_rn(in1, in0));
            const float fl0 = __fmul_rn(w.ce0[sl], __fsub_rn(in0, left));
            const float f1r = __shfl_down_sync(full, fl0, 1);
            const float h0 = __fsub_rn(f01, fl0), h1 = __fsub_rn(f1r, f01);
            const float v0 = __fmul_rn(w.cs0[sl], __fsub_rn(in0, w.L[s][0]));
            const float v1 = __fmul_rn(w.cs1[sl], __fsub_rn(in1, w.L[s][1]));
            float d0 = __fmul_rn(__fsub_rn(__fadd_rn(w.hs[s][0], v0), w.fv[s][0]), fs.step[s]);
            float d1 = __fmul_rn(__fsub_rn(__fadd_rn(w.hs[s][1], v1), w.fv[s][1]), fs.step[s]);
            if (!FAST) {                                                // the four image corners are left unchanged
                const int ro = r - s - 1;
                if ((ro == 0 || ro == H - 1) && c0 == 0) d0 = 0.f;
                if ((ro == 0 || ro == H - 1) && c0 + 1 == W - 1) d1 = 0.f;
            }
            const float o0 = __fadd_rn(w.L[s][0], d0), o1 = __fadd_rn(w.L[s][1], d1);
            w.L[s][0] = in0; w.L[s][1] = in1; w.hs[s][0] = h0; w.hs[s][1] = h1; w.fv[s][0] = v0; w.fv[s][1] = v1;
            in0 = o0; in1 = o1;
        }
        const int y = r - K;
        if (FAST ? col_ok : (col_ok && y >= y0 && y < y_end && i < n_rows))
            *(float2*)(lout + oy + (long long)ph * W) = make_float2(in0, in1);
    }
}

template <int K>
__global__ void __launch_bounds__(32, K >= 7 ? 12 : (K >= 5 ? 16 : 20))
k_fed_reg(const float* __restrict__ Lin, size_t in_stride, float* __restrict__ Lout, size_t out_stride,
          const float* __restrict__ Lflow, size_t flow_stride, int W, int H, int R, FedSteps fs) {
    constexpr int KH = (K + 1) / 2 * 2;                                 // column halo, even
    constexpr int OUTW = 64 - 2 * KH;
    constexpr int ROUND = K * (K <= 2 ? 4 : (K <= 4 ? 2 : 1));          // rows loaded per round, a multiple of K
    const int f = blockIdx.z, lane = threadIdx.x;
    const int xs = blockIdx.x * OUTW - KH, y0 = blockIdx.y * R;
    const int c0 = xs + 2 * lane;
    const float* lin = Lin + (size_t)f * in_stride;
    const float* lfl = Lflow + (size_t)f * flow_stride;
    float* lout = Lout + (size_t)f * out_stride;
    const int y_end = min(y0 + R, H);
    const int r_start = y0 - K, n_rows = y_end - y0 + 2 * K;
    const bool interior = xs >= 0 && xs + 64 <= W;
    const bool col_ok = lane >= KH / 2 && lane < 32 - KH / 2 && c0 < W;
    const int ca = clampi(c0, 0, W - 1), cb = clampi(c0 + 1, 0, W - 1);
    FedWin<K> w;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        w.L[s][0] = w.L[s][1] = w.hs[s][0] = w.hs[s][1] = w.fv[s][0] = w.fv[s][1] = 0.f;
        w.ce0[s] = w.ce1[s] = w.cs0[s] = w.cs1[s] = 0.f;
    }
    w.cp0 = w.cp1 = 0.f;
    for (int base = 0; base < n_rows; base += ROUND) {
        // outputs exist for walk indices [2K, n_rows); FAST also needs every input row of the round, the row above
        // the first one, and all stage rows inside the image
        const bool fast = interior && base >= 2 * K && base + ROUND <= n_rows && r_start + base - K >= 1 && r_start + base + ROUND <= H;
        if (fast) fed_round<K, ROUND, true>(w, lin, lfl, lout, W, H, base, r_start, n_rows, y0, y_end, c0, ca, cb, interior, col_ok, fs);
        else fed_round<K, ROUND, false>(w, lin, lfl, lout, W, H, base, r_start, n_rows, y0, y_end, c0, ca, cb, interior, col_ok, fs);
    }
}

// ---- fused Hessian: Lsmooth -> Lx, Ly (kept for orientation/descriptor) and Ldet ------------------
// kernels of compute_derivative_kernels(scale s): taps at -s, 0, +s: smoothing [w0, w1, w0],
// derivative [-1, 0, 1]; sepFilter2D = row pass (kx) then column pass (ky), BORDER_REFLECT_101.
__global__ void __launch_bounds__(kBX* kBY)
k_hessian_generic(const float* __restrict__ Lsm, size_t sm_stride, int W, int H, int s, float w0, float w1,
          float sigma_quat, float* __restrict__ Lx, float* __restrict__ Ly, float* __restrict__ Ldet,
          size_t pyr_stride) {
    extern __shared__ float smem[];
    const int SW = kTW + 4 * s, SH = kTH + 4 * s;      // Lsmooth region (halo 2s)
    const int DW = kTW + 2 * s, DH = kTH + 2 * s;      // first-derivative region (halo s)
    float* S0 = smem;                 // SH x SW
    float* DX = S0 + SW * SH;         // DH x DW
    float* DY = DX + DW * DH;         // DH x DW
    const int f = blockIdx.z;
    const float* src = Lsm + (size_t)f * sm_stride;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int ly = ty; ly < SH; ly += kBY)
        for (int lx = tx; lx < SW; lx += kBX)
            S0[ly * SW + lx] = src[(size_t)reflect101(y0 - 2 * s + ly, H) * W + reflect101(x0 - 2 * s + lx, W)];
    __syncthreads();
    for (int ly = ty; ly < DH; ly += kBY)
        for (int lx = tx; lx < DW; lx += kBX) {
            const float* c = &S0[(ly + s) * SW + (lx + s)];
            const float* u = c - s * SW;
            const float* d = c + s * SW;
            // Lx: rows [-1 0 1], columns [w0 w1 w0]
            DX[ly * DW + lx] = w0 * (u[s] - u[-s]) + w1 * (c[s] - c[-s]) + w0 * (d[s] - d[-s]);
            // Ly: rows [w0 w1 w0], columns [-1 0 1]
            DY[ly * DW + lx] = (w0 * d[-s] + w1 * d[0] + w0 * d[s]) - (w0 * u[-s] + w1 * u[0] + w0 * u[s]);
        }
    __syncthreads();
    float* ox = Lx + (size_t)f * pyr_stride;
    float* oy = Ly + (size_t)f * pyr_stride;
    float* od = Ldet + (size_t)f * pyr_stride;
    for (int ly = ty; ly < kTH; ly += kBY)
        for (int lx = tx; lx < kTW; lx += kBX) {
            const int gx = x0 + lx, gy = y0 + ly;
            if (gx >= W || gy >= H) continue;
            const float* cx = &DX[(ly + s) * DW + (lx + s)];
            const float* cy = &DY[(ly + s) * DW + (lx + s)];
            const float* ux = cx - s * DW;
            const float* dx = cx + s * DW;
            const float* uy = cy - s * DW;
            const float* dy = cy + s * DW;
            const float lxx = w0 * (ux[s] - ux[-s]) + w1 * (cx[s] - cx[-s]) + w0 * (dx[s] - dx[-s]);
            const float lxy = (w0 * dx[-s] + w1 * dx[0] + w0 * dx[s]) - (w0 * ux[-s] + w1 * ux[0] + w0 * ux[s]);
            const float lyy = (w0 * dy[-s] + w1 * dy[0] + w0 * dy[s]) - (w0 * uy[-s] + w1 * uy[0] + w0 * uy[s]);
            const size_t o = (size_t)gy * W + gx;
            ox[o] = cx[0];
            oy[o] = cy[0];
            od[o] = __fmul_rn(__fsub_rn(__fmul_rn(lxx, lyy), __fmul_rn(lxy, lxy)), sigma_quat);
        }
}

// The same kernel with the tap distance S as a template parameter (AKAZE's sublevels always give
// sigma_size 2, 3, 3, 4): every extent and tap offset is a compile-time constant, the three stages are
// flat 256-thread loops, and tiles away from the border skip the reflect-101 index work.  The
// arithmetic expressions are those of k_hessian_generic.
template <int S>
__global__ void __launch_bounds__(kBX* kBY)
k_hessian(const float* __restrict__ Lsm, size_t sm_stride, int W, int H, float w0, float w1,
          float sigma_quat, float* __restrict__ Lx, float* __restrict__ Ly, float* __restrict__ Ldet,
          size_t pyr_stride) {
    constexpr int SW = kTW + 4 * S, SH = kTH + 4 * S;      // Lsmooth region (halo 2S)
    constexpr int DW = kTW + 2 * S, DH = kTH + 2 * S;      // first-derivative region (halo S)
    extern __shared__ float smem[];
    float* S0 = smem;                 // SH x SW
    float* DX = S0 + SW * SH;         // DH x DW
    float* DY = DX + DW * DH;         // DH x DW
    const int f = blockIdx.z;
    const float* src = Lsm + (size_t)f * sm_stride;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const int tid = threadIdx.y * kBX + threadIdx.x;
    if (x0 >= 2 * S && y0 >= 2 * S && x0 + kTW + 2 * S <= W && y0 + kTH + 2 * S <= H) {
        const float* base = src + (size_t)(y0 - 2 * S) * W + (x0 - 2 * S);
#pragma unroll
        for (int it = 0; it < (SH * SW + 255) / 256; ++it) {
            const int i = tid + it * 256;
            if (i < SH * SW) {
                const int ly = i / SW, lx = i - ly * SW;
                S0[i] = base[(size_t)ly * W + lx];
            }
        }
    } else {
        for (int i = tid; i < SH * SW; i += 256) {
            const int ly = i / SW, lx = i - ly * SW;
            S0[i] = src[(size_t)reflect101(y0 - 2 * S + ly, H) * W + reflect101(x0 - 2 * S + lx, W)];
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < (DH * DW + 255) / 256; ++it) {
        const int i = tid + it * 256;
        if (i < DH * DW) {
            const int ly = i / DW, lx = i - ly * DW;
            const float* c = &S0[(ly + S) * SW + (lx + S)];
            const float* u = c - S * SW;
            const float* d = c + S * SW;
            // Lx: rows [-1 0 1], columns [w0 w1 w0]
            DX[i] = w0 * (u[S] - u[-S]) + w1 * (c[S] - c[-S]) + w0 * (d[S] - d[-S]);
            // Ly: rows [w0 w1 w0], columns [-1 0 1]
            DY[i] = (w0 * d[-S] + w1 * d[0] + w0 * d[S]) - (w0 * u[-S] + w1 * u[0] + w0 * u[S]);
        }
    }
    __syncthreads();
    float* ox = Lx + (size_t)f * pyr_stride;
    float* oy = Ly + (size_t)f * pyr_stride;
    float* od = Ldet + (size_t)f * pyr_stride;
#pragma unroll
    for (int it = 0; it < kTW * kTH / 256; ++it) {
        const int i = tid + it * 256;
        const int ly = i / kTW, lx = i - ly * kTW;
        const int gx = x0 + lx, gy = y0 + ly;
        if (gx >= W || gy >= H) continue;
        const float* cx = &DX[(ly + S) * DW + (lx + S)];
        const float* cy = &DY[(ly + S) * DW + (lx + S)];
        const float* ux = cx - S * DW;
        const float* dx = cx + S * DW;
        const float* uy = cy - S * DW;
        const float* dy = cy + S * DW;
        const float lxx = w0 * (ux[S] - ux[-S]) + w1 * (cx[S] - cx[-S]) + w0 * (dx[S] - dx[-S]);
        const float lxy = (w0 * dx[-S] + w1 * dx[0] + w0 * dx[S]) - (w0 * ux[-S] + w1 * ux[0] + w0 * ux[S]);
        const float lyy = (w0 * dy[-S] + w1 * dy[0] + w0 * dy[S]) - (w0 * uy[-S] + w1 * uy[0] + w0 * uy[S]);
        const size_t o = (size_t)gy * W + gx;
        ox[o] = cx[0];
        oy[o] = cy[0];
        od[o] = __fmul_rn(__fsub_rn(__fmul_rn(lxx, lyy), __fmul_rn(lxy, lxy)), sigma_quat);
    }
}

// Register-window version of k_hessian<S> for even W >= 128, H >= 32: no shared memory, no block barrier.
// A warp owns a span of 64 adjacent columns (two per lane, interleaved: lane t holds columns 2t, 2t + 1 of
// the span) and walks down R output rows.  Horizontal neighbours at distance S come from warp shuffles,
// vertical ones from rolling register windows of depth 2S + 1 (statically indexed: the row loop is unrolled
// by the window period).  Per input row a lane computes the row filters rd = L(x+S) - L(x-S) and
// cs = [w0 w1 w0] . L(x-S, x, x+S); S rows later Lx / Ly of that row exist (column filters over the windows),
// are stored, shuffled and row-filtered again; another S rows later Lxx, Lxy, Lyy and Ldet follow.  The span
// carries 2S halo columns either side (outputs: 64 - 4S columns) and the walk 2S rows above and below.
// Arithmetic: exactly the operations k_hessian<S> compiles to (t = w1 * mid; t = fma(w0, first, t);
// t = fma(w0, last, t)), so both kernels give the same bits.
template <int S>
struct HessWin {
    static constexpr int P = 2 * S + 1;
    float rd[2][P], cs[2][P], rdx[2][P], csx[2][P], csy[2][P];
};

// One round = P consecutive input rows (walk indices base .. base + P - 1).  FAST: all 64 columns and all P
// rows lie inside the image and every store of the round targets an output row, so the loads are plain
// float2 loads off a running pointer and the stores are predicated on the lane's column only.
// the P input rows of one round (walk indices base .. base + P - 1), reflect-101 beyond the image
template <int P>
__device__ __forceinline__ void hess_round_load(float2 (&dst)[P], const float* __restrict__ src, int W, int H, int base,
                                                int r_start, int n_rows, int c0, int ca, int cb, bool interior) {
#pragma unroll
    for (int ph = 0; ph < P; ++ph) {
        const float* row = src + (size_t)reflect101_once(r_start + min(base + ph, n_rows - 1), H) * W;
        if (interior) dst[ph] = __ldg((const float2*)(row + c0));
        else dst[ph] = make_float2(row[ca], row[cb]);
    }
}

template <int S, bool FAST>
__device__ __forceinline__ void hess_round(HessWin<S>& w, const float2 (&cur)[2 * S + 1], float* __restrict__ ox,
                                           float* __restrict__ oy, float* __restrict__ od, int W, int H, int base,
                                           int r_start, int n_rows, int y0, int y_end, int c0, bool col_ok, float w0,
                                           float w1, float sigma_quat) {
    constexpr int P = 2 * S + 1;
    constexpr int DA = (S + 1) / 2, DB = S / 2;            // shuffle distances (lanes) for the two columns
    auto tri = [&](float a, float b, float c) { return __fmaf_rn(w0, c, __fmaf_rn(w0, a, __fmul_rn(w1, b))); };
    auto left = [&](float v0, float v1, int which) {       // value at column - S for column `which` of the pair
        return which == 0 ? __shfl_up_sync(0xffffffffu, (S & 1) ? v1 : v0, DA)
                          : __shfl_up_sync(0xffffffffu, (S & 1) ? v0 : v1, DB);
    };
    auto right = [&](float v0, float v1, int which) {      // value at column + S
        return which == 0 ? __shfl_down_sync(0xffffffffu, (S & 1) ? v1 : v0, DB)
                          : __shfl_down_sync(0xffffffffu, (S & 1) ? v0 : v1, DA);
    };
    // output offsets of row q = r - S (Lx, Ly) for ph = 0; Ldet goes S rows further up
    const long long oq = (long long)(r_start + base - S) * W + c0;
#pragma unroll
    for (int ph = 0; ph < P; ++ph) {
        const int pm = (ph + P - S) % P, pu = (ph + 1) % P;          // rows S and 2S earlier in the windows
        const float l0 = cur[ph].x, l1 = cur[ph].y;
        {
            const float m0 = left(l0, l1, 0), p0 = right(l0, l1, 0), m1 = left(l0, l1, 1), p1 = right(l0, l1, 1);
            w.rd[0][ph] = p0 - m0; w.cs[0][ph] = tri(m0, l0, p0);
            w.rd[1][ph] = p1 - m1; w.cs[1][ph] = tri(m1, l1, p1);
        }
        // first derivatives of row q = r - S
        const float lx0 = tri(w.rd[0][pu], w.rd[0][pm], w.rd[0][ph]), lx1 = tri(w.rd[1][pu], w.rd[1][pm], w.rd[1][ph]);
        const float ly0 = w.cs[0][ph] - w.cs[0][pu], ly1 = w.cs[1][ph] - w.cs[1][pu];
        const int i = base + ph, q = r_start + i - S;
        const long long o = oq + (long long)ph * W;
        const bool st1 = FAST ? col_ok : (col_ok && q >= y0 && q < y_end && i < n_rows);
        if (st1) *(float2*)(ox + o) = make_float2(lx0, lx1);
        if (st1) *(float2*)(oy + o) = make_float2(ly0, ly1);
        {
            const float m0 = left(lx0, lx1, 0), p0 = right(lx0, lx1, 0), m1 = left(lx0, lx1, 1), p1 = right(lx0, lx1, 1);
            w.rdx[0][ph] = p0 - m0; w.csx[0][ph] = tri(m0, lx0, p0);
            w.rdx[1][ph] = p1 - m1; w.csx[1][ph] = tri(m1, lx1, p1);
        }
        {
            const float m0 = left(ly0, ly1, 0), p0 = right(ly0, ly1, 0), m1 = left(ly0, ly1, 1), p1 = right(ly0, ly1, 1);
            w.csy[0][ph] = tri(m0, ly0, p0);
            w.csy[1][ph] = tri(m1, ly1, p1);
        }
        // second derivatives of row p = r - 2S
        const int pr = q - S;
        float det[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float lxx = tri(w.rdx[c][pu], w.rdx[c][pm], w.rdx[c][ph]);
            const float lxy = w.csx[c][ph] - w.csx[c][pu];
            const float lyy = w.csy[c][ph] - w.csy[c][pu];
            det[c] = __fmul_rn(__fsub_rn(__fmul_rn(lxx, lyy), __fmul_rn(lxy, lxy)), sigma_quat);
        }
        if (FAST ? col_ok : (col_ok && pr >= y0 && pr < y_end && i < n_rows))
            *(float2*)(od + o - (long long)S * W) = make_float2(det[0], det[1]);
    }
}

template <int S>
__global__ void __launch_bounds__(32, S == 4 ? 12 : (S == 3 ? 16 : 20))
k_hessian_reg(const float* __restrict__ Lsm, size_t sm_stride, int W, int H, int R, int nspans, float w0, float w1,
              float sigma_quat, float* __restrict__ Lx, float* __restrict__ Ly, float* __restrict__ Ldet,
              size_t pyr_stride) {
    constexpr int P = 2 * S + 1;
    constexpr int OUTW = 64 - 4 * S;
    // one warp per block: every branch below depends on block-uniform values only, so the compiler keeps the
    // shuffles free of divergence checks and can interleave the P independent row steps of a round
    const int lane = threadIdx.x;
    const int span = blockIdx.x;
    const int f = blockIdx.z;
    const int cx0 = span * OUTW, y0 = blockIdx.y * R;
    const int c0 = cx0 - 2 * S + 2 * lane;                 // this lane's first column (may be outside the image)
    const float* src = Lsm + (size_t)f * sm_stride;
    float* ox = Lx + (size_t)f * pyr_stride;
    float* oy = Ly + (size_t)f * pyr_stride;
    float* od = Ldet + (size_t)f * pyr_stride;
    const int r_start = y0 - 2 * S;
    const int y_end = min(y0 + R, H);                      // output rows [y0, y_end)
    const int n_rows = y_end - y0 + 4 * S;                 // input rows walked
    const bool interior = cx0 - 2 * S >= 0 && cx0 - 2 * S + 64 <= W;      // all 64 columns inside the image
    const bool col_ok = lane >= S && lane < 32 - S && c0 < W;
    // border spans: reflected column indices, fixed over the walk
    const int ca = reflect101_once(c0, W), cb = reflect101_once(c0 + 1, W);
    HessWin<S> w;
#pragma unroll
    for (int k = 0; k < P; ++k)
#pragma unroll
        for (int c = 0; c < 2; ++c) w.rd[c][k] = w.cs[c][k] = w.rdx[c][k] = w.csx[c][k] = w.csy[c][k] = 0.f;
    // software pipeline: the next round's P rows are in flight while this round is processed
    float2 nxt[P];
    hess_round_load<P>(nxt, src, W, H, 0, r_start, n_rows, c0, ca, cb, interior);
    for (int base = 0; base < n_rows; base += P) {
        float2 cur[P];
#pragma unroll
        for (int k = 0; k < P; ++k) cur[k] = nxt[k];
        if (base + P < n_rows) hess_round_load<P>(nxt, src, W, H, base + P, r_start, n_rows, c0, ca, cb, interior);
        // Lx / Ly stores are valid for walk indices [3S, n_rows - S), Ldet stores for [4S, n_rows)
        const bool fast = base >= 4 * S && base + P <= n_rows - S;
        if (fast) hess_round<S, true>(w, cur, ox, oy, od, W, H, base, r_start, n_rows, y0, y_end, c0, col_ok, w0, w1, sigma_quat);
        else hess_round<S, false>(w, cur, ox, oy, od, W, H, base, r_start, n_rows, y0, y_end, c0, col_ok, w0, w1, sigma_quat);
    }
}

bool is_prime(int n) {
    if (n < 2) return false;
    for (int d = 2; d * d <= n; ++d)
        if (n % d == 0) return false;
    return true;
}

// kaze/fed.cpp fed_tau_by_process_time(T, 1, 0.25, reordering = true)
int fed_tau(float T, float* tau) {
    const float tau_max = 0.25f;
    const int n = (int)ceilf(sqrtf(3.0f * T / tau_max + 0.25f) - 0.5f - 1.0e-8f);
    if (n <= 0) return 0;
    const float scale = 3.0f * T / (tau_max * (float)(n * (n + 1)));
    const float c = 1.0f / (4.0f * (float)n + 2.0f);
    const float d = scale * tau_max / 2.0f;
    float tauh[kMaxFedSteps];
    for (int k = 0; k < n; ++k) {
        const float h = cosf((float)M_PI * (2.0f * (float)k + 1.0f) * c);
        tauh[k] = d / (h * h);
    }
    const int kappa = n / 2;
    int prime = n + 1;
    while (!is_prime(prime)) prime++;
    for (int k = 0, l = 0; l < n; ++k, ++l) {
        int index;
        while ((index = ((k + 1) * kappa) % prime - 1) >= n) k++;
        tau[l] = tauh[index];
    }
    return n;
}

void gaussian_kernel(int ksize, double sigma, float* out /*centre first*/) {
    double k[16], sum = 0;
    for (int i = 0; i < ksize; ++i) {
        const double x = i - (ksize - 1) * 0.5;
        k[i] = std::exp(-0.5 / (sigma * sigma) * x * x);
        sum += k[i];
    }
    for (int i = 0; i <= ksize / 2; ++i) out[i] = (float)(k[ksize / 2 + i] / sum);
}

}  // namespace

LevelTable make_level_table(int width, int height) {
    LevelTable t{};
    t.width = width;
    t.height = height;
    int n = 0;
    size_t off = 0;
    for (int o = 0; o < 4; ++o) {
        const int power = 1 << o;
        const float rf = 1.0f / power;
        const int lw = (int)(width * rf), lh = (int)(height * rf);
        if ((lw < 80 || lh < 40) && o != 0) break;
        for (int j = 0; j < 4; ++j) {
            LevelInfo& e = t.lv[n];
            e.w = lw; e.h = lh; e.octave = o; e.sublevel = j;
            e.esigma = 1.6f * powf(2.f, (float)j / 4.f + o);
            e.sigma_size = (int)lrint((double)(e.esigma * 1.5f / power));
            e.ratio = (float)power;
            e.border = (int)lrint((double)(10.0f * sqrtf(2.0f) * e.sigma_size)) + 1;
            e.plane_off = off;
            off += (size_t)lw * lh;
            e.n_tau = 0;
            ++n;
        }
    }
    t.n_levels = n;
    t.pyramid_floats = off;
    for (int i = 1; i < n; ++i) {
        const float et = 0.5f * (t.lv[i].esigma * t.lv[i].esigma), ep = 0.5f * (t.lv[i - 1].esigma * t.lv[i - 1].esigma);
        t.lv[i].n_tau = fed_tau(et - ep, t.lv[i].tau);
    }
    return t;
}

LevelsDev make_levels_dev(const LevelTable& lt) {
    LevelsDev d{};
    d.n = lt.n_levels;
    for (int i = 0; i < lt.n_levels; ++i) {
        d.lv[i].w = lt.lv[i].w; d.lv[i].h = lt.lv[i].h;
        d.lv[i].sigma_size = lt.lv[i].sigma_size; d.lv[i].border = lt.lv[i].border;
        d.lv[i].octave = lt.lv[i].octave; d.lv[i].ratio = lt.lv[i].ratio; d.lv[i].esigma = lt.lv[i].esigma;
        d.lv[i].plane_off = lt.lv[i].plane_off;
    }
    return d;
}

// per-device shared-memory opt-ins of the tiled kernels (run by dunk_ctx_create on the context's device)
static int scale_device_init(dunk_ctx*) {
    DUNK_CUDA(cudaFuncSetAttribute(k_fed, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kFedS * kFedS * 4));
    DUNK_CUDA(cudaFuncSetAttribute(k_hessian_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DUNK_CUDA(cudaFuncSetAttribute(k_hessian<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DUNK_CUDA(cudaFuncSetAttribute(k_hessian<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DUNK_CUDA(cudaFuncSetAttribute(k_hessian<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    return DUNK_OK;
}
static DeviceInitReg scale_device_init_reg(scale_device_init);

int akaze_build_scale_space(dunk_ctx* ctx, cudaStream_t st, const LevelTable& lt, const AkazeWorkspace& ws,
                            const unsigned char* images, size_t image_stride_bytes, int row_stride, int channels,
                            int frames) {
    const int W = lt.width, H = lt.height;
    const size_t pyr = lt.pyramid_floats, plane = (size_t)W * H;
    const dim3 blk(kBX, kBY);
    Gauss9 g9;
    Gauss5 g5;
    gaussian_kernel(9, 1.6, g9.k);
    gaussian_kernel(5, 1.0, g5.k);
    // level 0
    {
        const dim3 grid(div_up(W, kTW), div_up(H, kTH), frames);
        // debug switch for tools/ab_kernels.py (bit-for-bit comparison of the two kernel generations)
        static const bool l0_old = getenv("DUNK_LEVEL0_OLD") != nullptr;
        const bool l0_reg = !l0_old && W % 2 == 0 && W >= 128 && H >= 16 && pyr % 2 == 0 && plane % 2 == 0 && lt.lv[0].plane_off % 2 == 0;
        // vector loads of a pixel pair: 2 bytes (gray) or 8 bytes (BGRA)
        const int pair_bytes = channels == 1 ? 2 : 8;
        const bool l0_aligned = (channels == 1 || channels == 4) && ((uintptr_t)images % pair_bytes) == 0 &&
                                image_stride_bytes % pair_bytes == 0 && row_stride % pair_bytes == 0;
        const int l0_mode = !l0_aligned ? 2 : (channels == 1 ? 0 : 1);
        const int l0_spans = div_up(W, 56);
        int l0_R = 128;
        while (l0_R > 16 && (long long)l0_spans * div_up(H, l0_R) * frames < 4096) l0_R /= 2;
        {
            ProfScope ps(ctx, st, "scale.gray_gauss9", (double)frames * plane * (channels + 4));
            if (l0_reg) {
                const dim3 g0(l0_spans, div_up(H, l0_R), frames);
                float* lt0 = ws.Lt + lt.lv[0].plane_off;
                if (l0_mode == 0) k_gray_gauss9_reg<0><<<g0, 32, 0, st>>>(images, image_stride_bytes, row_stride, channels, W, H, l0_R, g9, lt0, pyr);
                else if (l0_mode == 1) k_gray_gauss9_reg<1><<<g0, 32, 0, st>>>(images, image_stride_bytes, row_stride, channels, W, H, l0_R, g9, lt0, pyr);
                else k_gray_gauss9_reg<2><<<g0, 32, 0, st>>>(images, image_stride_bytes, row_stride, channels, W, H, l0_R, g9, lt0, pyr);
            } else
            k_gray_gauss9<<<grid, blk, 0, st>>>(images, image_stride_bytes, row_stride, channels, W, H, g9,
                                                ws.Lt + lt.lv[0].plane_off, pyr);
            DUNK_KERNEL_CHECK(ctx);
        }
        if (lt.n_levels > 1) {
            DUNK_CUDA(cudaMemsetAsync(ws.hmax, 0, (size_t)frames * 4, st));
            DUNK_CUDA(cudaMemsetAsync(ws.hist, 0, (size_t)frames * kNBins * 4, st));
            {
                ProfScope ps(ctx, st, "scale.contrast_modg", (double)frames * plane * (channels + 4));
                if (l0_reg) {
                    const dim3 g0(l0_spans, div_up(H, l0_R), frames);
                    if (l0_mode == 0) k_contrast_modg_reg<0><<<g0, 32, 0, st>>>(images, image_stride_bytes, row_stride, channels, W, H, l0_R, g5, ws.Lflow, plane, ws.hmax);
                    else if (l0_mode == 1) k_contrast_modg_reg<1><<<g0, 32, 0, st>>>(images, image_stride_bytes, row_stride, channels, W, H, l0_R, g5, ws.Lflow, plane, ws.hmax);
                    else k_contrast_modg_reg<2><<<g0, 32, 0, st>>>(images, image_stride_bytes, row_stride, channels, W, H, l0_R, g5, ws.Lflow, plane, ws.hmax);
                } else
                k_contrast_modg<<<grid, blk, 0, st>>>(images, image_stride_bytes, row_stride, channels, W, H, g5,
                                                      ws.Lflow, plane, ws.hmax);
                DUNK_KERNEL_CHECK(ctx);
            }
            const int hb = std::max(1, std::min(64, div_up((long long)(W - 2) * (H - 2), 256 * 16)));
            {
                ProfScope ps(ctx, st, "scale.contrast_hist", (double)frames * plane * 4);
                k_contrast_hist<<<dim3(hb, frames), 256, 0, st>>>(ws.Lflow, plane, W, H, ws.hmax, ws.hist);
                DUNK_KERNEL_CHECK(ctx);
            }
            {
                ProfScope ps(ctx, st, "scale.contrast_final", 0.0);
                k_contrast_final<<<div_up(frames, 64), 64, 0, st>>>(ws.hist, ws.hmax, W, H, 0.7f, ws.kcontrast, frames);
                DUNK_KERNEL_CHECK(ctx);
            }
        }
    }
    auto hessian = [&](int i, const float* lsm, size_t lsm_stride) -> int {
        const LevelInfo& e = lt.lv[i];
        const int s = e.sigma_size;
        float w0, w1;
        if (s == 1) {
            w0 = 3.f / 32.f; w1 = 10.f / 32.f;
        } else {
            const float w = 10.0f / 3.0f;
            w0 = 1.0f / (2.0f * s * (w + 2.0f));
            w1 = w * w0;
        }
        const size_t smem = ((size_t)(kTW + 4 * s) * (kTH + 4 * s) + 2 * (size_t)(kTW + 2 * s) * (kTH + 2 * s)) * 4;
        const dim3 grid(div_up(e.w, kTW), div_up(e.h, kTH), frames);
        {
            ProfScope ps(ctx, st, "scale.hessian", (double)frames * e.w * e.h * 16);
            float* lx = ws.Lx + e.plane_off;
            float* ly = ws.Ly + e.plane_off;
            float* ld = ws.Ldet + e.plane_off;
            const float sq = (float)(s * s * s * s);
            // debug switch for tools/ab_kernels.py (bit-for-bit comparison of the two kernels)
            static const bool use_old = getenv("DUNK_HESSIAN_OLD") != nullptr;
            const bool reg_ok = !use_old && s >= 2 && s <= 4 && e.w % 2 == 0 && e.w >= 128 && e.h >= 32 && lsm_stride % 2 == 0 &&
                                pyr % 2 == 0 && e.plane_off % 2 == 0;
            if (reg_ok) {
                const int outw = 64 - 4 * s, nspans = div_up(e.w, outw);
                // rows per warp: long walks amortise the 4S halo rows, short ones keep small levels parallel
                int R = 128;
                while (R > 16 && (long long)nspans * div_up(e.h, R) * frames < 4096) R /= 2;
                const dim3 g2(nspans, div_up(e.h, R), frames);
                if (s == 2) k_hessian_reg<2><<<g2, 32, 0, st>>>(lsm, lsm_stride, e.w, e.h, R, nspans, w0, w1, sq, lx, ly, ld, pyr);
                else if (s == 3) k_hessian_reg<3><<<g2, 32, 0, st>>>(lsm, lsm_stride, e.w, e.h, R, nspans, w0, w1, sq, lx, ly, ld, pyr);
                else k_hessian_reg<4><<<g2, 32, 0, st>>>(lsm, lsm_stride, e.w, e.h, R, nspans, w0, w1, sq, lx, ly, ld, pyr);
            } else
            if (s == 2) k_hessian<2><<<grid, blk, smem, st>>>(lsm, lsm_stride, e.w, e.h, w0, w1, sq, lx, ly, ld, pyr);
            else if (s == 3) k_hessian<3><<<grid, blk, smem, st>>>(lsm, lsm_stride, e.w, e.h, w0, w1, sq, lx, ly, ld, pyr);
            else if (s == 4) k_hessian<4><<<grid, blk, smem, st>>>(lsm, lsm_stride, e.w, e.h, w0, w1, sq, lx, ly, ld, pyr);
            else k_hessian_generic<<<grid, blk, smem, st>>>(lsm, lsm_stride, e.w, e.h, s, w0, w1, sq, lx, ly, ld, pyr);
            DUNK_KERNEL_CHECK(ctx);
        }
        return DUNK_OK;
    };
    int rc = hessian(0, ws.Lt + lt.lv[0].plane_off, pyr);
    if (rc) return rc;

    float kscale = 1.0f;
    for (int i = 1; i < lt.n_levels; ++i) {
        const LevelInfo& e = lt.lv[i];
        const LevelInfo& p = lt.lv[i - 1];
        const int n = e.n_tau;
        static const int fed_maxk = getenv("DUNK_FED_MAXK") ? std::min(kFedMaxK, std::max(1, atoi(getenv("DUNK_FED_MAXK")))) : kFedMaxK;
        const int m = div_up(n, fed_maxk);   // launches; the n steps are split as evenly as possible
        float* P = ws.Lt + e.plane_off;   // final home of this level's Lt (stride pyr)
        float* Q = ws.Ltmp;               // ping-pong partner (stride plane)
        const float* init;
        size_t init_stride;
        if (e.octave > p.octave) {
            kscale *= 0.75f;
            // first FED launch writes to ((m-1) even ? P : Q); decimate into the other one
            const bool out0_is_P = ((m - 1) % 2 == 0);
            float* dstbuf = out0_is_P ? Q : P;
            const size_t dstride = out0_is_P ? plane : pyr;
            {
                ProfScope ps(ctx, st, "scale.halfsample", (double)frames * ((double)p.w * p.h + (double)e.w * e.h) * 4);
                const float* hsrc = ws.Lt + p.plane_off;
                if (p.w == 2 * e.w && p.h == 2 * e.h && p.w % 4 == 0 && pyr % 4 == 0 && dstride % 2 == 0 && e.w % 2 == 0 &&
                    ((uintptr_t)hsrc & 15) == 0 && ((uintptr_t)dstbuf & 7) == 0)
                    k_halfsample_x2<<<dim3(div_up(e.w / 2, 128), e.h, frames), 128, 0, st>>>(hsrc, pyr, p.w, dstbuf, dstride, e.w, e.h);
                else
                k_halfsample<<<dim3(div_up(e.w, 256), e.h, frames), 256, 0, st>>>(hsrc, pyr, p.w, p.h, dstbuf,
                                                                               dstride, e.w, e.h);
                DUNK_KERNEL_CHECK(ctx);
            }
            init = dstbuf;
            init_stride = dstride;
        } else {
            init = ws.Lt + p.plane_off;
            init_stride = pyr;
        }
        const dim3 grid(div_up(e.w, kTW), div_up(e.h, kTH), frames);
        {
            ProfScope ps(ctx, st, "scale.prep_level", (double)frames * e.w * e.h * 12);
            // debug switch for tools/ab_kernels.py (bit-for-bit comparison of the two kernels)
            static const bool prep_old = getenv("DUNK_PREP_OLD") != nullptr;
            if (!prep_old && e.w % 2 == 0 && e.w >= 128 && e.h >= 16 && init_stride % 2 == 0 && plane % 2 == 0 &&
                ((uintptr_t)init & 7) == 0) {
                const int nspans = div_up(e.w, 56);
                int R = 128;
                while (R > 16 && (long long)nspans * div_up(e.h, R) * frames < 4096) R /= 2;
                k_prep_level_reg<<<dim3(nspans, div_up(e.h, R), frames), 32, 0, st>>>(
                    init, init_stride, e.w, e.h, R, g5, ws.kcontrast, kscale, ws.Lsmooth, ws.Lflow, plane);
            } else
            k_prep_level<<<dim3(div_up(e.w, 8 * kPrepCols), div_up(e.h, kPrepRows), frames), 256, 0, st>>>(
                init, init_stride, e.w, e.h, g5, ws.kcontrast, kscale, ws.Lsmooth, ws.Lflow, plane);
            DUNK_KERNEL_CHECK(ctx);
        }
        if ((rc = hessian(i, ws.Lsmooth, plane))) return rc;
        const float* in = init;
        size_t in_stride = init_stride;
        int done = 0;
        for (int j = 0; j < m; ++j) {
            FedSteps fs{};
            fs.k = n / m + (j < n % m ? 1 : 0);
            for (int k = 0; k < fs.k; ++k) fs.step[k] = e.tau[done + k] * 0.5f;
            done += fs.k;
            const bool out_is_P = ((m - 1 - j) % 2 == 0);
            float* out = out_is_P ? P : Q;
            const size_t out_stride = out_is_P ? pyr : plane;
            const int T = kFedS - 2 * fs.k;
            const dim3 fgrid(div_up(e.w, T), div_up(e.h, T), frames);
            {
                ProfScope ps(ctx, st, "scale.fed", (double)frames * e.w * e.h * 12);
                // debug switch for tools/ab_kernels.py (bit-for-bit comparison of the two kernels)
                static const bool fed_old = getenv("DUNK_FED_OLD") != nullptr;
                // the row-walking cascade needs many spans x bands x frames to fill the GPU: measured faster than the
                // tiled kernel at 1024^2 and 512^2 (1.6x / 1.3x), slower at 256^2 and below (0.8x)
                // (64 frames); with 128 or more frames the 256-px levels have enough spans too (5.65 vs 6.02 ms / 256 frames)
                const int fed_minw = frames >= 128 ? 256 : 384;
                if (!fed_old && e.w % 2 == 0 && e.w >= fed_minw && e.h >= fed_minw / 2 && in_stride % 2 == 0 && out_stride % 2 == 0 && plane % 2 == 0 &&
                    ((uintptr_t)in & 7) == 0 && ((uintptr_t)out & 7) == 0) {
                    const int kh = (fs.k + 1) / 2 * 2, nspans = div_up(e.w, 64 - 2 * kh);
                    int R = 128;
                    while (R > 16 && (long long)nspans * div_up(e.h, R) * frames < 4096) R /= 2;
                    const dim3 g3(nspans, div_up(e.h, R), frames);
#define DUNK_FED_CASE(KK) case KK: k_fed_reg<KK><<<g3, 32, 0, st>>>(in, in_stride, out, out_stride, ws.Lflow, plane, e.w, e.h, R, fs); break;
                    switch (fs.k) {
                        DUNK_FED_CASE(1) DUNK_FED_CASE(2) DUNK_FED_CASE(3) DUNK_FED_CASE(4)
                        DUNK_FED_CASE(5) DUNK_FED_CASE(6) DUNK_FED_CASE(7) DUNK_FED_CASE(8)
                    }
#undef DUNK_FED_CASE
                } else
                k_fed<<<fgrid, dim3(16, 16), (size_t)2 * kFedS * kFedS * 4, st>>>(in, in_stride, out, out_stride, ws.Lflow, plane,
                                                                                 e.w, e.h, fs);
                DUNK_KERNEL_CHECK(ctx);
            }
            in = out;
            in_stride = out_stride;
        }
    }
    return DUNK_OK;
}

}  // namespace dunk
