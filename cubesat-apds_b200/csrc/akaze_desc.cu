// AKAZE stage 1c: dominant orientation + MLDB-486 descriptor, one warp per keypoint.
// Follows Compute_Main_Orientation / MLDB_Full_Descriptor_Invoker of OpenCV's AKAZEFeatures.cpp as
// restated in oracle/akaze_oracle.py: same sample tables, same f32 operation order (the sums are
// kept sequential per window / per grid cell so that rounding matches), fastAtan2 polynomial.
#include <algorithm>
#include "akaze.h"
#include <cfloat>
#include <cmath>

namespace dunk {

namespace {

struct OriTable {
    signed char xi[109], yi[109];
    float w[109];
};
// global, not __constant__: lanes index it with 32 distinct offsets (see g_cmp below)
__device__ OriTable c_ori;
// value indices (cell * 3 + channel) of the two operands of descriptor bit `dpos`, a | b << 8.  In GLOBAL memory
// on purpose: lane l of a warp reads entry w * 32 + l, a coalesced 64-byte read; the constant cache would replay
// the 32 distinct addresses one by one
__device__ unsigned short g_cmp[512];

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    // cv::hal::fastAtan32f (mathfuncs_core), degrees
    const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

constexpr int kWarpsPerBlock = 8;
constexpr int kMldbWarps = 4;          // 128-thread blocks: the kernel needs ~150 registers, small blocks pack the register file better
constexpr int kSlices = 42, kWin = 7, kAng = 109;

struct OriShared {
    float rx[kAng], ry[kAng];
    unsigned char sorted[kAng];
    unsigned char run[kSlices];
    int cum[kSlices + 1];
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_orientation(DunkKeyPoint* __restrict__ kps_all, int kp_cap, const int* __restrict__ kp_count,
              const float* __restrict__ Lx, const float* __restrict__ Ly, size_t pyr_stride, LevelsDev lv) {
    __shared__ OriShared sh_all[kWarpsPerBlock];
    const int f = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ki = blockIdx.x * kWarpsPerBlock + warp;
    if (ki >= kp_count[f]) return;
    OriShared& sh = sh_all[warp];
    DunkKeyPoint* kp = kps_all + (size_t)f * kp_cap + ki;
    const LevelDev& e = lv.lv[kp->class_id];
    const float ratio = e.ratio;
    const int scale = __float2int_rn(__fdiv_rn(__fmul_rn(0.5f, kp->size), ratio));
    const int x0 = __float2int_rn(__fdiv_rn(kp->x, ratio));
    const int y0 = __float2int_rn(__fdiv_rn(kp->y, ratio));
    const float* lx = Lx + (size_t)f * pyr_stride + e.plane_off;
    const float* ly = Ly + (size_t)f * pyr_stride + e.plane_off;
    const float ang_step = (float)(2.0 * M_PI / kSlices);
    // all 8 gathers of a lane are issued before any of them is used (4 samples x Lx, Ly)
    constexpr int kIt = (kAng + 31) / 32;
    float gx[kIt], gy[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int s = min(lane + 32 * it, kAng - 1);
        const int y = min(max(y0 + __ldg(&c_ori.yi[s]) * scale, 0), e.h - 1);
        const int x = min(max(x0 + __ldg(&c_ori.xi[s]) * scale, 0), e.w - 1);
        gx[it] = __ldg(lx + (size_t)y * e.w + x);
        gy[it] = __ldg(ly + (size_t)y * e.w + x);
    }
    for (int b = lane; b <= kSlices; b += 32) sh.cum[b] = 0;
    if (lane < kSlices) sh.run[lane] = 0;
    if (lane + 32 < kSlices) sh.run[lane + 32] = 0;
    __syncwarp();
    int bins[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int s = lane + 32 * it;
        bins[it] = 255;
        if (s < kAng) {
            const float w = __ldg(&c_ori.w[s]);
            const float rx = __fmul_rn(w, gx[it]);
            const float ry = __fmul_rn(w, gy[it]);
            sh.rx[s] = rx;
            sh.ry[s] = ry;
            const float ang = __fmul_rn(fast_atan2_deg(ry, rx), (float)(M_PI / 180.0));
            int b = (int)__fdiv_rn(ang, ang_step);
            if (b < 0 || b >= kSlices) b = 0;
            bins[it] = b;
            atomicAdd(&sh.cum[b + 1], 1);                  // slice populations
        }
    }
    __syncwarp();
    {   // inclusive scan of the 43 counters: cum[b] = first position of slice b, cum[kSlices] = kAng
        int v0 = sh.cum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v0, o);
            if (lane >= o) v0 += t;
        }
        const int tot = __shfl_sync(0xffffffffu, v0, 31);
        int v1 = lane + 32 <= kSlices ? sh.cum[lane + 32] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v1, o);
            if (lane >= o) v1 += t;
        }
        __syncwarp();
        sh.cum[lane] = v0;
        if (lane + 32 <= kSlices) sh.cum[lane + 32] = v1 + tot;
    }
    __syncwarp();
    // quantized_counting_sort places sample i at (end of its slice) - 1 - (same-slice samples before i), i.e. a slice
    // reads in DESCENDING sample index: position = slice start + number of same-slice samples with a larger index
#pragma unroll
    for (int it = kIt - 1; it >= 0; --it) {
        const int s = lane + 32 * it;
        const int b = bins[it];
        const unsigned m = __match_any_sync(0xffffffffu, b);
        if (b != 255) sh.sorted[sh.cum[b] + sh.run[b] + __popc(m & (0xfffffffeu << lane))] = (unsigned char)s;
        __syncwarp();
        if (b != 255 && (m & ((1u << lane) - 1)) == 0) sh.run[b] += (unsigned char)__popc(m);   // lowest lane of the group
        __syncwarp();
    }
    float bestN = -1.f, bestX = 0.f, bestY = 0.f;
    int bestW = 1 << 30;
    for (int sn = lane; sn < kSlices; sn += 32) {
        float sx = 0.f, sy = 0.f;
        if (sn <= kSlices - kWin) {
            for (int i = sh.cum[sn]; i < sh.cum[sn + kWin]; ++i) {
                sx = __fadd_rn(sx, sh.rx[sh.sorted[i]]);
                sy = __fadd_rn(sy, sh.ry[sh.sorted[i]]);
            }
        } else {
            const int remain = sn + kWin - kSlices;
            for (int i = sh.cum[sn]; i < sh.cum[kSlices]; ++i) {
                sx = __fadd_rn(sx, sh.rx[sh.sorted[i]]);
                sy = __fadd_rn(sy, sh.ry[sh.sorted[i]]);
            }
            for (int i = sh.cum[0]; i < sh.cum[remain]; ++i) {
                sx = __fadd_rn(sx, sh.rx[sh.sorted[i]]);
                sy = __fadd_rn(sy, sh.ry[sh.sorted[i]]);
            }
        }
        const float nrm = __fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy));
        if (nrm > bestN) { bestN = nrm; bestX = sx; bestY = sy; bestW = sn; }   // lower sn first per lane
    }
    // arg-max with first-window-wins ties (the source replaces only on strictly greater)
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float n2 = __shfl_xor_sync(0xffffffffu, bestN, o);
        const float x2 = __shfl_xor_sync(0xffffffffu, bestX, o);
        const float y2 = __shfl_xor_sync(0xffffffffu, bestY, o);
        const int w2 = __shfl_xor_sync(0xffffffffu, bestW, o);
        if (n2 > bestN || (n2 == bestN && w2 < bestW)) { bestN = n2; bestX = x2; bestY = y2; bestW = w2; }
    }
    if (lane == 0) kp->angle = fast_atan2_deg(bestY, bestX);   // degrees (cv2 4.13 stores fastAtan2)
}

__device__ __forceinline__ int toggle_flt(float v) {
    const int x = __float_as_int(v);
    return x ^ ((x < 0) ? 0x7fffffff : 0);   // CV_TOGGLE_FLT
}

__global__ void __launch_bounds__(kMldbWarps * 32)
k_mldb(const DunkKeyPoint* __restrict__ kps_all, int kp_cap, const int* __restrict__ kp_count,
       const float* __restrict__ Lt, const float* __restrict__ Lx, const float* __restrict__ Ly, size_t pyr_stride,
       LevelsDev lv, uint4* __restrict__ desc64_all) {
    __shared__ int vals_all[kMldbWarps][29 * 3];
    __shared__ float pri_all[kMldbWarps][441];     // Lt at the lattice points (NaN = outside the image)
    __shared__ float2 pxy_all[kMldbWarps][441];    // rotated (Lx, Ly)
    const int f = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a fixed, small grid strides over the frame's keypoints: a grid sized for the keypoint CAPACITY is > 90 %
    // empty blocks whose turnover (45 KB of shared memory each) would dominate the kernel
    const int count = kp_count[f];
    for (int ki = blockIdx.x * kMldbWarps + warp; ki < count; ki += gridDim.x * kMldbWarps) {
    int* vals = vals_all[warp];
    const DunkKeyPoint kp = kps_all[(size_t)f * kp_cap + ki];
    const LevelDev& e = lv.lv[kp.class_id];
    const float ratio = (float)(1 << kp.octave);
    const int scale = __float2int_rn(__fdiv_rn(__fmul_rn(0.5f, kp.size), ratio));
    const float fscale = (float)scale;
    const float xf = __fdiv_rn(kp.x, ratio), yf = __fdiv_rn(kp.y, ratio);
    const float angle = __fmul_rn(kp.angle, (float)(M_PI / 180.0));
    const float co = cosf(angle), si = sinf(angle);
    const float* lt = Lt + (size_t)f * pyr_stride + e.plane_off;
    const float* lx = Lx + (size_t)f * pyr_stride + e.plane_off;
    const float* ly = Ly + (size_t)f * pyr_stride + e.plane_off;
    // The three cell grids (2x2 of 10x10, 3x3 of 7x7, 4x4 of 5x5 samples) all sample the same 21 x 21 lattice
    // of offsets (k, l) in -10..10.  Phase 1 gathers every lattice point ONCE with all 32 lanes (441 x 3 plane
    // reads instead of 1241 x 3): the 14 rounds are fully unrolled with clamped (always valid) addresses, so all
    // 42 loads of a lane are in flight together — the kernel is bound by gather latency, not arithmetic.
    // Phase 2 lets the 29 cell lanes add their samples from shared memory in OpenCV's (k, l) order, so the
    // f32 sums are bit-identical.
    float* pri = pri_all[warp];
    float2* pxy = pxy_all[warp];
    {
        const bool kfast = fabsf(co) >= fabsf(si);
        float ri[14], rx[14], ry[14];
        bool ok[14];
#pragma unroll
        for (int it = 0; it < 14; ++it) {
            // lattice point of this lane: the index that moves mostly along image x (k when |cos| >= |sin|, else l) is the
            // fastest over the lanes, so neighbouring lanes read neighbouring pixels of one row and share 32-byte sectors
            // (with l always fastest every lane of an upright keypoint hit its own row: 32 sectors per gather, l1tex-bound)
            const int p = min(lane + 32 * it, 440);
            const int k = (kfast ? p % 21 : p / 21) - 10, l = (kfast ? p / 21 : p % 21) - 10;
            const float sy = __fadd_rn(yf, __fadd_rn(__fmul_rn(__fmul_rn((float)l, co), fscale),
                                                     __fmul_rn(__fmul_rn((float)k, si), fscale)));
            const float sx = __fadd_rn(xf, __fadd_rn(__fmul_rn(__fmul_rn((float)(-l), si), fscale),
                                                     __fmul_rn(__fmul_rn((float)k, co), fscale)));
            const int y1 = __float2int_rn(sy), x1 = __float2int_rn(sx);
            ok[it] = y1 >= 0 && y1 < e.h && x1 >= 0 && x1 < e.w;
            const size_t o = (size_t)min(max(y1, 0), e.h - 1) * e.w + min(max(x1, 0), e.w - 1);
            ri[it] = __ldg(lt + o);
            rx[it] = __ldg(lx + o);
            ry[it] = __ldg(ly + o);
        }
#pragma unroll
        for (int it = 0; it < 14; ++it) {
            const int p = lane + 32 * it;
            if (p < 441) {
                const int q = kfast ? (p % 21) * 21 + p / 21 : p;           // slot (k + 10) * 21 + (l + 10)
                pri[q] = ok[it] ? ri[it] : __int_as_float(0x7fc00000);     // NaN marks a sample outside the image
                pxy[q] = make_float2(__fadd_rn(__fmul_rn(-rx[it], si), __fmul_rn(ry[it], co)),
                                     __fadd_rn(__fmul_rn(rx[it], co), __fmul_rn(ry[it], si)));
            }
        }
    }
    __syncwarp();
    if (lane < 29) {
        // cell -> (grid, i, j): pattern 10; steps 10, 7, 5 -> 2x2, 3x3, 4x4 cells starting at -10
        int step, n, local;
        if (lane < 4) { step = 10; n = 2; local = lane; }
        else if (lane < 13) { step = 7; n = 3; local = lane - 4; }
        else { step = 5; n = 4; local = lane - 13; }
        const int i0 = -10 + (local / n) * step, j0 = -10 + (local % n) * step;
        float di = 0.f, dx = 0.f, dy = 0.f;
        int nsamples = 0;
        // uniform 10 x 10 trip counts (cells of 7 x 7 / 5 x 5 samples mask the rest): the inner loop is fully
        // unrolled with its 20 shared loads issued together; predicated adds keep OpenCV's (k, l) order exactly
        for (int kk = 0; kk < 10; ++kk) {
            const bool krow = kk < step;
            const int rowbase = ((krow ? i0 + kk : i0) + 10) * 21 + (j0 + 10);
            float r0[10];
            float2 xy[10];
#pragma unroll
            for (int l = 0; l < 10; ++l) {
                const int idx = rowbase + min(l, step - 1);
                r0[l] = pri[idx];
                xy[l] = pxy[idx];
            }
#pragma unroll
            for (int l = 0; l < 10; ++l) {
                const bool use = krow && l < step && r0[l] == r0[l];
                di = use ? __fadd_rn(di, r0[l]) : di;
                dx = use ? __fadd_rn(dx, xy[l].x) : dx;
                dy = use ? __fadd_rn(dy, xy[l].y) : dy;
                nsamples += use;
            }
        }
        if (nsamples > 0) {
            const float inv = __fdiv_rn(1.0f, (float)nsamples);
            di = __fmul_rn(di, inv); dx = __fmul_rn(dx, inv); dy = __fmul_rn(dy, inv);
        }
        vals[lane * 3 + 0] = toggle_flt(di);
        vals[lane * 3 + 1] = toggle_flt(dx);
        vals[lane * 3 + 2] = toggle_flt(dy);
    }
    __syncwarp();
    unsigned word[16];
#pragma unroll
    for (int w = 0; w < 16; ++w) {
        const int dpos = w * 32 + lane;
        bool bit = false;
        if (dpos < 486) {
            const unsigned ab = __ldg(&g_cmp[dpos]);
            bit = vals[ab & 255u] > vals[ab >> 8];
        }
        word[w] = __ballot_sync(0xffffffffu, bit);
    }
    if (lane < 4) {
        uint4 v;
        v.x = word[lane * 4 + 0]; v.y = word[lane * 4 + 1]; v.z = word[lane * 4 + 2]; v.w = word[lane * 4 + 3];
        // word[] is warp-uniform; select with a switch-free copy
        desc64_all[((size_t)f * kp_cap + ki) * 4 + lane] = v;
    }
    __syncwarp();
    }
}

// __device__ / __constant__ symbols exist once per DEVICE: uploaded by dunk_ctx_create on the context's device
int upload_tables(dunk_ctx*) {
    OriTable t;
    int k = 0;
    for (int i = -6; i <= 6; ++i)
        for (int j = -6; j <= 6; ++j)
            if (i * i + j * j < 36) {
                // gauss25[|i|][|j|]: 2-D Gaussian sigma 2.5, tabulated to 8 decimals in the source
                const double g = std::exp(-(double)(i * i + j * j) / (2 * 2.5 * 2.5)) / (2 * M_PI * 2.5 * 2.5);
                t.w[k] = (float)(std::round(g * 1e8) / 1e8);
                t.xi[k] = (signed char)i;
                t.yi[k] = (signed char)j;
                ++k;
            }
    if (k != 109) {
        set_error("orientation table has %d entries", k);
        return DUNK_ERR_ASSERT;
    }
    unsigned char a[486], b[486];
    int dpos = 0, base = 0;
    for (int z = 0; z < 3; ++z) {
        const int count = (z + 2) * (z + 2);
        for (int pos = 0; pos < 3; ++pos)
            for (int i = 0; i < count; ++i)
                for (int j = i + 1; j < count; ++j) {
                    a[dpos] = (unsigned char)((base + i) * 3 + pos);
                    b[dpos] = (unsigned char)((base + j) * 3 + pos);
                    ++dpos;
                }
        base += count;
    }
    if (dpos != 486) {
        set_error("MLDB comparison table has %d entries", dpos);
        return DUNK_ERR_ASSERT;
    }
    DUNK_CUDA(cudaMemcpyToSymbol(c_ori, &t, sizeof t));
    unsigned short ab[512] = {0};
    for (int i = 0; i < 486; ++i) ab[i] = (unsigned short)(a[i] | (b[i] << 8));
    DUNK_CUDA(cudaMemcpyToSymbol(g_cmp, ab, sizeof ab));
    return DUNK_OK;
}
DeviceInitReg desc_tables_reg(upload_tables);

}  // namespace

int akaze_describe(dunk_ctx* ctx, cudaStream_t st, const LevelTable& lt, const AkazeWorkspace& ws, int frames) {
    const LevelsDev lv = make_levels_dev(lt);
    const dim3 grid(div_up(ws.kp_cap, kWarpsPerBlock), frames);
    {
        ProfScope ps(ctx, st, "describe.orientation", 0.0);
        k_orientation<<<grid, kWarpsPerBlock * 32, 0, st>>>(ws.kps, ws.kp_cap, ws.kp_count, ws.Lx, ws.Ly, lt.pyramid_floats, lv);
        DUNK_KERNEL_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, st, "describe.mldb", 0.0);
        const dim3 mgrid(std::min(div_up(ws.kp_cap, kMldbWarps), 128), frames);
        k_mldb<<<mgrid, kMldbWarps * 32, 0, st>>>(ws.kps, ws.kp_cap, ws.kp_count, ws.Lt, ws.Lx, ws.Ly, lt.pyramid_floats, lv,
                                                      ws.desc64);
        DUNK_KERNEL_CHECK(ctx);
    }
    return DUNK_OK;
}

}  // namespace dunk
