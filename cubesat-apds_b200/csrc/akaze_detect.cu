// AKAZE stage 1b: scale-space extrema -> keypoints, batched over frames, reproducing OpenCV's
// sequential semantics exactly (oracle/akaze_oracle.py, established differentially against cv2):
//   1. strict 3x3 maxima of Ldet above the threshold inside the level border;
//   2. row-major ordered candidate lists (row buckets: count -> scan -> scatter -> sort rows);
//   3. same-level suppression within sigma_size (later-stronger replaces, later-weaker dropped);
//   4. two one-directional cross-level passes (a keypoint STRONGER than the first neighbour found
//      in the adjacent level deletes that neighbour);
//   5. 2x2 sub-pixel solve, |d| <= 1, ordered compaction, optional top-max_points by response.
// Order dependence is local, so candidates whose search window holds at most one partner are
// decided in parallel and only the rare multi-partner ones are replayed sequentially.
#include "akaze.h"
#include <climits>

namespace dunk {

namespace {

constexpr int kExtremaRows = 4;      // centre loads in flight per thread (8 measured slower: 1.81 vs 1.58 ms per 256 frames)

__global__ void __launch_bounds__(256)
k_extrema(const float* __restrict__ Ldet, size_t pyr_stride, LevelDev e, int level, int row_base, float thr,
          Cand* __restrict__ cand_raw, int cand_cap, int* __restrict__ cand_count, int* __restrict__ row_count,
          int total_rows) {
    // a block covers 32 x 8 kExtremaRows pixels; every thread issues its kExtremaRows centre loads (rows y, y+8, ...) before
    // testing any of them: almost every pixel fails the threshold, so the kernel is one streaming read and
    // needs the memory-level parallelism, not the arithmetic (the rest of its time is the divergent 3x3 re-test of the pixels above the threshold)
    const int f = blockIdx.z;
    const int x = e.border + blockIdx.x * 32 + threadIdx.x;
    const int y0 = e.border + blockIdx.y * (8 * kExtremaRows) + threadIdx.y;
    if (x >= e.w - e.border) return;
    const float* L = Ldet + (size_t)f * pyr_stride + e.plane_off;
    float v[kExtremaRows];
#pragma unroll
    for (int k = 0; k < kExtremaRows; ++k) {
        const int y = y0 + 8 * k;
        v[k] = y < e.h - e.border ? L[(size_t)y * e.w + x] : -1.f;
    }
#pragma unroll
    for (int k = 0; k < kExtremaRows; ++k) {
        const int y = y0 + 8 * k;
        if (!(v[k] > thr) || y >= e.h - e.border) continue;
        const float* c = L + (size_t)y * e.w + x;
        const float* u = c - e.w;
        const float* d = c + e.w;
        const float vv = v[k];
        if (vv <= c[-1] || vv <= c[1] || vv <= u[-1] || vv <= u[0] || vv <= u[1] || vv <= d[-1] || vv <= d[0] || vv <= d[1]) continue;
        const int i = atomicAdd(&cand_count[f], 1);
        if (i < cand_cap) {
            cand_raw[(size_t)f * cand_cap + i] = Cand{x, y, level, vv};
            atomicAdd(&row_count[(size_t)f * (total_rows + 1) + row_base + y], 1);
        }
    }
}

// exclusive scan of the per-row counts (one CTA per frame)
__global__ void __launch_bounds__(1024)
k_row_scan(const int* __restrict__ row_count, int* __restrict__ row_start, int* __restrict__ row_fill, int total_rows) {
    __shared__ int wsum[32];
    __shared__ int running;
    const int f = blockIdx.x;
    const int* rc = row_count + (size_t)f * (total_rows + 1);
    int* rs = row_start + (size_t)f * (total_rows + 1);
    int* rf = row_fill + (size_t)f * total_rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < total_rows; base += 1024) {
        const int i = base + tid;
        const int v = i < total_rows ? rc[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = wsum[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wsum[lane] = wi - w;
        }
        __syncthreads();
        const int excl = running + wsum[warp] + incl - v;
        if (i < total_rows) {
            rs[i] = excl;
            rf[i] = 0;
        }
        __syncthreads();
        if (tid == 1023) running = excl + v;
        __syncthreads();
    }
    if (tid == 0) rs[total_rows] = running;
}

__global__ void __launch_bounds__(256)
k_scatter(const Cand* __restrict__ cand_raw, Cand* __restrict__ cand, const int* __restrict__ cand_count, int cand_cap,
          const int* __restrict__ row_start, int* __restrict__ row_fill, int total_rows, LevelsDev lv, const int* __restrict__ row_base) {
    const int f = blockIdx.y;
    const int n = min(cand_count[f], cand_cap);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Cand c = cand_raw[(size_t)f * cand_cap + i];
    const int r = row_base[c.level] + c.y;
    const int pos = row_start[(size_t)f * (total_rows + 1) + r] + atomicAdd(&row_fill[(size_t)f * total_rows + r], 1);
    cand[(size_t)f * cand_cap + pos] = c;
}

// insertion sort of every row bucket by x (buckets hold a handful of entries)
__global__ void __launch_bounds__(256)
k_sort_rows(Cand* __restrict__ cand, int cand_cap, const int* __restrict__ row_start, int total_rows) {
    const int f = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= total_rows) return;
    const int* rs = row_start + (size_t)f * (total_rows + 1);
    const int a = rs[r], b = rs[r + 1];
    Cand* c = cand + (size_t)f * cand_cap;
    for (int i = a + 1; i < b; ++i) {
        const Cand v = c[i];
        int j = i - 1;
        while (j >= a && c[j].x > v.x) {
            c[j + 1] = c[j];
            --j;
        }
        c[j + 1] = v;
    }
}

struct FrameLists {
    const Cand* cand;
    const int* row_start;   // [total_rows + 1]
    LevelsDev lv;
    const int* row_base;    // [n_levels + 1]
};

// visit candidates of `level` inside the half-open window [c-r, c+r)^2 around (cx, cy) in window scan order
// (rows ascending, x ascending) that lie within L2 distance r; fn(idx) returns true to stop
template <class Fn>
__device__ __forceinline__ void visit_window(const FrameLists& fl, int level, int cx, int cy, int r, Fn fn) {
    const LevelDev& e = fl.lv.lv[level];
    // OpenCV scans the HALF-OPEN window [c - r, c + r) in both axes (a neighbour at exactly +r along
    // an axis is not seen), then tests dx^2 + dy^2 <= r^2
    const int y0 = max(cy - r, 0), y1 = min(cy + r - 1, e.h - 1);
    const int rb = fl.row_base[level];
    for (int y = y0; y <= y1; ++y) {
        const int a = fl.row_start[rb + y], b = fl.row_start[rb + y + 1];
        for (int i = a; i < b; ++i) {
            const int x = fl.cand[i].x;
            if (x < cx - r) continue;
            if (x >= cx + r) break;
            const int dx = x - cx, dy = y - cy;
            if (dx * dx + dy * dy <= r * r)
                if (fn(i)) return;
        }
    }
}

__device__ void sort_small(int* a, int n) {
    for (int i = 1; i < n; ++i) {
        const int v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) {
            a[j + 1] = a[j];
            --j;
        }
        a[j + 1] = v;
    }
}

// One CTA per frame: same-level suppression, lower pass, upper pass (see file header).
// aux layout per frame (ints): [0,cap) del_lower | [cap,2cap) del_upper | [2cap,3cap) work list
__global__ void __launch_bounds__(512)
k_suppress(const Cand* __restrict__ cand_all, const int* __restrict__ cand_count, int cand_cap,
           const int* __restrict__ row_start_all, int total_rows, LevelsDev lv, const int* __restrict__ row_base,
           unsigned char* __restrict__ state_all, int* __restrict__ aux_all) {
    __shared__ int list_n[kMaxLevels];
    const int f = blockIdx.x;
    const int n = min(cand_count[f], cand_cap);
    const Cand* cand = cand_all + (size_t)f * cand_cap;
    const int* row_start = row_start_all + (size_t)f * (total_rows + 1);
    unsigned char* alive = state_all + (size_t)f * cand_cap;
    int* del_lower = aux_all + (size_t)f * cand_cap * 3;
    int* del_upper = del_lower + cand_cap;
    int* work = del_upper + cand_cap;   // per-level lists live at [level_start, level_start + count)
    const FrameLists fl{cand, row_start, lv, row_base};
    const int tid = threadIdx.x;
    auto level_start = [&](int l) { return row_start[row_base[l]]; };

    // ---- same-level suppression ------------------------------------------------------------
    if (tid < kMaxLevels) list_n[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const Cand c = cand[i];
        bool conflicted = false;
        visit_window(fl, c.level, c.x, c.y, lv.lv[c.level].sigma_size, [&](int j) {
            if (j != i) { conflicted = true; return true; }
            return false;
        });
        alive[i] = conflicted ? 0 : 1;
        del_lower[i] = INT_MAX;
        del_upper[i] = INT_MAX;
        if (conflicted) work[level_start(c.level) + atomicAdd(&list_n[c.level], 1)] = i;
    }
    __syncthreads();
    if (tid < lv.n && list_n[tid] > 0) {
        int* lst = work + level_start(tid);
        const int m = list_n[tid];
        sort_small(lst, m);                    // row-major order = list order
        for (int k = 0; k < m; ++k) {
            const int i = lst[k];
            const Cand c = cand[i];
            int nb = -1;
            visit_window(fl, c.level, c.x, c.y, lv.lv[c.level].sigma_size, [&](int j) {
                if (j != i && alive[j]) { nb = j; return true; }
                return false;
            });
            if (nb < 0) alive[i] = 1;
            else if (c.resp > cand[nb].resp) { alive[nb] = 0; alive[i] = 1; }
        }
    }
    __syncthreads();

    // ---- cross-level passes -----------------------------------------------------------------
    // lower: keypoints of level i against level i-1;  upper: level i against level i+1
    for (int pass = 0; pass < 2; ++pass) {
        int* del = pass == 0 ? del_lower : del_upper;
        if (tid < kMaxLevels) list_n[tid] = 0;
        __syncthreads();
        auto is_alive_in = [&](int j) { return alive[j] && (pass == 0 || del_lower[j] == INT_MAX); };
        auto target = [&](const Cand& c, int& other, int& px, int& py, int& r) {
            if (pass == 0) {
                other = c.level - 1;
                if (other < 0) return false;
                const int diff = (int)lv.lv[c.level].ratio / (int)lv.lv[other].ratio;
                px = c.x * diff; py = c.y * diff;
                r = lv.lv[c.level].sigma_size * diff;
            } else {
                other = c.level + 1;
                if (other >= lv.n) return false;
                const int diff = (int)lv.lv[other].ratio / (int)lv.lv[c.level].ratio;
                px = c.x / diff; py = c.y / diff;
                r = lv.lv[other].sigma_size;
            }
            return true;
        };
        for (int i = tid; i < n; i += blockDim.x) {
            if (!is_alive_in(i)) continue;
            const Cand c = cand[i];
            int other, px, py, r;
            if (!target(c, other, px, py, r)) continue;
            int k = 0, first = -1;
            visit_window(fl, other, px, py, r, [&](int j) {
                if (is_alive_in(j)) {
                    if (k == 0) first = j;
                    ++k;
                    if (k >= 2) return true;
                }
                return false;
            });
            if (k == 1) {
                if (c.resp > cand[first].resp) atomicMin(&del[first], i);
            } else if (k >= 2) {
                work[level_start(c.level) + atomicAdd(&list_n[c.level], 1)] = i;
            }
        }
        __syncthreads();
        if (tid < lv.n && list_n[tid] > 0) {
            int* lst = work + level_start(tid);
            const int m = list_n[tid];
            sort_small(lst, m);
            for (int k = 0; k < m; ++k) {
                const int i = lst[k];
                const Cand c = cand[i];
                int other, px, py, r;
                target(c, other, px, py, r);
                int nb = -1;
                visit_window(fl, other, px, py, r, [&](int j) {
                    // present at c's turn: alive on entry and not deleted by an earlier keypoint
                    if (is_alive_in(j) && !(atomicAdd(&del[j], 0) < i)) { nb = j; return true; }
                    return false;
                });
                if (nb >= 0 && c.resp > cand[nb].resp) atomicMin(&del[nb], i);
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += blockDim.x)
        alive[i] = alive[i] && del_lower[i] == INT_MAX && del_upper[i] == INT_MAX;
}

// sub-pixel refinement + ordered compaction (Do_Subpixel_Refinement); one CTA per frame
__global__ void __launch_bounds__(1024)
k_refine(const Cand* __restrict__ cand_all, const int* __restrict__ cand_count, int cand_cap,
         const unsigned char* __restrict__ state_all, const float* __restrict__ Ldet, size_t pyr_stride, LevelsDev lv,
         DunkKeyPoint* __restrict__ kps_all, int kp_cap, int* __restrict__ kp_count) {
    __shared__ int wsum[32];
    __shared__ int running, chunk_total;
    const int f = blockIdx.x;
    const int n = min(cand_count[f], cand_cap);
    const Cand* cand = cand_all + (size_t)f * cand_cap;
    const unsigned char* alive = state_all + (size_t)f * cand_cap;
    DunkKeyPoint* kps = kps_all + (size_t)f * kp_cap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        bool keep = false;
        DunkKeyPoint kp;
        if (i < n && alive[i]) {
            const Cand c = cand[i];
            const LevelDev& e = lv.lv[c.level];
            const float* L = Ldet + (size_t)f * pyr_stride + e.plane_off + (size_t)c.y * e.w + c.x;
            const int w = e.w;
            const float v = L[0];
            const float Dx = __fmul_rn(0.5f, __fsub_rn(L[1], L[-1]));
            const float Dy = __fmul_rn(0.5f, __fsub_rn(L[w], L[-w]));
            const float Dxx = __fsub_rn(__fadd_rn(L[1], L[-1]), __fmul_rn(2.0f, v));
            const float Dyy = __fsub_rn(__fadd_rn(L[w], L[-w]), __fmul_rn(2.0f, v));
            const float Dxy = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(L[w + 1], L[-w - 1]), L[-w + 1]), L[w - 1]));
            // Matx22f solve (LU, 2x2 closed form): d = 1/det; x = (b0*a11 - b1*a01)*d ...
            float det = __fsub_rn(__fmul_rn(Dxx, Dyy), __fmul_rn(Dxy, Dxy));
            float dx = 0.f, dy = 0.f;
            if (det != 0.f) {
                det = __fdiv_rn(1.f, det);
                const float b0 = -Dx, b1 = -Dy;
                dx = __fmul_rn(__fsub_rn(__fmul_rn(b0, Dyy), __fmul_rn(b1, Dxy)), det);
                dy = __fmul_rn(__fsub_rn(__fmul_rn(b1, Dxx), __fmul_rn(b0, Dxy)), det);
            }
            if (fabsf(dx) <= 1.0f && fabsf(dy) <= 1.0f) {
                keep = true;
                const float ratio = e.ratio;
                const float half = __fmul_rn(0.5f, __fsub_rn(ratio, 1.0f));
                kp.x = __fadd_rn(__fadd_rn(__fmul_rn((float)c.x, ratio), __fmul_rn(dx, ratio)), half);
                kp.y = __fadd_rn(__fadd_rn(__fmul_rn((float)c.y, ratio), __fmul_rn(dy, ratio)), half);
                kp.size = __fmul_rn(2.0f, __fmul_rn(e.esigma, 1.5f));
                kp.angle = -1.f;
                kp.response = v;
                kp.octave = e.octave;
                kp.class_id = c.level;
            }
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
            const int pos = running + wsum[warp] + __popc(bal & ((1u << lane) - 1u));
            if (pos < kp_cap) kps[pos] = kp;
        }
        __syncthreads();
        if (tid == 0) running += chunk_total;
        __syncthreads();
    }
    if (tid == 0) kp_count[f] = min(running, kp_cap);
}

// max_points: keep the strongest K (std::partial_sort by response, descending); bitonic sort of
// (response desc, index asc) keys in global memory by one CTA per frame
__global__ void __launch_bounds__(1024)
k_top_k(DunkKeyPoint* __restrict__ kps_all, int kp_cap, int* __restrict__ kp_count, int max_points,
        unsigned long long* __restrict__ keys_all, DunkKeyPoint* __restrict__ tmp_all, int pow2_cap) {
    const int f = blockIdx.x;
    const int n = kp_count[f];
    if (n <= max_points) return;
    DunkKeyPoint* kps = kps_all + (size_t)f * kp_cap;
    DunkKeyPoint* tmp = tmp_all + (size_t)f * kp_cap;
    unsigned long long* keys = keys_all + (size_t)f * pow2_cap;
    int p2 = 1;
    while (p2 < n) p2 <<= 1;
    for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < n) k = ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(kps[i].response)) << 32) | (unsigned)i;
        keys[i] = k;
        if (i < n) tmp[i] = kps[i];
    }
    __syncthreads();
    for (int k = 2; k <= p2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], b = keys[l];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < max_points; i += blockDim.x) kps[i] = tmp[(unsigned)keys[i]];
    __syncthreads();
    if (threadIdx.x == 0) kp_count[f] = max_points;
}

}  // namespace

int akaze_detect(dunk_ctx* ctx, cudaStream_t st, const LevelTable& lt, const AkazeWorkspace& ws, int frames,
                 float dthreshold, int max_points) {
    const LevelsDev lv = make_levels_dev(lt);
    const size_t pyr = lt.pyramid_floats;
    // row bases (host) -> device copy lives at the tail of row_fill's allocation? keep it simple:
    int row_base_h[kMaxLevels + 1];
    int acc = 0;
    for (int i = 0; i < lt.n_levels; ++i) { row_base_h[i] = acc; acc += lt.lv[i].h; }
    row_base_h[lt.n_levels] = acc;
    int* row_base_d = ws.kp_level_start;   // (kMaxLevels + 1) ints reserved in the workspace
    DUNK_CUDA(cudaMemcpyAsync(row_base_d, row_base_h, sizeof(int) * (lt.n_levels + 1), cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemsetAsync(ws.cand_count, 0, (size_t)frames * 4, st));
    DUNK_CUDA(cudaMemsetAsync(ws.row_count, 0, (size_t)frames * (ws.total_rows + 1) * 4, st));
    for (int i = 0; i < lt.n_levels; ++i) {
        const LevelInfo& e = lt.lv[i];
        if (e.border + 1 >= e.h) continue;   // FindKeypointsSameScale: border too big
        const int iw = e.w - 2 * e.border, ih = e.h - 2 * e.border;
        if (iw <= 0 || ih <= 0) continue;
        {
            ProfScope ps(ctx, st, "detect.extrema", (double)frames * iw * ih * 4);
            k_extrema<<<dim3(div_up(iw, 32), div_up(ih, 8 * kExtremaRows), frames), dim3(32, 8), 0, st>>>(
                ws.Ldet, pyr, lv.lv[i], i, row_base_h[i], dthreshold, ws.cand_raw, ws.cand_cap, ws.cand_count, ws.row_count,
                ws.total_rows);
            DUNK_KERNEL_CHECK(ctx);
        }
    }
    {
        ProfScope ps(ctx, st, "detect.row_scan", 0.0);
        k_row_scan<<<frames, 1024, 0, st>>>(ws.row_count, ws.row_start, ws.row_fill, ws.total_rows);
        DUNK_KERNEL_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, st, "detect.scatter", 0.0);
        k_scatter<<<dim3(div_up(ws.cand_cap, 256), frames), 256, 0, st>>>(ws.cand_raw, ws.cand, ws.cand_count, ws.cand_cap,
                                                                          ws.row_start, ws.row_fill, ws.total_rows, lv, row_base_d);
        DUNK_KERNEL_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, st, "detect.sort_rows", 0.0);
        k_sort_rows<<<dim3(div_up(ws.total_rows, 256), frames), 256, 0, st>>>(ws.cand, ws.cand_cap, ws.row_start, ws.total_rows);
        DUNK_KERNEL_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, st, "detect.suppress", 0.0);
        k_suppress<<<frames, 512, 0, st>>>(ws.cand, ws.cand_count, ws.cand_cap, ws.row_start, ws.total_rows, lv, row_base_d,
                                           ws.state, ws.aux);
        DUNK_KERNEL_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, st, "detect.refine", 0.0);
        k_refine<<<frames, 1024, 0, st>>>(ws.cand, ws.cand_count, ws.cand_cap, ws.state, ws.Ldet, pyr, lv, ws.kps, ws.kp_cap,
                                          ws.kp_count);
        DUNK_KERNEL_CHECK(ctx);
    }
    if (max_points > 0 && max_points < ws.kp_cap) {
        int p2 = 1;
        while (p2 < ws.kp_cap) p2 <<= 1;
        // keys: reuse aux (3*cand_cap ints >= p2 u64 when cand_cap >= kp_cap); tmp: reuse cand_raw
        k_top_k<<<frames, 1024, 0, st>>>(ws.kps, ws.kp_cap, ws.kp_count, max_points, (unsigned long long*)ws.sort_keys,
                                         (DunkKeyPoint*)ws.desc64, p2);
        DUNK_KERNEL_CHECK(ctx);
    }
    return DUNK_OK;
}

}  // namespace dunk
