// Stage 2 of the hot path: brute-force Hamming 2-NN over 61-byte MLDB descriptors
// (replaces BFMatcher(NORM_HAMMING).knnMatch/match, feature_extraction/src/lib.rs:94-126).
//
// Layout in HBM: descriptors are 64-byte rows (16 x u32, bytes 61..63 zero).
// Kernel: lane = query (QT queries per thread held in registers), DB rows are streamed
// HBM -> smem with 1-D bulk async copies (cp.async.bulk + mbarrier, SASS UBLKCP) and read
// by every lane as broadcast LDS.128.  Each thread keeps a register-resident top-2 for its
// queries; rows are visited in increasing index order with strict '<', so a thread's top-2
// is exactly the two lexicographically smallest (distance, index) pairs of its slab — the
// stable-argsort rule OpenCV's batchDistance follows.  Slabs (grid.x)
// are merged by (distance, index) in a second tiny kernel; the same merge serves the
// multi-GPU shards after the allgather.
#include <algorithm>
#include <cstdlib>
#include "match.h"

namespace dunk {

namespace {

constexpr int kTileRows = 64;            // DB rows per smem stage (4 KB), one ring per warp
constexpr int kStages = 3;
constexpr int kWarpsPerCta = 8;
constexpr int kQT = 4;                   // queries per thread
#ifndef DUNK_MATCH_UNROLL
#define DUNK_MATCH_UNROLL 2
#endif
constexpr int kRowUnroll = DUNK_MATCH_UNROLL;   // DB rows per unrolled loop body
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void top2_insert_stream(uint32_t d, uint32_t gi, uint32_t& d1,
                                                   uint32_t& i1, uint32_t& d2, uint32_t& i2) {
    // rows arrive in increasing index order: strict '<' keeps the lowest index on ties
    if (d < d2) {
        if (d < d1) {
            d2 = d1; i2 = i1; d1 = d; i1 = gi;
        } else {
            d2 = d; i2 = gi;
        }
    }
}

__device__ __forceinline__ bool lex_less(uint32_t da, uint32_t ia, uint32_t db, uint32_t ib) {
    return da < db || (da == db && ia < ib);
}
__device__ __forceinline__ void top2_insert_lex(uint32_t d, uint32_t gi, uint32_t& d1, uint32_t& i1,
                                                uint32_t& d2, uint32_t& i2) {
    if (lex_less(d, gi, d2, i2)) {
        if (lex_less(d, gi, d1, i1)) {
            d2 = d1; i2 = i1; d1 = d; i1 = gi;
        } else {
            d2 = d; i2 = gi;
        }
    }
}

// carry-save adder on 32 bit lanes: (a, b, c) -> sum (weight 1) and carry (weight 2); 2 LOP3
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    // one LOP3 each: 0x96 = a ^ b ^ c, 0xE8 = majority(a, b, c)
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(sum) : "r"(a), "r"(b), "r"(c));
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(carry) : "r"(a), "r"(b), "r"(c));
}

// Hamming distance of two 512-bit rows.  The algorithmic unit of the matcher roofline is
// 16 x (XOR, POPC); POPC issues at 16 lanes/clk/SM against 64 for LOP3, so kCsa
// carry-save adders (Harley-Seal style) first compress the 16 XOR words to 16 - kCsa weighted
// words: the POPC pipe and the ALU pipe then carry about the same load.
#ifndef DUNK_MATCH_CSA
#define DUNK_MATCH_CSA 7
#endif
__device__ __forceinline__ uint32_t hamming512(const uint32_t (&q)[16], const uint4& a,
                                               const uint4& b, const uint4& c, const uint4& d) {
    uint32_t x[16];
    x[0] = q[0] ^ a.x; x[1] = q[1] ^ a.y; x[2] = q[2] ^ a.z; x[3] = q[3] ^ a.w;
    x[4] = q[4] ^ b.x; x[5] = q[5] ^ b.y; x[6] = q[6] ^ b.z; x[7] = q[7] ^ b.w;
    x[8] = q[8] ^ c.x; x[9] = q[9] ^ c.y; x[10] = q[10] ^ c.z; x[11] = q[11] ^ c.w;
    x[12] = q[12] ^ d.x; x[13] = q[13] ^ d.y; x[14] = q[14] ^ d.z; x[15] = q[15] ^ d.w;
#if DUNK_MATCH_CSA == 0
    uint32_t s0 = __popc(x[0]) + __popc(x[1]) + __popc(x[2]) + __popc(x[3]);
    uint32_t s1 = __popc(x[4]) + __popc(x[5]) + __popc(x[6]) + __popc(x[7]);
    uint32_t s2 = __popc(x[8]) + __popc(x[9]) + __popc(x[10]) + __popc(x[11]);
    uint32_t s3 = __popc(x[12]) + __popc(x[13]) + __popc(x[14]) + __popc(x[15]);
    return (s0 + s1) + (s2 + s3);
#else
    // level 1: five CSAs over x0..x14 (x15 left over)
    uint32_t s[5], t[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) csa(x[3 * i], x[3 * i + 1], x[3 * i + 2], s[i], t[i]);
    // level 2: one CSA on the weight-1 sums, one on the weight-2 carries
    uint32_t s5, t5, u0, f0;
    csa(s[0], s[1], s[2], s5, t5);
    csa(t[0], t[1], t[2], u0, f0);
#if DUNK_MATCH_CSA >= 8
    uint32_t s6, t6;
    csa(s5, s[3], s[4], s6, t6);
    const uint32_t w1 = __popc(s6) + __popc(x[15]);
    const uint32_t w2 = (__popc(u0) + __popc(t[3])) + (__popc(t[4]) + __popc(t5)) + __popc(t6);
#else
    const uint32_t w1 = (__popc(s5) + __popc(s[3])) + (__popc(s[4]) + __popc(x[15]));
    const uint32_t w2 = (__popc(u0) + __popc(t[3])) + (__popc(t[4]) + __popc(t5));
#endif
    return w1 + 2u * w2 + 4u * __popc(f0);
#endif
}

// Lower bound of the distance from the first 15 words (480 bits): d480 <= d, so a row whose d480 is not below a list's
// second-best distance cannot enter it (strict '<'), and the 16th word -- 6 bits of a 486-bit MLDB row -- is only
// looked at on the rare rows that pass the vote.  One XOR and one POPC less per pair than hamming512: 15 XOR + 7
// carry-save adders (14 LOP3) + 8 POPC, ALU 33/64 and XU 8/16 clocks per pair and SM.  Exact for any row content.
#ifndef DUNK_MATCH_LAZY15
#define DUNK_MATCH_LAZY15 1
#endif
__device__ __forceinline__ uint32_t hamming480(const uint32_t (&q)[16], const uint4& a, const uint4& b, const uint4& c,
                                               const uint4& d) {
    uint32_t x[15];
    x[0] = q[0] ^ a.x; x[1] = q[1] ^ a.y; x[2] = q[2] ^ a.z; x[3] = q[3] ^ a.w;
    x[4] = q[4] ^ b.x; x[5] = q[5] ^ b.y; x[6] = q[6] ^ b.z; x[7] = q[7] ^ b.w;
    x[8] = q[8] ^ c.x; x[9] = q[9] ^ c.y; x[10] = q[10] ^ c.z; x[11] = q[11] ^ c.w;
    x[12] = q[12] ^ d.x; x[13] = q[13] ^ d.y; x[14] = q[14] ^ d.z;
    uint32_t s[5], t[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) csa(x[3 * i], x[3 * i + 1], x[3 * i + 2], s[i], t[i]);
    uint32_t s5, t5, u0, f0;
    csa(s[0], s[1], s[2], s5, t5);
    csa(t[0], t[1], t[2], u0, f0);
    const uint32_t w1 = (__popc(s5) + __popc(s[3])) + __popc(s[4]);
    const uint32_t w2 = (__popc(u0) + __popc(t[3])) + (__popc(t[4]) + __popc(t5));
    return w1 + 2u * w2 + 4u * __popc(f0);
}

// One warp = one work item (query group of 32*QT queries, DB slab): the warp streams its slab through
// its own 3-stage smem ring (lane 0 issues the bulk copies, all lanes wait on the mbarrier) and keeps
// QT register-resident top-2 lists per lane.  Warps never synchronise with each other, so every
// CTA has 8 busy warps (2 per SM sub-partition) whatever the query count is.
template <int QT>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
hamming_top2_kernel(const uint4* __restrict__ db, uint32_t nt, const uint4* __restrict__ q, int nq,
                    int n_qgroups, int n_slabs, int tiles_per_slab, uint32_t index_base,
                    uint4* __restrict__ partial, const int* __restrict__ seg_counts, int seg_len) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (item >= (long long)n_qgroups * n_slabs) return;
    const int qg = (int)(item % n_qgroups), slab = (int)(item / n_qgroups);
    // sharded pipeline: the query array is W segments of seg_len rows, segment r holding seg_counts[2 r] real rows and
    // padding behind them; a query group that lies wholly in the padding of one segment has nothing to match
    if (seg_counts) {
        const int a = qg * 32 * QT, seg = a / seg_len;
        if ((a + 32 * QT - 1) / seg_len == seg && a - seg * seg_len >= seg_counts[2 * seg]) return;
    }
    uint4* tiles = reinterpret_cast<uint4*>(smem_raw) + (size_t)warp * kStages * kTileRows * 4;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kWarpsPerCta * kStages * kTileRows * 64) +
                     warp * kStages;

    const int q0 = (qg * 32 + lane) * QT;   // first query of this lane (queries are lane-contiguous)
    uint32_t qr[QT][16];
    uint32_t d1[QT], i1[QT], d2[QT], i2[QT];
#pragma unroll
    for (int s = 0; s < QT; ++s) {
        d1[s] = d2[s] = kEmpty;
        i1[s] = i2[s] = kEmpty;
        if (q0 + s < nq) {
            const uint4* p = q + (size_t)(q0 + s) * 4;
            uint4 a = p[0], b = p[1], c = p[2], d = p[3];
            qr[s][0] = a.x; qr[s][1] = a.y; qr[s][2] = a.z; qr[s][3] = a.w;
            qr[s][4] = b.x; qr[s][5] = b.y; qr[s][6] = b.z; qr[s][7] = b.w;
            qr[s][8] = c.x; qr[s][9] = c.y; qr[s][10] = c.z; qr[s][11] = c.w;
            qr[s][12] = d.x; qr[s][13] = d.y; qr[s][14] = d.z; qr[s][15] = d.w;
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) qr[s][j] = 0u;
        }
    }

    const int total_tiles = (int)((nt + kTileRows - 1) / kTileRows);
    const int t0 = slab * tiles_per_slab;
    const int ntile = max(0, min(t0 + tiles_per_slab, total_tiles) - t0);

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto issue = [&](int k) {
        const int st = k % kStages;
        const uint32_t row0 = (uint32_t)(t0 + k) * kTileRows;
        const uint32_t rows = min((uint32_t)kTileRows, nt - row0);
        mbar_expect_tx(&full[st], rows * 64u);
        bulk_g2s(tiles + (size_t)st * kTileRows * 4, db + (size_t)row0 * 4, rows * 64u, &full[st]);
    };
    if (lane == 0)
        for (int k = 0; k < min(kStages, ntile); ++k) issue(k);

    for (int k = 0; k < ntile; ++k) {
        const int st = k % kStages;
        mbar_wait(&full[st], (uint32_t)(k / kStages) & 1u);
        const uint32_t row0 = (uint32_t)(t0 + k) * kTileRows;
        const int rows = (int)min((uint32_t)kTileRows, nt - row0);
        const uint4* tp = tiles + (size_t)st * kTileRows * 4;
        const uint32_t g0 = index_base + row0;
#pragma unroll kRowUnroll
        for (int r = 0; r < rows; ++r) {
            const uint4 a = tp[r * 4 + 0], b = tp[r * 4 + 1], c = tp[r * 4 + 2], d = tp[r * 4 + 3];
            uint32_t dist[QT];
            bool hit = false;
#pragma unroll
            for (int s = 0; s < QT; ++s) {
#if DUNK_MATCH_LAZY15
                dist[s] = hamming480(qr[s], a, b, c, d);
#else
                dist[s] = hamming512(qr[s], a, b, c, d);
#endif
                hit |= dist[s] < d2[s];
            }
            // a row improves some lane's top-2 with probability ~256/rows_seen: keep the common
            // path to QT compares + one vote + one warp-uniform branch
            if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
                for (int s = 0; s < QT; ++s) {
#if DUNK_MATCH_LAZY15
                    dist[s] += __popc(qr[s][15] ^ d.w);     // the exact distance, only where it can matter
#endif
                    top2_insert_stream(dist[s], g0 + r, d1[s], i1[s], d2[s], i2[s]);
                }
            }
        }
        __syncwarp();  // every lane is done with stage st before it is refilled
        if (lane == 0 && k + kStages < ntile) issue(k + kStages);
    }

#pragma unroll
    for (int s = 0; s < QT; ++s)
        if (q0 + s < nq)
            partial[(size_t)slab * nq + (q0 + s)] = make_uint4(d1[s], i1[s], d2[s], i2[s]);
}

// lexicographic (distance, index) merge of n_parts top-2 arrays (part-major)
__global__ void top2_merge_kernel(const uint4* __restrict__ parts, int n_parts, long long part_stride, int nq,
                                  uint4* __restrict__ out) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    uint32_t d1 = kEmpty, i1 = kEmpty, d2 = kEmpty, i2 = kEmpty;
    for (int p = 0; p < n_parts; ++p) {
        const uint4 v = parts[(size_t)p * part_stride + qi];
        top2_insert_lex(v.x, v.y, d1, i1, d2, i2);
        top2_insert_lex(v.z, v.w, d1, i1, d2, i2);
    }
    out[qi] = make_uint4(d1, i1, d2, i2);
}

// ordered compaction by a single CTA: keep(qi, m) decides and fills the DMatch
template <class Keep>
__device__ __forceinline__ void ordered_compact(int nq, DunkDMatch* __restrict__ out,
                                                int* __restrict__ count, Keep keep_fn) {
    __shared__ int wsum[32];
    __shared__ int running, chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < nq; base += blockDim.x) {
        const int qi = base + tid;
        DunkDMatch m;
        const bool keep = qi < nq && keep_fn(qi, m);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            const int x = (lane < (int)(blockDim.x >> 5)) ? wsum[lane] : 0;
            int incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            wsum[lane] = incl - x;
            if (lane == 31) chunk_total = incl;
        }
        __syncthreads();
        if (keep) out[running + wsum[warp] + __popc(bal & ((1u << lane) - 1u))] = m;
        __syncthreads();
        if (tid == 0) running += chunk_total;
        __syncthreads();
    }
    if (tid == 0) *count = running;
}

// Lowe ratio test (f32, strict '<', lib.rs:107-111) + ordered compaction
__global__ void __launch_bounds__(1024)
top2_ratio_compact_kernel(const uint4* __restrict__ top2, int nq, float ratio,
                          DunkDMatch* __restrict__ out, int* __restrict__ count) {
    ordered_compact(nq, out, count, [&](int qi, DunkDMatch& m) {
        const uint4 v = top2[qi];
        m.query_idx = qi;
        m.train_idx = (int)v.y;
        m.img_idx = 0;
        m.distance = (float)v.x;
        return v.z != kEmpty && ((float)v.x < __fmul_rn((float)v.z, ratio));
    });
}

// cross-check (BFMatcher crossCheck=true, cv2 4.13): keep (q, i) iff i = nearest train row of
// q and q = nearest query row of i, both with lowest-index tie-breaks
__global__ void __launch_bounds__(1024)
crosscheck_compact_kernel(const uint4* __restrict__ q2t, const uint4* __restrict__ t2q, int nq,
                          DunkDMatch* __restrict__ out, int* __restrict__ count) {
    ordered_compact(nq, out, count, [&](int qi, DunkDMatch& m) {
        const uint4 v = q2t[qi];
        m.query_idx = qi;
        m.train_idx = (int)v.y;
        m.img_idx = 0;
        m.distance = (float)v.x;
        return v.y != kEmpty && t2q[v.y].y == (uint32_t)qi;
    });
}

// n x desc_bytes (byte rows) <-> n x 64-B rows
__global__ void pad_rows_kernel(const uint8_t* __restrict__ src, long long n, int desc_bytes,
                                uint32_t* __restrict__ dst) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // output word
    if (w >= n * 16) return;
    const long long row = w >> 4;
    const int j = (int)(w & 15) * 4;
    const uint8_t* p = src + row * desc_bytes;
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (j + b < desc_bytes) v |= (uint32_t)p[j + b] << (8 * b);
    dst[w] = v;
}
__global__ void unpad_rows_kernel(const uint8_t* __restrict__ src64, long long n, int desc_bytes,
                                  uint8_t* __restrict__ dst) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n * desc_bytes) return;
    const long long row = b / desc_bytes;
    const int j = (int)(b - row * desc_bytes);
    dst[b] = src64[row * 64 + j];
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// synthetic DB rows (bench config 3): word pair j of row r = splitmix64(seed + r*8 + j);
// byte 60 keeps 6 bits, bytes 61..63 are zero — like a real MLDB-486 row
__global__ void fill_random_rows_kernel(uint2* __restrict__ dst, long long n, unsigned long long seed,
                                        unsigned long long row_offset) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // u64 index
    if (w >= n * 8) return;
    const unsigned long long row = row_offset + (unsigned long long)(w >> 3);
    const int j = (int)(w & 7);
    unsigned long long v = splitmix64(seed + row * 8ull + (unsigned long long)j);
    if (j == 7) v &= 0x0000003FFFFFFFFFull;
    dst[w] = make_uint2((uint32_t)v, (uint32_t)(v >> 32));
}

}  // namespace

static constexpr size_t kHammingSmem = (size_t)kWarpsPerCta * kStages * kTileRows * 64 + kWarpsPerCta * kStages * sizeof(uint64_t);

// per-device opt-in to ~96 KB of dynamic shared memory + the occupancy the planner sizes its rounds with
static int hamming_device_init(dunk_ctx* ctx) {
    DUNK_CUDA(cudaFuncSetAttribute(hamming_top2_kernel<kQT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHammingSmem));
    int occ = 0;
    DUNK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hamming_top2_kernel<kQT>, kWarpsPerCta * 32, kHammingSmem));
    ctx->hamming_occ = std::max(occ, 1);
    return DUNK_OK;
}
static DeviceInitReg hamming_device_init_reg(hamming_device_init);

KnnPlan plan_knn2(dunk_ctx* ctx, int nq, uint32_t nt) {
    KnnPlan p;
    p.threads = kWarpsPerCta * 32;
    p.gy = div_up(nq, 32 * kQT);          // query groups (one warp each)
    p.smem = (size_t)kWarpsPerCta * kStages * kTileRows * 64 + kWarpsPerCta * kStages * sizeof(uint64_t);
    const int total_tiles = div_up(nt, kTileRows);
    const int occ = std::max(1, ctx->hamming_occ);      // per device, set by hamming_device_init
    // equal-sized slabs: (query groups x slabs) work items on `slots` resident warps.  All items cost the same,
    // so the launch takes ceil(items / slots) rounds; pick the slab count whose last round is fullest
    // (e.g. 1233 query groups x 4 slabs would be 2.08 rounds = 69 % efficiency, x 19 slabs is 9.9 rounds = 99 %).
    // Slabs stay >= 32 tiles (2048 rows) so the per-item set-up (64 query registers, barrier ring) is amortised.
    const long long slots = (long long)ctx->sm_count * occ * kWarpsPerCta;
    const long long max_slabs = std::max<long long>(1, std::min<long long>(total_tiles / 32, 4096));
    long long slabs = 1;
    double best_eff = -1.0;
    for (long long s = 1; s <= max_slabs; ++s) {
        const long long tps = div_up(total_tiles, s);
        const long long real_slabs = div_up(total_tiles, tps);
        const long long items = real_slabs * p.gy;
        const long long rounds = (items + slots - 1) / slots;
        // measured (3163 x 4 M rows): 1 full round 418 Gpairs/s, 2-6 full rounds 428-431 (warps that finish are
        // replaced while others still run, so pipeline fill / drain overlaps), a barely started extra round 287
        const double eff = (double)items / (double)(rounds * slots) * (rounds >= 3 ? 1.0 : 0.97);
        if (eff > best_eff + 0.005) { best_eff = eff; slabs = real_slabs; }
        if (items > 16 * slots) break;
    }
    static const int force_slabs = getenv("DUNK_MATCH_SLABS") ? atoi(getenv("DUNK_MATCH_SLABS")) : 0;      // A/B switch
    if (force_slabs > 0) slabs = std::min<long long>(force_slabs, max_slabs);
    p.tiles_per_cta = div_up(total_tiles, slabs);
    if (p.tiles_per_cta < 1) p.tiles_per_cta = 1;
    p.gx = div_up(total_tiles, p.tiles_per_cta);   // slabs
    if (p.gx < 1) p.gx = 1;
    return p;
}

// upper bound of plan.gx * nq * 16 over every query count <= nq_max (callers that size a workspace before the
// query count is known): the planner never creates more than max(query groups, 16 * resident warps) work items
size_t knn2_partial_bound(dunk_ctx* ctx, size_t nq_max) {
    const int occ = ctx->hamming_occ;
    const size_t slots = (size_t)ctx->sm_count * (size_t)std::max(occ, 1) * kWarpsPerCta;
    const size_t groups = (nq_max + 32 * kQT - 1) / (32 * kQT);
    return std::max(groups, 16 * slots + groups) * (32 * kQT) * 16;
}

#define DUNK_LAUNCH_CHECK(ctx)                                                        \
    do {                                                                              \
        (ctx)->launches.fetch_add(1);                                                 \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                      \
            return DUNK_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

int launch_pad_rows(dunk_ctx* ctx, cudaStream_t st, const uint8_t* src, int64_t n, int desc_bytes,
                    uint4* dst64) {
    if (n <= 0) return DUNK_OK;
    const long long words = n * 16;
    pad_rows_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(src, n, desc_bytes,
                                                                      (uint32_t*)dst64);
    DUNK_LAUNCH_CHECK(ctx);
    return DUNK_OK;
}

int launch_unpad_rows(dunk_ctx* ctx, cudaStream_t st, const uint4* src64, int64_t n, int desc_bytes,
                      uint8_t* dst) {
    if (n <= 0) return DUNK_OK;
    const long long bytes = n * desc_bytes;
    unpad_rows_kernel<<<(unsigned)((bytes + 255) / 256), 256, 0, st>>>((const uint8_t*)src64, n,
                                                                        desc_bytes, dst);
    DUNK_LAUNCH_CHECK(ctx);
    return DUNK_OK;
}

int launch_knn2(dunk_ctx* ctx, cudaStream_t st, const uint4* db64, uint32_t nt, const uint4* q64,
                int nq, uint32_t index_base, uint4* partial, uint4* top2_out, const KnnPlan& p, const int* seg_counts,
                int seg_len) {
    if (nq <= 0) return DUNK_OK;
    // with a single slab the kernel's partial IS the result
    uint4* dst = (p.gx == 1) ? top2_out : partial;
    {
        ProfScope ps(ctx, st, "match.hamming_top2", (double)nq * (double)nt);   // pairs
        const long long items = (long long)p.gx * p.gy;
        hamming_top2_kernel<kQT><<<(unsigned)div_up(items, kWarpsPerCta), p.threads, p.smem, st>>>(
            db64, nt, q64, nq, p.gy, p.gx, p.tiles_per_cta, index_base, dst, seg_len > 0 ? seg_counts : nullptr, seg_len);
        DUNK_LAUNCH_CHECK(ctx);
    }
    if (p.gx > 1) return launch_top2_merge(ctx, st, partial, p.gx, nq, top2_out);
    return DUNK_OK;
}

int launch_top2_merge(dunk_ctx* ctx, cudaStream_t st, const uint4* parts, int n_parts, int nq,
                      uint4* out) {
    if (nq <= 0) return DUNK_OK;
    ProfScope ps(ctx, st, "match.top2_merge", (double)nq * n_parts * 16);
    top2_merge_kernel<<<div_up(nq, 128), 128, 0, st>>>(parts, n_parts, (long long)nq, nq, out);
    DUNK_LAUNCH_CHECK(ctx);
    return DUNK_OK;
}

int launch_top2_merge_strided(dunk_ctx* ctx, cudaStream_t st, const uint4* parts, int n_parts, long long part_stride,
                              int nq, uint4* out) {
    if (nq <= 0) return DUNK_OK;
    top2_merge_kernel<<<div_up(nq, 128), 128, 0, st>>>(parts, n_parts, part_stride, nq, out);
    DUNK_LAUNCH_CHECK(ctx);
    return DUNK_OK;
}

int launch_top2_ratio(dunk_ctx* ctx, cudaStream_t st, const uint4* top2, int nq, float ratio,
                      DunkDMatch* out, int* count) {
    top2_ratio_compact_kernel<<<1, 1024, 0, st>>>(top2, nq, ratio, out, count);
    DUNK_LAUNCH_CHECK(ctx);
    return DUNK_OK;
}

int launch_crosscheck(dunk_ctx* ctx, cudaStream_t st, const uint4* q2t_top2, const uint4* t2q_top2,
                      int nq, DunkDMatch* out, int* count) {
    crosscheck_compact_kernel<<<1, 1024, 0, st>>>(q2t_top2, t2q_top2, nq, out, count);
    DUNK_LAUNCH_CHECK(ctx);
    return DUNK_OK;
}

int launch_fill_random_rows(dunk_ctx* ctx, cudaStream_t st, uint4* dst64, int64_t n, uint64_t seed,
                            uint64_t row_offset) {
    if (n <= 0) return DUNK_OK;
    const long long words = n * 8;
    fill_random_rows_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>((uint2*)dst64, n, seed,
                                                                              row_offset);
    DUNK_LAUNCH_CHECK(ctx);
    return DUNK_OK;
}

}  // namespace dunk
