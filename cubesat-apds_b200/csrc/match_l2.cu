// Float-descriptor matcher (north_star: "float descriptors use an L2 decomposition whose dot products
// run on tcgen05 tensor cores (the only dense contraction on the path)"; SURVEY 8d roofline row).
// Replaces cv::BFMatcher(NORM_L2).knnMatch(query, train, 2) for f32 descriptors of 64 or 128 floats
// (KAZE / SURF / SIFT shapes).  The reference itself only ever builds NORM_HAMMING matchers
// (feature_extraction/src/lib.rs:101,121), so this entry point has no call site there.
//
//   ||q - t||^2 = ||q||^2 + ||t||^2 - 2 q.t
//
// Stage A (tensor cores): one CTA = (128 queries) x (a slab of train rows).  Warp 0 streams 128-byte
// swizzled K-major tiles with TMA (cp.async.bulk.tensor) through an mbarrier ring, one elected
// thread of warp 1 issues tcgen05.mma.kind::tf32 (M 128, N 256 or 128, K 8) into a double-buffered
// TMEM accumulator, 16 epilogue warps (4 per TMEM lane quarter, one column group each) pull the
// accumulators back with tcgen05.ld (one TMEM lane = one query per thread) and keep the 4 smallest
// scores ||t||^2 - 2 q.t per query and column group.  TF32 keeps 10
// mantissa bits of the operands, so stage A only NOMINATES candidates.
// Stage B (CUDA cores): one warp per query recomputes the exact f32 distance of every nominated row
// (4 per slab), takes the two lexicographically smallest (distance, index) pairs, and proves the
// answer: every row stage A dropped scored >= the slab's 4th candidate, and |tf32 score - exact score|
// <= margin, so if the exact 2nd best beats (4th candidate - margin) of every slab nothing was
// missed.  Queries that cannot be proven are re-done by an exact brute-force kernel (rare).
#include <cuda.h>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "match.h"

namespace dunk {
namespace {

constexpr int kM = 128;            // queries per CTA = UMMA M = TMEM lanes
constexpr int kKBlock = 32;        // floats per 128-byte swizzle row
constexpr int kUmmaK = 8;          // K per tcgen05.mma.kind::tf32
// Tile shape at D = 64, measured with the chunked epilogue below (3 163 x 2 M rows, tf32 TFLOP/s of the main stage):
// N 256 x 2 accumulators x 2 smem stages 506; N 128 x 4 accumulators x 5 stages 467-476 (the single MMA-issuing thread
// and the barrier hand-shakes per tile weigh twice as much); with the epilogue arithmetic skipped 711 / 552.
#ifndef DUNK_L2_ACC
#define DUNK_L2_ACC 2
#endif
#ifndef DUNK_L2_N64
#define DUNK_L2_N64 256
#endif
constexpr int kAcc = DUNK_L2_ACC;  // TMEM accumulator ring
constexpr int kCand = 4;           // candidates kept per (query, slab)
constexpr int kSub = 4;            // epilogue warps per TMEM lane quarter = column groups = candidate lists per (query, slab)
constexpr int kEpiWarps = 4 * kSub;
constexpr int kThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA + TMEM owner, the rest epilogue
constexpr uint32_t kNoIdx = 0xFFFFFFFFu;

template <int D>
struct Cfg {
    static constexpr int KB = D / kKBlock;                 // 128-byte K blocks per row
    static constexpr int N = D <= 64 ? DUNK_L2_N64 : 128;  // train rows per MMA tile
    static constexpr int A_BYTES = KB * kM * 128;
    static constexpr int B_BYTES = KB * N * 128;
    static constexpr int TMEM_COLS = kAcc * N;             // 512 or 256 (power of two)
    static constexpr int kStages = D <= 64 ? (N == 256 ? 2 : 5) : 2;   // smem ring for the train tiles (227 KB per CTA)
    // ||t||^2 of a tile travels with it (one 1-D bulk copy on the tile's barrier) into a ring of kStages + kAcc slots:
    // the producer is at most kStages tiles ahead of the MMA warp, which is at most kAcc tiles ahead of the epilogue
    static constexpr int kTn = kStages + kAcc;
    static constexpr int TN_BYTES = N * 4;
    static constexpr int SMEM = A_BYTES + kStages * B_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + kTn * TN_BYTES;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded spin: a protocol bug traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both K-major, tf32 inputs, f32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major operand, 128-byte swizzle,
// rows of 128 bytes, 8-row groups 1024 bytes apart; version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);            // start address        bits [0,14)
    d |= (uint64_t)1 << 16;                                  // leading byte offset  bits [16,30) (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                                  // descriptor version   bits [46,48)
    d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B         bits [61,64)
    return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32: f32 accumulate, A and B TF32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct Top4 {
    float s[kCand];
    uint32_t i[kCand];
    __device__ void init() {
#pragma unroll
        for (int k = 0; k < kCand; ++k) { s[k] = INFINITY; i[k] = kNoIdx; }
    }
    // rows arrive in increasing index order: strict '<' keeps the lower index on ties
    __device__ __forceinline__ void insert(float v, uint32_t idx) {
        if (v < s[3]) {
            s[3] = v; i[3] = idx;
#pragma unroll
            for (int k = 3; k > 0; --k)
                if (s[k] < s[k - 1]) {
                    const float ts = s[k]; s[k] = s[k - 1]; s[k - 1] = ts;
                    const uint32_t ti = i[k]; i[k] = i[k - 1]; i[k - 1] = ti;
                }
        }
    }
};

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
l2_candidates_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t,
                     const float* __restrict__ tnorm /* padded to whole tiles, +inf past nt */, int nq, int total_tiles,
                     int tiles_per_slab, const float* __restrict__ tau /* per-query admission threshold or NULL */,
                     float4* __restrict__ cand_score, uint4* __restrict__ cand_idx, int dbg_skip) {
    using C = Cfg<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // 128B swizzle atoms need 1024-byte alignment
    uint8_t* sA = smem;
    uint8_t* sB = smem + C::A_BYTES;
    constexpr int kStages = C::kStages;
    uint64_t* bars = (uint64_t*)(smem + C::A_BYTES + kStages * C::B_BYTES);
    uint64_t* full_a = bars;                    // 1
    uint64_t* full_b = bars + 1;                // kStages
    uint64_t* empty_b = full_b + kStages;       // kStages
    uint64_t* acc_full = empty_b + kStages;     // kAcc
    uint64_t* acc_empty = acc_full + kAcc;      // kAcc
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + kAcc);
    float* sTn = (float*)(smem + C::A_BYTES + kStages * C::B_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // D = 128 (HBM-heavier): query tile fastest, so the CTAs that share a train slab are launched next to each other,
    // stream it in step and the slab is read from DRAM once (r1 ncu: 2x re-read with the slab index fastest; measured
    // 595 -> 631 TFLOP/s).  D = 64 is epilogue-bound and measured 7 % faster with the slab index fastest.
    const int slab = D >= 128 ? blockIdx.y : blockIdx.x, m0 = (D >= 128 ? blockIdx.x : blockIdx.y) * kM;
    const int tile0 = slab * tiles_per_slab;
    const int ntiles = min(tiles_per_slab, total_tiles - tile0);

    if (threadIdx.x == 0) {
        mbar_init(full_a, 1);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int s = 0; s < kAcc; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation is a warp-wide instruction; this warp also frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(full_a, C::A_BYTES);
            for (int kb = 0; kb < C::KB; ++kb) tma_load_2d(sA + kb * kM * 128, &map_q, kb * kKBlock, m0, full_a);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % kStages;
                mbar_wait(&empty_b[s], ((j / kStages) & 1) ^ 1);
                mbar_expect_tx(&full_b[s], C::B_BYTES + C::TN_BYTES);
                for (int kb = 0; kb < C::KB; ++kb)
                    tma_load_2d(sB + s * C::B_BYTES + kb * C::N * 128, &map_t, kb * kKBlock, (tile0 + j) * C::N, &full_b[s]);
                bulk_load_1d(sTn + (j % C::kTn) * C::N, tnorm + (size_t)(tile0 + j) * C::N, C::TN_BYTES, &full_b[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kM, C::N);
            mbar_wait(full_a, 0);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j % kStages, a = j % kAcc;
                mbar_wait(&acc_empty[a], ((j / kAcc) & 1) ^ 1);
                mbar_wait(&full_b[s], (j / kStages) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * C::N);
#pragma unroll
                for (int kb = 0; kb < C::KB; ++kb)
#pragma unroll
                    for (int k = 0; k < kKBlock / kUmmaK; ++k) {
                        const uint64_t da = umma_desc_sw128(smem_u32(sA + kb * kM * 128) + k * kUmmaK * 4);
                        const uint64_t db = umma_desc_sw128(smem_u32(sB + s * C::B_BYTES + kb * C::N * 128) + k * kUmmaK * 4);
                        umma_tf32(d_tmem, da, db, idesc, (kb | k) != 0);
                    }
                tc_commit(&empty_b[s]);     // the smem stage may be refilled once these MMAs have read it
                tc_commit(&acc_full[a]);    // ... and the accumulator is complete
            }
        }
    } else {
        // ===== epilogue warps; TMEM lane quarter = warp % 4 (hardware rule), column group = (warp-2)/4;
        // one query per thread, 4 smallest scores of its half kept in registers =====
        const int quarter = warp & 3, half = (warp - 2) >> 2;   // `half` = column group 0..kSub-1
        const int row = m0 + quarter * 32 + lane;
        constexpr int kCols = C::N / kSub, kChunks = kCols / 32;
        Top4 top;
        top.init();
        if (tau && row < nq) {
            // admission threshold from the seed pass: the list starts with 4 sentinels at tau, so only rows that beat
            // tau enter, and s[3] stays a valid lower bound for every row this list drops
            const float t0 = tau[row];
#pragma unroll
            for (int k = 0; k < kCand; ++k) top.s[k] = t0;
        }
        // 16 accumulator columns of this thread's TMEM lane -> registers (asynchronous until wait_ld)
        auto tmem_ld16 = [&](uint32_t taddr, uint32_t (&v)[16]) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
        };
        // tcgen05.wait::ld with the loaded registers as in/out operands: every later use of v depends on the wait
        auto wait_ld = [&](uint32_t (&v)[16]) {
            asm volatile("tcgen05.wait::ld.sync.aligned;"
                         : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                           "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                         :
                         : "memory");
        };
        // scores of one 16-column chunk, ||t||^2 - 2 q.t (||t||^2 from the tile's shared-memory slot: the same address for
        // all lanes), reduced on the fly to g[k] = min over the 4 columns c with ((c >> 1) & 3) == k; pure math, no warp sync
        auto score = [&](const uint32_t (&v)[16], const float* tn, float (&g)[4], float& all) {
            const float4* tnp = reinterpret_cast<const float4*>(tn);
            float m[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 x = tnp[k];
                m[2 * k] = fminf(fmaf(-2.f, __uint_as_float(v[4 * k]), x.x), fmaf(-2.f, __uint_as_float(v[4 * k + 1]), x.y));
                m[2 * k + 1] = fminf(fmaf(-2.f, __uint_as_float(v[4 * k + 2]), x.z), fmaf(-2.f, __uint_as_float(v[4 * k + 3]), x.w));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) g[k] = fminf(m[k], m[k + 4]);
            all = fminf(fminf(g[0], g[1]), fminf(g[2], g[3]));
        };
        // the common case (no lane of the warp improves its list) is a single vote; otherwise narrow down to
        // the column groups, then the columns, that improve some lane (32 lanes share every branch); the few scores
        // that are looked at are formed again from the accumulator registers
        auto maybe_insert = [&](const uint32_t (&v)[16], const float* tn, const float (&g)[4], float all, uint32_t t_first) {
            if (__any_sync(0xffffffffu, all < top.s[3])) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (__any_sync(0xffffffffu, g[k] < top.s[3])) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int c = 2 * k + (u & 1) + 8 * (u >> 1);
                            const float sc = fmaf(-2.f, __uint_as_float(v[c]), tn[c]);
                            if (__any_sync(0xffffffffu, sc < top.s[3])) top.insert(sc, t_first + c);
                        }
                    }
            }
        };
        // The warp's columns of a tile are walked in chunks of 16; chunk q = tile q / kPer, part q % kPer.  While the
        // scores of chunk q are reduced, the TMEM load of chunk q + 1 is in flight: one warp sees every tile, so its own
        // chain (barrier, TMEM load, math) bounds the CTA and the accumulator ring only decouples it from the MMA warp.
        constexpr int kPer = kCols / 16;
        const uint32_t taddr_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * kCols);
        auto issue_ld = [&](int q, uint32_t (&v)[16]) {
            const int j = q / kPer, part = q % kPer, a = j % kAcc;
            if (part == 0) {
                mbar_wait(&acc_full[a], (j / kAcc) & 1);
                tc_fence_after();
            }
            tmem_ld16(taddr_base + (uint32_t)(a * C::N + 16 * part), v);
        };
        // chunk q has landed in v (tcgen05.wait::ld waits for EVERY outstanding load of the thread, so it is called while
        // only this chunk's load is in flight, before the next one is issued)
        auto landed = [&](int q, uint32_t (&v)[16]) {
            wait_ld(v);
            const int j = q / kPer, part = q % kPer;
            if (part == kPer - 1) {     // every column of tile j is in registers: the accumulator stage goes back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[j % kAcc]);
            }
        };
        auto reduce = [&](int q, const uint32_t (&v)[16]) {
            const int j = q / kPer, part = q % kPer;
            const float* tn = sTn + (j % C::kTn) * C::N + half * kCols + 16 * part;
            float g[4], all;
            score(v, tn, g, all);
            maybe_insert(v, tn, g, all, (uint32_t)(tile0 + j) * C::N + half * kCols + 16 * part);
        };
        {
            uint32_t va[16], vb[16];
            const int nchunks = ntiles * kPer;
            if (nchunks > 0) issue_ld(0, va);
            for (int q = 0; q < nchunks; q += 2) {
                landed(q, va);
                if (q + 1 < nchunks) issue_ld(q + 1, vb);
                if (!dbg_skip) reduce(q, va);
                if (q + 1 < nchunks) {
                    landed(q + 1, vb);
                    if (q + 2 < nchunks) issue_ld(q + 2, va);
                    if (!dbg_skip) reduce(q + 1, vb);
                }
            }
        }
        if (row < nq) {
            const size_t o = ((size_t)slab * kSub + half) * nq + row;
            cand_score[o] = make_float4(top.s[0], top.s[1], top.s[2], top.s[3]);
            cand_idx[o] = make_uint4(top.i[0], top.i[1], top.i[2], top.i[3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

// seed pass -> per-query admission threshold: the 4th smallest tf32 score over the seed lists (a real row's
// score, hence an upper bound of the 4th best score over the whole train set)
__global__ void __launch_bounds__(256)
l2_tau_kernel(const float4* __restrict__ cand_score, int n_lists, int nq, float* __restrict__ tau) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    float s[kCand] = {INFINITY, INFINITY, INFINITY, INFINITY};
    for (int l = 0; l < n_lists; ++l) {
        const float4 c = cand_score[(size_t)l * nq + qi];
        const float v[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (v[k] < s[3]) {
                s[3] = v[k];
#pragma unroll
                for (int u = 3; u > 0; --u)
                    if (s[u] < s[u - 1]) { const float t = s[u]; s[u] = s[u - 1]; s[u - 1] = t; }
            }
    }
    tau[qi] = s[3];
}

// squared norms of n rows of `dim` floats; rows in [n, n_padded) get +inf (they are TMA zero-fill rows)
__global__ void __launch_bounds__(256)
row_norms_kernel(const float* __restrict__ x, long long n, long long n_padded, int dim, float* __restrict__ out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_padded) return;
    if (r >= n) { out[r] = INFINITY; return; }
    const float4* p = reinterpret_cast<const float4*>(x + r * dim);
    float s = 0.f;
    for (int k = 0; k < dim / 4; ++k) {
        const float4 v = p[k];
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    out[r] = s;
}

// exact squared L2 distance in f32 (sequential, unfused: the order a scalar normL2Sqr loop uses)
__device__ __forceinline__ float exact_d2(const float* __restrict__ q, const float* __restrict__ t, int dim) {
    float s = 0.f;
    for (int k = 0; k < dim; k += 4) {
        const float4 a = *reinterpret_cast<const float4*>(q + k);
        const float4 b = __ldg(reinterpret_cast<const float4*>(t + k));
        float d = __fsub_rn(a.x, b.x); s = __fadd_rn(s, __fmul_rn(d, d));
        d = __fsub_rn(a.y, b.y); s = __fadd_rn(s, __fmul_rn(d, d));
        d = __fsub_rn(a.z, b.z); s = __fadd_rn(s, __fmul_rn(d, d));
        d = __fsub_rn(a.w, b.w); s = __fadd_rn(s, __fmul_rn(d, d));
    }
    return s;
}

__device__ __forceinline__ bool lex_less(float da, uint32_t ia, float db, uint32_t ib) { return da < db || (da == db && ia < ib); }

// warp-wide two smallest (d, i) pairs; every lane contributes one pair (d = inf, i = kNoIdx if none)
__device__ __forceinline__ void warp_top2(float d, uint32_t i, float& d1, uint32_t& i1, float& d2, uint32_t& i2) {
    float bd = d; uint32_t bi = i;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (lex_less(od, oi, bd, bi)) { bd = od; bi = oi; }
    }
    d1 = bd; i1 = bi;
    if (i == bi && d == bd) { d = INFINITY; i = kNoIdx; }     // the winner drops out
    bd = d; bi = i;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (lex_less(od, oi, bd, bi)) { bd = od; bi = oi; }
    }
    d2 = bd; i2 = bi;
}

// stage B: one warp per query
__global__ void __launch_bounds__(256)
l2_rerank_kernel(const float* __restrict__ q, const float* __restrict__ t, int nq, uint32_t nt, int dim, int n_slabs,
                 const float4* __restrict__ cand_score, const uint4* __restrict__ cand_idx, const float* __restrict__ qnorm,
                 float tnorm_max, int32_t* __restrict__ idx_out, float* __restrict__ dist_out, int* __restrict__ n_flagged,
                 int* __restrict__ flagged) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const float* qp = q + (size_t)qi * dim;
    float bd1 = INFINITY, bd2 = INFINITY;
    uint32_t bi1 = kNoIdx, bi2 = kNoIdx;
    float dropped_min = INFINITY;      // smallest tf32 score a dropped row can have
    const int n_c = n_slabs * kCand;
    for (int base = 0; base < n_c; base += 32) {
        const int c = base + lane;
        float d = INFINITY;
        uint32_t i = kNoIdx;
        if (c < n_c) {
            const int slab = c / kCand, k = c % kCand;
            const uint4 ci = cand_idx[(size_t)slab * nq + qi];
            const float4 cs = cand_score[(size_t)slab * nq + qi];
            i = k == 0 ? ci.x : k == 1 ? ci.y : k == 2 ? ci.z : ci.w;
            if (k == kCand - 1) dropped_min = fminf(dropped_min, cs.w);   // a slab with < 4 rows leaves +inf here
            if (i != kNoIdx && i < nt) d = exact_d2(qp, t + (size_t)i * dim, dim);
            else i = kNoIdx;
        }
        float d1, d2;
        uint32_t i1, i2;
        warp_top2(d, i, d1, i1, d2, i2);
        // merge the chunk's pair into the running pair
        if (lex_less(d1, i1, bd1, bi1)) {
            if (lex_less(bd1, bi1, d2, i2)) { bd2 = bd1; bi2 = bi1; } else { bd2 = d2; bi2 = i2; }
            bd1 = d1; bi1 = i1;
        } else if (lex_less(d1, i1, bd2, bi2)) {
            bd2 = d1; bi2 = i1;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) dropped_min = fminf(dropped_min, __shfl_xor_sync(0xffffffffu, dropped_min, o));
    if (lane == 0) {
        idx_out[2 * qi] = (int32_t)bi1;
        idx_out[2 * qi + 1] = (int32_t)bi2;
        dist_out[2 * qi] = sqrtf(bd1);
        dist_out[2 * qi + 1] = sqrtf(bd2);
        // |tf32 score - exact score| <= 2 * (2 * 2^-10 + 2^-20) * |q|.|t| + accumulation noise < 2^-8 |q| |t|_max
        const float qn = qnorm[qi];
        const float margin = 0.00390625f * sqrtf(qn) * sqrtf(tnorm_max);
        const float second_score = bd2 - qn;                  // exact ||t||^2 - 2 q.t of the 2nd best (up to rounding)
        const bool proven = dropped_min == INFINITY || second_score < dropped_min - margin;
        if (!proven) flagged[atomicAdd(n_flagged, 1)] = qi;
    }
}

// exact brute force for the queries stage B could not prove: grid (flagged query, part of the train rows); every CTA
// leaves its two lexicographically smallest (distance, index) pairs in part_out, l2_exact_merge_kernel merges the parts
// (one CTA per query over all 2 M rows took 22 ms for a single flagged query)
__global__ void __launch_bounds__(256)
l2_exact_kernel(const float* __restrict__ q, const float* __restrict__ t, uint32_t nt, int dim, const int* __restrict__ flagged,
                float4* __restrict__ part_out /* [n_flagged][gridDim.y]: d1, i1 bits, d2, i2 bits */) {
    const int qi = flagged[blockIdx.x];
    const float* qp = q + (size_t)qi * dim;
    const uint32_t per = (nt + gridDim.y - 1) / gridDim.y;
    const uint32_t r0 = blockIdx.y * per, r1 = min(nt, r0 + per);
    float d1 = INFINITY, d2 = INFINITY;
    uint32_t i1 = kNoIdx, i2 = kNoIdx;
    for (uint32_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        const float d = exact_d2(qp, t + (size_t)r * dim, dim);
        if (lex_less(d, r, d2, i2)) {
            if (lex_less(d, r, d1, i1)) { d2 = d1; i2 = i1; d1 = d; i1 = r; } else { d2 = d; i2 = r; }
        }
    }
    // rare path: publish every thread's pair, one thread scans the 512 entries
    __shared__ float all_d[512];
    __shared__ uint32_t all_i[512];
    all_d[2 * threadIdx.x] = d1; all_i[2 * threadIdx.x] = i1;
    all_d[2 * threadIdx.x + 1] = d2; all_i[2 * threadIdx.x + 1] = i2;
    __syncthreads();
    if (threadIdx.x == 0) {
        float b1 = INFINITY, b2 = INFINITY;
        uint32_t c1 = kNoIdx, c2 = kNoIdx;
        for (int k = 0; k < 512; ++k) {
            const float d = all_d[k];
            const uint32_t i = all_i[k];
            if (lex_less(d, i, b2, c2)) {
                if (lex_less(d, i, b1, c1)) { b2 = b1; c2 = c1; b1 = d; c1 = i; } else { b2 = d; c2 = i; }
            }
        }
        part_out[(size_t)blockIdx.x * gridDim.y + blockIdx.y] = make_float4(b1, __uint_as_float(c1), b2, __uint_as_float(c2));
    }
}

__global__ void l2_exact_merge_kernel(const float4* __restrict__ part, int n_parts, const int* __restrict__ flagged, int n_flagged,
                                      int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_flagged) return;
    float b1 = INFINITY, b2 = INFINITY;
    uint32_t c1 = kNoIdx, c2 = kNoIdx;
    for (int p = 0; p < n_parts; ++p) {
        const float4 v = part[(size_t)k * n_parts + p];
        const float d[2] = {v.x, v.z};
        const uint32_t i[2] = {__float_as_uint(v.y), __float_as_uint(v.w)};
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (lex_less(d[u], i[u], b2, c2)) {
                if (lex_less(d[u], i[u], b1, c1)) { b2 = b1; c2 = c1; b1 = d[u]; c1 = i[u]; } else { b2 = d[u]; c2 = i[u]; }
            }
    }
    const int qi = flagged[k];
    idx_out[2 * qi] = (int32_t)c1; idx_out[2 * qi + 1] = (int32_t)c2;
    dist_out[2 * qi] = sqrtf(b1); dist_out[2 * qi + 1] = sqrtf(b2);
}

__global__ void max_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, x[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax((int*)out, __float_as_int(m));   // non-negative floats order like ints
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x dim f32, row-major; box = 32 floats (one 128-byte swizzle row) x box_rows
bool make_map(CUtensorMap* m, const float* base, uint64_t rows, int dim, int box_rows) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)dim, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)dim * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kKBlock, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// per-device opt-in to the kernels' dynamic shared memory (run by dunk_ctx_create on the context's device)
static int l2_device_init(dunk_ctx*) {
    DUNK_CUDA(cudaFuncSetAttribute(l2_candidates_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM));
    DUNK_CUDA(cudaFuncSetAttribute(l2_candidates_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM));
    return DUNK_OK;
}
static DeviceInitReg l2_device_init_reg(l2_device_init);

template <int D>
int launch_candidates(dunk_ctx* ctx, cudaStream_t st, const float* d_q, int nq, const float* d_t, uint32_t nt, const float* d_tnorm,
                      int total_tiles, int tiles_per_slab, int n_slabs, const float* d_tau, float4* cand_score, uint4* cand_idx, const char* label) {
    using C = Cfg<D>;
    alignas(64) CUtensorMap mq, mt;
    if (!make_map(&mq, d_q, (uint64_t)nq, D, kM) || !make_map(&mt, d_t, nt, D, C::N)) {
        set_error("dunk_knn2_l2: cuTensorMapEncodeTiled failed");
        return DUNK_ERR_CUDA;
    }
    ProfScope ps(ctx, st, label, (double)nq * (double)std::min<long long>((long long)total_tiles * C::N, (long long)nt));
    const dim3 grid = D >= 128 ? dim3(div_up(nq, kM), n_slabs) : dim3(n_slabs, div_up(nq, kM));
    static const int dbg_skip = getenv("DUNK_L2_SKIP_EPILOGUE") ? atoi(getenv("DUNK_L2_SKIP_EPILOGUE")) : 0;
    l2_candidates_kernel<D><<<grid, kThreads, C::SMEM, st>>>(mq, mt, d_tnorm, nq, total_tiles, tiles_per_slab,
                                                                                    d_tau, cand_score, cand_idx, dbg_skip);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    return DUNK_OK;
}

}  // namespace
}  // namespace dunk

using namespace dunk;

extern "C" {

/* device-resident inputs: q [nq][dim], t [nt][dim] f32 (16-byte aligned); outputs idx [nq][2] i32, dist [nq][2] f32 (device).
 * stats (host, may be NULL): {queries re-done by the exact fallback, slabs} */
int dunk_knn2_l2_dev(dunk_ctx* ctx, int slot, const void* q_dev, int nq, const void* t_dev, int64_t nt, int dim, void* idx_dev,
                     void* dist_dev, int* stats) {
    DUNK_REQUIRE(ctx && slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_knn2_l2_dev: bad ctx / slot");
    DUNK_REQUIRE(dim == 64 || dim == 128, DUNK_ERR_BAD_ARG, "dunk_knn2_l2: descriptors of %d floats unsupported (64 or 128)", dim);
    DUNK_REQUIRE(nq >= 0 && nt >= 0 && nt < 0xFFFFFFFFll, DUNK_ERR_BAD_ARG, "dunk_knn2_l2: bad row counts");
    if (stats) stats[0] = stats[1] = 0;
    if (nq == 0) return DUNK_OK;
    // knnMatch with fewer than 2 train rows yields short lists; the reference's wrapper indexes [1] (lib.rs:108)
    DUNK_REQUIRE(nt >= 2, DUNK_ERR_OUT_OF_RANGE, "dunk_knn2_l2: %lld train rows, 2 are needed", (long long)nt);
    DUNK_REQUIRE(q_dev && t_dev && idx_dev && dist_dev, DUNK_ERR_BAD_ARG, "dunk_knn2_l2_dev: NULL pointer");
    DUNK_REQUIRE(((uintptr_t)q_dev & 15) == 0 && ((uintptr_t)t_dev & 15) == 0, DUNK_ERR_BAD_ARG, "dunk_knn2_l2_dev: descriptors must be 16-byte aligned");
    DUNK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[slot].stream;
    const int N = dim <= 64 ? Cfg<64>::N : Cfg<128>::N;
    const int total_tiles = div_up(nt, N);
    const int q_tiles = div_up(nq, kM);
    // two full waves of CTAs (one CTA per SM: 160-197 KB of smem); every query tile re-reads its slab through L2
    int n_slabs = std::max(1, std::min(total_tiles, (2 * ctx->sm_count) / q_tiles));
    const int tiles_per_slab = div_up(total_tiles, n_slabs);
    n_slabs = div_up(total_tiles, tiles_per_slab);
    const long long padded = (long long)total_tiles * N;
    size_t need = Carver::need((size_t)padded * 4) + Carver::need((size_t)nq * 4) + 2 * Carver::need((size_t)n_slabs * kSub * nq * 16) +
                  Carver::need(16) + 2 * Carver::need((size_t)nq * 4);
    void* scratch = ctx->dev_scratch(slot, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    float* d_tnorm = cv.take<float>(padded);
    float* d_qnorm = cv.take<float>(nq);
    float4* d_cs = cv.take<float4>((size_t)n_slabs * kSub * nq);
    uint4* d_ci = cv.take<uint4>((size_t)n_slabs * kSub * nq);
    int* d_misc = cv.take<int>(4);      // [0] flagged count, [1] max ||t||^2 (float bits)
    int* d_flagged = cv.take<int>(nq);
    float* d_tau = cv.take<float>(nq);
    DUNK_CUDA(cudaMemsetAsync(d_misc, 0, 16, st));
    row_norms_kernel<<<div_up(padded, 256), 256, 0, st>>>((const float*)t_dev, nt, padded, dim, d_tnorm);
    row_norms_kernel<<<div_up(nq, 256), 256, 0, st>>>((const float*)q_dev, nq, nq, dim, d_qnorm);
    max_kernel<<<std::min(1024, div_up(nt, 256)), 256, 0, st>>>(d_tnorm, nt, (float*)(d_misc + 1));
    ctx->launches.fetch_add(3);
    DUNK_CUDA(cudaGetLastError());
    // seed pass over the first 1/16 of the train rows -> per-query admission threshold for the main pass
    // (without it every list converges harmonically and 32 lanes share each insertion branch)
    const float* d_tau_arg = nullptr;
    int rc = DUNK_OK;
    const int seed_tiles = total_tiles / 16;
    if (seed_tiles >= 1) {
        int seed_slabs = std::max(1, std::min(seed_tiles, std::min(n_slabs, ctx->sm_count / q_tiles)));
        const int seed_tps = div_up(seed_tiles, seed_slabs);
        seed_slabs = div_up(seed_tiles, seed_tps);
        rc = dim == 64 ? launch_candidates<64>(ctx, st, (const float*)q_dev, nq, (const float*)t_dev, (uint32_t)nt, d_tnorm, seed_tiles, seed_tps,
                                               seed_slabs, nullptr, d_cs, d_ci, "match.l2_tcgen05_seed")
                       : launch_candidates<128>(ctx, st, (const float*)q_dev, nq, (const float*)t_dev, (uint32_t)nt, d_tnorm, seed_tiles, seed_tps,
                                                seed_slabs, nullptr, d_cs, d_ci, "match.l2_tcgen05_seed");
        if (rc) return rc;
        l2_tau_kernel<<<div_up(nq, 256), 256, 0, st>>>(d_cs, seed_slabs * kSub, nq, d_tau);
        ctx->launches.fetch_add(1);
        DUNK_CUDA(cudaGetLastError());
        d_tau_arg = d_tau;
    }
    rc = dim == 64 ? launch_candidates<64>(ctx, st, (const float*)q_dev, nq, (const float*)t_dev, (uint32_t)nt, d_tnorm, total_tiles,
                                           tiles_per_slab, n_slabs, d_tau_arg, d_cs, d_ci, "match.l2_tcgen05")
                   : launch_candidates<128>(ctx, st, (const float*)q_dev, nq, (const float*)t_dev, (uint32_t)nt, d_tnorm, total_tiles,
                                            tiles_per_slab, n_slabs, d_tau_arg, d_cs, d_ci, "match.l2_tcgen05");
    if (rc) return rc;
    int h_misc[4] = {0, 0, 0, 0};
    DUNK_CUDA(cudaMemcpyAsync(h_misc, d_misc, 16, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));        // max ||t||^2 is a kernel argument of stage B
    float tmax;
    memcpy(&tmax, &h_misc[1], 4);
    {
        ProfScope ps(ctx, st, "match.l2_rerank", (double)nq * n_slabs * kSub * kCand);
        l2_rerank_kernel<<<div_up(nq, 8), 256, 0, st>>>((const float*)q_dev, (const float*)t_dev, nq, (uint32_t)nt, dim, n_slabs * kSub, d_cs, d_ci,
                                                      d_qnorm, tmax, (int32_t*)idx_dev, (float*)dist_dev, d_misc, d_flagged);
        ctx->launches.fetch_add(1);
        DUNK_CUDA(cudaGetLastError());
    }
    DUNK_CUDA(cudaMemcpyAsync(h_misc, d_misc, 4, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    if (h_misc[0] > 0) {
        ProfScope ps(ctx, st, "match.l2_exact_fallback", (double)h_misc[0] * (double)nt);
        // the candidate-score array is free again: it holds the per-part pairs (at most nq * n_slabs * kSub entries)
        const int parts = (int)std::max<long long>(1, std::min<long long>({64ll, (long long)n_slabs * kSub, ((long long)nt + 4095) / 4096}));
        l2_exact_kernel<<<dim3(h_misc[0], parts), 256, 0, st>>>((const float*)q_dev, (const float*)t_dev, (uint32_t)nt, dim, d_flagged, d_cs);
        l2_exact_merge_kernel<<<div_up(h_misc[0], 128), 128, 0, st>>>(d_cs, parts, d_flagged, h_misc[0], (int32_t*)idx_dev, (float*)dist_dev);
        ctx->launches.fetch_add(2);
        DUNK_CUDA(cudaGetLastError());
    }
    if (stats) { stats[0] = h_misc[0]; stats[1] = n_slabs; }
    return DUNK_OK;
}

int dunk_knn2_l2(dunk_ctx* ctx, const float* query, int nq, const float* train, int64_t nt, int dim, int32_t* idx, float* dist,
                 int* stats) {
    DUNK_REQUIRE(ctx, DUNK_ERR_BAD_ARG, "dunk_knn2_l2: ctx is NULL");
    DUNK_REQUIRE(dim == 64 || dim == 128, DUNK_ERR_BAD_ARG, "dunk_knn2_l2: descriptors of %d floats unsupported (64 or 128)", dim);
    DUNK_REQUIRE(nq >= 0 && nt >= 0, DUNK_ERR_BAD_ARG, "dunk_knn2_l2: negative row count");
    if (stats) stats[0] = stats[1] = 0;
    if (nq == 0) return DUNK_OK;
    DUNK_REQUIRE(nt >= 2, DUNK_ERR_OUT_OF_RANGE, "dunk_knn2_l2: %lld train rows, 2 are needed", (long long)nt);
    DUNK_REQUIRE(query && train && idx && dist, DUNK_ERR_BAD_ARG, "dunk_knn2_l2: NULL pointer");
    // private buffers + a reserved slot: the _dev entry point uses the slot's grow-only scratch itself
    const int slot = dunk_ctx_reserve_slot(ctx);
    if (slot < 0) return slot;
    float *d_q = nullptr, *d_t = nullptr, *d_dist = nullptr;
    int32_t* d_idx = nullptr;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->slots[slot].stream;
    int rc = DUNK_OK;
    if (cudaMalloc(&d_q, (size_t)nq * dim * 4) != cudaSuccess || cudaMalloc(&d_t, (size_t)nt * dim * 4) != cudaSuccess ||
        cudaMalloc(&d_idx, (size_t)nq * 8) != cudaSuccess || cudaMalloc(&d_dist, (size_t)nq * 8) != cudaSuccess) {
        cudaGetLastError();
        set_error("dunk_knn2_l2: device allocation failed");
        rc = DUNK_ERR_NO_MEM;
    }
    if (!rc && (cudaMemcpyAsync(d_q, query, (size_t)nq * dim * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                cudaMemcpyAsync(d_t, train, (size_t)nt * dim * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)) {
        set_error("dunk_knn2_l2: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = DUNK_ERR_CUDA;
    }
    if (!rc) rc = dunk_knn2_l2_dev(ctx, slot, d_q, nq, d_t, nt, dim, d_idx, d_dist, stats);
    if (!rc && (cudaMemcpyAsync(idx, d_idx, (size_t)nq * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaMemcpyAsync(dist, d_dist, (size_t)nq * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess)) {
        set_error("dunk_knn2_l2: kernel or download failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = DUNK_ERR_CUDA;
    }
    cudaStreamSynchronize(st);
    cudaFree(d_q); cudaFree(d_t); cudaFree(d_idx); cudaFree(d_dist);
    dunk_ctx_release_slot(ctx, slot);
    return rc;
}

}  // extern "C"
