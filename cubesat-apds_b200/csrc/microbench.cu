// Pipe-peak microbenchmarks used as roofline denominators (SURVEY 8d asks the builder to
// confirm the POPC rate on the box, the way MEASURED_PEAKS.json did for HBM).
#include "ctx.h"

namespace {

// 8 independent XOR->POPC->ADD chains per thread; the adds go to the ALU pipe, POPC to its own
__global__ void __launch_bounds__(256) popc_peak_kernel(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __popc(a[i] ^ seed) + (a[i] << 3);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345678u) out[0] = s;  // keep the chains alive
}

}  // namespace

extern "C" int dunk_microbench_popc(dunk_ctx* ctx, int iters, double* tpopc_per_s) {
    DUNK_REQUIRE(ctx && tpopc_per_s && iters > 0, DUNK_ERR_BAD_ARG, "dunk_microbench_popc: bad argument");
    dunk::SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    uint32_t* out = (uint32_t*)ctx->dev_scratch(g.s, 256);
    if (!out) return DUNK_ERR_NO_MEM;
    const int blocks = ctx->sm_count * 8;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        DUNK_CUDA(cudaEventRecord(g.slot().ev0, st));
        popc_peak_kernel<<<blocks, 256, 0, st>>>(out, 0x5bd1e995u + rep, iters);
        ctx->launches.fetch_add(1);
        DUNK_CUDA(cudaEventRecord(g.slot().ev1, st));
        DUNK_CUDA(cudaEventSynchronize(g.slot().ev1));
        float ms = 0;
        DUNK_CUDA(cudaEventElapsedTime(&ms, g.slot().ev0, g.slot().ev1));
        const double ops = (double)blocks * 256.0 * iters * 32.0;
        if (rep > 0 && ms > 0) best = std::max(best, ops / (ms * 1e-3) / 1e12);
    }
    *tpopc_per_s = best;
    return DUNK_OK;
}
