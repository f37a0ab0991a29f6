// f32_to_u8's tail (geotiff_extractor/src/image_extractor/mod.rs:346-378, 402-422):
//     u8 = floor((float)pow((double)fl, 1 / 2.2f) * 255 + 0.5),  fl in [0, 1]
// is monotone in fl, so it equals the number of thresholds T[k] (k = 1..255, T[k] = the smallest f32 whose
// value is >= k) that are <= fl.  The table is built on the device from the direct formula by bisection
// over the f32 bit patterns, and dunk_selftest_gamma_lut() compares table and formula on EVERY f32 in
// [0, 1]; the kernels then replace three f64 pow calls per pixel by 8 shared-memory compares per channel.
#pragma once
#include <cuda_runtime.h>

struct dunk_ctx;

namespace dunk {

__device__ __forceinline__ unsigned char gamma_u8_direct(float fl) {
    // f32::powf is correctly rounded in practice (glibc evaluates it in double): do the same
    const float g = (float)pow((double)fl, (double)(1.0f / 2.2f));
    const float x = __fmul_rn(g, 255.f);
    return (unsigned char)(int)floorf(__fadd_rn(x, 0.5f));   // round half away from zero, x >= 0
}

// s_thr: 256 floats in shared memory (s_thr[0] unused)
__device__ __forceinline__ unsigned char gamma_u8_lut(const float* s_thr, float fl) {
    int r = 0;
#pragma unroll
    for (int step = 128; step; step >>= 1)
        if (fl >= s_thr[r + step]) r += step;
    return (unsigned char)r;
}

// f32_to_u8(...).unwrap_or(0) with the table
__device__ __forceinline__ unsigned char f32_to_u8_lut(const float* s_thr, float v, float vmin, float vmax) {
    if (isnan(v)) return 0;
    const float fl = __fdiv_rn(__fsub_rn(v, vmin), __fsub_rn(vmax, vmin));
    if (!(fl >= 0.f && fl <= 1.f)) return 0;                 // gamma_correction: GammaOutOfRange
    return gamma_u8_lut(s_thr, fl);
}

// device pointer to the 256-entry table of the context's device, built on first use (stream-ordered on st)
const float* gamma_table(dunk_ctx* ctx, cudaStream_t st);

}  // namespace dunk
