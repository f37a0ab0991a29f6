// Internal: cross-stage launchers used by the end-to-end registration pipeline.
#pragma once
#include "ctx.h"

namespace dunk {
int launch_find_homography(dunk_ctx* ctx, cudaStream_t st, const float2* src, const float2* dst, const int* starts,
                           const int* counts, int n_problems, float thr, double* H, uint8_t* mask, int* info);
int launch_pnp_ransac(dunk_ctx* ctx, cudaStream_t st, const float* obj, const float* img, const int* starts, const int* counts,
                      int n_problems, const double* K, int k_stride, int method, int iters, float thr, double confidence,
                      double* rt, uint8_t* mask, int* info);
}
