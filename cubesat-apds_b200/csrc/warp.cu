// warp_image_perspective (homographier/src/homographier/mod.rs:271-300): cv::warpPerspective(src, dst,
// M, size, INTER_LINEAR, BORDER_CONSTANT, Scalar(1,1,1,1)) for 8-bit images, bit-exact with OpenCV's
// WarpPerspectiveInvoker + remapBilinear: M is inverted on the host exactly as
// cv::invert does, source coordinates are evaluated in f64 per 64 x 16 block the way OpenCV does
// (X0 at the block's first column + M0 * x1, no FMA contraction), snapped to the 1/32-pixel grid with
// round-half-even, and blended with the 15-bit fixed-point bilinear table.
// One thread per output pixel; a batch (blockIdx.z) warps the same source with one matrix per output,
// which is how config 1/5 query frames are cut from a scene without leaving the device.
#include <cmath>
#include "ctx.h"

namespace dunk {
namespace {

struct WarpMats {
    double m[9];
};

__device__ __forceinline__ int fetch(const unsigned char* __restrict__ src, int row_stride, int channels, int sw, int sh, int x,
                                     int y, int c, int border) {
    return ((unsigned)x < (unsigned)sw && (unsigned)y < (unsigned)sh) ? src[(size_t)y * row_stride + x * channels + c] : border;
}

__global__ void __launch_bounds__(256)
k_warp_perspective(const unsigned char* __restrict__ src, int sw, int sh, int channels, int row_stride,
                   const double* __restrict__ mats /* [batch][9], already inverted */, int dw, int dh, int bw,
                   int b0, int b1, int b2, int b3, unsigned char* __restrict__ dst, size_t dst_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const double* m = mats + (size_t)blockIdx.z * 9;
    const double xb = (double)((x / bw) * bw), x1 = (double)(x % bw), yd = (double)y;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(m[0], xb), __dmul_rn(m[1], yd)), m[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(m[3], xb), __dmul_rn(m[4], yd)), m[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(m[6], xb), __dmul_rn(m[7], yd)), m[8]);
    double W = __dadd_rn(W0, __dmul_rn(m[6], x1));
    W = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
    const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(m[0], x1)), W)));
    const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(m[3], x1)), W)));
    const int X = (int)__double2ll_rn(fX), Y = (int)__double2ll_rn(fY);
    int sx = X >> 5, sy = Y >> 5;
    sx = max(-32768, min(32767, sx));
    sy = max(-32768, min(32767, sy));
    const int ax = X & 31, ay = Y & 31;
    const int w0 = (32 - ax) * (32 - ay) * 32, w1 = ax * (32 - ay) * 32, w2 = (32 - ax) * ay * 32, w3 = ax * ay * 32;
    unsigned char* out = dst + (size_t)blockIdx.z * dst_stride + ((size_t)y * dw + x) * channels;
    const int border[4] = {b0, b1, b2, b3};
    for (int c = 0; c < channels; ++c) {
        const int v = fetch(src, row_stride, channels, sw, sh, sx, sy, c, border[c]) * w0 +
                      fetch(src, row_stride, channels, sw, sh, sx + 1, sy, c, border[c]) * w1 +
                      fetch(src, row_stride, channels, sw, sh, sx, sy + 1, c, border[c]) * w2 +
                      fetch(src, row_stride, channels, sw, sh, sx + 1, sy + 1, c, border[c]) * w3;
        out[c] = (unsigned char)min(255, max(0, (v + (1 << 14)) >> 15));
    }
}

// cv::invert of a 3 x 3 f64 matrix (closed form; a singular matrix inverts to zeros)
void invert3(const double* m, double* t) {
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0) {
        for (int i = 0; i < 9; ++i) t[i] = 0;
        return;
    }
    d = 1.0 / d;
    t[0] = (m[4] * m[8] - m[5] * m[7]) * d; t[1] = (m[2] * m[7] - m[1] * m[8]) * d; t[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    t[3] = (m[5] * m[6] - m[3] * m[8]) * d; t[4] = (m[0] * m[8] - m[2] * m[6]) * d; t[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    t[6] = (m[3] * m[7] - m[4] * m[6]) * d; t[7] = (m[1] * m[6] - m[0] * m[7]) * d; t[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

int block_width(int dw, int dh) {   // WarpPerspectiveInvoker's block shape
    int bh0 = std::min(16, dh);
    int bw0 = std::min(1024 / bh0, dw);
    return bw0;
}

int launch_warp(dunk_ctx* ctx, cudaStream_t st, const unsigned char* src_dev, int rows, int cols, int channels, int row_stride,
                const double* mats_dev, int n, int out_rows, int out_cols, const int* border, unsigned char* dst_dev) {
    ProfScope ps(ctx, st, "warp.perspective", (double)n * out_rows * out_cols * channels * 2.0);
    k_warp_perspective<<<dim3(div_up(out_cols, 256), out_rows, n), 256, 0, st>>>(
        src_dev, cols, rows, channels, row_stride, mats_dev, out_cols, out_rows, block_width(out_cols, out_rows), border[0],
        border[1], border[2], border[3], dst_dev, (size_t)out_rows * out_cols * channels);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    return DUNK_OK;
}

int check_warp_args(const char* fn, int rows, int cols, int channels, int row_stride, int n, int out_rows, int out_cols) {
    DUNK_REQUIRE(rows > 0 && cols > 0, DUNK_ERR_ASSERT, "%s: empty source image", fn);
    DUNK_REQUIRE(channels >= 1 && channels <= 4, DUNK_ERR_ASSERT, "%s: %d channels (1..4 x 8-bit supported)", fn, channels);
    DUNK_REQUIRE(row_stride >= cols * channels, DUNK_ERR_BAD_ARG, "%s: row stride < row bytes", fn);
    DUNK_REQUIRE(out_rows > 0 && out_cols > 0 && out_rows <= 65535, DUNK_ERR_ASSERT, "%s: output size %dx%d", fn, out_cols, out_rows);
    DUNK_REQUIRE(n >= 0 && n <= 65535, DUNK_ERR_BAD_ARG, "%s: batch of %d outputs (0..65535)", fn, n);
    return DUNK_OK;
}

}  // namespace
}  // namespace dunk

using namespace dunk;

extern "C" {

int dunk_warp_perspective_batch_dev(dunk_ctx* ctx, int slot, const void* src_dev, int rows, int cols, int channels,
                                    int row_stride_bytes, const double* M, int n, int out_rows, int out_cols,
                                    const double* border_value, void* dst_dev) {
    DUNK_REQUIRE(ctx && src_dev && M && dst_dev, DUNK_ERR_BAD_ARG, "dunk_warp_perspective_batch_dev: NULL argument");
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_warp_perspective_batch_dev: bad slot");
    int rc = check_warp_args("dunk_warp_perspective_batch_dev", rows, cols, channels, row_stride_bytes, n, out_rows, out_cols);
    if (rc) return rc;
    if (n == 0) return DUNK_OK;
    DUNK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[slot].stream;
    std::vector<double> inv((size_t)n * 9);
    for (int i = 0; i < n; ++i) invert3(M + (size_t)i * 9, inv.data() + (size_t)i * 9);
    double* d_m = (double*)ctx->dev_scratch(slot, (size_t)n * 72);
    if (!d_m) return DUNK_ERR_NO_MEM;
    DUNK_CUDA(cudaMemcpyAsync(d_m, inv.data(), (size_t)n * 72, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaStreamSynchronize(st));   // `inv` is pageable and goes out of scope
    int border[4] = {1, 1, 1, 1};           // the reference's Scalar(1, 1, 1, 1)
    if (border_value)
        for (int c = 0; c < 4; ++c) border[c] = (int)std::min(255.0, std::max(0.0, std::nearbyint(border_value[c])));
    return launch_warp(ctx, st, (const unsigned char*)src_dev, rows, cols, channels, row_stride_bytes, d_m, n, out_rows, out_cols,
                       border, (unsigned char*)dst_dev);
}

int dunk_warp_perspective(dunk_ctx* ctx, const uint8_t* src, int rows, int cols, int channels, int row_stride_bytes,
                          const double* M, int out_rows, int out_cols, const double* border_value, uint8_t* dst) {
    DUNK_REQUIRE(ctx && M && dst, DUNK_ERR_BAD_ARG, "dunk_warp_perspective: NULL argument");
    DUNK_REQUIRE(src, DUNK_ERR_ASSERT, "dunk_warp_perspective: empty source image");
    int rc = check_warp_args("dunk_warp_perspective", rows, cols, channels, row_stride_bytes, 1, out_rows, out_cols);
    if (rc) return rc;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const size_t src_bytes = (size_t)rows * row_stride_bytes, dst_bytes = (size_t)out_rows * out_cols * channels;
    void* scratch = ctx->dev_scratch(g.s, Carver::need(src_bytes) + Carver::need(dst_bytes) + Carver::need(72));
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    unsigned char* d_src = cv.take<unsigned char>(src_bytes);
    unsigned char* d_dst = cv.take<unsigned char>(dst_bytes);
    double* d_m = cv.take<double>(9);
    double inv[9];
    invert3(M, inv);
    DUNK_CUDA(cudaMemcpyAsync(d_src, src, src_bytes, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_m, inv, 72, cudaMemcpyHostToDevice, st));
    int border[4] = {1, 1, 1, 1};
    if (border_value)
        for (int c = 0; c < 4; ++c) border[c] = (int)std::min(255.0, std::max(0.0, std::nearbyint(border_value[c])));
    if ((rc = launch_warp(ctx, st, d_src, rows, cols, channels, row_stride_bytes, d_m, 1, out_rows, out_cols, border, d_dst))) return rc;
    DUNK_CUDA(cudaMemcpyAsync(dst, d_dst, dst_bytes, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    return DUNK_OK;
}

}  // extern "C"
