// The steps either side of the hot path, batched on the device:
//   * band_merger / f32_to_u8 / gamma_correction (geotiff_extractor/src/image_extractor/mod.rs:346-378,
//     402-422): three f32 bands -> RGBA8 (min-max normalise, gamma 1/2.2, round half away, NaN /
//     out-of-range -> 0, alpha 0 where every band is NaN);
//   * raster_to_mat (homographier/src/homographier/mod.rs:183-220): RGBA -> BGRA swizzle;
//   * get_world_coordinates (feature_database/src/elevationdb.rs:64-104): pixel -> geotransform ->
//     nearest elevation sample -> WGS-84 geodetic -> ECEF (EPSG:4326 -> EPSG:4978, the closed form PROJ
//     evaluates), i.e. the 3-D object points pnp_solver_ransac consumes.
// All HBM-bound streaming kernels: one thread per pixel / point, coalesced loads and stores.
#include <cmath>
#include <mutex>
#include "ctx.h"
#include "gamma_lut.cuh"
#include "geo.cuh"

namespace dunk {
namespace {

__device__ float g_gamma_thr[256];

__global__ void k_gamma_thresholds() {
    const int k = threadIdx.x;
    if (k == 0) { g_gamma_thr[0] = 0.f; return; }
    // smallest bit pattern in [0, bits(1.0f)] whose value is >= k (positive floats order like their bits)
    unsigned lo = 0, hi = 0x3f800000u;          // value(hi) = 255 >= k always
    while (lo < hi) {
        const unsigned mid = lo + (hi - lo) / 2;
        if (gamma_u8_direct(__uint_as_float(mid)) >= k) hi = mid; else lo = mid + 1;
    }
    g_gamma_thr[k] = __uint_as_float(lo);
}

// every f32 in [0, 1]: table result vs direct formula
__global__ void __launch_bounds__(256) k_gamma_selftest(unsigned long long* __restrict__ mismatches) {
    __shared__ float s_thr[256];
    s_thr[threadIdx.x] = g_gamma_thr[threadIdx.x];
    __syncthreads();
    unsigned long long bad = 0;
    for (unsigned long long bits = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; bits <= 0x3f800000ull;
         bits += (unsigned long long)gridDim.x * blockDim.x) {
        const float fl = __uint_as_float((unsigned)bits);
        bad += gamma_u8_lut(s_thr, fl) != gamma_u8_direct(fl);
    }
    if (bad) atomicAdd(mismatches, bad);
}

__global__ void __launch_bounds__(256)
k_band_merger(const float* __restrict__ r, const float* __restrict__ g, const float* __restrict__ b, long long n, float rmin,
              float rmax, float gmin, float gmax, float bmin, float bmax, const float* __restrict__ thr,
              uchar4* __restrict__ rgba, int bgra) {
    __shared__ float s_thr[256];
    s_thr[threadIdx.x] = thr[threadIdx.x];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float vr = r[i], vg = g[i], vb = b[i];
    uchar4 o;
    const unsigned char cr = f32_to_u8_lut(s_thr, vr, rmin, rmax), cg = f32_to_u8_lut(s_thr, vg, gmin, gmax),
                        cb = f32_to_u8_lut(s_thr, vb, bmin, bmax);
    o.x = bgra ? cb : cr;
    o.y = cg;
    o.z = bgra ? cr : cb;
    o.w = (isnan(vr) && isnan(vg) && isnan(vb)) ? 0 : 255;
    rgba[i] = o;
}

__global__ void __launch_bounds__(256) k_swizzle_rb(const uchar4* __restrict__ src, long long n, uchar4* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uchar4 v = src[i];
    dst[i] = make_uchar4(v.z, v.y, v.x, v.w);
}

__global__ void __launch_bounds__(256)
k_world_coordinates(const double* __restrict__ px, const double* __restrict__ py, long long n, GeoParams p,
                    const double* __restrict__ heights, double* __restrict__ xyz, int* __restrict__ n_missing) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o[3];
    if (!world_point(p, heights, px[i], py[i], o)) atomicAdd(n_missing, 1);
    xyz[3 * i + 0] = o[0];
    xyz[3 * i + 1] = o[1];
    xyz[3 * i + 2] = o[2];
}

// GDALInvGeoTransform
bool invert_geotransform(const double* gt, double* out) {
    if (gt[2] == 0.0 && gt[4] == 0.0 && gt[1] != 0.0 && gt[5] != 0.0) {
        out[0] = -gt[0] / gt[1]; out[1] = 1.0 / gt[1]; out[2] = 0.0;
        out[3] = -gt[3] / gt[5]; out[4] = 0.0; out[5] = 1.0 / gt[5];
        return true;
    }
    const double det = gt[1] * gt[5] - gt[2] * gt[4];
    const double mag = std::max(std::max(std::fabs(gt[1]), std::fabs(gt[2])), std::max(std::fabs(gt[4]), std::fabs(gt[5])));
    if (std::fabs(det) <= 1e-10 * mag * mag) return false;
    const double inv = 1.0 / det;
    out[1] = gt[5] * inv; out[4] = -gt[4] * inv; out[2] = -gt[2] * inv; out[5] = gt[1] * inv;
    out[0] = (gt[2] * gt[3] - gt[0] * gt[5]) * inv;
    out[3] = (-gt[1] * gt[3] + gt[0] * gt[4]) * inv;
    return true;
}

}  // namespace

const float* gamma_table(dunk_ctx* ctx, cudaStream_t st) {
    static std::mutex mu;
    static bool ready[64] = {};
    std::lock_guard<std::mutex> lk(mu);
    const int dev = ctx->device;
    void* p = nullptr;
    if (dev < 0 || dev >= 64 || cudaGetSymbolAddress(&p, g_gamma_thr) != cudaSuccess) return nullptr;
    if (!ready[dev]) {
        k_gamma_thresholds<<<1, 256, 0, st>>>();
        ctx->launches.fetch_add(1);
        // once per device and process; synchronous so that every other stream may read the table afterwards
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return nullptr;
        ready[dev] = true;
    }
    return (const float*)p;
}

}  // namespace dunk

using namespace dunk;

extern "C" {

/* compares the threshold table with the direct pow formula on every f32 in [0, 1]; *mismatches must be 0 */
int dunk_selftest_gamma_lut(dunk_ctx* ctx, uint64_t* mismatches) {
    DUNK_REQUIRE(ctx && mismatches, DUNK_ERR_BAD_ARG, "dunk_selftest_gamma_lut: NULL argument");
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    DUNK_REQUIRE(gamma_table(ctx, st), DUNK_ERR_CUDA, "dunk_selftest_gamma_lut: gamma table");
    unsigned long long* d = (unsigned long long*)ctx->dev_scratch(g.s, 256);
    if (!d) return DUNK_ERR_NO_MEM;
    DUNK_CUDA(cudaMemsetAsync(d, 0, 8, st));
    k_gamma_selftest<<<ctx->sm_count * 8, 256, 0, st>>>(d);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    DUNK_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    *mismatches = h;
    return DUNK_OK;
}

int dunk_band_merger(dunk_ctx* ctx, const float* red, const float* green, const float* blue, int64_t n, const double* min_max,
                     int bgra, uint8_t* out_rgba) {
    DUNK_REQUIRE(ctx && min_max && n >= 0, DUNK_ERR_BAD_ARG, "dunk_band_merger: bad argument");
    if (n == 0) return DUNK_OK;
    DUNK_REQUIRE(red && green && blue && out_rgba, DUNK_ERR_BAD_ARG, "dunk_band_merger: NULL band / output");
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    // chunked so a 10980^2 scene (3 x 482 MB of f32) streams through a bounded scratch
    const int64_t chunk = std::min<int64_t>(n, (int64_t)16 << 20);
    void* scratch = ctx->dev_scratch(g.s, 3 * Carver::need((size_t)chunk * 4) + Carver::need((size_t)chunk * 4));
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    float* d_r = cv.take<float>(chunk);
    float* d_g = cv.take<float>(chunk);
    float* d_b = cv.take<float>(chunk);
    uchar4* d_o = cv.take<uchar4>(chunk);
    const float* thr = gamma_table(ctx, st);
    DUNK_REQUIRE(thr, DUNK_ERR_CUDA, "dunk_band_merger: gamma table");
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t m = std::min(chunk, n - off);
        DUNK_CUDA(cudaMemcpyAsync(d_r, red + off, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        DUNK_CUDA(cudaMemcpyAsync(d_g, green + off, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        DUNK_CUDA(cudaMemcpyAsync(d_b, blue + off, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        {
            ProfScope ps(ctx, st, "geo.band_merger", (double)m * 16.0);
            k_band_merger<<<div_up(m, 256), 256, 0, st>>>(d_r, d_g, d_b, m, (float)min_max[0], (float)min_max[1], (float)min_max[2],
                                                        (float)min_max[3], (float)min_max[4], (float)min_max[5], thr, d_o, bgra);
            ctx->launches.fetch_add(1);
            DUNK_CUDA(cudaGetLastError());
        }
        DUNK_CUDA(cudaMemcpyAsync(out_rgba + off * 4, d_o, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
        DUNK_CUDA(cudaStreamSynchronize(st));
    }
    return DUNK_OK;
}

int dunk_band_merger_dev(dunk_ctx* ctx, int slot, const void* red_dev, const void* green_dev, const void* blue_dev, int64_t n,
                         const double* min_max, int bgra, void* out_dev) {
    DUNK_REQUIRE(ctx && min_max && n >= 0 && slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG,
                 "dunk_band_merger_dev: bad argument");
    if (n == 0) return DUNK_OK;
    DUNK_REQUIRE(red_dev && green_dev && blue_dev && out_dev, DUNK_ERR_BAD_ARG, "dunk_band_merger_dev: NULL pointer");
    DUNK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[slot].stream;
    const float* thr = gamma_table(ctx, st);
    DUNK_REQUIRE(thr, DUNK_ERR_CUDA, "dunk_band_merger_dev: gamma table");
    ProfScope ps(ctx, st, "geo.band_merger", (double)n * 16.0);
    k_band_merger<<<div_up(n, 256), 256, 0, st>>>((const float*)red_dev, (const float*)green_dev, (const float*)blue_dev, n,
                                                (float)min_max[0], (float)min_max[1], (float)min_max[2], (float)min_max[3],
                                                (float)min_max[4], (float)min_max[5], thr, (uchar4*)out_dev, bgra);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    return DUNK_OK;
}

int dunk_raster_to_mat(dunk_ctx* ctx, const uint8_t* rgba, int w, int h, uint8_t* bgra) {
    DUNK_REQUIRE(ctx, DUNK_ERR_BAD_ARG, "dunk_raster_to_mat: ctx is NULL");
    DUNK_REQUIRE(w > 0 && h > 0 && rgba && bgra, DUNK_ERR_ASSERT, "dunk_raster_to_mat: empty raster");
    const int64_t n = (int64_t)w * h;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    void* scratch = ctx->dev_scratch(g.s, 2 * Carver::need((size_t)n * 4));
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    uchar4* d_i = cv.take<uchar4>(n);
    uchar4* d_o = cv.take<uchar4>(n);
    DUNK_CUDA(cudaMemcpyAsync(d_i, rgba, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    k_swizzle_rb<<<div_up(n, 256), 256, 0, st>>>(d_i, n, d_o);
    ctx->launches.fetch_add(1);
    DUNK_CUDA(cudaGetLastError());
    DUNK_CUDA(cudaMemcpyAsync(bgra, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    return DUNK_OK;
}

int dunk_elevation_create(dunk_ctx* ctx, const double* gt_dataset, const double* gt_elevation, const double* heights, int x_size,
                          int y_size, dunk_elevation** out) {
    DUNK_REQUIRE(ctx && gt_dataset && out, DUNK_ERR_BAD_ARG, "dunk_elevation_create: NULL argument");
    *out = nullptr;
    dunk_elevation* e = new dunk_elevation();
    e->ctx = ctx;
    for (int i = 0; i < 6; ++i) e->gt_dataset[i] = gt_dataset[i];
    if (gt_elevation) {
        if (!(heights && x_size > 0 && y_size > 0)) {
            delete e;
            set_error("dunk_elevation_create: an elevation geotransform needs a height raster");
            return DUNK_ERR_BAD_ARG;
        }
        if (!invert_geotransform(gt_elevation, e->gt_elev_inv)) {
            delete e;
            set_error("dunk_elevation_create: the elevation geotransform is not invertible");   // the reference .expect()s here
            return DUNK_ERR_BAD_ARG;
        }
        e->has_elevation = 1;
        e->x_size = x_size;
        e->y_size = y_size;
        cudaSetDevice(ctx->device);
        const size_t bytes = (size_t)x_size * y_size * 8;
        if (cudaMalloc(&e->heights, bytes) != cudaSuccess) {
            cudaGetLastError();
            delete e;
            set_error("dunk_elevation_create: allocating %zu bytes failed", bytes);
            return DUNK_ERR_NO_MEM;
        }
        if (cudaMemcpy(e->heights, heights, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(e->heights);
            delete e;
            set_error("dunk_elevation_create: upload failed");
            return DUNK_ERR_CUDA;
        }
    }
    *out = e;
    return DUNK_OK;
}

void dunk_elevation_destroy(dunk_elevation* e) {
    if (!e) return;
    if (e->heights) {
        cudaSetDevice(e->ctx->device);
        cudaFree(e->heights);
    }
    delete e;
}

int dunk_world_coordinates(dunk_elevation* e, const double* px, const double* py, int64_t n, double* xyz, int* n_missing) {
    DUNK_REQUIRE(e && n >= 0, DUNK_ERR_BAD_ARG, "dunk_world_coordinates: bad argument");
    if (n_missing) *n_missing = 0;
    if (n == 0) return DUNK_OK;
    DUNK_REQUIRE(px && py && xyz, DUNK_ERR_BAD_ARG, "dunk_world_coordinates: NULL pointer");
    dunk_ctx* ctx = e->ctx;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    void* scratch = ctx->dev_scratch(g.s, 2 * Carver::need((size_t)n * 8) + Carver::need((size_t)n * 24) + Carver::need(16));
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    double* d_x = cv.take<double>(n);
    double* d_y = cv.take<double>(n);
    double* d_o = cv.take<double>((size_t)n * 3);
    int* d_miss = cv.take<int>(4);
    DUNK_CUDA(cudaMemcpyAsync(d_x, px, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemcpyAsync(d_y, py, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaMemsetAsync(d_miss, 0, 4, st));
    const GeoParams p = make_geo_params(e);
    {
        ProfScope ps(ctx, st, "geo.world_coordinates", (double)n * 40.0);
        k_world_coordinates<<<div_up(n, 256), 256, 0, st>>>(d_x, d_y, n, p, e->heights, d_o, d_miss);
        ctx->launches.fetch_add(1);
        DUNK_CUDA(cudaGetLastError());
    }
    int miss = 0;
    DUNK_CUDA(cudaMemcpyAsync(xyz, d_o, (size_t)n * 24, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaMemcpyAsync(&miss, d_miss, 4, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    if (n_missing) *n_missing = miss;
    return DUNK_OK;
}

}  // extern "C"
