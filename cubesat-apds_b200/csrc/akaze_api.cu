// C ABI for stage 1 (AKAZE extraction) + workspace management.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "akaze.h"
#include "match.h"

namespace dunk {

static size_t al(size_t b) { return (b + 255) & ~size_t(255); }

static int pow2_at_least(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

size_t akaze_workspace_bytes(const LevelTable& lt, int frames, int cand_cap, int kp_cap) {
    const size_t pyr = lt.pyramid_floats, plane = (size_t)lt.width * lt.height;
    int total_rows = 0;
    for (int i = 0; i < lt.n_levels; ++i) total_rows += lt.lv[i].h;
    size_t b = 0;
    b += 4 * al(frames * pyr * 4);                       // Lt Lx Ly Ldet
    b += 3 * al(frames * plane * 4);                     // Lsmooth Lflow Ltmp
    b += al(frames * 4) + al((size_t)frames * 300 * 4) + al(frames * 4);
    b += 2 * al((size_t)frames * cand_cap * sizeof(Cand));
    b += al(frames * 4);
    b += 2 * al((size_t)frames * (total_rows + 1) * 4) + al((size_t)frames * total_rows * 4);
    b += al((size_t)frames * cand_cap);
    b += al((size_t)frames * cand_cap * 3 * 4);
    b += al((size_t)frames * kp_cap * sizeof(DunkKeyPoint));
    b += al(frames * 4) + al((kMaxLevels + 1) * 4);
    b += al((size_t)frames * kp_cap * 64);
    b += al((size_t)frames * pow2_at_least(kp_cap) * 8);
    return b;
}

void akaze_carve_workspace(void* base, const LevelTable& lt, int frames, int cand_cap, int kp_cap, AkazeWorkspace* ws) {
    const size_t pyr = lt.pyramid_floats, plane = (size_t)lt.width * lt.height;
    int total_rows = 0;
    for (int i = 0; i < lt.n_levels; ++i) total_rows += lt.lv[i].h;
    char* p = (char*)base;
    auto take = [&](size_t bytes) { void* r = p; p += al(bytes); return r; };
    ws->frames = frames; ws->cand_cap = cand_cap; ws->kp_cap = kp_cap; ws->total_rows = total_rows;
    ws->Lt = (float*)take(frames * pyr * 4);
    ws->Lx = (float*)take(frames * pyr * 4);
    ws->Ly = (float*)take(frames * pyr * 4);
    ws->Ldet = (float*)take(frames * pyr * 4);
    ws->Lsmooth = (float*)take(frames * plane * 4);
    ws->Lflow = (float*)take(frames * plane * 4);
    ws->Ltmp = (float*)take(frames * plane * 4);
    ws->hmax = (float*)take(frames * 4);
    ws->hist = (int*)take((size_t)frames * 300 * 4);
    ws->kcontrast = (float*)take(frames * 4);
    ws->cand_raw = (Cand*)take((size_t)frames * cand_cap * sizeof(Cand));
    ws->cand = (Cand*)take((size_t)frames * cand_cap * sizeof(Cand));
    ws->cand_count = (int*)take(frames * 4);
    ws->row_count = (int*)take((size_t)frames * (total_rows + 1) * 4);
    ws->row_start = (int*)take((size_t)frames * (total_rows + 1) * 4);
    ws->row_fill = (int*)take((size_t)frames * total_rows * 4);
    ws->state = (unsigned char*)take((size_t)frames * cand_cap);
    ws->aux = (int*)take((size_t)frames * cand_cap * 3 * 4);
    ws->kps = (DunkKeyPoint*)take((size_t)frames * kp_cap * sizeof(DunkKeyPoint));
    ws->kp_count = (int*)take(frames * 4);
    ws->kp_level_start = (int*)take((kMaxLevels + 1) * 4);
    ws->desc64 = (uint4*)take((size_t)frames * kp_cap * 64);
    ws->sort_keys = (float*)take((size_t)frames * pow2_at_least(kp_cap) * 8);
}

// default raw-candidate capacity per frame: 3x3 maxima are sparse (<= 1/4 of the pixels in theory,
// a few 1e-3 in practice); 1/32 of the level-0 pixels leaves a wide margin
static int default_cand_cap(int w, int h) {
    long long c = (long long)w * h / 32;
    c = std::max<long long>(c, 2048);
    c = std::min<long long>(c, 1 << 20);
    return (int)c;
}

int akaze_run(dunk_ctx* ctx, cudaStream_t st, const LevelTable& lt, const AkazeWorkspace& ws,
              const unsigned char* images_dev, size_t image_stride, int row_stride, int channels, int frames,
              int max_points) {
    int rc = akaze_build_scale_space(ctx, st, lt, ws, images_dev, image_stride, row_stride, channels, frames);
    if (rc) return rc;
    if ((rc = akaze_detect(ctx, st, lt, ws, frames, 0.001f, max_points))) return rc;
    return akaze_describe(ctx, st, lt, ws, frames);
}

namespace {
// exclusive scan of the per-frame keypoint counts (frames <= 64 per sub-batch) + the largest raw-extrema count
__global__ void k_out_offsets(const int* __restrict__ counts, const int* __restrict__ cand_count, int frames, int* __restrict__ off) {
    if (threadIdx.x == 0) {
        int acc = 0, cmax = 0;
        for (int f = 0; f < frames; ++f) {
            off[f] = acc;
            acc += counts[f];
            cmax = max(cmax, cand_count[f]);
        }
        off[frames] = acc;
        off[frames + 1] = cmax;
    }
}
// every frame's keypoints and 61-byte descriptor rows, packed back to back in frame order
__global__ void __launch_bounds__(256)
k_pack_outputs(const DunkKeyPoint* __restrict__ kps, const uint4* __restrict__ desc64, int kp_cap, const int* __restrict__ counts,
               const int* __restrict__ off, DunkKeyPoint* __restrict__ out_kps, uint8_t* __restrict__ out_desc) {
    const int f = blockIdx.y, n = counts[f];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // byte-quad index inside the frame's descriptor block
    const size_t base = (size_t)off[f];
    if (i < n) out_kps[base + i] = kps[(size_t)f * kp_cap + i];
    const unsigned char* src = (const unsigned char*)(desc64 + (size_t)f * kp_cap * 4);
    for (int k = i; k < n * 61; k += gridDim.x * blockDim.x) out_desc[base * 61 + k] = src[(size_t)(k / 61) * 64 + k % 61];
}
}  // namespace

static int launch_pack_outputs(dunk_ctx* ctx, cudaStream_t st, const AkazeWorkspace& ws, int frames, int* d_off, DunkKeyPoint* d_kps,
                               uint8_t* d_desc61) {
    k_out_offsets<<<1, 32, 0, st>>>(ws.kp_count, ws.cand_count, frames, d_off);
    DUNK_KERNEL_CHECK(ctx);
    k_pack_outputs<<<dim3(div_up(ws.kp_cap, 256), frames), 256, 0, st>>>(ws.kps, ws.desc64, ws.kp_cap, ws.kp_count, d_off, d_kps, d_desc61);
    DUNK_KERNEL_CHECK(ctx);
    return DUNK_OK;
}

}  // namespace dunk

using namespace dunk;

extern "C" {

int dunk_akaze_extract_batch(dunk_ctx* ctx, const uint8_t* images, int n_frames, int rows, int cols, int channels,
                             int row_stride_bytes, size_t frame_stride_bytes, int max_points, DunkKeyPoint* kps,
                             uint8_t* desc, int cap_per_frame, int* counts) {
    DUNK_REQUIRE(ctx && counts, DUNK_ERR_BAD_ARG, "dunk_akaze_extract: NULL ctx / counts");
    DUNK_REQUIRE(n_frames >= 0, DUNK_ERR_BAD_ARG, "dunk_akaze_extract: n_frames < 0");
    if (n_frames == 0) return DUNK_OK;
    // OpenCV: detectAndCompute on an empty image asserts (-215)
    DUNK_REQUIRE(images && rows > 0 && cols > 0, DUNK_ERR_ASSERT, "dunk_akaze_extract: empty image");
    DUNK_REQUIRE(channels == 1 || channels == 3 || channels == 4, DUNK_ERR_ASSERT,
                 "dunk_akaze_extract: %d channels (8UC1, 8UC3 BGR or 8UC4 BGRA expected)", channels);
    DUNK_REQUIRE(rows >= 16 && cols >= 16, DUNK_ERR_ASSERT, "dunk_akaze_extract: image %dx%d too small", cols, rows);
    DUNK_REQUIRE(row_stride_bytes >= cols * channels, DUNK_ERR_BAD_ARG, "dunk_akaze_extract: row stride < row bytes");
    DUNK_REQUIRE(kps && desc && cap_per_frame > 0, DUNK_ERR_BAD_ARG, "dunk_akaze_extract: NULL / empty output");
    if (max_points <= 0) max_points = 0;   // AKAZE: max_points <= 0 means unlimited
    const size_t img_bytes = (size_t)rows * row_stride_bytes;
    if (frame_stride_bytes == 0) frame_stride_bytes = img_bytes;
    DUNK_REQUIRE(frame_stride_bytes >= img_bytes, DUNK_ERR_BAD_ARG, "dunk_akaze_extract: frame stride < frame bytes");

    const LevelTable lt = make_level_table(cols, rows);
    int cand_cap = default_cand_cap(cols, rows);
    static const bool trace = getenv("DUNK_TRACE") != nullptr;     // host-side phase times on stderr (diagnostics)
    auto ms_since = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    const auto t_entry = std::chrono::steady_clock::now();
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    // workspace budget: what the slot already holds is used without asking the driver (cudaMemGetInfo takes up to
    // 13 ms while other streams are busy); a call that needs more queries the free memory once
    size_t budget = g.slot().dev_bytes;
    {
        const size_t want = (akaze_workspace_bytes(lt, 1, cand_cap, cand_cap) + 2 * frame_stride_bytes + 4096) * (size_t)std::min(n_frames, 128);
        if (budget < want) {
            size_t free_b = 0, total_b = 0;
            DUNK_CUDA(cudaMemGetInfo(&free_b, &total_b));
            budget = std::min<size_t>((free_b + g.slot().dev_bytes) / 2, (size_t)16 << 30);
        }
    }
    if (trace) fprintf(stderr, "[dunk] +%.1f ms slot %d acquired\n", ms_since(t_entry), g.s);
    cudaEvent_t done[2] = {g.slot().ev0, g.slot().ev1};
    cudaEvent_t copied[2] = {g.slot().ev2, g.slot().ev3};
    static thread_local cudaEvent_t t_start[2] = {nullptr, nullptr};      // trace only
    int f0 = 0;
    // The raw-candidate capacity (w*h/32 by default) is exceeded only by pathological textures (random 4x4 blocks
    // reach w*h/29); k_extrema keeps counting past the capacity, so on overflow the sub-batch is simply re-run
    // with a workspace carved for the observed maximum — the result never depends on which candidates were dropped.
    while (f0 < n_frames) {
        const int kp_cap = cand_cap;
        // sub-batch so the workspace stays bounded (~100 MB per 1024^2 frame)
        const size_t per_frame = akaze_workspace_bytes(lt, 1, cand_cap, kp_cap) + 2 * frame_stride_bytes + 4096;
        int sub = (int)std::max<size_t>(1, std::min<size_t>(n_frames - f0, budget / per_frame));
        // the small octaves are latency-bound at 64 frames (T(n) ~ 1.4 ms + 0.089 ms * n on B200): 128-frame sub-batches
        sub = std::min(sub, 128);
        const size_t ws_bytes = akaze_workspace_bytes(lt, sub, cand_cap, kp_cap);
        // Two-deep software pipeline over sub-batches (one stream): while the device works on sub-batch i the calling
        // thread stages sub-batch i + 1 into the other pinned half (pageable caller memory would otherwise go through
        // the driver's blocking staged copy at ~11 GB/s) and unpacks the results of sub-batch i - 1.  The packed
        // outputs of a sub-batch leave in ONE D2H copy each (keypoints, descriptors).
        const size_t in_half = al((size_t)sub * frame_stride_bytes);
        const size_t out_rows = (size_t)sub * kp_cap;
        const size_t out_half = al(out_rows * 61) + al(out_rows * sizeof(DunkKeyPoint)) + al((size_t)(sub + 2) * 4);
        void* scratch = ctx->dev_scratch(g.s, al(ws_bytes) + 2 * in_half + 2 * out_half);
        if (!scratch) return DUNK_ERR_NO_MEM;
        // pinned: two input halves, two offset blocks, one output block sized for the typical density (grown on demand)
        const size_t pin_off = 2 * in_half, pin_out = pin_off + 2 * al((size_t)(sub + 2) * 4);
        size_t pin_out_cap = std::max<size_t>((size_t)sub * 4096 * 89, 1 << 20);
        unsigned char* pin = (unsigned char*)ctx->pin_scratch(g.s, pin_out + pin_out_cap);
        if (!pin) return DUNK_ERR_NO_MEM;
        AkazeWorkspace ws;
        akaze_carve_workspace(scratch, lt, sub, cand_cap, kp_cap, &ws);
        unsigned char* d_in = (unsigned char*)scratch + al(ws_bytes);
        unsigned char* d_out = d_in + 2 * in_half;
        auto d_desc61 = [&](int h) { return (uint8_t*)(d_out + h * out_half); };
        auto d_kps = [&](int h) { return (DunkKeyPoint*)(d_out + h * out_half + al(out_rows * 61)); };
        auto d_off = [&](int h) { return (int*)(d_out + h * out_half + al(out_rows * 61) + al(out_rows * sizeof(DunkKeyPoint))); };
        auto h_off = [&](int h) { return (int*)(pin + pin_off + h * al((size_t)(sub + 2) * 4)); };
        bool overflow = false;
        // chunk schedule: a first chunk of at most 64 frames (the device starts after half a staging copy), full
        // sub-batches after it
        const int first_f0 = f0;
        auto chunk_len = [&](int fs) {
            const int left = n_frames - fs;
            return (fs == first_f0 && left > 64) ? std::min(sub, 64) : std::min(sub, left);
        };
        cudaStream_t st2 = g.slot().stream2;      // copies in both directions run beside the kernels of the other half
        // stage sub-batch starting at frame `fs` into half h (asynchronous after the host memcpy)
        auto stage = [&](int fs, int h) -> int {
            const int nf = chunk_len(fs);
            const size_t in_bytes = (size_t)nf * frame_stride_bytes;
            const auto t0 = std::chrono::steady_clock::now();
            par_memcpy(pin + h * in_half, images + (size_t)fs * frame_stride_bytes, in_bytes);
            if (trace) fprintf(stderr, "[dunk] +%.1f ms stage %d: memcpy %.1f MB in %.2f ms\n", ms_since(t_entry), fs, in_bytes / 1e6, ms_since(t0));
            // the copy rides the slot's second stream so that it overlaps the previous sub-batch's kernels; the
            // compute stream waits for it (the input half was released when sub-batch i - 1 finished, see below)
            DUNK_CUDA(cudaMemcpyAsync(d_in + h * in_half, pin + h * in_half, in_bytes, cudaMemcpyHostToDevice, st2));
            DUNK_CUDA(cudaEventRecord(copied[h], st2));
            DUNK_CUDA(cudaStreamWaitEvent(st, copied[h], 0));
            if (trace) {
                if (!t_start[h]) cudaEventCreate(&t_start[h]);
                cudaEventRecord(t_start[h], st);
            }
            int rc = akaze_run(ctx, st, lt, ws, d_in + h * in_half, frame_stride_bytes, row_stride_bytes, channels, nf, max_points);
            if (rc) return rc;
            if ((rc = launch_pack_outputs(ctx, st, ws, nf, d_off(h), d_kps(h), d_desc61(h)))) return rc;
            DUNK_CUDA(cudaMemcpyAsync(h_off(h), d_off(h), (size_t)(nf + 2) * 4, cudaMemcpyDeviceToHost, st));
            DUNK_CUDA(cudaEventRecord(done[h], st));
            return DUNK_OK;
        };
        int rc = stage(f0, 0);
        if (rc) return rc;
        for (int h = 0; f0 < n_frames && !overflow; h ^= 1) {
            const int nf = chunk_len(f0);
            const int next = f0 + nf;
            // host stages sub-batch i + 1 while the device runs sub-batch i (its pinned half and packed-output half
            // were released when sub-batch i - 1 was unpacked)
            if (next < n_frames && (rc = stage(next, h ^ 1))) return rc;
            const auto tw = std::chrono::steady_clock::now();
            DUNK_CUDA(cudaEventSynchronize(done[h]));
            if (trace) {
                float dev_ms = 0;
                cudaEventElapsedTime(&dev_ms, t_start[h], done[h]);
                fprintf(stderr, "[dunk] +%.1f ms sub-batch %d (%d frames): waited %.2f ms for the device, its kernels took %.2f ms\n",
                        ms_since(t_entry), f0, nf, ms_since(tw), dev_ms);
            }
            const int* off = h_off(h);
            if (off[nf + 1] > cand_cap) {              // re-run from this sub-batch with a larger candidate capacity
                DUNK_REQUIRE(off[nf + 1] <= (1 << 22), DUNK_ERR_NO_MEM, "dunk_akaze_extract: %d raw extrema in one frame", off[nf + 1]);
                DUNK_CUDA(cudaStreamSynchronize(st));
                DUNK_CUDA(cudaStreamSynchronize(st2));
                cand_cap = off[nf + 1] + off[nf + 1] / 4;
                overflow = true;
                break;
            }
            const int total = off[nf];
            for (int f = 0; f < nf; ++f) {
                const int n = off[f + 1] - off[f];
                DUNK_REQUIRE(n <= cap_per_frame, DUNK_ERR_NO_MEM,
                             "dunk_akaze_extract: frame %d has %d keypoints, output capacity %d", f0 + f, n, cap_per_frame);
                counts[f0 + f] = n;
            }
            if (total > 0) {
                if ((size_t)total * 89 > pin_out_cap) {     // denser than typical: grow the pinned block (rare, slow path)
                    DUNK_CUDA(cudaStreamSynchronize(st));
                    DUNK_CUDA(cudaStreamSynchronize(st2));
                    std::vector<int> keep0(h_off(0), h_off(0) + sub + 2), keep1(h_off(1), h_off(1) + sub + 2);
                    pin_out_cap = (size_t)total * 89 * 2;
                    pin = (unsigned char*)ctx->pin_scratch(g.s, pin_out + pin_out_cap);
                    if (!pin) return DUNK_ERR_NO_MEM;
                    memcpy(h_off(0), keep0.data(), keep0.size() * 4);
                    memcpy(h_off(1), keep1.data(), keep1.size() * 4);
                    off = h_off(h);
                }
                DunkKeyPoint* p_kps = (DunkKeyPoint*)(pin + pin_out);
                uint8_t* p_desc = (uint8_t*)(p_kps + total);
                DUNK_CUDA(cudaMemcpyAsync(p_kps, d_kps(h), (size_t)total * sizeof(DunkKeyPoint), cudaMemcpyDeviceToHost, st2));
                DUNK_CUDA(cudaMemcpyAsync(p_desc, d_desc61(h), (size_t)total * 61, cudaMemcpyDeviceToHost, st2));
                const auto td = std::chrono::steady_clock::now();
                DUNK_CUDA(cudaStreamSynchronize(st2));
                const double d2h_ms = ms_since(td);
                const auto tu = std::chrono::steady_clock::now();
                par_for((size_t)nf, [&](size_t f) {
                    const int n = off[f + 1] - off[f];
                    if (n == 0) return;
                    memcpy(kps + (size_t)(f0 + f) * cap_per_frame, p_kps + off[f], (size_t)n * sizeof(DunkKeyPoint));
                    memcpy(desc + (size_t)(f0 + f) * cap_per_frame * 61, p_desc + (size_t)off[f] * 61, (size_t)n * 61);
                });
                if (trace) fprintf(stderr, "[dunk] sub-batch %d: D2H %.1f MB in %.2f ms, unpack %.2f ms\n", f0, total * 89 / 1e6, d2h_ms, ms_since(tu));
            }
            f0 = next;
        }
        DUNK_CUDA(cudaStreamSynchronize(st));
    }
    if (trace) fprintf(stderr, "[dunk] +%.1f ms done\n", ms_since(t_entry));
    return DUNK_OK;
}

int dunk_akaze_extract(dunk_ctx* ctx, const uint8_t* image, int rows, int cols, int channels, int row_stride_bytes,
                       int max_points, DunkKeyPoint* kps, uint8_t* desc, int cap, int* n_out) {
    DUNK_REQUIRE(n_out, DUNK_ERR_BAD_ARG, "dunk_akaze_extract: n_out is NULL");
    *n_out = 0;
    return dunk_akaze_extract_batch(ctx, image, 1, rows, cols, channels, row_stride_bytes, 0, max_points, kps, desc, cap,
                                    n_out);
}

}  // extern "C"

/* per-stage parity hook: scale space of one frame, planes of one evolution level copied back */
extern "C" int dunk_akaze_debug_level(dunk_ctx* ctx, const uint8_t* image, int rows, int cols, int channels,
                                      int row_stride_bytes, int level, float* Lt, float* Lx, float* Ly, float* Ldet,
                                      float* kcontrast, int* level_w, int* level_h, int* n_levels) {
    DUNK_REQUIRE(ctx && image && rows >= 16 && cols >= 16, DUNK_ERR_BAD_ARG, "dunk_akaze_debug_level: bad argument");
    const LevelTable lt = make_level_table(cols, rows);
    if (n_levels) *n_levels = lt.n_levels;
    DUNK_REQUIRE(level >= 0 && level < lt.n_levels, DUNK_ERR_OUT_OF_RANGE, "dunk_akaze_debug_level: level %d of %d",
                 level, lt.n_levels);
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const int cand_cap = 2048;
    const size_t ws_bytes = akaze_workspace_bytes(lt, 1, cand_cap, cand_cap);
    const size_t img_bytes = (size_t)rows * row_stride_bytes;
    void* scratch = ctx->dev_scratch(g.s, al(ws_bytes) + al(img_bytes));
    if (!scratch) return DUNK_ERR_NO_MEM;
    AkazeWorkspace ws;
    akaze_carve_workspace(scratch, lt, 1, cand_cap, cand_cap, &ws);
    unsigned char* d_img = (unsigned char*)scratch + al(ws_bytes);
    DUNK_CUDA(cudaMemcpyAsync(d_img, image, img_bytes, cudaMemcpyHostToDevice, st));
    int rc = akaze_build_scale_space(ctx, st, lt, ws, d_img, img_bytes, row_stride_bytes, channels, 1);
    if (rc) return rc;
    const LevelInfo& e = lt.lv[level];
    const size_t n = (size_t)e.w * e.h * 4;
    if (level_w) *level_w = e.w;
    if (level_h) *level_h = e.h;
    if (Lt) DUNK_CUDA(cudaMemcpyAsync(Lt, ws.Lt + e.plane_off, n, cudaMemcpyDeviceToHost, st));
    if (Lx) DUNK_CUDA(cudaMemcpyAsync(Lx, ws.Lx + e.plane_off, n, cudaMemcpyDeviceToHost, st));
    if (Ly) DUNK_CUDA(cudaMemcpyAsync(Ly, ws.Ly + e.plane_off, n, cudaMemcpyDeviceToHost, st));
    if (Ldet) DUNK_CUDA(cudaMemcpyAsync(Ldet, ws.Ldet + e.plane_off, n, cudaMemcpyDeviceToHost, st));
    if (kcontrast) DUNK_CUDA(cudaMemcpyAsync(kcontrast, ws.kcontrast, 4, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    return DUNK_OK;
}
