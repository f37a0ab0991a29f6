// Internal: AKAZE extraction (stage 1) — level table, workspace layout, kernel launchers.
// Replaces cv::AKAZE::detectAndCompute as configured at feature_extraction/src/lib.rs:64-79.
#pragma once
#include <vector>
#include "ctx.h"

namespace dunk {

constexpr int kMaxLevels = 16;   // omax 4 x nsublevels 4
constexpr int kMaxFedSteps = 32;

// AKAZEFeatures::Allocate_Memory_Evolution (one entry per evolution level)
struct LevelInfo {
    int w, h;             // level size
    int octave, sublevel;
    int sigma_size, border;
    float esigma, ratio;
    int n_tau;
    float tau[kMaxFedSteps];
    size_t plane_off;     // offset (in floats) of this level's plane inside a pyramid array
};

struct LevelTable {
    int n_levels;
    int width, height;
    size_t pyramid_floats;   // sum of w*h over levels
    LevelInfo lv[kMaxLevels];
};

LevelTable make_level_table(int width, int height);

// compact per-level constants handed to detection / description kernels by value
struct LevelDev {
    int w, h, sigma_size, border, octave;
    float ratio, esigma;
    unsigned long long plane_off;
};
struct LevelsDev {
    int n;
    LevelDev lv[kMaxLevels];
};

// Per-frame raw candidate (3x3 maximum above threshold)
struct Cand {
    int x, y, level;
    float resp;
};

// Workspace for a sub-batch of frames (all device memory, carved from one allocation)
struct AkazeWorkspace {
    int frames;            // capacity in frames
    int cand_cap;          // raw candidates per frame
    int kp_cap;            // keypoints per frame (<= cand_cap)
    int total_rows;        // sum of level heights (row buckets per frame)
    // pyramids: [frames][pyramid_floats]
    float *Lt, *Lx, *Ly, *Ldet;
    // per-level temporaries sized for level 0: [frames][w0*h0]
    float *Lsmooth, *Lflow, *Ltmp;
    // contrast
    float* hmax;           // [frames]
    int* hist;             // [frames][300]
    float* kcontrast;      // [frames]
    // detection lists
    Cand* cand_raw;        // [frames][cand_cap] unordered
    Cand* cand;            // [frames][cand_cap] ordered (level, y, x)
    int* cand_count;       // [frames]
    int* row_count;        // [frames][total_rows + 1]
    int* row_start;        // [frames][total_rows + 1] (exclusive scan)
    int* row_fill;         // [frames][total_rows]
    unsigned char* state;  // [frames][cand_cap] 1 = alive
    int* aux;              // [frames][cand_cap * 6] scratch for the suppression passes
    DunkKeyPoint* kps;     // [frames][kp_cap]
    int* kp_count;         // [frames]
    int* kp_level_start;   // unused for now
    uint4* desc64;         // [frames][kp_cap] x 64 B
    float* sort_keys;      // [frames][kp_cap] (max_points path)
};

size_t akaze_workspace_bytes(const LevelTable& lt, int frames, int cand_cap, int kp_cap);
void akaze_carve_workspace(void* base, const LevelTable& lt, int frames, int cand_cap, int kp_cap,
                           AkazeWorkspace* ws);

// stage launchers (async on st); images: [frames] u8, `channels` interleaved, row stride bytes
int akaze_build_scale_space(dunk_ctx* ctx, cudaStream_t st, const LevelTable& lt, const AkazeWorkspace& ws,
                            const unsigned char* images, size_t image_stride_bytes, int row_stride,
                            int channels, int frames);
int akaze_detect(dunk_ctx* ctx, cudaStream_t st, const LevelTable& lt, const AkazeWorkspace& ws, int frames,
                 float dthreshold, int max_points);
int akaze_describe(dunk_ctx* ctx, cudaStream_t st, const LevelTable& lt, const AkazeWorkspace& ws, int frames);

LevelsDev make_levels_dev(const LevelTable& lt);

#define DUNK_KERNEL_CHECK(ctx)                                                             \
    do {                                                                                   \
        (ctx)->launches.fetch_add(1);                                                      \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            dunk::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                     \
            return DUNK_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

}  // namespace dunk
