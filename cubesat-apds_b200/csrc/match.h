// Internal declarations for the Hamming matcher (not part of the ABI).
#pragma once
#include <vector>
#include "ctx.h"

// One HBM-resident shard of the reference descriptor database: SoA columns of the reference's
// `keypoint` table (feature_database/src/models.rs:30-41, schema.rs:27-40)
struct dunk_db {
    dunk_ctx* ctx = nullptr;
    int desc_bytes = 61;
    int64_t capacity = 0;
    int64_t size = 0;
    uint4* desc64 = nullptr;        // capacity x 64 B descriptor rows
    DunkKeyPoint* kps = nullptr;    // capacity x 28 B (x, y, size, angle, response, octave, class_id)
    int32_t* image_id = nullptr;    // capacity
    int32_t* row_id = nullptr;      // capacity; only in DBs produced by dunk_db_select (the `id` column,
                                    // = 1 + row index in the source DB); a base DB's id is 1 + row index
    // `ref_image` table (feature_database/src/models.rs:5-15): tiny, host-resident; ids are 1-based
    // like a Postgres SERIAL.  image_lod_dev mirrors level_of_detail for the select kernel.
    std::vector<DunkImage> images;
    int32_t* image_lod_dev = nullptr;
    int64_t image_lod_cap = 0;
    bool image_lod_dirty = true;
    std::mutex mu;
};

namespace dunk {

struct KnnPlan {
    int threads;        // CTA size (warps chosen to minimise query padding)
    int gx, gy;         // DB slabs x query groups
    int tiles_per_cta;  // 128-row tiles per slab
    size_t smem;
};

KnnPlan plan_knn2(dunk_ctx* ctx, int nq, uint32_t nt);
size_t knn2_partial_bound(dunk_ctx* ctx, size_t nq_max);
// scratch needed for slab partials
inline size_t knn2_partial_bytes(const KnnPlan& p, int nq) { return (size_t)p.gx * nq * 16; }

// all launches are asynchronous on `st`
int launch_pad_rows(dunk_ctx* ctx, cudaStream_t st, const uint8_t* src, int64_t n, int desc_bytes,
                    uint4* dst64);
int launch_unpad_rows(dunk_ctx* ctx, cudaStream_t st, const uint4* src64, int64_t n, int desc_bytes,
                      uint8_t* dst);
// seg_counts / seg_len (sharded pipeline only): the queries are segments of seg_len rows, segment r holds
// seg_counts[2 * r] real rows (device array); groups wholly inside a segment's padding are skipped
int launch_knn2(dunk_ctx* ctx, cudaStream_t st, const uint4* db64, uint32_t nt, const uint4* q64,
                int nq, uint32_t index_base, uint4* partial, uint4* top2_out, const KnnPlan& plan,
                const int* seg_counts = nullptr, int seg_len = 0);
int launch_top2_merge(dunk_ctx* ctx, cudaStream_t st, const uint4* parts, int n_parts, int nq,
                      uint4* out);
int launch_top2_merge_strided(dunk_ctx* ctx, cudaStream_t st, const uint4* parts, int n_parts, long long part_stride,
                              int nq, uint4* out);
int launch_top2_ratio(dunk_ctx* ctx, cudaStream_t st, const uint4* top2, int nq, float ratio,
                      DunkDMatch* out, int* count);
// cross-check (BFMatcher crossCheck=true): mutual nearest neighbours from both 1-NN passes
int launch_crosscheck(dunk_ctx* ctx, cudaStream_t st, const uint4* q2t_top2, const uint4* t2q_top2,
                      int nq, DunkDMatch* out, int* count);
int launch_fill_random_rows(dunk_ctx* ctx, cudaStream_t st, uint4* dst64, int64_t n, uint64_t seed,
                            uint64_t row_offset);

}  // namespace dunk
