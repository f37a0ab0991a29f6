// C ABI for stage 2 (matching) and the HBM-resident descriptor database.
#include <algorithm>
#include "match.h"

using namespace dunk;

namespace {

int check_desc_args(const char* fn, const void* q, int nq, const void* t, int64_t nt, int desc_bytes) {
    DUNK_REQUIRE(desc_bytes >= 1 && desc_bytes <= 64, DUNK_ERR_BAD_ARG,
                 "%s: desc_bytes=%d unsupported (1..64; MLDB-486 is 61)", fn, desc_bytes);
    DUNK_REQUIRE(nq >= 0 && nt >= 0, DUNK_ERR_BAD_ARG, "%s: negative row count", fn);
    DUNK_REQUIRE((nq == 0 || q) && (nt == 0 || t), DUNK_ERR_BAD_ARG, "%s: NULL descriptor pointer", fn);
    DUNK_REQUIRE(nt < 0xFFFFFFFFll, DUNK_ERR_BAD_ARG, "%s: train rows exceed 32-bit index space", fn);
    return DUNK_OK;
}

// upload host rows and pad to 64 B on device: returns device pointer inside scratch
int upload_padded(dunk_ctx* ctx, cudaStream_t st, const uint8_t* host, int64_t n, int desc_bytes,
                  uint8_t* raw_dev, uint4* dst64) {
    if (n == 0) return DUNK_OK;
    DUNK_CUDA(cudaMemcpyAsync(raw_dev, host, (size_t)n * desc_bytes, cudaMemcpyHostToDevice, st));
    return launch_pad_rows(ctx, st, raw_dev, n, desc_bytes, dst64);
}

// shared body of knn_match / knn2: leaves merged top-2 in `top2` (device)
struct KnnScratch {
    uint8_t *q_raw, *t_raw;
    uint4 *q64, *t64, *partial, *top2;
    DunkDMatch* matches;
    int* count;
    unsigned long long* keys;
};

}  // namespace

extern "C" {

int dunk_knn2_hamming(dunk_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int64_t nt,
                      int desc_bytes, int32_t* idx, int32_t* dist) {
    DUNK_REQUIRE(ctx, DUNK_ERR_BAD_ARG, "dunk_knn2_hamming: ctx is NULL");
    int rc = check_desc_args("dunk_knn2_hamming", query, nq, train, nt, desc_bytes);
    if (rc) return rc;
    if (nq == 0) return DUNK_OK;
    DUNK_REQUIRE(idx && dist, DUNK_ERR_BAD_ARG, "dunk_knn2_hamming: NULL output");
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const KnnPlan plan = plan_knn2(ctx, nq, (uint32_t)nt);
    size_t need = Carver::need((size_t)nq * desc_bytes) + Carver::need((size_t)nt * desc_bytes) +
                  Carver::need((size_t)nq * 64) + Carver::need((size_t)nt * 64) +
                  Carver::need(knn2_partial_bytes(plan, nq)) + Carver::need((size_t)nq * 16);
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    uint8_t* q_raw = cv.take<uint8_t>((size_t)nq * desc_bytes);
    uint8_t* t_raw = cv.take<uint8_t>((size_t)nt * desc_bytes);
    uint4* q64 = cv.take<uint4>((size_t)nq * 4);
    uint4* t64 = cv.take<uint4>((size_t)nt * 4);
    uint4* partial = cv.take<uint4>((size_t)plan.gx * nq);
    uint4* top2 = cv.take<uint4>(nq);
    if ((rc = upload_padded(ctx, st, query, nq, desc_bytes, q_raw, q64))) return rc;
    if ((rc = upload_padded(ctx, st, train, nt, desc_bytes, t_raw, t64))) return rc;
    if (nt == 0) {
        DUNK_CUDA(cudaMemsetAsync(top2, 0xFF, (size_t)nq * 16, st));
    } else if ((rc = launch_knn2(ctx, st, t64, (uint32_t)nt, q64, nq, 0, partial, top2, plan))) {
        return rc;
    }
    std::vector<DunkTop2> h(nq);
    DUNK_CUDA(cudaMemcpyAsync(h.data(), top2, (size_t)nq * 16, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < nq; ++i) {
        idx[2 * i] = h[i].i1 == 0xFFFFFFFFu ? -1 : (int32_t)h[i].i1;
        dist[2 * i] = h[i].i1 == 0xFFFFFFFFu ? -1 : (int32_t)h[i].d1;
        idx[2 * i + 1] = h[i].i2 == 0xFFFFFFFFu ? -1 : (int32_t)h[i].i2;
        dist[2 * i + 1] = h[i].i2 == 0xFFFFFFFFu ? -1 : (int32_t)h[i].d2;
    }
    return DUNK_OK;
}

int dunk_knn_match_hamming(dunk_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train,
                           int64_t nt, int desc_bytes, int k, float ratio, DunkDMatch* out,
                           int out_cap, int* n_out) {
    DUNK_REQUIRE(ctx && n_out, DUNK_ERR_BAD_ARG, "dunk_knn_match_hamming: NULL ctx / n_out");
    *n_out = 0;
    int rc = check_desc_args("dunk_knn_match_hamming", query, nq, train, nt, desc_bytes);
    if (rc) return rc;
    DUNK_REQUIRE(k >= 1, DUNK_ERR_BAD_ARG, "dunk_knn_match_hamming: k=%d", k);
    if (nq == 0) return DUNK_OK;
    // reference: `i.get(1)?` on a neighbour list shorter than 2 (lib.rs:108) -> StsOutOfRange
    DUNK_REQUIRE(k >= 2 && nt >= 2, DUNK_ERR_OUT_OF_RANGE,
                 "dunk_knn_match_hamming: neighbour list has %lld entries, the ratio test needs 2 "
                 "(k=%d, train rows=%lld)",
                 (long long)std::min<int64_t>(k, nt), k, (long long)nt);
    DUNK_REQUIRE(out && out_cap >= nq, DUNK_ERR_BAD_ARG,
                 "dunk_knn_match_hamming: output capacity %d < query rows %d", out_cap, nq);
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const KnnPlan plan = plan_knn2(ctx, nq, (uint32_t)nt);
    size_t need = Carver::need((size_t)nq * desc_bytes) + Carver::need((size_t)nt * desc_bytes) +
                  Carver::need((size_t)nq * 64) + Carver::need((size_t)nt * 64) +
                  Carver::need(knn2_partial_bytes(plan, nq)) + Carver::need((size_t)nq * 16) +
                  Carver::need((size_t)nq * sizeof(DunkDMatch)) + Carver::need(sizeof(int));
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    uint8_t* q_raw = cv.take<uint8_t>((size_t)nq * desc_bytes);
    uint8_t* t_raw = cv.take<uint8_t>((size_t)nt * desc_bytes);
    uint4* q64 = cv.take<uint4>((size_t)nq * 4);
    uint4* t64 = cv.take<uint4>((size_t)nt * 4);
    uint4* partial = cv.take<uint4>((size_t)plan.gx * nq);
    uint4* top2 = cv.take<uint4>(nq);
    DunkDMatch* matches = cv.take<DunkDMatch>(nq);
    int* count = cv.take<int>(1);
    if ((rc = upload_padded(ctx, st, query, nq, desc_bytes, q_raw, q64))) return rc;
    if ((rc = upload_padded(ctx, st, train, nt, desc_bytes, t_raw, t64))) return rc;
    if ((rc = launch_knn2(ctx, st, t64, (uint32_t)nt, q64, nq, 0, partial, top2, plan))) return rc;
    if ((rc = launch_top2_ratio(ctx, st, top2, nq, ratio, matches, count))) return rc;
    int n = 0;
    DUNK_CUDA(cudaMemcpyAsync(&n, count, sizeof(int), cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    if (n > 0) {
        DUNK_CUDA(cudaMemcpyAsync(out, matches, (size_t)n * sizeof(DunkDMatch), cudaMemcpyDeviceToHost, st));
        DUNK_CUDA(cudaStreamSynchronize(st));
    }
    *n_out = n;
    return DUNK_OK;
}

int dunk_match_crosscheck_hamming(dunk_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train,
                                  int64_t nt, int desc_bytes, DunkDMatch* out, int out_cap,
                                  int* n_out) {
    DUNK_REQUIRE(ctx && n_out, DUNK_ERR_BAD_ARG, "dunk_match_crosscheck_hamming: NULL ctx / n_out");
    *n_out = 0;
    int rc = check_desc_args("dunk_match_crosscheck_hamming", query, nq, train, nt, desc_bytes);
    if (rc) return rc;
    if (nq == 0 || nt == 0) return DUNK_OK;
    DUNK_REQUIRE(nt <= 0x7FFFFFFF, DUNK_ERR_BAD_ARG, "dunk_match_crosscheck_hamming: too many train rows");
    DUNK_REQUIRE(out && out_cap >= nq, DUNK_ERR_BAD_ARG,
                 "dunk_match_crosscheck_hamming: output capacity %d < query rows %d", out_cap, nq);
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    // two 1-NN passes: query -> train and (roles swapped) train -> query
    const KnnPlan plan_q = plan_knn2(ctx, nq, (uint32_t)nt);
    const KnnPlan plan_t = plan_knn2(ctx, (int)nt, (uint32_t)nq);
    const size_t partial_bytes = std::max(knn2_partial_bytes(plan_q, nq), knn2_partial_bytes(plan_t, (int)nt));
    size_t need = Carver::need((size_t)nq * desc_bytes) + Carver::need((size_t)nt * desc_bytes) +
                  Carver::need((size_t)nq * 64) + Carver::need((size_t)nt * 64) +
                  Carver::need(partial_bytes) + Carver::need((size_t)nt * 16) +
                  Carver::need((size_t)nq * 16) + Carver::need((size_t)nq * sizeof(DunkDMatch)) +
                  Carver::need(sizeof(int));
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    uint8_t* q_raw = cv.take<uint8_t>((size_t)nq * desc_bytes);
    uint8_t* t_raw = cv.take<uint8_t>((size_t)nt * desc_bytes);
    uint4* q64 = cv.take<uint4>((size_t)nq * 4);
    uint4* t64 = cv.take<uint4>((size_t)nt * 4);
    uint4* partial = cv.take<uint4>(partial_bytes / 16);
    uint4* t2q = cv.take<uint4>(nt);
    uint4* q2t = cv.take<uint4>(nq);
    DunkDMatch* matches = cv.take<DunkDMatch>(nq);
    int* count = cv.take<int>(1);
    if ((rc = upload_padded(ctx, st, query, nq, desc_bytes, q_raw, q64))) return rc;
    if ((rc = upload_padded(ctx, st, train, nt, desc_bytes, t_raw, t64))) return rc;
    if ((rc = launch_knn2(ctx, st, t64, (uint32_t)nt, q64, nq, 0, partial, q2t, plan_q))) return rc;
    if ((rc = launch_knn2(ctx, st, q64, (uint32_t)nq, t64, (int)nt, 0, partial, t2q, plan_t))) return rc;
    if ((rc = launch_crosscheck(ctx, st, q2t, t2q, nq, matches, count))) return rc;
    int n = 0;
    DUNK_CUDA(cudaMemcpyAsync(&n, count, sizeof(int), cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    if (n > 0) {
        DUNK_CUDA(cudaMemcpyAsync(out, matches, (size_t)n * sizeof(DunkDMatch), cudaMemcpyDeviceToHost, st));
        DUNK_CUDA(cudaStreamSynchronize(st));
    }
    *n_out = n;
    return DUNK_OK;
}

/* ---- HBM-resident DB ----------------------------------------------------------------- */

int dunk_db_create(dunk_ctx* ctx, int64_t capacity_rows, int desc_bytes, dunk_db** out) {
    DUNK_REQUIRE(ctx && out, DUNK_ERR_BAD_ARG, "dunk_db_create: NULL argument");
    *out = nullptr;
    DUNK_REQUIRE(capacity_rows > 0 && capacity_rows < 0xFFFFFFFFll, DUNK_ERR_BAD_ARG,
                 "dunk_db_create: capacity %lld out of range", (long long)capacity_rows);
    DUNK_REQUIRE(desc_bytes >= 1 && desc_bytes <= 64, DUNK_ERR_BAD_ARG, "dunk_db_create: desc_bytes=%d",
                 desc_bytes);
    DUNK_CUDA(cudaSetDevice(ctx->device));
    dunk_db* db = new dunk_db();
    db->ctx = ctx;
    db->desc_bytes = desc_bytes;
    db->capacity = capacity_rows;
    cudaError_t e = cudaMalloc(&db->desc64, (size_t)capacity_rows * 64);
    if (e == cudaSuccess) e = cudaMalloc(&db->kps, (size_t)capacity_rows * sizeof(DunkKeyPoint));
    if (e == cudaSuccess) e = cudaMalloc(&db->image_id, (size_t)capacity_rows * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("dunk_db_create: allocating %lld rows failed: %s", (long long)capacity_rows,
                  cudaGetErrorString(e));
        dunk_db_destroy(db);
        return DUNK_ERR_NO_MEM;
    }
    *out = db;
    return DUNK_OK;
}

void dunk_db_destroy(dunk_db* db) {
    if (!db) return;
    cudaSetDevice(db->ctx->device);
    if (db->desc64) cudaFree(db->desc64);
    if (db->kps) cudaFree(db->kps);
    if (db->image_id) cudaFree(db->image_id);
    if (db->row_id) cudaFree(db->row_id);
    if (db->image_lod_dev) cudaFree(db->image_lod_dev);
    delete db;
}

int64_t dunk_db_size(dunk_db* db) { return db ? db->size : 0; }
int dunk_db_desc_bytes(dunk_db* db) { return db ? db->desc_bytes : 0; }

int dunk_db_clear(dunk_db* db) {
    DUNK_REQUIRE(db, DUNK_ERR_BAD_ARG, "dunk_db_clear: db is NULL");
    std::lock_guard<std::mutex> lk(db->mu);
    db->size = 0;
    db->images.clear();
    db->image_lod_dirty = true;
    return DUNK_OK;
}

int dunk_db_append(dunk_db* db, const uint8_t* desc, const DunkKeyPoint* kps, const int32_t* image_ids,
                   int64_t n) {
    DUNK_REQUIRE(db, DUNK_ERR_BAD_ARG, "dunk_db_append: db is NULL");
    DUNK_REQUIRE(n >= 0 && (n == 0 || desc), DUNK_ERR_BAD_ARG, "dunk_db_append: bad rows");
    if (n == 0) return DUNK_OK;
    std::lock_guard<std::mutex> lk(db->mu);
    DUNK_REQUIRE(db->size + n <= db->capacity, DUNK_ERR_NO_MEM,
                 "dunk_db_append: %lld + %lld rows exceed capacity %lld", (long long)db->size,
                 (long long)n, (long long)db->capacity);
    dunk_ctx* ctx = db->ctx;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    // stage raw rows through scratch in chunks (bounded scratch for multi-GB uploads)
    const int64_t chunk = 4 << 20;
    void* scratch = ctx->dev_scratch(g.s, (size_t)std::min(chunk, n) * db->desc_bytes);
    if (!scratch) return DUNK_ERR_NO_MEM;
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t m = std::min(chunk, n - off);
        DUNK_CUDA(cudaMemcpyAsync(scratch, desc + off * db->desc_bytes, (size_t)m * db->desc_bytes,
                                  cudaMemcpyHostToDevice, st));
        int rc = launch_pad_rows(ctx, st, (const uint8_t*)scratch, m, db->desc_bytes,
                                 db->desc64 + (db->size + off) * 4);
        if (rc) return rc;
        DUNK_CUDA(cudaStreamSynchronize(st));
    }
    if (kps)
        DUNK_CUDA(cudaMemcpyAsync(db->kps + db->size, kps, (size_t)n * sizeof(DunkKeyPoint),
                                  cudaMemcpyHostToDevice, st));
    else
        DUNK_CUDA(cudaMemsetAsync(db->kps + db->size, 0, (size_t)n * sizeof(DunkKeyPoint), st));
    if (image_ids)
        DUNK_CUDA(cudaMemcpyAsync(db->image_id + db->size, image_ids, (size_t)n * 4,
                                  cudaMemcpyHostToDevice, st));
    else
        DUNK_CUDA(cudaMemsetAsync(db->image_id + db->size, 0, (size_t)n * 4, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    db->size += n;
    return DUNK_OK;
}

int dunk_db_append_random_at(dunk_db* db, int64_t n, uint64_t seed, uint64_t global_row_offset);
int dunk_db_append_random(dunk_db* db, int64_t n, uint64_t seed) {
    return dunk_db_append_random_at(db, n, seed, db ? (uint64_t)db->size : 0);
}

int dunk_db_append_random_at(dunk_db* db, int64_t n, uint64_t seed, uint64_t global_row_offset) {
    DUNK_REQUIRE(db && n >= 0, DUNK_ERR_BAD_ARG, "dunk_db_append_random: bad argument");
    if (n == 0) return DUNK_OK;
    std::lock_guard<std::mutex> lk(db->mu);
    DUNK_REQUIRE(db->size + n <= db->capacity, DUNK_ERR_NO_MEM,
                 "dunk_db_append_random: exceeds capacity %lld", (long long)db->capacity);
    dunk_ctx* ctx = db->ctx;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    int rc = launch_fill_random_rows(ctx, st, db->desc64 + db->size * 4, n, seed, global_row_offset);
    if (rc) return rc;
    DUNK_CUDA(cudaMemsetAsync(db->kps + db->size, 0, (size_t)n * sizeof(DunkKeyPoint), st));
    DUNK_CUDA(cudaMemsetAsync(db->image_id + db->size, 0, (size_t)n * 4, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    db->size += n;
    return DUNK_OK;
}

int dunk_db_read(dunk_db* db, int64_t first, int64_t n, uint8_t* desc, DunkKeyPoint* kps,
                 int32_t* image_ids) {
    DUNK_REQUIRE(db, DUNK_ERR_BAD_ARG, "dunk_db_read: db is NULL");
    DUNK_REQUIRE(first >= 0 && n >= 0 && first + n <= db->size, DUNK_ERR_OUT_OF_RANGE,
                 "dunk_db_read: rows [%lld, %lld) outside 0..%lld", (long long)first,
                 (long long)(first + n), (long long)db->size);
    if (n == 0) return DUNK_OK;
    dunk_ctx* ctx = db->ctx;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    if (desc) {
        void* scratch = ctx->dev_scratch(g.s, (size_t)n * db->desc_bytes);
        if (!scratch) return DUNK_ERR_NO_MEM;
        int rc = launch_unpad_rows(ctx, st, db->desc64 + first * 4, n, db->desc_bytes, (uint8_t*)scratch);
        if (rc) return rc;
        DUNK_CUDA(cudaMemcpyAsync(desc, scratch, (size_t)n * db->desc_bytes, cudaMemcpyDeviceToHost, st));
    }
    if (kps)
        DUNK_CUDA(cudaMemcpyAsync(kps, db->kps + first, (size_t)n * sizeof(DunkKeyPoint),
                                  cudaMemcpyDeviceToHost, st));
    if (image_ids)
        DUNK_CUDA(cudaMemcpyAsync(image_ids, db->image_id + first, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    return DUNK_OK;
}

int dunk_db_knn2_dev(dunk_db* db, int slot, const void* query64_dev, int nq, uint32_t index_base,
                     void* top2_dev) {
    DUNK_REQUIRE(db && query64_dev && top2_dev && nq >= 0, DUNK_ERR_BAD_ARG, "dunk_db_knn2_dev: bad argument");
    dunk_ctx* ctx = db->ctx;
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_db_knn2_dev: bad slot");
    if (nq == 0) return DUNK_OK;
    DUNK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[slot].stream;
    if (db->size == 0) {
        DUNK_CUDA(cudaMemsetAsync(top2_dev, 0xFF, (size_t)nq * 16, st));
        return DUNK_OK;
    }
    const KnnPlan plan = plan_knn2(ctx, nq, (uint32_t)db->size);
    void* scratch = ctx->dev_scratch(slot, knn2_partial_bytes(plan, nq));
    if (!scratch) return DUNK_ERR_NO_MEM;
    return launch_knn2(ctx, st, db->desc64, (uint32_t)db->size, (const uint4*)query64_dev, nq,
                       index_base, (uint4*)scratch, (uint4*)top2_dev, plan);
}

int dunk_top2_merge_dev(dunk_ctx* ctx, int slot, const void* parts_dev, int n_parts, int nq,
                        void* merged_dev) {
    DUNK_REQUIRE(ctx && parts_dev && merged_dev && n_parts >= 1 && nq >= 0, DUNK_ERR_BAD_ARG,
                 "dunk_top2_merge_dev: bad argument");
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_top2_merge_dev: bad slot");
    DUNK_CUDA(cudaSetDevice(ctx->device));
    return launch_top2_merge(ctx, ctx->slots[slot].stream, (const uint4*)parts_dev, n_parts, nq,
                             (uint4*)merged_dev);
}

int dunk_top2_ratio_dev(dunk_ctx* ctx, int slot, const void* merged_dev, int nq, float ratio,
                        void* matches_dev, void* count_dev) {
    DUNK_REQUIRE(ctx && merged_dev && matches_dev && count_dev && nq >= 0, DUNK_ERR_BAD_ARG,
                 "dunk_top2_ratio_dev: bad argument");
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_top2_ratio_dev: bad slot");
    DUNK_CUDA(cudaSetDevice(ctx->device));
    return launch_top2_ratio(ctx, ctx->slots[slot].stream, (const uint4*)merged_dev, nq, ratio,
                             (DunkDMatch*)matches_dev, (int*)count_dev);
}

int dunk_pad_desc_dev(dunk_ctx* ctx, int slot, const void* src_dev, int64_t n, int desc_bytes,
                      void* dst64_dev) {
    DUNK_REQUIRE(ctx && src_dev && dst64_dev && n >= 0, DUNK_ERR_BAD_ARG, "dunk_pad_desc_dev: bad argument");
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_pad_desc_dev: bad slot");
    DUNK_CUDA(cudaSetDevice(ctx->device));
    return launch_pad_rows(ctx, ctx->slots[slot].stream, (const uint8_t*)src_dev, n, desc_bytes,
                           (uint4*)dst64_dev);
}

int dunk_db_knn2(dunk_db* db, const uint8_t* query, int nq, uint32_t index_base, DunkTop2* out) {
    DUNK_REQUIRE(db && nq >= 0 && (nq == 0 || (query && out)), DUNK_ERR_BAD_ARG, "dunk_db_knn2: bad argument");
    if (nq == 0) return DUNK_OK;
    dunk_ctx* ctx = db->ctx;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const KnnPlan plan = plan_knn2(ctx, nq, (uint32_t)std::max<int64_t>(db->size, 1));
    size_t need = Carver::need((size_t)nq * db->desc_bytes) + Carver::need((size_t)nq * 64) +
                  Carver::need(knn2_partial_bytes(plan, nq)) + Carver::need((size_t)nq * 16);
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    uint8_t* q_raw = cv.take<uint8_t>((size_t)nq * db->desc_bytes);
    uint4* q64 = cv.take<uint4>((size_t)nq * 4);
    uint4* partial = cv.take<uint4>((size_t)plan.gx * nq);
    uint4* top2 = cv.take<uint4>(nq);
    int rc;
    if ((rc = upload_padded(ctx, st, query, nq, db->desc_bytes, q_raw, q64))) return rc;
    if (db->size == 0) {
        DUNK_CUDA(cudaMemsetAsync(top2, 0xFF, (size_t)nq * 16, st));
    } else if ((rc = launch_knn2(ctx, st, db->desc64, (uint32_t)db->size, q64, nq, index_base, partial,
                                 top2, plan))) {
        return rc;
    }
    DUNK_CUDA(cudaMemcpyAsync(out, top2, (size_t)nq * 16, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    return DUNK_OK;
}

int dunk_db_match(dunk_db* db, const uint8_t* query, int nq, float ratio, DunkDMatch* out, int out_cap,
                  int* n_out) {
    DUNK_REQUIRE(db && n_out, DUNK_ERR_BAD_ARG, "dunk_db_match: NULL argument");
    *n_out = 0;
    DUNK_REQUIRE(nq >= 0 && (nq == 0 || query), DUNK_ERR_BAD_ARG, "dunk_db_match: bad query");
    if (nq == 0) return DUNK_OK;
    DUNK_REQUIRE(db->size >= 2, DUNK_ERR_OUT_OF_RANGE,
                 "dunk_db_match: database holds %lld rows, the ratio test needs 2", (long long)db->size);
    DUNK_REQUIRE(out && out_cap >= nq, DUNK_ERR_BAD_ARG, "dunk_db_match: output capacity %d < %d", out_cap, nq);
    dunk_ctx* ctx = db->ctx;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const KnnPlan plan = plan_knn2(ctx, nq, (uint32_t)db->size);
    size_t need = Carver::need((size_t)nq * db->desc_bytes) + Carver::need((size_t)nq * 64) +
                  Carver::need(knn2_partial_bytes(plan, nq)) + Carver::need((size_t)nq * 16) +
                  Carver::need((size_t)nq * sizeof(DunkDMatch)) + Carver::need(sizeof(int));
    void* scratch = ctx->dev_scratch(g.s, need);
    if (!scratch) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    uint8_t* q_raw = cv.take<uint8_t>((size_t)nq * db->desc_bytes);
    uint4* q64 = cv.take<uint4>((size_t)nq * 4);
    uint4* partial = cv.take<uint4>((size_t)plan.gx * nq);
    uint4* top2 = cv.take<uint4>(nq);
    DunkDMatch* matches = cv.take<DunkDMatch>(nq);
    int* count = cv.take<int>(1);
    int rc;
    if ((rc = upload_padded(ctx, st, query, nq, db->desc_bytes, q_raw, q64))) return rc;
    if ((rc = launch_knn2(ctx, st, db->desc64, (uint32_t)db->size, q64, nq, 0, partial, top2, plan))) return rc;
    if ((rc = launch_top2_ratio(ctx, st, top2, nq, ratio, matches, count))) return rc;
    int n = 0;
    DUNK_CUDA(cudaMemcpyAsync(&n, count, sizeof(int), cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    if (n > 0) {
        DUNK_CUDA(cudaMemcpyAsync(out, matches, (size_t)n * sizeof(DunkDMatch), cudaMemcpyDeviceToHost, st));
        DUNK_CUDA(cudaStreamSynchronize(st));
    }
    *n_out = n;
    return DUNK_OK;
}

}  // extern "C"
