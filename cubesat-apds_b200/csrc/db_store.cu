// feature_database on the device: the `ref_image` table, the three keyed reads of
// KeypointDatabase (feature_database/src/keypointdb.rs:38-90: filter, ORDER BY response DESC,
// LIMIT 2^18-1) as predicate + stable radix sort + gather over the HBM SoA columns, and a flat
// binary dump / load of a shard.  Postgres / diesel themselves are out of scope.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include "match.h"

using namespace dunk;

namespace {

struct FilterDev {
    int image_id, lod, use_box;
    float x_lo, x_hi, y_lo, y_hi;
};

// flags[i] = row i passes the WHERE clause; keys[i] = response; idx[i] = i
__global__ void __launch_bounds__(256)
k_select_flags(const DunkKeyPoint* __restrict__ kps, const int32_t* __restrict__ image_id, const int32_t* __restrict__ image_lod,
               int n_images, long long n, FilterDev f, unsigned char* __restrict__ flags, float* __restrict__ keys,
               uint32_t* __restrict__ idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DunkKeyPoint k = kps[i];
    const int im = image_id[i];
    bool ok = true;
    if (f.image_id >= 0) ok = ok && im == f.image_id;
    if (f.lod >= 0) ok = ok && im >= 1 && im <= n_images && image_lod[im - 1] == f.lod;   // inner join on ref_image
    if (f.use_box) ok = ok && k.x >= f.x_lo && k.x <= f.x_hi && k.y >= f.y_lo && k.y <= f.y_hi;
    flags[i] = ok;
    keys[i] = k.response;
    idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
k_gather_rows(const uint4* __restrict__ desc64, const DunkKeyPoint* __restrict__ kps, const int32_t* __restrict__ image_id,
              const int32_t* __restrict__ row_id, const uint32_t* __restrict__ order, long long n,
              uint4* __restrict__ o_desc, DunkKeyPoint* __restrict__ o_kps, int32_t* __restrict__ o_image, int32_t* __restrict__ o_id) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // 4 threads per row (one uint4 each)
    const long long r = t >> 2;
    if (r >= n) return;
    const uint32_t src = order[r];
    const int q = (int)(t & 3);
    o_desc[r * 4 + q] = desc64[(size_t)src * 4 + q];
    if (q == 0) {
        o_kps[r] = kps[src];
        o_image[r] = image_id[src];
        o_id[r] = row_id ? row_id[src] : (int32_t)(src + 1);
    }
}

int sync_image_lod(dunk_db* db, cudaStream_t st) {
    const int64_t n = (int64_t)db->images.size();
    if (!db->image_lod_dirty && db->image_lod_dev) return DUNK_OK;
    if (n > db->image_lod_cap) {
        if (db->image_lod_dev) cudaFree(db->image_lod_dev);
        db->image_lod_cap = std::max<int64_t>(n * 2, 1024);
        DUNK_CUDA(cudaMalloc(&db->image_lod_dev, (size_t)db->image_lod_cap * 4));
    } else if (!db->image_lod_dev) {
        db->image_lod_cap = 1024;
        DUNK_CUDA(cudaMalloc(&db->image_lod_dev, (size_t)db->image_lod_cap * 4));
    }
    std::vector<int32_t> lod((size_t)n);
    for (int64_t i = 0; i < n; ++i) lod[i] = db->images[i].level_of_detail;
    if (n) DUNK_CUDA(cudaMemcpyAsync(db->image_lod_dev, lod.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    db->image_lod_dirty = false;
    return DUNK_OK;
}

struct FileHeader {
    char magic[8];            // "DUNKDB01"
    int64_t rows;
    int32_t desc_bytes, n_images;
    int32_t has_row_id, reserved;
};

}  // namespace

extern "C" {

int dunk_db_create_image(dunk_db* db, int32_t x_start, int32_t y_start, int32_t x_end, int32_t y_end,
                         int32_t level_of_detail, int32_t* id_out) {
    DUNK_REQUIRE(db, DUNK_ERR_BAD_ARG, "dunk_db_create_image: db is NULL");
    std::lock_guard<std::mutex> lk(db->mu);
    DunkImage im{(int32_t)db->images.size() + 1, x_start, y_start, x_end, y_end, level_of_detail};
    db->images.push_back(im);
    db->image_lod_dirty = true;
    if (id_out) *id_out = im.id;
    return DUNK_OK;
}

int dunk_db_image_count(dunk_db* db) { return db ? (int)db->images.size() : 0; }

int dunk_db_read_image(dunk_db* db, int32_t id, DunkImage* out) {
    DUNK_REQUIRE(db && out, DUNK_ERR_BAD_ARG, "dunk_db_read_image: NULL argument");
    std::lock_guard<std::mutex> lk(db->mu);
    DUNK_REQUIRE(id >= 1 && id <= (int32_t)db->images.size(), DUNK_ERR_OUT_OF_RANGE,
                 "dunk_db_read_image: no image with id %d (diesel NotFound)", id);
    *out = db->images[id - 1];
    return DUNK_OK;
}

int dunk_db_find_images(dunk_db* db, int use_box, int32_t x_start, int32_t y_start, int32_t x_end, int32_t y_end,
                        int32_t level_of_detail, int32_t* ids, int cap, int* n_out) {
    DUNK_REQUIRE(db && n_out, DUNK_ERR_BAD_ARG, "dunk_db_find_images: NULL argument");
    std::lock_guard<std::mutex> lk(db->mu);
    int n = 0;
    for (const DunkImage& im : db->images) {
        if (im.level_of_detail != level_of_detail) continue;
        // imagedb.rs:47-51: x_end >= x_start AND x_start <= x_end AND y likewise
        if (use_box && !(im.x_end >= x_start && im.x_start <= x_end && im.y_end >= y_start && im.y_start <= y_end)) continue;
        if (ids && n < cap) ids[n] = im.id;
        ++n;
    }
    *n_out = n;
    DUNK_REQUIRE(!ids || n <= cap, DUNK_ERR_NO_MEM, "dunk_db_find_images: %d ids, capacity %d", n, cap);
    return DUNK_OK;
}

int dunk_db_select(dunk_db* db, const DunkRowFilter* filter, int64_t limit, dunk_db** out) {
    DUNK_REQUIRE(db && filter && out, DUNK_ERR_BAD_ARG, "dunk_db_select: NULL argument");
    *out = nullptr;
    DUNK_REQUIRE(limit >= 0, DUNK_ERR_BAD_ARG, "dunk_db_select: negative limit");
    dunk_ctx* ctx = db->ctx;
    std::lock_guard<std::mutex> lk(db->mu);
    const int64_t n = db->size;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    int rc = sync_image_lod(db, st);
    if (rc) return rc;
    FilterDev f{filter->image_id, filter->level_of_detail, filter->use_box, floorf(filter->x_start), ceilf(filter->x_end),
                floorf(filter->y_start), ceilf(filter->y_end)};
    int64_t n_sel = 0;
    uint32_t* d_order = nullptr;
    // the CUB select / radix-sort entry points used below take 32-bit item counts
    DUNK_REQUIRE(n <= 0x7fffffffll, DUNK_ERR_BAD_ARG, "dunk_db_select: %lld rows in one shard (at most 2^31 - 1 per keyed read)", (long long)n);
    if (n > 0) {
        size_t tmp_select = 0, tmp_sort = 0;
        cub::DeviceSelect::Flagged(nullptr, tmp_select, (float*)nullptr, (unsigned char*)nullptr, (float*)nullptr, (int*)nullptr, (int)n, st);
        cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_sort, (float*)nullptr, (float*)nullptr, (uint32_t*)nullptr,
                                                  (uint32_t*)nullptr, (int)n, 0, 32, st);
        const size_t tmp = std::max(tmp_select, tmp_sort);
        const size_t need = Carver::need((size_t)n) + 4 * Carver::need((size_t)n * 4) + 2 * Carver::need((size_t)n * 4) +
                            Carver::need(16) + Carver::need(tmp);
        void* scratch = ctx->dev_scratch(g.s, need);
        if (!scratch) return DUNK_ERR_NO_MEM;
        Carver cv(scratch);
        unsigned char* d_flags = cv.take<unsigned char>(n);
        float* d_keys = cv.take<float>(n);
        uint32_t* d_idx = cv.take<uint32_t>(n);
        float* d_keys_sel = cv.take<float>(n);
        uint32_t* d_idx_sel = cv.take<uint32_t>(n);
        float* d_keys_sorted = cv.take<float>(n);
        uint32_t* d_idx_sorted = cv.take<uint32_t>(n);
        int* d_count = cv.take<int>(4);
        void* d_tmp = cv.take<char>(tmp);
        k_select_flags<<<div_up(n, 256), 256, 0, st>>>(db->kps, db->image_id, db->image_lod_dev, (int)db->images.size(), n, f, d_flags,
                                                       d_keys, d_idx);
        ctx->launches.fetch_add(1);
        DUNK_CUDA(cudaGetLastError());
        size_t t1 = tmp;
        DUNK_CUDA(cub::DeviceSelect::Flagged(d_tmp, t1, d_keys, d_flags, d_keys_sel, d_count, (int)n, st));
        t1 = tmp;
        DUNK_CUDA(cub::DeviceSelect::Flagged(d_tmp, t1, d_idx, d_flags, d_idx_sel, d_count, (int)n, st));
        int h_count = 0;
        DUNK_CUDA(cudaMemcpyAsync(&h_count, d_count, 4, cudaMemcpyDeviceToHost, st));
        DUNK_CUDA(cudaStreamSynchronize(st));
        if (h_count > 0) {
            t1 = tmp;
            // stable: equal responses keep ascending row order (Postgres leaves ties unspecified)
            DUNK_CUDA(cub::DeviceRadixSort::SortPairsDescending(d_tmp, t1, d_keys_sel, d_keys_sorted, d_idx_sel, d_idx_sorted, h_count,
                                                                0, 32, st));
        }
        n_sel = std::min<int64_t>(h_count, limit);
        d_order = d_idx_sorted;
    }
    dunk_db* sub = nullptr;
    rc = dunk_db_create(ctx, std::max<int64_t>(n_sel, 1), db->desc_bytes, &sub);
    if (rc) return rc;
    cudaError_t e = cudaMalloc(&sub->row_id, (size_t)std::max<int64_t>(n_sel, 1) * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        dunk_db_destroy(sub);
        set_error("dunk_db_select: row id column allocation failed");
        return DUNK_ERR_NO_MEM;
    }
    sub->images = db->images;
    if (n_sel > 0) {
        k_gather_rows<<<div_up(n_sel * 4, 256), 256, 0, st>>>(db->desc64, db->kps, db->image_id, db->row_id, d_order, n_sel, sub->desc64,
                                                            sub->kps, sub->image_id, sub->row_id);
        ctx->launches.fetch_add(1);
        cudaError_t ge = cudaGetLastError();
        if (ge == cudaSuccess) ge = cudaStreamSynchronize(st);
        if (ge != cudaSuccess) {
            dunk_db_destroy(sub);
            set_error("dunk_db_select: gather failed: %s", cudaGetErrorString(ge));
            return DUNK_ERR_CUDA;
        }
    }
    sub->size = n_sel;
    *out = sub;
    return DUNK_OK;
}

int dunk_db_read_ids(dunk_db* db, int64_t first, int64_t n, int32_t* ids) {
    DUNK_REQUIRE(db && (ids || n == 0), DUNK_ERR_BAD_ARG, "dunk_db_read_ids: NULL argument");
    DUNK_REQUIRE(first >= 0 && n >= 0 && first + n <= db->size, DUNK_ERR_OUT_OF_RANGE, "dunk_db_read_ids: rows [%lld, %lld) outside 0..%lld",
                 (long long)first, (long long)(first + n), (long long)db->size);
    if (n == 0) return DUNK_OK;
    if (!db->row_id) {
        for (int64_t i = 0; i < n; ++i) ids[i] = (int32_t)(first + i + 1);
        return DUNK_OK;
    }
    SlotGuard g(db->ctx);
    DUNK_CUDA(cudaMemcpyAsync(ids, db->row_id + first, (size_t)n * 4, cudaMemcpyDeviceToHost, g.stream()));
    DUNK_CUDA(cudaStreamSynchronize(g.stream()));
    return DUNK_OK;
}

int dunk_db_save(dunk_db* db, const char* path) {
    DUNK_REQUIRE(db && path, DUNK_ERR_BAD_ARG, "dunk_db_save: NULL argument");
    std::lock_guard<std::mutex> lk(db->mu);
    FILE* fp = fopen(path, "wb");
    DUNK_REQUIRE(fp, DUNK_ERR_BAD_ARG, "dunk_db_save: cannot open %s for writing", path);
    FileHeader h{};
    memcpy(h.magic, "DUNKDB01", 8);
    h.rows = db->size;
    h.desc_bytes = db->desc_bytes;
    h.n_images = (int32_t)db->images.size();
    h.has_row_id = db->row_id ? 1 : 0;
    bool ok = fwrite(&h, sizeof h, 1, fp) == 1;
    if (ok && h.n_images) ok = fwrite(db->images.data(), sizeof(DunkImage), h.n_images, fp) == (size_t)h.n_images;
    dunk_ctx* ctx = db->ctx;
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const size_t chunk = (size_t)64 << 20;
    void* pin = ctx->pin_scratch(g.s, chunk);
    if (!pin) { fclose(fp); return DUNK_ERR_NO_MEM; }
    auto dump = [&](const void* dev, size_t bytes) -> int {
        for (size_t off = 0; off < bytes && ok; off += chunk) {
            const size_t m = std::min(chunk, bytes - off);
            DUNK_CUDA(cudaMemcpyAsync(pin, (const char*)dev + off, m, cudaMemcpyDeviceToHost, st));
            DUNK_CUDA(cudaStreamSynchronize(st));
            ok = fwrite(pin, 1, m, fp) == m;
        }
        return DUNK_OK;
    };
    int rc = DUNK_OK;
    if (ok && !(rc = dump(db->desc64, (size_t)db->size * 64)) && !(rc = dump(db->kps, (size_t)db->size * sizeof(DunkKeyPoint))) &&
        !(rc = dump(db->image_id, (size_t)db->size * 4)) && db->row_id)
        rc = dump(db->row_id, (size_t)db->size * 4);
    ok = (fclose(fp) == 0) && ok;
    if (rc) return rc;
    DUNK_REQUIRE(ok, DUNK_ERR_BAD_ARG, "dunk_db_save: write to %s failed", path);
    return DUNK_OK;
}

int dunk_db_load(dunk_ctx* ctx, const char* path, int64_t min_capacity_rows, dunk_db** out) {
    DUNK_REQUIRE(ctx && path && out, DUNK_ERR_BAD_ARG, "dunk_db_load: NULL argument");
    *out = nullptr;
    FILE* fp = fopen(path, "rb");
    DUNK_REQUIRE(fp, DUNK_ERR_BAD_ARG, "dunk_db_load: cannot open %s", path);
    FileHeader h{};
    if (fread(&h, sizeof h, 1, fp) != 1 || memcmp(h.magic, "DUNKDB01", 8) != 0 || h.rows < 0 || h.desc_bytes < 1 || h.desc_bytes > 64 ||
        h.n_images < 0 || h.n_images > (1 << 26) || h.reserved != 0) {       // `reserved` doubles as the format version (0)
        fclose(fp);
        set_error("dunk_db_load: %s is not a dunk_b200 database dump", path);
        return DUNK_ERR_BAD_ARG;
    }
    {   // the file must hold what the header announces before anything is allocated from it
        const long pos = ftell(fp);
        fseek(fp, 0, SEEK_END);
        const long long file_bytes = ftell(fp);
        fseek(fp, pos, SEEK_SET);
        const long long per_row = 64 + (long long)sizeof(DunkKeyPoint) + 4 + (h.has_row_id ? 4 : 0);
        const long long want = (long long)sizeof h + (long long)h.n_images * (long long)sizeof(DunkImage) + (long long)h.rows * per_row;
        if (file_bytes < want) {
            fclose(fp);
            set_error("dunk_db_load: %s is truncated (%lld bytes, the header announces %lld)", path, file_bytes, want);
            return DUNK_ERR_BAD_ARG;
        }
    }
    dunk_db* db = nullptr;
    int rc = dunk_db_create(ctx, std::max<int64_t>(std::max<int64_t>(h.rows, min_capacity_rows), 1), h.desc_bytes, &db);
    if (rc) { fclose(fp); return rc; }
    auto fail = [&](int code, const char* what) {
        fclose(fp);
        dunk_db_destroy(db);
        set_error("dunk_db_load: %s (%s)", what, path);
        return code;
    };
    db->images.resize(h.n_images);
    if (h.n_images && fread(db->images.data(), sizeof(DunkImage), h.n_images, fp) != (size_t)h.n_images) return fail(DUNK_ERR_BAD_ARG, "truncated image table");
    if (h.has_row_id && cudaMalloc(&db->row_id, (size_t)db->capacity * 4) != cudaSuccess) {
        cudaGetLastError();
        return fail(DUNK_ERR_NO_MEM, "row id column allocation failed");
    }
    SlotGuard g(ctx);
    cudaStream_t st = g.stream();
    const size_t chunk = (size_t)64 << 20;
    // two pinned halves: the file read of chunk k+1 overlaps the H2D copy of chunk k
    char* pin = (char*)ctx->pin_scratch(g.s, 2 * chunk);
    if (!pin) return fail(DUNK_ERR_NO_MEM, "pinned staging allocation failed");
    cudaEvent_t ev[2];
    cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
    int half = 0;
    bool used[2] = {false, false};
    auto fill = [&](void* dev, size_t bytes) -> bool {
        for (size_t off = 0; off < bytes; off += chunk) {
            const size_t m = std::min(chunk, bytes - off);
            if (used[half]) cudaEventSynchronize(ev[half]);
            if (fread(pin + half * chunk, 1, m, fp) != m) return false;
            if (cudaMemcpyAsync((char*)dev + off, pin + half * chunk, m, cudaMemcpyHostToDevice, st) != cudaSuccess) return false;
            cudaEventRecord(ev[half], st);
            used[half] = true;
            half ^= 1;
        }
        return true;
    };
    bool ok = fill(db->desc64, (size_t)h.rows * 64) && fill(db->kps, (size_t)h.rows * sizeof(DunkKeyPoint)) &&
              fill(db->image_id, (size_t)h.rows * 4) && (!h.has_row_id || fill(db->row_id, (size_t)h.rows * 4));
    cudaStreamSynchronize(st);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (!ok) return fail(DUNK_ERR_BAD_ARG, "truncated column data");
    fclose(fp);
    db->size = h.rows;
    db->image_lod_dirty = true;
    *out = db;
    return DUNK_OK;
}

}  // extern "C"
