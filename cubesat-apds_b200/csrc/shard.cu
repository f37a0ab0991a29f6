// The shard group: NCCL over NVLink / NVSwitch for the one exchange step of the path.
//
// The reference descriptor database is sharded over the GPUs of a box by contiguous row range (global row
// index = shard base + local row).  north_star: "each GPU computes a local top-2 per query and the shards are
// merged with an NCCL allgather over NVLink"; extraction and RANSAC / PnP partition by frame with no collective.
// This file owns the communicator and the collectives; the kernels either side live in match_hamming.cu
// (local top-2, lexicographic merge, ratio test) and pipeline.cu (the sharded registration step).
//
// NCCL is resolved with dlopen("libnccl.so.2") at the first group creation instead of a DT_NEEDED entry: a host
// process that also loads PyTorch must end up with ONE libnccl (the SONAME lookup returns the copy that is
// already mapped), whichever of the two libraries was loaded first.
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstring>
#include <mutex>
#include "shard.h"

namespace dunk {
namespace {

struct NcclApi {
    ncclResult_t (*GetVersion)(int*);
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
};

const NcclApi* nccl_api() {
    static NcclApi api;
    static bool ok = false, tried = false;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (tried) return ok ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        set_error("dunk_shard_group: libnccl.so.2 not found (%s)", dlerror());
        return nullptr;
    }
    bool all = true;
    auto sym = [&](const char* name) {
        void* p = dlsym(h, name);
        if (!p) all = false;
        return p;
    };
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    if (!all) {
        set_error("dunk_shard_group: libnccl.so.2 lacks a required entry point");
        return nullptr;
    }
    ok = true;
    return &api;
}

#define DUNK_NCCL(expr)                                                                                  \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) {                                                                         \
            dunk::set_error("%s failed: %s (%s:%d)", #expr, nccl_api()->GetErrorString(_r), __FILE__, __LINE__); \
            return DUNK_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

}  // namespace

int shard_all_gather(dunk_shard_group* g, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st) {
    if (g->world == 1) {
        if (send != recv) DUNK_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, st));
        return DUNK_OK;
    }
    const NcclApi* n = nccl_api();
    DUNK_NCCL(n->AllGather(send, recv, bytes_per_rank, ncclChar, (ncclComm_t)g->comm, st));
    return DUNK_OK;
}

int shard_all_to_all(dunk_shard_group* g, const void* send, void* recv, size_t bytes_per_peer, cudaStream_t st) {
    if (g->world == 1) {
        if (send != recv) DUNK_CUDA(cudaMemcpyAsync(recv, send, bytes_per_peer, cudaMemcpyDeviceToDevice, st));
        return DUNK_OK;
    }
    const NcclApi* n = nccl_api();
    DUNK_NCCL(n->GroupStart());
    for (int r = 0; r < g->world; ++r) {
        DUNK_NCCL(n->Send((const char*)send + (size_t)r * bytes_per_peer, bytes_per_peer, ncclChar, r, (ncclComm_t)g->comm, st));
        DUNK_NCCL(n->Recv((char*)recv + (size_t)r * bytes_per_peer, bytes_per_peer, ncclChar, r, (ncclComm_t)g->comm, st));
    }
    DUNK_NCCL(n->GroupEnd());
    return DUNK_OK;
}

}  // namespace dunk

void* dunk_shard_group::ensure(int i, size_t bytes, cudaStream_t st) {
    if (bytes <= cap[i]) return buf[i];
    cudaStreamSynchronize(st);
    if (buf[i]) cudaFree(buf[i]);
    buf[i] = nullptr;
    cap[i] = 0;
    const size_t want = bytes + bytes / 8 + (1 << 20);
    if (cudaMalloc(&buf[i], want) != cudaSuccess) {
        cudaGetLastError();
        dunk::set_error("dunk_shard_group: exchange buffer of %zu bytes failed", want);
        return nullptr;
    }
    cap[i] = want;
    return buf[i];
}

using namespace dunk;

extern "C" {

int dunk_shard_unique_id(uint8_t* id128) {
    DUNK_REQUIRE(id128, DUNK_ERR_BAD_ARG, "dunk_shard_unique_id: NULL argument");
    static_assert(sizeof(ncclUniqueId) == DUNK_SHARD_ID_BYTES, "ncclUniqueId is 128 bytes");
    const NcclApi* n = nccl_api();
    if (!n) return DUNK_ERR_CUDA;
    ncclUniqueId id;
    DUNK_NCCL(n->GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return DUNK_OK;
}

void dunk_shard_group_destroy(dunk_shard_group* g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    cudaDeviceSynchronize();
    if (g->comm) nccl_api()->CommDestroy((ncclComm_t)g->comm);
    for (int i = 0; i < 4; ++i)
        if (g->buf[i]) cudaFree(g->buf[i]);
    if (g->kps_all) cudaFree(g->kps_all);
    if (g->d_counts) cudaFree(g->d_counts);
    if (g->h_counts) cudaFreeHost(g->h_counts);
    delete g;
}

int dunk_shard_group_create(dunk_ctx* ctx, int rank, int world, const uint8_t* id128, dunk_shard_group** out) {
    DUNK_REQUIRE(ctx && out, DUNK_ERR_BAD_ARG, "dunk_shard_group_create: NULL argument");
    *out = nullptr;
    DUNK_REQUIRE(world >= 1 && rank >= 0 && rank < world, DUNK_ERR_BAD_ARG, "dunk_shard_group_create: rank %d of %d", rank, world);
    DUNK_REQUIRE(world == 1 || id128, DUNK_ERR_BAD_ARG, "dunk_shard_group_create: world > 1 needs the unique id of rank 0");
    DUNK_CUDA(cudaSetDevice(ctx->device));
    dunk_shard_group* g = new dunk_shard_group();
    g->ctx = ctx;
    g->rank = rank;
    g->world = world;
    g->bases.assign(world + 1, 0);
    if (cudaMalloc(&g->d_counts, (size_t)2 * world * 8) != cudaSuccess || cudaMallocHost(&g->h_counts, (size_t)2 * world * 8) != cudaSuccess) {
        cudaGetLastError();
        dunk_shard_group_destroy(g);
        set_error("dunk_shard_group_create: count buffers");
        return DUNK_ERR_NO_MEM;
    }
    if (world > 1) {
        const NcclApi* n = nccl_api();
        if (!n) {
            dunk_shard_group_destroy(g);
            return DUNK_ERR_CUDA;
        }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        ncclComm_t comm = nullptr;
        const ncclResult_t r = n->CommInitRank(&comm, world, id, rank);
        if (r != ncclSuccess) {
            set_error("ncclCommInitRank(rank %d of %d) failed: %s", rank, world, n->GetErrorString(r));
            dunk_shard_group_destroy(g);
            return DUNK_ERR_CUDA;
        }
        g->comm = comm;
    }
    *out = g;
    return DUNK_OK;
}

int dunk_shard_group_rank(dunk_shard_group* g) { return g ? g->rank : -1; }
int dunk_shard_group_world(dunk_shard_group* g) { return g ? g->world : 0; }
int64_t dunk_shard_group_total_rows(dunk_shard_group* g) { return g ? g->total_rows : 0; }
int64_t dunk_shard_group_base(dunk_shard_group* g, int rank) {
    return (g && rank >= 0 && rank <= g->world) ? g->bases[rank] : -1;
}
int dunk_nccl_version(void) {
    const NcclApi* n = nccl_api();
    int v = 0;
    if (n) n->GetVersion(&v);
    return v;
}

/* Re-cut the rows the ranks built locally (rank r holds built->size rows; their global order is rank-major) into
 * `world` equal contiguous row ranges — tiles of coarser LoDs carry more keypoints, equal ROW counts are what
 * balances the matcher — and replicate the keypoint column on every rank.  Descriptors and image ids move with
 * exact-size ncclSend / ncclRecv pairs between the ranks whose old and new ranges overlap (no full replication
 * of the 64-byte rows); the 28-byte keypoint rows are gathered with one ncclBroadcast per source rank. */
int dunk_shard_group_balance(dunk_shard_group* g, dunk_db* built, dunk_db** shard_out) {
    DUNK_REQUIRE(g && built && shard_out, DUNK_ERR_BAD_ARG, "dunk_shard_group_balance: NULL argument");
    *shard_out = nullptr;
    dunk_ctx* ctx = g->ctx;
    DUNK_REQUIRE(built->ctx == ctx, DUNK_ERR_BAD_ARG, "dunk_shard_group_balance: the DB lives on another context");
    SlotGuard sg(ctx);
    cudaStream_t st = sg.stream();
    const int W = g->world, me = g->rank;
    const NcclApi* n = W > 1 ? nccl_api() : nullptr;
    // (1) row counts of every rank
    g->h_counts[0] = built->size;
    DUNK_CUDA(cudaMemcpyAsync(g->d_counts + W, g->h_counts, 8, cudaMemcpyHostToDevice, st));
    int rc = shard_all_gather(g, g->d_counts + W, g->d_counts, 8, st);
    if (rc) return rc;
    DUNK_CUDA(cudaMemcpyAsync(g->h_counts, g->d_counts, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    std::vector<int64_t> eb(W + 1, 0);
    for (int r = 0; r < W; ++r) eb[r + 1] = eb[r] + g->h_counts[r];
    const int64_t total = eb[W];
    DUNK_REQUIRE(total < 0xFFFFFFFFll, DUNK_ERR_BAD_ARG, "dunk_shard_group_balance: %lld rows exceed 32-bit row indices", (long long)total);
    for (int r = 0; r <= W; ++r) g->bases[r] = total * r / W;
    g->total_rows = total;
    const int64_t lo = g->bases[me], hi = g->bases[me + 1];
    // (2) replicated keypoint column
    if (g->kps_all) cudaFree(g->kps_all);
    g->kps_all = nullptr;
    if (cudaMalloc(&g->kps_all, (size_t)std::max<int64_t>(total, 1) * sizeof(DunkKeyPoint)) != cudaSuccess) {
        cudaGetLastError();
        set_error("dunk_shard_group_balance: keypoint column of %lld rows", (long long)total);
        return DUNK_ERR_NO_MEM;
    }
    if (W == 1) {
        DUNK_CUDA(cudaMemcpyAsync(g->kps_all, built->kps, (size_t)total * sizeof(DunkKeyPoint), cudaMemcpyDeviceToDevice, st));
    } else {
        DUNK_NCCL(n->GroupStart());
        for (int r = 0; r < W; ++r) {
            const int64_t cnt = eb[r + 1] - eb[r];
            if (cnt == 0) continue;
            DUNK_NCCL(n->Broadcast(built->kps, g->kps_all + eb[r], (size_t)cnt * sizeof(DunkKeyPoint), ncclChar, r, (ncclComm_t)g->comm, st));
        }
        DUNK_NCCL(n->GroupEnd());
    }
    // (3) the rank's new shard: descriptor and image-id rows from the ranks whose built range overlaps [lo, hi)
    dunk_db* shard = nullptr;
    if ((rc = dunk_db_create(ctx, std::max<int64_t>(1, hi - lo), built->desc_bytes, &shard))) return rc;
    if (W > 1) DUNK_NCCL(n->GroupStart());
    for (int s = 0; s < W; ++s)
        for (int d = 0; d < W; ++d) {
            const int64_t a = std::max(eb[s], g->bases[d]), b = std::min(eb[s + 1], g->bases[d + 1]);
            if (a >= b || (s != me && d != me)) continue;
            const size_t rows = (size_t)(b - a);
            if (s == me && d == me) {
                DUNK_CUDA(cudaMemcpyAsync(shard->desc64 + (a - lo) * 4, built->desc64 + (a - eb[me]) * 4, rows * 64, cudaMemcpyDeviceToDevice, st));
                DUNK_CUDA(cudaMemcpyAsync(shard->image_id + (a - lo), built->image_id + (a - eb[me]), rows * 4, cudaMemcpyDeviceToDevice, st));
            } else if (s == me) {
                DUNK_NCCL(n->Send(built->desc64 + (a - eb[me]) * 4, rows * 64, ncclChar, d, (ncclComm_t)g->comm, st));
                DUNK_NCCL(n->Send(built->image_id + (a - eb[me]), rows * 4, ncclChar, d, (ncclComm_t)g->comm, st));
            } else {
                DUNK_NCCL(n->Recv(shard->desc64 + (a - lo) * 4, rows * 64, ncclChar, s, (ncclComm_t)g->comm, st));
                DUNK_NCCL(n->Recv(shard->image_id + (a - lo), rows * 4, ncclChar, s, (ncclComm_t)g->comm, st));
            }
        }
    if (W > 1) DUNK_NCCL(n->GroupEnd());
    DUNK_CUDA(cudaMemcpyAsync(shard->kps, g->kps_all + lo, (size_t)(hi - lo) * sizeof(DunkKeyPoint), cudaMemcpyDeviceToDevice, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    shard->size = hi - lo;
    shard->images = built->images;      // ref_image rows of the tiles this rank extracted (ids are the caller's)
    shard->image_lod_dirty = true;
    *shard_out = shard;
    return DUNK_OK;
}

/* config 3 / get_knn_matches over a sharded DB: local top-2 of the (replicated) queries on every shard, ONE
 * ncclAllGather of the 16-byte records, lexicographic (distance, index) merge, ratio test after the merge. */
int dunk_db_match_sharded_dev(dunk_shard_group* g, dunk_db* shard, int slot, const void* query64_dev, int nq, uint32_t index_base,
                              float ratio, void* top2_merged_dev, void* matches_dev, void* count_dev) {
    DUNK_REQUIRE(g && shard && query64_dev && nq >= 0, DUNK_ERR_BAD_ARG, "dunk_db_match_sharded_dev: bad argument");
    dunk_ctx* ctx = g->ctx;
    DUNK_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), DUNK_ERR_BAD_ARG, "dunk_db_match_sharded_dev: bad slot");
    if (nq == 0) return DUNK_OK;
    DUNK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[slot].stream;
    const int W = g->world;
    uint4* local = (uint4*)g->ensure(0, (size_t)nq * 16, st);
    uint4* parts = (uint4*)g->ensure(1, (size_t)W * nq * 16, st);
    uint4* merged = top2_merged_dev ? (uint4*)top2_merged_dev : (uint4*)g->ensure(2, (size_t)nq * 16, st);
    if (!local || !parts || !merged) return DUNK_ERR_NO_MEM;
    int rc;
    if ((rc = dunk_db_knn2_dev(shard, slot, query64_dev, nq, index_base, W == 1 ? (void*)merged : (void*)local))) return rc;
    if (W > 1) {
        {
            ProfScope ps(ctx, st, "shard.all_gather_top2", (double)W * nq * 16);
            if ((rc = shard_all_gather(g, local, parts, (size_t)nq * 16, st))) return rc;
        }
        ProfScope ps(ctx, st, "match.top2_merge", (double)W * nq * 16);
        if ((rc = launch_top2_merge(ctx, st, parts, W, nq, merged))) return rc;
    }
    if (matches_dev && count_dev) return launch_top2_ratio(ctx, st, merged, nq, ratio, (DunkDMatch*)matches_dev, (int*)count_dev);
    return DUNK_OK;
}

/* host-buffer form: what get_knn_matches (feature_extraction/src/lib.rs:94-114) becomes when the train set is the
 * sharded reference DB; every rank passes the same queries and receives the same matches */
int dunk_db_match_sharded(dunk_shard_group* g, dunk_db* shard, const uint8_t* query, int nq, uint32_t index_base, float ratio,
                          DunkDMatch* out, int out_cap, int* n_out) {
    DUNK_REQUIRE(g && shard && n_out && nq >= 0, DUNK_ERR_BAD_ARG, "dunk_db_match_sharded: bad argument");
    *n_out = 0;
    if (nq == 0) return DUNK_OK;
    DUNK_REQUIRE(query && out, DUNK_ERR_BAD_ARG, "dunk_db_match_sharded: NULL buffers");
    DUNK_REQUIRE(g->total_rows >= 2 || shard->size >= 2 || g->world > 1, DUNK_ERR_OUT_OF_RANGE,
                 "dunk_db_match_sharded: fewer than 2 reference rows (lib.rs:108 `get(1)?`)");
    dunk_ctx* ctx = g->ctx;
    const int slot = ctx->acquire();
    struct Rel { dunk_ctx* c; int s; ~Rel() { c->release(s); } } rel{ctx, slot};
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->slots[slot].stream;
    const int db = shard->desc_bytes;
    // (the slot's own device scratch holds the matcher's slab partials, so the staging rows live in a group buffer)
    void* scratch = g->ensure(3, Carver::need((size_t)nq * db) + Carver::need((size_t)nq * 64) + Carver::need((size_t)nq * 16) + 256, st);
    unsigned char* pin = (unsigned char*)ctx->pin_scratch(slot, (size_t)nq * 64 + (size_t)nq * 16 + 16);
    if (!scratch || !pin) return DUNK_ERR_NO_MEM;
    Carver cv(scratch);
    uint8_t* d_raw = cv.take<uint8_t>((size_t)nq * db);
    uint4* d_q = cv.take<uint4>((size_t)nq * 4);
    DunkDMatch* d_m = cv.take<DunkDMatch>(nq);
    int* d_c = cv.take<int>(1);
    memcpy(pin, query, (size_t)nq * db);
    DUNK_CUDA(cudaMemcpyAsync(d_raw, pin, (size_t)nq * db, cudaMemcpyHostToDevice, st));
    int rc;
    if ((rc = launch_pad_rows(ctx, st, d_raw, nq, db, d_q))) return rc;
    if ((rc = dunk_db_match_sharded_dev(g, shard, slot, d_q, nq, index_base, ratio, nullptr, d_m, d_c))) return rc;
    DunkDMatch* h_m = (DunkDMatch*)(pin + (size_t)nq * 64);
    int* h_c = (int*)(pin + (size_t)nq * 64 + (size_t)nq * 16);
    DUNK_CUDA(cudaMemcpyAsync(h_m, d_m, (size_t)nq * 16, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaMemcpyAsync(h_c, d_c, 4, cudaMemcpyDeviceToHost, st));
    DUNK_CUDA(cudaStreamSynchronize(st));
    DUNK_REQUIRE(*h_c <= out_cap, DUNK_ERR_NO_MEM, "dunk_db_match_sharded: %d matches exceed the output capacity %d", *h_c, out_cap);
    memcpy(out, h_m, (size_t)*h_c * 16);
    *n_out = *h_c;
    return DUNK_OK;
}

}  // extern "C"
