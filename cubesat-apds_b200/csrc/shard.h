// Internal: the shard group — one NCCL communicator over the ranks that hold the row-range shards of the
// reference descriptor database, plus the exchange buffers of the sharded matcher.
#pragma once
#include <vector>
#include "ctx.h"
#include "match.h"

struct dunk_shard_group {
    dunk_ctx* ctx = nullptr;
    int rank = 0, world = 1;
    void* comm = nullptr;                 // ncclComm_t (NULL when world == 1)
    // keypoint column of the WHOLE database, replicated on every rank: global row -> keypoint (the frame
    // owner turns merged train indices into reference points without another exchange)
    DunkKeyPoint* kps_all = nullptr;
    int64_t total_rows = 0;
    std::vector<int64_t> bases;           // world + 1: first global row of every shard
    // grow-only exchange buffers
    void* buf[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t cap[4] = {0, 0, 0, 0};
    long long* d_counts = nullptr;        // 2 * world
    long long* h_counts = nullptr;        // pinned, 2 * world
    void* ensure(int i, size_t bytes, cudaStream_t st);
};

namespace dunk {
// thin wrappers over the NCCL entry points (resolved from libnccl.so.2 at first use); all asynchronous on st
int shard_all_gather(dunk_shard_group* g, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st);
// block r of `send` (bytes_per_peer) goes to rank r; block r of `recv` comes from rank r
int shard_all_to_all(dunk_shard_group* g, const void* send, void* recv, size_t bytes_per_peer, cudaStream_t st);
}  // namespace dunk
