// Internal: context, stream/workspace slots, error plumbing.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/dunk_b200.h"

namespace dunk {

void set_error(const char* fmt, ...);
// host memcpy on the persistent host workers (a single core moves ~15 GB/s, less than PCIe 5 or the kernels consume)
void par_memcpy(void* dst, const void* src, size_t n);
// f(0) ... f(items - 1) on the library's persistent host workers (and the calling thread); returns when all are done
void par_for(size_t items, const std::function<void(size_t)>& f);

// One stream + growable device / pinned-host scratch.  A host-API call owns exactly one
// slot for its duration (SURVEY 8b "Threading").
struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // copy stream of the pipelined host-buffer calls (ordered against `stream` by events)
    void* dev = nullptr;
    size_t dev_bytes = 0;
    void* pin = nullptr;
    size_t pin_bytes = 0;
    void* ring = nullptr;             // 2 x 64 MB pinned halves of upload_pageable
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    bool busy = false;
};

}  // namespace dunk

namespace dunk {
// optional per-kernel-class CUDA-event profiler (bench.py's per-stage times and roofline)
struct ProfAccum {
    std::string name;
    double alg = 0;   // algorithmic bytes (or ops) accumulated by the launch sites
    int count = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
};
}  // namespace dunk

struct dunk_ctx {
    int device = 0;
    int sm_count = 0;
    std::vector<dunk::Slot> slots;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> launches{0};
    int hamming_occ = 1;    // resident CTAs per SM of the Hamming matcher on THIS device (set by its device-init hook)
    bool prof_on = false;
    std::vector<dunk::ProfAccum> prof;
    std::mutex prof_mu;

    int acquire();          // blocks until a slot is free
    void release(int s);
    // grow-only scratch; returns nullptr (and sets error) on failure
    void* dev_scratch(int s, size_t bytes);
    void* pin_scratch(int s, size_t bytes);
    // H2D copy of PAGEABLE caller memory at pinned speed: multi-threaded memcpy into two pinned halves of the slot,
    // each shipped with an async copy while the other half is being filled.  Asynchronous on the slot's stream for
    // the last chunk only (callers order later work on the same stream).  Uses stream2-free events ev0 / ev1.
    int upload_pageable(int s, void* dst_dev, const void* src_host, size_t nbytes);
    void* upload_ring(int s);          // the slot's 2 x 64 MB pinned ring (allocated on first use); nullptr = no memory
};

namespace dunk {

// Per-device kernel set-up (cudaFuncSetAttribute opt-ins, occupancy queries).  Function attributes are per
// DEVICE, so every translation unit registers a hook and dunk_ctx_create runs all of them on the context's
// device (a process may hold contexts on several GPUs).
using DeviceInitFn = int (*)(dunk_ctx*);
struct DeviceInitReg {
    explicit DeviceInitReg(DeviceInitFn fn);
};

struct SlotGuard {
    dunk_ctx* ctx;
    int s;
    explicit SlotGuard(dunk_ctx* c) : ctx(c), s(c->acquire()) { cudaSetDevice(c->device); }
    ~SlotGuard() { ctx->release(s); }
    Slot& slot() { return ctx->slots[s]; }
    cudaStream_t stream() { return ctx->slots[s].stream; }
};

// RAII: when profiling is on, brackets the launches issued in its scope with two CUDA events
struct ProfScope {
    dunk_ctx* ctx;
    cudaStream_t st;
    cudaEvent_t e1 = nullptr;
    ProfScope(dunk_ctx* c, cudaStream_t s, const char* name, double alg) : ctx(c), st(s) {
        if (!c->prof_on) return;
        std::lock_guard<std::mutex> lk(c->prof_mu);
        ProfAccum* a = nullptr;
        for (auto& p : c->prof)
            if (p.name == name) a = &p;
        if (!a) {
            c->prof.push_back(ProfAccum{});
            a = &c->prof.back();
            a->name = name;
        }
        cudaEvent_t e0;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        a->ev.emplace_back(e0, e1);
        a->alg += alg;
        a->count += 1;
    }
    ~ProfScope() {
        if (e1) cudaEventRecord(e1, st);
    }
};

// bump allocator over a scratch block (256-B aligned pieces)
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base((char*)b) {}
    template <class T>
    T* take(size_t n) {
        T* p = (T*)(base + off);
        off += ((n * sizeof(T)) + 255) & ~size_t(255);
        return p;
    }
    static size_t need(size_t bytes) { return (bytes + 255) & ~size_t(255); }
};

#define DUNK_CUDA(expr)                                                                    \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            dunk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                     \
            return DUNK_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define DUNK_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            dunk::set_error(__VA_ARGS__);  \
            return (code);                 \
        }                                  \
    } while (0)

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Phase timestamps of the latency-bound tail kernels (RANSAC homography, PnP): compiled in only for the
// `make timing` variant (libdunk_b200_timing.so, tools/phase_times.py); a no-op in the product build.
#if defined(DUNK_PHASE_TIMING) && defined(__CUDACC__)
static __device__ long long g_phase[64];     // one copy per translation unit (no relocatable device code)
#endif
#if defined(DUNK_PHASE_TIMING) && defined(__CUDA_ARCH__)
#define DUNK_PHASE(k)                                                     \
    do {                                                                  \
        if (blockIdx.x == 0 && threadIdx.x == 0) g_phase[k] = clock64();  \
    } while (0)
#else
#define DUNK_PHASE(k) \
    do {              \
    } while (0)
#endif

}  // namespace dunk
