// Context / slot pool / error string / timers.
#include "ctx.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <thread>
#include <unistd.h>

namespace dunk {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

// Persistent host workers: a host memcpy moves ~15 GB/s per core on the GPU boxes and scales to ~85 GB/s on 12 cores;
// spawning threads per call cost more than the 64 MB copies they served.  Work items are claimed from an atomic
// cursor, so a core that is busy elsewhere simply takes fewer of them.  One job at a time; a second concurrent caller
// runs its items inline.
namespace {
class HostPool {
    std::mutex mu, call_mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    const std::function<void(size_t)>* fn = nullptr;
    size_t n_items = 0;
    std::atomic<size_t> cursor{0};
    unsigned long long generation = 0;
    int active = 0;
    bool stop = false;
    pid_t owner = 0;

    void drain() {
        for (;;) {
            const size_t i = cursor.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_items) return;
            (*fn)(i);
        }
    }
    void worker() {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            lk.unlock();
            drain();
            lk.lock();
            if (--active == 0) cv_done.notify_one();
        }
    }

public:
    HostPool() {
        unsigned want = std::thread::hardware_concurrency();
        want = want > 3 ? std::min(12u, want - 2) : 0;
        if (const char* e = getenv("DUNK_COPY_THREADS")) want = (unsigned)std::max(0, atoi(e) - 1);
        owner = getpid();
        try {
            for (unsigned i = 0; i < want; ++i) workers.emplace_back([this] { worker(); });
        } catch (...) {
            // thread creation refused (resource limit): whatever was started serves; with none the caller copies inline
        }
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto& t : workers) t.join();
    }
    void run(size_t items, const std::function<void(size_t)>& f) {
        // a forked child inherits the object but not the worker threads: it runs its items inline
        if (items < 2 || workers.empty() || getpid() != owner || !call_mu.try_lock()) {
            for (size_t i = 0; i < items; ++i) f(i);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = &f;
            n_items = items;
            cursor.store(0, std::memory_order_relaxed);
            active = (int)workers.size();
            ++generation;
        }
        cv_work.notify_all();
        drain();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return active == 0; });
        }
        call_mu.unlock();
    }
};
HostPool& host_pool() {
    static HostPool pool;
    return pool;
}
}  // namespace

void par_for(size_t items, const std::function<void(size_t)>& f) { host_pool().run(items, f); }

void par_memcpy(void* dst, const void* src, size_t n) {
    constexpr size_t kChunk = (size_t)2 << 20;
    if (n < 2 * kChunk) {
        memcpy(dst, src, n);
        return;
    }
    par_for((n + kChunk - 1) / kChunk, [=](size_t i) {
        const size_t o = i * kChunk;
        memcpy((char*)dst + o, (const char*)src + o, std::min(kChunk, n - o));
    });
}

static std::vector<DeviceInitFn>& device_init_hooks() {
    static std::vector<DeviceInitFn> hooks;
    return hooks;
}
DeviceInitReg::DeviceInitReg(DeviceInitFn fn) { device_init_hooks().push_back(fn); }

}  // namespace dunk

int dunk_ctx::acquire() {
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
        for (size_t i = 0; i < slots.size(); ++i)
            if (!slots[i].busy) {
                slots[i].busy = true;
                return (int)i;
            }
        cv.wait(lk);
    }
}

void dunk_ctx::release(int s) {
    {
        std::lock_guard<std::mutex> lk(mu);
        slots[s].busy = false;
    }
    cv.notify_one();
}

void* dunk_ctx::dev_scratch(int s, size_t bytes) {
    dunk::Slot& sl = slots[s];
    if (bytes <= sl.dev_bytes) return sl.dev;
    // the old block may still be referenced by work queued on this slot's stream
    cudaStreamSynchronize(sl.stream);
    if (sl.dev) cudaFree(sl.dev);
    sl.dev = nullptr;
    sl.dev_bytes = 0;
    size_t want = bytes + bytes / 4 + (1 << 20);
    if (cudaMalloc(&sl.dev, want) != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        if (cudaMalloc(&sl.dev, want) != cudaSuccess) {
            cudaGetLastError();
            dunk::set_error("device scratch allocation of %zu bytes failed", bytes);
            return nullptr;
        }
    }
    sl.dev_bytes = want;
    return sl.dev;
}

void* dunk_ctx::pin_scratch(int s, size_t bytes) {
    dunk::Slot& sl = slots[s];
    if (bytes <= sl.pin_bytes) return sl.pin;
    cudaStreamSynchronize(sl.stream);
    if (sl.pin) cudaFreeHost(sl.pin);
    sl.pin = nullptr;
    sl.pin_bytes = 0;
    size_t want = bytes + bytes / 4 + (1 << 16);
    if (cudaMallocHost(&sl.pin, want) != cudaSuccess) {
        cudaGetLastError();
        dunk::set_error("pinned host scratch allocation of %zu bytes failed", bytes);
        return nullptr;
    }
    sl.pin_bytes = want;
    return sl.pin;
}

void* dunk_ctx::upload_ring(int s) {
    dunk::Slot& sl = slots[s];
    if (!sl.ring && cudaMallocHost(&sl.ring, (size_t)128 << 20) != cudaSuccess) {
        cudaGetLastError();
        sl.ring = nullptr;
        dunk::set_error("upload ring allocation failed");
    }
    return sl.ring;
}

int dunk_ctx::upload_pageable(int s, void* dst_dev, const void* src_host, size_t nbytes) {
    dunk::Slot& sl = slots[s];
    const size_t half = (size_t)64 << 20;
    if (nbytes < ((size_t)1 << 20)) {
        DUNK_CUDA(cudaMemcpyAsync(dst_dev, src_host, nbytes, cudaMemcpyHostToDevice, sl.stream));
        return DUNK_OK;
    }
    // a dedicated pinned ring (the slot's pin scratch may hold the caller's other staging data)
    if (!upload_ring(s)) return DUNK_ERR_NO_MEM;
    cudaEvent_t ev[2] = {sl.ev0, sl.ev1};
    int h = 0;
    for (size_t off = 0; off < nbytes; off += half, h ^= 1) {
        const size_t n = std::min(half, nbytes - off);
        // the half's previous copy (of this or an earlier call) must have left it; a never-recorded event is complete
        DUNK_CUDA(cudaEventSynchronize(ev[h]));
        dunk::par_memcpy((char*)sl.ring + h * half, (const char*)src_host + off, n);
        DUNK_CUDA(cudaMemcpyAsync((char*)dst_dev + off, (char*)sl.ring + h * half, n, cudaMemcpyHostToDevice, sl.stream));
        DUNK_CUDA(cudaEventRecord(ev[h], sl.stream));
    }
    return DUNK_OK;
}

extern "C" {

const char* dunk_last_error(void) { return dunk::g_err.c_str(); }
const char* dunk_version(void) { return "dunk_b200 0.1 (sm_100a)"; }

int dunk_ctx_create(int device, int n_slots, dunk_ctx** out) {
    DUNK_REQUIRE(out != nullptr, DUNK_ERR_BAD_ARG, "dunk_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        dunk::set_error("dunk_ctx_create: no CUDA device (%s); there is no CPU fallback",
                        e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return DUNK_ERR_CUDA;
    }
    DUNK_REQUIRE(device >= 0 && device < ndev, DUNK_ERR_BAD_ARG,
                 "dunk_ctx_create: device %d out of range (0..%d)", device, ndev - 1);
    if (n_slots < 1) n_slots = 1;
    if (n_slots > 64) n_slots = 64;
    DUNK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DUNK_CUDA(cudaGetDeviceProperties(&prop, device));
    DUNK_REQUIRE(prop.major >= 10, DUNK_ERR_CUDA,
                 "dunk_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                 device, prop.major, prop.minor);
    dunk_ctx* c = new dunk_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->slots.resize(n_slots);
    for (auto& s : c->slots) {
        DUNK_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        DUNK_CUDA(cudaEventCreate(&s.ev0));
        DUNK_CUDA(cudaEventCreate(&s.ev1));
        DUNK_CUDA(cudaEventCreateWithFlags(&s.ev2, cudaEventDisableTiming));
        DUNK_CUDA(cudaEventCreateWithFlags(&s.ev3, cudaEventDisableTiming));
        DUNK_CUDA(cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking));
    }
    for (dunk::DeviceInitFn fn : dunk::device_init_hooks()) {
        const int rc = fn(c);
        if (rc != DUNK_OK) {
            dunk_ctx_destroy(c);
            return rc;
        }
    }
    *out = c;
    return DUNK_OK;
}

void dunk_ctx_destroy(dunk_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& s : c->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.dev) cudaFree(s.dev);
        if (s.pin) cudaFreeHost(s.pin);
        if (s.ring) cudaFreeHost(s.ring);
        if (s.ev0) cudaEventDestroy(s.ev0);
        if (s.ev1) cudaEventDestroy(s.ev1);
        if (s.ev2) cudaEventDestroy(s.ev2);
        if (s.ev3) cudaEventDestroy(s.ev3);
        if (s.stream2) { cudaStreamSynchronize(s.stream2); cudaStreamDestroy(s.stream2); }
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete c;
}

void* dunk_ctx_stream(dunk_ctx* c, int slot) {
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return nullptr;
    return (void*)c->slots[slot].stream;
}
int dunk_ctx_device(dunk_ctx* c) { return c ? c->device : -1; }
int dunk_ctx_sm_count(dunk_ctx* c) { return c ? c->sm_count : 0; }
uint64_t dunk_ctx_launch_count(dunk_ctx* c) { return c ? c->launches.load() : 0; }

int dunk_timer_begin(dunk_ctx* c, int slot) {
    DUNK_REQUIRE(c && slot >= 0 && slot < (int)c->slots.size(), DUNK_ERR_BAD_ARG, "bad slot");
    DUNK_CUDA(cudaSetDevice(c->device));
    DUNK_CUDA(cudaEventRecord(c->slots[slot].ev0, c->slots[slot].stream));
    return DUNK_OK;
}
int dunk_timer_end(dunk_ctx* c, int slot, float* ms) {
    DUNK_REQUIRE(c && slot >= 0 && slot < (int)c->slots.size() && ms, DUNK_ERR_BAD_ARG, "bad slot");
    DUNK_CUDA(cudaSetDevice(c->device));
    DUNK_CUDA(cudaEventRecord(c->slots[slot].ev1, c->slots[slot].stream));
    DUNK_CUDA(cudaEventSynchronize(c->slots[slot].ev1));
    DUNK_CUDA(cudaEventElapsedTime(ms, c->slots[slot].ev0, c->slots[slot].ev1));
    return DUNK_OK;
}
int dunk_sync(dunk_ctx* c, int slot) {
    DUNK_REQUIRE(c && slot >= 0 && slot < (int)c->slots.size(), DUNK_ERR_BAD_ARG, "bad slot");
    DUNK_CUDA(cudaSetDevice(c->device));
    DUNK_CUDA(cudaStreamSynchronize(c->slots[slot].stream));
    return DUNK_OK;
}

int dunk_profile_begin(dunk_ctx* c) {
    DUNK_REQUIRE(c, DUNK_ERR_BAD_ARG, "dunk_profile_begin: ctx is NULL");
    std::lock_guard<std::mutex> lk(c->prof_mu);
    for (auto& p : c->prof)
        for (auto& e : p.ev) {
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
    c->prof.clear();
    c->prof_on = true;
    return DUNK_OK;
}

int dunk_profile_end(dunk_ctx* c, char* names, int names_cap, double* ms, int* launches, double* alg, int cap) {
    DUNK_REQUIRE(c && names && ms && launches && alg && names_cap > 0, DUNK_ERR_BAD_ARG, "dunk_profile_end: NULL argument");
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(c->prof_mu);
    c->prof_on = false;
    std::string all;
    int n = 0;
    for (auto& p : c->prof) {
        if (n >= cap) break;
        double t = 0;
        for (auto& e : p.ev) {
            float x = 0;
            if (cudaEventElapsedTime(&x, e.first, e.second) == cudaSuccess) t += x;
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
        p.ev.clear();
        ms[n] = t;
        launches[n] = p.count;
        alg[n] = p.alg;
        all += p.name;
        all += ';';
        ++n;
    }
    c->prof.clear();
    snprintf(names, names_cap, "%s", all.c_str());
    return n;
}

int dunk_memcpy_h2d(dunk_ctx* c, int slot, void* dst_dev, const void* src_host, size_t nbytes) {
    DUNK_REQUIRE(c && slot >= 0 && slot < (int)c->slots.size(), DUNK_ERR_BAD_ARG, "dunk_memcpy_h2d: bad ctx / slot");
    if (nbytes == 0) return DUNK_OK;
    DUNK_REQUIRE(dst_dev && src_host, DUNK_ERR_BAD_ARG, "dunk_memcpy_h2d: NULL pointer");
    DUNK_CUDA(cudaSetDevice(c->device));
    DUNK_CUDA(cudaMemcpyAsync(dst_dev, src_host, nbytes, cudaMemcpyHostToDevice, c->slots[slot].stream));
    return DUNK_OK;
}
int dunk_memcpy_d2h(dunk_ctx* c, int slot, void* dst_host, const void* src_dev, size_t nbytes) {
    DUNK_REQUIRE(c && slot >= 0 && slot < (int)c->slots.size(), DUNK_ERR_BAD_ARG, "dunk_memcpy_d2h: bad ctx / slot");
    if (nbytes == 0) return DUNK_OK;
    DUNK_REQUIRE(dst_host && src_dev, DUNK_ERR_BAD_ARG, "dunk_memcpy_d2h: NULL pointer");
    DUNK_CUDA(cudaSetDevice(c->device));
    DUNK_CUDA(cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, c->slots[slot].stream));
    return DUNK_OK;
}
int dunk_dev_alloc(dunk_ctx* c, size_t nbytes, void** out_dev) {
    DUNK_REQUIRE(c && out_dev, DUNK_ERR_BAD_ARG, "dunk_dev_alloc: NULL argument");
    *out_dev = nullptr;
    DUNK_CUDA(cudaSetDevice(c->device));
    if (cudaMalloc(out_dev, nbytes ? nbytes : 1) != cudaSuccess) {
        cudaGetLastError();
        dunk::set_error("dunk_dev_alloc: %zu bytes", nbytes);
        return DUNK_ERR_NO_MEM;
    }
    return DUNK_OK;
}
int dunk_dev_free(dunk_ctx* c, void* dev) {
    DUNK_REQUIRE(c, DUNK_ERR_BAD_ARG, "dunk_dev_free: ctx is NULL");
    if (!dev) return DUNK_OK;
    DUNK_CUDA(cudaSetDevice(c->device));
    DUNK_CUDA(cudaFree(dev));
    return DUNK_OK;
}
int dunk_host_alloc(dunk_ctx* c, size_t nbytes, void** out_host) {
    DUNK_REQUIRE(c && out_host, DUNK_ERR_BAD_ARG, "dunk_host_alloc: NULL argument");
    *out_host = nullptr;
    DUNK_CUDA(cudaSetDevice(c->device));
    if (cudaMallocHost(out_host, nbytes ? nbytes : 1) != cudaSuccess) {
        cudaGetLastError();
        dunk::set_error("dunk_host_alloc: %zu pinned bytes", nbytes);
        return DUNK_ERR_NO_MEM;
    }
    return DUNK_OK;
}
int dunk_host_free(dunk_ctx* c, void* host) {
    DUNK_REQUIRE(c, DUNK_ERR_BAD_ARG, "dunk_host_free: ctx is NULL");
    if (!host) return DUNK_OK;
    DUNK_CUDA(cudaSetDevice(c->device));
    DUNK_CUDA(cudaFreeHost(host));
    return DUNK_OK;
}

int dunk_ctx_reserve_slot(dunk_ctx* c) {
    DUNK_REQUIRE(c, DUNK_ERR_BAD_ARG, "dunk_ctx_reserve_slot: ctx is NULL");
    std::lock_guard<std::mutex> lk(c->mu);
    int free_slots = 0;
    for (auto& s : c->slots) free_slots += !s.busy;
    DUNK_REQUIRE(free_slots >= 2, DUNK_ERR_NO_MEM,
                 "dunk_ctx_reserve_slot: would leave no slot for host-API calls");
    for (int i = (int)c->slots.size() - 1; i >= 0; --i)
        if (!c->slots[i].busy) {
            c->slots[i].busy = true;
            return i;
        }
    return DUNK_ERR_NO_MEM;
}
int dunk_ctx_release_slot(dunk_ctx* c, int slot) {
    DUNK_REQUIRE(c && slot >= 0 && slot < (int)c->slots.size(), DUNK_ERR_BAD_ARG, "bad slot");
    c->release(slot);
    return DUNK_OK;
}

}  // extern "C"
