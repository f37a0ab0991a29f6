// Context / slot pool / error string / timers.
#include "ctx.h"
#include <cstdarg>
#include <cstdio>

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
        if (s.ev0) cudaEventDestroy(s.ev0);
        if (s.ev1) cudaEventDestroy(s.ev1);
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
