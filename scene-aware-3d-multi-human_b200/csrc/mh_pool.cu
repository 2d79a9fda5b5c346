// Process-level pool of device buffers.  A job that fits many sequences (mhmocap/predict.py builds one optimiser per sequence) creates and
// destroys contexts of the same shape over and over; a C3-size context is 8.6 GB of device memory in ~60 buffers, and cudaMalloc / cudaFree
// of that much were measured at 0.05-0.33 s / 0.05-0.70 s per context on a B200 box (tools/time_ctor.py) -- as much as 15 optimisation
// cycles.  Buffers of 256 KB and more that a context frees are kept here, keyed by (device, size), and handed to the next request of
// exactly that size (NOT cleared: every buffer the library needs zeroed is cleared where it is allocated; MH_POOL_POISON=1 fills recycled
// buffers with 0xff bytes to prove it).  When the device runs out of memory the pool is emptied and the request repeated.  MH_POOL=0 turns
// the pool off; mh_pool_trim() returns everything to the driver; at most 32 GB (MH_POOL_MAX_GB) are parked at a time, so that other
// allocators of the process (torch's) are not starved by contexts of many different shapes.
#include "mh_ctx.h"

#include <map>
#include <mutex>
#include <unordered_map>

namespace {
struct Pool {
    std::mutex m;
    std::unordered_map<void*, std::pair<int, size_t>> live;          // pooled-size buffers in use: pointer -> (device, bytes)
    std::multimap<std::pair<int, size_t>, void*> idle;                // free buffers
    size_t idle_bytes = 0;
    size_t cap = (size_t)32 << 30;                                    // most bytes parked at a time (MH_POOL_MAX_GB)
    bool on = true, poison = false;
    Pool() {
        const char* v = getenv("MH_POOL");
        on = !(v && atoi(v) == 0);
        v = getenv("MH_POOL_MAX_GB");
        if (v && atof(v) >= 0) cap = (size_t)(atof(v) * (double)(1 << 30));
        v = getenv("MH_POOL_POISON");
        poison = v && atoi(v) != 0;
    }
};
Pool& pool() { static Pool* p = new Pool(); return *p; }             // never destroyed: the CUDA context may be gone at exit
const size_t kMinPooled = 256 << 10;

void trim_locked(Pool& P) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : P.idle) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
    P.idle.clear();
    P.idle_bytes = 0;
    cudaSetDevice(cur);
}
}  // namespace

cudaError_t mh_dev_alloc(void** p, size_t bytes) {
    Pool& P = pool();
    if (!P.on || bytes < kMinPooled) return cudaMalloc(p, bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> g(P.m);
        auto it = P.idle.find({dev, bytes});
        if (it != P.idle.end()) {
            *p = it->second;
            P.idle.erase(it);
            P.idle_bytes -= bytes;
            P.live[*p] = {dev, bytes};
        } else {
            *p = nullptr;
        }
    }
    if (*p) {
        if (!P.poison) return cudaSuccess;
        cudaError_t e = cudaMemsetAsync(*p, 0xff, bytes, 0);
        if (e == cudaSuccess) e = cudaStreamSynchronize(0);
        return e;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        { std::lock_guard<std::mutex> g(P.m); trim_locked(P); }
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> g(P.m); P.live[*p] = {dev, bytes}; }
    return e;
}

cudaError_t mh_dev_free(void* p) {
    if (!p) return cudaSuccess;
    Pool& P = pool();
    if (P.on) {
        bool pooled;
        { std::lock_guard<std::mutex> g(P.m); pooled = P.live.count(p) != 0; }
        // cudaFree waits for the work that may still use the buffer; a pooled buffer must be just as idle before its next owner gets it
        if (pooled) cudaDeviceSynchronize();
    }
    {
        std::lock_guard<std::mutex> g(P.m);
        auto it = P.live.find(p);
        if (it != P.live.end() && P.idle_bytes + it->second.second > P.cap) {      // the pool is full: back to the driver
            P.live.erase(it);
            it = P.live.end();
        }
        if (it != P.live.end()) {
            P.idle.insert({it->second, p});
            P.idle_bytes += it->second.second;
            P.live.erase(it);
            return cudaSuccess;
        }
    }
    return cudaFree(p);
}

extern "C" void mh_pool_trim(void) {
    Pool& P = pool();
    std::lock_guard<std::mutex> g(P.m);
    trim_locked(P);
}

extern "C" int64_t mh_pool_bytes(void) {
    Pool& P = pool();
    std::lock_guard<std::mutex> g(P.m);
    return (int64_t)P.idle_bytes;
}
