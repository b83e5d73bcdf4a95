// common.cuh -- shared host/device helpers of librasr_b200 (error plumbing, launch accounting,
// device buffers, PTX wrappers for cp.async.bulk / mbarrier used by the SIMT kernels).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rasr_b200.h"

namespace rb {

// ---------------------------------------------------------------- errors
void        set_error(const char* fmt, ...);
const char* get_error();

#define RB_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            rb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return RB_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

#define RB_CHECK(expr)            \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != RB_OK)        \
            return rc__;          \
    } while (0)

#define RB_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            rb::set_error(__VA_ARGS__); \
            return RB_ERR_INVALID;     \
        }                              \
    } while (0)

// ---------------------------------------------------------------- launch accounting
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) {
    g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed);
}
// checks the launch itself (not completion)
#define RB_LAUNCH_CHECK()                                                                          \
    do {                                                                                           \
        rb::count_launch();                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess) {                                                                  \
            rb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return RB_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

// ---------------------------------------------------------------- device selection
struct DeviceInfo {
    int    ordinal = -1;
    int    sm_count = 0;
    int    cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
};
// selects `device`, verifies it is sm_100, fills info.  RB_ERR_NO_DEVICE otherwise.
int use_device(int device, DeviceInfo* info);

// ---------------------------------------------------------------- RAII device / pinned buffers
template<typename T>
struct DevBuf {
    T*     p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&)            = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
    // grows only; contents are not preserved
    int reserve(size_t count) {
        if (count <= n)
            return RB_OK;
        release();
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return RB_ERR_NOMEM;
        }
        n = count;
        return RB_OK;
    }
    int upload(const T* host, size_t count, cudaStream_t s) {
        RB_CHECK(reserve(count));
        if (count)
            RB_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
        return RB_OK;
    }
    int upload(const std::vector<T>& v, cudaStream_t s) { return upload(v.data(), v.size(), s); }
};

// The two copy streams and the events of a host-buffer call (H2D of slab i+1 / kernels of slab i / D2H of slab i-1),
// created once per handle: creating and destroying them per call costs ~0.1 ms.
struct CopyStreams {
    cudaStream_t             in = nullptr, out = nullptr;
    std::vector<cudaEvent_t> pool;
    CopyStreams() {}
    CopyStreams(const CopyStreams&)            = delete;
    CopyStreams& operator=(const CopyStreams&) = delete;
    int ensure(size_t nEvents) {
        if (!in)
            RB_CUDA(cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking));
        if (!out)
            RB_CUDA(cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking));
        while (pool.size() < nEvents) {
            cudaEvent_t e = nullptr;
            RB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            pool.push_back(e);
        }
        return RB_OK;
    }
    ~CopyStreams() {
        for (cudaEvent_t e : pool)
            cudaEventDestroy(e);
        if (in)
            cudaStreamDestroy(in);
        if (out)
            cudaStreamDestroy(out);
    }
};

template<typename T>
struct PinnedBuf {
    T*     p = nullptr;
    size_t n = 0;
    PinnedBuf() {}
    PinnedBuf(const PinnedBuf&)            = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() {
        if (p)
            cudaFreeHost(p);
    }
    int reserve(size_t count) {
        if (count <= n)
            return RB_OK;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMallocHost((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMallocHost of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return RB_ERR_NOMEM;
        }
        n = count;
        return RB_OK;
    }
};

inline size_t round_up(size_t v, size_t m) {
    return (v + m - 1) / m * m;
}

// host-to-host copy on several cores (the stager's worker pool + the calling thread)
void parallel_memcpy(void* dst, const void* src, size_t bytes);

// D2H into PAGEABLE caller memory.  cudaMemcpyAsync to pageable memory is staged by the driver and blocks the calling
// thread until the copy is done, so the H2D / kernel / D2H slab pipeline of a host-buffer call degenerates to one step
// after the other (C2: 14.9 M frames/s against 50.8 M from page-locked buffers).  The stager keeps the pipeline
// asynchronous: the DMA goes into a ring of page-locked slots, and a small process-wide pool of worker threads waits
// for each slot's event and copies it on into the caller's buffer while the next slabs are in flight.
struct HostStager {
    PinnedBuf<unsigned char> ring;
    size_t                   slotBytes = 0;
    int                      nSlots = 0, next = 0, device = 0;
    std::vector<cudaEvent_t> done;       // DMA into the slot has completed
    std::atomic<int>*        busy = nullptr;  // per slot: 1 while a worker still has to copy it out
    HostStager() {}
    HostStager(const HostStager&)            = delete;
    HostStager& operator=(const HostStager&) = delete;
    ~HostStager();
    int ensure(size_t slot_bytes, int slots, int dev);
    // enqueue: device -> slot(s) on stream s, then slot -> dst by a worker.  Returns once everything is enqueued.
    int d2h(void* dst, const void* d_src, size_t bytes, cudaStream_t s);
    // block until every enqueued copy has reached the caller's buffer; RB_ERR_CUDA if a DMA failed
    int drain();
    std::atomic<int> failed{0};
};

}  // namespace rb

// ---------------------------------------------------------------- device-side PTX wrappers
#ifdef __CUDACC__
namespace rbdev {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the init visible to the async proxy (TMA / tcgen05.commit)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as
// a hung GPU.  try_wait itself suspends the thread for a HW-defined time slice per call.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    long long t0   = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ffu) == 0) {
            long long now = clock64();
            if (t0 == 0)
                t0 = now;
            else if (now - t0 > 4000000000ll) {  // ~2 s at 2 GHz
                printf("rasr_b200: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x,
                       threadIdx.x, parity);
                __trap();
            }
        }
    }
}
// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP); bytes % 16 == 0, both
// addresses 16-byte aligned; completion is signalled as transaction bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace rbdev
#endif
