// common.cu -- error plumbing, device selection and the library-level entry points.
#include "common.cuh"

#include <condition_variable>
#include <deque>
#include <mutex>
#include <pthread.h>
#include <thread>

namespace rb {

// ---------------------------------------------------------------------------------------------------------------
// worker pool of the host stager (see common.cuh).  Created on first use, never destroyed: the threads sleep on a
// condition variable and die with the process (a static destructor could race with them at exit).
namespace {
struct CopyJob {
    cudaEvent_t       ev;
    const void*       src;
    void*             dst;
    size_t            bytes;
    std::atomic<int>* busy;
    std::atomic<int>* failed;
    int               device;
};
struct CopyPool {
    std::mutex              m;
    std::condition_variable cv;
    std::deque<CopyJob>     q;
    CopyPool() {
        unsigned n = std::thread::hardware_concurrency();
        n          = n >= 16 ? 4 : (n >= 4 ? 2 : 1);
        if (const char* e = getenv("RB_COPY_THREADS"))
            n = (unsigned)std::max(1, atoi(e));
        for (unsigned i = 0; i < n; ++i)
            std::thread([this] { run(); }).detach();
    }
    void run() {
        int dev = -1;
        for (;;) {
            CopyJob j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [this] { return !q.empty(); });
                j = q.front();
                q.pop_front();
            }
            if (j.device >= 0 && j.device != dev) {
                cudaSetDevice(j.device);
                dev = j.device;
            }
            if (j.ev && cudaEventSynchronize(j.ev) != cudaSuccess) {
                cudaGetLastError();
                j.failed->store(1);
            }
            else
                std::memcpy(j.dst, j.src, j.bytes);
            if (j.ev)
                j.busy->store(0, std::memory_order_release);
            else
                j.busy->fetch_sub(1, std::memory_order_release);  // plain copy: one part of a parallel_memcpy done
        }
    }
    // urgent: input copies gate the GPU's work and overtake the queued output copies
    void submit(const CopyJob& j, bool urgent = false) {
        {
            std::lock_guard<std::mutex> lk(m);
            if (urgent)
                q.push_front(j);
            else
                q.push_back(j);
        }
        cv.notify_one();
    }
};
// One pool per process, intentionally leaked.  A forked child inherits the pointer but not the threads: the atfork
// handler drops the pointer so that the child builds its own pool on first use.
CopyPool*  g_pool = nullptr;
std::mutex g_poolMutex;
CopyPool&  copy_pool() {
    std::lock_guard<std::mutex> lk(g_poolMutex);
    if (!g_pool) {
        static bool hooked = false;
        if (!hooked) {
            pthread_atfork(nullptr, nullptr, [] {
                g_pool = nullptr;
                new (&g_poolMutex) std::mutex();
            });
            hooked = true;
        }
        g_pool = new CopyPool();
    }
    return *g_pool;
}
}  // namespace

// host-to-host copy split over the pool's workers and the calling thread (a single core of the GPU boxes' hosts copies
// ~9 GB/s; pageable features have to reach page-locked memory faster than that to keep up with the DMA)
void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t kPart = (size_t)1 << 20;
    if (bytes < 4 * kPart) {
        std::memcpy(dst, src, bytes);
        return;
    }
    const int        parts = 4;
    const size_t     step  = round_up((bytes + parts - 1) / parts, 4096);
    std::atomic<int> pending{0};
    std::atomic<int> failed{0};
    for (int i = 1; i < parts; ++i) {
        const size_t off = (size_t)i * step;
        if (off >= bytes)
            break;
        pending.fetch_add(1);
        copy_pool().submit(CopyJob{nullptr, (const unsigned char*)src + off, (unsigned char*)dst + off,
                                   std::min(step, bytes - off), &pending, &failed, -1},
                           true);
    }
    std::memcpy(dst, src, std::min(step, bytes));
    while (pending.load(std::memory_order_acquire))
        std::this_thread::yield();
}

HostStager::~HostStager() {
    drain();
    for (cudaEvent_t e : done)
        cudaEventDestroy(e);
    delete[] busy;
}

int HostStager::ensure(size_t slot_bytes, int slots, int dev) {
    if (slot_bytes <= slotBytes && slots <= nSlots)
        return RB_OK;
    RB_CHECK(drain());
    slot_bytes = std::max(slot_bytes, slotBytes);
    slots      = std::max(slots, nSlots);
    RB_CHECK(ring.reserve(slot_bytes * (size_t)slots));
    while ((int)done.size() < slots) {
        cudaEvent_t e = nullptr;
        RB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        done.push_back(e);
    }
    delete[] busy;
    busy = new std::atomic<int>[slots];
    for (int i = 0; i < slots; ++i)
        busy[i].store(0);
    slotBytes = slot_bytes;
    nSlots    = slots;
    next      = 0;
    device    = dev;
    return RB_OK;
}

int HostStager::d2h(void* dst, const void* d_src, size_t bytes, cudaStream_t s) {
    for (size_t off = 0; off < bytes; off += slotBytes) {
        const size_t n = std::min(slotBytes, bytes - off);
        const int    k = next;
        next           = (next + 1) % nSlots;
        while (busy[k].load(std::memory_order_acquire))  // the worker is still copying this slot's previous content out
            std::this_thread::yield();
        unsigned char* slot = ring.p + (size_t)k * slotBytes;
        RB_CUDA(cudaMemcpyAsync(slot, (const unsigned char*)d_src + off, n, cudaMemcpyDeviceToHost, s));
        RB_CUDA(cudaEventRecord(done[k], s));
        busy[k].store(1, std::memory_order_release);
        copy_pool().submit(CopyJob{done[k], slot, (unsigned char*)dst + off, n, &busy[k], &failed, device});
    }
    return RB_OK;
}

int HostStager::drain() {
    for (int k = 0; k < nSlots; ++k)
        while (busy && busy[k].load(std::memory_order_acquire))
            std::this_thread::yield();
    if (failed.exchange(0)) {
        set_error("a device-to-host copy into the staging ring failed");
        return RB_ERR_CUDA;
    }
    return RB_OK;
}

static thread_local char t_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

const char* get_error() {
    return t_error;
}

std::atomic<uint64_t> g_launches{0};

int use_device(int device, DeviceInfo* info) {
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); librasr_b200 has no CPU path",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return RB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device ordinal %d out of range (0..%d)", device, n - 1);
        return RB_ERR_NO_DEVICE;
    }
    cudaDeviceProp p;
    RB_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d (%s) is sm_%d%d; librasr_b200 is built for sm_100a only", device, p.name, p.major,
                  p.minor);
        return RB_ERR_NO_DEVICE;
    }
    RB_CUDA(cudaSetDevice(device));
    if (info) {
        info->ordinal    = device;
        info->sm_count   = p.multiProcessorCount;
        info->cc_major   = p.major;
        info->cc_minor   = p.minor;
        info->smem_optin = p.sharedMemPerBlockOptin;
    }
    return RB_OK;
}

}  // namespace rb

extern "C" const char* rb_last_error(void) {
    return rb::get_error();
}

extern "C" const char* rb_version(void) {
    return "rasr_b200 0.1.0 (sm_100a)";
}

extern "C" int rb_device_count(void) {
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10)
            ++ok;
    }
    return ok;
}

extern "C" uint64_t rb_launch_count(void) {
    return rb::g_launches.load();
}

// ---------------------------------------------------------------------------------------------------------------
// page-locked host memory for the caller's feature / score / sample buffers (the host-buffer entry points accept any
// host pointer; from pageable memory every cudaMemcpyAsync is staged by the driver and synchronous)
extern "C" int rb_host_alloc(size_t bytes, void** out) {
    RB_REQUIRE(out != nullptr, "NULL output pointer");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        rb::set_error("no CUDA device available; librasr_b200 has no CPU path");
        return RB_ERR_NO_DEVICE;
    }
    const cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        rb::set_error("cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return RB_ERR_NOMEM;
    }
    return RB_OK;
}

extern "C" void rb_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess)
        cudaGetLastError();
}

extern "C" int rb_host_register(void* p, size_t bytes) {
    RB_REQUIRE(p != nullptr && bytes > 0, "bad argument");
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        rb::set_error("cudaHostRegister of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? RB_ERR_NOMEM : RB_ERR_CUDA;
    }
    return RB_OK;
}

extern "C" int rb_host_unregister(void* p) {
    RB_REQUIRE(p != nullptr, "bad argument");
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        rb::set_error("cudaHostUnregister failed: %s", cudaGetErrorString(e));
        return RB_ERR_CUDA;
    }
    return RB_OK;
}

// 1: page-locked (rb_host_alloc / rb_host_register / cudaHostAlloc), 0: pageable or unknown
extern "C" int rb_host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return a.type == cudaMemoryTypeHost ? 1 : 0;
}
