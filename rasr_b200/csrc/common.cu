// common.cu -- error plumbing, device selection and the library-level entry points.
#include "common.cuh"

namespace rb {

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
