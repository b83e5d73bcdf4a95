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
