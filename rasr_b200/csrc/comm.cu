// comm.cu -- exchange of the score matrix between the GPUs of one box.
//
// The reference has no collective at all: its only parallelism is N independent processes over corpus partitions
// (src/Bliss/CorpusDescription.cc:173-180).  The exchange exists for ONE case -- a single decoder rank consumes the
// frames of every shard -- and is ours to define:
//
//   * every rank owns a WINDOW in its HBM that holds the gathered matrix [rows x row_len] f32; the windows are mapped
//     into every process of the box (CUDA IPC), so a kernel on GPU a can store into GPU b's window over NVLink /
//     NVSwitch.  Two ways to fill it:
//       - fused: the scorer's epilogue writes there itself -- rb_comm_window_ptr() of the consumer rank, offset by this
//         rank's first row, is passed as `d_scores` to rb_gmm_score_dev / rb_pipeline_score_dev / rb_nn_score_dev; the
//         scores cross NVLink tile by tile while the kernel is still computing and never exist in local HBM;
//       - push: gather_push_kernel copies this rank's rows with 128-bit loads and stores into the window of the root
//         (gather) or of every peer (all-gather).
//   * arrival is signalled on the device: comm_barrier_kernel publishes this rank's epoch into every peer's flag words
//     (system-scope release) and spins (bounded) until every peer has published its own -- no host round trip, the
//     consumer's kernels are simply enqueued behind it on the same stream.
//   * RB_COMM_NCCL is the library baseline for the same call: ncclAllGather when the shards are equal, else one
//     ncclBroadcast per rank in a group (no padding, no staging copy), on the caller's stream.  libnccl.so.2 is
//     resolved at run time (the copy torch already loaded, else the system one): the library does not link NCCL.
#include <dlfcn.h>

#include "common.cuh"

namespace {

constexpr int    kMaxWorld   = 16;
constexpr size_t kFlagBytes  = 4096;  // flag words in front of the window's data
constexpr int    kPushThreads = 256;

struct PushParams {
    const float* src;
    float*       dst[kMaxWorld];  // window data pointers (already offset to this rank's first row) of the targets
    int          nDst;
    size_t       n;  // floats
};

// every element is read once from local HBM and stored to every target (remote stores are posted: no round trip)
__global__ void __launch_bounds__(kPushThreads) gather_push_kernel(const PushParams p) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    bool         vec = (p.n % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.src) & 15) == 0);
    for (int d = 0; d < p.nDst; ++d)
        vec = vec && ((reinterpret_cast<uintptr_t>(p.dst[d]) & 15) == 0);
    if (vec) {
        const float4* s  = reinterpret_cast<const float4*>(p.src);
        const size_t  n4 = p.n / 4;
        // 4 independent loads in flight per thread
        for (size_t i = tid; i < n4; i += 4 * nth) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k * nth < n4)
                    v[k] = __ldcs(s + i + k * nth);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k * nth < n4)
                    for (int d = 0; d < p.nDst; ++d)
                        reinterpret_cast<float4*>(p.dst[d])[i + k * nth] = v[k];
        }
    }
    else {
        for (size_t i = tid; i < p.n; i += nth) {
            const float v = p.src[i];
            for (int d = 0; d < p.nDst; ++d)
                p.dst[d][i] = v;
        }
    }
}

struct BarrierParams {
    unsigned long long* peerFlags[kMaxWorld];  // flag array of every rank's window, as mapped here
    int                 world, rank;
    unsigned long long  epoch;
};

// thread i: publish `epoch` in peer i's slot for this rank, then wait for peer i's epoch in the local slot
__global__ void comm_barrier_kernel(const BarrierParams p) {
    const int i = threadIdx.x;
    if (i >= p.world)
        return;
    __threadfence_system();  // everything this GPU stored before the barrier is visible before the flag
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peerFlags[i] + p.rank), "l"(p.epoch) : "memory");
    const unsigned long long* mine = p.peerFlags[p.rank] + i;
    long long                 t0   = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= p.epoch)
            break;
        if (clock64() - t0 > 20000000000ll) {  // ~10 s: a missing peer must surface as an error, not as a hung GPU
            printf("rasr_b200: rank %d gave up waiting for rank %d at barrier %llu (saw %llu)\n", p.rank, i, p.epoch, v);
            __trap();
        }
        __nanosleep(200);
    }
}

// ---- NCCL, resolved at run time ---------------------------------------------------------------
struct NcclId {
    char bytes[128];
};
typedef void* NcclComm;
struct NcclApi {
    void* so = nullptr;
    int (*GetUniqueId)(NcclId*)                                                                 = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int)                                            = nullptr;
    int (*CommDestroy)(NcclComm)                                                                = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t)                   = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t)              = nullptr;
    int (*GroupStart)()                                                                         = nullptr;
    int (*GroupEnd)()                                                                           = nullptr;
    const char* (*GetErrorString)(int)                                                          = nullptr;
    int (*GetVersion)(int*)                                                                     = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.so)
        return RB_OK;
    void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host process already uses (torch's)
    if (!so)
        so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) {
        rb::set_error("libnccl.so.2 cannot be loaded: %s", dlerror());
        return RB_ERR_UNSUPPORTED;
    }
    NcclApi a;
    a.so = so;
#define RB_SYM(field, name)                                              \
    *reinterpret_cast<void**>(&a.field) = dlsym(so, name);               \
    if (!a.field) {                                                      \
        rb::set_error("libnccl.so.2 does not export %s", name);         \
        return RB_ERR_UNSUPPORTED;                                       \
    }
    RB_SYM(GetUniqueId, "ncclGetUniqueId")
    RB_SYM(CommInitRank, "ncclCommInitRank")
    RB_SYM(CommDestroy, "ncclCommDestroy")
    RB_SYM(AllGather, "ncclAllGather")
    RB_SYM(Broadcast, "ncclBroadcast")
    RB_SYM(GroupStart, "ncclGroupStart")
    RB_SYM(GroupEnd, "ncclGroupEnd")
    RB_SYM(GetErrorString, "ncclGetErrorString")
    RB_SYM(GetVersion, "ncclGetVersion")
#undef RB_SYM
    g_nccl = a;
    return RB_OK;
}

#define RB_NCCL(call)                                                                                  \
    do {                                                                                               \
        int r__ = (call);                                                                              \
        if (r__ != 0) {                                                                                \
            rb::set_error("%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
            return RB_ERR_CUDA;                                                                        \
        }                                                                                              \
    } while (0)

}  // namespace

struct rb_comm {
    rb::DeviceInfo dev;
    int            world = 1, rank = 0;
    cudaStream_t   stream = nullptr;
    // window
    unsigned char*     local = nullptr;  // flags + data
    size_t             bytes = 0;        // data bytes
    unsigned char*     peer[kMaxWorld] = {};
    bool               attached = false;
    unsigned long long epoch = 0;
    NcclComm           nccl = nullptr;
    ~rb_comm() {
        if (nccl && g_nccl.CommDestroy)
            g_nccl.CommDestroy(nccl);
        for (int r = 0; r < world && r < kMaxWorld; ++r)
            if (peer[r] && r != rank)
                cudaIpcCloseMemHandle(peer[r]);
        if (local)
            cudaFree(local);
        if (stream)
            cudaStreamDestroy(stream);
        cudaGetLastError();
    }
};

extern "C" int rb_comm_create(int world, int rank, int device, rb_comm** out) {
    RB_REQUIRE(out != nullptr, "NULL output pointer");
    *out = nullptr;
    RB_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "bad world size %d / rank %d (at most %d ranks)",
               world, rank, kMaxWorld);
    rb::DeviceInfo dev;
    RB_CHECK(rb::use_device(device, &dev));
    rb_comm* c = new (std::nothrow) rb_comm();
    if (!c) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    c->dev   = dev;
    c->world = world;
    c->rank  = rank;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rb::set_error("cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return RB_ERR_CUDA;
    }
    *out = c;
    return RB_OK;
}

extern "C" void rb_comm_destroy(rb_comm* c) {
    if (!c)
        return;
    cudaSetDevice(c->dev.ordinal);
    delete c;
}

extern "C" int rb_comm_world(const rb_comm* c) {
    return c ? c->world : 0;
}
extern "C" int rb_comm_rank(const rb_comm* c) {
    return c ? c->rank : -1;
}

extern "C" int rb_comm_window_alloc(rb_comm* c, size_t bytes, void** d_window, void* handle) {
    RB_REQUIRE(c && d_window && handle, "NULL argument");
    RB_REQUIRE(!c->local, "the communicator already owns a window");
    RB_CUDA(cudaSetDevice(c->dev.ordinal));
    const size_t total = kFlagBytes + rb::round_up(std::max<size_t>(bytes, 16), 256);
    cudaError_t  e     = cudaMalloc((void**)&c->local, total);
    if (e != cudaSuccess) {
        c->local = nullptr;
        rb::set_error("cudaMalloc of the %zu-byte window failed: %s", total, cudaGetErrorString(e));
        return RB_ERR_NOMEM;
    }
    RB_CUDA(cudaMemset(c->local, 0, kFlagBytes));
    RB_CUDA(cudaDeviceSynchronize());
    c->bytes         = bytes;
    c->peer[c->rank] = c->local;
    static_assert(sizeof(cudaIpcMemHandle_t) == RB_COMM_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    if (c->world > 1) {
        e = cudaIpcGetMemHandle(&h, c->local);
        if (e != cudaSuccess) {
            cudaGetLastError();
            rb::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
            return RB_ERR_UNSUPPORTED;
        }
    }
    else
        memset(&h, 0, sizeof(h));
    memcpy(handle, &h, sizeof(h));
    *d_window = c->local + kFlagBytes;
    return RB_OK;
}

extern "C" int rb_comm_window_attach(rb_comm* c, const void* handles) {
    RB_REQUIRE(c && handles, "NULL argument");
    RB_REQUIRE(c->local, "rb_comm_window_alloc first");
    RB_REQUIRE(!c->attached, "the window is already attached");
    RB_CUDA(cudaSetDevice(c->dev.ordinal));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank)
            continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
        void*       p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            rb::set_error("cudaIpcOpenMemHandle for the window of rank %d failed: %s (no peer access between the devices?)",
                          r, cudaGetErrorString(e));
            return RB_ERR_UNSUPPORTED;
        }
        c->peer[r] = (unsigned char*)p;
    }
    c->attached = true;
    return RB_OK;
}

extern "C" int rb_comm_window_ptr(const rb_comm* c, int peer, void** d_ptr) {
    RB_REQUIRE(c && d_ptr, "NULL argument");
    RB_REQUIRE(peer >= 0 && peer < c->world, "rank %d out of range", peer);
    RB_REQUIRE(c->peer[peer] != nullptr, peer == c->rank ? "rb_comm_window_alloc first" : "rb_comm_window_attach first");
    *d_ptr = c->peer[peer] + kFlagBytes;
    return RB_OK;
}

extern "C" int rb_comm_barrier_dev(rb_comm* c, void* stream) {
    RB_REQUIRE(c != nullptr, "NULL communicator");
    RB_REQUIRE(c->local && (c->attached || c->world == 1), "the window is not attached");
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    BarrierParams p;
    for (int r = 0; r < c->world; ++r)
        p.peerFlags[r] = reinterpret_cast<unsigned long long*>(c->peer[r]);
    p.world = c->world;
    p.rank  = c->rank;
    p.epoch = ++c->epoch;
    comm_barrier_kernel<<<1, 32, 0, s>>>(p);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

extern "C" int rb_comm_nccl_unique_id(void* id) {
    RB_REQUIRE(id != nullptr, "NULL argument");
    RB_CHECK(load_nccl());
    NcclId u;
    RB_NCCL(g_nccl.GetUniqueId(&u));
    static_assert(sizeof(NcclId) == RB_COMM_ID_BYTES, "id size");
    memcpy(id, &u, sizeof(u));
    return RB_OK;
}

extern "C" int rb_comm_nccl_init(rb_comm* c, const void* id) {
    RB_REQUIRE(c && id, "NULL argument");
    RB_REQUIRE(!c->nccl, "NCCL is already initialised on this communicator");
    RB_CHECK(load_nccl());
    RB_CUDA(cudaSetDevice(c->dev.ordinal));
    NcclId u;
    memcpy(&u, id, sizeof(u));
    RB_NCCL(g_nccl.CommInitRank(&c->nccl, c->world, u, c->rank));
    return RB_OK;
}

extern "C" int rb_comm_nccl_version(void) {
    if (load_nccl() != RB_OK)
        return 0;
    int v = 0;
    g_nccl.GetVersion(&v);
    return v;
}

// floats [first, first + n) of the gathered matrix, read from d_send, stored into the window of `root` (>= 0) or of every rank
static int push_range(rb_comm* c, const float* d_send, size_t first, size_t n, int root, cudaStream_t s) {
    PushParams p;
    p.src  = d_send;
    p.n    = n;
    p.nDst = 0;
    for (int r = 0; r < c->world; ++r) {
        if (root >= 0 && r != root)
            continue;
        float* dst = reinterpret_cast<float*>(c->peer[r] + kFlagBytes) + first;
        if (r == c->rank && dst == d_send)
            continue;  // the scorer already wrote into the local window
        p.dst[p.nDst++] = dst;
    }
    if (p.nDst == 0 || n == 0)
        return RB_OK;
    const size_t work   = (n + 3) / 4;
    const int    blocks = (int)std::min<size_t>((work + kPushThreads * 4 - 1) / (kPushThreads * 4), (size_t)c->dev.sm_count * 4);
    gather_push_kernel<<<std::max(blocks, 1), kPushThreads, 0, s>>>(p);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

extern "C" int rb_comm_gather_scores_dev(rb_comm* c, const float* d_send, const int64_t* row_offsets, int row_len,
                                         int root, int transport, void* stream) {
    RB_REQUIRE(c && row_offsets, "NULL argument");
    RB_REQUIRE(row_len > 0, "row length must be positive");
    RB_REQUIRE(root < c->world, "root %d out of range", root);
    RB_REQUIRE(c->local && (c->attached || c->world == 1), "the window is not attached");
    for (int r = 0; r < c->world; ++r)
        RB_REQUIRE(row_offsets[r] <= row_offsets[r + 1], "row offsets must not decrease");
    RB_REQUIRE(row_offsets[0] == 0 && (size_t)row_offsets[c->world] * row_len * 4 <= c->bytes,
               "the gathered matrix (%lld rows x %d) does not fit the %zu-byte window", (long long)row_offsets[c->world],
               row_len, c->bytes);
    cudaStream_t s     = stream ? (cudaStream_t)stream : c->stream;
    const size_t first = (size_t)row_offsets[c->rank] * row_len;
    const size_t n     = (size_t)(row_offsets[c->rank + 1] - row_offsets[c->rank]) * row_len;
    float*       mine  = reinterpret_cast<float*>(c->local + kFlagBytes) + first;
    RB_REQUIRE(d_send != nullptr || n == 0, "NULL send buffer");

    if (transport == RB_COMM_NCCL) {
        RB_REQUIRE(c->nccl != nullptr, "rb_comm_nccl_init first");
        float* win = reinterpret_cast<float*>(c->local + kFlagBytes);
        if (d_send != mine && n)
            RB_CUDA(cudaMemcpyAsync(mine, d_send, n * 4, cudaMemcpyDeviceToDevice, s));
        bool equal = root < 0;
        for (int r = 1; r < c->world && equal; ++r)
            equal = row_offsets[r + 1] - row_offsets[r] == row_offsets[1];
        if (equal) {
            RB_NCCL(g_nccl.AllGather(mine, win, n, /*ncclFloat*/ 7, c->nccl, s));
        }
        else {
            // unequal shards: one in-place broadcast per rank, fused into one NCCL group; a gather to `root` has no
            // NCCL primitive of its own and is served by the same all-gather
            RB_NCCL(g_nccl.GroupStart());
            for (int r = 0; r < c->world; ++r) {
                float*       at = win + (size_t)row_offsets[r] * row_len;
                const size_t nr = (size_t)(row_offsets[r + 1] - row_offsets[r]) * row_len;
                if (nr)
                    RB_NCCL(g_nccl.Broadcast(at, at, nr, 7, r, c->nccl, s));
            }
            RB_NCCL(g_nccl.GroupEnd());
        }
        rb::count_launch();
        return RB_OK;
    }
    RB_REQUIRE(transport == RB_COMM_P2P, "unknown transport %d", transport);
    return push_range(c, d_send, first, n, root, s);
}

extern "C" int rb_comm_push_rows_dev(rb_comm* c, const float* d_send, int64_t first_row, int64_t n_rows, int row_len,
                                     int root, void* stream) {
    RB_REQUIRE(c != nullptr, "NULL communicator");
    RB_REQUIRE(row_len > 0 && first_row >= 0 && n_rows >= 0, "bad row range");
    RB_REQUIRE(root < c->world, "root %d out of range", root);
    RB_REQUIRE(c->local && (c->attached || c->world == 1), "the window is not attached");
    RB_REQUIRE((size_t)(first_row + n_rows) * row_len * 4 <= c->bytes, "rows [%lld, %lld) x %d do not fit the %zu-byte window",
               (long long)first_row, (long long)(first_row + n_rows), row_len, c->bytes);
    RB_REQUIRE(d_send != nullptr || n_rows == 0, "NULL send buffer");
    return push_range(c, d_send, (size_t)first_row * row_len, (size_t)n_rows * row_len, root,
                      stream ? (cudaStream_t)stream : c->stream);
}
