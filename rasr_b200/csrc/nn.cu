// nn.cu -- legacy feed-forward acoustic model (Nn::NeuralNetwork<f32>::forward) on sm_100a.
//
// Replaces, for a whole block of frames at once:
//   NeuralNetwork<T>::forward / forwardLayers      src/Nn/NeuralNetwork.cc:313-331,409-425
//   LinearLayer<T>::_forward (C = W^T X, + bias)   src/Nn/LinearLayer.cc:298-321  (cblas_sgemm / cublasSgemm)
//   sigmoid / ensureMinimalValue(0) / tanh / softmax  src/Math/FastMatrix.hh:802-836, CudaMatrixKernels.cu:176-294
//   Nn::BatchFeatureScorer score definition         src/Nn/BatchFeatureScorer.cc:148-171
//   BiasLayer::removeLogPriorFromBias               src/Nn/LinearLayer.cc:499-519
//
// The reference keeps activations as dim x T column-major matrices (a frame is a contiguous
// column) and weights as in x out column-major, multiplied with transposedA: both operands are
// therefore "K-major" rows (frame row x K, output-unit row x K) -- exactly the TN layout tcgen05
// wants, so no transposition is ever materialised.
//
// BF16 path: every affine layer is one launch of the tcgen05 GEMM (gemm_sm100.cuh) with the bias
// add, the activation and the bf16 re-quantisation for the next layer fused into the TMEM epilogue;
// the last layer writes f32 scores (negated, prior removed) straight from TMEM.
// F32 path: a CUDA-core tiled sgemm with the same fused epilogues, for 1e-4 parity with cblas_sgemm.
#include <cfloat>
#include <cmath>

#include "gemm_sm100.cuh"

namespace {

using namespace rbdev;

__device__ __forceinline__ float activate(float x, int act) {
    switch (act) {
        case RB_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));  // FastMatrix::sigmoid, gamma = 1
        case RB_ACT_RELU: return x < 0.0f ? 0.0f : x;          // ensureMinimalValue(0)
        case RB_ACT_TANH: return tanhf(x);
        default: return x;  // linear; softmax is a separate row pass
    }
}

// ---- tcgen05 epilogues -------------------------------------------------------------------
// bias of 32 consecutive columns: 8 vector loads issued together (the scalar form serialised 32 load latencies
// inside the exposed, single-buffered TMEM epilogue); columns >= N read as 0
__device__ __forceinline__ void load_bias32(const float* bias, int col0, int N, float (&b)[32]) {
    if (bias && col0 + 32 <= N && (((uintptr_t)(bias + col0)) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(bias + col0) + j);
            b[4 * j] = q.x; b[4 * j + 1] = q.y; b[4 * j + 2] = q.z; b[4 * j + 3] = q.w;
        }
    }
    else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            b[j] = (bias && col0 + j < N) ? __ldg(bias + col0 + j) : 0.0f;
    }
}

// v[j] = act(v[j] + b[j]) with the activation switch hoisted out of the element loop
__device__ __forceinline__ void bias_activate32(const float (&v)[32], const float (&b)[32], int act, float (&o)[32]) {
    if (act == RB_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float x = v[j] + b[j];
            o[j]          = x < 0.0f ? 0.0f : x;
        }
    }
    else if (act == RB_ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            o[j] = 1.0f / (1.0f + expf(-(v[j] + b[j])));
    }
    else if (act == RB_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            o[j] = tanhf(v[j] + b[j]);
    }
    else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            o[j] = v[j] + b[j];
    }
}

struct EpiHiddenBf16 {  // bf16 activations for the next layer, row pitch ldo (multiple of 64)
    const float*   bias;
    __nv_bfloat16* out;
    int            ldo, N, act;
    struct State {};
    __device__ void begin(State&, int) const {}
    __device__ void chunk(State&, int row, int col0, const float (&v)[32]) const {
        if (col0 >= ldo)
            return;
        float b[32], o[32];
        load_bias32(bias, col0, N, b);
        bias_activate32(v, b, act, o);
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(col0 + j < N ? o[j] : 0.0f, col0 + j + 1 < N ? o[j + 1] : 0.0f);
            packed[j >> 1]         = *reinterpret_cast<const uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)row * ldo + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    }
};

struct EpiFinalF32 {  // f32 output [M x N], out = sign * act(acc + bias)
    const float* bias;
    float*       out;
    int          ldo, N, act;
    float        sign;
    struct State {};
    __device__ void begin(State&, int) const {}
    // TMA-store path of gemm16_kernel: values only, the kernel stages and stores the 32 x 32 box
    __device__ void transform(int col0, const float (&v)[32], float (&o)[32]) const {
        float b[32];
        load_bias32(bias, col0, N, b);
        bias_activate32(v, b, act, o);
#pragma unroll
        for (int j = 0; j < 32; ++j)
            o[j] = sign * o[j];
    }
    __device__ void chunk(State&, int row, int col0, const float (&v)[32]) const {
        if (col0 >= N)
            return;
        float b[32], o[32];
        load_bias32(bias, col0, N, b);
        bias_activate32(v, b, act, o);
        float* dst = out + (size_t)row * ldo + col0;
        if (col0 + 32 <= N && ((ldo & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) =
                        make_float4(sign * o[j], sign * o[j + 1], sign * o[j + 2], sign * o[j + 3]);
        }
        else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < N)
                    dst[j] = sign * o[j];
        }
    }
};

// ---- f32 -> bf16 with zero padding of K ---------------------------------------------------
__global__ void convert_pad_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long rows,
                                        int cols, int ldo) {
    const long total = rows * (long)(ldo / 2);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / (ldo / 2);
        const int  c = (int)(i - r * (ldo / 2)) * 2;
        const float a = c < cols ? in[r * cols + c] : 0.0f;
        const float b = c + 1 < cols ? in[r * cols + c + 1] : 0.0f;
        reinterpret_cast<__nv_bfloat162*>(out)[i] = __floats2bfloat162_rn(a, b);
    }
}

// ---- CUDA-core sgemm: C[M x N] = A[M x K] * B[N x K]^T, 128x128 tile, 8x8 per thread -------
constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16;

__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ A, int lda,
                                                       const float* __restrict__ B, int ldb, int M, int N, int K,
                                                       const float* __restrict__ bias, float* __restrict__ C, int ldc,
                                                       int act, float sign) {
    __shared__ float As[SG_BK][SG_BM + 4];
    __shared__ float Bs[SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    float     acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            acc[i][j] = 0.0f;
    for (int k0 = 0; k0 < K; k0 += SG_BK) {
        // 128 rows x 16 k: thread loads 8 elements of A and of B (coalesced along k)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * 256;
            const int r = e >> 4, k = e & 15;
            const int gm = m0 + r, gn = n0 + r, gk = k0 + k;
            As[k][r] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : 0.0f;
            Bs[k][r] = (gn < N && gk < K) ? B[(size_t)gn * ldb + gk] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            float a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = As[k][ty * 8 + i];
                b[i] = Bs[k][tx * 8 + i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + ty * 8 + i;
        if (gm >= M)
            continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int gn = n0 + tx * 8 + j;
            if (gn < N)
                C[(size_t)gm * ldc + gn] = sign * activate(acc[i][j] + (bias ? bias[gn] : 0.0f), act);
        }
    }
}

// ---- row softmax (FastMatrix::softmax: subtract the column max, exp, divide by the sum) ---
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, long rows, int n) {
    __shared__ float red[8];
    for (long r = blockIdx.x; r < rows; r += gridDim.x) {
        float* row = x + r * (long)n;
        float  mx  = -FLT_MAX;
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            mx = fmaxf(mx, row[i]);
        for (int o = 16; o; o >>= 1)
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0)
            red[threadIdx.x >> 5] = mx;
        __syncthreads();
        mx = red[0];
        for (int w = 1; w < 8; ++w)
            mx = fmaxf(mx, red[w]);
        __syncthreads();
        float sum = 0.0f;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float e = expf(row[i] - mx);
            row[i]        = e;
            sum += e;
        }
        for (int o = 16; o; o >>= 1)
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if ((threadIdx.x & 31) == 0)
            red[threadIdx.x >> 5] = sum;
        __syncthreads();
        sum = 0.0f;
        for (int w = 0; w < 8; ++w)
            sum += red[w];
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            row[i] = row[i] / sum;
    }
}

}  // namespace

// ==========================================================================================
// host side
// ==========================================================================================

struct NnLayer {
    int in = 0, out = 0, act = 0;
    int kPad = 0;  // K padded to a multiple of 64 (bf16 path)
    rb::DevBuf<__nv_bfloat16> wBf16;  // [out x kPad]
    rb::DevBuf<float>         wF32;   // [out x in]
    rb::DevBuf<float>         bias;   // [out]
    rb::DevBuf<float>         biasScore;  // top layer only: bias - priorScale * logPrior
    CUtensorMap               mapW;
    CUtensorMap               mapW128;  // the same weights with 128-row boxes: hidden layers of small batches
};

struct rb_nn {
    rb::DeviceInfo       dev;
    int                  precision = RB_NN_BF16;
    int                  nLayers = 0;
    std::vector<NnLayer*> layers;
    cudaStream_t         stream = nullptr;
    long                 chunk = 18944;  // frames per pass: 74 row blocks of 256 -> whole waves on 148 SMs
    // bf16 path: ping-pong activation buffers [chunk x maxKPad] and their TMA maps (128-row boxes) per layer input
    std::vector<CUtensorMap>  mapIn128;
    rb::DevBuf<__nv_bfloat16> actA, actB;
    // f32 path
    rb::DevBuf<float> actFA, actFB;
    // host-pointer staging
    rb::DevBuf<float> dIn, dOut;
    // optional Nn::ClassLabelWrapper mapping of the scores (rb_nn_set_class_mapping)
    int               nClasses = 0;
    rb::DevBuf<int>   dClassMap;
    rb::DevBuf<float> dUnmapped;

    ~rb_nn() {
        for (NnLayer* l : layers)
            delete l;
        if (stream)
            cudaStreamDestroy(stream);
    }
};

namespace {

int forward_chunk(rb_nn* h, const float* dFeats, long T, float* dOut, bool scoreMode, cudaStream_t s) {
    const int L = h->nLayers;
    if (h->precision == RB_NN_BF16) {
        // input -> bf16 with K padding
        NnLayer* l0 = h->layers[0];
        {
            const long pairs  = T * (long)(l0->kPad / 2);
            const int  blocks = (int)std::min<long>((pairs + 255) / 256, (long)h->dev.sm_count * 16);
            convert_pad_bf16_kernel<<<blocks, 256, 0, s>>>(dFeats, h->actA.p, T, l0->in, l0->kPad);
            RB_LAUNCH_CHECK();
        }
        __nv_bfloat16* cur = h->actA.p;
        __nv_bfloat16* nxt = h->actB.p;
        for (int l = 0; l < L; ++l) {
            NnLayer* ly   = h->layers[l];
            const bool last = l == L - 1;
            if (!last) {
                EpiHiddenBf16 epi;
                epi.bias = ly->bias.p;
                epi.out  = nxt;
                epi.ldo  = h->layers[l + 1]->kPad;
                epi.N    = ly->out;
                epi.act  = ly->act;
                // a segment-sized batch (1000 frames) makes 8 x 8 tiles of 128 x 256 out of a 2048-wide layer: 64 CTAs for
                // 148 SMs.  With 128-column tiles there are twice as many (hidden layers 20 -> ~11 us each)
                const long tiles256 = ((T + rbgemm::BM - 1) / rbgemm::BM) * ((ly->out + rbgemm::BN - 1) / rbgemm::BN);
                if (tiles256 < h->dev.sm_count && getenv("RB_NN_WIDE_TILES") == nullptr)
                    RB_CHECK((rbgemm::launch<EpiHiddenBf16, 128>(h->mapIn128[l], ly->mapW128, (int)T, ly->out, ly->kPad,
                                                                 rbgemm::FMT_BF16, epi, h->dev.sm_count, s)));
                else
                    RB_CHECK(rbgemm::launch(h->mapIn128[l], ly->mapW, (int)T, ly->out, ly->kPad, rbgemm::FMT_BF16, epi,
                                            h->dev.sm_count, s));
                std::swap(cur, nxt);
            }
            else {
                EpiFinalF32 epi;
                epi.bias = scoreMode ? ly->biasScore.p : ly->bias.p;
                epi.out  = dOut;
                epi.ldo  = ly->out;
                epi.N    = ly->out;
                epi.act  = (scoreMode || ly->act == RB_ACT_SOFTMAX) ? RB_ACT_LINEAR : ly->act;
                epi.sign = scoreMode ? -1.0f : 1.0f;
                // the 128 x 256 kernel double-buffers its accumulator in TMEM: the epilogue of tile i (here 48 KB of
                // f32 scores per frame) overlaps the MMAs of tile i+1
                const bool tmaStore = (ly->out % 4) == 0 && ((uintptr_t)dOut % 16) == 0 &&
                                      getenv("RB_NN_NO_TMA_STORE") == nullptr;
                if (tmaStore) {  // wide f32 rows: whole-line stores by the TMA unit instead of row-strided STG
                    CUtensorMap mapOut;
                    RB_CHECK(rbgemm::make_map_out_f32(&mapOut, dOut, (uint64_t)T, (uint64_t)ly->out, (uint64_t)ly->out));
                    RB_CHECK(rbgemm::launch_tma_store(h->mapIn128[l], ly->mapW, mapOut, (int)T, ly->out, ly->kPad,
                                                      rbgemm::FMT_BF16, epi, h->dev.sm_count, s));
                }
                else
                    RB_CHECK(rbgemm::launch(h->mapIn128[l], ly->mapW, (int)T, ly->out, ly->kPad, rbgemm::FMT_BF16, epi,
                                            h->dev.sm_count, s));
            }
        }
    }
    else {
        const float* cur = dFeats;
        float*       bufs[2] = {h->actFA.p, h->actFB.p};
        for (int l = 0; l < L; ++l) {
            NnLayer*   ly   = h->layers[l];
            const bool last = l == L - 1;
            float*     dst  = last ? dOut : bufs[l & 1];
            const int  act  = last ? ((scoreMode || ly->act == RB_ACT_SOFTMAX) ? RB_ACT_LINEAR : ly->act) : ly->act;
            const float* b  = (last && scoreMode) ? ly->biasScore.p : ly->bias.p;
            dim3 grid((ly->out + SG_BN - 1) / SG_BN, (unsigned)((T + SG_BM - 1) / SG_BM));
            sgemm_tn_kernel<<<grid, 256, 0, s>>>(cur, ly->in, ly->wF32.p, ly->in, (int)T, ly->out, ly->in, b, dst,
                                                 ly->out, act, (last && scoreMode) ? -1.0f : 1.0f);
            RB_LAUNCH_CHECK();
            cur = dst;
        }
    }
    NnLayer* top = h->layers[L - 1];
    if (!scoreMode && top->act == RB_ACT_SOFTMAX) {
        const int blocks = (int)std::min<long>(T, (long)h->dev.sm_count * 8);
        softmax_rows_kernel<<<blocks, 256, 0, s>>>(dOut, T, top->out);
        RB_LAUNCH_CHECK();
    }
    return RB_OK;
}

// Nn::ClassLabelWrapper (src/Nn/ClassLabelWrapper.cc:57-100): emission class -> network output, -1 = disregarded class,
// whose score is FLT_MAX (Nn::BatchFeatureScorer::ContextScorer::score, src/Nn/BatchFeatureScorer.cc:163-169)
__global__ void __launch_bounds__(256) nn_map_classes_kernel(const float* __restrict__ in, int nOut, const int* __restrict__ map,
                                                             int nClasses, long T, float* __restrict__ out) {
    const long total = T * nClasses;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long t = i / nClasses;
        const int  o = map[i - t * nClasses];
        out[i]       = o >= 0 ? in[t * nOut + o] : FLT_MAX;
    }
}

int run_dev(rb_nn* h, const float* dFeats, long T, float* dOut, bool scoreMode, cudaStream_t s) {
    const int  in = h->layers[0]->in, out = h->layers[h->nLayers - 1]->out;
    const bool mapped = scoreMode && h->nClasses > 0;
    if (mapped)
        RB_CHECK(h->dUnmapped.reserve((size_t)std::min(h->chunk, T) * out));
    for (long a = 0; a < T; a += h->chunk) {
        const long n = std::min(h->chunk, T - a);
        RB_CHECK(forward_chunk(h, dFeats + a * in, n, mapped ? h->dUnmapped.p : dOut + a * out, scoreMode, s));
        if (mapped) {
            const long total = n * h->nClasses;
            nn_map_classes_kernel<<<(int)std::min<long>((total + 255) / 256, (long)h->dev.sm_count * 16), 256, 0, s>>>(
                    h->dUnmapped.p, out, h->dClassMap.p, h->nClasses, n, dOut + a * h->nClasses);
            RB_LAUNCH_CHECK();
        }
    }
    return RB_OK;
}

int run_host(rb_nn* h, const float* feats, long T, float* outp, bool scoreMode) {
    RB_REQUIRE(h != nullptr, "nn handle is NULL");
    RB_REQUIRE(T >= 0, "negative frame count");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(feats && outp, "NULL host buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    const size_t in = h->layers[0]->in,
                 out = scoreMode && h->nClasses > 0 ? (size_t)h->nClasses : (size_t)h->layers[h->nLayers - 1]->out;
    // bounded staging: the score matrix of a long segment (48 KB / frame for 12k senones) is streamed
    const long slab = std::min<long>(T, 4 * h->chunk);
    RB_CHECK(h->dIn.reserve((size_t)slab * in));
    RB_CHECK(h->dOut.reserve((size_t)slab * out));
    for (long a = 0; a < T; a += slab) {
        const long n = std::min(slab, T - a);
        RB_CUDA(cudaMemcpyAsync(h->dIn.p, feats + a * in, (size_t)n * in * 4, cudaMemcpyHostToDevice, h->stream));
        RB_CHECK(run_dev(h, h->dIn.p, n, h->dOut.p, scoreMode, h->stream));
        RB_CUDA(cudaMemcpyAsync(outp + a * out, h->dOut.p, (size_t)n * out * 4, cudaMemcpyDeviceToHost, h->stream));
        RB_CUDA(cudaStreamSynchronize(h->stream));
    }
    return RB_OK;
}

}  // namespace

extern "C" int rb_nn_create(int n_layers, const int* dims, const int* act, const float* const* weights,
                            const float* const* bias, const float* log_prior, float prior_scale, int precision,
                            int device, rb_nn** out) {
    RB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RB_REQUIRE(n_layers >= 1 && dims && act && weights, "bad network description");
    RB_REQUIRE(precision == RB_NN_F32 || precision == RB_NN_BF16, "unknown precision %d", precision);
    for (int l = 0; l <= n_layers; ++l)
        RB_REQUIRE(dims[l] >= 1, "layer dimension %d is %d", l, dims[l]);
    for (int l = 0; l < n_layers; ++l) {
        RB_REQUIRE(weights[l] != nullptr, "weights of layer %d are NULL", l);
        RB_REQUIRE(act[l] >= RB_ACT_LINEAR && act[l] <= RB_ACT_TANH, "unknown activation %d in layer %d", act[l], l);
        if (act[l] == RB_ACT_SOFTMAX && l != n_layers - 1) {
            rb::set_error("softmax is only supported as the top-layer activation (layer %d)", l);
            return RB_ERR_UNSUPPORTED;
        }
    }
    rb_nn* h = new (std::nothrow) rb_nn();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    auto fail = [&](int code) {
        delete h;
        return code;
    };
    int rc = rb::use_device(device, &h->dev);
    if (rc != RB_OK)
        return fail(rc);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rb::set_error("cudaStreamCreate failed");
        return fail(RB_ERR_CUDA);
    }
    h->precision = precision;
    h->nLayers   = n_layers;
    h->chunk     = 256L * std::max(1, h->dev.sm_count / 2);  // 18944 frames on 148 SMs: whole waves of 128-row tiles
    int maxKPad = 0, maxDim = 0;
    for (int l = 0; l < n_layers; ++l) {
        NnLayer* ly = new NnLayer();
        h->layers.push_back(ly);
        ly->in   = dims[l];
        ly->out  = dims[l + 1];
        ly->act  = act[l];
        ly->kPad = (int)rb::round_up(ly->in, 64);
        maxKPad  = std::max(maxKPad, ly->kPad);
        maxDim   = std::max(maxDim, std::max(ly->in, ly->out));
        std::vector<float> b(ly->out, 0.0f);
        if (bias && bias[l])
            std::copy(bias[l], bias[l] + ly->out, b.begin());
        if (ly->bias.upload(b, h->stream) != RB_OK)
            return fail(RB_ERR_CUDA);
        if (l == n_layers - 1) {
            std::vector<float> bs(b);
            if (log_prior && prior_scale != 0.0f)
                for (int c = 0; c < ly->out; ++c)
                    bs[c] -= prior_scale * log_prior[c];  // CPU branch of removeLogPriorFromBias
            if (ly->biasScore.upload(bs, h->stream) != RB_OK)
                return fail(RB_ERR_CUDA);
        }
        if (precision == RB_NN_BF16) {
            std::vector<__nv_bfloat16> w((size_t)ly->out * ly->kPad, __float2bfloat16(0.0f));
            for (int o = 0; o < ly->out; ++o)
                for (int i = 0; i < ly->in; ++i)
                    w[(size_t)o * ly->kPad + i] = __float2bfloat16(weights[l][(size_t)o * ly->in + i]);
            if (ly->wBf16.upload(w.data(), w.size(), h->stream) != RB_OK)
                return fail(RB_ERR_CUDA);
            if (cudaStreamSynchronize(h->stream) != cudaSuccess) {  // w is a local staging vector
                rb::set_error("weight upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                return fail(RB_ERR_CUDA);
            }
            rc = rbgemm::make_map(&ly->mapW, ly->wBf16.p, (uint64_t)ly->out, (uint64_t)ly->kPad, (uint64_t)ly->kPad,
                                  rbgemm::BN, true);
            if (rc == RB_OK)
                rc = rbgemm::make_map(&ly->mapW128, ly->wBf16.p, (uint64_t)ly->out, (uint64_t)ly->kPad, (uint64_t)ly->kPad,
                                      128, true);
            if (rc != RB_OK)
                return fail(rc);
        }
        else {
            if (ly->wF32.upload(weights[l], (size_t)ly->out * ly->in, h->stream) != RB_OK)
                return fail(RB_ERR_CUDA);
        }
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
            rb::set_error("parameter upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(RB_ERR_CUDA);
        }
    }
    if (precision == RB_NN_BF16) {
        if (h->actA.reserve((size_t)h->chunk * maxKPad) != RB_OK || h->actB.reserve((size_t)h->chunk * maxKPad) != RB_OK)
            return fail(RB_ERR_NOMEM);
        cudaMemset(h->actA.p, 0, (size_t)h->chunk * maxKPad * 2);
        cudaMemset(h->actB.p, 0, (size_t)h->chunk * maxKPad * 2);
        // padding columns of the first layer's input are written by the converter; hidden activations
        // are written up to the next layer's kPad by the epilogue
        h->mapIn128.resize(n_layers);
        for (int l = 0; l < n_layers; ++l) {
            const __nv_bfloat16* buf = (l % 2 == 0) ? h->actA.p : h->actB.p;
            rc = rbgemm::make_map(&h->mapIn128[l], buf, (uint64_t)h->chunk, (uint64_t)h->layers[l]->kPad,
                                  (uint64_t)h->layers[l]->kPad, rbgemm::BM, true);
            if (rc != RB_OK)
                return fail(rc);
        }
    }
    else {
        if (h->actFA.reserve((size_t)h->chunk * maxDim) != RB_OK || h->actFB.reserve((size_t)h->chunk * maxDim) != RB_OK)
            return fail(RB_ERR_NOMEM);
    }
    *out = h;
    return RB_OK;
}

extern "C" void rb_nn_destroy(rb_nn* h) {
    if (!h)
        return;
    cudaSetDevice(h->dev.ordinal);
    delete h;
}

extern "C" int rb_nn_n_outputs(const rb_nn* h) {
    return h ? h->layers[h->nLayers - 1]->out : 0;
}

extern "C" int rb_nn_n_emissions(const rb_nn* h) {
    return h ? (h->nClasses > 0 ? h->nClasses : h->layers[h->nLayers - 1]->out) : 0;
}

extern "C" int rb_nn_set_class_mapping(rb_nn* h, int n_classes, const int32_t* class_to_output) {
    RB_REQUIRE(h != nullptr, "nn handle is NULL");
    if (!class_to_output || n_classes <= 0) {
        h->nClasses = 0;
        return RB_OK;
    }
    const int nOut = h->layers[h->nLayers - 1]->out;
    for (int c = 0; c < n_classes; ++c)
        RB_REQUIRE(class_to_output[c] >= -1 && class_to_output[c] < nOut, "class %d maps to output %d of %d", c,
                   class_to_output[c], nOut);
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    RB_CHECK(h->dClassMap.upload(class_to_output, (size_t)n_classes, h->stream));
    RB_CUDA(cudaStreamSynchronize(h->stream));
    h->nClasses = n_classes;
    return RB_OK;
}

extern "C" int rb_nn_n_inputs(const rb_nn* h) {
    return h ? h->layers[0]->in : 0;
}

extern "C" int rb_nn_score_dev(rb_nn* h, const float* d_feats, long T, float* d_scores, void* stream) {
    RB_REQUIRE(h && T >= 0, "bad argument");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(d_feats && d_scores, "NULL device buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    return run_dev(h, d_feats, T, d_scores, true, stream ? (cudaStream_t)stream : h->stream);
}

extern "C" int rb_nn_forward_dev(rb_nn* h, const float* d_feats, long T, float* d_out, void* stream) {
    RB_REQUIRE(h && T >= 0, "bad argument");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(d_feats && d_out, "NULL device buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    return run_dev(h, d_feats, T, d_out, false, stream ? (cudaStream_t)stream : h->stream);
}

extern "C" int rb_nn_score(rb_nn* h, const float* feats, long T, float* scores) {
    return run_host(h, feats, T, scores, true);
}

extern "C" int rb_nn_forward(rb_nn* h, const float* feats, long T, float* out) {
    return run_host(h, feats, T, out, false);
}

// ---- test hook: one tcgen05 GEMM -----------------------------------------------------------
extern "C" int rb_test_gemm_bf16(const float* a, const float* b, const float* bias, int M, int N, int K, int act,
                                 float* d, int device) {
    RB_REQUIRE(a && b && d && M > 0 && N > 0 && K > 0, "bad argument");
    rb::DeviceInfo dev;
    RB_CHECK(rb::use_device(device, &dev));
    const int kPad = (int)rb::round_up(K, 64);
    rb::DevBuf<float>         dA32, dB32, dBias, dD;
    rb::DevBuf<__nv_bfloat16> dA, dB;
    cudaStream_t              s = nullptr;
    RB_CUDA(cudaStreamCreate(&s));
    int rc = RB_OK;
    do {
        if ((rc = dA32.upload(a, (size_t)M * K, s)) != RB_OK) break;
        if ((rc = dB32.upload(b, (size_t)N * K, s)) != RB_OK) break;
        if (bias && (rc = dBias.upload(bias, N, s)) != RB_OK) break;
        if ((rc = dA.reserve((size_t)M * kPad)) != RB_OK) break;
        if ((rc = dB.reserve((size_t)N * kPad)) != RB_OK) break;
        if ((rc = dD.reserve((size_t)M * N)) != RB_OK) break;
        convert_pad_bf16_kernel<<<256, 256, 0, s>>>(dA32.p, dA.p, M, K, kPad);
        convert_pad_bf16_kernel<<<256, 256, 0, s>>>(dB32.p, dB.p, N, K, kPad);
        rb::count_launch(2);
        CUtensorMap mA, mB;
        if ((rc = rbgemm::make_map(&mA, dA.p, M, kPad, kPad, rbgemm::BM, true)) != RB_OK) break;
        if ((rc = rbgemm::make_map(&mB, dB.p, N, kPad, kPad, rbgemm::BN, true)) != RB_OK) break;
        EpiFinalF32 epi;
        epi.bias = bias ? dBias.p : nullptr;
        epi.out  = dD.p;
        epi.ldo  = N;
        epi.N    = N;
        epi.act  = act;
        epi.sign = 1.0f;
        rc = rbgemm::launch(mA, mB, M, N, kPad, rbgemm::FMT_BF16, epi, dev.sm_count, s);
        if (rc != RB_OK) break;
        cudaError_t e = cudaMemcpyAsync(d, dD.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            rb::set_error("tcgen05 GEMM failed: %s", cudaGetErrorString(e));
            rc = RB_ERR_CUDA;
        }
    } while (0);
    cudaStreamDestroy(s);
    return rc;
}

// ---- test hook: time one GEMM variant on device-resident operands (bf16 hidden-layer epilogue) -----
extern "C" int rb_test_gemm_bench(int M, int N, int K, int variant, int iters, float* ms_per_iter, int device) {
    RB_REQUIRE(M > 0 && N > 0 && K > 0 && iters > 0 && ms_per_iter, "bad argument");
    rb::DeviceInfo dev;
    RB_CHECK(rb::use_device(device, &dev));
    const int kPad = (int)rb::round_up(K, 64), nPad = (int)rb::round_up(N, 64);
    rb::DevBuf<__nv_bfloat16> dA, dB, dC;
    rb::DevBuf<float>         dBias;
    RB_CHECK(dA.reserve((size_t)M * kPad));
    RB_CHECK(dB.reserve((size_t)N * kPad));
    RB_CHECK(dC.reserve((size_t)M * nPad));
    RB_CHECK(dBias.reserve(N));
    // 0x3c00 bit pattern = 0.0078 in bf16: finite, non-trivial operands
    RB_CUDA(cudaMemset(dA.p, 0x3c, (size_t)M * kPad * 2));
    RB_CUDA(cudaMemset(dB.p, 0x3c, (size_t)N * kPad * 2));
    RB_CUDA(cudaMemset(dBias.p, 0, (size_t)N * 4));
    CUtensorMap mA, mB;
    RB_REQUIRE(variant == 0, "GEMM variant %d does not exist (0: 128 x 256 tile, double-buffered TMEM)", variant);
    RB_CHECK(rbgemm::make_map(&mA, dA.p, M, kPad, kPad, rbgemm::BM, true));
    RB_CHECK(rbgemm::make_map(&mB, dB.p, N, kPad, kPad, rbgemm::BN, true));
    EpiHiddenBf16 epi;
    epi.bias = dBias.p;
    epi.out  = dC.p;
    epi.ldo  = nPad;
    epi.N    = N;
    epi.act  = RB_ACT_RELU;
    cudaStream_t s = nullptr;
    RB_CUDA(cudaStreamCreate(&s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int rc = RB_OK;
    for (int i = 0; i < iters + 2 && rc == RB_OK; ++i) {
        if (i == 2)
            cudaEventRecord(e0, s);
        rc = rbgemm::launch(mA, mB, M, N, kPad, rbgemm::FMT_BF16, epi, dev.sm_count, s);
    }
    cudaEventRecord(e1, s);
    cudaError_t e = cudaStreamSynchronize(s);
    float       ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    if (rc != RB_OK)
        return rc;
    if (e != cudaSuccess) {
        rb::set_error("GEMM benchmark failed: %s", cudaGetErrorString(e));
        return RB_ERR_CUDA;
    }
    *ms_per_iter = ms / iters;
    return RB_OK;
}
