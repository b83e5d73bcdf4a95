// postproc.cu -- the feature post-processing nodes between the MFCC front-end and the scorers:
//   signal-normalization (mean / mean-and-variance, sliding window or whole segment)
//                                              src/Signal/Normalization.cc:41-190, src/Signal/SlidingWindow.hh:397-470
//   signal-vector-f32-sequence-concatenation   src/Signal/VectorSequenceConcatenation.hh:89-103 (DelayNode window,
//                                              margin policy copy)
//   signal-matrix-multiplication-f32           src/Signal/MatrixMult.hh, src/Math/Matrix.hh:487-494, Vector.hh:95-101
// i.e. cepstral mean (and variance) normalisation, the +-k frame splice and the LDA matrix of lda.flow:11-19 /
// processing.standard_system.flow:25-27.  Features never leave HBM between the stages.
//
// Kernel 1 (normalize_kernel): one thread per (segment, dimension) walks the frames in order and keeps the running
// f64 sums exactly as Normalization::update does (add the new frame, subtract the one pushed out of the window,
// re-derive mean / standard deviation whenever the statistics changed, freeze them during the flush), so the output
// is bit-identical to the CPU path for any window.  The walk is inherently sequential per (segment, dimension);
// segments x dimensions provide the parallelism.
// Kernel 2 (splice_matmul_kernel): y[t][n] = sum_k M[n][k] * S[t][k] with S the spliced (edge-replicated) window,
// gathered on the fly into shared memory; 64 x 64 output tiles, 4 x 4 outputs per thread, every output accumulates
// in ascending k in f32 (fused or not, as the reference build does) -- bit-identical to the sequential dot product.
#include <cmath>

#include "common.cuh"

namespace {

struct NormParams {
    const float*   in;
    float*         out;
    const int64_t* frameOff;  // [U+1] device
    int            nUtt, dim, type;
    long           L, R;
};

template<bool FUSE>
__global__ void __launch_bounds__(64) normalize_kernel(const NormParams p) {
    const int u = blockIdx.x, d = blockIdx.y * blockDim.x + threadIdx.x;
    if (d >= p.dim)
        return;
    const long   a = (long)p.frameOff[u], T = (long)p.frameOff[u + 1] - a;
    const float* x = p.in + a * p.dim + d;
    float*       y = p.out + a * p.dim + d;
    const long   D = p.dim;
    double       sum = 0.0, sumSq = 0.0, w = 0.0;
    float        mean = 0.0f, sd = 1.0f;
    bool         changed = true;
    auto finalize = [&]() {
        if (w > 0) {
            mean = (float)__ddiv_rn(sum, w);
            if (p.type == 2) {
                sd = (float)__dsqrt_rn(__ddiv_rn(__dsub_rn(sumSq, __ddiv_rn(__dmul_rn(sum, sum), w)), w));
                if (sd == 0.0f)
                    sd = 1.0f;
            }
        }
        changed = false;
    };
    auto apply = [&](long t) {
        if (changed)
            finalize();
        float v = __fsub_rn(x[t * D], mean);
        if (p.type == 2)
            v = __fdiv_rn(v, sd);
        y[t * D] = v;
    };
    long emitted = 0;
    for (long i = 0; i < T; ++i) {
        const double v = (double)x[i * D];
        sum            = __dadd_rn(sum, v);
        if (p.type == 2)
            sumSq = FUSE ? __fma_rn(v, v, sumSq) : __dadd_rn(sumSq, __dmul_rn(v, v));
        if (i >= p.L) {
            const double r = (double)x[(i - p.L) * D];
            sum            = __dsub_rn(sum, r);
            if (p.type == 2)
                sumSq = FUSE ? __fma_rn(-r, r, sumSq) : __dsub_rn(sumSq, __dmul_rn(r, r));
        }
        else
            w += 1.0;
        changed = true;
        if (i >= p.R) {
            apply(i - p.R);
            emitted = i - p.R + 1;
        }
    }
    for (long t = emitted; t < T; ++t)
        apply(t);
}

struct MatParams {
    const float*   in;   // [T * dim] (normalised) frames
    float*         out;  // [T * rows]
    const float*   M;    // [rows * K] row-major; null: splice only
    const int64_t* frameOff;
    const int*     tileUtt;  // utterance of each 64-frame tile
    const int64_t* tileT0;   // first frame (global) of each tile
    int            nTiles, dim, past, length, K, rows;
};

constexpr int kTile = 64, kKc = 16;

template<bool FUSE>
__global__ void __launch_bounds__(256) splice_matmul_kernel(const MatParams p) {
    __shared__ float sS[kKc][kTile + 4];  // [k][frame]
    __shared__ float sM[kKc][kTile + 4];  // [k][output]
    const int     tile = blockIdx.x, nb = blockIdx.y;
    const int     u    = p.tileUtt[tile];
    const int64_t a = p.frameOff[u], e = p.frameOff[u + 1], t0 = p.tileT0[tile];
    const int     tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx: outputs, ty: frames
    float         acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j] = 0.0f;
    for (int k0 = 0; k0 < p.K; k0 += kKc) {
        // gather the spliced window: element k of frame t is dim (k % D) of frame clamp(t - past + k / D)
        for (int idx = threadIdx.x; idx < kKc * kTile; idx += 256) {
            const int     f = idx / kKc, kk = idx - f * kKc, k = k0 + kk;
            const int64_t t = t0 + f;
            float         v = 0.0f;
            if (k < p.K && t < e) {
                const int     j = k / p.dim, dd = k - j * p.dim;
                int64_t       s = t - p.past + j;
                s               = s < a ? a : (s > e - 1 ? e - 1 : s);
                v               = p.in[s * p.dim + dd];
            }
            sS[kk][f] = v;
        }
        for (int idx = threadIdx.x; idx < kKc * kTile; idx += 256) {
            const int n = idx / kKc, kk = idx - n * kKc, k = k0 + kk, row = nb * kTile + n;
            sM[kk][n]   = (k < p.K && row < p.rows) ? p.M[(size_t)row * p.K + k] : 0.0f;
        }
        __syncthreads();
        const int kn = min(kKc, p.K - k0);
        for (int kk = 0; kk < kn; ++kk) {
            float xs[4], ms[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                xs[i] = sS[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                ms[j] = sM[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] = FUSE ? __fmaf_rn(ms[j], xs[i], acc[i][j]) : __fadd_rn(acc[i][j], __fmul_rn(ms[j], xs[i]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t t = t0 + ty * 4 + i;
        if (t >= e)
            continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = nb * kTile + tx * 4 + j;
            if (row < p.rows)
                p.out[t * p.rows + row] = acc[i][j];
        }
    }
}

// splice without a matrix: plain gather
__global__ void __launch_bounds__(256) splice_kernel(const MatParams p) {
    const int     tile = blockIdx.x;
    const int     u    = p.tileUtt[tile];
    const int64_t a = p.frameOff[u], e = p.frameOff[u + 1], t0 = p.tileT0[tile];
    for (int64_t idx = threadIdx.x; idx < (int64_t)kTile * p.K; idx += 256) {
        const int64_t f = idx / p.K, t = t0 + f;
        const int     k = (int)(idx - f * p.K);
        if (t >= e)
            break;
        const int j = k / p.dim, dd = k - j * p.dim;
        int64_t   s = t - p.past + j;
        s           = s < a ? a : (s > e - 1 ? e - 1 : s);
        p.out[t * p.K + k] = p.in[s * p.dim + dd];
    }
}

}  // namespace

struct rb_postproc {
    rb::DeviceInfo   dev;
    rb_postproc_cfg  cfg;
    int              dimIn = 0, dimMid = 0, dimOut = 0;  // input, after splice, output
    cudaStream_t     stream = nullptr;
    rb::DevBuf<float> dMatrix, dNorm, dIn, dOut;
    static constexpr int kSlots = 4;
    struct Slot {
        rb::PinnedBuf<char> host;
        rb::DevBuf<char>    dev;
        cudaEvent_t         ev = nullptr;
    } slots[kSlots];
    int nextSlot = 0;
    ~rb_postproc() {
        for (Slot& s : slots)
            if (s.ev) {
                cudaEventSynchronize(s.ev);
                cudaEventDestroy(s.ev);
            }
        if (stream)
            cudaStreamDestroy(stream);
    }
};

extern "C" int rb_postproc_create(const rb_postproc_cfg* cfg, int dim_in, rb_postproc** out) {
    RB_REQUIRE(cfg && out, "NULL argument");
    *out = nullptr;
    RB_REQUIRE(dim_in >= 1, "input dimension must be positive");
    RB_REQUIRE(cfg->norm_type >= 0 && cfg->norm_type <= 2, "unknown normalization type %d", cfg->norm_type);
    if (cfg->norm_type) {
        const long INF = 2147483647L;
        long       L = cfg->norm_length < 0 ? INF : cfg->norm_length, R = cfg->norm_right < 0 ? INF : cfg->norm_right;
        if (L >= INF && R >= INF)
            --R;
        RB_REQUIRE(L > R, "normalization: cannot initialize with length (%ld) and right (%ld)", cfg->norm_length,
                   cfg->norm_right);
    }
    const bool splice = cfg->splice_length > 0;
    if (splice)
        RB_REQUIRE(cfg->splice_right >= 0 && cfg->splice_length > cfg->splice_right,
                   "sequence concatenation: max-size (%d) must exceed right (%d)", cfg->splice_length,
                   cfg->splice_right);
    const int dimMid = dim_in * (splice ? cfg->splice_length : 1);
    if (cfg->matrix)
        RB_REQUIRE(cfg->matrix_cols == dimMid && cfg->matrix_rows >= 1,
                   "matrix has %d columns, the input vectors have dimension %d", cfg->matrix_cols, dimMid);
    rb_postproc* h = new (std::nothrow) rb_postproc();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    auto fail = [&](int code) {
        delete h;
        return code;
    };
    h->cfg    = *cfg;
    h->dimIn  = dim_in;
    h->dimMid = dimMid;
    h->dimOut = cfg->matrix ? cfg->matrix_rows : dimMid;
    int rc    = rb::use_device(cfg->device, &h->dev);
    if (rc != RB_OK)
        return fail(rc);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rb::set_error("cudaStreamCreate failed");
        return fail(RB_ERR_CUDA);
    }
    if (cfg->matrix) {
        if (h->dMatrix.upload(cfg->matrix, (size_t)cfg->matrix_rows * cfg->matrix_cols, h->stream) != RB_OK ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) {
            rb::set_error("matrix upload failed");
            return fail(RB_ERR_CUDA);
        }
    }
    h->cfg.matrix = nullptr;  // the caller's memory is not kept
    *out          = h;
    return RB_OK;
}

extern "C" void rb_postproc_destroy(rb_postproc* h) {
    if (!h)
        return;
    cudaSetDevice(h->dev.ordinal);
    delete h;
}

extern "C" int rb_postproc_dim_out(const rb_postproc* h) {
    return h ? h->dimOut : 0;
}

extern "C" int rb_postproc_process_dev(rb_postproc* h, const float* d_feats, const int64_t* frame_offsets, int n_utt,
                                       float* d_out, void* stream) {
    RB_REQUIRE(h && frame_offsets && n_utt >= 0, "bad argument");
    if (n_utt == 0)
        return RB_OK;
    const int64_t base = frame_offsets[0], T = frame_offsets[n_utt] - base;
    RB_REQUIRE(T >= 0, "negative frame count");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(d_feats && d_out, "NULL device buffer");
    RB_REQUIRE(d_feats != d_out, "input and output must not alias");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    const bool   hasNorm = h->cfg.norm_type != 0, hasSplice = h->cfg.splice_length > 0, hasMat = h->dMatrix.p != nullptr;

    // staging: [frame offsets | tile utterances | tile first frames]
    size_t nTiles = 0;
    for (int u = 0; u < n_utt; ++u) {
        RB_REQUIRE(frame_offsets[u + 1] >= frame_offsets[u], "frame offsets not monotone at utterance %d", u);
        nTiles += (size_t)(frame_offsets[u + 1] - frame_offsets[u] + kTile - 1) / kTile;
    }
    rb_postproc::Slot& slot = h->slots[h->nextSlot];
    h->nextSlot             = (h->nextSlot + 1) % rb_postproc::kSlots;
    if (!slot.ev)
        RB_CUDA(cudaEventCreateWithFlags(&slot.ev, cudaEventDisableTiming));
    else
        RB_CUDA(cudaEventSynchronize(slot.ev));
    const size_t offBytes = sizeof(int64_t) * (size_t)(n_utt + 1), t0Bytes = sizeof(int64_t) * nTiles;
    const size_t bytes    = offBytes + t0Bytes + sizeof(int) * nTiles;
    RB_CHECK(slot.host.reserve(bytes));
    RB_CHECK(slot.dev.reserve(bytes));
    int64_t* fo = reinterpret_cast<int64_t*>(slot.host.p);
    int64_t* t0 = reinterpret_cast<int64_t*>(slot.host.p + offBytes);
    int*     tu = reinterpret_cast<int*>(slot.host.p + offBytes + t0Bytes);
    size_t   ti = 0;
    for (int u = 0; u <= n_utt; ++u)
        fo[u] = frame_offsets[u] - base;
    for (int u = 0; u < n_utt; ++u)
        for (int64_t t = fo[u]; t < fo[u + 1]; t += kTile) {
            t0[ti] = t;
            tu[ti] = u;
            ++ti;
        }
    RB_CUDA(cudaMemcpyAsync(slot.dev.p, slot.host.p, bytes, cudaMemcpyHostToDevice, s));
    const int64_t* dFo = reinterpret_cast<const int64_t*>(slot.dev.p);

    const float* cur = d_feats;
    if (hasNorm) {
        float* dst = d_out;
        if (hasSplice || hasMat) {
            RB_CHECK(h->dNorm.reserve((size_t)T * h->dimIn));
            dst = h->dNorm.p;
        }
        const long INF = 2147483647L;
        NormParams np;
        np.in       = d_feats;
        np.out      = dst;
        np.frameOff = dFo;
        np.nUtt     = n_utt;
        np.dim      = h->dimIn;
        np.type     = h->cfg.norm_type;
        np.L        = h->cfg.norm_length < 0 ? INF : h->cfg.norm_length;
        np.R        = h->cfg.norm_right < 0 ? INF : h->cfg.norm_right;
        if (np.L >= INF && np.R >= INF)
            --np.R;  // SlidingWindow::init special case of collecting a whole segment
        const dim3 grid((unsigned)n_utt, (unsigned)((h->dimIn + 63) / 64));
        if (h->cfg.contraction)
            normalize_kernel<true><<<grid, 64, 0, s>>>(np);
        else
            normalize_kernel<false><<<grid, 64, 0, s>>>(np);
        RB_LAUNCH_CHECK();
        cur = dst;
    }
    if (hasSplice || hasMat) {
        MatParams mp;
        mp.in       = cur;
        mp.out      = d_out;
        mp.M        = h->dMatrix.p;
        mp.frameOff = dFo;
        mp.tileT0   = reinterpret_cast<const int64_t*>(slot.dev.p + offBytes);
        mp.tileUtt  = reinterpret_cast<const int*>(slot.dev.p + offBytes + t0Bytes);
        mp.nTiles   = (int)nTiles;
        mp.dim      = h->dimIn;
        mp.length   = hasSplice ? h->cfg.splice_length : 1;
        mp.past     = hasSplice ? h->cfg.splice_length - h->cfg.splice_right - 1 : 0;
        mp.K        = h->dimMid;
        mp.rows     = h->dimOut;
        if (hasMat) {
            const dim3 grid((unsigned)nTiles, (unsigned)((h->dimOut + kTile - 1) / kTile));
            if (h->cfg.contraction)
                splice_matmul_kernel<true><<<grid, 256, 0, s>>>(mp);
            else
                splice_matmul_kernel<false><<<grid, 256, 0, s>>>(mp);
        }
        else {
            splice_kernel<<<(unsigned)nTiles, 256, 0, s>>>(mp);
        }
        RB_LAUNCH_CHECK();
    }
    else if (!hasNorm) {
        RB_CUDA(cudaMemcpyAsync(d_out, d_feats, (size_t)T * h->dimIn * 4, cudaMemcpyDeviceToDevice, s));
    }
    RB_CUDA(cudaEventRecord(slot.ev, s));
    return RB_OK;
}

extern "C" int rb_postproc_process(rb_postproc* h, const float* feats, const int64_t* frame_offsets, int n_utt,
                                   float* out) {
    RB_REQUIRE(h && frame_offsets && n_utt >= 0, "bad argument");
    if (n_utt == 0)
        return RB_OK;
    const int64_t base = frame_offsets[0], T = frame_offsets[n_utt] - base;
    RB_REQUIRE(T >= 0, "negative frame count");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(feats && out, "NULL host buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    RB_CHECK(h->dIn.reserve((size_t)T * h->dimIn));
    RB_CHECK(h->dOut.reserve((size_t)T * h->dimOut));
    RB_CUDA(cudaMemcpyAsync(h->dIn.p, feats + base * h->dimIn, (size_t)T * h->dimIn * 4, cudaMemcpyHostToDevice,
                            h->stream));
    RB_CHECK(rb_postproc_process_dev(h, h->dIn.p, frame_offsets, n_utt, h->dOut.p, h->stream));
    RB_CUDA(cudaMemcpyAsync(out, h->dOut.p, (size_t)T * h->dimOut * 4, cudaMemcpyDeviceToHost, h->stream));
    RB_CUDA(cudaStreamSynchronize(h->stream));
    return RB_OK;
}
