// frontend.cu -- the Flow MFCC pipeline as two frame-batched kernels (sm_100a).
//
// Replaces the per-frame pull chain of src/Tools/FeatureExtraction/share/mfcc.flow:8-34 and
// derivationWithRegression.flow:7-27:
//   Signal::Preemphasis::apply                         src/Signal/Preemphasis.cc:51-74
//   Signal::WindowBuffer get/flush + Window::transform  src/Signal/WindowBuffer.cc:50-126, Window.cc:84-96
//   HammingWindowFunction                               src/Signal/WindowFunction.cc:92-101
//   RealFastFourierTransform (zero pad, radix-2, split) src/Signal/FastFourierTransform.cc:51-100,
//                                                       src/Math/FastFourierTransform.cc:59-146
//   alternatingComplexVectorAmplitude                   src/Signal/ComplexVectorFunction.hh:29-45
//   FilterBank::apply (mel, triangular)                 src/Signal/Filterbank.cc:65-71,640-672
//   VectorLogFunction (log10)                           src/Flow/SimpleFunction.hh:39-48
//   CosineTransform::apply                              src/Signal/CosineTransform.cc:76-83
//   Delay + Regression (first/second order) + concat    src/Signal/Delay.cc:137-183, Regression.cc:25-63
//
// Kernel 1 (mfcc_static_kernel): persistent CTAs; a CTA takes a tile of 32 consecutive frames of one
// utterance, stages the 21 KB of samples the tile touches with one TMA bulk copy (the 2.5x frame
// overlap is served from shared memory, HBM sees each sample once), keeps window / twiddles / mel
// weights / DCT matrix in shared memory (one bulk copy per CTA), and each warp turns frames into
// cepstra entirely on chip: pre-emphasis+window on the fly, 256-point complex FFT in shared memory,
// real split, |.|, sparse mel taps, log10, DCT.  Only 13 floats per frame leave the SM.
// Kernel 2 (mfcc_derivative_kernel): delta / delta-delta over the +-2 frame window with edge
// replication, concatenated output, arithmetic in the reference's operation order.
#include <cmath>
#include <functional>

#include "common.cuh"
#include "internal.h"

namespace {

using namespace rbdev;

constexpr int kWarps      = 8;
constexpr int kThreads    = kWarps * 32;
constexpr int kTileFrames = 32;

struct Tile {
    int utt;  // utterance index
    int f0;   // first frame of the tile within the utterance
    int nf;   // frames in the tile (<= kTileFrames)
    int pad;
};

struct FeParams {
    const float*   samples;     // concatenated utterances
    // static features: segment u covers samples [sampleOff[u], sampleEnd[u]) and writes frames from frameOff[u] on.
    // Plain calls: segments = utterances, sampleEnd = sampleOff + 1.  With DC detection: segments = kept sample runs.
    const int64_t* sampleOff;
    const int64_t* sampleEnd;
    const int64_t* frameOff;
    const Tile*    tiles;
    int            nTiles;
    // derivatives / concatenation run over whole utterances (the delay node does not look at time stamps):
    // group g covers frames [groupFrameOff[g], groupFrameOff[g + 1]); same tables as above for plain calls
    const int64_t* groupFrameOff;
    const Tile*    groupTiles;
    int            nGroupTiles;
    const float*   tables;      // blob, see TableLayout
    int            tableFloats;
    // geometry
    int   L, S, N, nBins, nFilters, nCep, featDim, derivatives;
    float alpha, scale;
    // table offsets (in floats) inside the blob
    int oWindow, oTw, oTws, oFbStart, oFbEnd, oFbOff, oFbW, oDct;
    // outputs
    float* cep;        // [T * nCep]
    float* feats;      // [T * featDim]
    float* dbgAmp;     // [T * nBins] or null
    float* dbgFbank;   // [T * nFilters] or null
    int64_t totalSamples;
};

__device__ __forceinline__ int bitrev(int x, int bits) {
    return (int)(__brev((unsigned)x) >> (32 - bits));
}

// dynamic shared memory layout (floats):
//   tables[tableFloats] | samples[sampleCap] | per warp: z[N] amp[nBinsPad] fb[fbPad]
__global__ void __launch_bounds__(kThreads) mfcc_static_kernel(const FeParams p, int sampleCap, int nBinsPad,
                                                               int fbPad, int log2M) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar;

    float* sTab     = smem;
    float* sSamples = sTab + ((p.tableFloats + 3) & ~3);
    float* sWarp    = sSamples + sampleCap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int perWarp = p.N + nBinsPad + fbPad;
    float*  z   = sWarp + warp * perWarp;
    float*  amp = z + p.N;
    float*  fb  = amp + nBinsPad;
    float2* zc  = reinterpret_cast<float2*>(z);

    const float*  sWindow = sTab + p.oWindow;
    const float2* sTw     = reinterpret_cast<const float2*>(sTab + p.oTw);
    const float2* sTws    = reinterpret_cast<const float2*>(sTab + p.oTws);
    const int*    sFbStart = reinterpret_cast<const int*>(sTab + p.oFbStart);
    const int*    sFbEnd   = reinterpret_cast<const int*>(sTab + p.oFbEnd);
    const int*    sFbOff   = reinterpret_cast<const int*>(sTab + p.oFbOff);
    const float*  sFbW     = sTab + p.oFbW;
    const float*  sDct     = sTab + p.oDct;

    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)p.tableFloats * 4u;  // blob is padded to 16 bytes on the host
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(sTab, p.tables, bytes, &bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;

    const int M = p.N >> 1;  // complex points

    for (int tileIdx = blockIdx.x; tileIdx < p.nTiles; tileIdx += gridDim.x) {
        const Tile    tile   = p.tiles[tileIdx];
        const int64_t uBeg   = p.sampleOff[tile.utt];
        const int64_t uLen   = p.sampleEnd[tile.utt] - uBeg;
        const int64_t fOut   = p.frameOff[tile.utt] + tile.f0;
        // samples the tile touches, relative to the utterance: [s0 - 1, s1)
        const int64_t s0 = (int64_t)tile.f0 * p.S;
        int64_t       s1 = s0 + (int64_t)(tile.nf - 1) * p.S + p.L;
        if (s1 > uLen)
            s1 = uLen;
        const int64_t gFirst = uBeg + (s0 > 0 ? s0 - 1 : 0);   // first global sample needed
        const int64_t gLast  = uBeg + s1;                       // one past the last
        const int64_t gA     = gFirst & ~(int64_t)3;            // 16-byte aligned start of the staged span
        const int64_t gB     = gLast & ~(int64_t)3;             // aligned end of the bulk part
        const int     lead   = (int)(gFirst - gA);              // smem index of sample gFirst
        __syncthreads();  // previous tile fully consumed before the staging buffer is overwritten
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)(gB - gA) * 4u;
            if (bytes) {
                mbar_expect_tx(&bar, bytes);
                bulk_g2s(sSamples, p.samples + gA, bytes, &bar);
            }
            else {
                mbar_arrive(&bar);
            }
        }
        // ragged tail (< 4 samples) with plain loads
        if (threadIdx.x < (int)(gLast - gB))
            sSamples[(int)(gB - gA) + threadIdx.x] = p.samples[gB + threadIdx.x];
        mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        // smem index of utterance sample i is  i - (s0>0 ? s0-1 : 0) + lead
        const int base = lead - (int)(s0 > 0 ? s0 - 1 : 0);

        for (int fi = warp; fi < tile.nf; fi += kWarps) {
            const int64_t fs  = (int64_t)(tile.f0 + fi) * p.S;  // first sample of the frame
            int           len = p.L;
            if (fs + len > uLen)
                len = (int)(uLen - fs);  // short last frame (WindowBuffer::flush)
            // ---- pre-emphasis + window + zero padding, written in bit-reversed complex order
            for (int n = lane; n < p.N; n += 32) {
                float v = 0.0f;
                if (n < len) {
                    const int64_t i   = fs + n;
                    const float   cur = sSamples[base + (int)i];
                    // the first sample of a segment is its own predecessor (Preemphasis.cc:54-55)
                    const float prev = i > 0 ? sSamples[base + (int)i - 1] : cur;
                    const float e    = p.alpha == 1.0f ? __fsub_rn(cur, prev) : __fmaf_rn(-p.alpha, prev, cur);
                    v                = __fmul_rn(sWindow[n], e);
                }
                z[2 * bitrev(n >> 1, log2M) + (n & 1)] = v;
            }
            __syncwarp();
            // ---- radix-2 decimation-in-time, twiddles e^{+i 2 pi k / M}
            for (int h = 1; h < M; h <<= 1) {
                const int tstep = (M >> 1) / h;
                for (int bf = lane; bf < (M >> 1); bf += 32) {
                    const int    pos = bf & (h - 1);
                    const int    i   = ((bf - pos) << 1) + pos;
                    const int    j   = i + h;
                    const float2 w   = sTw[pos * tstep];
                    const float2 a = zc[i], b = zc[j];
                    const float  tR = __fmaf_rn(w.x, b.x, -__fmul_rn(w.y, b.y));
                    const float  tI = __fmaf_rn(w.x, b.y, __fmul_rn(w.y, b.x));
                    zc[j]           = make_float2(__fsub_rn(a.x, tR), __fsub_rn(a.y, tI));
                    zc[i]           = make_float2(__fadd_rn(a.x, tR), __fadd_rn(a.y, tI));
                }
                __syncwarp();
            }
            // ---- split into the spectrum of the real sequence, scale by 1/sampleRate, amplitude
            for (int k = lane; k <= (M >> 1); k += 32) {
                if (k == 0) {
                    const float2 a = zc[0];
                    const float  x0 = __fmul_rn(__fadd_rn(a.x, a.y), p.scale);
                    const float  xn = __fmul_rn(__fsub_rn(a.x, a.y), p.scale);
                    amp[0]          = fabsf(x0);
                    amp[M]          = fabsf(xn);
                }
                else if (k == (M >> 1)) {
                    const float2 a  = zc[k];
                    const float  re = __fmul_rn(a.x, p.scale), im = __fmul_rn(a.y, p.scale);
                    amp[k]          = __fsqrt_rn(__fmaf_rn(re, re, __fmul_rn(im, im)));
                }
                else {
                    const float2 a = zc[k], b = zc[M - k];
                    const float2 w = sTws[k];
                    const float  h1R = 0.5f * (a.x + b.x), h1I = 0.5f * (a.y - b.y);
                    const float  h2R = 0.5f * (a.y + b.y), h2I = -0.5f * (a.x - b.x);
                    const float  uR = __fmaf_rn(w.x, h2R, -__fmul_rn(w.y, h2I));  // wR*h2R - wI*h2I
                    const float  uI = __fmaf_rn(w.x, h2I, __fmul_rn(w.y, h2R));   // wR*h2I + wI*h2R
                    float        re = __fmul_rn(h1R + uR, p.scale), im = __fmul_rn(h1I + uI, p.scale);
                    amp[k]          = __fsqrt_rn(__fmaf_rn(re, re, __fmul_rn(im, im)));
                    re              = __fmul_rn(h1R - uR, p.scale);
                    im              = __fmul_rn(uI - h1I, p.scale);
                    amp[M - k]      = __fsqrt_rn(__fmaf_rn(re, re, __fmul_rn(im, im)));
                }
            }
            __syncwarp();
            const int64_t t = fOut + fi;
            if (p.dbgAmp)
                for (int k = lane; k < p.nBins; k += 32)
                    p.dbgAmp[t * p.nBins + k] = amp[k];
            // ---- mel filter bank: sequential f32 multiply-add over the taps of each filter
            for (int f = lane; f < p.nFilters; f += 32) {
                const int    a = sFbStart[f], e = sFbEnd[f];
                const float* w = sFbW + sFbOff[f];
                float        r = 0.0f;
                for (int k = a; k < e; ++k)
                    r = __fmaf_rn(amp[k], w[k - a], r);
                if (p.dbgFbank)
                    p.dbgFbank[t * p.nFilters + f] = r;
                fb[f] = log10f(r);
            }
            __syncwarp();
            // ---- DCT-II (unnormalised), sequential dot product per cepstral coefficient
            for (int c = lane; c < p.nCep; c += 32) {
                const float* row = sDct + c * p.nFilters;
                float        r   = 0.0f;
                for (int n = 0; n < p.nFilters; ++n)
                    r = __fmaf_rn(row[n], fb[n], r);
                p.cep[t * p.nCep + c] = r;
                if (!p.derivatives)
                    p.feats[t * p.featDim + c] = r;
            }
            __syncwarp();
        }
    }
}

// one thread per (frame, coefficient): static | delta | delta-delta
__global__ void __launch_bounds__(256) mfcc_derivative_kernel(const FeParams p) {
    for (int tileIdx = blockIdx.x; tileIdx < p.nGroupTiles; tileIdx += gridDim.x) {
        const Tile    tile = p.groupTiles[tileIdx];
        const int64_t u0   = p.groupFrameOff[tile.utt];
        const int64_t nU   = p.groupFrameOff[tile.utt + 1] - u0;
        const int     K    = p.nCep;
        for (int idx = threadIdx.x; idx < tile.nf * K; idx += blockDim.x) {
            const int     fi = idx / K, c = idx - fi * K;
            const int64_t ft = tile.f0 + fi;  // frame within the utterance
            float         f[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                int64_t u = ft + i - 2;
                u         = u < 0 ? 0 : (u > nU - 1 ? nU - 1 : u);  // margin-policy copy
                f[i]      = p.cep[(u0 + u) * K + c];
            }
            // regressFirstOrder (Regression.cc:25-39): sum dt*f / sum dt^2, dt = i - 2
            float d = 0.0f, tm = 0.0f;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float dt = (float)i - 2.0f;
                d              = __fmaf_rn(dt, f[i], d);
                tm             = __fmaf_rn(dt, dt, tm);
            }
            d = __fdiv_rn(d, tm);
            // regressSecondOrder (Regression.cc:41-63)
            float ns = 0.0f;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float dt = (float)i - 2.0f;
                ns             = __fmaf_rn(__fmul_rn(__fmul_rn(dt, dt), dt), dt, ns);
            }
            ns       = __fmaf_rn(-5.0f, ns, __fmul_rn(tm, tm));
            float dd = 0.0f;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float dt = (float)i - 2.0f;
                dd             = __fmaf_rn(f[i], tm, dd);
                const float t  = __fmul_rn(__fmul_rn(f[i], dt), dt);
                dd             = __fmaf_rn(-t, 5.0f, dd);
            }
            dd = (float)((double)dd * (2.0 / (double)ns));
            float* o  = p.feats + (u0 + ft) * p.featDim;
            o[c]         = f[2];
            o[K + c]     = d;
            o[2 * K + c] = dd;
        }
    }
}

// generic-vector-s16-demultiplex (track of an interleaved stream) + generic-convert-vector-s16-to-vector-f32
// (src/Tools/FeatureExtraction/share/samples.flow:13-18, src/Flow/TypeConverter.hh:113): plain value conversion
__global__ void __launch_bounds__(256) convert_s16_kernel(const int16_t* __restrict__ in, float* __restrict__ out,
                                                          long n, int channels, int track) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = (float)in[i * channels + track];
}

#include "frontend_fft256.cuh"

}  // namespace

// ==========================================================================================
// host side
// ==========================================================================================

struct rb_frontend {
    rb::DeviceInfo  dev;
    rb_frontend_cfg cfg;
    cudaStream_t    stream = nullptr;
    // geometry
    double sampleRate = 0;  // as read back from the text attribute
    int    L = 0, S = 0, N = 0, nBins = 0, nFilters = 0, nWeights = 0, featDim = 0, log2M = 0;
    // host tables
    std::vector<float> window, dct, fbWeights;  // fbWeights: concatenated taps
    std::vector<int>   fbStart, fbEnd, fbOff;
    std::vector<float> blob;
    int oWindow = 0, oTw = 0, oTws = 0, oFbStart = 0, oFbEnd = 0, oFbOff = 0, oFbW = 0, oDct = 0;
    // launch configuration
    int    sampleCap = 0, nBinsPad = 0, fbPad = 0, grid = 0;
    size_t smemBytes = 0;
    // register-resident kernel for the 512-point geometry (frontend_fft256.cuh)
    bool       fast = false;
    F256Tables f256;
    std::vector<float> fastBlob;
    rb::DevBuf<float>  dFastTables;
    int        fastSampleCap = 0, fastGrid = 0, fastTpl = 0;
    void (*fastKernel)(const FeParams, const F256Tables, int) = nullptr;
    size_t     fastSmemBytes = 0;
    // device buffers
    rb::DevBuf<float>   dTables, dSamples, dCep, dFeats, dDbgAmp, dDbgFbank;
    rb::DevBuf<int16_t> dPcm;  // interleaved s16 input of rb_frontend_process_s16
    // signal-dc-detection (rb_frontend_process_dc): flag bits, run table on the device, runs of the last call
    rb::DevBuf<uint32_t> dDcBits;
    rb::DevBuf<int64_t>  dDcOff, dDcRunBeg, dDcRunEnd;
    rb::DevBuf<double>   dDcRunStart;
    rb::DevBuf<int>      dDcCount;
    std::vector<int64_t> dcRunUtt, dcRunBeg, dcRunEnd;
    std::vector<double>  dcRunStart;
    int                  dcSlowPath = 0;
    bool                 dcEnabled  = false;  // streaming API: rb_frontend_set_dc_detection
    rb_dc_cfg            dcCfg{};
    rb::DevBuf<double>   dDcUttStart;
    static constexpr int kSlots = 4;
    struct StageSlot {
        rb::PinnedBuf<char> host;
        rb::DevBuf<char>    dev;
        cudaEvent_t         ev = nullptr, evUp = nullptr;
    } slots[kSlots];
    int nextSlot = 0;
    // Host-buffer calls set this to their H2D stream: the tile tables are then uploaded on it, in line with the bulk
    // sample copies.  Issued on the kernel stream the small copy only reaches the copy engine once the stream's
    // dependencies have resolved -- by then the bulk copies of the NEXT slabs are queued in front of it, and the
    // kernels of slab i wait for the transfers of slabs i+1.. (measured: 0.45 ms per call).
    cudaStream_t uploadStream = nullptr;
    rb::CopyStreams copy;  // streams / events of the host-buffer calls
    // streaming state
    std::vector<float> pending;
    double             pendingStart = 0;
    bool               havePending  = false;
    // results of the last finish()/process()
    long                 lastFrames = 0;
    std::vector<float>   lastFeats;
    std::vector<double>  lastStart, lastEnd;
    bool                 debug = false;
    long                 dbgFrames = 0;

    ~rb_frontend() {
        for (StageSlot& sl : slots)
            if (sl.ev) {
                cudaEventSynchronize(sl.ev);
                cudaEventDestroy(sl.ev);
            }
        for (StageSlot& sl : slots)
            if (sl.evUp)
                cudaEventDestroy(sl.evUp);
        if (stream)
            cudaStreamDestroy(stream);
    }
};

namespace {

// Flow attributes travel as text with default ostream precision (src/Flow/Attributes.hh:104-113)
double attribute_round_trip(double v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%g", v);  // "%g" == default ostream formatting (6 significant digits)
    return atof(buf);
}

bool almost_equal(double a, double b) {  // Core::isAlmostEqual src/Core/Utility.hh:322-327
    const double eps = 2.2204460492503131e-16, tiny = 2.2250738585072014e-308;
    return std::fabs(a - b) < (std::fabs(a) + std::fabs(b) + tiny) * eps;
}

bool almost_integer(double x) {  // FilterBank::isAlmostInteger src/Signal/Filterbank.cc:690-693
    return std::fabs(x - std::round(x)) < 1e-10;
}

struct MelScale {
    double hzPerIndex;  // continuous frequency of one FFT bin
    double toMel(double index) const { return 2595.0 * std::log10(1.0 + hzPerIndex * index / 700.0); }
    double toIndex(double mel) const { return (1.0 / hzPerIndex) * ((std::pow(10.0, (1.0 / 2595.0) * mel) - 1.0) * 700.0); }
    double slope(double index) const { return 2595.0 * (1.0 / std::log(10.0) / (700.0 + hzPerIndex * index)); }
};

int build_geometry(rb_frontend* h) {
    const rb_frontend_cfg& c = h->cfg;
    RB_REQUIRE(c.sample_rate > 0, "sample rate must be positive");
    RB_REQUIRE(c.n_cepstra >= 1, "nr-outputs of the cosine transform must be >= 1");
    h->sampleRate = attribute_round_trip(c.sample_rate);
    h->L          = (int)(unsigned)rint(c.window_length_s * h->sampleRate);  // Window.cc:73-79
    h->S          = (int)(unsigned)rint(c.window_shift_s * h->sampleRate);
    RB_REQUIRE(h->L >= 1 && h->S >= 1, "window length/shift round to zero samples");
    const unsigned maxLen = (unsigned)ceil(c.fft_max_input_s * h->sampleRate);
    RB_REQUIRE(maxLen >= 1, "maximum-input-size of the FFT is zero");
    double power = std::log((double)maxLen) / std::log(2.0);  // FastFourierTransform.cc:30-41
    power        = almost_equal(power, rint(power)) ? rint(power) : ceil(power);
    RB_REQUIRE(power <= 12, "FFT length 2^%d is outside the supported range (<= 4096)", (int)power);
    h->N = 1 << (unsigned)power;
    if (h->N < 64) {
        rb::set_error("FFT length %d < 64 is not supported by the warp-level transform", h->N);
        return RB_ERR_UNSUPPORTED;
    }
    RB_REQUIRE(h->L <= h->N, "window of %d samples does not fit the %d-point FFT", h->L, h->N);
    h->nBins = h->N / 2 + 1;
    h->log2M = (int)power - 1;
    return RB_OK;
}

// FilterBankNode::init + StretchToCover + FilterBuilder (src/Signal/Filterbank.cc:144-244,523-569,765-820)
int build_filterbank(rb_frontend* h) {
    MelScale mel;
    const double fftOutRate = attribute_round_trip(h->N / h->sampleRate);  // "sample-rate" after the FFT node
    mel.hzPerIndex          = 1.0 / fftOutRate;
    const double lo = 0.0, hi = mel.toMel(h->nBins - 1);
    double       width = h->cfg.filter_width;
    RB_REQUIRE(width > 0, "filter-width must be positive");
    double spacing = 0.5 * width;
    double count   = (hi - lo - width) / spacing + 1;
    if (count < 1)
        count = 1;
    else if (almost_integer(count))
        count = std::round(count);
    const size_t nF       = (size_t)std::floor(count);
    const double coverage = (spacing * (nF - 1) + width) / (hi - lo);
    if (!(nF == 1 && coverage > 1 && !almost_equal(coverage, 1))) {
        width /= coverage;
        spacing /= coverage;
    }
    h->nFilters = (int)nF;
    h->fbStart.resize(nF);
    h->fbEnd.resize(nF);
    h->fbOff.resize(nF);
    h->fbWeights.clear();
    for (size_t i = 0; i < nF; ++i) {
        const double center = lo + spacing * i + 0.5 * width;
        double       a      = mel.toIndex(std::max(center - 0.5 * width, lo));
        a                   = almost_integer(a) ? std::round(a) : std::ceil(a);
        double e            = mel.toIndex(std::min(center + 0.5 * width, hi));
        e                   = almost_integer(e) ? std::round(e) + 1 : std::ceil(e);
        h->fbStart[i]       = (int)(size_t)a;
        h->fbEnd[i]         = (int)(size_t)e;
        RB_REQUIRE(h->fbStart[i] >= 0 && h->fbEnd[i] <= h->nBins && h->fbStart[i] <= h->fbEnd[i],
                   "filter %zu covers bins [%d,%d) outside the spectrum", i, h->fbStart[i], h->fbEnd[i]);
        h->fbOff[i] = (int)h->fbWeights.size();
        for (int k = h->fbStart[i]; k < h->fbEnd[i]; ++k) {
            float tri = (float)(1.0 - std::fabs(mel.toMel(k) - center) / (width / 2));
            if (!(tri >= 0))
                tri = 0;
            h->fbWeights.push_back((float)(tri * mel.slope(k)));
        }
    }
    h->nWeights = (int)h->fbWeights.size();
    return RB_OK;
}

int build_tables(rb_frontend* h) {
    // window functions of src/Signal/WindowFunction.cc:62-132 (f64 expressions narrowed to f32; symmetric fill)
    h->window.assign(h->L, 0.0f);
    const unsigned size = (unsigned)h->L;
    switch (h->cfg.window_type) {
        case RB_WINDOW_RECTANGULAR:
            std::fill(h->window.begin(), h->window.end(), 1.0f);
            break;
        case RB_WINDOW_HAMMING:
        case RB_WINDOW_BARTLETT:
        case RB_WINDOW_BLACKMAN:
            if (size > 1) {
                const unsigned Mw = size - 1;
                for (unsigned n = 0; n <= Mw / 2; ++n) {
                    double v;
                    if (h->cfg.window_type == RB_WINDOW_HAMMING)
                        v = 0.54 - 0.46 * cos(2.0 * M_PI * n / Mw);
                    else if (h->cfg.window_type == RB_WINDOW_BARTLETT)
                        v = 2.0 * (float)n / (float)Mw;
                    else
                        v = 0.42 - 0.5 * cos(2.0 * M_PI * n / Mw) + 0.08 * cos(4.0 * M_PI * n / Mw);
                    h->window[n] = h->window[Mw - n] = (float)v;
                }
            }
            break;
        case RB_WINDOW_HANNING:
        case RB_WINDOW_PERIODIC_HANNING:
            if (size > 1) {
                const unsigned Mw = size - (h->cfg.window_type == RB_WINDOW_PERIODIC_HANNING ? 0 : 1);
                for (unsigned n = 0; n <= Mw / 2; ++n) {
                    h->window[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / Mw));
                    if (Mw - n < size)
                        h->window[Mw - n] = h->window[n];
                }
            }
            break;
        case RB_WINDOW_KAISER:
            // KaiserWindowFunction::init (src/Signal/KaiserWindowFunction.cc:22-33) with Math::Nr::bessi0
            // (src/Math/Nr/BesselFunctions.cc:22-37).  WindowFunction::create builds it with beta = 0 and no node
            // parameter reaches setBeta, so the window the reference produces is bessi0(0) / bessi0(0) everywhere
            if (size > 1) {
                const double   beta = 0.0;
                const unsigned Mw   = size - 1;
                auto bessi0 = [](double x) {
                    const double ax = fabs(x);
                    if (ax < 3.75) {
                        double y = x / 3.75;
                        y *= y;
                        return 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
                    }
                    const double y = 3.75 / ax;
                    return (exp(ax) / sqrt(ax)) *
                           (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
                };
                for (unsigned n = 0; n <= Mw / 2; ++n) {
                    const double u = (double)n / (Mw / 2.0) - 1.0;
                    h->window[n] = h->window[Mw - n] = (float)(bessi0(beta * sqrt(1.0 - u * u)) / bessi0(beta));
                }
            }
            break;
        default:
            rb::set_error("unknown window type %d", h->cfg.window_type);
            return RB_ERR_UNSUPPORTED;
    }
    RB_CHECK(build_filterbank(h));
    // DCT-II, even about N-1/2 (CosineTransform.cc:62-74)
    const int K = h->cfg.n_cepstra, F = h->nFilters;
    h->dct.resize((size_t)K * F);
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < F; ++n)
            h->dct[(size_t)k * F + n] = (float)cos(M_PI * (n + 0.5) / F * k);
    // twiddles: e^{+i 2 pi k / M} for the complex transform (k < M/2) and e^{+i 2 pi k / N} for the split
    const int M = h->N / 2;
    std::vector<float> tw((size_t)M, 0.0f), tws((size_t)M, 0.0f);
    for (int k = 0; k < M / 2; ++k) {
        tw[2 * k]      = (float)cos(2.0 * M_PI * k / M);
        tw[2 * k + 1]  = (float)sin(2.0 * M_PI * k / M);
        tws[2 * k]     = (float)cos(2.0 * M_PI * k / h->N);
        tws[2 * k + 1] = (float)sin(2.0 * M_PI * k / h->N);
    }
    // blob
    std::vector<float>& b = h->blob;
    b.clear();
    auto align4 = [&]() {
        while (b.size() % 4)
            b.push_back(0.0f);
    };
    auto put_f = [&](const std::vector<float>& v) {
        align4();
        int o = (int)b.size();
        b.insert(b.end(), v.begin(), v.end());
        return o;
    };
    auto put_i = [&](const std::vector<int>& v) {
        align4();
        int o = (int)b.size();
        for (int x : v) {
            float f;
            std::memcpy(&f, &x, 4);
            b.push_back(f);
        }
        return o;
    };
    h->oWindow  = put_f(h->window);
    h->oTw      = put_f(tw);
    h->oTws     = put_f(tws);
    h->oFbStart = put_i(h->fbStart);
    h->oFbEnd   = put_i(h->fbEnd);
    h->oFbOff   = put_i(h->fbOff);
    h->oFbW     = put_f(h->fbWeights);
    h->oDct     = put_f(h->dct);
    align4();

    // ---- tables of the register-resident 512-point kernel: their own blob (the generic tables are not needed then)
    h->fast = h->N == 512 && (h->S % 2) == 0 && F <= 32 && K <= 32 &&               (kTileFrames - 1) * h->S + h->N <= kPreRounds * kF256Threads &&
              getenv("RB_FRONTEND_GENERIC") == nullptr;
    if (h->fast) {
        F256Tables& t = h->f256;
        std::vector<float> twA(7 * 32 * 2), twB(7 * 4 * 2), tws512(256 * 2), win(512, 0.0f);
        for (int k1 = 1; k1 < 8; ++k1)
            for (int l = 0; l < 32; ++l) {
                twA[((k1 - 1) * 32 + l) * 2]     = (float)cos(2.0 * M_PI * (l * k1) / 256.0);
                twA[((k1 - 1) * 32 + l) * 2 + 1] = (float)sin(2.0 * M_PI * (l * k1) / 256.0);
            }
        for (int j1 = 1; j1 < 8; ++j1)
            for (int m0 = 0; m0 < 4; ++m0) {
                twB[((j1 - 1) * 4 + m0) * 2]     = (float)cos(2.0 * M_PI * (m0 * j1) / 32.0);
                twB[((j1 - 1) * 4 + m0) * 2 + 1] = (float)sin(2.0 * M_PI * (m0 * j1) / 32.0);
            }
        for (int k = 0; k < 256; ++k) {
            tws512[2 * k]     = (float)cos(2.0 * M_PI * k / 512.0);
            tws512[2 * k + 1] = (float)sin(2.0 * M_PI * k / 512.0);
        }
        std::copy(h->window.begin(), h->window.end(), win.begin());
        // mel: every lane owns one contiguous piece of one filter; smallest piece length that needs <= 32 lanes
        int tpl = 0;
        for (int cand : {16, 24, 32}) {
            int lanes = 0, worst = 0;
            for (int f = 0; f < F; ++f) {
                const int n = h->fbEnd[f] - h->fbStart[f], pcs = std::max(1, (n + cand - 1) / cand);
                lanes += pcs;
                worst = std::max(worst, pcs);
            }
            if (lanes <= 32 && worst <= 4) {
                tpl = cand;
                break;
            }
        }
        if (!tpl)
            h->fast = false;
        h->fastTpl = tpl ? tpl : 32;
        std::vector<float> melW((size_t)h->fastTpl * 32, 0.0f);
        std::vector<int>   melBin(32, 257), pieceOf(32, 0), pcBin, pcLen, pcTap;
        for (int f = 0; f < F && tpl; ++f) {
            const int n = h->fbEnd[f] - h->fbStart[f], pcs = std::max(1, (n + tpl - 1) / tpl);
            pieceOf[f] = (int)pcBin.size() | (pcs << 8);
            for (int q = 0; q < pcs; ++q) {  // pieces of (almost) equal length, in bin order
                const int a0 = (int)((long)n * q / pcs), a1 = (int)((long)n * (q + 1) / pcs);
                pcBin.push_back(h->fbStart[f] + a0);
                pcLen.push_back(a1 - a0);
                pcTap.push_back(h->fbOff[f] + a0);
            }
        }
        // a lane may start reading up to (tpl - length) bins early (weight 0 there): pick the starts so that the
        // 32 lanes hit 32 different banks (bipartite matching pieces -> residues mod 32; kept unshifted if none exists)
        const int          nPc = (int)pcBin.size();
        std::vector<int>   owner(32, -1), shift(nPc, 0);
        std::function<bool(int, std::vector<char>&)> augment = [&](int i, std::vector<char>& seen) {
            for (int sft = 0; sft <= tpl - pcLen[i] && sft <= pcBin[i]; ++sft) {
                const int r = (pcBin[i] - sft) & 31;
                if (seen[r])
                    continue;
                seen[r] = 1;
                if (owner[r] < 0 || augment(owner[r], seen)) {
                    owner[r] = i;
                    return true;
                }
            }
            return false;
        };
        bool perfect = nPc > 0;
        for (int i = 0; i < nPc && perfect; ++i) {
            std::vector<char> seen(32, 0);
            perfect = augment(i, seen);
        }
        if (perfect)
            for (int r = 0; r < 32; ++r)
                if (owner[r] >= 0)
                    shift[owner[r]] = (pcBin[owner[r]] - r) & 31;
        for (int i = 0; i < nPc; ++i) {
            melBin[i] = pcBin[i] - shift[i];
            for (int k = 0; k < pcLen[i]; ++k)
                melW[(size_t)(k + shift[i]) * 32 + i] = h->fbWeights[pcTap[i] + k];
        }
        if (perfect) {  // idle lanes read from the residues nobody uses
            int lane = nPc;
            for (int r = 0; r < 32 && lane < 32; ++r)
                if (owner[r] < 0)
                    melBin[lane++] = 257 + ((r - 257) & 31);
        }
        std::vector<float> dctT((size_t)F * 32, 0.0f);
        for (int c = 0; c < K; ++c)
            for (int n = 0; n < F; ++n)
                dctT[(size_t)n * 32 + c] = h->dct[(size_t)c * F + n];
        std::vector<float> keep;
        keep.swap(b);  // b is a reference to h->blob: build the fast blob in it, then swap back
        t.oTwA         = put_f(twA);
        t.oTwB         = put_f(twB);
        t.oMelW        = put_f(melW);
        t.oMelBin      = put_i(melBin);
        t.oPiece       = put_i(pieceOf);
        t.oDctT        = put_f(dctT);
        align4();
        t.stagedFloats = (int)b.size();
        t.oTws         = put_f(tws512);
        t.oWin         = put_f(win);
        align4();
        h->fastBlob.swap(b);
        b.swap(keep);
    }
    return RB_OK;
}

long frames_for(const rb_frontend* h, long n) {
    if (n <= 0)
        return 0;
    const long M = std::max(h->S, h->L);
    if (n <= M)
        return 1;
    return (n - M + h->S - 1) / h->S + 1;
}

// Timestamps exactly as the nodes produce them: WindowBuffer accumulates bufferStart by repeated
// += shift/sampleRate (WindowBuffer.cc:94); merged packets span [min start, max end] of the delay window
void timestamps(const rb_frontend* h, long nSamples, double start0, long T, double* ts, double* te) {
    if (!ts && !te)
        return;
    std::vector<double> s(T), e(T);
    double              cur = start0;
    for (long t = 0; t < T; ++t) {
        long len = h->L;
        if ((long)t * h->S + len > nSamples)
            len = nSamples - (long)t * h->S;
        s[t] = cur;
        e[t] = cur + (double)len / (double)h->sampleRate;
        cur += (double)h->S / (double)h->sampleRate;
    }
    for (long t = 0; t < T; ++t) {
        double a = s[t], b = e[t];
        if (h->cfg.derivatives)
            for (int i = -2; i <= 2; ++i) {
                long u = std::min<long>(std::max<long>(t + i, 0), T - 1);
                a      = std::min(a, s[u]);
                b      = std::max(b, e[u]);
            }
        if (ts)
            ts[t] = a;
        if (te)
            te[t] = b;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// signal-dc-detection (src/Signal/DcDetection.{hh,cc}; wired in front of the MFCC chain by samples.flow:34-37).
//
// The reference walks the samples one by one: a sample is "non-DC" if it differs from the LAST non-DC sample by at
// least max-dc-increment; a stretch of at least min-dc-length DC samples is discarded, non-DC segments shorter than
// min-non-dc-segment-length as well, and what is kept leaves in blocks of at most max(min-non-dc-segment-length,
// maximal-output-size) samples whose start times are accumulated block by block.
//
// The chain "last non-DC sample" is sequential, but when every sample either equals its predecessor or differs from
// it by at least the increment -- always true for 16-bit audio, whose values are integers, and the increment 0.9 --
// the reference sample always equals the previous sample (induction over the stream), so
//     non-DC(i)  <=>  |x[i] - x[i-1]| >= increment.
// dc_flags_kernel computes these flags as one bit per sample (and reports whether the premise holds);
// dc_runs_kernel then replays the block logic per utterance with one warp that jumps from event to event -- the
// next DC sample, the next non-DC sample, the next block cut -- by scanning 1024 flag bits per step.  If the premise
// fails (arbitrary float input), dc_runs_sequential_kernel restates the reference literally, one thread per utterance.
// Results (kept sample runs, their start times) are identical to the reference's in both cases.
// ---------------------------------------------------------------------------------------------------------------
struct DcParams {
    const float*   x;
    const int64_t* uOff;    // [nUtt + 1] sample ranges
    const double*  uStart;  // [nUtt] start time of the utterances, or null (0)
    int            nUtt;
    float          inc;
    uint32_t       minDc, minSeg, cut;  // samples; cut = max(minSeg, maximal-output-size)
    double         sampleRate;
    const uint32_t* bits;
    const int64_t* runOff;  // [nUtt + 1] capacity ranges in the run arrays
    int64_t*       runBeg;
    int64_t*       runEnd;
    double*        runStart;
    int*           nRuns;   // [nUtt]
};

__global__ void __launch_bounds__(256) dc_flags_kernel(const float* __restrict__ x, int64_t n, float inc,
                                                       const int64_t* __restrict__ uOff, int nUtt,
                                                       uint32_t* __restrict__ bits, int* violation) {
    const int     lane  = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x / 32);
    for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * 1024; base < n; base += warps * 1024) {
        uint32_t mine = 0;
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const int64_t i   = base + k * 32 + lane;
            bool          nd  = false;
            if (i < n) {
                const float cur = x[i], prev = i > 0 ? x[i - 1] : cur;
                nd = i == 0 || fabsf(cur - prev) >= inc;
                if (!nd && cur != prev) {
                    // premise broken unless i starts an utterance (its flag is never looked at)
                    int lo = 0, hi = nUtt;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (uOff[mid] < i)
                            lo = mid + 1;
                        else
                            hi = mid;
                    }
                    if (uOff[lo] != i)
                        *violation = 1;
                }
            }
            const uint32_t word = __ballot_sync(0xffffffffu, nd);
            if (lane == k)
                mine = word;
        }
        if (base + (int64_t)lane * 32 < n)
            bits[base / 32 + lane] = mine;
    }
}

struct DcRunWriter {
    int64_t* beg;
    int64_t* end;
    double*  start;
    int      cap, n;
    int64_t  lastEnd;
    uint32_t segLen, minSeg;
    double   time, sampleRate;
    // flushBlock = copyBlock + eraseBlock (DcDetection.cc:168-210)
    __device__ void emit(int64_t b, uint32_t nonDc, uint32_t dc, bool write) {
        segLen += nonDc;
        if (segLen >= minSeg) {
            if (n > 0 && lastEnd == b) {  // continues the previous block without a gap: same run
                if (write && n <= cap)
                    end[n - 1] = b + nonDc;
            }
            else {
                if (write && n < cap) {
                    beg[n]   = b;
                    end[n]   = b + nonDc;
                    start[n] = time;
                }
                ++n;
            }
            lastEnd = b + nonDc;
        }
        if (dc > 0)
            segLen = 0;
        time += (double)(nonDc + dc) / sampleRate;
    }
};

// smallest i in [pos, end) whose flag equals `want`, else end; cooperative over the warp: 4096 flags per step
// (one 16-byte load per lane; the flag buffer is padded so that the last load stays inside it)
__device__ __forceinline__ int64_t dc_next_flag(const uint32_t* __restrict__ bits, int64_t pos, int64_t end, bool want) {
    const int lane = threadIdx.x & 31;
    for (int64_t w = (pos >> 5) & ~(int64_t)3; w * 32 < end; w += 128) {
        const int64_t wi = w + lane * 4;
        uint32_t      v[4] = {0, 0, 0, 0};
        if (wi * 32 < end) {
            const uint4 raw = *reinterpret_cast<const uint4*>(bits + wi);
            v[0] = raw.x;
            v[1] = raw.y;
            v[2] = raw.z;
            v[3] = raw.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t s0 = (wi + j) * 32;  // first sample of the word
                v[j]             = want ? v[j] : ~v[j];
                if (s0 + 32 <= pos || s0 >= end)
                    v[j] = 0;
                else {
                    if (s0 < pos)
                        v[j] &= 0xffffffffu << (pos - s0);
                    if (end - s0 < 32)
                        v[j] &= (1u << (end - s0)) - 1u;
                }
            }
        }
        const uint32_t any = __ballot_sync(0xffffffffu, (v[0] | v[1] | v[2] | v[3]) != 0);
        if (any) {
            const int first = __ffs(any) - 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t fv = __shfl_sync(0xffffffffu, v[j], first);
                if (fv)
                    return (w + first * 4 + j) * 32 + (__ffs(fv) - 1);
            }
        }
    }
    return end;
}

__global__ void __launch_bounds__(128) dc_runs_kernel(const DcParams p) {
    const int u = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (u >= p.nUtt)
        return;
    const bool    write = (threadIdx.x & 31) == 0;
    const int64_t uBeg = p.uOff[u], uEnd = p.uOff[u + 1];
    DcRunWriter   out;
    out.beg        = p.runBeg + p.runOff[u];
    out.end        = p.runEnd + p.runOff[u];
    out.start      = p.runStart + p.runOff[u];
    out.cap        = (int)(p.runOff[u + 1] - p.runOff[u]);
    out.n          = 0;
    out.lastEnd    = -1;
    out.segLen     = 0;
    out.minSeg     = p.minSeg;
    out.time       = p.uStart ? p.uStart[u] : 0.0;
    out.sampleRate = p.sampleRate;
    if (uEnd > uBeg) {
        int64_t b   = uBeg;      // block start: the reference sample of the block (nonDcLength_ = 1)
        int64_t pos = uBeg + 1;  // next sample to look at; pos - 1 is a non-DC sample
        while (true) {
            // [pos, z0): non-DC samples only -- the block is cut at the first one that is `cut` samples from its start
            const int64_t z0 = dc_next_flag(p.bits, pos, uEnd, false);
            while (true) {
                const int64_t q = b + p.cut > pos ? b + p.cut : pos;
                if (q >= z0)
                    break;
                out.emit(b, (uint32_t)(q - b), 0, write);
                b   = q;
                pos = q + 1;
            }
            if (z0 == uEnd) {  // end of stream: lastBlock()
                out.emit(b, (uint32_t)(uEnd - b), 0, write);
                break;
            }
            // [z0, z1): DC hypotheses behind the non-DC sample z0 - 1
            const int64_t  z1 = dc_next_flag(p.bits, z0, uEnd, true);
            const uint32_t dc = (uint32_t)(z1 - z0);
            if (z1 == uEnd) {
                if (dc >= p.minDc)
                    out.emit(b, (uint32_t)(z0 - b), dc, write);
                else
                    out.emit(b, (uint32_t)(uEnd - b), 0, write);
                break;
            }
            if (dc >= p.minDc) {  // DC detected: the hypotheses are discarded, z1 starts the next block
                out.emit(b, (uint32_t)(z0 - b), dc, write);
                b = z1;
            }
            else if ((uint32_t)(z1 - b) >= p.cut) {  // hypotheses absorbed; block full
                out.emit(b, (uint32_t)(z1 - b), 0, write);
                b = z1;
            }
            pos = z1 + 1;
        }
    }
    if (write)
        p.nRuns[u] = out.n;
}

// literal restatement of nextBlock / lastBlock (DcDetection.cc:133-166) on the samples themselves
__global__ void __launch_bounds__(64) dc_runs_sequential_kernel(const DcParams p) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= p.nUtt)
        return;
    DcRunWriter out;
    out.beg        = p.runBeg + p.runOff[u];
    out.end        = p.runEnd + p.runOff[u];
    out.start      = p.runStart + p.runOff[u];
    out.cap        = (int)(p.runOff[u + 1] - p.runOff[u]);
    out.n          = 0;
    out.lastEnd    = -1;
    out.segLen     = 0;
    out.minSeg     = p.minSeg;
    out.time       = p.uStart ? p.uStart[u] : 0.0;
    out.sampleRate = p.sampleRate;
    int64_t  b = p.uOff[u], n = p.uOff[u + 1] - b;
    uint32_t nonDc = 1, dc = 0;
    if (n > 0) {
        float ref = p.x[b];
        while ((int64_t)nonDc + dc < n) {
            const float v = p.x[b + nonDc + dc];
            if (fabsf(v - ref) >= p.inc) {
                bool cutHere = dc >= p.minDc;
                if (!cutHere) {
                    nonDc += dc;
                    dc = 0;
                    cutHere = nonDc >= p.cut;
                }
                if (cutHere) {
                    out.emit(b, nonDc, dc, true);
                    b += nonDc + dc;
                    n -= nonDc + dc;
                    nonDc = 1;
                    dc    = 0;
                }
                else
                    ++nonDc;
                ref = v;  // v is the last non-DC sample now (of the old block, or the first of the new one)
            }
            else
                ++dc;
        }
        if (dc < p.minDc) {
            nonDc += dc;
            dc = 0;
        }
        out.emit(b, nonDc, dc, true);
    }
    p.nRuns[u] = out.n;
}

// One call = one staging slot: [sample offsets | frame offsets | tile table] built in pinned memory, sent with a
// single H2D copy.  Slots form a ring guarded by events, so the call only enqueues work (no stream synchronise):
// back-to-back calls keep the GPU queue full.
// Segments [segBeg[i], segEnd[i]) of the sample buffer produce the static features; derivative windows run over
// groups of consecutive segments (groupOff[g] .. groupOff[g + 1], segment indices).  groupOff == nullptr: every
// segment is its own group (plain calls: segments = utterances).
int run_segments(rb_frontend* h, const float* dSamples, const int64_t* segBeg, const int64_t* segEnd, int nSeg,
                 const int64_t* groupOff, int nGroups, float* dFeats, cudaStream_t s, int64_t* totalFramesOut) {
    size_t nTiles = 0, nGroupTiles = 0;
    {
        int64_t acc = 0;
        for (int u = 0; u < nSeg; ++u) {
            RB_REQUIRE(segEnd[u] >= segBeg[u], "sample offsets not monotone at segment %d", u);
            const long T = frames_for(h, (long)(segEnd[u] - segBeg[u]));
            acc += T;
            nTiles += (size_t)(T + kTileFrames - 1) / kTileFrames;
        }
        if (totalFramesOut)
            *totalFramesOut = acc;
        if (acc == 0)
            return RB_OK;
        if (groupOff)
            for (int g = 0; g < nGroups; ++g) {
                int64_t T = 0;
                for (int64_t u = groupOff[g]; u < groupOff[g + 1]; ++u)
                    T += frames_for(h, (long)(segEnd[u] - segBeg[u]));
                nGroupTiles += (size_t)(T + kTileFrames - 1) / kTileFrames;
            }
    }
    RB_REQUIRE(nTiles < (size_t)1 << 31 && nGroupTiles < (size_t)1 << 31, "too many frames in one call");
    rb_frontend::StageSlot& slot = h->slots[h->nextSlot];
    h->nextSlot                  = (h->nextSlot + 1) % rb_frontend::kSlots;
    if (!slot.ev)
        RB_CUDA(cudaEventCreateWithFlags(&slot.ev, cudaEventDisableTiming));
    else
        RB_CUDA(cudaEventSynchronize(slot.ev));  // the call that used this slot last has finished with it
    // [seg begin | seg end | seg frame offsets | group frame offsets | tiles | group tiles]
    const size_t segBytes = sizeof(int64_t) * (size_t)(nSeg + 1);
    const size_t grpBytes = groupOff ? sizeof(int64_t) * (size_t)(nGroups + 1) : 0;
    const size_t bytes    = 3 * segBytes + grpBytes + sizeof(Tile) * (nTiles + nGroupTiles);
    RB_CHECK(slot.host.reserve(bytes));
    RB_CHECK(slot.dev.reserve(bytes));
    int64_t* sOff   = reinterpret_cast<int64_t*>(slot.host.p);
    int64_t* sEnd   = reinterpret_cast<int64_t*>(slot.host.p + segBytes);
    int64_t* fOff   = reinterpret_cast<int64_t*>(slot.host.p + 2 * segBytes);
    int64_t* gOff   = reinterpret_cast<int64_t*>(slot.host.p + 3 * segBytes);
    Tile*    tiles  = reinterpret_cast<Tile*>(slot.host.p + 3 * segBytes + grpBytes);
    Tile*    gTiles = tiles + nTiles;
    auto addTiles = [](Tile* out, size_t& ti, int u, long T) {
        for (long f0 = 0; f0 < T; f0 += kTileFrames) {
            Tile t;
            t.utt = u;
            t.f0  = (int)f0;
            t.nf  = (int)std::min<long>(kTileFrames, T - f0);
            t.pad = 0;
            out[ti++] = t;
        }
    };
    fOff[0]   = 0;
    size_t ti = 0;
    for (int u = 0; u < nSeg; ++u) {
        sOff[u]      = segBeg[u];
        sEnd[u]      = segEnd[u];
        const long T = frames_for(h, (long)(segEnd[u] - segBeg[u]));
        fOff[u + 1]  = fOff[u] + T;
        addTiles(tiles, ti, u, T);
    }
    sOff[nSeg] = sEnd[nSeg] = nSeg ? segEnd[nSeg - 1] : 0;
    if (groupOff) {
        size_t gi = 0;
        for (int g = 0; g <= nGroups; ++g)
            gOff[g] = fOff[groupOff[g]];
        for (int g = 0; g < nGroups; ++g)
            addTiles(gTiles, gi, g, (long)(gOff[g + 1] - gOff[g]));
    }
    const int64_t total = fOff[nSeg];
    RB_CHECK(h->dCep.reserve((size_t)total * h->cfg.n_cepstra));
    if (h->uploadStream && h->uploadStream != s) {
        if (!slot.evUp)
            RB_CUDA(cudaEventCreateWithFlags(&slot.evUp, cudaEventDisableTiming));
        RB_CUDA(cudaMemcpyAsync(slot.dev.p, slot.host.p, bytes, cudaMemcpyHostToDevice, h->uploadStream));
        RB_CUDA(cudaEventRecord(slot.evUp, h->uploadStream));
        RB_CUDA(cudaStreamWaitEvent(s, slot.evUp, 0));
    }
    else
        RB_CUDA(cudaMemcpyAsync(slot.dev.p, slot.host.p, bytes, cudaMemcpyHostToDevice, s));

    FeParams p;
    p.samples      = dSamples;
    p.sampleOff    = reinterpret_cast<const int64_t*>(slot.dev.p);
    p.sampleEnd    = reinterpret_cast<const int64_t*>(slot.dev.p + segBytes);
    p.frameOff     = reinterpret_cast<const int64_t*>(slot.dev.p + 2 * segBytes);
    p.tiles        = reinterpret_cast<const Tile*>(slot.dev.p + 3 * segBytes + grpBytes);
    p.nTiles       = (int)nTiles;
    p.groupFrameOff = groupOff ? reinterpret_cast<const int64_t*>(slot.dev.p + 3 * segBytes) : p.frameOff;
    p.groupTiles    = groupOff ? p.tiles + nTiles : p.tiles;
    p.nGroupTiles   = groupOff ? (int)nGroupTiles : (int)nTiles;
    p.tables       = h->dTables.p;
    p.tableFloats  = (int)h->blob.size();
    p.L            = h->L;
    p.S            = h->S;
    p.N            = h->N;
    p.nBins        = h->nBins;
    p.nFilters     = h->nFilters;
    p.nCep         = h->cfg.n_cepstra;
    p.featDim      = h->featDim;
    p.derivatives  = h->cfg.derivatives;
    p.alpha        = h->cfg.preemphasis_alpha;
    p.scale        = h->sampleRate != 1 ? 1 / (float)h->sampleRate : 1.0f;
    p.oWindow      = h->oWindow;
    p.oTw          = h->oTw;
    p.oTws         = h->oTws;
    p.oFbStart     = h->oFbStart;
    p.oFbEnd       = h->oFbEnd;
    p.oFbOff       = h->oFbOff;
    p.oFbW         = h->oFbW;
    p.oDct         = h->oDct;
    p.cep          = h->dCep.p;
    p.feats        = dFeats;
    p.dbgAmp       = nullptr;
    p.dbgFbank     = nullptr;
    p.totalSamples = nSeg ? segEnd[nSeg - 1] : 0;
    if (h->debug) {
        RB_CHECK(h->dDbgAmp.reserve((size_t)total * h->nBins));
        RB_CHECK(h->dDbgFbank.reserve((size_t)total * h->nFilters));
        p.dbgAmp     = h->dDbgAmp.p;
        p.dbgFbank   = h->dDbgFbank.p;
        h->dbgFrames = (long)total;
    }
    if (h->fast) {
        const int grid = (int)std::min<size_t>(nTiles, (size_t)h->fastGrid);
        p.tables       = h->dFastTables.p;
        h->fastKernel<<<grid, kF256Threads, h->fastSmemBytes, s>>>(p, h->f256, h->fastSampleCap);
    }
    else {
        const int grid = (int)std::min<size_t>(nTiles, (size_t)h->grid);
        mfcc_static_kernel<<<grid, kThreads, h->smemBytes, s>>>(p, h->sampleCap, h->nBinsPad, h->fbPad, h->log2M);
    }
    RB_LAUNCH_CHECK();
    if (h->cfg.derivatives) {
        const int grid2 = (int)std::min<size_t>((size_t)p.nGroupTiles, (size_t)h->dev.sm_count * 8);
        mfcc_derivative_kernel<<<grid2, 256, 0, s>>>(p);
        RB_LAUNCH_CHECK();
    }
    RB_CUDA(cudaEventRecord(slot.ev, s));
    return RB_OK;
}

// utterances [offsets[u], offsets[u + 1])
int run_device(rb_frontend* h, const float* dSamples, const int64_t* offsets, int nUtt, float* dFeats,
               cudaStream_t s, int64_t* totalFramesOut) {
    return run_segments(h, dSamples, offsets, offsets + 1, nUtt, nullptr, nUtt, dFeats, s, totalFramesOut);
}

}  // namespace

extern "C" void rb_frontend_default_cfg(rb_frontend_cfg* cfg) {
    if (!cfg)
        return;
    cfg->sample_rate       = 16000.0;
    cfg->window_length_s   = 0.025;
    cfg->window_shift_s    = 0.01;
    cfg->fft_max_input_s   = 0.025;
    cfg->filter_width      = 268.258;
    cfg->preemphasis_alpha = 1.0f;
    cfg->n_cepstra         = 13;
    cfg->derivatives       = 1;
    cfg->device            = 0;
    cfg->window_type       = RB_WINDOW_HAMMING;
}

extern "C" int rb_frontend_create(const rb_frontend_cfg* cfg, rb_frontend** out) {
    RB_REQUIRE(cfg && out, "NULL argument");
    *out = nullptr;
    rb_frontend* h = new (std::nothrow) rb_frontend();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    h->cfg = *cfg;
    auto fail = [&](int code) {
        delete h;
        return code;
    };
    // geometry and tables first: configuration errors are reported even without a device
    int rc = build_geometry(h);
    if (rc == RB_OK)
        rc = build_tables(h);
    if (rc != RB_OK)
        return fail(rc);
    h->featDim = cfg->n_cepstra * (cfg->derivatives ? 3 : 1);
    rc = rb::use_device(cfg->device, &h->dev);
    if (rc != RB_OK)
        return fail(rc);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rb::set_error("cudaStreamCreate failed");
        return fail(RB_ERR_CUDA);
    }
    // shared memory plan
    h->sampleCap = (int)rb::round_up((size_t)(kTileFrames - 1) * h->S + h->L + 1 + 8, 4);
    h->nBinsPad  = (int)rb::round_up(h->nBins, 4);
    h->fbPad     = (int)rb::round_up(std::max(h->nFilters, 4), 4);
    const size_t floats = rb::round_up(h->blob.size(), 4) + h->sampleCap +
                          (size_t)kWarps * (h->N + h->nBinsPad + h->fbPad);
    h->smemBytes = floats * sizeof(float);
    if (h->smemBytes > h->dev.smem_optin) {
        rb::set_error("front-end configuration needs %zu bytes of shared memory per CTA (limit %zu)", h->smemBytes,
                      h->dev.smem_optin);
        return fail(RB_ERR_UNSUPPORTED);
    }
    if (cudaFuncSetAttribute(mfcc_static_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBytes) !=
        cudaSuccess) {
        rb::set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(RB_ERR_CUDA);
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mfcc_static_kernel, kThreads, h->smemBytes) !=
                cudaSuccess ||
        occ < 1) {
        rb::set_error("front-end kernel does not fit on the device");
        return fail(RB_ERR_CUDA);
    }
    h->grid = h->dev.sm_count * occ;
    if (h->fast) {
        h->fastSampleCap = (int)rb::round_up((size_t)(kTileFrames - 1) * h->S + h->N + 16, 4);
        h->fastSmemBytes = sizeof(float) * (rb::round_up((size_t)h->f256.stagedFloats, 4) + (size_t)h->fastSampleCap +
                                            (size_t)kF256Warps * kWarpScratch);
        const bool nf20 = h->nFilters == 20;
        switch (h->fastTpl) {
            case 16: h->fastKernel = nf20 ? mfcc_fft256_kernel<20, 16> : mfcc_fft256_kernel<0, 16>; break;
            case 24: h->fastKernel = nf20 ? mfcc_fft256_kernel<20, 24> : mfcc_fft256_kernel<0, 24>; break;
            default: h->fastKernel = nf20 ? mfcc_fft256_kernel<20, 32> : mfcc_fft256_kernel<0, 32>; break;
        }
        bool ok = h->fastSmemBytes <= h->dev.smem_optin;
        ok = ok && cudaFuncSetAttribute(h->fastKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)h->fastSmemBytes) == cudaSuccess;
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->fastKernel, kF256Threads,
                                                                 h->fastSmemBytes) == cudaSuccess;
        if (!ok || occ < 1) {
            cudaGetLastError();
            h->fast = false;  // the generic kernel still fits
        }
        else {
            h->fastGrid = h->dev.sm_count * occ;
            if (h->dFastTables.upload(h->fastBlob, h->stream) != RB_OK) {
                rb::set_error("table upload failed");
                return fail(RB_ERR_CUDA);
            }
        }
    }
    if (h->dTables.upload(h->blob, h->stream) != RB_OK || cudaStreamSynchronize(h->stream) != cudaSuccess) {
        rb::set_error("table upload failed");
        return fail(RB_ERR_CUDA);
    }
    *out = h;
    return RB_OK;
}

extern "C" void rb_frontend_destroy(rb_frontend* h) {
    if (!h)
        return;
    cudaSetDevice(h->dev.ordinal);
    rb_pipeline_forget(h);  // scratch buffers of the rb_pipeline_* calls that used this handle
    delete h;
}

extern "C" int rb_frontend_get_geometry(const rb_frontend* h, rb_frontend_geometry* g) {
    RB_REQUIRE(h && g, "NULL argument");
    g->win_length = h->L;
    g->win_shift  = h->S;
    g->fft_length = h->N;
    g->n_bins     = h->nBins;
    g->n_filters  = h->nFilters;
    g->n_weights  = h->nWeights;
    g->feat_dim   = h->featDim;
    return RB_OK;
}

extern "C" int rb_frontend_get_tables(const rb_frontend* h, float* window, int* fb_start, int* fb_end,
                                      float* fb_weights, float* dct) {
    RB_REQUIRE(h != nullptr, "NULL handle");
    if (window)
        std::copy(h->window.begin(), h->window.end(), window);
    for (int f = 0; f < h->nFilters; ++f) {
        if (fb_start)
            fb_start[f] = h->fbStart[f];
        if (fb_end)
            fb_end[f] = h->fbEnd[f];
        if (fb_weights) {
            for (int k = 0; k < h->nBins; ++k)
                fb_weights[(size_t)f * h->nBins + k] = 0.0f;
            for (int k = h->fbStart[f]; k < h->fbEnd[f]; ++k)
                fb_weights[(size_t)f * h->nBins + k] = h->fbWeights[h->fbOff[f] + k - h->fbStart[f]];
        }
    }
    if (dct)
        std::copy(h->dct.begin(), h->dct.end(), dct);
    return RB_OK;
}

extern "C" long rb_frontend_nframes_for(const rb_frontend* h, long n_samples) {
    return h ? frames_for(h, n_samples) : 0;
}

extern "C" long rb_frontend_timestamps(const rb_frontend* h, long n_samples, double start_time, double* t_start,
                                       double* t_end) {
    if (!h || n_samples < 0)
        return -1;
    const long T = frames_for(h, n_samples);
    if (T > 0)
        timestamps(h, n_samples, start_time, T, t_start, t_end);
    return T;
}

extern "C" long rb_frontend_count_frames(const rb_frontend* h, const int64_t* offsets, int n_utt,
                                         int64_t* frame_offsets) {
    if (!h || !offsets || n_utt < 0)
        return -1;
    int64_t acc = 0;
    if (frame_offsets)
        frame_offsets[0] = 0;
    for (int u = 0; u < n_utt; ++u) {
        acc += frames_for(h, (long)(offsets[u + 1] - offsets[u]));
        if (frame_offsets)
            frame_offsets[u + 1] = acc;
    }
    return (long)acc;
}

extern "C" int rb_frontend_process_dev(rb_frontend* h, const float* d_samples, const int64_t* offsets, int n_utt,
                                       float* d_feats, void* stream) {
    RB_REQUIRE(h && offsets && n_utt >= 0, "bad argument");
    if (n_utt == 0)
        return RB_OK;
    RB_REQUIRE(d_samples && d_feats, "NULL device buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    return run_device(h, d_samples, offsets, n_utt, d_feats, stream ? (cudaStream_t)stream : h->stream, nullptr);
}

namespace {
// samples: f32 mono (channels == 0) or interleaved s16 with `channels` channels of which `track` is used
int process_host(rb_frontend* h, const void* samplesRaw, int channels, int track, const int64_t* offsets, int n_utt,
                 float* feats, double* t_start, double* t_end) {
    RB_REQUIRE(h && offsets && n_utt >= 0, "bad argument");
    if (n_utt == 0)
        return RB_OK;
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    const float*   samples = channels ? nullptr : static_cast<const float*>(samplesRaw);
    const int16_t* pcm     = channels ? static_cast<const int16_t*>(samplesRaw) : nullptr;
    const int64_t base = offsets[0], nS = offsets[n_utt] - base;
    RB_REQUIRE(nS >= 0, "negative sample count");
    RB_REQUIRE(samplesRaw || nS == 0, "NULL sample buffer");
    if (channels)
        RB_CHECK(h->dPcm.reserve((size_t)nS * channels));
    // slack so that the aligned bulk copies never leave the allocation
    RB_CHECK(h->dSamples.reserve((size_t)nS + 8));
    std::vector<int64_t> rel(n_utt + 1), fo(n_utt + 1);
    for (int u = 0; u <= n_utt; ++u)
        rel[u] = offsets[u] - base;
    const long total = rb_frontend_count_frames(h, rel.data(), n_utt, fo.data());
    RB_CHECK(h->dFeats.reserve((size_t)total * h->featDim));
    // slabs of whole utterances (~8 per call): H2D of slab i+1 overlaps the kernels of slab i and the D2H of i-1.
    // Debug dumps index frames from 0, so a debug run is a single slab.
    const long       target = h->debug ? total + 1 : std::max<long>(8192, (total + 7) / 8);
    std::vector<int> cut(1, 0);
    for (int u = 1; u <= n_utt; ++u)
        if (u == n_utt || fo[u] - fo[cut.back()] >= target)
            cut.push_back(u);
    const int    nSlabs = (int)cut.size() - 1;
    RB_CHECK(h->copy.ensure(2 * (size_t)nSlabs));
    cudaStream_t sIn = h->copy.in, sOut = h->copy.out;
    cudaEvent_t* evIn = h->copy.pool.data();
    cudaEvent_t* evK  = h->copy.pool.data() + nSlabs;
    int rc = RB_OK;
    for (int i = 0; i < nSlabs && rc == RB_OK; ++i) {
        const int     u0 = cut[i], u1 = cut[i + 1];
        const int64_t sA = rel[u0], sB = rel[u1], fA = fo[u0], fB = fo[u1];
        if (sB > sA) {
            const cudaError_t e =
                    channels ? cudaMemcpyAsync(h->dPcm.p + sA * channels, pcm + (base + sA) * channels,
                                               (size_t)(sB - sA) * channels * 2, cudaMemcpyHostToDevice, sIn)
                             : cudaMemcpyAsync(h->dSamples.p + sA, samples + base + sA, (size_t)(sB - sA) * 4,
                                               cudaMemcpyHostToDevice, sIn);
            if (e != cudaSuccess)
                rc = RB_ERR_CUDA;
        }
        cudaEventRecord(evIn[i], sIn);
        cudaStreamWaitEvent(h->stream, evIn[i], 0);
        if (rc == RB_OK && channels && sB > sA) {
            const long n = (long)(sB - sA);
            convert_s16_kernel<<<(int)std::min<long>((n + 255) / 256, (long)h->dev.sm_count * 8), 256, 0, h->stream>>>(
                    h->dPcm.p + sA * channels, h->dSamples.p + sA, n, channels, track);
            rb::count_launch();
        }
        if (rc == RB_OK) {
            h->uploadStream = sIn;
            rc = run_device(h, h->dSamples.p, rel.data() + u0, u1 - u0, h->dFeats.p + fA * h->featDim, h->stream,
                            nullptr);
            h->uploadStream = nullptr;
        }
        cudaEventRecord(evK[i], h->stream);
        cudaStreamWaitEvent(sOut, evK[i], 0);
        if (rc == RB_OK && feats && fB > fA &&
            cudaMemcpyAsync(feats + fA * h->featDim, h->dFeats.p + fA * h->featDim, (size_t)(fB - fA) * h->featDim * 4,
                            cudaMemcpyDeviceToHost, sOut) != cudaSuccess)
            rc = RB_ERR_CUDA;
    }
    const cudaError_t e1 = cudaStreamSynchronize(sIn), e2 = cudaStreamSynchronize(h->stream),
                      e3 = cudaStreamSynchronize(sOut);
    if (rc == RB_ERR_CUDA && rb::get_error()[0] == 0)
        rb::set_error("asynchronous copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != RB_OK)
        return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        const cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
        rb::set_error("front-end failed on the device: %s", cudaGetErrorString(e));
        return RB_ERR_CUDA;
    }
    h->lastFrames = total;
    if (t_start || t_end) {
        long f = 0;
        for (int u = 0; u < n_utt; ++u) {
            const long n = (long)(rel[u + 1] - rel[u]);
            const long T = frames_for(h, n);
            timestamps(h, n, 0.0, T, t_start ? t_start + f : nullptr, t_end ? t_end + f : nullptr);
            f += T;
        }
    }
    return RB_OK;
}
}  // namespace

extern "C" int rb_frontend_process(rb_frontend* h, const float* samples, const int64_t* offsets, int n_utt,
                                   float* feats, double* t_start, double* t_end) {
    return process_host(h, samples, 0, 0, offsets, n_utt, feats, t_start, t_end);
}

extern "C" int rb_frontend_process_s16(rb_frontend* h, const int16_t* samples, int n_channels, int track,
                                       const int64_t* offsets, int n_utt, float* feats, double* t_start,
                                       double* t_end) {
    RB_REQUIRE(n_channels >= 1 && track >= 0 && track < n_channels, "track %d of %d channels", track, n_channels);
    return process_host(h, samples, n_channels, track, offsets, n_utt, feats, t_start, t_end);
}

extern "C" int rb_frontend_reset(rb_frontend* h) {
    RB_REQUIRE(h != nullptr, "NULL handle");
    h->pending.clear();
    h->havePending  = false;
    h->pendingStart = 0;
    h->lastFrames   = 0;
    h->lastFeats.clear();
    h->lastStart.clear();
    h->lastEnd.clear();
    return RB_OK;
}

extern "C" int rb_frontend_push(rb_frontend* h, const float* samples, long n, double start_time) {
    RB_REQUIRE(h != nullptr, "NULL handle");
    RB_REQUIRE(n >= 0 && (samples || n == 0), "bad packet");
    if (!h->havePending) {
        h->pendingStart = start_time;  // WindowBuffer::put on an empty buffer (WindowBuffer.cc:50-58)
        h->havePending  = true;
    }
    h->pending.insert(h->pending.end(), samples, samples + n);
    return RB_OK;
}

namespace {
int process_dc_impl(rb_frontend* h, const rb_dc_cfg* dc, const float* samples, const int64_t* offsets, int n_utt,
                    float* feats, long capacity, int64_t* frame_offsets, double* t_start, double* t_end,
                    const double* utt_start);
}

extern "C" int rb_frontend_finish(rb_frontend* h) {
    RB_REQUIRE(h != nullptr, "NULL handle");
    const long    n      = (long)h->pending.size();
    const int64_t off[2] = {0, n};
    if (h->dcEnabled) {  // signal-dc-detection between the samples and the chain
        const long cap = rb_frontend_dc_max_frames(h, &h->dcCfg, off, 1);
        if (cap < 0)
            return (int)cap;
        h->lastFeats.assign((size_t)cap * h->featDim, 0.0f);
        h->lastStart.assign(cap, 0.0);
        h->lastEnd.assign(cap, 0.0);
        int64_t fo[2] = {0, 0};
        RB_CHECK(process_dc_impl(h, &h->dcCfg, h->pending.data(), off, 1, h->lastFeats.data(), cap, fo,
                                 h->lastStart.data(), h->lastEnd.data(), &h->pendingStart));
        h->lastFrames = (long)fo[1];
        h->lastFeats.resize((size_t)fo[1] * h->featDim);
        h->lastStart.resize(fo[1]);
        h->lastEnd.resize(fo[1]);
        h->pending.clear();
        h->havePending = false;
        return RB_OK;
    }
    const long T = frames_for(h, n);
    h->lastFeats.assign((size_t)T * h->featDim, 0.0f);
    h->lastStart.assign(T, 0.0);
    h->lastEnd.assign(T, 0.0);
    if (T > 0) {
        RB_CHECK(rb_frontend_process(h, h->pending.data(), off, 1, h->lastFeats.data(), nullptr, nullptr));
        timestamps(h, n, h->pendingStart, T, h->lastStart.data(), h->lastEnd.data());
    }
    h->lastFrames = T;
    h->pending.clear();
    h->havePending = false;
    return RB_OK;
}

extern "C" long rb_frontend_nframes(const rb_frontend* h) {
    return h ? h->lastFrames : 0;
}

extern "C" int rb_frontend_read(rb_frontend* h, float* feats, double* t_start, double* t_end) {
    RB_REQUIRE(h != nullptr, "NULL handle");
    if ((size_t)h->lastFrames * h->featDim != h->lastFeats.size()) {
        rb::set_error("rb_frontend_read without a preceding rb_frontend_finish");
        return RB_ERR_STATE;
    }
    if (feats)
        std::copy(h->lastFeats.begin(), h->lastFeats.end(), feats);
    if (t_start)
        std::copy(h->lastStart.begin(), h->lastStart.end(), t_start);
    if (t_end)
        std::copy(h->lastEnd.begin(), h->lastEnd.end(), t_end);
    return RB_OK;
}

extern "C" int rb_frontend_set_debug(rb_frontend* h, int on) {
    RB_REQUIRE(h != nullptr, "NULL handle");
    h->debug = on != 0;
    return RB_OK;
}

extern "C" int rb_frontend_read_stages(rb_frontend* h, float* amplitude, float* fbank, float* cepstra) {
    RB_REQUIRE(h != nullptr, "NULL handle");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    const size_t T = (size_t)h->lastFrames;
    if (amplitude || fbank) {
        if (!h->debug || (size_t)h->dbgFrames != T) {
            rb::set_error("stage dumps were not recorded: call rb_frontend_set_debug(h, 1) before processing");
            return RB_ERR_STATE;
        }
    }
    if (amplitude && T)
        RB_CUDA(cudaMemcpy(amplitude, h->dDbgAmp.p, T * h->nBins * 4, cudaMemcpyDeviceToHost));
    if (fbank && T)
        RB_CUDA(cudaMemcpy(fbank, h->dDbgFbank.p, T * h->nFilters * 4, cudaMemcpyDeviceToHost));
    if (cepstra && T)
        RB_CUDA(cudaMemcpy(cepstra, h->dCep.p, T * h->cfg.n_cepstra * 4, cudaMemcpyDeviceToHost));
    return RB_OK;
}

// s16 -> f32 of one track on the device (pipeline.cu)
int rb_frontend_convert_s16_dev(const rb_frontend* h, const int16_t* d_pcm, float* d_out, long n, int channels,
                                int track, cudaStream_t s) {
    if (n <= 0)
        return RB_OK;
    convert_s16_kernel<<<(int)std::min<long>((n + 255) / 256, (long)h->dev.sm_count * 8), 256, 0, s>>>(d_pcm, d_out, n,
                                                                                                      channels, track);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

// accessors for pipeline.cu
rb::DeviceInfo rb_frontend_device(const rb_frontend* h) {
    return h->dev;
}
void rb_frontend_set_upload_stream(rb_frontend* h, cudaStream_t s) {
    h->uploadStream = s;
}
cudaStream_t rb_frontend_stream(const rb_frontend* h) {
    return h->stream;
}
int rb_frontend_feat_dim(const rb_frontend* h) {
    return h->featDim;
}

// ---------------------------------------------------------------------------------------------------------------
// signal-dc-detection + front-end
// ---------------------------------------------------------------------------------------------------------------
extern "C" void rb_dc_default_cfg(rb_dc_cfg* cfg) {
    if (!cfg)
        return;
    // src/Tools/FeatureExtraction/share/samples.flow:34-35 (the node's own defaults differ only in .02 for the last)
    cfg->min_dc_length_s             = 0.0125;
    cfg->max_dc_increment            = 0.9f;
    cfg->min_non_dc_segment_length_s = 0.026;
    cfg->maximal_output_size         = 4096;
}

namespace {
struct DcSizes {
    uint32_t minDc, minSeg, cut;
};
int dc_sizes(const rb_frontend* h, const rb_dc_cfg* dc, DcSizes* out) {
    RB_REQUIRE(dc->min_dc_length_s >= 0 && dc->min_non_dc_segment_length_s >= 0 && dc->max_dc_increment >= 0 &&
                       dc->maximal_output_size >= 1,
               "bad dc-detection parameters");
    out->minDc  = (uint32_t)rint(dc->min_dc_length_s * h->sampleRate);  // DcDetection::init, DcDetection.cc:75-85
    out->minSeg = (uint32_t)rint(dc->min_non_dc_segment_length_s * h->sampleRate);
    out->cut    = std::max(out->minSeg, (uint32_t)dc->maximal_output_size);
    return RB_OK;
}
// every kept run but the last of an utterance is followed by at least minDc discarded samples
long dc_max_runs(const DcSizes& z, long n) {
    return n <= 0 ? 0 : n / (long)std::max<uint32_t>(1, z.minDc + 1) + 1;
}
}  // namespace

extern "C" long rb_frontend_dc_max_frames(const rb_frontend* h, const rb_dc_cfg* dc, const int64_t* offsets, int n_utt) {
    if (!h || !dc || !offsets || n_utt < 0) {
        rb::set_error("bad argument");
        return RB_ERR_INVALID;
    }
    DcSizes z;
    if (dc_sizes(h, dc, &z) != RB_OK)
        return RB_ERR_INVALID;
    long total = 0;
    for (int u = 0; u < n_utt; ++u) {
        const long n = (long)(offsets[u + 1] - offsets[u]);
        if (n > 0)  // a run of m samples gives at most m / S + 1 frames
            total += n / h->S + dc_max_runs(z, n);
    }
    return total;
}

namespace {
int process_dc_impl(rb_frontend* h, const rb_dc_cfg* dc, const float* samples, const int64_t* offsets, int n_utt,
                    float* feats, long capacity, int64_t* frame_offsets, double* t_start, double* t_end,
                    const double* utt_start) {
    RB_REQUIRE(h && dc && offsets && frame_offsets && n_utt >= 0 && capacity >= 0, "bad argument");
    DcSizes z;
    RB_CHECK(dc_sizes(h, dc, &z));
    h->dcRunUtt.clear();
    h->dcRunBeg.clear();
    h->dcRunEnd.clear();
    h->dcRunStart.clear();
    frame_offsets[0] = 0;
    if (n_utt == 0)
        return RB_OK;
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    const int64_t base = offsets[0], nS = offsets[n_utt] - base;
    RB_REQUIRE(nS >= 0, "negative sample count");
    RB_REQUIRE(samples || nS == 0, "NULL sample buffer");
    std::vector<int64_t> off(2 * (size_t)(n_utt + 1));  // [sample ranges | run capacity ranges]
    int64_t*             rel    = off.data();
    int64_t*             runOff = off.data() + n_utt + 1;
    runOff[0]                   = 0;
    for (int u = 0; u <= n_utt; ++u) {
        rel[u] = offsets[u] - base;
        RB_REQUIRE(u == 0 || rel[u] >= rel[u - 1], "sample offsets not monotone at utterance %d", u - 1);
        if (u > 0)
            runOff[u] = runOff[u - 1] + dc_max_runs(z, (long)(rel[u] - rel[u - 1]));
    }
    const size_t runCap = (size_t)runOff[n_utt];
    cudaStream_t s      = h->stream;
    RB_CHECK(h->dSamples.reserve((size_t)nS + 8));
    RB_CHECK(h->dDcBits.reserve((size_t)(nS + 31) / 32 + 256));
    RB_CHECK(h->dDcOff.upload(off.data(), off.size(), s));
    RB_CHECK(h->dDcRunBeg.reserve(runCap));
    RB_CHECK(h->dDcRunEnd.reserve(runCap));
    RB_CHECK(h->dDcRunStart.reserve(runCap));
    RB_CHECK(h->dDcCount.reserve((size_t)n_utt + 1));
    if (nS > 0)
        RB_CUDA(cudaMemcpyAsync(h->dSamples.p, samples + base, (size_t)nS * 4, cudaMemcpyHostToDevice, s));
    RB_CUDA(cudaMemsetAsync(h->dDcCount.p, 0, sizeof(int) * ((size_t)n_utt + 1), s));
    DcParams p;
    p.x          = h->dSamples.p;
    p.uOff       = h->dDcOff.p;
    p.uStart     = nullptr;
    if (utt_start) {
        RB_CHECK(h->dDcUttStart.upload(utt_start, (size_t)n_utt, s));
        p.uStart = h->dDcUttStart.p;
    }
    p.nUtt       = n_utt;
    p.inc        = dc->max_dc_increment;
    p.minDc      = z.minDc;
    p.minSeg     = z.minSeg;
    p.cut        = z.cut;
    p.sampleRate = h->sampleRate;
    p.bits       = h->dDcBits.p;
    p.runOff     = h->dDcOff.p + n_utt + 1;
    p.runBeg     = h->dDcRunBeg.p;
    p.runEnd     = h->dDcRunEnd.p;
    p.runStart   = h->dDcRunStart.p;
    p.nRuns      = h->dDcCount.p;
    int* dViolation = h->dDcCount.p + n_utt;
    std::vector<int> counts((size_t)n_utt + 1, 0);
    // a zero min-dc-length makes every non-DC sample a block of its own: only the literal restatement covers that
    bool sequential = z.minDc == 0 || getenv("RB_DC_SEQUENTIAL") != nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        if (!sequential) {
            if (nS > 0) {
                const int grid = (int)std::min<int64_t>((nS + 8191) / 8192, (int64_t)h->dev.sm_count * 8);
                dc_flags_kernel<<<grid, 256, 0, s>>>(h->dSamples.p, nS, p.inc, p.uOff, n_utt, h->dDcBits.p, dViolation);
                RB_LAUNCH_CHECK();
            }
            dc_runs_kernel<<<(n_utt + 3) / 4, 128, 0, s>>>(p);
        }
        else
            dc_runs_sequential_kernel<<<(n_utt + 63) / 64, 64, 0, s>>>(p);
        RB_LAUNCH_CHECK();
        RB_CUDA(cudaMemcpyAsync(counts.data(), h->dDcCount.p, sizeof(int) * counts.size(), cudaMemcpyDeviceToHost, s));
        RB_CUDA(cudaStreamSynchronize(s));
        if (sequential || counts[n_utt] == 0)
            break;
        sequential = true;  // the input is not "equal or an increment apart" everywhere: replay the reference chain
    }
    h->dcSlowPath = sequential;
    // the run table comes back to the host: it shapes the launches that follow
    std::vector<int64_t> rb(runCap), re(runCap);
    std::vector<double>  rs(runCap);
    if (runCap) {
        RB_CUDA(cudaMemcpyAsync(rb.data(), h->dDcRunBeg.p, 8 * runCap, cudaMemcpyDeviceToHost, s));
        RB_CUDA(cudaMemcpyAsync(re.data(), h->dDcRunEnd.p, 8 * runCap, cudaMemcpyDeviceToHost, s));
        RB_CUDA(cudaMemcpyAsync(rs.data(), h->dDcRunStart.p, 8 * runCap, cudaMemcpyDeviceToHost, s));
        RB_CUDA(cudaStreamSynchronize(s));
    }
    std::vector<int64_t> groupOff(1, 0);
    for (int u = 0; u < n_utt; ++u) {
        RB_REQUIRE(counts[u] <= runOff[u + 1] - runOff[u], "internal: run table of utterance %d overflowed", u);
        long T = 0;
        for (int r = 0; r < counts[u]; ++r) {
            const size_t i = (size_t)runOff[u] + r;
            h->dcRunUtt.push_back(u);
            h->dcRunBeg.push_back(rb[i]);
            h->dcRunEnd.push_back(re[i]);
            h->dcRunStart.push_back(rs[i]);
            T += frames_for(h, (long)(re[i] - rb[i]));
        }
        groupOff.push_back((int64_t)h->dcRunBeg.size());
        frame_offsets[u + 1] = frame_offsets[u] + T;
    }
    const long total = (long)frame_offsets[n_utt];
    if (total > capacity) {
        rb::set_error("feature buffer holds %ld frames, %ld are needed (rb_frontend_dc_max_frames gives a bound)", capacity,
                      total);
        return RB_ERR_INVALID;
    }
    h->lastFrames = total;
    if (total == 0)
        return RB_OK;
    RB_CHECK(h->dFeats.reserve((size_t)total * h->featDim));
    const int nSeg = (int)h->dcRunBeg.size();
    RB_CHECK(run_segments(h, h->dSamples.p, h->dcRunBeg.data(), h->dcRunEnd.data(), nSeg, groupOff.data(), n_utt,
                          h->dFeats.p, s, nullptr));
    if (feats)
        RB_CUDA(cudaMemcpyAsync(feats, h->dFeats.p, (size_t)total * h->featDim * 4, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaStreamSynchronize(s));
    if (t_start || t_end) {
        // static frames: the window buffer restarts at every run with the run's start time (WindowBuffer.cc:56-59);
        // the merged packets span the delay window, which runs across the runs of an utterance
        std::vector<double> ws, we;
        for (int u = 0; u < n_utt; ++u) {
            ws.clear();
            we.clear();
            for (int64_t r = groupOff[u]; r < groupOff[u + 1]; ++r) {
                const long n = (long)(h->dcRunEnd[r] - h->dcRunBeg[r]), T = frames_for(h, n);
                double     cur = h->dcRunStart[r];
                for (long t = 0; t < T; ++t) {
                    const long len = std::min<long>(h->L, n - t * h->S);
                    ws.push_back(cur);
                    we.push_back(cur + (double)len / (double)h->sampleRate);
                    cur += (double)h->S / (double)h->sampleRate;
                }
            }
            const long T = (long)ws.size(), f0 = (long)frame_offsets[u];
            for (long t = 0; t < T; ++t) {
                double a = ws[t], b = we[t];
                if (h->cfg.derivatives)
                    for (int i = -2; i <= 2; ++i) {
                        const long v = std::min<long>(std::max<long>(t + i, 0), T - 1);
                        a            = std::min(a, ws[v]);
                        b            = std::max(b, we[v]);
                    }
                if (t_start)
                    t_start[f0 + t] = a;
                if (t_end)
                    t_end[f0 + t] = b;
            }
        }
    }
    // run positions are reported relative to the caller's buffer
    for (auto& v : h->dcRunBeg)
        v += base;
    for (auto& v : h->dcRunEnd)
        v += base;
    return RB_OK;
}
}  // namespace

extern "C" int rb_frontend_process_dc(rb_frontend* h, const rb_dc_cfg* dc, const float* samples, const int64_t* offsets,
                                      int n_utt, float* feats, long capacity, int64_t* frame_offsets, double* t_start,
                                      double* t_end) {
    return process_dc_impl(h, dc, samples, offsets, n_utt, feats, capacity, frame_offsets, t_start, t_end, nullptr);
}

extern "C" int rb_frontend_set_dc_detection(rb_frontend* h, const rb_dc_cfg* dc) {
    RB_REQUIRE(h != nullptr, "NULL handle");
    h->dcEnabled = dc != nullptr;
    if (dc) {
        DcSizes z;
        RB_CHECK(dc_sizes(h, dc, &z));
        h->dcCfg = *dc;
    }
    return RB_OK;
}

extern "C" long rb_frontend_dc_runs(const rb_frontend* h, int64_t* run_utt, int64_t* run_begin, int64_t* run_end,
                                    double* run_start, long capacity, int* sequential_path) {
    if (!h) {
        rb::set_error("NULL handle");
        return RB_ERR_INVALID;
    }
    const long n = (long)h->dcRunBeg.size();
    for (long i = 0; i < std::min(n, capacity); ++i) {
        if (run_utt)
            run_utt[i] = h->dcRunUtt[i];
        if (run_begin)
            run_begin[i] = h->dcRunBeg[i];
        if (run_end)
            run_end[i] = h->dcRunEnd[i];
        if (run_start)
            run_start[i] = h->dcRunStart[i];
    }
    if (sequential_path)
        *sequential_path = h->dcSlowPath;
    return n;
}
