// frontend_fft256.cuh -- register-resident MFCC kernel for the standard 16 kHz geometry (512-point real FFT).
// Included by frontend.cu (inside its anonymous namespace, after FeParams / Tile).
//
// Same arithmetic chain as mfcc_static_kernel (pre-emphasis, Hamming window, N/2-point complex FFT of the
// even/odd packed frame with e^{+i theta} twiddles, real split, 1/sampleRate scale, amplitude, mel taps, log10,
// DCT-II; reference lines cited in frontend.cu), different schedule:
//   * a warp owns a frame; the 256 complex points live in registers, 8 per lane.  The transform is the
//     Cooley-Tukey factorisation 256 = 8 x 8 x 4: radix-8 over n1 (n = n0 + 32 n1, lane = n0), twiddle W256^{n0 k1},
//     transpose through shared memory, radix-8 over m1 (n0 = m0 + 4 m1, lane = 4 k1 + m0), twiddle W32^{m0 j1},
//     transpose, radix-4 over m0.  Two shared-memory transposes replace the eight read-modify-write passes of the
//     radix-2 kernel.
//   * after the second transpose lane l holds the radix-4 groups g = l and 64 - l, i.e. the spectrum bins
//     k = l + 64 j and their mirror images 256 - k: the real split and the amplitude need no further exchange.
//   * the 479 mel taps are dealt to the 32 lanes in contiguous runs of <= 16 (filter-major order); a lane leaves one
//     partial sum per filter it touches, lane f adds the <= 6 partials of filter f, log10, and the DCT is 20
//     shuffles + FMAs per lane.
//   * pre-emphasis runs once per 32-frame tile (every sample is used by 2.5 frames), out of the TMA-staged raw
//     samples into an 8-byte aligned, zero-filled buffer: short last frames need no masking.
// Window (16 values), split twiddles (8) stay in registers for the whole kernel.

constexpr int kF256Warps   = 8;
constexpr int kF256Threads = kF256Warps * 32;
constexpr int kTapsPerLane = 16;
constexpr int kT1Stride    = 36;  // float2 units; = 4 mod 16 makes the transposed 64-bit reads conflict free
constexpr int kT2Stride    = 5;   // float2 units per radix-4 group (odd: conflict-free 64-bit reads)
constexpr int kTransFloats = 2 * 64 * kT2Stride;           // 64 groups x 5 float2 (>= 8 x 36 float2)
constexpr int kWarpScratch = kTransFloats + 264 + 72;     // floats: transposes | amp[257] | partial sums

struct F256Tables {  // offsets in floats inside the table blob
    int oTwA;      // [7][32] float2: W256^{l k1}, k1 = 1..7
    int oTwB;      // [7][4]  float2: W32^{m0 j1}, j1 = 1..7
    int oTws;      // [256]   float2: W512^{k}
    int oWin;      // [512]   window, zero padded
    int oMelMeta;  // [16][32] int: bin | lastOfSegment << 16 | partialIndex << 20
    int oMelW;     // [16][32] float
    int oPartOff;  // [33] int: partials of filter f are [partOff[f], partOff[f+1])
    int oDctT;     // [nFilters][32] float: dct[c][n] stored at [n][c]
    int maxPart;   // longest partial list of a filter
};

__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return make_float2(__fmaf_rn(a.x, w.x, -__fmul_rn(a.y, w.y)), __fmaf_rn(a.x, w.y, __fmul_rn(a.y, w.x)));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    return make_float2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    return make_float2(a.x - b.x, a.y - b.y);
}
__device__ __forceinline__ float2 cmul_i(float2 a) {  // a * i
    return make_float2(-a.y, a.x);
}

// 4-point DFT with e^{+i 2 pi jk/4}, in place, natural order
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s0 = cadd(a0, a2), s1 = csub(a0, a2), s2 = cadd(a1, a3), s3 = cmul_i(csub(a1, a3));
    a0 = cadd(s0, s2);
    a2 = csub(s0, s2);
    a1 = cadd(s1, s3);
    a3 = csub(s1, s3);
}

// 8-point DFT with e^{+i 2 pi jk/8}, in place, natural order
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    const float r = 0.70710678118654752440f;
    float2      a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    float2      b0 = csub(v[0], v[4]), t1 = csub(v[1], v[5]), t2 = csub(v[2], v[6]), t3 = csub(v[3], v[7]);
    float2      b1 = make_float2((t1.x - t1.y) * r, (t1.x + t1.y) * r);    // * (1+i)/sqrt2
    float2      b2 = cmul_i(t2);                                           // * i
    float2      b3 = make_float2((-t3.x - t3.y) * r, (t3.x - t3.y) * r);   // * (-1+i)/sqrt2
    dft4(a0, a1, a2, a3);
    dft4(b0, b1, b2, b3);
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// spectrum of the real sequence from Z[k], Z[M-k] (same formulas as mfcc_static_kernel), scaled amplitudes
__device__ __forceinline__ void split_pair(float2 a, float2 b, float2 w, float scale, float& ampK, float& ampMK) {
    const float h1R = 0.5f * (a.x + b.x), h1I = 0.5f * (a.y - b.y);
    const float h2R = 0.5f * (a.y + b.y), h2I = -0.5f * (a.x - b.x);
    const float uR  = __fmaf_rn(w.x, h2R, -__fmul_rn(w.y, h2I));
    const float uI  = __fmaf_rn(w.x, h2I, __fmul_rn(w.y, h2R));
    float       re = __fmul_rn(h1R + uR, scale), im = __fmul_rn(h1I + uI, scale);
    ampK  = __fsqrt_rn(__fmaf_rn(re, re, __fmul_rn(im, im)));
    re    = __fmul_rn(h1R - uR, scale);
    im    = __fmul_rn(uI - h1I, scale);
    ampMK = __fsqrt_rn(__fmaf_rn(re, re, __fmul_rn(im, im)));
}

// dynamic shared memory (floats): tables | raw samples (TMA) | pre-emphasised samples | per-warp scratch
__global__ void __launch_bounds__(kF256Threads, 2)
        mfcc_fft256_kernel(const FeParams p, const F256Tables tb, int sampleCap) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar;

    float* sTab  = smem;
    float* sRaw  = sTab + ((p.tableFloats + 3) & ~3);
    float* sEmph = sRaw + sampleCap;
    float* sWarp = sEmph + sampleCap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2*   sT   = reinterpret_cast<float2*>(sWarp + warp * kWarpScratch);
    float*    sAmp = sWarp + warp * kWarpScratch + kTransFloats;
    float*    sPart = sAmp + 264;

    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)p.tableFloats * 4u;
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(sTab, p.tables, bytes, &bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;

    const float2* sTwA     = reinterpret_cast<const float2*>(sTab + tb.oTwA);
    const float2* sTwB     = reinterpret_cast<const float2*>(sTab + tb.oTwB);
    const int*    sMelMeta = reinterpret_cast<const int*>(sTab + tb.oMelMeta);
    const float*  sMelW    = sTab + tb.oMelW;
    const int*    sPartOff = reinterpret_cast<const int*>(sTab + tb.oPartOff);
    const float*  sDctT    = sTab + tb.oDctT;

    // lane constants: window pairs of the 8 complex points n = lane + 32 j, split twiddles of this lane's bins
    float2 win[8], tws[4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        win[j] = reinterpret_cast<const float2*>(sTab + tb.oWin)[lane + 32 * j];
    const int gA = lane ? lane : 32;  // radix-4 groups of this lane: gA and 64 - gA (lane 0: 32 twice)
#pragma unroll
    for (int j = 0; j < 4; ++j)
        tws[j] = reinterpret_cast<const float2*>(sTab + tb.oTws)[gA + 64 * j];
    const int k1 = lane >> 2, m0 = lane & 3;
    const int partBeg = sPartOff[lane < p.nFilters ? lane : 0];
    const int partEnd = lane < p.nFilters ? sPartOff[lane + 1] : partBeg;

    for (int tileIdx = blockIdx.x; tileIdx < p.nTiles; tileIdx += gridDim.x) {
        const Tile    tile = p.tiles[tileIdx];
        const int64_t uBeg = p.sampleOff[tile.utt];
        const int64_t uLen = p.sampleOff[tile.utt + 1] - uBeg;
        const int64_t fOut = p.frameOff[tile.utt] + tile.f0;
        const int64_t s0   = (int64_t)tile.f0 * p.S;
        const int     span = (tile.nf - 1) * p.S + p.L;                    // samples the tile's frames cover
        const int64_t sEnd = s0 + span < uLen ? s0 + span : uLen;          // real samples end here
        const int64_t gFirst = uBeg + (s0 > 0 ? s0 - 1 : 0);
        const int64_t gLast  = uBeg + sEnd;
        const int64_t gAl    = gFirst & ~(int64_t)3;
        const int64_t gBl    = gLast & ~(int64_t)3;
        __syncthreads();  // the previous tile's pre-emphasis pass has consumed sRaw, its frames have consumed sEmph
        if (threadIdx.x == 0) {
            const uint32_t bytes = gBl > gAl ? (uint32_t)(gBl - gAl) * 4u : 0u;
            if (bytes) {
                mbar_expect_tx(&bar, bytes);
                bulk_g2s(sRaw, p.samples + gAl, bytes, &bar);
            }
            else {
                mbar_arrive(&bar);
            }
        }
        if (threadIdx.x < (int)(gLast - gBl))  // ragged tail (< 4 samples) with plain loads
            sRaw[(int)(gBl - gAl) + threadIdx.x] = p.samples[gBl + threadIdx.x];
        mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        // pre-emphasis (Preemphasis.cc:51-74): e[i] = x[i] - alpha x[i-1], the first sample of a segment is its own
        // predecessor; zero beyond the end of the utterance (the window node zero-pads short frames)
        {
            const int off   = (int)(uBeg + s0 - gAl);  // sRaw index of utterance sample s0
            const int nReal = (int)(sEnd - s0);
            // the frames read N samples each (the window table is zero beyond L, but 0 * stale NaN would poison)
            const int spanE = (tile.nf - 1) * p.S + p.N;
            for (int i = threadIdx.x; i < spanE; i += kF256Threads) {
                float e = 0.0f;
                if (i < nReal) {
                    const float cur  = sRaw[off + i];
                    const float prev = (s0 + i > 0) ? sRaw[off + i - 1] : cur;
                    e                = p.alpha == 1.0f ? __fsub_rn(cur, prev) : __fmaf_rn(-p.alpha, prev, cur);
                }
                sEmph[i] = e;
            }
        }
        __syncthreads();

        for (int fi = warp; fi < tile.nf; fi += kF256Warps) {
            const float2* e2 = reinterpret_cast<const float2*>(sEmph + fi * p.S);  // S is even on this path
            float2        v[8];
            // ---- window; lane = n0, register = n1
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 e = e2[lane + 32 * j];
                v[j]           = make_float2(__fmul_rn(win[j].x, e.x), __fmul_rn(win[j].y, e.y));
            }
            // ---- radix-8 over n1, twiddle W256^{n0 k1}, transpose: (n0, k1) -> lane 4 k1 + m0, register m1
            dft8(v);
#pragma unroll
            for (int k = 1; k < 8; ++k)
                v[k] = cmul(v[k], sTwA[(k - 1) * 32 + lane]);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                sT[k * kT1Stride + lane] = v[k];
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 8; ++m)
                v[m] = sT[k1 * kT1Stride + m0 + 4 * m];
            __syncwarp();
            // ---- radix-8 over m1, twiddle W32^{m0 j1}, transpose: group g = k1 + 8 j1 holds its four m0
            dft8(v);
#pragma unroll
            for (int j = 1; j < 8; ++j)
                v[j] = cmul(v[j], sTwB[(j - 1) * 4 + m0]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                sT[(k1 + 8 * j) * kT2Stride + m0] = v[j];
            __syncwarp();
            const int gB = 64 - gA;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                v[m]     = sT[gA * kT2Stride + m];
                v[4 + m] = sT[gB * kT2Stride + m];
            }
            // ---- radix-4 over m0: v[j] = Z[gA + 64 j], v[4 + j] = Z[gB + 64 j]
            dft4(v[0], v[1], v[2], v[3]);
            dft4(v[4], v[5], v[6], v[7]);
            // ---- real split + amplitude: bin k = gA + 64 j pairs with 256 - k = gB + 64 (3 - j)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float aK, aMK;
                split_pair(v[j], v[4 + (3 - j)], tws[j], p.scale, aK, aMK);
                sAmp[gA + 64 * j]       = aK;
                sAmp[256 - gA - 64 * j] = aMK;
            }
            if (lane == 0) {  // group 0: bins 0 / 256 (no partner), 64 <-> 192, 128 (its own partner)
                float2 z[4];
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    z[m] = sT[m];
                dft4(z[0], z[1], z[2], z[3]);
                sAmp[0]   = fabsf(__fmul_rn(__fadd_rn(z[0].x, z[0].y), p.scale));
                sAmp[256] = fabsf(__fmul_rn(__fsub_rn(z[0].x, z[0].y), p.scale));
                const float re = __fmul_rn(z[2].x, p.scale), im = __fmul_rn(z[2].y, p.scale);
                sAmp[128] = __fsqrt_rn(__fmaf_rn(re, re, __fmul_rn(im, im)));
                float aK, aMK;
                split_pair(z[1], z[3], reinterpret_cast<const float2*>(sTab + tb.oTws)[64], p.scale, aK, aMK);
                sAmp[64]  = aK;
                sAmp[192] = aMK;
            }
            __syncwarp();
            const int64_t t = fOut + fi;
            if (p.dbgAmp)
                for (int k = lane; k < p.nBins; k += 32)
                    p.dbgAmp[t * p.nBins + k] = sAmp[k];
            // ---- mel filter bank: this lane's run of taps, one partial sum per filter touched
            {
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < kTapsPerLane; ++i) {
                    const int meta = sMelMeta[i * 32 + lane];
                    acc            = __fmaf_rn(sAmp[meta & 0xffff], sMelW[i * 32 + lane], acc);
                    if (meta & 0x10000) {
                        sPart[meta >> 20] = acc;
                        acc               = 0.0f;
                    }
                }
            }
            __syncwarp();
            float fbv = 0.0f;
            for (int i = 0; i < tb.maxPart; ++i)
                if (partBeg + i < partEnd)
                    fbv = __fadd_rn(fbv, sPart[partBeg + i]);
            if (p.dbgFbank && lane < p.nFilters)
                p.dbgFbank[t * p.nFilters + lane] = fbv;
            fbv = log10f(fbv);
            // ---- DCT-II: lane c accumulates dct[c][n] * fb[n], fb[n] lives in lane n
            float r = 0.0f;
            for (int n = 0; n < p.nFilters; ++n)
                r = __fmaf_rn(sDctT[n * 32 + lane], __shfl_sync(0xffffffffu, fbv, n), r);
            if (lane < p.nCep) {
                p.cep[t * p.nCep + lane] = r;
                if (!p.derivatives)
                    p.feats[t * p.featDim + lane] = r;
            }
            __syncwarp();  // scratch is reused by the next frame
        }
    }
}
