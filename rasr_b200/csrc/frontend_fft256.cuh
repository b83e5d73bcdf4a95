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
//   * mel filter bank: every lane owns one contiguous piece (<= TPL taps) of ONE filter, wide filters are cut into
//     up to 4 pieces (20 filters -> 32 pieces of <= 24 taps for the standard geometry); the taps of a piece are
//     consecutive bins, so the inner loop is LDS amp / LDS weight / FFMA with immediate offsets and no control flow.
//     Lane f then adds the pieces of filter f in order, log10, and the DCT is 20 shuffles + FMAs per lane.
//   * pre-emphasis runs once per 32-frame tile (every sample is used by 2.5 frames), out of the TMA-staged raw
//     samples into an 8-byte aligned, zero-filled buffer: short last frames need no masking.
// Window (16 values), split twiddles (8) stay in registers for the whole kernel.

constexpr int kF256Warps   = 8;
constexpr int kF256Threads = kF256Warps * 32;
constexpr int kT1Stride    = 36;  // float2 units; = 4 mod 16 makes the transposed 64-bit reads conflict free
constexpr int kTransFloats = 2 * 8 * kT1Stride;            // 8 x 36 float2 (second transpose: 64 x 4 float2)
constexpr int kAmpFloats   = 257 + 67;                     // amp[257] + zero pad read by padding taps / idle lanes
constexpr int kFrameScratch = kTransFloats + kAmpFloats + 32;  // floats: transposes | amp | piece sums
constexpr int kFramesPerWarp = 2;
constexpr int kWarpScratch = kFramesPerWarp * kFrameScratch;

constexpr int kPreRounds   = 22;  // pre-emphasis pass: <= 22 samples per thread (tile span <= 5632)

struct F256Tables {  // offsets in floats inside the table blob; [0, stagedFloats) is staged into shared memory
    int stagedFloats;
    int oTwA;      // [7][32] float2: W256^{l k1}, k1 = 1..7
    int oTwB;      // [7][4]  float2: W32^{m0 j1}, j1 = 1..7
    int oMelW;     // [TPL][32] float: weight of tap i of lane's piece (0 beyond the piece)
    int oMelBin;   // [32] int: first bin of the lane's piece
    int oPiece;    // [32] int: first piece of filter f | number of pieces << 8
    int oDctT;     // [nFilters][32] float: dct[c][n] stored at [n][c]
    int oTws;      // global only: [256] float2 W512^{k}
    int oWin;      // global only: [512] window, zero padded
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

__device__ __forceinline__ float sqrt_approx(float x) {  // MUFU.SQRT: no denormal fix-up path, <= 1 ulp
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// spectrum of the real sequence from Z[k] = a, Z[M-k] = b (src/Math/FastFourierTransform.cc:108-131):
//   X[k]   = h1 + w h2,  X[M-k] = conj(h1 - w h2),  h1 = (a + conj b) / 2,  h2 = -i (a - conj b) / 2,  w = W_N^k.
// The factors 1/2 and 1/sampleRate are linear: they are applied once to the amplitude (halfScale = scale / 2).
__device__ __forceinline__ void split_pair(float2 a, float2 b, float2 w, float halfScale, float& ampK, float& ampMK) {
    const float h1R = a.x + b.x, h1I = a.y - b.y;
    const float h2R = a.y + b.y, h2I = b.x - a.x;
    const float uR  = __fmaf_rn(w.x, h2R, -__fmul_rn(w.y, h2I));
    const float uI  = __fmaf_rn(w.x, h2I, __fmul_rn(w.y, h2R));
    float       re = h1R + uR, im = h1I + uI;
    ampK  = __fmul_rn(sqrt_approx(__fmaf_rn(re, re, __fmul_rn(im, im))), halfScale);
    re    = h1R - uR;
    im    = uI - h1I;
    ampMK = __fmul_rn(sqrt_approx(__fmaf_rn(re, re, __fmul_rn(im, im))), halfScale);
}

// dynamic shared memory (floats): tables | samples (TMA-staged raw, pre-emphasised in place) | per-warp scratch
// NF: number of mel filters when known at compile time (0: runtime); TPL: mel taps per lane;
// a warp works on FR = 2 frames at a time: twiddles, mel weights and DCT rows are loaded once for both
template<int NF, int TPL>
__global__ void __launch_bounds__(kF256Threads, 2)
        mfcc_fft256_kernel(const FeParams p, const F256Tables tb, int sampleCap) {
    constexpr int FR = kFramesPerWarp;
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t bar;

    float* sTab  = smem;
    float* sSmp  = sTab + ((tb.stagedFloats + 3) & ~3);
    float* sWarp = sSmp + sampleCap + (threadIdx.x >> 5) * kWarpScratch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nFilters = NF ? NF : p.nFilters;
    // per frame slot f: transposes sT(f), amplitudes sAmp(f), mel piece sums sPart(f)
    auto sT    = [&](int f) { return reinterpret_cast<float2*>(sWarp + f * kFrameScratch); };
    auto sAmp  = [&](int f) { return sWarp + f * kFrameScratch + kTransFloats; };
    auto sPart = [&](int f) { return sWarp + f * kFrameScratch + kTransFloats + kAmpFloats; };

    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)((tb.stagedFloats + 3) & ~3) * 4u;
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(sTab, p.tables, bytes, &bar);
    }

    // lane constants: window pairs of the 8 complex points n = lane + 32 j, split twiddles of this lane's bins
    float2 win[8], tws[4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        win[j] = __ldg(reinterpret_cast<const float2*>(p.tables + tb.oWin) + lane + 32 * j);
    const int gA = lane ? lane : 32;  // radix-4 groups of this lane: gA and 64 - gA (lane 0: 32 twice)
    const int gB = 64 - gA;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        tws[j] = __ldg(reinterpret_cast<const float2*>(p.tables + tb.oTws) + gA + 64 * j);
    const float2 tws64 = __ldg(reinterpret_cast<const float2*>(p.tables + tb.oTws) + 64);
    const int    k1 = lane >> 2, m0 = lane & 3;
    // second transpose: group g at 4 g + (m ^ ((g >> 2) & 3)): 64-bit accesses conflict free both ways
    const int t2w  = 4 * k1 + (m0 ^ (k1 >> 2));           // + 32 j, swizzle term (2 j) & 3 added per j
    const int t2rA = 4 * gA, xA = (gA >> 2) & 3, t2rB = 4 * gB, xB = (gB >> 2) & 3;

    mbar_wait(&bar, phase);
    phase ^= 1;
    const float2* sTwA     = reinterpret_cast<const float2*>(sTab + tb.oTwA) + lane;
    const float2* sTwB     = reinterpret_cast<const float2*>(sTab + tb.oTwB) + m0;
    const float*  sMelW    = sTab + tb.oMelW + lane;
    const float*  sDctT    = sTab + tb.oDctT + lane;
    const int     melBin   = reinterpret_cast<const int*>(sTab + tb.oMelBin)[lane];
    const int     piece    = reinterpret_cast<const int*>(sTab + tb.oPiece)[lane];
    const int     piece0   = piece & 0xff;
    const int     nPieces  = lane < nFilters ? (piece >> 8) : 0;
    const float   halfScale = 0.5f * p.scale;
#pragma unroll
    for (int f = 0; f < FR; ++f) {
        sAmp(f)[257 + lane] = 0.0f;  // zero pad behind the spectrum: padding taps (weight 0) of the pieces read it
        sAmp(f)[257 + 32 + lane] = 0.0f;
        if (lane < kAmpFloats - 257 - 64)
            sAmp(f)[257 + 64 + lane] = 0.0f;
    }
    __syncwarp();

    for (int tileIdx = blockIdx.x; tileIdx < p.nTiles; tileIdx += gridDim.x) {
        const Tile    tile = p.tiles[tileIdx];
        const int64_t uBeg = p.sampleOff[tile.utt];
        const int64_t uLen = p.sampleEnd[tile.utt] - uBeg;
        const int64_t fOut = p.frameOff[tile.utt] + tile.f0;
        const int64_t s0   = (int64_t)tile.f0 * p.S;
        const int     span = (tile.nf - 1) * p.S + p.L;                    // samples the tile's frames cover
        const int64_t sEnd = s0 + span < uLen ? s0 + span : uLen;          // real samples end here
        const int64_t gFirst = uBeg + (s0 > 0 ? s0 - 1 : 0);
        const int64_t gLast  = uBeg + sEnd;
        const int64_t gAl    = gFirst & ~(int64_t)3;
        const int64_t gBl    = gLast & ~(int64_t)3;
        __syncthreads();  // the previous tile's frames have consumed the sample buffer
        if (threadIdx.x == 0) {
            const uint32_t bytes = gBl > gAl ? (uint32_t)(gBl - gAl) * 4u : 0u;
            if (bytes) {
                mbar_expect_tx(&bar, bytes);
                bulk_g2s(sSmp, p.samples + gAl, bytes, &bar);
            }
            else {
                mbar_arrive(&bar);
            }
        }
        if (threadIdx.x < (int)(gLast - gBl))  // ragged tail (< 4 samples) with plain loads
            sSmp[(int)(gBl - gAl) + threadIdx.x] = p.samples[gBl + threadIdx.x];
        mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        // pre-emphasis (Preemphasis.cc:51-74) in place: e[i] = x[i] - alpha x[i-1], the first sample of a segment is
        // its own predecessor; zero beyond the end of the utterance (the window node zero-pads short frames) and up
        // to the N samples every frame reads (the window is zero beyond L, but 0 * stale NaN would poison).
        // Raw sample s0 + i sits at sSmp[off + i]; e[i] goes to sSmp[i] (8-byte aligned frames).
        {
            const int off   = (int)(uBeg + s0 - gAl);
            const int nReal = (int)(sEnd - s0);
            const int spanE = (tile.nf - 1) * p.S + p.N;
            float     e[kPreRounds];
#pragma unroll
            for (int r = 0; r < kPreRounds; ++r) {
                const int i = threadIdx.x + kF256Threads * r;
                float     d = 0.0f;
                if (i < nReal) {
                    const float cur  = sSmp[off + i];
                    const float prev = (s0 + i > 0) ? sSmp[off + i - 1] : cur;
                    d                = p.alpha == 1.0f ? __fsub_rn(cur, prev) : __fmaf_rn(-p.alpha, prev, cur);
                }
                e[r] = d;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < kPreRounds; ++r) {
                const int i = threadIdx.x + kF256Threads * r;
                if (i < spanE)
                    sSmp[i] = e[r];
            }
        }
        __syncthreads();

        for (int fi = warp * FR; fi < tile.nf; fi += kF256Warps * FR) {
            float2 v[FR][8];
            // ---- window; lane = n0, register = n1.  A frame slot beyond the tile recomputes the last frame
#pragma unroll
            for (int f = 0; f < FR; ++f) {
                const int     ff = fi + f < tile.nf ? fi + f : tile.nf - 1;
                const float2* e2 = reinterpret_cast<const float2*>(sSmp + ff * p.S) + lane;  // S is even
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 e = e2[32 * j];
                    v[f][j]        = make_float2(__fmul_rn(win[j].x, e.x), __fmul_rn(win[j].y, e.y));
                }
            }
            // ---- radix-8 over n1, twiddle W256^{n0 k1}, transpose: (n0, k1) -> lane 4 k1 + m0, register m1
#pragma unroll
            for (int f = 0; f < FR; ++f)
                dft8(v[f]);
#pragma unroll
            for (int k = 1; k < 8; ++k) {
                const float2 w = sTwA[(k - 1) * 32];
#pragma unroll
                for (int f = 0; f < FR; ++f)
                    v[f][k] = cmul(v[f][k], w);
            }
#pragma unroll
            for (int f = 0; f < FR; ++f)
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    sT(f)[k * kT1Stride + lane] = v[f][k];
            __syncwarp();
#pragma unroll
            for (int f = 0; f < FR; ++f)
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    v[f][m] = sT(f)[k1 * kT1Stride + m0 + 4 * m];
            __syncwarp();
            // ---- radix-8 over m1, twiddle W32^{m0 j1}, transpose: group g = k1 + 8 j1 holds its four m0
#pragma unroll
            for (int f = 0; f < FR; ++f)
                dft8(v[f]);
#pragma unroll
            for (int j = 1; j < 8; ++j) {
                const float2 w = sTwB[(j - 1) * 4];
#pragma unroll
                for (int f = 0; f < FR; ++f)
                    v[f][j] = cmul(v[f][j], w);
            }
#pragma unroll
            for (int f = 0; f < FR; ++f)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sT(f)[(t2w ^ ((2 * j) & 3)) + 32 * j] = v[f][j];
            __syncwarp();
#pragma unroll
            for (int f = 0; f < FR; ++f)
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    v[f][m]     = sT(f)[t2rA + (m ^ xA)];
                    v[f][4 + m] = sT(f)[t2rB + (m ^ xB)];
                }
#pragma unroll
            for (int f = 0; f < FR; ++f) {
                // ---- radix-4 over m0: v[j] = Z[gA + 64 j], v[4 + j] = Z[gB + 64 j]
                dft4(v[f][0], v[f][1], v[f][2], v[f][3]);
                dft4(v[f][4], v[f][5], v[f][6], v[f][7]);
                // ---- real split + amplitude: bin k = gA + 64 j pairs with 256 - k = gB + 64 (3 - j)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float aK, aMK;
                    split_pair(v[f][j], v[f][4 + (3 - j)], tws[j], halfScale, aK, aMK);
                    sAmp(f)[gA + 64 * j]       = aK;
                    sAmp(f)[256 - gA - 64 * j] = aMK;
                }
            }
            if (lane == 0) {  // group 0: bins 0 / 256 (no partner), 64 <-> 192, 128 (its own partner)
#pragma unroll
                for (int f = 0; f < FR; ++f) {
                    float2 z[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        z[m] = sT(f)[m];
                    dft4(z[0], z[1], z[2], z[3]);
                    sAmp(f)[0]   = fabsf(__fmul_rn(__fadd_rn(z[0].x, z[0].y), p.scale));
                    sAmp(f)[256] = fabsf(__fmul_rn(__fsub_rn(z[0].x, z[0].y), p.scale));
                    sAmp(f)[128] =
                            __fmul_rn(sqrt_approx(__fmaf_rn(z[2].x, z[2].x, __fmul_rn(z[2].y, z[2].y))), p.scale);
                    float aK, aMK;
                    split_pair(z[1], z[3], tws64, halfScale, aK, aMK);
                    sAmp(f)[64]  = aK;
                    sAmp(f)[192] = aMK;
                }
            }
            __syncwarp();
            if (p.dbgAmp)
                for (int f = 0; f < FR; ++f)
                    if (fi + f < tile.nf)
                        for (int k = lane; k < p.nBins; k += 32)
                            p.dbgAmp[(fOut + fi + f) * p.nBins + k] = sAmp(f)[k];
            // ---- mel filter bank: this lane's piece of one filter, then lane f adds the pieces of filter f
            {
                float acc[FR];
#pragma unroll
                for (int f = 0; f < FR; ++f)
                    acc[f] = 0.0f;
#pragma unroll
                for (int i = 0; i < TPL; ++i) {
                    const float w = sMelW[i * 32];
#pragma unroll
                    for (int f = 0; f < FR; ++f)
                        acc[f] = __fmaf_rn(sAmp(f)[melBin + i], w, acc[f]);
                }
#pragma unroll
                for (int f = 0; f < FR; ++f)
                    sPart(f)[lane] = acc[f];
            }
            __syncwarp();
            float fbv[FR], r[FR];
#pragma unroll
            for (int f = 0; f < FR; ++f) {
                fbv[f] = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < nPieces)
                        fbv[f] = __fadd_rn(fbv[f], sPart(f)[piece0 + i]);
                if (p.dbgFbank && lane < nFilters && fi + f < tile.nf)
                    p.dbgFbank[(fOut + fi + f) * nFilters + lane] = fbv[f];
                fbv[f] = log10f(fbv[f]);
                r[f]   = 0.0f;
            }
            // ---- DCT-II: lane c accumulates dct[c][n] * fb[n], fb[n] lives in lane n
            if (NF) {
#pragma unroll
                for (int n = 0; n < (NF ? NF : 1); ++n) {
                    const float d = sDctT[n * 32];
#pragma unroll
                    for (int f = 0; f < FR; ++f)
                        r[f] = __fmaf_rn(d, __shfl_sync(0xffffffffu, fbv[f], n), r[f]);
                }
            }
            else {
                for (int n = 0; n < nFilters; ++n) {
                    const float d = sDctT[n * 32];
#pragma unroll
                    for (int f = 0; f < FR; ++f)
                        r[f] = __fmaf_rn(d, __shfl_sync(0xffffffffu, fbv[f], n), r[f]);
                }
            }
#pragma unroll
            for (int f = 0; f < FR; ++f)
                if (lane < p.nCep && fi + f < tile.nf) {
                    const int64_t t = fOut + fi + f;
                    p.cep[t * p.nCep + lane] = r[f];
                    if (!p.derivatives)
                        p.feats[t * p.featDim + lane] = r[f];
                }
            __syncwarp();  // scratch is reused by the next frames
        }
    }
}
