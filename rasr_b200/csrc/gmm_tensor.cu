// gmm_tensor.cu -- RB_GMM_BATCH_TENSOR: the pooled-covariance max-approximation GMM scorer
// (Mm::BatchFloatFeatureScorer, src/Mm/BatchFeatureScorer.cc:164-253) as ONE tcgen05 GEMM with the
// min over the densities of a mixture fused into the TMEM epilogue.
//
//   score[t][m] = 0.5 * ( |x|^2 + min_{k in m} ( c_k + |mu_k|^2 - 2 x.mu_k ) ),   x, mu scaled by 1/sigma
//
// The direct form costs 2 CUDA-core ops per (frame, density, dim) and tops out at ~2 % of the HBM
// roofline.  Here the inner product runs on the tensor cores in SPLIT PRECISION so
// that f32 accuracy survives the fp16 operands:
//   x = xh + xl,  -2 mu = mh + ml   (fp16 pairs: 22 significant bits each)
//   A row  = [ xh | xh | xl | 1 1 1 0.. ]      (K = 3*40 + 8 = 128 for D = 39)
//   B row  = [ mh | ml | mh | c1 c2 c3 0.. ]   (c1+c2+c3 = c_k + |mu_k|^2, three fp16 terms)
//   A.B^T  = xh.mh + xh.ml + xl.mh + (c_k + |mu_k|^2)   -- dropped term xl.ml ~ 2^-22 relative
// Both x and mu are first centred on the mean of all mu (distances are translation invariant), which
// keeps |x|^2 and the cross term small, so the cancellation in the expanded form costs < 1e-6
// relative on realistic data.  Scores agree with the reference within 1e-4 relative (north_star
// tolerance; measured ~1e-6), but are not bit-identical -- RB_GMM_BATCH_FLOAT is.
//
// The same product also serves RB_GMM_BATCH_FLOAT (the exact scorer) on large batches: with the EpiGmmScreen epilogue
// the kernel does not emit the minimum but, per (frame, mixture), the set of densities within a proven error bound of
// it; gmm_refine_kernel (gmm.cu) evaluates those in the reference's operation order (rb_gmm_tensor_screen below,
// DESIGN.md 4.1b).
#include <cfloat>
#include <cmath>

#include "gemm_sm100.cuh"

namespace {

using namespace rbdev;

// ---- epilogue: min over column segments -----------------------------------------------------
// SEG > 0: every mixture has exactly SEG densities (SEG in {8,16,32}), segments never straddle a chunk.
// SEG == 0: ragged mixtures described by per-chunk end masks; segments never straddle a 256-column tile.
template<int N>
__device__ __forceinline__ float tree_min(const float* v) {
    if constexpr (N == 1)
        return v[0];
    else
        return fminf(tree_min<N / 2>(v), tree_min<N / 2>(v + N / 2));
}

template<int SEG>
struct EpiGmmMin {
    const uint32_t* endMask;   // [nChunks] bit j: column 32*c+j is the last density of its mixture
    const int*      mixStart;  // [nChunks] mixture that the first end flag of the chunk closes
    const float*    xnorm;     // [T] |x|^2 of the centred, scaled feature
    float*          scores;    // [T * nMix]
    int             nMix;
    float           invS2;     // 1 / s^2 (power of two), undoes the operand scaling
    static constexpr bool kUniform = SEG > 0;
    struct State {
        float best, xn;
    };
    __device__ void begin(State& st, int row) const {
        st.best = FLT_MAX;
        st.xn   = __ldg(xnorm + row);
    }
    // uniform mixtures: 64 columns at once, balanced min trees (no long dependent chains), 16-byte stores
    __device__ void emit64(const State& st, int row, int col0, const float (&v)[64]) const {
        constexpr int S  = SEG > 0 ? SEG : 32;
        constexpr int NO = 64 / S;
        const int     m0 = col0 / S;
        if (m0 >= nMix)
            return;
        float out[NO];
#pragma unroll
        for (int g = 0; g < NO; ++g)
            out[g] = 0.5f * __fmaf_rn(tree_min<S>(v + g * S), invS2, st.xn);
        float* dst = scores + (size_t)row * nMix + m0;
        if ((nMix % NO) == 0) {
            if constexpr (NO == 2)
                *reinterpret_cast<float2*>(dst) = make_float2(out[0], out[1]);
            else {
#pragma unroll
                for (int g = 0; g < NO; g += 4)
                    *reinterpret_cast<float4*>(dst + g) = make_float4(out[g], out[g + 1], out[g + 2], out[g + 3]);
            }
        }
        else {
#pragma unroll
            for (int g = 0; g < NO; ++g)
                if (m0 + g < nMix)
                    dst[g] = out[g];
        }
    }
    __device__ void chunk(State& st, int row, int col0, const float (&v)[32]) const {
        if (SEG > 0) {
            constexpr int S = SEG > 0 ? SEG : 32;
            const int     m0 = col0 / S;
            if (m0 >= nMix)
                return;
            float out[32 / S];
#pragma unroll
            for (int g = 0; g < 32 / S; ++g) {
                float b = v[g * S];
#pragma unroll
                for (int j = 1; j < S; ++j)
                    b = fminf(b, v[g * S + j]);
                out[g] = 0.5f * __fmaf_rn(b, invS2, st.xn);
            }
            float* dst = scores + (size_t)row * nMix + m0;
            if (S == 32) {
                dst[0] = out[0];
            }
            else if (S == 16) {
                if ((nMix & 1) == 0)
                    *reinterpret_cast<float2*>(dst) = make_float2(out[0], out[32 / S - 1]);
                else {
                    dst[0] = out[0];
                    if (m0 + 1 < nMix)
                        dst[1] = out[32 / S - 1];
                }
            }
            else {
                if ((nMix & 3) == 0)
                    *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1 % (32 / S)], out[2 % (32 / S)], out[3 % (32 / S)]);
                else {
#pragma unroll
                    for (int g = 0; g < 32 / S; ++g)
                        if (m0 + g < nMix)
                            dst[g] = out[g];
                }
            }
        }
        else {
            const int      c    = col0 >> 5;
            const uint32_t mask = __ldg(endMask + c);
            int            mix  = __ldg(mixStart + c);
            if ((col0 & (rbgemm::BN - 1)) == 0)
                st.best = FLT_MAX;  // a new tile row: nothing carries over
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                st.best = fminf(st.best, v[j]);
                if ((mask >> j) & 1u) {
                    scores[(size_t)row * nMix + mix] = 0.5f * __fmaf_rn(st.best, invS2, st.xn);
                    ++mix;
                    st.best = FLT_MAX;
                }
            }
        }
    }
};

// ---- screening epilogue (exact batch-float scoring) ------------------------------
// Instead of the minimum itself, every (frame, mixture) gets the SET of densities whose approximate score lies within
// thr[frame] of the smallest one -- one bit per density of the mixture (mixtures of at most 32 densities).  thr bounds
// twice the error of the split-precision product plus twice the rounding error of the reference's own f32 sum, so the
// density that wins in the reference's arithmetic is always in the set; gmm_refine_kernel (gmm.cu) then evaluates
// only those densities in the reference's operation order.  thr = +inf (frame with non-finite or out-of-range
// values): every density is a candidate.  The words go where the scores will be (same buffer, overwritten in place).
template<int SEG>
struct EpiGmmScreen {
    const uint32_t* endMask;
    const int*      mixStart;
    const float*    thr;    // [T] in accumulator units
    uint32_t*       masks;  // [nMix / 4][pitch][4]: the words of mixtures 4q .. 4q+3 of frame t form one uint4 at (q, t) --
                            // both this epilogue (lane = frame) and gmm_refine_kernel (thread = frame) access it coalesced
    long            pitch;
    int             nMix;
    static constexpr bool kUniform = SEG > 0;
    struct State {
        float    best, thr;
        uint32_t mask;
        int      pos;
    };
    __device__ __forceinline__ uint32_t* word(int row, int mix) const {
        return masks + (((size_t)(mix >> 2) * (size_t)pitch + (size_t)row) << 2) + (mix & 3);
    }
    __device__ void begin(State& st, int row) const {
        st.best = FLT_MAX;
        st.thr  = __ldg(thr + row);
        st.mask = 0;
        st.pos  = 0;
    }
    __device__ void emit64(const State& st, int row, int col0, const float (&v)[64]) const {
        constexpr int S  = SEG > 0 ? SEG : 32;
        constexpr int NO = 64 / S;
        const int     m0 = col0 / S;
        if (m0 >= nMix)
            return;
        constexpr uint32_t kAll = S == 32 ? 0xffffffffu : ((1u << (S & 31)) - 1u);
        const bool         all  = !(st.thr < FLT_MAX);
        uint32_t           out[NO];
#pragma unroll
        for (int g = 0; g < NO; ++g) {
            // v <= lim  <=>  lim - v >= 0  <=>  the sign bit of (lim - v) is clear (the difference of equal values is
            // +0).  The subtraction runs on the FMA pipe, one funnel shift per column collects the sign bits: the
            // epilogue warps are bound by the half-rate ALU pipe, where compare + select + or cost three slots.
            const float lim = tree_min<S>(v + g * S) + st.thr;
            uint32_t    mk  = 0;
#pragma unroll
            for (int j = S - 1; j >= 0; --j)
                mk = __funnelshift_l(__float_as_uint(__fsub_rn(lim, v[g * S + j])), mk, 1);
            mk     = ~mk & kAll;
            out[g] = (all || mk == 0) ? kAll : mk;
        }
        // nMix % 4 == 0 (a condition of the exact route), so whole quads are always inside the matrix
        if constexpr (NO == 2)
            *reinterpret_cast<uint2*>(word(row, m0)) = make_uint2(out[0], out[1]);
        else {
#pragma unroll
            for (int g = 0; g < NO; g += 4)
                if (m0 + g < nMix)
                    *reinterpret_cast<uint4*>(word(row, m0 + g)) = make_uint4(out[g], out[g + 1], out[g + 2], out[g + 3]);
        }
    }
    // ragged mixtures: one pass over the columns.  A density enters the set if it lies within thr of the running
    // minimum; the set is emptied when a new minimum undercuts the old one by more than thr (nothing seen before can
    // then be within thr of the final minimum).  The result is a superset of the exact candidate set.
    __device__ void chunk(State& st, int row, int col0, const float (&v)[32]) const {
        if (SEG > 0)
            return;
        const int      c    = col0 >> 5;
        const uint32_t ends = __ldg(endMask + c);
        int            mix  = __ldg(mixStart + c);
        if ((col0 & (rbgemm::BN - 1)) == 0) {
            st.best = FLT_MAX;
            st.mask = 0;
            st.pos  = 0;
        }
        const bool all = !(st.thr < FLT_MAX);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float    x   = v[j];
            const uint32_t bit = 1u << (st.pos & 31);
            if (x < st.best - st.thr) {
                st.mask = bit;
                st.best = x;
            }
            else if (x <= st.best + st.thr) {
                st.mask |= bit;
                st.best = fminf(st.best, x);
            }
            ++st.pos;
            if ((ends >> j) & 1u) {
                const uint32_t full = st.pos >= 32 ? 0xffffffffu : ((1u << st.pos) - 1u);
                *word(row, mix) = (all || st.mask == 0) ? full : st.mask;
                ++mix;
                st.best = FLT_MAX;
                st.mask = 0;
                st.pos  = 0;
            }
        }
    }
};

// ---- features -> split fp16 A operand + |x|^2 -------------------------------------------------
// row layout [ xh(dp) | xh(dp) | xl(dp) | 1 1 1 0... ] padded to kPad halves.  8 lanes per frame, a
// lane owns 8 consecutive dims and writes whole 16-byte chunks (dp <= 64).
__global__ void __launch_bounds__(256) gmm_split_features_kernel(const float* __restrict__ feats,
                                                                 const float* __restrict__ isd,
                                                                 const float* __restrict__ centre, long T, int dim,
                                                                 int dp, int kPad, float scale, __half* __restrict__ A,
                                                                 float* __restrict__ xnorm, float* __restrict__ thr,
                                                                 float thrA, float thrB, float* __restrict__ xT,
                                                                 long pitch) {
    const int  sub = threadIdx.x & 7;  // lane within the frame group
    const long g0  = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 3;
    const long nG  = ((long)gridDim.x * blockDim.x) >> 3;
    const int  nChunk = dp >> 3;  // 16-byte chunks per operand copy
    for (long t0 = g0; t0 < ((T + 3) & ~3L); t0 += nG) {  // whole warps stay converged for the shuffles
        const long t   = t0 < T ? t0 : T - 1;
        float      acc = 0.0f;
        uint32_t   hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
        int        bad = 0;  // non-finite or outside the fp16 range: the screening must not trust this frame
        if (sub < nChunk) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d = sub * 8 + j;
                float     x = 0.0f, xr = 0.0f;
                if (d < dim) {
                    xr = __fmul_rn(__ldg(feats + (size_t)t * dim + d), __ldg(isd + d));  // the reference's scaled feature
                    x  = __fsub_rn(xr, __ldg(centre + d));
                }
                if (xT && t0 < T)
                    xT[(size_t)d * pitch + t] = xr;
                acc      = __fmaf_rn(x, x, acc);
                bad |= !(fabsf(x * scale) <= 60000.0f);
                float xs = fminf(fmaxf(x * scale, -60000.0f), 60000.0f);
                const __half h = __float2half_rn(xs);
                const __half l = __float2half_rn(xs - __half2float(h));
                hi[j >> 1] |= (uint32_t)__half_as_ushort(h) << ((j & 1) * 16);
                lo[j >> 1] |= (uint32_t)__half_as_ushort(l) << ((j & 1) * 16);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 1);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 2);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 4);
        if (t0 < T) {
            uint4* row = reinterpret_cast<uint4*>(A + (size_t)t * kPad);
            if (sub < nChunk) {
                const uint4 h4 = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                row[sub]              = h4;
                row[nChunk + sub]     = h4;
                row[2 * nChunk + sub] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            // tail chunks: three ones (0x3C00) then zeros
            for (int c = 3 * nChunk + sub; c < (kPad >> 3); c += 8)
                row[c] = c == 3 * nChunk ? make_uint4(0x3C003C00u, 0x00003C00u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
            if (sub == 0) {
                xnorm[t] = acc;
                if (thr)
                    thr[t] = (bad || !(acc < FLT_MAX)) ? __int_as_float(0x7f800000) : __fmaf_rn(acc, thrA, thrB);
            }
        }
    }
}

// ---- diagonal scorers (per-density covariance): features -> split fp16 A operand + threshold -------------------------
//   dist_k(x) = sum_d v_kd^2 (x_d - mu_kd)^2 = sum_d  xc_d^2 v_kd^2  -  2 xc_d (v_kd^2 mu_c,kd)  +  v_kd^2 mu_c,kd^2
// (xc, mu_c centred on the mean of all means).  A row = [ X2h | X2h | X2l | X1h | X1h | X1l | 1 1 1 0.. ] with X2 = xc^2,
// X1 = xc; B row = [ V2h | V2l | V2h | Mh | Ml | Mh | c1 c2 c3 ] with V2 = v^2, M = -2 v^2 mu_c (built on the host).
// vmax2[d] = the largest v_kd^2 of any density: qx = sum_d vmax2_d xc_d^2 bounds the feature's share of every density's
// distance and sets the frame's screening threshold.  xT receives the RAW feature (the reference subtracts it from the
// unscaled mean, src/Mm/GaussDiagonalMaximumFeatureScorer.cc:144-233).
__global__ void __launch_bounds__(256) gmm_split_features_diag_kernel(const float* __restrict__ feats,
                                                                      const float* __restrict__ centre,
                                                                      const float* __restrict__ vmax2, long T, int dim,
                                                                      int dp, int kPad, __half* __restrict__ A,
                                                                      float* __restrict__ thr, float thrA, float thrB,
                                                                      float* __restrict__ xT, long pitch, int pooled) {
    const int  sub = threadIdx.x & 7;
    const long g0  = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 3;
    const long nG  = ((long)gridDim.x * blockDim.x) >> 3;
    const int  nChunk = dp >> 3;
    for (long t0 = g0; t0 < ((T + 3) & ~3L); t0 += nG) {
        const long t   = t0 < T ? t0 : T - 1;
        float      qx  = 0.0f;
        uint32_t   h2[4] = {0, 0, 0, 0}, l2[4] = {0, 0, 0, 0}, h1[4] = {0, 0, 0, 0}, l1[4] = {0, 0, 0, 0};
        int        bad = 0;
        if (sub < nChunk) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d = sub * 8 + j;
                float     xr = 0.0f, x = 0.0f, w = 0.0f;
                if (d < dim) {
                    xr = __ldg(feats + (size_t)t * dim + d);
                    x  = __fsub_rn(xr, __ldg(centre + d));
                    w  = __ldg(vmax2 + d);
                }
                if (xT && t0 < T)
                    xT[(size_t)d * pitch + t] = xr;
                const float x2 = __fmul_rn(x, x);
                qx             = __fmaf_rn(x2, w, qx);
                bad |= !(x2 <= 57600.0f);
                const float  c2 = fminf(x2, 60000.0f), c1 = fminf(fmaxf(x, -60000.0f), 60000.0f);
                const __half a  = __float2half_rn(c2), b = __float2half_rn(c2 - __half2float(a));
                const __half c  = __float2half_rn(c1), e = __float2half_rn(c1 - __half2float(c));
                h2[j >> 1] |= (uint32_t)__half_as_ushort(a) << ((j & 1) * 16);
                l2[j >> 1] |= (uint32_t)__half_as_ushort(b) << ((j & 1) * 16);
                h1[j >> 1] |= (uint32_t)__half_as_ushort(c) << ((j & 1) * 16);
                l1[j >> 1] |= (uint32_t)__half_as_ushort(e) << ((j & 1) * 16);
            }
        }
        qx += __shfl_xor_sync(0xffffffffu, qx, 1);
        qx += __shfl_xor_sync(0xffffffffu, qx, 2);
        qx += __shfl_xor_sync(0xffffffffu, qx, 4);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 1);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 2);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 4);
        if (t0 < T) {
            uint4* row = reinterpret_cast<uint4*>(A + (size_t)t * kPad);
            // pooled covariance: sum_d v_d^2 xc_d^2 is the same for every density of the frame and cannot change which
            // densities are candidates, so the x^2 group of the expansion is left out (K = 3 dp + 3 instead of 6 dp + 3)
            const int first1 = pooled ? 0 : 3 * nChunk, ones = first1 + 3 * nChunk;
            if (sub < nChunk) {
                const uint4 a2 = make_uint4(h2[0], h2[1], h2[2], h2[3]), a1 = make_uint4(h1[0], h1[1], h1[2], h1[3]);
                if (!pooled) {
                    row[sub]              = a2;
                    row[nChunk + sub]     = a2;
                    row[2 * nChunk + sub] = make_uint4(l2[0], l2[1], l2[2], l2[3]);
                }
                row[first1 + sub]              = a1;
                row[first1 + nChunk + sub]     = a1;
                row[first1 + 2 * nChunk + sub] = make_uint4(l1[0], l1[1], l1[2], l1[3]);
            }
            for (int c = ones + sub; c < (kPad >> 3); c += 8)
                row[c] = c == ones ? make_uint4(0x3C003C00u, 0x00003C00u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
            if (sub == 0)
                thr[t] = (bad || !(qx < FLT_MAX)) ? __int_as_float(0x7f800000) : __fmaf_rn(qx, thrA, thrB);
        }
    }
}

// ---- B-stationary tcgen05 kernel ------------------------------------------------------------------
// The model slice of NBRES column blocks (256 densities x K each) is loaded into shared memory ONCE per
// CTA; only the 128-frame A tiles stream through a TMA ring.  Streaming both operands per tile (the
// generic gemm16_kernel) needs 96 KB of L2->SM traffic per 1024 tensor cycles and is L2-bound at
// ~8 TB/s; this kernel needs 16 KB per 1024 cycles.
// Warp roles as in gemm16_kernel: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue.
constexpr int kTensorThreads = 384;  // 4 control warps + 8 epilogue warps

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
}

template<class Epi, int KB, int NBRES, bool SPLIT>
__global__ void __launch_bounds__(kTensorThreads, 1)
        gmm_tensor_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M,
                          int nNB, uint32_t idesc, int aStages, const Epi epi) {
    using namespace rbgemm;
    constexpr int B_TILE = BN * BK * 2;  // 32 KB: one column block, one k block
    constexpr int A_TILE = BM * BK * 2;  // 16 KB
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw   = smem_u32(smem_dyn);
    const uint32_t pad   = (1024u - (raw & 1023u)) & 1023u;
    unsigned char* base  = smem_dyn + pad;
    const uint32_t sB    = raw + pad;
    const uint32_t sA    = sB + NBRES * KB * B_TILE;
    uint64_t* afull      = reinterpret_cast<uint64_t*>(base + NBRES * KB * B_TILE + aStages * A_TILE);
    uint64_t* aempty     = afull + 8;
    uint64_t* tfull      = aempty + 8;
    uint64_t* tempty     = tfull + 2;
    uint64_t* bfull      = tempty + 2;
    uint64_t* bempty     = bfull + 1;
    uint32_t* tmemPtr    = reinterpret_cast<uint32_t*>(bempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // column-block groups are dealt round-robin to the CTAs; the CTAs of a group stride over the frame
    // blocks.  Models with more groups than CTAs take several rounds (B is reloaded per round).
    const int  nGroups  = (nNB + NBRES - 1) / NBRES;
    const bool oneRound = nGroups <= (int)gridDim.x;
    const int  group0   = oneRound ? (int)blockIdx.x % nGroups : (int)blockIdx.x;
    const int  gstep    = oneRound ? nGroups : (int)gridDim.x;
    const int  member   = oneRound ? (int)blockIdx.x / nGroups : 0;
    const int  nMembers = oneRound ? ((int)gridDim.x - group0 + nGroups - 1) / nGroups : 1;
    const int  nMB      = (M + BM - 1) / BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < aStages; ++s) {
            mbar_init(&afull[s], 1);
            mbar_init(&aempty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], SPLIT ? 8 : 4);
        }
        mbar_init(bfull, 1);
        mbar_init(bempty, 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemPtr)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmemBase = *tmemPtr;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0, round = 0;
            for (int group = group0; group < nGroups; group += gstep, ++round) {
                const int nb0 = group * NBRES, nbCount = min(NBRES, nNB - nb0);
                mbar_wait(bempty, (round & 1u) ^ 1u);  // previous round's MMAs have finished reading B
                mbar_expect_tx(bfull, (uint32_t)(nbCount * KB * B_TILE));
                for (int nb = 0; nb < nbCount; ++nb)
                    for (int kb = 0; kb < KB; ++kb)
                        tma_load_2d(sB + (nb * KB + kb) * B_TILE, &tmB, kb * BK, (nb0 + nb) * BN, bfull);
                for (int mb = member; mb < nMB; mb += nMembers) {
                    for (int kb = 0; kb < KB; ++kb, ++it) {
                        const uint32_t s = it % aStages, ph = (it / aStages) & 1u;
                        mbar_wait(&aempty[s], ph ^ 1u);
                        mbar_expect_tx(&afull[s], A_TILE);
                        tma_load_2d(sA + s * A_TILE, &tmA, kb * BK, mb * BM, &afull[s]);
                    }
                }
            }
        }
    }
    else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, tc = 0, round = 0;
            for (int group = group0; group < nGroups; group += gstep, ++round) {
            const int nbCount = min(NBRES, nNB - group * NBRES);
            mbar_wait(bfull, round & 1u);
            tc_fence_after();
            for (int mb = member; mb < nMB; mb += nMembers, it += KB) {
                for (int nb = 0; nb < nbCount; ++nb, ++tc) {
                    const uint32_t a = tc & 1u, aph = (tc >> 1) & 1u;
                    mbar_wait(&tempty[a], aph ^ 1u);
                    tc_fence_after();
                    const uint32_t dTmem = tmemBase + a * BN;
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint32_t i2 = it + kb, s = i2 % aStages, ph = (i2 / aStages) & 1u;
                        if (nb == 0) {
                            mbar_wait(&afull[s], ph);
                            tc_fence_after();
                        }
                        const uint64_t ad = smem_desc(sA + s * A_TILE);
                        const uint64_t bd = smem_desc(sB + (nb * KB + kb) * B_TILE);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            tc_mma(dTmem, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                        if (nb == nbCount - 1)
                            tc_commit(&aempty[s]);  // the frame tile is no longer needed
                    }
                    tc_commit(&tfull[a]);
                }
            }
            tc_commit(bempty);
            }
        }
    }
    else if (warp >= 4 && (SPLIT || warp < 8)) {
        // epilogue: warp w may read TMEM lanes 32*(w%4)..+31; with SPLIT two warps share a lane quarter and
        // take 128 accumulator columns each (two warps per scheduler hide the TMEM-load latency)
        const int q    = warp & 3;
        const int half = (warp - 4) >> 2;
        uint32_t  tc   = 0;
        for (int group = group0; group < nGroups; group += gstep) {
            const int nb0 = group * NBRES, nbCount = min(NBRES, nNB - nb0);
            for (int mb = member; mb < nMB; mb += nMembers) {
                const int row = mb * BM + q * 32 + lane;
                for (int nb = 0; nb < nbCount; ++nb, ++tc) {
                    const uint32_t a = tc & 1u, aph = (tc >> 1) & 1u;
                    mbar_wait(&tfull[a], aph);
                    tc_fence_after();
                    typename Epi::State st;
                    if (row < M)
                        epi.begin(st, row);
                    const uint32_t tbase = tmemBase + ((uint32_t)(q * 32) << 16) + a * BN;
                    if constexpr (SPLIT) {
#pragma unroll 1
                        for (int c = 0; c < 2; ++c) {
                            const int col = half * 128 + c * 64;
                            float     v[64];
                            tmem_ld32_nowait(tbase + col, v);
                            tmem_ld32_nowait(tbase + col + 32, v + 32);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            if (row < M)
                                epi.emit64(st, row, (nb0 + nb) * BN + col, v);
                        }
                    }
                    else {
#pragma unroll 1
                        for (int c = 0; c < BN / 32; ++c) {
                            float v[32];
                            tmem_ld32(tbase + c * 32, v);
                            if (row < M)
                                epi.chunk(st, row, (nb0 + nb) * BN + c * 32, v);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(&tempty[a]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

}  // namespace

// ==========================================================================================
// host side
// ==========================================================================================

struct rb_gmm_tensor {
    rb::DeviceInfo dev;
    int            dim = 0, dp = 0, kPad = 0, nMix = 0, nCols = 0, seg = 0;
    float          scale = 1.0f;
    long           chunk = 262144;  // frames per GEMM launch at most (64 MB of fp16 A operand)
    long           cap   = 0;       // frames the per-call buffers below hold (grown on demand up to chunk)
    // screening (exact batch-float scoring): thr = thrA * |x|^2 + thrB in accumulator units, see rb_gmm_tensor_create
    bool           screenable = false;
    float          thrA = 0.0f, thrB = 0.0f;
    rb::DevBuf<__half>   dB, dA;
    rb::DevBuf<float>    dIsd, dCentre, dXnorm, dThr, dXT;
    rb::DevBuf<uint32_t> dWords;  // candidate sets of the screening pass, [nMix / 4][cap][4]
    bool                 diag = false;   // operands of the diagonal scorers (rb_gmm_tensor_create_diag)
    bool                 pooled = false; // ... with one covariance for all densities: no x^2 group (K = 3 dp + 3)
    rb::DevBuf<float>    dVmax2;
    rb::DevBuf<uint32_t> dEndMask;
    rb::DevBuf<int>      dMixStart;
    CUtensorMap          mapA, mapB;
};

namespace {

double log_norm_factor(const float* var, unsigned dim) {
    double s = 0;
    for (unsigned d = 0; d < dim; ++d)
        s += std::log(std::fabs((double)var[d]));
    return (double)dim * std::log(2.0 * M_PI) + s;
}

template<class Epi, int KB, int NBRES>
int launch_kernel(rb_gmm_tensor* t, long T, const Epi& epi, cudaStream_t s) {
    constexpr int B_TILE = rbgemm::BN * rbgemm::BK * 2, A_TILE = rbgemm::BM * rbgemm::BK * 2;
    const int     budget = 220 * 1024 - 1024 - 256 - NBRES * KB * B_TILE;
    const int     aStages = std::max(KB, std::min(8, budget / A_TILE));
    const int     smem = NBRES * KB * B_TILE + aStages * A_TILE + 256 + 1024;
    constexpr bool SPLIT = Epi::kUniform;
    RB_CUDA(cudaFuncSetAttribute(gmm_tensor_kernel<Epi, KB, NBRES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 smem));
    const int nNB  = t->nCols / rbgemm::BN;
    const int nMB  = (int)((T + rbgemm::BM - 1) / rbgemm::BM);
    const int nGroups = (nNB + NBRES - 1) / NBRES;
    const int grid = std::min(t->dev.sm_count, nGroups * nMB);
    gmm_tensor_kernel<Epi, KB, NBRES, SPLIT><<<grid, kTensorThreads, smem, s>>>(
            t->mapA, t->mapB, (int)T, nNB, rbgemm::instr_desc(rbgemm::FMT_F16), aStages, epi);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

template<class Epi>
int launch_epi(rb_gmm_tensor* t, long T, const Epi& epi, cudaStream_t s) {
    switch (t->kPad / rbgemm::BK) {
        case 1: return launch_kernel<Epi, 1, 2>(t, T, epi, s);
        case 2: return launch_kernel<Epi, 2, 2>(t, T, epi, s);
        case 3: return launch_kernel<Epi, 3, 1>(t, T, epi, s);
        case 4: return launch_kernel<Epi, 4, 1>(t, T, epi, s);
    }
    rb::set_error("unsupported K padding %d", t->kPad);
    return RB_ERR_UNSUPPORTED;
}

template<int SEG>
int launch_seg(rb_gmm_tensor* t, long T, float* dScores, cudaStream_t s) {
    EpiGmmMin<SEG> epi;
    epi.endMask  = t->dEndMask.p;
    epi.mixStart = t->dMixStart.p;
    epi.xnorm    = t->dXnorm.p;
    epi.scores   = dScores;
    epi.nMix     = t->nMix;
    epi.invS2    = 1.0f / (t->scale * t->scale);
    return launch_epi(t, T, epi, s);
}

template<int SEG>
int launch_screen(rb_gmm_tensor* t, long T, cudaStream_t s) {
    EpiGmmScreen<SEG> epi;
    epi.endMask  = t->dEndMask.p;
    epi.mixStart = t->dMixStart.p;
    epi.thr      = t->dThr.p;
    epi.masks    = t->dWords.p;
    epi.pitch    = t->cap;
    epi.nMix     = t->nMix;
    return launch_epi(t, T, epi, s);
}

// per-call buffers for up to min(T, chunk) frames; the TMA map of the A operand follows the allocation
int ensure_capacity(rb_gmm_tensor* t, long T, bool screen) {
    const long need = std::min(T, t->chunk);
    if (need > t->cap) {
        const long cap = std::max(need, t->cap);
        RB_CHECK(t->dA.reserve((size_t)cap * t->kPad));
        RB_CHECK(t->dXnorm.reserve((size_t)cap));
        RB_CHECK(rbgemm::make_map(&t->mapA, t->dA.p, (uint64_t)cap, (uint64_t)t->kPad, (uint64_t)t->kPad, rbgemm::BM, false));
        t->cap = cap;
    }
    if (screen) {
        RB_CHECK(t->dThr.reserve((size_t)t->cap));
        RB_CHECK(t->dXT.reserve((size_t)t->cap * t->dp));
        RB_CHECK(t->dWords.reserve((size_t)t->cap * t->nMix));
    }
    return RB_OK;
}

}  // namespace

int rb_gmm_tensor_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream,
                         rb_gmm_tensor** out) {
    *out = nullptr;
    if (ms->n_covariances != 1) {
        rb::set_error("tensor GMM scorer supports only one globally pooled covariance (got %u)", ms->n_covariances);
        return RB_ERR_UNSUPPORTED;
    }
    const unsigned D  = ms->dim;
    const int      dp = (int)rb::round_up(D, 8);
    const int      kPad = (int)rb::round_up((size_t)3 * dp + 3, 64);
    // mixture sizes: uniform power-of-two fast path or ragged
    const uint32_t n0 = ms->mix_offsets[1] - ms->mix_offsets[0];
    bool uniform = true;
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        const uint32_t n = ms->mix_offsets[m + 1] - ms->mix_offsets[m];
        if (n == 0) {
            rb::set_error("tensor GMM scorer does not support mixtures without densities (mixture %u)", m);
            return RB_ERR_UNSUPPORTED;
        }
        if (n > (uint32_t)rbgemm::BN) {
            rb::set_error("tensor GMM scorer supports at most %d densities per mixture (mixture %u has %u)",
                          rbgemm::BN, m, n);
            return RB_ERR_UNSUPPORTED;
        }
        uniform = uniform && n == n0;
    }
    rb_gmm_tensor* t = new (std::nothrow) rb_gmm_tensor();
    if (!t) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    auto fail = [&](int code) {
        delete t;
        return code;
    };
    t->dev  = dev;
    t->dim  = (int)D;
    t->dp   = dp;
    t->kPad = kPad;
    t->nMix = (int)ms->n_mixtures;
    t->seg  = (uniform && (n0 == 8 || n0 == 16 || n0 == 32)) ? (int)n0 : 0;

    // column layout: densities in mixture order; a mixture never straddles a 256-column tile
    std::vector<uint32_t> colEntry;  // mixture entry index per column, 0xffffffff = padding
    std::vector<uint32_t> colEnd;    // 1 if the column closes its mixture
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        const uint32_t e0 = ms->mix_offsets[m], n = ms->mix_offsets[m + 1] - e0;
        const size_t   used = colEntry.size() % rbgemm::BN;
        if (used + n > (size_t)rbgemm::BN)
            while (colEntry.size() % rbgemm::BN) {
                colEntry.push_back(0xffffffffu);
                colEnd.push_back(0);
            }
        for (uint32_t i = 0; i < n; ++i) {
            colEntry.push_back(e0 + i);
            colEnd.push_back(i + 1 == n);
        }
    }
    while (colEntry.size() % rbgemm::BN) {
        colEntry.push_back(0xffffffffu);
        colEnd.push_back(0);
    }
    t->nCols = (int)colEntry.size();

    // scaled means, centre, constants (f64 on the host)
    std::vector<float> isd(dp, 0.0f);
    for (unsigned d = 0; d < D; ++d)
        isd[d] = 1.0f / (float)std::sqrt((double)ms->variances[d]);
    const float         logNorm = (float)log_norm_factor(ms->variances, D);
    const uint32_t      nEntries = ms->mix_offsets[ms->n_mixtures];
    std::vector<float>  mu((size_t)nEntries * D);
    std::vector<double> centre64(D, 0.0);
    for (uint32_t e = 0; e < nEntries; ++e) {
        const uint32_t dns = ms->mix_density[e];
        const float*   src = ms->means + (size_t)ms->dens_mean[dns] * D;
        for (unsigned d = 0; d < D; ++d) {
            mu[(size_t)e * D + d] = src[d] * isd[d];  // f32 product, as the reference's init
            centre64[d] += mu[(size_t)e * D + d];
        }
    }
    std::vector<float> centre(dp, 0.0f);
    for (unsigned d = 0; d < D; ++d)
        centre[d] = (float)(centre64[d] / std::max<uint32_t>(nEntries, 1));
    // centred means (f32, the same subtraction the feature kernel performs) and c_k + |mu_k|^2
    double maxAbs = 0, maxC = 0;
    std::vector<double> cc(nEntries);
    for (uint32_t e = 0; e < nEntries; ++e) {
        double n2 = 0;
        for (unsigned d = 0; d < D; ++d) {
            float& v = mu[(size_t)e * D + d];
            v        = v - centre[d];
            n2 += (double)v * (double)v;
            maxAbs = std::max(maxAbs, std::fabs(2.0 * (double)v));
        }
        const float c = (float)((double)logNorm - 2 * ms->mix_log_weight[e]);  // the reference's f32 constant
        cc[e]         = (double)c + n2;
        maxC          = std::max(maxC, std::fabs(cc[e]));
    }
    // power-of-two operand scale keeping everything inside fp16 range
    float scale = 1.0f;
    while ((maxAbs * scale > 16384.0 || maxC * scale * scale > 32768.0) && scale > 1e-12f)
        scale *= 0.5f;
    t->scale = scale;

    std::vector<__half> B((size_t)t->nCols * kPad, __float2half_rn(0.0f));
    for (int col = 0; col < t->nCols; ++col) {
        const uint32_t e = colEntry[col];
        if (e == 0xffffffffu)
            continue;
        __half* row = B.data() + (size_t)col * kPad;
        for (unsigned d = 0; d < D; ++d) {
            const float  v = -2.0f * mu[(size_t)e * D + d] * scale;
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            row[d]          = h;
            row[dp + d]     = l;
            row[2 * dp + d] = h;
        }
        double r = cc[e] * (double)scale * (double)scale;
        for (int j = 0; j < 3; ++j) {
            const __half h   = __float2half_rn((float)r);
            row[3 * dp + j]  = h;
            r -= (double)__half2float(h);
        }
    }
    // per-chunk segment metadata
    const int             nChunks = t->nCols / 32;
    std::vector<uint32_t> endMask(nChunks, 0);
    std::vector<int>      mixStart(nChunks, 0);
    int                   mix = 0;
    for (int c = 0; c < nChunks; ++c) {
        mixStart[c] = mix;
        for (int j = 0; j < 32; ++j)
            if (colEnd[c * 32 + j]) {
                endMask[c] |= 1u << j;
                ++mix;
            }
    }
    if (t->dB.upload(B.data(), B.size(), stream) != RB_OK || t->dIsd.upload(isd, stream) != RB_OK ||
        t->dCentre.upload(centre, stream) != RB_OK || t->dEndMask.upload(endMask, stream) != RB_OK ||
        t->dMixStart.upload(mixStart, stream) != RB_OK)
        return fail(RB_ERR_CUDA);
    {
        // Screening threshold.  With Q = |xc|^2 + |mu_c|^2 + |c + |mu_c|^2| (centred, scaled values)
        //   - the reference's own f32 sum (<= 5 fused accumulations + 3 adds per lane, differences rounded once) is
        //     within 11 u (|c| + |x - mu|^2) <= 33 u Q of the exact value, u = 2^-24;
        //   - the split-precision product misses the exact cross term by the centring roundings (4 u Q), the fp16
        //     representation of both operands and the dropped lo * lo products (3 * 2^-22 Q, plus 2^-25 per element
        //     where a low part falls into the fp16 subnormal range), the residual of the three-term constant
        //     (measured below) and the tensor core's f32 accumulation over K / 16 steps (taken as <= 2^-20 Q).
        // A density can win in the reference's arithmetic only if its approximate score is within twice the sum of the
        // two of the approximate minimum: <= 2^-17 Q.  kappa = 2^-16 doubles that again; tests/test_gpu_gmm_exact.py
        // measures the actual error of the product (~2^-22 Q) on the device.
        double qmu = 0, resid = 0;
        bool   finite = std::isfinite((double)logNorm);
        uint32_t nMax = 0;
        for (uint32_t m = 0; m < ms->n_mixtures; ++m)
            nMax = std::max(nMax, ms->mix_offsets[m + 1] - ms->mix_offsets[m]);
        for (uint32_t e = 0; e < nEntries; ++e) {
            double n2 = 0;
            for (unsigned d = 0; d < D; ++d) {
                const double v = mu[(size_t)e * D + d];
                n2 += v * v;
                finite = finite && std::isfinite(v);
            }
            finite = finite && std::isfinite(cc[e]);
            qmu    = std::max(qmu, n2 + std::fabs(cc[e]));
        }
        for (int col = 0; col < t->nCols; ++col) {
            const uint32_t e = colEntry[col];
            if (e == 0xffffffffu)
                continue;
            const __half* row = B.data() + (size_t)col * kPad;
            const double  got = ((double)__half2float(row[3 * dp]) + (double)__half2float(row[3 * dp + 1]) +
                                (double)__half2float(row[3 * dp + 2])) / ((double)scale * (double)scale);
            resid = std::max(resid, std::fabs(got - cc[e]));
        }
        const double kappa = std::ldexp(1.0, -16) + std::ldexp(1.0, -21) * std::sqrt((double)dp) / (double)scale;
        const double s2    = (double)scale * (double)scale;
        t->thrA       = (float)(s2 * kappa);
        t->thrB       = (float)(s2 * (kappa * (qmu + 2.0) + 2.0 * resid) * 1.0001);
        t->screenable = finite && nMax <= 32 && std::isfinite(t->thrA) && std::isfinite(t->thrB) && t->thrA > 0;
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess) {
        rb::set_error("tensor GMM model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(RB_ERR_CUDA);
    }
    int rc = rbgemm::make_map(&t->mapB, t->dB.p, (uint64_t)t->nCols, (uint64_t)kPad, (uint64_t)kPad, rbgemm::BN, false);
    if (rc != RB_OK)
        return fail(rc);
    *out = t;
    return RB_OK;
}

// Operands of the screening pass for the DIAGONAL scorers (per-density covariance, Mm::GaussDiagonalMaximumFeatureScorer).
// `rows` are the rows gmm.cu built for its direct kernel, one per mixture entry in mixture order:
// [ mu (4 nq) | isd (4 nq) | w | logNorm | flags | 0 ], score = w + logNorm + sum_d ((mu_d - x_d) isd_d)^2.
int rb_gmm_tensor_create_diag(const rb_mixture_set* ms, const float* rows, int rowf, int nq, const rb::DeviceInfo& dev,
                              cudaStream_t stream, rb_gmm_tensor** out) {
    *out = nullptr;
    const unsigned D  = ms->dim;
    const int      dp = (int)rb::round_up(D, 8);
    // one covariance for all densities: the x^2 group of the expansion is a per-frame constant and is dropped
    bool pooled = ms->n_covariances == 1 && getenv("RB_GMM_DIAG_FULL_SCREEN") == nullptr;
    for (uint32_t i = 0; i < ms->n_densities && pooled; ++i)
        pooled = ms->dens_cov[i] == 0;
    const int groups = pooled ? 3 : 6;
    const int kPad   = (int)rb::round_up((size_t)groups * dp + 3, 64);
    if (kPad > 4 * rbgemm::BK) {
        rb::set_error("diagonal screening supports at most 40 dimensions (got %u)", D);
        return RB_ERR_UNSUPPORTED;
    }
    const uint32_t n0 = ms->mix_offsets[1] - ms->mix_offsets[0];
    bool           uniform = true;
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        const uint32_t n = ms->mix_offsets[m + 1] - ms->mix_offsets[m];
        if (n == 0 || n > 32) {
            rb::set_error("diagonal screening needs 1..32 densities per mixture (mixture %u has %u)", m, n);
            return RB_ERR_UNSUPPORTED;
        }
        uniform = uniform && n == n0;
    }
    rb_gmm_tensor* t = new (std::nothrow) rb_gmm_tensor();
    if (!t) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    auto fail = [&](int code) {
        delete t;
        return code;
    };
    t->dev  = dev;
    t->dim  = (int)D;
    t->dp   = dp;
    t->kPad = kPad;
    t->nMix = (int)ms->n_mixtures;
    t->seg  = (uniform && (n0 == 8 || n0 == 16 || n0 == 32)) ? (int)n0 : 0;
    t->diag   = true;
    t->pooled = pooled;
    std::vector<uint32_t> colEntry, colEnd;
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        const uint32_t e0 = ms->mix_offsets[m], n = ms->mix_offsets[m + 1] - e0;
        if (colEntry.size() % rbgemm::BN + n > (size_t)rbgemm::BN)
            while (colEntry.size() % rbgemm::BN) {
                colEntry.push_back(0xffffffffu);
                colEnd.push_back(0);
            }
        for (uint32_t i = 0; i < n; ++i) {
            colEntry.push_back(e0 + i);
            colEnd.push_back(i + 1 == n);
        }
    }
    while (colEntry.size() % rbgemm::BN) {
        colEntry.push_back(0xffffffffu);
        colEnd.push_back(0);
    }
    t->nCols = (int)colEntry.size();
    const uint32_t nEntries = ms->mix_offsets[ms->n_mixtures];
    // centre = mean of all means (f32, the value the feature kernel subtracts)
    std::vector<double> centre64(D, 0.0);
    for (uint32_t e = 0; e < nEntries; ++e)
        for (unsigned d = 0; d < D; ++d)
            centre64[d] += rows[(size_t)e * rowf + d];
    std::vector<float> centre(dp, 0.0f), vmax2(dp, 0.0f);
    for (unsigned d = 0; d < D; ++d)
        centre[d] = (float)(centre64[d] / std::max<uint32_t>(nEntries, 1));
    // per entry: V2 = v^2, M = -2 v^2 mu_c, constant = w + logNorm + sum v^2 mu_c^2 (f64)
    std::vector<double> v2((size_t)nEntries * D), mm((size_t)nEntries * D), cc(nEntries);
    double              maxAbs = 0, maxC = 0, qmu = 0;
    bool                finite = true;
    for (uint32_t e = 0; e < nEntries; ++e) {
        const float* row = rows + (size_t)e * rowf;
        double       acc = 0;
        for (unsigned d = 0; d < D; ++d) {
            const double v   = (double)row[4 * nq + d];
            const float  muc = row[d] - centre[d];  // f32, like the feature's centring
            const double a = v * v, b = a * (double)muc;
            v2[(size_t)e * D + d] = a;
            mm[(size_t)e * D + d] = -2.0 * b;
            acc += b * (double)muc;
            maxAbs   = std::max(maxAbs, std::max(a, std::fabs(2.0 * b)));
            vmax2[d] = std::max(vmax2[d], (float)(a * (1.0 + 1e-6)));
            finite   = finite && std::isfinite(a) && std::isfinite(b);
        }
        cc[e]  = (double)row[8 * nq] + (double)row[8 * nq + 1] + acc;
        finite = finite && std::isfinite(cc[e]);
        maxC   = std::max(maxC, std::fabs(cc[e]));
        qmu    = std::max(qmu, acc + std::fabs(cc[e]));
    }
    float scale = 1.0f;
    while ((maxAbs * scale > 16384.0 || maxC * scale > 32768.0) && scale > 1e-12f)
        scale *= 0.5f;
    while (maxAbs * scale < 64.0 && maxC * scale < 128.0 && scale < 1e12f && maxAbs > 0)  // keep the low halves normal
        scale *= 2.0f;
    t->scale = scale;
    std::vector<__half> B((size_t)t->nCols * kPad, __float2half_rn(0.0f));
    double              resid = 0;
    for (int col = 0; col < t->nCols; ++col) {
        const uint32_t e = colEntry[col];
        if (e == 0xffffffffu)
            continue;
        __half* row = B.data() + (size_t)col * kPad;
        for (unsigned d = 0; d < D; ++d) {
            const float  a = (float)(v2[(size_t)e * D + d] * scale), b = (float)(mm[(size_t)e * D + d] * scale);
            const __half ah = __float2half_rn(a), al = __float2half_rn(a - __half2float(ah));
            const __half bh = __float2half_rn(b), bl = __float2half_rn(b - __half2float(bh));
            const int    first1 = pooled ? 0 : 3 * dp;
            if (!pooled) {
                row[d]          = ah;
                row[dp + d]     = al;
                row[2 * dp + d] = ah;
            }
            row[first1 + d]          = bh;
            row[first1 + dp + d]     = bl;
            row[first1 + 2 * dp + d] = bh;
        }
        double r = cc[e] * (double)scale, got = 0;
        for (int j = 0; j < 3; ++j) {
            const __half h  = __float2half_rn((float)r);
            row[groups * dp + j] = h;
            r -= (double)__half2float(h);
            got += (double)__half2float(h);
        }
        resid = std::max(resid, std::fabs(got / (double)scale - cc[e]));
    }
    const int             nChunks = t->nCols / 32;
    std::vector<uint32_t> endMask(nChunks, 0);
    std::vector<int>      mixStart(nChunks, 0);
    int                   mix = 0;
    for (int c = 0; c < nChunks; ++c) {
        mixStart[c] = mix;
        for (int j = 0; j < 32; ++j)
            if (colEnd[c * 32 + j]) {
                endMask[c] |= 1u << j;
                ++mix;
            }
    }
    // Threshold: the bound of the batch scorer with the two product groups of the expansion, sixteen
    // K steps and the reference's longer rounding chain (difference, scaling by 1/sigma, up to 10 accumulations per
    // lane, the f64 sum narrowed to f32): twice the sum of all terms stays below 2^-16 Q, kappa = 2^-15.
    const double kappa = std::ldexp(1.0, -15) + std::ldexp(1.0, -20) * std::sqrt((double)dp) / (double)scale;
    t->thrA       = (float)((double)scale * kappa);
    t->thrB       = (float)((double)scale * (kappa * (qmu + 2.0) + 2.0 * resid) * 1.0001);
    t->screenable = finite && std::isfinite(t->thrA) && std::isfinite(t->thrB) && t->thrA > 0;
    if (t->dB.upload(B.data(), B.size(), stream) != RB_OK || t->dCentre.upload(centre, stream) != RB_OK ||
        t->dVmax2.upload(vmax2, stream) != RB_OK || t->dEndMask.upload(endMask, stream) != RB_OK ||
        t->dMixStart.upload(mixStart, stream) != RB_OK)
        return fail(RB_ERR_CUDA);
    if (cudaStreamSynchronize(stream) != cudaSuccess) {
        rb::set_error("tensor GMM model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(RB_ERR_CUDA);
    }
    const int rc = rbgemm::make_map(&t->mapB, t->dB.p, (uint64_t)t->nCols, (uint64_t)kPad, (uint64_t)kPad, rbgemm::BN, false);
    if (rc != RB_OK)
        return fail(rc);
    *out = t;
    return RB_OK;
}

void rb_gmm_tensor_destroy(rb_gmm_tensor* t) {
    delete t;
}

// size the per-call buffers once for calls of up to `frames` frames (a host-buffer call that scores slab after slab must
// not reallocate -- cudaFree synchronises the device -- in the middle of its copy / compute pipeline)
int rb_gmm_tensor_reserve(rb_gmm_tensor* t, long frames, bool screen) {
    return ensure_capacity(t, frames, screen);
}

bool rb_gmm_tensor_screenable(const rb_gmm_tensor* t) {
    return t && t->screenable;
}
long rb_gmm_tensor_chunk(const rb_gmm_tensor* t) {
    return t->chunk;
}

// The operand split alone: the reference's scaled features, transposed (*xT, [dp x pitch]) and the (uninitialised) word
// buffer [nMix / 4][pitch][4] for a caller that decides the candidate sets itself (density preselection, gmm.cu).
int rb_gmm_tensor_split(rb_gmm_tensor* t, const float* dFeats, long n, uint32_t** words, const float** xT, long* pitch,
                        cudaStream_t s) {
    RB_REQUIRE(t->screenable && !t->diag, "this mixture set has no batch operand split");
    RB_REQUIRE(n >= 1 && n <= t->chunk, "bad frame count for one pass");
    RB_CHECK(ensure_capacity(t, n, true));
    const int blocks = (int)std::min<long>((n + 31) / 32, (long)t->dev.sm_count * 16);
    gmm_split_features_kernel<<<blocks, 256, 0, s>>>(dFeats, t->dIsd.p, t->dCentre.p, n, t->dim, t->dp, t->kPad, t->scale,
                                                     t->dA.p, t->dXnorm.p, t->dThr.p, t->thrA, t->thrB, t->dXT.p, t->cap);
    RB_LAUNCH_CHECK();
    *xT    = t->dXT.p;
    *words = t->dWords.p;
    *pitch = t->cap;
    return RB_OK;
}

// Exact batch-float scoring, first half: for n <= chunk frames compute the candidate-density words (*words, laid out
// [nMix / 4][pitch][4]) and the reference's scaled features, transposed ([dp x pitch], x' = fl(feat * isd)) (*xT).
int rb_gmm_tensor_screen(rb_gmm_tensor* t, const float* dFeats, long n, const uint32_t** words, const float** xT,
                         long* pitch, cudaStream_t s, cudaEvent_t afterSplit) {
    RB_REQUIRE(t->screenable, "this mixture set cannot be screened");
    RB_REQUIRE(n >= 1 && n <= t->chunk, "bad frame count for one screening pass");
    RB_CHECK(ensure_capacity(t, n, true));
    const int blocks = (int)std::min<long>((n + 31) / 32, (long)t->dev.sm_count * 16);
    if (t->diag)
        gmm_split_features_diag_kernel<<<blocks, 256, 0, s>>>(dFeats, t->dCentre.p, t->dVmax2.p, n, t->dim, t->dp, t->kPad,
                                                              t->dA.p, t->dThr.p, t->thrA, t->thrB, t->dXT.p, t->cap,
                                                              t->pooled ? 1 : 0);
    else
        gmm_split_features_kernel<<<blocks, 256, 0, s>>>(dFeats, t->dIsd.p, t->dCentre.p, n, t->dim, t->dp, t->kPad, t->scale,
                                                         t->dA.p, t->dXnorm.p, t->dThr.p, t->thrA, t->thrB, t->dXT.p,
                                                         t->cap);
    RB_LAUNCH_CHECK();
    if (afterSplit)
        cudaEventRecord(afterSplit, s);
    *xT    = t->dXT.p;
    *words = t->dWords.p;
    *pitch = t->cap;
    switch (t->seg) {
        case 8: return launch_screen<8>(t, n, s);
        case 16: return launch_screen<16>(t, n, s);
        case 32: return launch_screen<32>(t, n, s);
        default: return launch_screen<0>(t, n, s);
    }
}

int rb_gmm_tensor_score(rb_gmm_tensor* t, const float* dFeats, long T, float* dScores, cudaStream_t s) {
    RB_CHECK(ensure_capacity(t, T, false));
    for (long a = 0; a < T; a += t->chunk) {
        const long n      = std::min(t->chunk, T - a);
        const int  blocks = (int)std::min<long>((n + 31) / 32, (long)t->dev.sm_count * 16);
        gmm_split_features_kernel<<<blocks, 256, 0, s>>>(dFeats + (size_t)a * t->dim, t->dIsd.p, t->dCentre.p, n,
                                                         t->dim, t->dp, t->kPad, t->scale, t->dA.p, t->dXnorm.p,
                                                         nullptr, 0.0f, 0.0f, nullptr, 0);
        RB_LAUNCH_CHECK();
        float* out = dScores + (size_t)a * t->nMix;
        int    rc;
        switch (t->seg) {
            case 8: rc = launch_seg<8>(t, n, out, s); break;
            case 16: rc = launch_seg<16>(t, n, out, s); break;
            case 32: rc = launch_seg<32>(t, n, out, s); break;
            default: rc = launch_seg<0>(t, n, out, s); break;
        }
        RB_CHECK(rc);
    }
    return RB_OK;
}
