// gmm.cu -- diagonal-covariance GMM emission scorer on CUDA cores (sm_100a).
//
// Replaces, for ALL mixtures and ALL frames at once (dense T x nMix score matrix):
//   Mm::BatchFloatFeatureScorer::fillScoreCacheTpl          src/Mm/BatchFeatureScorer.cc:207-253
//   Mm::GaussDiagonalMaximumFeatureScorer::calculateScoreAndDensity  src/Mm/GaussDiagonalMaximumFeatureScorer.cc:116-142
//   Mm::GaussDiagonalSumFeatureScorer::calculateScoreAndDensity      same file :263-290
//
// Design (B200-first, not a port of the SSE loops):
//   * work item = (block of 512 frames) x (group of mixtures); a persistent grid of
//     sm_count x occupancy CTAs walks the items, so the grid is always a multiple of the SM count.
//   * a thread owns TWO frames whose (scaled) feature vectors live in registers for the whole item;
//     a warp walks the densities of the group in mixture order, so the density row is a
//     warp-uniform shared-memory broadcast (one LDS.128 feeds 8 sub+fma pairs) and the min /
//     log-sum-exp over the densities of a mixture is a private register recurrence: no cross-thread
//     reduction, ragged mixtures cost nothing.
//   * density rows (scaled mean | constant | flags) are streamed global->shared by the TMA unit
//     (cp.async.bulk, SASS UBLKCP) through a 3-stage mbarrier ring; the model (0.7 MB for C2) stays
//     L2-resident, HBM traffic is the algorithmic 156 B in + 1024 B out per frame.
//   * arithmetic follows the reference's SSE lane structure exactly (lane j of the two 4-wide
//     accumulators sums dims 8b+j / 8b+4+j, constant added first in lane 0, same horizontal
//     add tree), so BATCH_FLOAT scores are bit-identical to the CPU path, contraction on or off.
//   * the kernel is FP32-ALU bound (2 issue slots per (frame,density,dim)); see DESIGN.md.
//
// BATCH_FLOAT on models the screening covers does not run that kernel (at any batch size: the route below is faster
// from one frame on): the reference's result for a mixture is a minimum,
// so gmm_tensor.cu screens the densities that can win with a split-precision tcgen05 product and gmm_refine_kernel
// (below) evaluates only those, in the reference's operation order -- same bits, a sixteenth of the FP32 work
// (rb_gmm_score_fanout_dev stores the result into several GPUs' windows at once).
#include <cfloat>
#include <cmath>

#include "common.cuh"
#include "f32x2.cuh"
#include "internal.h"

namespace {

using namespace rbdev;
using namespace rbf32x2;

constexpr int kThreads        = 256;
constexpr int kFramesPerThr   = 2;
constexpr int kFramesPerBlock = kThreads * kFramesPerThr;  // 512
constexpr int kStages         = 3;
constexpr int kChunkRows      = 64;  // density rows per pipeline stage (64: +1 % over 32, 16: -2 %)

struct GmmParams {
    const float*    rows;      // [nRows * rowf]
    const int*      grp_row;   // [G+1] row boundaries of the mixture groups
    const int*      grp_mix;   // [G+1] mixture boundaries
    const float*    isd;       // [dpad] inverse std-dev of the pooled covariance (BATCH only)
    const float*    feats;     // [T * dim]
    float*          scores;    // [T * nMix]
    uint32_t*       best;      // [T * nMix] or null
    long            T;
    int             dim;
    int             nMix;
    int             nGroups;
    int             nFrameBlocks;
    int             vec4;  // 1: nMix % 4 == 0 and scores 16-byte aligned -> float4 stores
};

__device__ __forceinline__ float sq_acc(float d, float a, bool fuse) {
    return fuse ? __fmaf_rn(d, d, a) : __fadd_rn(a, __fmul_rn(d, d));
}

// ------------------------------------------------------------------------------------------
// BATCH_FLOAT: row = [ mu' (NB*8) | c | 0 | flags | 0 ],  rowf = NB*8+4
// ------------------------------------------------------------------------------------------
template<int NB, bool FUSE, int FPT>
__global__ void __launch_bounds__(kThreads, (NB <= 5) ? (FPT == 1 ? 3 : (FPT == 2 ? 2 : 1)) : 1) gmm_batch_kernel(const GmmParams p) {
    constexpr int ROWF  = NB * 8 + 4;
    constexpr int CHUNK = kChunkRows * ROWF;  // floats per stage
    constexpr int FPB   = kThreads * FPT;     // frames per block
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float*    buf = reinterpret_cast<float*>(smem_raw);                       // kStages * CHUNK
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + sizeof(float) * kStages * CHUNK);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s)
            mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    uint32_t  seq    = 0;  // chunks consumed so far by this CTA (stage = seq % kStages)
    const int nItems = p.nGroups * p.nFrameBlocks;

    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int  g    = item / p.nFrameBlocks;
        const int  fb   = item - g * p.nFrameBlocks;
        const int  row0 = p.grp_row[g], row1 = p.grp_row[g + 1];
        int        mix  = p.grp_mix[g];
        const int  nCh  = (row1 - row0 + kChunkRows - 1) / kChunkRows;
        long       t[FPT];
#pragma unroll
        for (int f = 0; f < FPT; ++f)
            t[f] = (long)fb * FPB + f * kThreads + tid;

        // producer prologue
        if (tid == 0) {
            for (int c = 0; c < kStages - 1 && c < nCh; ++c) {
                const int      r0    = row0 + c * kChunkRows;
                const int      nr    = min(kChunkRows, row1 - r0);
                const uint32_t bytes = (uint32_t)nr * ROWF * 4u;
                const uint32_t st    = (seq + c) % kStages;
                mbar_expect_tx(&bar[st], bytes);
                bulk_g2s(buf + st * CHUNK, p.rows + (size_t)r0 * ROWF, bytes, &bar[st]);
            }
        }

        // the feature vectors of this thread, scaled by 1/sigma (setFeature :157-162); dims (2i, 2i+1) packed
        uint64_t x[FPT][NB * 4];
#pragma unroll
        for (int f = 0; f < FPT; ++f) {
            const long   tc = t[f] < p.T ? t[f] : p.T - 1;
            const float* fp = p.feats + (size_t)tc * p.dim;
#pragma unroll
            for (int i = 0; i < NB * 4; ++i) {
                float v[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int d = 2 * i + k;
                    v[k]        = d < p.dim ? __fmul_rn(__ldg(fp + d), __ldg(p.isd + d)) : 0.0f;
                }
                x[f][i] = pack2(v[0], v[1]);
            }
        }

        float best[FPT], o[FPT][4];  // running minimum / staged outputs
#pragma unroll
        for (int f = 0; f < FPT; ++f) {
            best[f] = FLT_MAX;
            o[f][0] = o[f][1] = o[f][2] = o[f][3] = 0.0f;
        }
        int nStaged = 0;

        for (int c = 0; c < nCh; ++c) {
            if (tid == 0 && c + kStages - 1 < nCh) {
                const int      cc    = c + kStages - 1;
                const int      r0    = row0 + cc * kChunkRows;
                const int      nr    = min(kChunkRows, row1 - r0);
                const uint32_t bytes = (uint32_t)nr * ROWF * 4u;
                const uint32_t st    = (seq + cc) % kStages;
                mbar_expect_tx(&bar[st], bytes);
                bulk_g2s(buf + st * CHUNK, p.rows + (size_t)r0 * ROWF, bytes, &bar[st]);
            }
            const uint32_t n  = seq + c;
            const uint32_t st = n % kStages;
            mbar_wait(&bar[st], (n / kStages) & 1u);
            const float* chunk = buf + st * CHUNK;
            const int    nr    = min(kChunkRows, row1 - (row0 + c * kChunkRows));

            for (int r = 0; r < nr; ++r) {
                // packed f32x2 arithmetic (FADD2/FFMA2): lanes (2i, 2i+1) of the reference's 8 partial sums share a
                // 64-bit register pair; every lane is still one IEEE f32 sub / fma, so the result is unchanged
                const ulonglong2* row  = reinterpret_cast<const ulonglong2*>(chunk + r * ROWF);
                const ulonglong2  tail = row[NB * 2];  // (c, 0 | flags, 0)
                uint64_t          a[FPT][4];
#pragma unroll
                for (int f = 0; f < FPT; ++f) {
                    a[f][0] = tail.x;  // constant first, in lane 0 of the first accumulator (:217-219)
                    a[f][1] = a[f][2] = a[f][3] = 0ull;
                }
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const ulonglong2 m0 = row[2 * b], m1 = row[2 * b + 1];
                    const uint64_t   m[4] = {m0.x, m0.y, m1.x, m1.y};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#pragma unroll
                        for (int f = 0; f < FPT; ++f) {
                            const uint64_t d = sub2(m[j], x[f][4 * b + j]);
                            a[f][j]          = FUSE ? fma2(d, d, a[f][j]) : sqadd2(d, a[f][j]);
                        }
                    }
                }
                const bool last = (uint32_t)tail.y & 1u;  // last density of its mixture: emit
                float      v[FPT];
#pragma unroll
                for (int f = 0; f < FPT; ++f) {
                    // s1 += s2; then the two shuffle/add steps of :236-243: (l0+l2)+(l1+l3)
                    const uint64_t q  = add2(add2(a[f][0], a[f][2]), add2(a[f][1], a[f][3]));
                    const float    rs = __fadd_rn(lo2(q), hi2(q));
                    best[f]           = best[f] < rs ? best[f] : rs;
                }
                if (last) {
#pragma unroll
                    for (int f = 0; f < FPT; ++f) {
                        v[f]    = best[f] < FLT_MAX ? __fmul_rn(best[f], 0.5f) : best[f];
                        best[f] = FLT_MAX;
                    }
                    if (p.vec4) {
#pragma unroll
                        for (int f = 0; f < FPT; ++f) {
                            o[f][0] = o[f][1];
                            o[f][1] = o[f][2];
                            o[f][2] = o[f][3];
                            o[f][3] = v[f];
                        }
                        if (++nStaged == 4) {
                            nStaged = 0;
#pragma unroll
                            for (int f = 0; f < FPT; ++f)
                                if (t[f] < p.T)
                                    *reinterpret_cast<float4*>(p.scores + (size_t)t[f] * p.nMix + (mix - 3)) =
                                            make_float4(o[f][0], o[f][1], o[f][2], o[f][3]);
                        }
                    }
                    else {
#pragma unroll
                        for (int f = 0; f < FPT; ++f)
                            if (t[f] < p.T)
                                p.scores[(size_t)t[f] * p.nMix + mix] = v[f];
                    }
                    ++mix;
                }
            }
            __syncthreads();  // everyone is done with stage st before it is refilled
        }
        seq += (uint32_t)nCh;
    }
}

// ------------------------------------------------------------------------------------------
// BATCH_FLOAT, second half of the exact two-pass scorer.  gmm_tensor.cu's screening pass left, for
// every (frame, mixture), the set of densities that can still win -- one bit per density, almost always a single bit --
// in the score matrix itself.  Here only those densities are evaluated, in the reference's operation order (the same
// lanes, fused multiply-adds and add tree as gmm_batch_kernel above), and the word is overwritten with the score.
//   * a CTA keeps the rows of one mixture group in shared memory for its whole life (one bulk copy) and strides over
//     blocks of 256 frames; a thread owns one frame, its scaled feature vector (read coalesced from the transposed
//     copy the screening pass wrote) stays in registers as packed f32x2 pairs;
//   * per mixture a lane picks ITS candidate row (per-lane LDS.128 addresses), candidates in ascending density order
//     with the reference's `min_ps` operand order, so a frame whose set is "every density" (non-finite input) replays
//     the reference loop exactly.
// ------------------------------------------------------------------------------------------
struct RefineParams {
    const float* rows;      // [nRows * refine_pitch(NB)], mixture order
    const int*   grp_row;   // [G+1]
    const int*   grp_mix;   // [G+1]
    const int*   mix_row;   // [nMix+1] first row of every mixture
    const float* xT;        // [NB*8 x pitch] scaled features, transposed
    long         pitch;
    const uint32_t* words;  // [nMix / 4][pitch][4] candidate sets (EpiGmmScreen, gmm_tensor.cu)
    float*       scores;    // [T * nMix]
    float*       extra[15]; // further destinations that receive the same rows (peer windows, rb_gmm_score_fanout_dev)
    int          nExtra;
    uint32_t*    best;      // DIAG_MAX: [T * nMix] index of the winning density within its mixture, or null
    int          dim;       // DIAG_MAX: feature dimension (tail dims are summed sequentially)
    long         T;
    int          nMix, nGroups, nFrameBlocks;
    float        emptyScore;  // score of a mixture without a candidate: FLT_MAX, or the preselection scorer's back-off score
    const float* pooledIsd;   // DIAG_MAX, one covariance for all densities: its NQ*4 scaled 1/sqrt(var) (rows hold no copy)
};

constexpr int kStageQuads = 4;                    // mixtures staged per flush = 16
constexpr int kStagePitch = kStageQuads * 4 + 4;  // floats per frame row of the staging tile (80 B: conflict-free STS.128)
constexpr int kStageBytes = (kThreads / 32) * 32 * kStagePitch * 4;

template<int NB, bool FUSE>
__global__ void __launch_bounds__(kThreads, 3) gmm_refine_kernel(const RefineParams p) {
    constexpr int ROWF = refine_pitch(NB);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int g    = blockIdx.x % p.nGroups;
    const int row0 = p.grp_row[g], row1 = p.grp_row[g + 1];
    const int mix0 = p.grp_mix[g], mix1 = p.grp_mix[g + 1];
    // shared memory: [ mbarrier | first row offset of every mixture of the group | the group's rows ]
    uint64_t*      bar    = reinterpret_cast<uint64_t*>(smem_raw);
    int*           first  = reinterpret_cast<int*>(smem_raw + 16);
    unsigned char* region = smem_raw + 16 + (((size_t)(mix1 - mix0) * sizeof(int) + 15) & ~(size_t)15);
    // a row is 8-byte aligned only: the bulk copy starts at the 16-byte boundary below the first row and ends at the
    // one above the last (the device buffer carries a spare row for the overhang)
    const size_t   gBegin = (size_t)row0 * ROWF * 4, gEnd = (size_t)row1 * ROWF * 4;
    const size_t   cBegin = gBegin & ~(size_t)15, cEnd = (gEnd + 15) & ~(size_t)15;
    const float*   rows   = reinterpret_cast<const float*>(region + (gBegin - cBegin));
    float*         stage  = reinterpret_cast<float*>(region + (((gEnd - cBegin) + 47) & ~(size_t)15));  // after the rows
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    for (int m = tid; m < mix1 - mix0; m += kThreads)
        first[m] = (p.mix_row[mix0 + m] - row0) * ROWF;
    __syncthreads();
    if (tid == 0) {
        const size_t total = cEnd - cBegin;
        mbar_expect_tx(bar, (uint32_t)total);
        for (size_t off = 0; off < total; off += 65536) {
            const uint32_t bytes = (uint32_t)(total - off < 65536 ? total - off : 65536);
            bulk_g2s(region + off, reinterpret_cast<const unsigned char*>(p.rows) + cBegin + off, bytes, bar);
        }
    }
    mbar_wait(bar, 0);

    const int member   = blockIdx.x / p.nGroups;
    const int nMembers = ((int)gridDim.x - g + p.nGroups - 1) / p.nGroups;
    for (int fb = member; fb < p.nFrameBlocks; fb += nMembers) {
        const long t  = (long)fb * kThreads + tid;
        const long tc = t < p.T ? t : p.T - 1;
        uint64_t   x[NB * 4];
#pragma unroll
        for (int i = 0; i < NB * 4; ++i)
            x[i] = pack2(__ldg(p.xT + (size_t)(2 * i) * p.pitch + tc), __ldg(p.xT + (size_t)(2 * i + 1) * p.pitch + tc));
        const uint4* w  = reinterpret_cast<const uint4*>(p.words) + tc;  // + (m4 / 4) * pitch
        uint4        mk = t < p.T ? __ldg(w + (size_t)(mix0 >> 2) * p.pitch) : make_uint4(0u, 0u, 0u, 0u);
        const long   warpFrame0 = (long)fb * kThreads + (tid & ~31);
        int          staged = 0;  // quads waiting in this warp's staging tile
        for (int m4 = mix0; m4 < mix1; m4 += 4) {
            const uint32_t sets[4] = {mk.x, mk.y, mk.z, mk.w};
            if (m4 + 4 < mix1 && t < p.T)  // the next quad's sets are on their way while this one is evaluated
                mk = __ldg(w + (size_t)((m4 + 4) >> 2) * p.pitch);
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* base = rows + first[m4 - mix0 + q];
                uint32_t     set  = sets[q];
                float        best = FLT_MAX;
                while (set) {
                    const int j = __ffs((int)set) - 1;
                    set &= set - 1;
                    const float rs = batch_row_score<NB, FUSE>(base + j * ROWF, x);
                    best           = best < rs ? best : rs;
                }
                // (a mixture none of whose candidates was scored keeps FLT_MAX -- or gets the preselection scorer's back-off
                // score, `if (s == max) s = backoffScore`; +inf / NaN pass through as the direct kernel leaves them)
                o[q] = best < FLT_MAX ? __fmul_rn(best, 0.5f) : (best == FLT_MAX ? p.emptyScore : best);
            }
            // Scores leave through a per-warp staging tile [32 frames][kStageQuads quads]: a thread's own 16 bytes per
            // quad would be a store instruction touching 32 different lines, 1 KB apart; after the transpose four lanes
            // share a frame and an instruction writes 64 contiguous bytes into each of 8 rows (a quarter of the LSU
            // wavefronts, whole sectors -- which is also what a peer's window behind NVLink wants to see).
            float* mine = stage + (tid >> 5) * (32 * kStagePitch);
            *reinterpret_cast<float4*>(mine + (tid & 31) * kStagePitch + staged * 4) = make_float4(o[0], o[1], o[2], o[3]);
            ++staged;
            if (staged == kStageQuads || m4 + 4 >= mix1) {
                __syncwarp();
                const int mFirst = m4 + 4 - staged * 4;
                for (int k = tid & 31; k < 32 * staged; k += 32) {
                    const int  r = k / staged, c = k - r * staged;
                    const long tf = warpFrame0 + r;
                    if (tf < p.T) {
                        const float4 val = *reinterpret_cast<const float4*>(mine + r * kStagePitch + c * 4);
                        const size_t at  = (size_t)tf * p.nMix + mFirst + c * 4;
                        *reinterpret_cast<float4*>(p.scores + at) = val;
                        for (int e = 0; e < p.nExtra; ++e)  // the same 64 contiguous bytes per frame into every peer's window
                            *reinterpret_cast<float4*>(p.extra[e] + at) = val;
                    }
                }
                __syncwarp();
                staged = 0;
            }
        }
    }
}

typedef void (*GmmRefineKernel)(const RefineParams);
template<int N>
GmmRefineKernel pick_refine(bool fuse) {
    return fuse ? gmm_refine_kernel<N, true> : gmm_refine_kernel<N, false>;
}
GmmRefineKernel refine_kernel_for(int nb, bool fuse) {
    switch (nb) {
        case 1: return pick_refine<1>(fuse);
        case 2: return pick_refine<2>(fuse);
        case 3: return pick_refine<3>(fuse);
        case 4: return pick_refine<4>(fuse);
        case 5: return pick_refine<5>(fuse);
        case 6: return pick_refine<6>(fuse);
        case 7: return pick_refine<7>(fuse);
        case 8: return pick_refine<8>(fuse);
    }
    return nullptr;
}

// ------------------------------------------------------------------------------------------
// DIAG_MAX, second half of the exact two-pass scorer: as gmm_refine_kernel, for Mm::GaussDiagonalMaximumFeatureScorer.
// Refinement rows [ mu (4 NQ) | 1/sigma (4 NQ) | w | logNorm ] = 8 NQ + 2 floats (an odd number of 8-byte words).  The
// arithmetic is gmm_diag_kernel's for one frame: packed f32x2 over the full quads in the reference's SSE lane order,
// the horizontal add, the tail dimensions sequentially, then the f64 sum of the three f32 terms compared against the
// f32 running best (calculateScoreAndDensity :130-137).  That comparison is order dependent (the best is narrowed to
// f32 after every update), so the candidates are walked in ascending density order: a density outside the candidate
// set can only hold the best before the first candidate is seen, and never after (its score exceeds every
// candidate's by more than the narrowing error), so the final (score, density) pair is the reference's.
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int refine_diag_pitch(int nq) {
    return nq * 8 + 2;
}
// POOLED (every density uses the same covariance -- the usual RASR model): a row is [ mean | (w, logNorm) ] and the
// common 1/sqrt(var) is read from one shared array -- the same address for every lane, a broadcast costing one wavefront
// where the per-lane gather of a row segment costs two; 60 instead of 82 wavefronts per candidate, and half the rows'
// shared memory (fewer mixture groups).  The arithmetic is unchanged.
__host__ __device__ constexpr int refine_diag_pooled_pitch(int nq) {
    return nq * 4 + 2;  // an odd number (2 NQ + 1) of 8-byte words: conflict-free per-lane row gathers
}

template<int NQ, bool FUSE, bool POOLED>
__global__ void __launch_bounds__(kThreads, 2) gmm_refine_diag_kernel(const RefineParams p) {
    constexpr int ROWF = POOLED ? refine_diag_pooled_pitch(NQ) : refine_diag_pitch(NQ);
    constexpr int TAILW = POOLED ? 2 * NQ : 4 * NQ;  // 8-byte word of (w, logNorm) within a row
    __shared__ uint64_t sIsd[NQ * 2];
    if (POOLED && threadIdx.x < NQ * 2)
        sIsd[threadIdx.x] = pack2(p.pooledIsd[2 * threadIdx.x], p.pooledIsd[2 * threadIdx.x + 1]);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int g    = blockIdx.x % p.nGroups;
    const int row0 = p.grp_row[g], row1 = p.grp_row[g + 1];
    const int mix0 = p.grp_mix[g], mix1 = p.grp_mix[g + 1];
    uint64_t*      bar    = reinterpret_cast<uint64_t*>(smem_raw);
    int*           first  = reinterpret_cast<int*>(smem_raw + 16);
    unsigned char* region = smem_raw + 16 + (((size_t)(mix1 - mix0) * sizeof(int) + 15) & ~(size_t)15);
    const size_t   gBegin = (size_t)row0 * ROWF * 4, gEnd = (size_t)row1 * ROWF * 4;
    const size_t   cBegin = gBegin & ~(size_t)15, cEnd = (gEnd + 15) & ~(size_t)15;
    const float*   rows   = reinterpret_cast<const float*>(region + (gBegin - cBegin));
    float*         stage  = reinterpret_cast<float*>(region + (((gEnd - cBegin) + 47) & ~(size_t)15));
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    for (int m = tid; m < mix1 - mix0; m += kThreads)
        first[m] = (p.mix_row[mix0 + m] - row0) * ROWF;
    __syncthreads();
    if (tid == 0) {
        const size_t total = cEnd - cBegin;
        mbar_expect_tx(bar, (uint32_t)total);
        for (size_t off = 0; off < total; off += 65536) {
            const uint32_t bytes = (uint32_t)(total - off < 65536 ? total - off : 65536);
            bulk_g2s(region + off, reinterpret_cast<const unsigned char*>(p.rows) + cBegin + off, bytes, bar);
        }
    }
    mbar_wait(bar, 0);

    const int nFullQ = p.dim >> 2, nTail = p.dim & 3;
    const int member   = blockIdx.x / p.nGroups;
    const int nMembers = ((int)gridDim.x - g + p.nGroups - 1) / p.nGroups;
    for (int fb = member; fb < p.nFrameBlocks; fb += nMembers) {
        const long t  = (long)fb * kThreads + tid;
        const long tc = t < p.T ? t : p.T - 1;
        uint64_t   xp[NQ * 2];
        float      xt[4];  // scalar copies of the last quad for the tail dimensions
#pragma unroll
        for (int i = 0; i < NQ * 2; ++i) {
            const float a = __ldg(p.xT + (size_t)(2 * i) * p.pitch + tc), b = __ldg(p.xT + (size_t)(2 * i + 1) * p.pitch + tc);
            xp[i] = pack2(a, b);
            if (i >= NQ * 2 - 2) {
                xt[2 * (i - (NQ * 2 - 2))]     = a;
                xt[2 * (i - (NQ * 2 - 2)) + 1] = b;
            }
        }
        const uint4* w  = reinterpret_cast<const uint4*>(p.words) + tc;
        uint4        mk = t < p.T ? __ldg(w + (size_t)(mix0 >> 2) * p.pitch) : make_uint4(0u, 0u, 0u, 0u);
        const long   warpFrame0 = (long)fb * kThreads + (tid & ~31);
        int          staged = 0;
        for (int m4 = mix0; m4 < mix1; m4 += 4) {
            const uint32_t sets[4] = {mk.x, mk.y, mk.z, mk.w};
            if (m4 + 4 < mix1 && t < p.T)
                mk = __ldg(w + (size_t)((m4 + 4) >> 2) * p.pitch);
            float    o[4];
            uint32_t ob[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* base = rows + first[m4 - mix0 + q];
                uint32_t     set  = sets[q];
                float        best = FLT_MAX;
                uint32_t     bd   = 0xffffffffu;
                while (set) {
                    const int j = __ffs((int)set) - 1;
                    set &= set - 1;
                    const uint64_t* r    = reinterpret_cast<const uint64_t*>(base + j * ROWF);
                    uint64_t        s[2] = {0ull, 0ull};
#pragma unroll
                    for (int qd = 0; qd < NQ; ++qd) {
                        if (qd < nFullQ) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint64_t e = mul2(sub2(r[2 * qd + h], xp[2 * qd + h]),
                                                        POOLED ? sIsd[2 * qd + h] : r[2 * NQ + 2 * qd + h]);
                                s[h]             = FUSE ? fma2(e, e, s[h]) : sqadd2(e, s[h]);
                            }
                        }
                    }
                    float d = __fadd_rn(0.0f, __fadd_rn(__fadd_rn(lo2(s[0]), hi2(s[0])), __fadd_rn(lo2(s[1]), hi2(s[1]))));
                    if (nTail) {
                        const uint64_t ma = r[2 * NQ - 2], mb = r[2 * NQ - 1];
                        const uint64_t va = POOLED ? sIsd[2 * NQ - 2] : r[4 * NQ - 2], vb = POOLED ? sIsd[2 * NQ - 1] : r[4 * NQ - 1];
                        const float    mt[4] = {lo2(ma), hi2(ma), lo2(mb), hi2(mb)};
                        const float    vt[4] = {lo2(va), hi2(va), lo2(vb), hi2(vb)};
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (k < nTail) {
                                const float e = __fmul_rn(__fsub_rn(mt[k], xt[k]), vt[k]);
                                d             = sq_acc(e, d, FUSE);
                            }
                    }
                    const uint64_t tail = r[TAILW];  // (w, logNorm)
                    const double   sc   = ((double)lo2(tail) + (double)hi2(tail)) + (double)d;
                    if ((double)best > sc) {
                        best = (float)sc;
                        bd   = (uint32_t)j;
                    }
                }
                o[q]  = __fmul_rn(0.5f, best);
                ob[q] = bd;
            }
            if (p.best && t < p.T)
                *reinterpret_cast<uint4*>(p.best + (size_t)t * p.nMix + m4) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
            float* mine = stage + (tid >> 5) * (32 * kStagePitch);
            *reinterpret_cast<float4*>(mine + (tid & 31) * kStagePitch + staged * 4) = make_float4(o[0], o[1], o[2], o[3]);
            ++staged;
            if (staged == kStageQuads || m4 + 4 >= mix1) {
                __syncwarp();
                const int mFirst = m4 + 4 - staged * 4;
                for (int k = tid & 31; k < 32 * staged; k += 32) {
                    const int  r = k / staged, c = k - r * staged;
                    const long tf = warpFrame0 + r;
                    if (tf < p.T) {
                        const float4 val = *reinterpret_cast<const float4*>(mine + r * kStagePitch + c * 4);
                        const size_t at  = (size_t)tf * p.nMix + mFirst + c * 4;
                        *reinterpret_cast<float4*>(p.scores + at) = val;
                        for (int e = 0; e < p.nExtra; ++e)
                            *reinterpret_cast<float4*>(p.extra[e] + at) = val;
                    }
                }
                __syncwarp();
                staged = 0;
            }
        }
    }
}

template<int N>
GmmRefineKernel pick_refine_diag(bool fuse, bool pooled) {
    if (pooled)
        return fuse ? gmm_refine_diag_kernel<N, true, true> : gmm_refine_diag_kernel<N, false, true>;
    return fuse ? gmm_refine_diag_kernel<N, true, false> : gmm_refine_diag_kernel<N, false, false>;
}
GmmRefineKernel refine_diag_kernel_for(int nq, bool fuse, bool pooled) {
    switch (nq) {
        case 1: return pick_refine_diag<1>(fuse, pooled);
        case 2: return pick_refine_diag<2>(fuse, pooled);
        case 3: return pick_refine_diag<3>(fuse, pooled);
        case 4: return pick_refine_diag<4>(fuse, pooled);
        case 5: return pick_refine_diag<5>(fuse, pooled);
        case 6: return pick_refine_diag<6>(fuse, pooled);
        case 7: return pick_refine_diag<7>(fuse, pooled);
        case 8: return pick_refine_diag<8>(fuse, pooled);
        case 9: return pick_refine_diag<9>(fuse, pooled);
        case 10: return pick_refine_diag<10>(fuse, pooled);
    }
    return nullptr;
}

// ------------------------------------------------------------------------------------------
// DIAG_MAX / DIAG_SUM: row = [ mu (NQ*4) | isd (NQ*4) | -2 log w * scale | logNorm | flags | 0 ]
// distance(): 4 SSE lanes over the first (dim & ~3) dims, hadd, then the tail dims sequentially
// (src/Mm/GaussDiagonalMaximumFeatureScorer.cc:144-233, SSE3 branch)
// ------------------------------------------------------------------------------------------
template<int NQ, bool FUSE, bool SUM>
__global__ void __launch_bounds__(kThreads, (NQ <= 10) ? 2 : 1) gmm_diag_kernel(const GmmParams p) {
    constexpr int ROWF  = NQ * 8 + 4;
    constexpr int CHUNK = kChunkRows * ROWF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float*    buf = reinterpret_cast<float*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + sizeof(float) * kStages * CHUNK);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s)
            mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    uint32_t  seq    = 0;
    const int nItems = p.nGroups * p.nFrameBlocks;
    const int nFullQ = p.dim >> 2;  // quads handled by the SSE loop
    const int nTail  = p.dim & 3;

    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int  g    = item / p.nFrameBlocks;
        const int  fb   = item - g * p.nFrameBlocks;
        const int  row0 = p.grp_row[g], row1 = p.grp_row[g + 1];
        int        mix  = p.grp_mix[g];
        const int  nCh  = (row1 - row0 + kChunkRows - 1) / kChunkRows;
        const long t0   = (long)fb * kFramesPerBlock + tid;
        const long t1   = t0 + kThreads;

        if (tid == 0) {
            for (int c = 0; c < kStages - 1 && c < nCh; ++c) {
                const int      r0    = row0 + c * kChunkRows;
                const int      nr    = min(kChunkRows, row1 - r0);
                const uint32_t bytes = (uint32_t)nr * ROWF * 4u;
                const uint32_t st    = (seq + c) % kStages;
                mbar_expect_tx(&bar[st], bytes);
                bulk_g2s(buf + st * CHUNK, p.rows + (size_t)r0 * ROWF, bytes, &bar[st]);
            }
        }

        uint64_t xp0[NQ * 2], xp1[NQ * 2];  // dims (2i, 2i+1) packed
        float    x0[4], x1[4];              // scalar copies of the last quad for the tail dims
        {
            const long   ta  = t0 < p.T ? t0 : p.T - 1;
            const long   tb  = t1 < p.T ? t1 : p.T - 1;
            const float* fa  = p.feats + (size_t)ta * p.dim;
            const float* fbp = p.feats + (size_t)tb * p.dim;
#pragma unroll
            for (int i = 0; i < NQ * 2; ++i) {
                const int   d  = 2 * i;
                const float a0 = d < p.dim ? __ldg(fa + d) : 0.0f, a1 = d + 1 < p.dim ? __ldg(fa + d + 1) : 0.0f;
                const float b0 = d < p.dim ? __ldg(fbp + d) : 0.0f, b1 = d + 1 < p.dim ? __ldg(fbp + d + 1) : 0.0f;
                xp0[i] = pack2(a0, a1);
                xp1[i] = pack2(b0, b1);
                if (i >= NQ * 2 - 2) {
                    x0[2 * (i - (NQ * 2 - 2))]     = a0;
                    x0[2 * (i - (NQ * 2 - 2)) + 1] = a1;
                    x1[2 * (i - (NQ * 2 - 2))]     = b0;
                    x1[2 * (i - (NQ * 2 - 2)) + 1] = b1;
                }
            }
        }

        // running state over the densities of the current mixture
        float    best0 = FLT_MAX, best1 = FLT_MAX;
        uint32_t bd0 = 0xffffffffu, bd1 = 0xffffffffu;
        float    se0 = 0.0f, se1 = 0.0f;  // SUM: sum of exp(best - s_k), rescaled when best moves
        uint32_t k = 0;                   // density index within the mixture

        for (int c = 0; c < nCh; ++c) {
            if (tid == 0 && c + kStages - 1 < nCh) {
                const int      cc    = c + kStages - 1;
                const int      r0    = row0 + cc * kChunkRows;
                const int      nr    = min(kChunkRows, row1 - r0);
                const uint32_t bytes = (uint32_t)nr * ROWF * 4u;
                const uint32_t st    = (seq + cc) % kStages;
                mbar_expect_tx(&bar[st], bytes);
                bulk_g2s(buf + st * CHUNK, p.rows + (size_t)r0 * ROWF, bytes, &bar[st]);
            }
            const uint32_t n  = seq + c;
            const uint32_t st = n % kStages;
            mbar_wait(&bar[st], (n / kStages) & 1u);
            const float* chunk = buf + st * CHUNK;
            const int    nr    = min(kChunkRows, row1 - (row0 + c * kChunkRows));

            for (int r = 0; r < nr; ++r) {
                const float4* row  = reinterpret_cast<const float4*>(chunk + r * ROWF);
                const float4  tail = row[NQ * 2];
                const int     flags = __float_as_int(tail.z);
                float         d0 = 0.0f, d1 = 0.0f;
                if (!(flags & 2)) {  // not a placeholder row of an empty mixture
                    // packed f32x2 (FADD2 / FMUL2 / FFMA2): SSE lanes (0,1) and (2,3) share a register pair; every lane
                    // is still one IEEE sub, mul and fma (or mul + add) in the reference's order
                    uint64_t s0[2] = {0ull, 0ull}, s1[2] = {0ull, 0ull};
                    const ulonglong2* rowp = reinterpret_cast<const ulonglong2*>(row);
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        if (q < nFullQ) {
                            const ulonglong2 m4 = rowp[q], v4 = rowp[NQ + q];
                            const uint64_t   m[2] = {m4.x, m4.y}, v[2] = {v4.x, v4.y};
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint64_t e0 = mul2(sub2(m[j], xp0[2 * q + j]), v[j]);
                                const uint64_t e1 = mul2(sub2(m[j], xp1[2 * q + j]), v[j]);
                                s0[j]             = FUSE ? fma2(e0, e0, s0[j]) : sqadd2(e0, s0[j]);
                                s1[j]             = FUSE ? fma2(e1, e1, s1[j]) : sqadd2(e1, s1[j]);
                            }
                        }
                    }
                    // hadd(sum,sum) -> (s0+s1, s2+s3); result = 0 + ((s0+s1) + (s2+s3))
                    d0 = __fadd_rn(0.0f, __fadd_rn(__fadd_rn(lo2(s0[0]), hi2(s0[0])), __fadd_rn(lo2(s0[1]), hi2(s0[1]))));
                    d1 = __fadd_rn(0.0f, __fadd_rn(__fadd_rn(lo2(s1[0]), hi2(s1[0])), __fadd_rn(lo2(s1[1]), hi2(s1[1]))));
                    if (nTail) {
                        const float4 m4 = row[NQ - 1], v4 = row[2 * NQ - 1];
                        const float  m[4] = {m4.x, m4.y, m4.z, m4.w};
                        const float  v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            if (j < nTail) {
                                const float e0 = __fmul_rn(__fsub_rn(m[j], x0[j]), v[j]);
                                const float e1 = __fmul_rn(__fsub_rn(m[j], x1[j]), v[j]);
                                d0             = sq_acc(e0, d0, FUSE);
                                d1             = sq_acc(e1, d1, FUSE);
                            }
                        }
                    }
                    if (!SUM) {
                        // f64 sum of the three f32 terms, compared against the f32 running best (:130-137)
                        const double base = (double)tail.x + (double)tail.y;
                        const double sc0 = base + (double)d0, sc1 = base + (double)d1;
                        if ((double)best0 > sc0) {
                            best0 = (float)sc0;
                            bd0   = k;
                        }
                        if ((double)best1 > sc1) {
                            best1 = (float)sc1;
                            bd1   = k;
                        }
                    }
                    else {
                        // s_k = 0.5 * (w + logNorm + dist) in f32 (:272-276); log-sum-exp around the running
                        // minimum (the reference's two passes folded into one; exp/log are libm anyway)
                        const float a0 = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(tail.x, tail.y), d0));
                        const float a1 = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(tail.x, tail.y), d1));
                        if (best0 > a0) {
                            se0   = se0 * expf(a0 - best0) + 1.0f;
                            best0 = a0;
                            bd0   = k;
                        }
                        else
                            se0 += expf(best0 - a0);
                        if (best1 > a1) {
                            se1   = se1 * expf(a1 - best1) + 1.0f;
                            best1 = a1;
                            bd1   = k;
                        }
                        else
                            se1 += expf(best1 - a1);
                    }
                    ++k;
                }
                if (flags & 1) {  // emit
                    float v0, v1;
                    if (!SUM) {
                        v0 = __fmul_rn(0.5f, best0);
                        v1 = __fmul_rn(0.5f, best1);
                    }
                    else {
                        v0 = best0 - logf(se0);
                        v1 = best1 - logf(se1);
                    }
                    if (t0 < p.T) {
                        p.scores[(size_t)t0 * p.nMix + mix] = v0;
                        if (p.best)
                            p.best[(size_t)t0 * p.nMix + mix] = bd0;
                    }
                    if (t1 < p.T) {
                        p.scores[(size_t)t1 * p.nMix + mix] = v1;
                        if (p.best)
                            p.best[(size_t)t1 * p.nMix + mix] = bd1;
                    }
                    best0 = FLT_MAX;
                    best1 = FLT_MAX;
                    bd0 = bd1 = 0xffffffffu;
                    se0 = se1 = 0.0f;
                    k         = 0;
                    ++mix;
                }
            }
            __syncthreads();
        }
        seq += (uint32_t)nCh;
    }
}

typedef void (*GmmKernel)(const GmmParams);

template<int N>
GmmKernel pick_batch(bool fuse, int fpt) {
    if (fpt == 1)
        return fuse ? gmm_batch_kernel<N, true, 1> : gmm_batch_kernel<N, false, 1>;
    if constexpr (N <= 5) {
        if (fpt == 4)
            return fuse ? gmm_batch_kernel<N, true, 4> : gmm_batch_kernel<N, false, 4>;
    }
    return fuse ? gmm_batch_kernel<N, true, 2> : gmm_batch_kernel<N, false, 2>;
}
template<int N>
GmmKernel pick_diag(bool fuse, bool sum) {
    if (sum)
        return fuse ? gmm_diag_kernel<N, true, true> : gmm_diag_kernel<N, false, true>;
    return fuse ? gmm_diag_kernel<N, true, false> : gmm_diag_kernel<N, false, false>;
}

GmmKernel batch_kernel_for(int nb, bool fuse, int fpt) {
    switch (nb) {
        case 1: return pick_batch<1>(fuse, fpt);
        case 2: return pick_batch<2>(fuse, fpt);
        case 3: return pick_batch<3>(fuse, fpt);
        case 4: return pick_batch<4>(fuse, fpt);
        case 5: return pick_batch<5>(fuse, fpt);
        case 6: return pick_batch<6>(fuse, fpt);
        case 7: return pick_batch<7>(fuse, fpt);
        case 8: return pick_batch<8>(fuse, fpt);
    }
    return nullptr;
}
GmmKernel diag_kernel_for(int nq, bool fuse, bool sum) {
    switch (nq) {
        case 1: return pick_diag<1>(fuse, sum);
        case 2: return pick_diag<2>(fuse, sum);
        case 3: return pick_diag<3>(fuse, sum);
        case 4: return pick_diag<4>(fuse, sum);
        case 5: return pick_diag<5>(fuse, sum);
        case 6: return pick_diag<6>(fuse, sum);
        case 7: return pick_diag<7>(fuse, sum);
        case 8: return pick_diag<8>(fuse, sum);
        case 9: return pick_diag<9>(fuse, sum);
        case 10: return pick_diag<10>(fuse, sum);
        case 11: return pick_diag<11>(fuse, sum);
        case 12: return pick_diag<12>(fuse, sum);
        case 13: return pick_diag<13>(fuse, sum);
        case 14: return pick_diag<14>(fuse, sum);
        case 15: return pick_diag<15>(fuse, sum);
        case 16: return pick_diag<16>(fuse, sum);
    }
    return nullptr;
}

}  // namespace

// ==========================================================================================
// host side
// ==========================================================================================

struct rb_gmm_int;  // gmm_int.cu
int  rb_gmm_int_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream, rb_gmm_int** out,
                       bool simd = false);
void rb_gmm_int_destroy(rb_gmm_int* h);
int  rb_gmm_int_score(rb_gmm_int* h, const float* d_feats, long T, float* d_scores, cudaStream_t stream);

struct rb_gmm_presel;  // gmm_presel.cu
int  rb_gmm_presel_create(const rb_mixture_set* ms, bool fuse, const rb::DeviceInfo& dev, cudaStream_t stream,
                          rb_gmm_presel** out);
void rb_gmm_presel_destroy(rb_gmm_presel* h);
int  rb_gmm_presel_score(rb_gmm_presel* h, const float* d_feats, long T, float* d_scores, cudaStream_t stream);
int  rb_gmm_presel_configure(rb_gmm_presel* h, int clusters, int select, int iterations, float backoff, cudaStream_t s);
void rb_gmm_presel_clustering(const rb_gmm_presel* h, uint32_t* cluster_of, float* cluster_means, int* n_clusters);
int  rb_gmm_presel_select(rb_gmm_presel* h, const float* d_feats, long T, const uint32_t** active,
                          const uint8_t** cluster_of, const uint32_t** offsets, float* backoff, cudaStream_t stream);

struct rb_gmm_presel_int;  // gmm_presel_int.cu
int  rb_gmm_presel_int_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream,
                              rb_gmm_presel_int** out);
void rb_gmm_presel_int_destroy(rb_gmm_presel_int* h);
int  rb_gmm_presel_int_score(rb_gmm_presel_int* h, const float* d_feats, long T, float* d_scores, cudaStream_t stream);
int  rb_gmm_presel_int_configure(rb_gmm_presel_int* h, int clusters, int select, int iterations, cudaStream_t s);
void rb_gmm_presel_int_clustering(const rb_gmm_presel_int* h, uint32_t* cluster_of, float* cluster_means,
                                  int* n_clusters);

struct rb_gmm_simd;  // gmm_simd.cu
int  rb_gmm_simd_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream, rb_gmm_simd** out);
void rb_gmm_simd_destroy(rb_gmm_simd* h);
int  rb_gmm_simd_score(rb_gmm_simd* h, const float* d_feats, long T, float* d_scores, uint32_t* d_best, cudaStream_t stream);

struct rb_gmm_tensor;  // gmm_tensor.cu
int  rb_gmm_tensor_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream,
                          rb_gmm_tensor** out);
void rb_gmm_tensor_destroy(rb_gmm_tensor* t);
int  rb_gmm_tensor_score(rb_gmm_tensor* t, const float* d_feats, long T, float* d_scores, cudaStream_t stream);
bool rb_gmm_tensor_screenable(const rb_gmm_tensor* t);
int  rb_gmm_tensor_reserve(rb_gmm_tensor* t, long frames, bool screen);
int  rb_gmm_tensor_create_diag(const rb_mixture_set* ms, const float* rows, int rowf, int nq, const rb::DeviceInfo& dev,
                               cudaStream_t stream, rb_gmm_tensor** out);
long rb_gmm_tensor_chunk(const rb_gmm_tensor* t);
int  rb_gmm_tensor_screen(rb_gmm_tensor* t, const float* d_feats, long n, const uint32_t** words, const float** xT,
                          long* pitch, cudaStream_t stream, cudaEvent_t after_split);
int  rb_gmm_tensor_split(rb_gmm_tensor* t, const float* d_feats, long n, uint32_t** words, const float** xT, long* pitch,
                         cudaStream_t stream);

struct rb_gmm {
    rb::DeviceInfo dev;
    int            mode = 0;
    bool           fuse = true;
    int            dim = 0, nMix = 0;
    int            nUnits = 0;  // NB (batch) or NQ (diag)
    int            rowf = 0;
    int            nRows = 0;
    cudaStream_t   stream = nullptr;
    GmmKernel      kernel = nullptr;
    size_t         smemBytes = 0;
    int            ctasPerSm = 1;
    int            framesPerBlock = kFramesPerBlock;
    // BATCH_FLOAT: launch variants with 1 / 2 / 4 frames per thread, picked per call by predicted efficiency
    struct Variant {
        GmmKernel kernel = nullptr;
        int       framesPerBlock = 0, ctasPerSm = 0;
        double    rate = 0;  // measured relative arithmetic rate at full occupancy (C2 shape, B200)
    } variants[3];
    int nVariants = 0, curVariant = -1;

    std::vector<int> rowsOfMixture;  // rows each mixture occupies (>= 1)
    int              curGroups = -1;
    std::vector<int> groupsOf;       // number of groups the table built for G actually has
    // host-pointer entry point: copy streams and events are created once
    cudaStream_t             sIn = nullptr, sOut = nullptr;
    std::vector<cudaEvent_t> events;

    rb::DevBuf<float>    dRows, dIsd;
    rb::DevBuf<int>      dGrpRow, dGrpMix;
    rb::DevBuf<float>    dFeats, dScores;  // staging for the host-pointer entry point
    rb::DevBuf<uint32_t> dBest;
    rb_gmm_tensor*       tensor = nullptr;
    // BATCH_FLOAT on large batches: tensor-core screening + exact evaluation of the surviving densities (same bits as
    // the direct kernel).  refGroups = 0: not available for this model.
    GmmRefineKernel      refine = nullptr;
    int                  refGroups = 0, refSlots = 0;
    size_t               refSmem = 0;
    long                 exactMinFrames = 1;  // measured: the two-pass route is faster than the direct kernel from 1 frame on
                                              // (23 vs 31 us per call up to 1000 frames, 27 vs 117 us for DIAG_MAX); RB_GMM_EXACT_MIN_FRAMES
    rb::DevBuf<int>      dMixRow;
    rb::DevBuf<float>    dRefRows, dRefIsd;
    bool                 refPooled = false;  // DIAG_MAX refinement rows without the per-row 1/sqrt(var) copy
    rb::HostStager       stager;  // host-buffer calls with pageable score buffers
    rb::PinnedBuf<float> hFeats;  // ... and pageable feature buffers: copied here by several cores, then DMA
    cudaEvent_t          tev[4] = {nullptr, nullptr, nullptr, nullptr};  // rb_gmm_set_timing: around the three kernels
    bool                 timed = false;
    rb_gmm_int*          quantised = nullptr;
    rb_gmm_presel*       presel    = nullptr;
    rb_gmm_presel_int*   preselInt = nullptr;
    rb_gmm_simd*         simd      = nullptr;

    ~rb_gmm() {
        if (simd)
            rb_gmm_simd_destroy(simd);
        if (tensor)
            rb_gmm_tensor_destroy(tensor);
        if (quantised)
            rb_gmm_int_destroy(quantised);
        if (presel)
            rb_gmm_presel_destroy(presel);
        if (preselInt)
            rb_gmm_presel_int_destroy(preselInt);
        for (cudaEvent_t e : events)
            cudaEventDestroy(e);
        for (cudaEvent_t e : tev)
            if (e)
                cudaEventDestroy(e);
        if (sIn)
            cudaStreamDestroy(sIn);
        if (sOut)
            cudaStreamDestroy(sOut);
        if (stream)
            cudaStreamDestroy(stream);
    }
};

namespace {

// Mm::gaussLogNormFactor: D*log(2 pi) + sum log|var|  (src/Mm/Utilities.hh:53-76), f64
double log_norm_factor(const float* var, unsigned dim) {
    double s = 0;
    for (unsigned d = 0; d < dim; ++d)
        s += std::log(std::fabs((double)var[d]));
    return (double)dim * std::log(2.0 * M_PI) + s;
}

// Mm::inverseSquareRoot<f32> (src/Mm/Utilities.hh:86-91)
inline float inv_sqrt(float v) {
    return 1.0f / (float)std::sqrt((double)v);
}

int validate(const rb_mixture_set* ms) {
    RB_REQUIRE(ms != nullptr, "mixture set is NULL");
    RB_REQUIRE(ms->dim >= 1, "mixture set dimension is 0");
    RB_REQUIRE(ms->n_mixtures >= 1, "mixture set has no mixtures");
    RB_REQUIRE(ms->mix_offsets && ms->dens_mean && ms->dens_cov && ms->means && ms->variances,
               "mixture set has NULL tables");
    const uint32_t nEntries = ms->mix_offsets[ms->n_mixtures];
    RB_REQUIRE(nEntries == 0 || (ms->mix_density && ms->mix_log_weight), "mixture entries missing");
    for (uint32_t m = 0; m < ms->n_mixtures; ++m)
        RB_REQUIRE(ms->mix_offsets[m] <= ms->mix_offsets[m + 1], "mix_offsets not monotone at mixture %u", m);
    for (uint32_t e = 0; e < nEntries; ++e) {
        const uint32_t d = ms->mix_density[e];
        RB_REQUIRE(d < ms->n_densities, "mixture entry %u refers to density %u >= %u", e, d, ms->n_densities);
    }
    // every density, referenced by a mixture or not: the quantisation scale and the clusterings walk all of them
    for (uint32_t d = 0; d < ms->n_densities; ++d) {
        RB_REQUIRE(ms->dens_mean[d] < ms->n_means, "density %u refers to mean %u >= %u", d, ms->dens_mean[d],
                   ms->n_means);
        RB_REQUIRE(ms->dens_cov[d] < ms->n_covariances, "density %u refers to covariance %u >= %u", d,
                   ms->dens_cov[d], ms->n_covariances);
    }
    return RB_OK;
}

inline float int_bits(int v) {
    float f;
    std::memcpy(&f, &v, 4);
    return f;
}

// BATCH_FLOAT rows, following BatchFloatFeatureScorer::init (src/Mm/BatchFeatureScorer.cc:164-197)
int build_batch_rows(rb_gmm* h, const rb_mixture_set* ms, std::vector<float>& rows, std::vector<float>& isd) {
    if (ms->n_covariances != 1) {
        rb::set_error("batch feature scorer supports only one globally pooled covariance (got %u); use "
                      "RB_GMM_DIAG_MAX",
                      ms->n_covariances);
        return RB_ERR_UNSUPPORTED;
    }
    const unsigned D    = ms->dim;
    const int      NB   = (int)((D + 7) / 8);
    const int      rowf = NB * 8 + 4;
    h->nUnits           = NB;
    h->rowf             = rowf;
    isd.assign((size_t)NB * 8, 0.0f);
    for (unsigned d = 0; d < D; ++d)
        isd[d] = inv_sqrt(ms->variances[d]);
    const float logNorm = (float)log_norm_factor(ms->variances, D);
    h->rowsOfMixture.assign(ms->n_mixtures, 0);
    rows.clear();
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        const uint32_t e0 = ms->mix_offsets[m], e1 = ms->mix_offsets[m + 1];
        const uint32_t n  = e1 - e0;
        const uint32_t nr = n ? n : 1;  // an empty mixture keeps one placeholder row scoring FLT_MAX
        h->rowsOfMixture[m] = (int)nr;
        const size_t base   = rows.size();
        rows.resize(base + (size_t)nr * rowf, 0.0f);
        for (uint32_t i = 0; i < nr; ++i) {
            float* row = rows.data() + base + (size_t)i * rowf;
            if (n) {
                const uint32_t dns = ms->mix_density[e0 + i];
                const float*   mu  = ms->means + (size_t)ms->dens_mean[dns] * D;
                for (unsigned d = 0; d < D; ++d)
                    row[d] = mu[d] * isd[d];
                // f32 logNorm minus f64 2*logWeight, narrowed (:193)
                row[NB * 8] = (float)((double)logNorm - 2 * ms->mix_log_weight[e0 + i]);
            }
            else {
                row[NB * 8] = FLT_MAX;
            }
            row[NB * 8 + 2] = int_bits(i + 1 == nr ? 1 : 0);
        }
    }
    return RB_OK;
}

// DIAG rows, following GaussDiagonalMaximumFeatureScorer's element caches
// (MixtureFeatureScorerElement.cc:21-34, CovarianceFeatureScorerElement.cc:21-51)
int build_diag_rows(rb_gmm* h, const rb_mixture_set* ms, float mixtureWeightScale, float gaussianScaleParam,
                    std::vector<float>& rows) {
    const unsigned D    = ms->dim;
    const int      NQ   = (int)((D + 3) / 4);
    const int      rowf = NQ * 8 + 4;
    h->nUnits           = NQ;
    h->rowf             = rowf;
    const float gaussianScale = (float)std::sqrt((double)gaussianScaleParam);
    std::vector<float> isd((size_t)ms->n_covariances * D);
    std::vector<float> logNorm(ms->n_covariances);
    for (uint32_t c = 0; c < ms->n_covariances; ++c) {
        const float* var = ms->variances + (size_t)c * D;
        for (unsigned d = 0; d < D; ++d)
            isd[(size_t)c * D + d] = inv_sqrt(var[d]) * gaussianScale;
        const float lnf = (float)log_norm_factor(var, D);
        logNorm[c]      = lnf * (gaussianScale * gaussianScale);
    }
    h->rowsOfMixture.assign(ms->n_mixtures, 0);
    rows.clear();
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        const uint32_t e0 = ms->mix_offsets[m], e1 = ms->mix_offsets[m + 1];
        const uint32_t n  = e1 - e0;
        const uint32_t nr = n ? n : 1;
        h->rowsOfMixture[m] = (int)nr;
        const size_t base   = rows.size();
        rows.resize(base + (size_t)nr * rowf, 0.0f);
        for (uint32_t i = 0; i < nr; ++i) {
            float* row   = rows.data() + base + (size_t)i * rowf;
            int    flags = (i + 1 == nr) ? 1 : 0;
            if (n) {
                const uint32_t dns = ms->mix_density[e0 + i];
                const uint32_t cov = ms->dens_cov[dns];
                const float*   mu  = ms->means + (size_t)ms->dens_mean[dns] * D;
                for (unsigned d = 0; d < D; ++d) {
                    row[d]          = mu[d];
                    row[NQ * 4 + d] = isd[(size_t)cov * D + d];
                }
                const float v   = (float)(-2 * ms->mix_log_weight[e0 + i]);
                row[NQ * 8]     = v * mixtureWeightScale;
                row[NQ * 8 + 1] = logNorm[cov];
            }
            else {
                flags |= 2;
            }
            row[NQ * 8 + 2] = int_bits(flags);
        }
    }
    return RB_OK;
}

// split the mixtures into G groups of ~equal row count; all cuts on multiples of 4 mixtures
void make_groups(const rb_gmm* h, int G, std::vector<int>& grpRow, std::vector<int>& grpMix) {
    const int nMix = h->nMix;
    grpRow.assign(1, 0);
    grpMix.assign(1, 0);
    long total = 0;
    for (int m = 0; m < nMix; ++m)
        total += h->rowsOfMixture[m];
    long acc = 0;
    int  g   = 1;
    for (int m = 0; m < nMix; ++m) {
        acc += h->rowsOfMixture[m];
        const bool boundaryOk = ((m + 1) % 4 == 0) && (m + 1 < nMix);
        if (g < G && boundaryOk && acc * G >= total * g) {
            grpRow.push_back((int)acc);
            grpMix.push_back(m + 1);
            ++g;
        }
    }
    grpRow.push_back((int)total);
    grpMix.push_back(nMix);
}

constexpr int kMaxGroups = 64, kGroupStride = kMaxGroups + 2;

int choose_groups(const rb_gmm* h, long T, int slots, int framesPerBlock, double* effOut = nullptr) {
    const long FB   = (T + framesPerBlock - 1) / framesPerBlock;
    const int  gmax = std::max(1, std::min(64, std::min(h->nMix / 4, h->nRows / 128)));
    int        best = 1;
    double     bestEff = -1;
    for (int G = 1; G <= gmax; ++G) {
        const long   items = FB * G;
        const long   waves = (items + slots - 1) / slots;
        const double eff   = (double)items / (double)(waves * slots);
        if (eff > bestEff + 0.02) {  // prefer few groups unless the wave efficiency gain is real
            bestEff = eff;
            best    = G;
        }
    }
    if (effOut)  // wave efficiency x lane fill of the last frame block
        *effOut = bestEff * (double)T / (double)(FB * framesPerBlock);
    return best;
}

int launch_simt(rb_gmm* h, const float* dFeats, long T, float* dScores, uint32_t* dBest, cudaStream_t s) {
    if (h->nVariants > 0) {
        int    pick = 0;
        double bestScore = -1;
        for (int v = 0; v < h->nVariants; ++v) {
            double eff = 0;
            choose_groups(h, T, h->dev.sm_count * h->variants[v].ctasPerSm, h->variants[v].framesPerBlock, &eff);
            if (eff * h->variants[v].rate > bestScore) {
                bestScore = eff * h->variants[v].rate;
                pick      = v;
            }
        }
        if (pick != h->curVariant) {
            h->curVariant     = pick;
            h->kernel         = h->variants[pick].kernel;
            h->framesPerBlock = h->variants[pick].framesPerBlock;
            h->ctasPerSm      = h->variants[pick].ctasPerSm;
            h->curGroups      = -1;
        }
    }
    const int slots = h->dev.sm_count * h->ctasPerSm;
    const int G     = choose_groups(h, T, slots, h->framesPerBlock);
    h->curGroups    = h->groupsOf[G];  // make_groups may return fewer groups than asked for
    GmmParams p;
    p.rows         = h->dRows.p;
    p.grp_row      = h->dGrpRow.p + (size_t)G * kGroupStride;  // tables of every G live on the device (rb_gmm_create)
    p.grp_mix      = h->dGrpMix.p + (size_t)G * kGroupStride;
    p.isd          = h->dIsd.p;
    p.feats        = dFeats;
    p.scores       = dScores;
    p.best         = dBest;
    p.T            = T;
    p.dim          = h->dim;
    p.nMix         = h->nMix;
    p.nGroups      = h->curGroups;
    p.nFrameBlocks = (int)((T + h->framesPerBlock - 1) / h->framesPerBlock);
    p.vec4         = (h->nMix % 4 == 0 && ((uintptr_t)dScores % 16 == 0)) ? 1 : 0;
    const long items = (long)p.nGroups * p.nFrameBlocks;
    const int  grid  = (int)std::min<long>(items, slots);
    h->kernel<<<grid, kThreads, h->smemBytes, s>>>(p);
    RB_LAUNCH_CHECK();
    return RB_OK;
}


// BATCH_FLOAT: can large batches take the screening + refinement route?  Needs what the tensor scorer needs (no empty
// mixture, at most 256 densities per mixture), at most 32 densities per mixture (one bit each), a mixture count that is
// a multiple of 4 (16-byte words), finite parameters, and a mixture grouping whose rows fit shared memory.  When it does
// not apply the direct kernel serves every call (same scores, only slower).  RB_GMM_EXACT=0 switches the route off.
int setup_exact_two_pass(rb_gmm* h, const rb_mixture_set* ms, const float* rowsHost, bool diag) {
    const char* env = getenv("RB_GMM_EXACT");
    if (env && atoi(env) == 0)
        return RB_OK;
    if (const char* e = getenv("RB_GMM_EXACT_MIN_FRAMES"))
        h->exactMinFrames = std::max(1, atoi(e));
    if (env && atoi(env) > 1)
        h->exactMinFrames = 1;  // tests: every call takes the two-pass route
    if (h->nMix % 4 != 0)
        return RB_OK;
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        const uint32_t n = ms->mix_offsets[m + 1] - ms->mix_offsets[m];
        if (n == 0 || n > 32)
            return RB_OK;
    }
    bool pooled = diag && ms->n_covariances == 1 && getenv("RB_GMM_DIAG_ROWS_WITH_ISD") == nullptr;
    for (uint32_t i = 0; i < ms->n_densities && pooled; ++i)
        pooled = ms->dens_cov[i] == 0;
    h->refPooled = pooled;
    h->refine = diag ? refine_diag_kernel_for(h->nUnits, h->fuse, pooled) : refine_kernel_for(h->nUnits, h->fuse);
    if (!h->refine)
        return RB_OK;
    rb_gmm_tensor* t = nullptr;
    const int      made = diag ? rb_gmm_tensor_create_diag(ms, rowsHost, h->rowf, h->nUnits, h->dev, h->stream, &t)
                               : rb_gmm_tensor_create(ms, h->dev, h->stream, &t);
    if (made != RB_OK || !rb_gmm_tensor_screenable(t)) {
        if (t)
            rb_gmm_tensor_destroy(t);
        return RB_OK;  // e.g. non-finite parameters: the direct kernel reproduces the reference on those too
    }
    // fewest groups whose rows fit 48 KB (3-4 CTAs per SM); else up to 200 KB at lower occupancy
    const int        pitch = diag ? (pooled ? refine_diag_pooled_pitch(h->nUnits) : refine_diag_pitch(h->nUnits))
                                  : refine_pitch(h->nUnits);
    const int        used  = diag ? h->nUnits * 8 + 2 : h->nUnits * 8 + 1;  // leading floats of a direct-kernel row
    std::vector<int> mixRow(h->nMix + 1, 0);
    for (int m = 0; m < h->nMix; ++m)
        mixRow[m + 1] = mixRow[m] + h->rowsOfMixture[m];
    int    pickG = 0;
    size_t pickSmem = 0;
    for (size_t budget : {(size_t)72 * 1024, (size_t)110 * 1024, (size_t)200 * 1024}) {
        for (int G = 1; G <= kMaxGroups && !pickG; ++G) {
            std::vector<int> grpRow, grpMix;
            make_groups(h, G, grpRow, grpMix);
            size_t need = 0;
            for (size_t g = 0; g + 1 < grpRow.size(); ++g)
                need = std::max(need, 16 + rb::round_up(sizeof(int) * (size_t)(grpMix[g + 1] - grpMix[g]), 16) +
                                              rb::round_up(sizeof(float) * (size_t)(grpRow[g + 1] - grpRow[g]) * pitch, 16) + 64 +
                                              kStageBytes);
            if (need <= budget) {
                pickG    = G;
                pickSmem = need;
            }
        }
        if (pickG)
            break;
    }
    int occ = 0;
    if (!pickG ||
        cudaFuncSetAttribute(h->refine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pickSmem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->refine, kThreads, pickSmem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        rb_gmm_tensor_destroy(t);
        return RB_OK;
    }
    // the rows again at the refinement pitch (+ one spare row for the 16-byte overhang of the bulk copy)
    std::vector<float> rrows((size_t)(h->nRows + 1) * pitch, 0.0f);
    for (int r = 0; r < h->nRows; ++r) {
        const float* src = rowsHost + (size_t)r * h->rowf;
        if (pooled) {  // [ mean | (w, logNorm) ]; the 1/sqrt(var) segment is the same in every row
            std::copy(src, src + h->nUnits * 4, rrows.begin() + (size_t)r * pitch);
            std::copy(src + h->nUnits * 8, src + h->nUnits * 8 + 2, rrows.begin() + (size_t)r * pitch + h->nUnits * 4);
        }
        else
            std::copy(src, src + used, rrows.begin() + (size_t)r * pitch);
    }
    if (pooled) {
        // (placeholder rows of empty mixtures carry no 1/sqrt(var): take it from the first real density)
        std::vector<float> isdRow(h->nUnits * 4, 0.0f);
        for (uint32_t m = 0, r = 0; m < ms->n_mixtures; r += h->rowsOfMixture[m], ++m)
            if (ms->mix_offsets[m + 1] > ms->mix_offsets[m]) {
                std::copy(rowsHost + (size_t)r * h->rowf + h->nUnits * 4, rowsHost + (size_t)r * h->rowf + h->nUnits * 8, isdRow.begin());
                break;
            }
        if (h->dRefIsd.upload(isdRow, h->stream) != RB_OK) {
            rb_gmm_tensor_destroy(t);
            return RB_ERR_CUDA;
        }
    }
    if (h->dRefRows.upload(rrows, h->stream) != RB_OK || h->dMixRow.upload(mixRow, h->stream) != RB_OK ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
        rb_gmm_tensor_destroy(t);
        return RB_ERR_CUDA;
    }
    h->tensor    = t;
    h->refGroups = pickG;
    h->refSmem   = pickSmem;
    h->refSlots  = h->dev.sm_count * occ;
    return RB_OK;
}

int launch_exact_two_pass(rb_gmm* h, const float* dFeats, long T, float* dScores, float* const* extra, int nExtra,
                          uint32_t* dBest, cudaStream_t s) {
    const long chunk = rb_gmm_tensor_chunk(h->tensor);
    for (long a = 0; a < T; a += chunk) {
        const long   n = std::min(chunk, T - a);
        RefineParams p;
        const bool   timing = h->tev[0] != nullptr && a == 0;  // the first chunk of the call is the one that is timed
        if (timing)
            cudaEventRecord(h->tev[0], s);
        RB_CHECK(rb_gmm_tensor_screen(h->tensor, dFeats + (size_t)a * h->dim, n, &p.words, &p.xT, &p.pitch, s,
                                      timing ? h->tev[1] : nullptr));
        if (timing)
            cudaEventRecord(h->tev[2], s);
        p.scores       = dScores + (size_t)a * h->nMix;
        p.nExtra       = nExtra;
        for (int e = 0; e < nExtra; ++e)
            p.extra[e] = extra[e] + (size_t)a * h->nMix;
        p.best         = dBest ? dBest + (size_t)a * h->nMix : nullptr;
        p.dim          = h->dim;
        const int G    = h->refGroups;
        p.rows         = h->dRefRows.p;
        p.grp_row      = h->dGrpRow.p + (size_t)G * kGroupStride;
        p.grp_mix      = h->dGrpMix.p + (size_t)G * kGroupStride;
        p.mix_row      = h->dMixRow.p;
        p.T            = n;
        p.nMix         = h->nMix;
        p.nGroups      = h->groupsOf[G];
        p.nFrameBlocks = (int)((n + kThreads - 1) / kThreads);
        p.emptyScore   = FLT_MAX;
        p.pooledIsd    = h->refPooled ? h->dRefIsd.p : nullptr;
        const long items = (long)p.nGroups * p.nFrameBlocks;
        const int  grid  = (int)std::max<long>(p.nGroups, std::min<long>(items, h->refSlots));
        h->refine<<<grid, kThreads, h->refSmem, s>>>(p);
        RB_LAUNCH_CHECK();
        if (timing) {
            cudaEventRecord(h->tev[3], s);
            h->timed = true;
        }
    }
    return RB_OK;
}

// ---- density preselection through the refinement kernel ------------------------------------------------------------
// "preselection-batch-float" scores, per frame, only the densities of the selected clusters with the arithmetic of
// the batch-float scorer and gives a mixture without a scored density the back-off score
// (src/Mm/BatchFeatureScorer.cc:286-315).  That is the refinement kernel's job description with a different source of
// candidate sets: bit j of a (frame, mixture) word = "the cluster of the mixture's j-th density is active".
constexpr int kMaskFrames = 64;  // frames per CTA of presel_masks_kernel
size_t presel_masks_smem(const rb_gmm* h) {  // cluster bytes + mixture offsets + the active bits of 64 frames
    return (size_t)((h->nRows + 15) & ~15) + sizeof(uint32_t) * (((size_t)h->nMix + 1 + 3) & ~(size_t)3) +
           sizeof(uint32_t) * kMaskFrames * 9;
}
__global__ void __launch_bounds__(256) presel_masks_kernel(const uint32_t* __restrict__ active, const uint8_t* __restrict__ clusterOf,
                                                           const uint32_t* __restrict__ offsets, uint32_t* __restrict__ words,
                                                           long pitch, long T, int nMix, int nDens) {
    // shared: the cluster of every density (a byte each), the first density of every mixture, and the active-cluster
    // bits of this CTA's 64 frames with a row stride of 9 words (bank = frame + word: conflict free for lanes = frames)
    extern __shared__ __align__(16) unsigned char smemMask[];
    uint8_t*  sCluster = smemMask;
    uint32_t* sOff     = reinterpret_cast<uint32_t*>(smemMask + ((nDens + 15) & ~15));
    uint32_t* sActive  = sOff + ((nMix + 1 + 3) & ~3);
    for (int i = threadIdx.x; i < nDens; i += 256)
        sCluster[i] = clusterOf[i];
    for (int i = threadIdx.x; i <= nMix; i += 256)
        sOff[i] = offsets[i];
    const int nQuads = nMix >> 2;
    for (long t0 = (long)blockIdx.x * kMaskFrames; t0 < T; t0 += (long)gridDim.x * kMaskFrames) {
        __syncthreads();
        for (int i = threadIdx.x; i < kMaskFrames * 8; i += 256) {
            const long t = t0 + (i >> 3);
            sActive[(i >> 3) * 9 + (i & 7)] = t < T ? active[t * 8 + (i & 7)] : 0u;
        }
        __syncthreads();
        const int  f = threadIdx.x & (kMaskFrames - 1);  // lanes of a warp = consecutive frames: coalesced 16-byte stores
        const long t = t0 + f;
        const uint32_t* sel = sActive + f * 9;
        for (int q = threadIdx.x / kMaskFrames; q < nQuads; q += 256 / kMaskFrames) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t a = sOff[4 * q + k], b = sOff[4 * q + k + 1];
                uint32_t       m = 0;
                for (uint32_t e = a; e < b; ++e) {
                    const uint32_t c = sCluster[e];  // the same density for the whole warp: a broadcast
                    m |= ((sel[c >> 5] >> (c & 31u)) & 1u) << (e - a);
                }
                w[k] = m;
            }
            if (t < T)
                reinterpret_cast<uint4*>(words)[(size_t)q * pitch + t] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

int launch_presel_two_pass(rb_gmm* h, const float* dFeats, long T, float* dScores, float* const* extra, int nExtra,
                           cudaStream_t s) {
    const long chunk = rb_gmm_tensor_chunk(h->tensor);
    for (long a = 0; a < T; a += chunk) {
        const long      n = std::min(chunk, T - a);
        const uint32_t* active = nullptr;
        const uint8_t*  clusterOf = nullptr;
        const uint32_t* offsets = nullptr;
        uint32_t*       words = nullptr;
        RefineParams    p;
        RB_CHECK(rb_gmm_presel_select(h->presel, dFeats + (size_t)a * h->dim, n, &active, &clusterOf, &offsets, &p.emptyScore, s));
        RB_CHECK(rb_gmm_tensor_split(h->tensor, dFeats + (size_t)a * h->dim, n, &words, &p.xT, &p.pitch, s));
        const int    nDens    = h->nRows;  // (every mixture has >= 1 density on this route: rows = densities)
        const size_t smemMask = presel_masks_smem(h);
        RB_CUDA(cudaFuncSetAttribute(presel_masks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemMask));
        presel_masks_kernel<<<(int)std::min<long>((n + kMaskFrames - 1) / kMaskFrames, (long)h->dev.sm_count * 8), 256, smemMask, s>>>(
                active, clusterOf, offsets, words, p.pitch, n, h->nMix, nDens);
        RB_LAUNCH_CHECK();
        p.words        = words;
        p.scores       = dScores + (size_t)a * h->nMix;
        p.nExtra       = nExtra;
        for (int e = 0; e < nExtra; ++e)
            p.extra[e] = extra[e] + (size_t)a * h->nMix;
        p.best         = nullptr;
        p.pooledIsd    = nullptr;
        p.dim          = h->dim;
        const int G    = h->refGroups;
        p.rows         = h->dRefRows.p;
        p.grp_row      = h->dGrpRow.p + (size_t)G * kGroupStride;
        p.grp_mix      = h->dGrpMix.p + (size_t)G * kGroupStride;
        p.mix_row      = h->dMixRow.p;
        p.T            = n;
        p.nMix         = h->nMix;
        p.nGroups      = h->groupsOf[G];
        p.nFrameBlocks = (int)((n + kThreads - 1) / kThreads);
        const long items = (long)p.nGroups * p.nFrameBlocks;
        const int  grid  = (int)std::max<long>(p.nGroups, std::min<long>(items, h->refSlots));
        h->refine<<<grid, kThreads, h->refSmem, s>>>(p);
        RB_LAUNCH_CHECK();
    }
    return RB_OK;
}

}  // namespace

extern "C" int rb_gmm_create(const rb_mixture_set* ms, int mode, float mixture_weight_scale, float gaussian_scale,
                             int contraction, int device, rb_gmm** out) {
    RB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    RB_CHECK(validate(ms));
    RB_REQUIRE(mode >= RB_GMM_BATCH_FLOAT && mode <= RB_GMM_SIMD_DIAG_MAX, "unknown gmm mode %d", mode);
    rb_gmm* h = new (std::nothrow) rb_gmm();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    int rc = rb::use_device(device, &h->dev);
    if (rc != RB_OK) {
        delete h;
        return rc;
    }
    auto fail = [&](int code) {
        delete h;
        return code;
    };
    h->mode = mode;
    h->fuse = contraction != 0;
    h->dim  = (int)ms->dim;
    h->nMix = (int)ms->n_mixtures;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rb::set_error("cudaStreamCreate failed");
        return fail(RB_ERR_CUDA);
    }
    std::vector<float> rows, isd;
    if (mode == RB_GMM_BATCH_INT) {
        rc = rb_gmm_int_create(ms, h->dev, h->stream, &h->quantised);
        if (rc != RB_OK)
            return fail(rc);
        *out = h;
        return RB_OK;
    }
    if (mode == RB_GMM_SIMD_DIAG_MAX) {
        if (mixture_weight_scale != 1.0f || gaussian_scale != 1.0f) {  // the reference's SIMD scorer has no such parameters
            rb::set_error("RB_GMM_SIMD_DIAG_MAX takes no mixture-weight / gaussian scale (got %g, %g)", mixture_weight_scale,
                          gaussian_scale);
            return fail(RB_ERR_INVALID);
        }
        rc = rb_gmm_simd_create(ms, h->dev, h->stream, &h->simd);
        if (rc != RB_OK)
            return fail(rc);
        *out = h;
        return RB_OK;
    }
    if (mode == RB_GMM_BATCH_PRESELECT) {
        rc = rb_gmm_presel_create(ms, h->fuse, h->dev, h->stream, &h->presel);
        if (rc != RB_OK)
            return fail(rc);
        // models the refinement kernel covers (<= 32 densities per mixture, dimension <= 64 ...) are scored through it
        // (launch_presel_two_pass); the batch-float tables below are set up for that.  Others keep presel_score_kernel.
        if (ms->dim > 64 || ms->n_covariances != 1 || getenv("RB_GMM_PRESEL_DIRECT") != nullptr) {
            *out = h;
            return RB_OK;
        }
    }
    if (mode == RB_GMM_BATCH_PRESELECT_INT) {
        rc = rb_gmm_presel_int_create(ms, h->dev, h->stream, &h->preselInt);
        if (rc != RB_OK)
            return fail(rc);
        *out = h;
        return RB_OK;
    }
    if (mode == RB_GMM_BATCH_FLOAT || mode == RB_GMM_BATCH_TENSOR || mode == RB_GMM_BATCH_PRESELECT) {
        if (ms->dim > 64) {
            rb::set_error("batch scorer supports feature dimension <= 64 (got %u)", ms->dim);
            return fail(RB_ERR_UNSUPPORTED);
        }
        rc = build_batch_rows(h, ms, rows, isd);
        if (rc != RB_OK)
            return fail(rc);
        // more frames per thread = fewer LDS.128 of the density rows per FFMA2 (each costs ~3 FMA-pipe cycles of
        // register-file bandwidth, scripts/micro/fp32_rate.cu) but larger frame blocks; 4 needs dim <= 40
        static const int    fpts[3]  = {1, 2, 4};
        static const double rates[3] = {0.947, 0.936, 1.0};
        const char*         force    = getenv("RB_GMM_FPT");
        for (int i = 0; i < 3; ++i) {
            if (force && atoi(force) != fpts[i])
                continue;
            GmmKernel k = (fpts[i] == 4 && h->nUnits > 5) ? nullptr : batch_kernel_for(h->nUnits, h->fuse, fpts[i]);
            if (!k)
                continue;
            rb_gmm::Variant& v = h->variants[h->nVariants++];
            v.kernel           = k;
            v.framesPerBlock   = kThreads * fpts[i];
            v.rate             = rates[i];
        }
        h->kernel = h->nVariants ? h->variants[0].kernel : nullptr;
    }
    else {
        if (ms->dim > 64) {
            rb::set_error("diagonal scorer supports feature dimension <= 64 (got %u)", ms->dim);
            return fail(RB_ERR_UNSUPPORTED);
        }
        rc = build_diag_rows(h, ms, mixture_weight_scale, gaussian_scale, rows);
        if (rc != RB_OK)
            return fail(rc);
        isd.assign(4, 0.0f);
        h->kernel = diag_kernel_for(h->nUnits, h->fuse, mode == RB_GMM_DIAG_SUM);
    }
    if (!h->kernel) {
        rb::set_error("no kernel instance for dimension %d", h->dim);
        return fail(RB_ERR_UNSUPPORTED);
    }
    h->nRows     = (int)(rows.size() / h->rowf);
    h->smemBytes = sizeof(float) * kStages * kChunkRows * h->rowf + sizeof(uint64_t) * kStages;
    if (cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBytes) !=
        cudaSuccess) {
        rb::set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", h->smemBytes,
                      cudaGetErrorString(cudaGetLastError()));
        return fail(RB_ERR_CUDA);
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->kernel, kThreads, h->smemBytes) != cudaSuccess ||
        occ < 1) {
        rb::set_error("gmm kernel does not fit on the device: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(RB_ERR_CUDA);
    }
    h->ctasPerSm = occ;
    for (int v = 0; v < h->nVariants; ++v) {
        if (cudaFuncSetAttribute(h->variants[v].kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)h->smemBytes) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->variants[v].kernel, kThreads, h->smemBytes) !=
                    cudaSuccess ||
            occ < 1) {
            rb::set_error("gmm kernel variant does not fit on the device: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(RB_ERR_CUDA);
        }
        h->variants[v].ctasPerSm = occ;
    }
    {  // group tables for every group count, so that a launch never has to touch them
        std::vector<int> allRow((size_t)(kMaxGroups + 1) * kGroupStride, 0), allMix(allRow.size(), 0);
        h->groupsOf.assign(kMaxGroups + 1, 1);
        for (int G = 1; G <= kMaxGroups; ++G) {
            std::vector<int> grpRow, grpMix;
            make_groups(h, G, grpRow, grpMix);
            std::copy(grpRow.begin(), grpRow.end(), allRow.begin() + (size_t)G * kGroupStride);
            std::copy(grpMix.begin(), grpMix.end(), allMix.begin() + (size_t)G * kGroupStride);
            h->groupsOf[G] = (int)grpRow.size() - 1;
        }
        if (h->dGrpRow.upload(allRow, h->stream) != RB_OK || h->dGrpMix.upload(allMix, h->stream) != RB_OK)
            return fail(RB_ERR_CUDA);
    }
    if (h->dRows.upload(rows, h->stream) != RB_OK || h->dIsd.upload(isd, h->stream) != RB_OK)
        return fail(RB_ERR_CUDA);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
        rb::set_error("model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(RB_ERR_CUDA);
    }
    if (mode == RB_GMM_BATCH_TENSOR) {
        rc = rb_gmm_tensor_create(ms, h->dev, h->stream, &h->tensor);
        if (rc != RB_OK)
            return fail(rc);
    }
    if (mode == RB_GMM_DIAG_MAX) {
        rc = setup_exact_two_pass(h, ms, rows.data(), true);
        if (rc != RB_OK)
            return fail(rc);
    }
    if (mode == RB_GMM_BATCH_FLOAT || mode == RB_GMM_BATCH_PRESELECT) {
        rc = setup_exact_two_pass(h, ms, rows.data(), false);
        if (rc != RB_OK)
            return fail(rc);
    }
    *out = h;
    return RB_OK;
}

extern "C" void rb_gmm_destroy(rb_gmm* h) {
    if (!h)
        return;
    cudaSetDevice(h->dev.ordinal);
    delete h;
}

extern "C" int rb_gmm_n_mixtures(const rb_gmm* h) {
    return h ? h->nMix : 0;
}

extern "C" int rb_gmm_dim(const rb_gmm* h) {
    return h ? h->dim : 0;
}

namespace {
// scores into d_scores and, row for row, into the nExtra further destinations (windows of peer GPUs): the exact two-pass
// route stores them from the refinement kernel itself, every other route copies the finished matrix
int score_dev_impl(rb_gmm* h, const float* d_feats, long T, float* d_scores, float* const* extra, int nExtra,
                   uint32_t* d_best, cudaStream_t s) {
    if ((h->mode == RB_GMM_BATCH_FLOAT || h->mode == RB_GMM_DIAG_MAX) && h->refGroups > 0 && T >= h->exactMinFrames &&
        ((uintptr_t)d_scores % 16) == 0 && ((uintptr_t)d_best % 16) == 0) {
        bool aligned = true;
        for (int e = 0; e < nExtra; ++e)
            aligned = aligned && ((uintptr_t)extra[e] % 16) == 0;
        if (aligned)
            return launch_exact_two_pass(h, d_feats, T, d_scores, extra, nExtra, d_best, s);
    }
    if (h->mode == RB_GMM_BATCH_PRESELECT && h->refGroups > 0 && h->tensor && ((uintptr_t)d_scores % 16) == 0 &&
        presel_masks_smem(h) + 1024 <= h->dev.smem_optin) {
        bool aligned = true;
        for (int e = 0; e < nExtra; ++e)
            aligned = aligned && ((uintptr_t)extra[e] % 16) == 0;
        if (aligned)
            return launch_presel_two_pass(h, d_feats, T, d_scores, extra, nExtra, s);
    }
    int rc;
    if (h->mode == RB_GMM_BATCH_TENSOR)
        rc = rb_gmm_tensor_score(h->tensor, d_feats, T, d_scores, s);
    else if (h->mode == RB_GMM_BATCH_INT)
        rc = rb_gmm_int_score(h->quantised, d_feats, T, d_scores, s);
    else if (h->mode == RB_GMM_BATCH_PRESELECT)
        rc = rb_gmm_presel_score(h->presel, d_feats, T, d_scores, s);
    else if (h->mode == RB_GMM_BATCH_PRESELECT_INT)
        rc = rb_gmm_presel_int_score(h->preselInt, d_feats, T, d_scores, s);
    else if (h->mode == RB_GMM_SIMD_DIAG_MAX)
        rc = rb_gmm_simd_score(h->simd, d_feats, T, d_scores, d_best, s);
    else
        rc = launch_simt(h, d_feats, T, d_scores, d_best, s);
    RB_CHECK(rc);
    for (int e = 0; e < nExtra; ++e)
        RB_CUDA(cudaMemcpyAsync(extra[e], d_scores, (size_t)T * h->nMix * sizeof(float), cudaMemcpyDefault, s));
    return RB_OK;
}
}  // namespace

extern "C" int rb_gmm_score_dev(rb_gmm* h, const float* d_feats, long T, float* d_scores, uint32_t* d_best,
                                void* stream) {
    RB_REQUIRE(h != nullptr, "gmm handle is NULL");
    RB_REQUIRE(T >= 0, "negative frame count");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(d_feats && d_scores, "NULL device buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    if (h->mode != RB_GMM_DIAG_MAX && h->mode != RB_GMM_DIAG_SUM && h->mode != RB_GMM_SIMD_DIAG_MAX)
        RB_REQUIRE(d_best == nullptr, "this scorer mode does not report densities; use RB_GMM_DIAG_MAX");
    return score_dev_impl(h, d_feats, T, d_scores, nullptr, 0, d_best, s);
}

extern "C" int rb_gmm_score_fanout_dev(rb_gmm* h, const float* d_feats, long T, int n_dst, float* const* d_dst,
                                       void* stream) {
    RB_REQUIRE(h != nullptr, "gmm handle is NULL");
    RB_REQUIRE(T >= 0, "negative frame count");
    RB_REQUIRE(n_dst >= 1 && n_dst <= 16 && d_dst, "between 1 and 16 destinations");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(d_feats != nullptr, "NULL device buffer");
    for (int e = 0; e < n_dst; ++e)
        RB_REQUIRE(d_dst[e] != nullptr, "destination %d is NULL", e);
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    return score_dev_impl(h, d_feats, T, d_dst[0], d_dst + 1, n_dst - 1, nullptr, s);
}

int rb_gmm_reserve(rb_gmm* h, long frames) {
    if (!h || !h->tensor || frames <= 0)
        return RB_OK;
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    return rb_gmm_tensor_reserve(h->tensor, frames, h->mode == RB_GMM_BATCH_FLOAT || h->mode == RB_GMM_DIAG_MAX ||
                                                            h->mode == RB_GMM_BATCH_PRESELECT);
}

// Host-pointer entry point: frames are cut into slabs; H2D of slab i+1, scoring of slab i and D2H of
// slab i-1 overlap on three streams (PCIe is the end-to-end bound: 1 KB of scores per frame).
extern "C" int rb_gmm_score(rb_gmm* h, const float* feats, long T, float* scores, uint32_t* best_density) {
    RB_REQUIRE(h != nullptr, "gmm handle is NULL");
    RB_REQUIRE(T >= 0, "negative frame count");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(feats && scores, "NULL host buffer");
    if (h->mode == RB_GMM_BATCH_FLOAT || h->mode == RB_GMM_BATCH_TENSOR || h->mode == RB_GMM_BATCH_INT ||
        h->mode == RB_GMM_BATCH_PRESELECT || h->mode == RB_GMM_BATCH_PRESELECT_INT)
        RB_REQUIRE(best_density == nullptr, "this scorer mode does not report densities; use RB_GMM_DIAG_MAX");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    const size_t D = h->dim, M = h->nMix;
    RB_CHECK(h->dFeats.reserve((size_t)T * D));
    RB_CHECK(h->dScores.reserve((size_t)T * M));
    if (best_density)
        RB_CHECK(h->dBest.reserve((size_t)T * M));

    // slabs: H2D of slab i+1, scoring of slab i and D2H of slab i-1 overlap.  The call is bound by the D2H stream
    // (1 KB of scores per frame), so the first slabs are small (the D2H stream starts early) and grow geometrically
    // to 16384 frames (large copies run closer to the PCIe rate)
    std::vector<long> cut(1, 0);
    for (long n = 2048; cut.back() < T; n = std::min<long>(2 * n, 16384))
        cut.push_back(std::min(T, cut.back() + n));
    const int nSlabs = (int)cut.size() - 1;
    if (!h->sIn) {
        RB_CUDA(cudaStreamCreateWithFlags(&h->sIn, cudaStreamNonBlocking));
        RB_CUDA(cudaStreamCreateWithFlags(&h->sOut, cudaStreamNonBlocking));
    }
    while ((int)h->events.size() < 2 * nSlabs) {
        cudaEvent_t e;
        RB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->events.push_back(e);
    }
    cudaStream_t sIn = h->sIn, sOut = h->sOut;
    int          rc = RB_OK;
    // pageable score buffer: through the page-locked staging ring + worker threads (common.cuh HostStager)
    const bool staged = !rb_host_is_pinned(scores) && getenv("RB_NO_HOST_STAGER") == nullptr;
    if (staged)
        RB_CHECK(h->stager.ensure((size_t)4 << 20, 8, h->dev.ordinal));
    const bool stagedIn = !rb_host_is_pinned(feats) && getenv("RB_NO_HOST_STAGER") == nullptr && (size_t)T * D * 4 >= ((size_t)4 << 20);
    if (stagedIn)
        RB_CHECK(h->hFeats.reserve((size_t)T * D));
    {  // the largest slab's scratch once, before the pipeline starts
        long largest = 0;
        for (int i = 0; i < nSlabs; ++i)
            largest = std::max(largest, cut[i + 1] - cut[i]);
        RB_CHECK(rb_gmm_reserve(h, largest));
    }
    for (int i = 0; i < nSlabs && rc == RB_OK; ++i) {
        const long  a = cut[i], n = cut[i + 1] - cut[i];
        cudaEvent_t evIn = h->events[2 * i], evK = h->events[2 * i + 1];
        const float* src = feats + a * D;
        if (stagedIn) {  // (the staging area holds the whole call: no slot has to be waited for)
            rb::parallel_memcpy(h->hFeats.p + a * D, src, (size_t)n * D * 4);
            src = h->hFeats.p + a * D;
        }
        if (cudaMemcpyAsync(h->dFeats.p + a * D, src, (size_t)n * D * 4, cudaMemcpyHostToDevice, sIn) != cudaSuccess)
            rc = RB_ERR_CUDA;
        cudaEventRecord(evIn, sIn);
        cudaStreamWaitEvent(h->stream, evIn, 0);
        if (rc == RB_OK)
            rc = rb_gmm_score_dev(h, h->dFeats.p + a * D, n, h->dScores.p + a * M,
                                  best_density ? h->dBest.p + a * M : nullptr, h->stream);
        cudaEventRecord(evK, h->stream);
        cudaStreamWaitEvent(sOut, evK, 0);
        if (rc == RB_OK && staged)
            rc = h->stager.d2h(scores + a * M, h->dScores.p + a * M, (size_t)n * M * 4, sOut);
        else if (rc == RB_OK &&
                 cudaMemcpyAsync(scores + a * M, h->dScores.p + a * M, (size_t)n * M * 4, cudaMemcpyDeviceToHost, sOut) !=
                         cudaSuccess)
            rc = RB_ERR_CUDA;
        if (rc == RB_OK && best_density &&
            cudaMemcpyAsync(best_density + a * M, h->dBest.p + a * M, (size_t)n * M * 4, cudaMemcpyDeviceToHost,
                            sOut) != cudaSuccess)
            rc = RB_ERR_CUDA;
    }
    cudaError_t e1 = cudaStreamSynchronize(sIn);
    cudaError_t e2 = cudaStreamSynchronize(h->stream);
    cudaError_t e3 = cudaStreamSynchronize(sOut);
    if (staged) {  // the workers' copies into the caller's buffer
        const int rs = h->stager.drain();
        if (rc == RB_OK)
            rc = rs;
    }
    if (rc == RB_ERR_CUDA && rb::get_error()[0] == 0)
        rb::set_error("asynchronous copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != RB_OK)
        return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
        rb::set_error("gmm scoring failed on the device: %s", cudaGetErrorString(e));
        return RB_ERR_CUDA;
    }
    return RB_OK;
}

// Measurement hook for bench.py's roofline: events around the three kernels of the exact two-pass route (first chunk of
// a call).  rb_gmm_get_timing waits for the last timed call and returns its split / screen / refine durations in ms;
// RB_ERR_STATE if the last calls took the direct kernel.
extern "C" int rb_gmm_set_timing(rb_gmm* h, int on) {
    RB_REQUIRE(h != nullptr, "gmm handle is NULL");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    for (cudaEvent_t& e : h->tev) {
        if (on && !e)
            RB_CUDA(cudaEventCreate(&e));
        if (!on && e) {
            cudaEventDestroy(e);
            e = nullptr;
        }
    }
    h->timed = false;
    return RB_OK;
}

extern "C" int rb_gmm_get_timing(rb_gmm* h, float* ms3) {
    RB_REQUIRE(h && ms3, "NULL argument");
    if (!h->tev[0] || !h->timed) {
        rb::set_error("no timed two-pass call on this handle");
        return RB_ERR_STATE;
    }
    RB_CUDA(cudaEventSynchronize(h->tev[3]));
    for (int i = 0; i < 3; ++i)
        RB_CUDA(cudaEventElapsedTime(ms3 + i, h->tev[i], h->tev[i + 1]));
    return RB_OK;
}

// density preselection (RB_GMM_BATCH_PRESELECT): re-cluster with other parameters / read the clustering back
extern "C" int rb_gmm_configure_preselection(rb_gmm* h, int clusters, int select, int iterations, float backoff_score) {
    RB_REQUIRE(h && (h->presel || h->preselInt), "not a preselection scorer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    if (h->preselInt)  // the int variant has no back-off score
        return rb_gmm_presel_int_configure(h->preselInt, clusters, select, iterations, h->stream);
    return rb_gmm_presel_configure(h->presel, clusters, select, iterations, backoff_score, h->stream);
}

extern "C" int rb_gmm_get_clustering(const rb_gmm* h, uint32_t* cluster_of_density, float* cluster_means, int* n_clusters) {
    RB_REQUIRE(h && (h->presel || h->preselInt), "not a preselection scorer");
    if (h->preselInt) {
        rb_gmm_presel_int_clustering(h->preselInt, cluster_of_density, cluster_means, n_clusters);
        return RB_OK;
    }
    rb_gmm_presel_clustering(h->presel, cluster_of_density, cluster_means, n_clusters);
    return RB_OK;
}
