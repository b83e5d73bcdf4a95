// gmm_int.cu -- RB_GMM_BATCH_INT: Mm::BatchIntFeatureScorer ("batch-diagonal-maximum-int", and its unrolled twin
// "batch-diagonal-maximum-fast") for ALL mixtures and ALL frames at once, bit-identical to the CPU path.
//
//   quantize / quantizationScale / init / setFeature   src/Mm/Utilities.hh:190-202, src/Mm/BatchFeatureScorer.cc:355-424
//   addDistance / horizontalAdd / fillScoreCacheTpl     src/Mm/BatchFeatureScorer.cc:427-510
//
// The reference scores a frame against a density as  sum_d (m_d - x_d)^2 + c  over u8-quantised means and
// features in s32 and keeps the minimum over the densities of a mixture; the result is (f32)best / scale_.
// Integer arithmetic is exact, so any evaluation order gives the reference's bits.  Here
//     sum_d (m_d - x_d)^2 + c  =  |x|^2 + (c + |m|^2) - 2 x.m
// and the u8 x u8 -> s32 inner products run on the tensor cores through the warp-level IMMA path
// (mma.sync.m16n8k32.s32.u8.u8.s32) with the accumulators in REGISTERS: the min over a mixture, the |x|^2 term,
// the int -> float conversion and the IEEE division are applied there and only one f32 per (frame, mixture)
// leaves the SM.  Why not tcgen05: its accumulators live in TMEM and must come back through tcgen05.ld at
// 64 B/clk/SM (B300_MICROARCH.md; the fp16 tensor scorer in gmm_tensor.cu sits exactly on that bound, 98 us per
// 100k frames), i.e. >= 88 us for 100k x 4096 accumulators, while the IMMA path needs 44-60 us for the same products
// (scripts/micro/mma_rate.cu: 2036 u8 MAC/clk/SM peak, 0.37-0.39 IMMA/clk/SM in this kernel's operand pattern) and
// has no read-back at all.  Measured: 125 us per 100k frames (tensor pipe 38 % busy, the rest is the per-mixture
// epilogue and its latency).  Tried and rejected: folding the factor 2 into the products ([x|x].[m|m], 12 IMMAs
// per tile, no subtract: 150 us) and 3 CTAs per SM (register spills: 236 us).
//
// Layout: densities in mixture order, every mixture padded to whole 8-column tiles (dummy columns can never win
// the min).  A tile is one 688-byte block [8 rows x 80 B (64 B of means + 16 B pad: conflict-free ldmatrix) |
// c + |m|^2 of the 8 columns | flags], streamed global -> shared by the TMA unit (cp.async.bulk) through a
// 3-stage mbarrier ring.  Work item = 512 frames x one group of mixtures, persistent grid of 148 x 2 CTAs; a
// warp owns 64 frames whose quantised features stay in registers as IMMA A fragments.
#include <climits>
#include <cmath>

#include "common.cuh"

namespace {

using namespace rbdev;

constexpr int kThreads     = 256;
constexpr int kWarpFrames  = 64;                             // 4 m16 tiles
constexpr int kBlockFrames = (kThreads / 32) * kWarpFrames;  // 512
constexpr int kKBytes      = 64;                             // quantised dims per row (2 k-steps of 32)
constexpr int kRowBytes    = 80;
constexpr int kTileBytes   = 8 * kRowBytes + 32 + 16;        // 688
constexpr int kChunkTiles  = 32;
constexpr int kStages      = 3;
constexpr int kDummy       = 0x3fffffff;  // c + |m|^2 of a padding column: larger than any real score

struct IntParams {
    const unsigned char* tiles;    // [nTiles * kTileBytes]
    const int*           grpTile;  // [G+1] tile boundaries of the mixture groups
    const int*           grpMix;   // [G+1] mixture boundaries
    const unsigned char* xq;       // [T * 64] quantised features
    const int*           xsq;      // [T] |x|^2
    float*               scores;   // [T * nMix]
    long                 T;
    int                  nMix, nGroups, nFrameBlocks, vec4;
    float                scale;     // scale_ = 2 * quantisation scale^2 (simd: the quantisation scale^2)
    float                rcpScale;  // RN(1 / scale_); 0 if the three-instruction division is not proven for this scale
    int                  simd;      // scores as Mm::SimdGaussDiagonalMaximumFeatureScorer forms them (gmm_simd.cu)
};

// setFeature: u8 = clip(round(f * isd * scale) + 128); 16 lanes per frame, 4 dims (one u32) per lane
__global__ void __launch_bounds__(256) gmm_int_quantize_kernel(const float* __restrict__ feats,
                                                               const float* __restrict__ variance, long T, int dim,
                                                               unsigned char* __restrict__ xq, int* __restrict__ xsq) {
    const int  sub = threadIdx.x & 15;
    const long g0  = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 4;
    const long nG  = ((long)gridDim.x * blockDim.x) >> 4;
    for (long t0 = g0; t0 < ((T + 1) & ~1L); t0 += nG) {  // whole warps stay converged for the shuffles
        const long t = t0 < T ? t0 : T - 1;
        uint32_t   packed = 0;
        int        sq = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = sub * 4 + j;
            int       q = 0;  // padding dims are 0 in features and means (memset / zero-initialised)
            if (d < dim) {
                const float r = roundf(__fmul_rn(__ldg(feats + (size_t)t * dim + d), __ldg(variance + d)));
                // (int) of an out-of-range or NaN float is INT_MIN on the reference's x86 (cvttss2si)
                const int   i = fabsf(r) < 2147483648.0f ? __float2int_rz(r) : INT_MIN;
                q             = min(max(i + (i < INT_MAX - 128 ? 128 : 0), 0), 255);
            }
            packed |= (uint32_t)q << (8 * j);
            sq += q * q;
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1)
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (t0 < T) {
            reinterpret_cast<uint32_t*>(xq + (size_t)t * kKBytes)[sub] = packed;
            if (sub == 0)
                xsq[t] = sq;
        }
    }
}

__device__ __forceinline__ void imma_u8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1,
                                        const int (&c)[4]) {
    asm volatile(
            "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
            : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
}

// (f32)b / scale_, correctly rounded.  For |b| < 2^24 the quotient is formed as q0 = a r, q = fma(fma(-s, q0, a), r, q0)
// with r = RN(1/s) (Markstein); the host has checked this sequence against IEEE division for EVERY integer in that
// range for this particular scale (rb_gmm_int_create), otherwise rcpScale is 0 and div.rn is used throughout.
__device__ __forceinline__ float score_of(int b, float scale, float rcp, int simd) {
    if (simd)  // result.score = 0.5 * quantizedResult.first / scalingSquared_ in f64 (src/Mm/SimdFeatureScorer.cc:143)
        return (float)(0.5 * (double)b / (double)scale);
    const float a = (float)b;
    if (rcp != 0.0f && (unsigned)(b + (1 << 24)) < (2u << 24)) {
        const float q0 = __fmul_rn(a, rcp);
        return __fmaf_rn(__fmaf_rn(-scale, q0, a), rcp, q0);
    }
    return __fdiv_rn(a, scale);
}

__device__ __forceinline__ int min3(int a, int b, int c) {
    return __vimin3_s32(a, b, c);
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&b)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
                 : "r"(addr));
}
__device__ __forceinline__ int2 lds_int2(uint32_t addr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_int(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__global__ void __launch_bounds__(kThreads, 2) gmm_int_kernel(const IntParams p) {
    constexpr int CHUNK = kChunkTiles * kTileBytes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* buf  = smem_raw;
    const uint32_t sbuf = smem_u32(smem_raw);
    uint64_t*      bar  = reinterpret_cast<uint64_t*>(smem_raw + kStages * CHUNK);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s)
            mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // ldmatrix.x4 row addresses: lane i -> row (i & 7) of the 16-byte k-chunk (i >> 3) of the tile
    const uint32_t ldsmOff = (uint32_t)((lane & 7) * kRowBytes + (lane >> 3) * 16);
    const uint32_t ccOff   = (uint32_t)(8 * kRowBytes + q * 8);  // c + |m|^2 of columns 2q, 2q+1
    // after the quad butterfly this lane owns the two rows  mtOwn * 16 + {0, 8} + g  of the warp's 64 frames
    const int mtOwn = 2 * (q & 1) + (q >> 1);

    uint32_t  seq    = 0;
    const int nItems = p.nGroups * p.nFrameBlocks;
    const int zero[4] = {0, 0, 0, 0};

    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int  grp   = item / p.nFrameBlocks;
        const int  fb    = item - grp * p.nFrameBlocks;
        const int  tile0 = p.grpTile[grp], tile1 = p.grpTile[grp + 1];
        const int  nCh   = (tile1 - tile0 + kChunkTiles - 1) / kChunkTiles;
        const long f0    = (long)fb * kBlockFrames + warp * kWarpFrames;

        if (tid == 0) {
            for (int c = 0; c < kStages - 1 && c < nCh; ++c) {
                const int      t0    = tile0 + c * kChunkTiles;
                const uint32_t bytes = (uint32_t)min(kChunkTiles, tile1 - t0) * kTileBytes;
                const uint32_t st    = (seq + c) % kStages;
                mbar_expect_tx(&bar[st], bytes);
                bulk_g2s(buf + st * CHUNK, p.tiles + (size_t)t0 * kTileBytes, bytes, &bar[st]);
            }
        }

        // A fragments of the warp's 4 x 16 frames (m16n8k32, 8-bit): a0 (row g, k 4q..), a1 (row g+8, same k),
        // a2 (row g, k + 16), a3 (row g+8, k + 16); two k-steps of 32 dims
        uint32_t a[4][2][4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const long r0 = min(f0 + mt * 16 + g, p.T - 1), r1 = min(f0 + mt * 16 + 8 + g, p.T - 1);
            const uint32_t* x0 = reinterpret_cast<const uint32_t*>(p.xq + (size_t)r0 * kKBytes);
            const uint32_t* x1 = reinterpret_cast<const uint32_t*>(p.xq + (size_t)r1 * kKBytes);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                a[mt][ks][0] = __ldg(x0 + ks * 8 + q);
                a[mt][ks][1] = __ldg(x1 + ks * 8 + q);
                a[mt][ks][2] = __ldg(x0 + ks * 8 + 4 + q);
                a[mt][ks][3] = __ldg(x1 + ks * 8 + 4 + q);
            }
        }
        const long rowA = f0 + mtOwn * 16 + g, rowB = rowA + 8;
        const int  xsA = __ldg(p.xsq + min(rowA, p.T - 1)), xsB = __ldg(p.xsq + min(rowB, p.T - 1));
        // output cursors: next mixture of this lane's two rows
        float* outA = p.scores + (size_t)min(rowA, p.T - 1) * p.nMix + p.grpMix[grp];
        float* outB = p.scores + (size_t)min(rowB, p.T - 1) * p.nMix + p.grpMix[grp];
        const bool liveA = rowA < p.T, liveB = rowB < p.T;

        int   best[8];  // index mt * 2 + h: running min of (c + |m|^2 - 2 x.m) for row mt * 16 + h * 8 + g
        float oA[4] = {0, 0, 0, 0}, oB[4] = {0, 0, 0, 0};
        int   nStaged = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            best[j] = INT_MAX;

        // products of one tile: 8 IMMAs into acc
        auto products = [&](uint32_t tileAddr, int (&acc)[4][4]) {
            uint32_t b[4];
            ldsm_x4(tileAddr + ldsmOff, b);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
                imma_u8(acc[mt], a[mt][0], b[0], b[1], zero);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
                imma_u8(acc[mt], a[mt][1], b[2], b[3], acc[mt]);
        };
        // min over the tile's columns, and at the end of a mixture the reduction over the quad + emit
        auto consume = [&](uint32_t tileAddr, const int (&acc)[4][4]) {
            const int2 cc    = lds_int2(tileAddr + ccOff);
            const int  flags = lds_int(tileAddr + 8 * kRowBytes + 32);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                best[mt * 2]     = min3(best[mt * 2], cc.x - 2 * acc[mt][0], cc.y - 2 * acc[mt][1]);
                best[mt * 2 + 1] = min3(best[mt * 2 + 1], cc.x - 2 * acc[mt][2], cc.y - 2 * acc[mt][3]);
            }
            if (flags & 1) {  // last tile of its mixture (warp-uniform)
                // butterfly: after xor 1 a lane keeps half of the 8 rows, after xor 2 a quarter
                int v4[4], v2[2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int keep = (q & 1) ? best[j + 4] : best[j];
                    const int send = (q & 1) ? best[j] : best[j + 4];
                    v4[j]          = min(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int keep = (q & 2) ? v4[j + 2] : v4[j];
                    const int send = (q & 2) ? v4[j] : v4[j + 2];
                    v2[j]          = min(keep, __shfl_xor_sync(0xffffffffu, send, 2));
                }
                // empty mixture: the reference's running minimum stays at INT_MAX (:479-481)
                const int   bA = (flags & 2) ? INT_MAX : v2[0] + xsA, bB = (flags & 2) ? INT_MAX : v2[1] + xsB;
                const float sA = score_of(bA, p.scale, p.rcpScale, p.simd), sB = score_of(bB, p.scale, p.rcpScale, p.simd);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    best[j] = INT_MAX;
                if (p.vec4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // nStaged is warp-uniform; static register indices
                        if (nStaged == k) {
                            oA[k] = sA;
                            oB[k] = sB;
                        }
                    if (++nStaged == 4) {
                        nStaged = 0;
                        if (liveA)
                            *reinterpret_cast<float4*>(outA) = make_float4(oA[0], oA[1], oA[2], oA[3]);
                        if (liveB)
                            *reinterpret_cast<float4*>(outB) = make_float4(oB[0], oB[1], oB[2], oB[3]);
                        outA += 4;
                        outB += 4;
                    }
                }
                else {
                    if (liveA)
                        *outA = sA;
                    if (liveB)
                        *outB = sB;
                    ++outA;
                    ++outB;
                }
            }
        };

        for (int c = 0; c < nCh; ++c) {
            if (tid == 0 && c + kStages - 1 < nCh) {
                const int      cc    = c + kStages - 1;
                const int      t0    = tile0 + cc * kChunkTiles;
                const uint32_t bytes = (uint32_t)min(kChunkTiles, tile1 - t0) * kTileBytes;
                const uint32_t st    = (seq + cc) % kStages;
                mbar_expect_tx(&bar[st], bytes);
                bulk_g2s(buf + st * CHUNK, p.tiles + (size_t)t0 * kTileBytes, bytes, &bar[st]);
            }
            const uint32_t n  = seq + c;
            const uint32_t st = n % kStages;
            mbar_wait(&bar[st], (n / kStages) & 1u);
            const int nt   = min(kChunkTiles, tile1 - (tile0 + c * kChunkTiles));
            uint32_t  addr = sbuf + st * CHUNK;

            // software pipeline inside the chunk: the IMMAs of tile t+1 are in flight while tile t is reduced
            int accA[4][4], accB[4][4];
            products(addr, accA);
            int t = 0;
            for (; t + 2 <= nt; t += 2) {
                products(addr + kTileBytes, accB);  // tile t+1 exists
                consume(addr, accA);
                if (t + 2 < nt)
                    products(addr + 2 * kTileBytes, accA);
                consume(addr + kTileBytes, accB);
                addr += 2 * kTileBytes;
            }
            if (t < nt)
                consume(addr, accA);
            __syncthreads();  // everyone is done with stage st before it is refilled
        }
        seq += (uint32_t)nCh;
    }
}

}  // namespace

// ==========================================================================================
// host side
// ==========================================================================================

struct rb_gmm_int {
    rb::DeviceInfo dev;
    int            dim = 0, nMix = 0, nTiles = 0, ctasPerSm = 1, curGroups = -1;
    float          scale = 1.0f, rcpScale = 0.0f;
    bool           simd = false;
    size_t         smemBytes = 0;
    std::vector<int> tilesOfMixture, groupsOf;
    rb::DevBuf<unsigned char> dTiles, dXq;
    rb::DevBuf<float>         dVariance;
    rb::DevBuf<int>           dXsq, dGrpTile, dGrpMix;
};

namespace {

// Mm::quantize<f32, u8> (src/Mm/Utilities.hh:190-202)
unsigned char quantize_u8(float x) {
    const int v = (int)std::round(x) + 128;
    return (unsigned char)std::min(std::max(v, 0), 255);
}

// does q0 = a r, q = fma(fma(-s, q0, a), r, q0) equal the IEEE quotient a / s for every integer |a| <= 2^24 ?
__attribute__((target("fma"))) bool fast_division_exact_fma(float sc, float r) {
    bool ok = true;
    for (int b = -(1 << 24); ok && b <= (1 << 24); ++b) {
        const float a = (float)b, q0 = a * r;
        ok = __builtin_fmaf(__builtin_fmaf(-sc, q0, a), r, q0) == a / sc;
    }
    return ok;
}
bool fast_division_exact_soft(float sc, float r) {
    bool ok = true;
    for (int b = -(1 << 24); ok && b <= (1 << 24); ++b) {
        const float a = (float)b, q0 = a * r;
        ok = std::fmaf(std::fmaf(-sc, q0, a), r, q0) == a / sc;
    }
    return ok;
}

constexpr int kMaxGroups = 64, kGroupStride = kMaxGroups + 2;

void make_groups(const rb_gmm_int* h, int G, std::vector<int>& grpTile, std::vector<int>& grpMix) {
    grpTile.assign(1, 0);
    grpMix.assign(1, 0);
    long acc = 0;
    int  g   = 1;
    for (int m = 0; m < h->nMix; ++m) {
        acc += h->tilesOfMixture[m];
        const bool boundaryOk = ((m + 1) % 4 == 0) && (m + 1 < h->nMix);  // keeps the float4 stores aligned
        if (g < G && boundaryOk && acc * G >= (long)h->nTiles * g) {
            grpTile.push_back((int)acc);
            grpMix.push_back(m + 1);
            ++g;
        }
    }
    grpTile.push_back(h->nTiles);
    grpMix.push_back(h->nMix);
}

int choose_groups(const rb_gmm_int* h, long T, int slots) {
    const long FB   = (T + kBlockFrames - 1) / kBlockFrames;
    const int  gmax = std::max(1, std::min(64, std::min(h->nMix / 4, h->nTiles / 16)));
    int        best = 1;
    double     bestEff = -1;
    for (int G = 1; G <= gmax; ++G) {
        const long   items = FB * G, waves = (items + slots - 1) / slots;
        const double eff   = (double)items / (double)(waves * slots);
        if (eff > bestEff + 0.02) {
            bestEff = eff;
            best    = G;
        }
    }
    return best;
}

}  // namespace

int rb_gmm_int_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream, rb_gmm_int** out, bool simd) {
    *out = nullptr;
    if (ms->n_covariances != 1) {
        rb::set_error("int feature scorer supports only globally pooled variance (got %u covariances)",
                      ms->n_covariances);
        return RB_ERR_UNSUPPORTED;
    }
    if (ms->dim > (unsigned)kKBytes) {
        rb::set_error("int feature scorer supports feature dimension <= %d (got %u)", kKBytes, ms->dim);
        return RB_ERR_UNSUPPORTED;
    }
    rb_gmm_int* h = new (std::nothrow) rb_gmm_int();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    auto fail = [&](int code) {
        delete h;
        return code;
    };
    const unsigned D = ms->dim;
    h->dev  = dev;
    h->dim  = (int)D;
    h->nMix = (int)ms->n_mixtures;

    // BatchIntFeatureScorer::init (:375-416) with the f32 / f64 mixing of the reference
    std::vector<float> variance(kKBytes, 0.0f);
    for (unsigned d = 0; d < D; ++d)
        variance[d] = 1.0f / (float)std::sqrt((double)ms->variances[d]);
    float minMean = 3.40282347e+38f, maxMean = -3.40282347e+38f;  // quantizationScale (:355-373): all densities
    for (uint32_t i = 0; i < ms->n_densities; ++i) {
        const float* mu = ms->means + (size_t)ms->dens_mean[i] * D;
        for (unsigned d = 0; d < D; ++d) {
            const float divided = mu[d] * variance[d];
            minMean             = std::min(minMean, divided);
            maxMean             = std::max(maxMean, divided);
        }
    }
    const float intervalSize = 2 * std::max(std::fabs(minMean), std::fabs(maxMean));
    const float scale        = (float)((double)255.0f / (1.25 * (double)intervalSize));
    const float scaleSquared = scale * scale;
    h->scale                 = (float)(2.0 * (double)scaleSquared);
    h->simd                  = simd;
    if (simd)
        h->scale = scaleSquared;  // scalingSquared_; the score is 0.5 * int / scalingSquared_ in f64
    else {  // exhaustive proof of the fast division for this scale (33.5 M quotients, 0.06-0.2 s, once per model)
        const float r  = 1.0f / h->scale;
        const bool  ok = std::isfinite(r) && r != 0.0f &&
                        (__builtin_cpu_supports("fma") ? fast_division_exact_fma(h->scale, r)
                                                       : fast_division_exact_soft(h->scale, r));
        h->rcpScale = ok ? r : 0.0f;
    }
    for (unsigned d = 0; d < D; ++d)
        variance[d] = variance[d] * scale;
    double sumLog = 0;
    for (unsigned d = 0; d < D; ++d)
        sumLog += std::log(std::fabs((double)ms->variances[d]));
    const float logNorm       = (float)((double)D * std::log(2.0 * M_PI) + sumLog);
    const float logNormFactor = logNorm * scaleSquared;

    // tiles: mixtures in order, padded to whole 8-column tiles
    std::vector<unsigned char> tiles;
    h->tilesOfMixture.assign(ms->n_mixtures, 0);
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        const uint32_t e0 = ms->mix_offsets[m], n = ms->mix_offsets[m + 1] - e0;
        const uint32_t nt = std::max<uint32_t>(1, (n + 7) / 8);
        h->tilesOfMixture[m] = (int)nt;
        for (uint32_t t = 0; t < nt; ++t) {
            const size_t base = tiles.size();
            tiles.resize(base + kTileBytes, 0);
            unsigned char* tile = tiles.data() + base;
            int*           cc   = reinterpret_cast<int*>(tile + 8 * kRowBytes);
            for (uint32_t r = 0; r < 8; ++r) {
                const uint32_t i = t * 8 + r;
                if (i >= n) {
                    cc[r] = kDummy;
                    continue;
                }
                const uint32_t dns = ms->mix_density[e0 + i];
                if (ms->dens_cov[dns] != 0) {
                    rb::set_error("density %u does not use covariance 0", dns);
                    return fail(RB_ERR_INVALID);
                }
                const float* mu = ms->means + (size_t)ms->dens_mean[dns] * D;
                int          m2 = 0;
                for (unsigned d = 0; d < D; ++d) {
                    const unsigned char qv = quantize_u8(mu[d] * variance[d]);
                    tile[r * kRowBytes + d] = qv;
                    m2 += (int)qv * (int)qv;
                }
                int c;
                if (simd) {
                    // createDensityElement (src/Mm/IntelOptimization.cc:39-49): Weight (f64) = f32 * -2 * f64, handed over
                    // as Score (f32); constantWeight_ = (s32)(f32 + f32)
                    const float sum = (float)((double)(scaleSquared * -2.0f) * ms->mix_log_weight[e0 + i]) + logNormFactor;
                    c = std::fabs(sum) < 2147483648.0f ? (int)sum : INT_MIN;
                }
                else
                    c = (int)((double)logNormFactor - (double)h->scale * ms->mix_log_weight[e0 + i]);
                cc[r] = c + m2;
            }
            int flags = (t + 1 == nt ? 1 : 0) | (n == 0 ? 2 : 0);
            std::memcpy(tile + 8 * kRowBytes + 32, &flags, 4);
        }
    }
    h->nTiles    = (int)(tiles.size() / kTileBytes);
    h->smemBytes = (size_t)kStages * kChunkTiles * kTileBytes + sizeof(uint64_t) * kStages;
    int occ = 0;
    if (cudaFuncSetAttribute(gmm_int_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBytes) !=
                cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gmm_int_kernel, kThreads, h->smemBytes) != cudaSuccess ||
        occ < 1) {
        rb::set_error("int gmm kernel does not fit on the device: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(RB_ERR_CUDA);
    }
    h->ctasPerSm = occ;
    {  // group tables for every group count: a launch never has to touch them
        std::vector<int> allTile((size_t)(kMaxGroups + 1) * kGroupStride, 0), allMix(allTile.size(), 0);
        h->groupsOf.assign(kMaxGroups + 1, 1);
        for (int G = 1; G <= kMaxGroups; ++G) {
            std::vector<int> grpTile, grpMix;
            make_groups(h, G, grpTile, grpMix);
            std::copy(grpTile.begin(), grpTile.end(), allTile.begin() + (size_t)G * kGroupStride);
            std::copy(grpMix.begin(), grpMix.end(), allMix.begin() + (size_t)G * kGroupStride);
            h->groupsOf[G] = (int)grpTile.size() - 1;
        }
        if (h->dGrpTile.upload(allTile, stream) != RB_OK || h->dGrpMix.upload(allMix, stream) != RB_OK)
            return fail(RB_ERR_CUDA);
    }
    if (h->dTiles.upload(tiles, stream) != RB_OK || h->dVariance.upload(variance, stream) != RB_OK)
        return fail(RB_ERR_CUDA);
    if (cudaStreamSynchronize(stream) != cudaSuccess) {
        rb::set_error("int gmm model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(RB_ERR_CUDA);
    }
    *out = h;
    return RB_OK;
}

void rb_gmm_int_destroy(rb_gmm_int* h) {
    delete h;
}

int rb_gmm_int_score(rb_gmm_int* h, const float* dFeats, long T, float* dScores, cudaStream_t s) {
    RB_CHECK(h->dXq.reserve((size_t)T * kKBytes));
    RB_CHECK(h->dXsq.reserve((size_t)T));
    const int qblocks = (int)std::min<long>((T + 15) / 16, (long)h->dev.sm_count * 16);
    gmm_int_quantize_kernel<<<qblocks, 256, 0, s>>>(dFeats, h->dVariance.p, T, h->dim, h->dXq.p, h->dXsq.p);
    RB_LAUNCH_CHECK();
    const int slots = h->dev.sm_count * h->ctasPerSm;
    const int G     = choose_groups(h, T, slots);
    h->curGroups    = h->groupsOf[G];
    IntParams p;
    p.tiles        = h->dTiles.p;
    p.grpTile      = h->dGrpTile.p + (size_t)G * kGroupStride;  // tables of every G live on the device
    p.grpMix       = h->dGrpMix.p + (size_t)G * kGroupStride;
    p.xq           = h->dXq.p;
    p.xsq          = h->dXsq.p;
    p.scores       = dScores;
    p.T            = T;
    p.nMix         = h->nMix;
    p.nGroups      = h->curGroups;
    p.nFrameBlocks = (int)((T + kBlockFrames - 1) / kBlockFrames);
    p.vec4         = (h->nMix % 4 == 0 && ((uintptr_t)dScores % 16 == 0)) ? 1 : 0;
    p.scale        = h->scale;
    p.rcpScale     = h->rcpScale;
    p.simd         = h->simd ? 1 : 0;
    const long items = (long)p.nGroups * p.nFrameBlocks;
    gmm_int_kernel<<<(int)std::min<long>(items, slots), kThreads, h->smemBytes, s>>>(p);
    RB_LAUNCH_CHECK();
    return RB_OK;
}
