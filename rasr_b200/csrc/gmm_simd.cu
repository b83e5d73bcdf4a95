// gmm_simd.cu -- RB_GMM_SIMD_DIAG_MAX: Mm::SimdGaussDiagonalMaximumFeatureScorer ("SIMD-diagonal-maximum") for ALL
// mixtures and ALL frames at once, scores and best densities bit-identical to the CPU path.
//
//   init / getScaling / quantizationScalingFactor   src/Mm/SimdFeatureScorer.cc:62-135
//   buildMixtureTable / createDensityElement        src/Mm/SimdFeatureScorer.cc:79-104, src/Mm/IntelOptimization.cc:39-49
//   multiplyAndQuantize / quantize                  src/Mm/IntelOptimization.cc:51-69, src/Mm/Utilities.hh:190-202
//   calculateScoreAndDensity / quantizedScore       src/Mm/SimdFeatureScorer.cc:137-176
//   distance                                        the machine code src/Mm/SSE2CodeGenerator.cc emits: sum (m - x)^2
//
// The reference quantises the feature vector once PER COVARIANCE (x_d / sqrt(var_c,d) * scaling -> u8), scores a
// density as  c_k + sum_d (m_kd - x_cd)^2  in s32 over u8 means and keeps the first minimum over the densities of a
// mixture; the result is (f32)(0.5 * int / scaling^2), evaluated in f64.  Integer arithmetic is exact, so
//     c_k + sum_d (m_kd - x_cd)^2  =  (c_k + |m_k|^2) + |x_c|^2 - 2 m_k.x_c
// Two kernels.  (1) Pooled covariance and scores only -- what the recognizer asks for: the IMMA kernel of gmm_int.cu
// (mma.sync u8 x u8 -> s32 on the tensor cores) with this scorer's density constants and score formula.  (2) Everything
// else (several covariances, best densities wanted): the kernel below,
// with the u8 x u8 inner product on DP4A (4 products per instruction).  A thread owns two frames whose quantised
// features stay in registers (pooled covariance, the usual RASR model) and walks the densities of a group of
// mixtures staged in shared memory: every mean word is one broadcast LDS.128 for the whole CTA, so a (frame, density)
// pair costs ~12 instructions against 2 x 39 for the float scorer.  With several covariances the quantised features
// of the density's covariance are fetched per density from a [covariance][frame] table in global memory -- correct for
// any model, fast for few covariances; frames are processed in slices so that the table stays below 256 MB.
#include <climits>
#include <cmath>

#include "common.cuh"

struct rb_gmm_simd;
int  rb_gmm_simd_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream, rb_gmm_simd** out);
void rb_gmm_simd_destroy(rb_gmm_simd* h);
int  rb_gmm_simd_score(rb_gmm_simd* h, const float* d_feats, long T, float* d_scores, uint32_t* d_best, cudaStream_t stream);

// gmm_int.cu: the tensor-core (IMMA) kernel of the batch-int scorer, with this scorer's constants and score formula
struct rb_gmm_int;
int  rb_gmm_int_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream, rb_gmm_int** out, bool simd);
void rb_gmm_int_destroy(rb_gmm_int* h);
int  rb_gmm_int_score(rb_gmm_int* h, const float* d_feats, long T, float* d_scores, cudaStream_t stream);

namespace {

constexpr int    kThreads     = 256;
constexpr int    kFpt         = 2;                  // frames per thread
constexpr int    kBlockFrames = kThreads * kFpt;    // 512
constexpr int    kMaxDim      = 64;
constexpr size_t kGroupBytes  = 16 * 1024;          // shared memory of one mixture group (means + constants): small groups = many
                                                    // work items also for the 2048..16384-frame slabs of the host-buffer calls
constexpr size_t kTableBytes  = 256u << 20;         // quantised features of one slice of frames, all covariances

struct SimdParams {
    const uint4*    means;    // [nDens][NQ] quantised means in mixture order, zero padded to NQ * 16 dims
    const int*      consts;   // [nDens] c_k + |m_k|^2
    const uint32_t* cov;      // [nDens] covariance of the density
    const uint32_t* mixOff;   // [nMix + 1] densities of a mixture
    const int*      grpMix;   // [nGroups + 1] mixtures of a group
    const uint4*    xq;       // [nCov][T][NQ] quantised features of this slice
    const int*      xsq;      // [nCov][T]
    float*          scores;   // [T][nMix]
    uint32_t*       best;     // [T][nMix] or nullptr
    long            T;        // frames of this slice
    int             nMix, nGroups, nFrameBlocks;
    int             vec4;     // nMix % 4 == 0, 16-byte aligned outputs: four mixtures per store
    float           scalingSquared;
};

// multiplyAndQuantize: u8 = clip((int)round(f * isd_c * scaling) + 128); one thread per (covariance, frame, 16 dims)
__global__ void __launch_bounds__(256) gmm_simd_quantize_kernel(const float* __restrict__ feats, const float* __restrict__ isd,
                                                                long T, int dim, int nq, int nCov,
                                                                uint4* __restrict__ xq, int* __restrict__ xsq) {
    const long total = (long)nCov * T;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long   t = i % T;
        const int    c = (int)(i / T);
        const float* f = feats + (size_t)t * dim;
        const float* v = isd + (size_t)c * dim;
        int          sq = 0;
        for (int q4 = 0; q4 < nq; ++q4) {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int d = q4 * 16 + j;
                int       q = 0;  // padding dims are 0 in features and means
                if (d < dim) {
                    const float r = roundf(__fmul_rn(__ldg(f + d), __ldg(v + d)));
                    // (int) of an out-of-range or NaN float is INT_MIN on the reference's x86 (cvttss2si)
                    const int k = fabsf(r) < 2147483648.0f ? __float2int_rz(r) : INT_MIN;
                    q           = min(max(k + (k < INT_MAX - 128 ? 128 : 0), 0), 255);
                }
                w[j >> 2] |= (uint32_t)q << (8 * (j & 3));
                sq += q * q;
            }
            xq[(size_t)i * nq + q4] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        xsq[i] = sq;
    }
}

__device__ __forceinline__ int dot16(const uint4& a, const uint4& b, int acc) {
    unsigned u = __dp4a(a.x, b.x, (unsigned)acc);  // unsigned x unsigned
    u = __dp4a(a.y, b.y, u);
    u = __dp4a(a.z, b.z, u);
    return (int)__dp4a(a.w, b.w, u);
}

template<int NQ, bool ONE_COV>
__global__ void __launch_bounds__(kThreads) gmm_simd_kernel(const SimdParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int  tid = threadIdx.x;
    const long items = (long)p.nGroups * p.nFrameBlocks;
    int        staged = -1;
    for (long item = blockIdx.x; item < items; item += gridDim.x) {
        // consecutive items of a CTA share the mixture group while the grid is smaller than the frame blocks
        const int g  = (int)(item / p.nFrameBlocks);
        const int fb = (int)(item % p.nFrameBlocks);
        const int m0 = p.grpMix[g], m1 = p.grpMix[g + 1];
        const uint32_t k0 = p.mixOff[m0], k1 = p.mixOff[m1];
        uint4*    sMeans = reinterpret_cast<uint4*>(smem);
        int*      sConst = reinterpret_cast<int*>(sMeans + (size_t)(k1 - k0) * NQ);
        uint32_t* sCov   = reinterpret_cast<uint32_t*>(sConst + (k1 - k0));
        if (g != staged) {
            __syncthreads();
            for (uint32_t i = tid; i < (k1 - k0) * NQ; i += kThreads)
                sMeans[i] = p.means[(size_t)k0 * NQ + i];
            for (uint32_t i = tid; i < k1 - k0; i += kThreads) {
                sConst[i] = p.consts[k0 + i];
                sCov[i]   = p.cov[k0 + i];
            }
            __syncthreads();
            staged = g;
        }
        long t[kFpt];
        bool live[kFpt];
#pragma unroll
        for (int f = 0; f < kFpt; ++f) {
            t[f]    = (long)fb * kBlockFrames + f * kThreads + tid;
            live[f] = t[f] < p.T;
            if (!live[f])
                t[f] = p.T - 1;
        }
        uint4 x[kFpt][NQ];
        int   xs[kFpt];
        if (ONE_COV) {
#pragma unroll
            for (int f = 0; f < kFpt; ++f) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    x[f][q] = __ldg(p.xq + (size_t)t[f] * NQ + q);
                xs[f] = __ldg(p.xsq + t[f]);
            }
        }
        // four mixtures at a time when the rows of the score matrix allow 16-byte stores (groups then start at
        // multiples of 4): one STG.128 per frame instead of four scattered 4-byte stores
        const int step = p.vec4 ? 4 : 1;
        for (int mq = m0; mq < m1; mq += step) {
            float    outS[kFpt][4];
            uint32_t outB[kFpt][4];
            for (int j = 0; j < step; ++j) {
                const int      m = mq + j;
                const uint32_t a = p.mixOff[m] - k0, b = p.mixOff[m + 1] - k0;
                int      bestScore[kFpt];
                uint32_t bestDns[kFpt];
#pragma unroll
                for (int f = 0; f < kFpt; ++f) {
                    bestScore[f] = INT_MAX;  // an empty mixture keeps Core::Type<int>::max and bestDensity = max
                    bestDns[f]   = 0xffffffffu;
                }
                for (uint32_t k = a; k < b; ++k) {
                    uint4 mean[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        mean[q] = sMeans[(size_t)k * NQ + q];
                    const int c = sConst[k];
                    if (!ONE_COV) {
                        const size_t row = (size_t)sCov[k] * p.T;
#pragma unroll
                        for (int f = 0; f < kFpt; ++f) {
#pragma unroll
                            for (int q = 0; q < NQ; ++q)
                                x[f][q] = __ldg(p.xq + (row + t[f]) * NQ + q);
                            xs[f] = __ldg(p.xsq + row + t[f]);
                        }
                    }
#pragma unroll
                    for (int f = 0; f < kFpt; ++f) {
                        int dot = 0;
#pragma unroll
                        for (int q = 0; q < NQ; ++q)
                            dot = dot16(mean[q], x[f][q], dot);
                        const int score = c + xs[f] - 2 * dot;
                        if (score < bestScore[f]) {  // the first minimum wins (:166-169)
                            bestScore[f] = score;
                            bestDns[f]   = k - a;
                        }
                    }
                }
#pragma unroll
                for (int f = 0; f < kFpt; ++f) {
                    // result.score = 0.5 * quantizedResult.first / scalingSquared_  (f64, then Score = f32)
                    const float sc = (float)(0.5 * (double)bestScore[f] / (double)p.scalingSquared);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)  // static register indices
                        if (jj == j) {
                            outS[f][jj] = sc;
                            outB[f][jj] = bestDns[f];
                        }
                }
            }
#pragma unroll
            for (int f = 0; f < kFpt; ++f)
                if (live[f]) {
                    const size_t at = (size_t)t[f] * p.nMix + mq;
                    if (p.vec4) {
                        *reinterpret_cast<float4*>(p.scores + at) = make_float4(outS[f][0], outS[f][1], outS[f][2], outS[f][3]);
                        if (p.best)
                            *reinterpret_cast<uint4*>(p.best + at) = make_uint4(outB[f][0], outB[f][1], outB[f][2], outB[f][3]);
                    }
                    else {
                        p.scores[at] = outS[f][0];
                        if (p.best)
                            p.best[at] = outB[f][0];
                    }
                }
        }
    }
}

// Mm::quantize<f32, u8> (src/Mm/Utilities.hh:190-202)
unsigned char quantize_u8(float x) {
    const int v = (int)std::round(x) + 128;
    return (unsigned char)std::min(std::max(v, 0), 255);
}

}  // namespace

struct rb_gmm_simd {
    rb::DeviceInfo        dev;
    int                   dim = 0, nMix = 0, nCov = 0, nq = 0, nGroups = 0;
    float                 scalingSquared = 1.0f;
    size_t                smemBytes = 0;
    rb::DevBuf<uint4>     dMeans, dXq;
    rb::DevBuf<int>       dConsts, dGrpMix, dXsq;
    rb::DevBuf<uint32_t>  dCov, dMixOff;
    rb::DevBuf<float>     dIsd;
    // pooled covariance and no density output wanted: the same integers come out of the IMMA kernel of gmm_int.cu
    // (u8 x u8 products on the tensor cores, 3.5 x the DP4A rate); the kernel above serves the rest
    rb_gmm_int*           imma = nullptr;
    ~rb_gmm_simd() {
        if (imma)
            rb_gmm_int_destroy(imma);
    }
};

int rb_gmm_simd_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream, rb_gmm_simd** out) {
    *out = nullptr;
    if (ms->dim > (unsigned)kMaxDim) {
        rb::set_error("SIMD feature scorer supports feature dimension <= %d (got %u)", kMaxDim, ms->dim);
        return RB_ERR_UNSUPPORTED;
    }
    rb_gmm_simd* h = new (std::nothrow) rb_gmm_simd();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    auto fail = [&](int code) {
        delete h;
        return code;
    };
    const unsigned D = ms->dim, nCov = ms->n_covariances;
    h->dev  = dev;
    h->dim  = (int)D;
    h->nMix = (int)ms->n_mixtures;
    h->nCov = (int)nCov;
    h->nq   = (int)((D + 15) / 16);
    const size_t rowBytes = (size_t)h->nq * 16;

    // init (:62-77): per covariance 1 / sqrt(var) and the log normalisation factor, f32 / f64 as the reference mixes them
    std::vector<float> isd((size_t)nCov * D), logNorm(nCov);
    for (unsigned c = 0; c < nCov; ++c) {
        const float* var    = ms->variances + (size_t)c * D;
        double       sumLog = 0;
        for (unsigned d = 0; d < D; ++d) {
            if (!(var[d] > 0.0f)) {
                rb::set_error("covariance %u has a variance <= 0 in dimension %u", c, d);
                return fail(RB_ERR_INVALID);
            }
            isd[(size_t)c * D + d] = 1.0f / (float)std::sqrt((double)var[d]);
            sumLog += std::log(std::fabs((double)var[d]));
        }
        logNorm[c] = (float)((double)D * std::log(2.0 * M_PI) + sumLog);
    }
    // getScaling (:106-127): range of mean / sqrt(var) over ALL densities of the set
    float minMean = 3.40282347e+38f, maxMean = -3.40282347e+38f;
    for (uint32_t i = 0; i < ms->n_densities; ++i) {
        if (ms->dens_mean[i] >= ms->n_means || ms->dens_cov[i] >= nCov) {
            rb::set_error("density %u refers to mean %u / covariance %u outside the tables", i, ms->dens_mean[i], ms->dens_cov[i]);
            return fail(RB_ERR_INVALID);
        }
        const float* mu = ms->means + (size_t)ms->dens_mean[i] * D;
        const float* sd = isd.data() + (size_t)ms->dens_cov[i] * D;
        for (unsigned d = 0; d < D; ++d) {
            const float divided = mu[d] * sd[d];
            minMean             = std::min(minMean, divided);
            maxMean             = std::max(maxMean, divided);
        }
    }
    const float intervalSize = 2 * std::max(std::fabs(minMean), std::fabs(maxMean));
    const float scaling      = (float)((double)255.0f / (1.25 * (double)intervalSize));
    h->scalingSquared        = scaling * scaling;
    for (unsigned c = 0; c < nCov; ++c) {  // CovarianceFeatureScorerElement::scale
        for (unsigned d = 0; d < D; ++d)
            isd[(size_t)c * D + d] = isd[(size_t)c * D + d] * scaling;
        logNorm[c] = logNorm[c] * (scaling * scaling);
    }

    // buildMixtureTable (:79-104): densities in mixture order
    const uint32_t             nDens = ms->mix_offsets[ms->n_mixtures];
    std::vector<unsigned char> means((size_t)std::max<uint32_t>(nDens, 1) * rowBytes, 0);
    std::vector<int>           consts(std::max<uint32_t>(nDens, 1), 0);
    std::vector<uint32_t>      cov(std::max<uint32_t>(nDens, 1), 0);
    for (uint32_t e = 0; e < nDens; ++e) {
        const uint32_t dns = ms->mix_density[e];
        const uint32_t c   = ms->dens_cov[dns];
        const float*   mu  = ms->means + (size_t)ms->dens_mean[dns] * D;
        int            m2  = 0;
        for (unsigned d = 0; d < D; ++d) {
            const unsigned char q = quantize_u8(mu[d] * isd[(size_t)c * D + d]);
            means[e * rowBytes + d] = q;
            m2 += (int)q * (int)q;
        }
        // Weight (f64) = f32 * -2 * f64; handed over as Score (f32); constantWeight_ = (s32)(f32 + f32)
        const double scaledMinus2LogWeight = (double)(h->scalingSquared * -2.0f) * ms->mix_log_weight[e];
        const float  sum                   = (float)scaledMinus2LogWeight + logNorm[c];
        const int    constantWeight        = std::fabs(sum) < 2147483648.0f ? (int)sum : INT_MIN;  // cvttss2si
        consts[e] = constantWeight + m2;
        cov[e]    = c;
    }
    // mixture groups: as many whole mixtures as fit the shared-memory budget
    std::vector<int> grpMix(1, 0);
    size_t           used = 0, largest = 0;
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        const size_t bytes = (size_t)(ms->mix_offsets[m + 1] - ms->mix_offsets[m]) * (rowBytes + 8);
        if (used && used + bytes > kGroupBytes && (ms->n_mixtures % 4 != 0 || m % 4 == 0)) {
            grpMix.push_back((int)m);
            used = 0;
        }
        used += bytes;
        largest = std::max(largest, used);
    }
    grpMix.push_back((int)ms->n_mixtures);
    h->nGroups   = (int)grpMix.size() - 1;
    h->smemBytes = std::max<size_t>(largest, 16);
    if (h->smemBytes + 1024 > dev.smem_optin) {
        rb::set_error("a mixture of the model needs %zu bytes of shared memory", h->smemBytes);
        return fail(RB_ERR_UNSUPPORTED);
    }
    if (h->dMeans.upload(reinterpret_cast<const uint4*>(means.data()), means.size() / 16, stream) != RB_OK ||
        h->dConsts.upload(consts, stream) != RB_OK || h->dCov.upload(cov, stream) != RB_OK ||
        h->dMixOff.upload(ms->mix_offsets, ms->n_mixtures + 1, stream) != RB_OK ||
        h->dGrpMix.upload(grpMix, stream) != RB_OK || h->dIsd.upload(isd, stream) != RB_OK ||
        cudaStreamSynchronize(stream) != cudaSuccess) {
        rb::set_error("SIMD gmm model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(RB_ERR_CUDA);
    }
    bool pooled = nCov == 1;
    for (uint32_t i = 0; i < ms->n_densities && pooled; ++i)
        pooled = ms->dens_cov[i] == 0;
    if (pooled && getenv("RB_GMM_SIMD_DP4A") == nullptr) {
        int rc = rb_gmm_int_create(ms, dev, stream, &h->imma, true);
        if (rc != RB_OK)
            return fail(rc);
    }
    *out = h;
    return RB_OK;
}

void rb_gmm_simd_destroy(rb_gmm_simd* h) {
    delete h;
}

template<int NQ>
static int launch_simd(rb_gmm_simd* h, const SimdParams& p, cudaStream_t s) {
    auto k = h->nCov == 1 ? gmm_simd_kernel<NQ, true> : gmm_simd_kernel<NQ, false>;
    RB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBytes));
    const long items = (long)p.nGroups * p.nFrameBlocks;
    const int  perSm = (int)std::max<size_t>(1, std::min<size_t>(8, h->dev.smem_optin / (h->smemBytes + 1024)));
    k<<<(int)std::min<long>(items, (long)h->dev.sm_count * perSm), kThreads, h->smemBytes, s>>>(p);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

int rb_gmm_simd_score(rb_gmm_simd* h, const float* dFeats, long T, float* dScores, uint32_t* dBest, cudaStream_t s) {
    if (h->imma && !dBest)
        return rb_gmm_int_score(h->imma, dFeats, T, dScores, s);
    // slices of frames: the table of quantised features [covariance][frame] stays below kTableBytes
    const size_t perFrame = (size_t)h->nCov * ((size_t)h->nq * 16 + 4);
    long         slice    = (long)std::max<size_t>(kBlockFrames, kTableBytes / perFrame / kBlockFrames * kBlockFrames);
    slice                 = std::min(slice, T);
    RB_CHECK(h->dXq.reserve((size_t)h->nCov * slice * h->nq));
    RB_CHECK(h->dXsq.reserve((size_t)h->nCov * slice));
    for (long t0 = 0; t0 < T; t0 += slice) {
        const long n      = std::min(slice, T - t0);
        const long total  = (long)h->nCov * n;
        const int  blocks = (int)std::min<long>((total + 255) / 256, (long)h->dev.sm_count * 16);
        gmm_simd_quantize_kernel<<<blocks, 256, 0, s>>>(dFeats + (size_t)t0 * h->dim, h->dIsd.p, n, h->dim, h->nq, h->nCov,
                                                        h->dXq.p, h->dXsq.p);
        RB_LAUNCH_CHECK();
        SimdParams p;
        p.means          = h->dMeans.p;
        p.consts         = h->dConsts.p;
        p.cov            = h->dCov.p;
        p.mixOff         = h->dMixOff.p;
        p.grpMix         = h->dGrpMix.p;
        p.xq             = h->dXq.p;
        p.xsq            = h->dXsq.p;
        p.scores         = dScores + (size_t)t0 * h->nMix;
        p.best           = dBest ? dBest + (size_t)t0 * h->nMix : nullptr;
        p.T              = n;
        p.nMix           = h->nMix;
        p.nGroups        = h->nGroups;
        p.nFrameBlocks   = (int)((n + kBlockFrames - 1) / kBlockFrames);
        p.scalingSquared = h->scalingSquared;
        p.vec4           = (h->nMix % 4 == 0 && (uintptr_t)p.scores % 16 == 0 && (uintptr_t)p.best % 16 == 0) ? 1 : 0;
        int rc = h->nq == 1 ? launch_simd<1>(h, p, s)
                            : (h->nq == 2 ? launch_simd<2>(h, p, s) : (h->nq == 3 ? launch_simd<3>(h, p, s) : launch_simd<4>(h, p, s)));
        if (rc != RB_OK)
            return rc;
    }
    return RB_OK;
}
