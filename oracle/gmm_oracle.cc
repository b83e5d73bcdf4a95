/*
 * gmm_oracle.cc -- CPU restatement of the reference's diagonal-covariance GMM feature scorers.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY PINNED: the reference has no unit test for any Mm
 * scorer, so its own scorers -- src/Mm compiled into oracle/_ref, created by Mm::Module's factory --
 * are the check: every function here equals them bit for bit (tests/test_ref_parity.py: C2 model,
 * ragged / multi-covariance models, preselection parameters, tied distances, buffer sizes).  The
 * restatement follows the cited lines, including accumulation order and the f32/f64 mixing.
 *
 * Build with -ffp-contract=off.  use_fma selects the contraction the reference's default build
 * (gcc -O2 -march=native, -ffp-contract=fast) applies to `s += x * x`.
 */
#include "oracle.h"
#include "std_sort_restated.h"

#include <immintrin.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <set>
#include <vector>

namespace {

/* Mm::gaussLogNormFactor src/Mm/Utilities.hh:71-76 (+ logNorm :53-59) */
double gaussLogNormFactor(const float* var, unsigned dim) {
    double sumLog = 0;
    for (unsigned d = 0; d < dim; ++d)
        sumLog += log(std::fabs(var[d]));
    return (double)dim * log((double)2 * M_PI) + sumLog;
}

/* Mm::inverseSquareRoot<f32> src/Mm/Utilities.hh:86-91 */
inline float inverseSquareRoot(float x) {
    return (float)1 / (float)sqrt(x);
}

float* alignedFloats(size_t n) {
    void* p = 0;
    if (posix_memalign(&p, 32, std::max<size_t>(n, 8) * sizeof(float)) != 0)
        return 0;
    std::memset(p, 0, std::max<size_t>(n, 8) * sizeof(float));
    return (float*)p;
}

/* ------------------------------------------------------------------ Mm::BatchFloatFeatureScorer
 * init src/Mm/BatchFeatureScorer.cc:164-197, setFeature :157-162, fillScoreCacheTpl :207-253 */
struct BatchFloat {
    unsigned              dim, padded, nMix, nDens;
    std::vector<unsigned> offsets;
    float *               isd, *means, *consts;

    BatchFloat() : isd(0), means(0), consts(0) {}
    ~BatchFloat() {
        free(isd);
        free(means);
        free(consts);
    }

    int init(const orc_mixture_set& ms) {
        if (ms.n_covariances != 1)
            return -2; /* "feature scorer supports only globally pooled covariance" */
        dim    = ms.dim;
        padded = ((dim + 7) / 8) * 8;
        nMix   = ms.n_mixtures;
        offsets.assign(nMix + 1, 0);
        nDens = 0;
        for (unsigned m = 0; m < nMix; ++m) {
            offsets[m] = nDens;
            nDens += ms.mix_offsets[m + 1] - ms.mix_offsets[m];
        }
        offsets[nMix] = nDens;
        isd           = alignedFloats(padded);
        means         = alignedFloats((size_t)nDens * padded);
        consts        = alignedFloats(nDens);
        for (unsigned d = 0; d < dim; ++d)
            isd[d] = inverseSquareRoot(ms.variances[d]);
        const float logNormFactor = gaussLogNormFactor(ms.variances, dim);
        for (unsigned m = 0; m < nMix; ++m) {
            float* mean = means + (size_t)offsets[m] * padded;
            float* c    = consts + offsets[m];
            for (unsigned e = ms.mix_offsets[m]; e < ms.mix_offsets[m + 1]; ++e) {
                unsigned dns = ms.mix_density[e];
                if (ms.dens_cov[dns] != 0)
                    return -3;
                const float* mu = ms.means + (size_t)ms.dens_mean[dns] * dim;
                for (unsigned d = 0; d < dim; ++d)
                    mean[d] = mu[d] * isd[d];
                mean += padded;
                *c = logNormFactor - 2 * ms.mix_log_weight[e];
                ++c;
            }
        }
        return 0;
    }

    template<bool Fuse>
    void scoreFrames(const float* feats, long t0, long t1, float* scores) const {
        float* x = alignedFloats(padded);
        for (long t = t0; t < t1; ++t) {
            std::memset(x, 0, sizeof(float) * padded);
            const float* f = feats + (size_t)t * dim;
            for (unsigned d = 0; d < dim; ++d)
                x[d] = f[d] * isd[d];
            for (unsigned m = 0; m < nMix; ++m) {
                float best = FLT_MAX;
                for (unsigned dns = offsets[m]; dns < offsets[m + 1]; ++dns) {
                    const float* mean = means + (size_t)dns * padded;
                    __m128       s1   = _mm_load_ss(consts + dns);
                    __m128       s2   = _mm_setzero_ps();
                    for (unsigned d = 0; d < padded; d += 8) {
                        __m128 x1 = _mm_sub_ps(_mm_load_ps(mean + d), _mm_load_ps(x + d));
                        __m128 x2 = _mm_sub_ps(_mm_load_ps(mean + d + 4), _mm_load_ps(x + d + 4));
                        if (Fuse) {
                            s1 = _mm_fmadd_ps(x1, x1, s1);
                            s2 = _mm_fmadd_ps(x2, x2, s2);
                        }
                        else {
                            s1 = _mm_add_ps(s1, _mm_mul_ps(x1, x1));
                            s2 = _mm_add_ps(s2, _mm_mul_ps(x2, x2));
                        }
                    }
                    s1 = _mm_add_ps(s1, s2);
                    s2 = s1;
                    s1 = _mm_shuffle_ps(s1, s2, _MM_SHUFFLE(1, 0, 3, 2));
                    s1 = _mm_add_ps(s1, s2);
                    s2 = s1;
                    s1 = _mm_shuffle_ps(s1, s2, _MM_SHUFFLE(2, 3, 0, 1));
                    s1 = _mm_add_ps(s1, s2);
                    _mm_store_ss(&best, _mm_min_ps(_mm_load_ss(&best), s1));
                }
                if (best < FLT_MAX)
                    best *= 0.5;
                scores[(size_t)t * nMix + m] = best;
            }
        }
        free(x);
    }

    /* one density against one (scaled, padded) feature: the arithmetic of fillScoreCacheTpl :225-243 */
    template<bool Fuse>
    float densityScore(unsigned dns, const float* x) const {
        const float* mean = means + (size_t)dns * padded;
        __m128       s1   = _mm_load_ss(consts + dns);
        __m128       s2   = _mm_setzero_ps();
        for (unsigned d = 0; d < padded; d += 8) {
            __m128 x1 = _mm_sub_ps(_mm_load_ps(mean + d), _mm_load_ps(x + d));
            __m128 x2 = _mm_sub_ps(_mm_load_ps(mean + d + 4), _mm_load_ps(x + d + 4));
            if (Fuse) {
                s1 = _mm_fmadd_ps(x1, x1, s1);
                s2 = _mm_fmadd_ps(x2, x2, s2);
            }
            else {
                s1 = _mm_add_ps(s1, _mm_mul_ps(x1, x1));
                s2 = _mm_add_ps(s2, _mm_mul_ps(x2, x2));
            }
        }
        s1 = _mm_add_ps(s1, s2);
        s2 = s1;
        s1 = _mm_shuffle_ps(s1, s2, _MM_SHUFFLE(1, 0, 3, 2));
        s1 = _mm_add_ps(s1, s2);
        s2 = s1;
        s1 = _mm_shuffle_ps(s1, s2, _MM_SHUFFLE(2, 3, 0, 1));
        s1 = _mm_add_ps(s1, s2);
        float r;
        _mm_store_ss(&r, s1);
        return r;
    }
};

/* ------------------------------------------------------------------ Mm::DensityClustering<f32, f32>
 * src/Mm/DensityClustering.{hh,cc,tcc}: k-means over the (scaled, padded) density means -- initializeClusters
 * .tcc:62-75 (srand(1), rand() % nDensities, distinct densities), assignDensities :82-100, updateClusterMeans
 * :103-123 (f64 sums), build :126-161; selectClusters :164-189 (std::sort by distance, the first `select`
 * clusters are active); distances by Mm::unrolledVectorDistance src/Mm/Utilities.hh:254-296 (sequential
 * score += df * df).  Parameters DensityClustering.cc:20-34: clusters 256, select-clusters 32, iterations 5,
 * backoff-score 40000. */
struct DensityClusteringF {
    unsigned              nClusters, nSelected, dim, nDens;
    bool                  fuse;
    std::vector<float>    clusterMeans;
    std::vector<unsigned> clusterOf;

    float distance(const float* a, const float* b) const {
        float score = 0;
        for (unsigned d = 0; d < dim; ++d) {
            float df = a[d] - b[d];
            score    = fuse ? std::fmaf(df, df, score) : (score + df * df);
        }
        return score;
    }
    void build(const float* densities, unsigned dimension, unsigned nDensities, unsigned clusters, unsigned select,
               unsigned iterations, bool useFma) {
        dim       = dimension;
        nDens     = nDensities;
        nClusters = std::min(clusters, nDensities); /* DensityClusteringBase::init, .cc:52-56 */
        nSelected = select;
        fuse      = useFma;
        clusterMeans.assign((size_t)nClusters * dim, 0.0f);
        clusterOf.assign(nDens, 0);
        std::set<unsigned> used;
        srand(1);
        for (unsigned c = 0; c < nClusters; ++c) {
            unsigned pick = 0;
            do {
                pick = rand() % nDens;
            } while (used.count(pick));
            used.insert(pick);
            std::copy(densities + (size_t)pick * dim, densities + (size_t)(pick + 1) * dim, clusterMeans.begin() + (size_t)c * dim);
        }
        for (unsigned it = 0; it < iterations; ++it) {
            std::vector<std::vector<unsigned>> assigned(nClusters);
            for (unsigned dns = 0; dns < nDens; ++dns) {
                float    best = FLT_MAX;
                unsigned bc   = 0;
                for (unsigned c = 0; c < nClusters; ++c) {
                    float dist = distance(&clusterMeans[(size_t)c * dim], densities + (size_t)dns * dim);
                    if (dist < best) {
                        best = dist;
                        bc   = c;
                    }
                }
                clusterOf[dns] = bc;
                assigned[bc].push_back(dns);
            }
            for (unsigned c = 0; c < nClusters; ++c) {
                if (assigned[c].empty())
                    continue;
                std::vector<double> sums(dim, 0.0);
                for (unsigned a : assigned[c])
                    for (unsigned d = 0; d < dim; ++d)
                        sums[d] += densities[(size_t)a * dim + d];
                for (unsigned d = 0; d < dim; ++d)
                    clusterMeans[(size_t)c * dim + d] = sums[d] / (double)assigned[c].size();
            }
        }
    }
    void select(std::vector<char>& active, const float* feature) const {
        std::vector<std::pair<float, unsigned>> byDistance(nClusters);
        for (unsigned c = 0; c < nClusters; ++c)
            byDistance[c] = std::make_pair(distance(feature, &clusterMeans[(size_t)c * dim]), c);
        std::sort(byDistance.begin(), byDistance.end(),
                  [](const std::pair<float, unsigned>& a, const std::pair<float, unsigned>& b) { return a.first < b.first; });
        active.assign(nClusters, 0);
        for (unsigned i = 0; i < nSelected && i < nClusters; ++i)
            active[byDistance[i].second] = 1;
    }
};

/* ------------------------------------------------------------------ Mm::GaussDiagonalMaximumFeatureScorer
 * init src/Mm/GaussDiagonalMaximumFeatureScorer.cc:65-87 (+ MixtureFeatureScorerElement.cc:21-34,
 * CovarianceFeatureScorerElement.cc:21-51), distance :144-233 (SSE3 branch: the reference always
 * builds with -msse3), calculateScoreAndDensity :116-142; Sum variant :239-290. */
struct DiagonalScorer {
    unsigned                        dim;
    const orc_mixture_set*          ms;
    std::vector<std::vector<float>> minus2LogWeights;  // per mixture
    std::vector<std::vector<float>> isd;               // per covariance
    std::vector<float>              logNorm;           // per covariance

    void init(const orc_mixture_set& m, float mixtureWeightScale, float gaussianScaleParam) {
        ms  = &m;
        dim = m.dim;
        /* gaussianScale_ = std::sqrt(paramGaussianScale(c)) :49 -- f64 parameter, f32 member */
        float gaussianScale = std::sqrt((double)gaussianScaleParam);
        minus2LogWeights.resize(m.n_mixtures);
        for (unsigned i = 0; i < m.n_mixtures; ++i) {
            unsigned n = m.mix_offsets[i + 1] - m.mix_offsets[i];
            minus2LogWeights[i].resize(n);
            for (unsigned k = 0; k < n; ++k) {
                float v                = -2 * m.mix_log_weight[m.mix_offsets[i] + k];
                minus2LogWeights[i][k] = v * mixtureWeightScale;
            }
        }
        isd.resize(m.n_covariances);
        logNorm.resize(m.n_covariances);
        for (unsigned c = 0; c < m.n_covariances; ++c) {
            const float* var = m.variances + (size_t)c * dim;
            isd[c].resize(dim);
            for (unsigned d = 0; d < dim; ++d)
                isd[c][d] = inverseSquareRoot(var[d]) * gaussianScale;
            float lnf  = gaussLogNormFactor(var, dim);
            logNorm[c] = lnf * (gaussianScale * gaussianScale);
        }
    }

    template<bool Fuse>
    float distance(const float* feature, const float* mean, const float* isrv) const {
        unsigned cmp = 0;
        float    result = 0;
        __m128   sum    = _mm_setzero_ps();
        unsigned eff    = dim & (~3u);
        while (cmp < eff) {
            __m128 m  = _mm_loadu_ps(mean + cmp);
            __m128 f  = _mm_loadu_ps(feature + cmp);
            __m128 v  = _mm_loadu_ps(isrv + cmp);
            __m128 df = _mm_mul_ps(_mm_sub_ps(m, f), v);
            sum       = Fuse ? _mm_fmadd_ps(df, df, sum) : _mm_add_ps(sum, _mm_mul_ps(df, df));
            cmp += 4u;
        }
        sum = _mm_hadd_ps(sum, sum);
        float buffer[4];
        _mm_storeu_ps(buffer, sum);
        result += buffer[0] + buffer[1];
        for (; cmp < dim; ++cmp) {
            float df = (mean[cmp] - feature[cmp]) * isrv[cmp];
            result   = Fuse ? std::fmaf(df, df, result) : (result + df * df);
        }
        return result;
    }

    template<bool Fuse>
    void scoreMax(const float* feats, long T, float* scores, uint32_t* best) const {
        for (long t = 0; t < T; ++t) {
            const float* x = feats + (size_t)t * dim;
            for (unsigned m = 0; m < ms->n_mixtures; ++m) {
                float    bestScore   = FLT_MAX;
                uint32_t bestDensity = 0xffffffffu;
                unsigned n           = ms->mix_offsets[m + 1] - ms->mix_offsets[m];
                for (unsigned k = 0; k < n; ++k) {
                    unsigned dns = ms->mix_density[ms->mix_offsets[m] + k];
                    unsigned cov = ms->dens_cov[dns];
                    double   score = (double)minus2LogWeights[m][k] + (double)logNorm[cov] +
                                   (double)distance<Fuse>(x, ms->means + (size_t)ms->dens_mean[dns] * dim,
                                                          isd[cov].data());
                    if (bestScore > score) {
                        bestScore   = score;
                        bestDensity = k;
                    }
                }
                scores[(size_t)t * ms->n_mixtures + m] = 0.5 * bestScore;
                if (best)
                    best[(size_t)t * ms->n_mixtures + m] = bestDensity;
            }
        }
    }

    template<bool Fuse>
    void scoreSum(const float* feats, long T, float* scores, uint32_t* best) const {
        std::vector<float> s;
        for (long t = 0; t < T; ++t) {
            const float* x = feats + (size_t)t * dim;
            for (unsigned m = 0; m < ms->n_mixtures; ++m) {
                unsigned n = ms->mix_offsets[m + 1] - ms->mix_offsets[m];
                s.resize(n);
                for (unsigned k = 0; k < n; ++k) {
                    unsigned dns   = ms->mix_density[ms->mix_offsets[m] + k];
                    unsigned cov   = ms->dens_cov[dns];
                    float    score = minus2LogWeights[m][k] + logNorm[cov] +
                                  distance<Fuse>(x, ms->means + (size_t)ms->dens_mean[dns] * dim, isd[cov].data());
                    s[k] = 0.5 * score;
                }
                float    bestScore   = FLT_MAX;
                uint32_t bestDensity = 0xffffffffu;
                for (unsigned k = 0; k < n; ++k)
                    if (bestScore > s[k]) {
                        bestScore   = s[k];
                        bestDensity = k;
                    }
                float sumExp = 0;
                for (unsigned k = 0; k < n; ++k)
                    sumExp += std::exp(bestScore - s[k]);
                scores[(size_t)t * ms->n_mixtures + m] = bestScore - std::log(sumExp);
                if (best)
                    best[(size_t)t * ms->n_mixtures + m] = bestDensity;
            }
        }
    }
};

/* ------------------------------------------------------------------ Mm::BatchIntFeatureScorer
 * ("batch-diagonal-maximum-int"; BatchUnrolledIntFeatureScorer "batch-diagonal-maximum-fast" computes the same
 * numbers with the loop over three 16-byte blocks unrolled, src/Mm/BatchFeatureScorer.cc:581-650)
 * quantize src/Mm/Utilities.hh:190-202, quantizationScale src/Mm/BatchFeatureScorer.cc:355-373, init :375-416,
 * setFeature :418-424, addDistance / horizontalAdd :427-456, fillScoreCacheTpl :458-510.
 * Integer arithmetic is exact, so only the f32/f64 mixing of init() and the final division matter. */
struct BatchInt {
    unsigned                   dim, padded, nMix, nDens;
    std::vector<unsigned>      offsets;
    std::vector<float>         variance; /* isd * scale, zero padded */
    std::vector<int32_t>       consts;
    uint8_t*                   means;
    float                      scale_;

    BatchInt() : means(0), scale_(1) {}
    ~BatchInt() { free(means); }

    static uint8_t quantize(float x) {
        /* clip((int)round(x) + offset), offset = (int)round((255 + 0 + 1) / 2) = 128 */
        int v = (int)roundf(x) + 128;
        return (uint8_t)std::min(std::max(v, 0), 255);
    }

    int init(const orc_mixture_set& ms) {
        if (ms.n_covariances != 1)
            return -2;
        dim    = ms.dim;
        padded = ((dim + 15) / 16) * 16;
        nMix   = ms.n_mixtures;
        offsets.assign(nMix + 1, 0);
        nDens = 0;
        for (unsigned m = 0; m < nMix; ++m) {
            offsets[m] = nDens;
            nDens += ms.mix_offsets[m + 1] - ms.mix_offsets[m];
        }
        offsets[nMix] = nDens;
        variance.assign(padded, 0.0f);
        for (unsigned d = 0; d < dim; ++d)
            variance[d] = inverseSquareRoot(ms.variances[d]);
        /* quantizationScale(): range of mean * isd over ALL densities of the set */
        float minMean = FLT_MAX, maxMean = -FLT_MAX;
        for (unsigned i = 0; i < ms.n_densities; ++i) {
            const float* mu = ms.means + (size_t)ms.dens_mean[i] * dim;
            for (unsigned d = 0; d < dim; ++d) {
                float divided = mu[d] * variance[d];
                minMean       = std::min(minMean, divided);
                maxMean       = std::max(maxMean, divided);
            }
        }
        const int   quantizedIntervalSize = 255 - 0;
        const float intervalSize          = 2 * std::max(std::fabs(minMean), std::fabs(maxMean));
        const float scale                 = static_cast<float>(quantizedIntervalSize) / (1.25 * intervalSize);
        const float scaleSquared          = scale * scale;
        scale_                            = 2.0 * scaleSquared;
        for (unsigned d = 0; d < padded; ++d)
            variance[d] = variance[d] * scale;
        const float logNorm       = gaussLogNormFactor(ms.variances, dim); /* Score logNormalizationFactor_ */
        const float logNormFactor = logNorm * scaleSquared;
        if (posix_memalign((void**)&means, 16, std::max<size_t>((size_t)nDens * padded, 16)) != 0)
            return -4;
        std::memset(means, 0, std::max<size_t>((size_t)nDens * padded, 16));
        consts.assign(nDens, 0);
        for (unsigned m = 0; m < nMix; ++m) {
            uint8_t* mean = means + (size_t)offsets[m] * padded;
            int32_t* c    = consts.data() + offsets[m];
            for (unsigned e = ms.mix_offsets[m]; e < ms.mix_offsets[m + 1]; ++e) {
                unsigned dns = ms.mix_density[e];
                if (ms.dens_cov[dns] != 0)
                    return -3;
                const float* mu = ms.means + (size_t)ms.dens_mean[dns] * dim;
                for (unsigned d = 0; d < dim; ++d)
                    mean[d] = quantize(mu[d] * variance[d]);
                mean += padded;
                *c = static_cast<int32_t>(logNormFactor - scale_ * ms.mix_log_weight[e]);
                ++c;
            }
        }
        return 0;
    }

    void scoreFrames(const float* feats, long t0, long t1, float* scores) const {
        uint8_t* x = 0;
        if (posix_memalign((void**)&x, 16, padded) != 0)
            return;
        for (long t = t0; t < t1; ++t) {
            std::memset(x, 0, padded);
            const float* f = feats + (size_t)t * dim;
            for (unsigned d = 0; d < dim; ++d)
                x[d] = quantize(f[d] * variance[d]);
            for (unsigned m = 0; m < nMix; ++m) {
                int32_t best = 2147483647;
                for (unsigned dns = offsets[m]; dns < offsets[m + 1]; ++dns) {
                    const uint8_t* mean = means + (size_t)dns * padded;
                    __m128i        sum  = _mm_setzero_si128();
                    for (unsigned d = 0; d < padded; d += 16) {
                        __m128i mv = _mm_load_si128((const __m128i*)(mean + d));
                        __m128i xv = _mm_load_si128((const __m128i*)(x + d));
                        /* |m - x| per byte, widened to 16 bit, squared and pair-summed to 32 bit */
                        __m128i ad = _mm_or_si128(_mm_subs_epu8(mv, xv), _mm_subs_epu8(xv, mv));
                        __m128i hi = _mm_unpackhi_epi8(ad, _mm_setzero_si128());
                        __m128i lo = _mm_unpacklo_epi8(ad, _mm_setzero_si128());
                        sum        = _mm_add_epi32(sum, _mm_madd_epi16(hi, hi));
                        sum        = _mm_add_epi32(sum, _mm_madd_epi16(lo, lo));
                    }
                    __m128i s   = _mm_add_epi32(_mm_shuffle_epi32(sum, _MM_SHUFFLE(3, 2, 3, 2)),
                                                _mm_shuffle_epi32(sum, _MM_SHUFFLE(1, 0, 1, 0)));
                    int32_t tmp = _mm_cvtsi128_si32(_mm_add_epi32(s, _mm_shuffle_epi32(s, _MM_SHUFFLE(2, 3, 0, 1))));
                    tmp += consts[dns];
                    if (tmp < best)
                        best = tmp;
                }
                scores[(size_t)t * nMix + m] = static_cast<float>(best) / scale_;
            }
        }
        free(x);
    }
};

/* ------------------------------------------------------------------------------------------------------------------
 * Mm::SimdGaussDiagonalMaximumFeatureScorer ("SIMD-diagonal-maximum", src/Mm/SimdFeatureScorer.{hh,cc} +
 * src/Mm/IntelOptimization.{hh,cc}): u8-quantised means and features like the batch-int scorer, but with one
 * quantised copy of the feature vector PER COVARIANCE (each scaled by that covariance's 1/sqrt(var)), the constant of
 * a density truncated from f32, the best density kept, and the score 0.5 * int / scaling^2 evaluated in double.
 *   init / getScaling / quantizationScalingFactor   SimdFeatureScorer.cc:62-135
 *   buildMixtureTable / createDensityElement        SimdFeatureScorer.cc:79-104, IntelOptimization.cc:39-49
 *   multiplyAndQuantize / quantize                  IntelOptimization.cc:51-69, Utilities.hh:190-202
 *   calculateScoreAndDensity / quantizedScore       SimdFeatureScorer.cc:137-176
 *   distance                                        the code SSE2CodeGenerator.cc emits: sum_d (m_d - x_d)^2 in s32
 * (The reference's cmake build generates 32-bit code on x86-64 and crashes in this scorer; oracle/refbuild compiles
 * the generator with -DPROC_x86_64, as the reference's old Makefiles did, to run it.) */
struct SimdDiagMax {
    unsigned                          dim, nMix, nCov;
    std::vector<std::vector<float>>   isd;     /* [cov][dim] 1/sqrt(var) * scaling */
    std::vector<unsigned>             offsets; /* [nMix + 1] */
    std::vector<uint8_t>              means;   /* [nDens][dim] */
    std::vector<int32_t>              consts;
    std::vector<unsigned>             cov;
    float                             scalingSquared;

    int init(const orc_mixture_set& ms) {
        dim  = ms.dim;
        nMix = ms.n_mixtures;
        nCov = ms.n_covariances;
        isd.assign(nCov, std::vector<float>(dim));
        std::vector<float> logNorm(nCov);
        for (unsigned c = 0; c < nCov; ++c) {
            const float* var = ms.variances + (size_t)c * dim;
            for (unsigned d = 0; d < dim; ++d) {
                if (!(var[d] > 0))
                    return -2; /* require(checkDiagonal(diagonal)) */
                isd[c][d] = inverseSquareRoot(var[d]);
            }
            logNorm[c] = (float)gaussLogNormFactor(var, dim); /* Score logNormalizationFactor_ */
        }
        /* getScaling(): range of mean / sqrt(var) over ALL densities of the set */
        float minMean = FLT_MAX, maxMean = -FLT_MAX; /* Core::Type<f32>::max, ::min = -FLT_MAX */
        for (unsigned i = 0; i < ms.n_densities; ++i) {
            const float* mu = ms.means + (size_t)ms.dens_mean[i] * dim;
            const float* sd = isd[ms.dens_cov[i]].data();
            for (unsigned d = 0; d < dim; ++d) {
                const float divided = mu[d] * sd[d];
                minMean             = std::min(minMean, divided);
                maxMean             = std::max(maxMean, divided);
            }
        }
        const float intervalSize = 2 * std::max(std::fabs(minMean), std::fabs(maxMean));
        const float scaling      = (float)255 / (1.25 * intervalSize);
        scalingSquared           = scaling * scaling;
        for (unsigned c = 0; c < nCov; ++c) { /* CovarianceFeatureScorerElement::scale */
            for (unsigned d = 0; d < dim; ++d)
                isd[c][d] = isd[c][d] * scaling;
            logNorm[c] *= scaling * scaling;
        }
        offsets.assign(nMix + 1, 0);
        for (unsigned m = 0; m < nMix; ++m)
            offsets[m + 1] = offsets[m] + (ms.mix_offsets[m + 1] - ms.mix_offsets[m]);
        means.assign((size_t)offsets[nMix] * dim, 0);
        consts.assign(offsets[nMix], 0);
        cov.assign(offsets[nMix], 0);
        for (unsigned m = 0; m < nMix; ++m) {
            unsigned k = offsets[m];
            for (unsigned e = ms.mix_offsets[m]; e < ms.mix_offsets[m + 1]; ++e, ++k) {
                const unsigned dns = ms.mix_density[e];
                const unsigned c   = ms.dens_cov[dns];
                const float*   mu  = ms.means + (size_t)ms.dens_mean[dns] * dim;
                cov[k]             = c;
                for (unsigned d = 0; d < dim; ++d)
                    means[(size_t)k * dim + d] = BatchInt::quantize(mu[d] * isd[c][d]);
                /* Weight (f64) = f32 * -2 * f64; passed as Score (f32); (s32)(f32 + f32) */
                const double scaledMinus2LogWeight = scalingSquared * -2 * ms.mix_log_weight[e];
                consts[k] = (int32_t)((float)scaledMinus2LogWeight + logNorm[c]);
            }
        }
        return 0;
    }

    void scoreFrames(const float* feats, long t0, long t1, float* scores, uint32_t* best) const {
        std::vector<uint8_t> x((size_t)nCov * dim);
        for (long t = t0; t < t1; ++t) {
            const float* f = feats + (size_t)t * dim;
            for (unsigned c = 0; c < nCov; ++c)
                for (unsigned d = 0; d < dim; ++d)
                    x[(size_t)c * dim + d] = BatchInt::quantize(f[d] * isd[c][d]);
            for (unsigned m = 0; m < nMix; ++m) {
                int      minScore = 2147483647;
                uint32_t bestDns  = 0xffffffffu;
                for (unsigned k = offsets[m]; k < offsets[m + 1]; ++k) {
                    const uint8_t* mean = means.data() + (size_t)k * dim;
                    const uint8_t* xq   = x.data() + (size_t)cov[k] * dim;
                    int            dist = 0;
                    for (unsigned d = 0; d < dim; ++d) {
                        const int df = (int)mean[d] - (int)xq[d];
                        dist += df * df;
                    }
                    const int score = consts[k] + dist;
                    if (score < minScore) {
                        minScore = score;
                        bestDns  = k - offsets[m];
                    }
                }
                scores[(size_t)t * nMix + m] = (float)(0.5 * minScore / scalingSquared);
                if (best)
                    best[(size_t)t * nMix + m] = bestDns;
            }
        }
    }
};

}  // namespace

extern "C" int orc_gmm_batch_int(const orc_mixture_set* ms, const float* feats, long T, float* scores, int n_threads) {
    BatchInt s;
    int      rc = s.init(*ms);
    if (rc)
        return rc;
    if (n_threads < 1)
        n_threads = 1;
    std::vector<std::thread> pool;
    for (int i = 0; i < n_threads; ++i) {
        long a = T * i / n_threads, b = T * (i + 1) / n_threads;
        pool.emplace_back([&s, feats, scores, a, b]() { s.scoreFrames(feats, a, b, scores); });
    }
    for (auto& th : pool)
        th.join();
    return 0;
}

/* scores [T * n_mixtures], best (optional) [T * n_mixtures] density-in-mixture indices */
extern "C" int orc_gmm_simd_diag_max(const orc_mixture_set* ms, const float* feats, long T, float* scores, uint32_t* best,
                                     int n_threads) {
    SimdDiagMax s;
    int         rc = s.init(*ms);
    if (rc)
        return rc;
    if (n_threads < 1)
        n_threads = 1;
    std::vector<std::thread> pool;
    for (int i = 0; i < n_threads; ++i) {
        long a = T * i / n_threads, b = T * (i + 1) / n_threads;
        pool.emplace_back([&s, feats, scores, best, a, b]() { s.scoreFrames(feats, a, b, scores, best); });
    }
    for (auto& th : pool)
        th.join();
    return 0;
}

/* the quantised model (for known-answer tests and the CUDA side's checks): means [n_dens * dim], consts [n_dens],
 * isd [n_cov * dim] (scaled), scaling^2 */
extern "C" int orc_gmm_simd_model(const orc_mixture_set* ms, uint8_t* means, int32_t* consts, float* isd, float* scaling_squared) {
    SimdDiagMax s;
    int         rc = s.init(*ms);
    if (rc)
        return rc;
    if (means)
        std::memcpy(means, s.means.data(), s.means.size());
    if (consts)
        std::memcpy(consts, s.consts.data(), sizeof(int32_t) * s.consts.size());
    if (isd)
        for (unsigned c = 0; c < s.nCov; ++c)
            std::memcpy(isd + (size_t)c * s.dim, s.isd[c].data(), sizeof(float) * s.dim);
    if (scaling_squared)
        *scaling_squared = s.scalingSquared;
    return 0;
}

/* the quantised model as the scorer holds it (for known-answer tests): means [n_dens * padded], consts [n_dens] */
extern "C" int orc_gmm_batch_int_model(const orc_mixture_set* ms, uint8_t* means, int32_t* consts, float* variance,
                                       float* scale, int* padded) {
    BatchInt s;
    int      rc = s.init(*ms);
    if (rc)
        return rc;
    if (means)
        std::memcpy(means, s.means, (size_t)s.nDens * s.padded);
    if (consts)
        std::memcpy(consts, s.consts.data(), sizeof(int32_t) * s.nDens);
    if (variance)
        std::memcpy(variance, s.variance.data(), sizeof(float) * s.padded);
    if (scale)
        *scale = s.scale_;
    if (padded)
        *padded = (int)s.padded;
    return 0;
}

extern "C" int orc_gmm_batch_float(const orc_mixture_set* ms, const float* feats, long T, float* scores, int use_fma) {
    BatchFloat s;
    int        rc = s.init(*ms);
    if (rc)
        return rc;
    if (use_fma)
        s.scoreFrames<true>(feats, 0, T, scores);
    else
        s.scoreFrames<false>(feats, 0, T, scores);
    return 0;
}

/* Mm::BatchPreselectionFloatFeatureScorer src/Mm/BatchFeatureScorer.cc:257-315 ("preselection-batch-float"): only
 * densities of the `select` clusters closest to the feature are scored; a mixture without one gets the back-off
 * score.  cluster_of [n_densities] / cluster_means [n_clusters * padded] are optional outputs. */
extern "C" int orc_gmm_preselect_float(const orc_mixture_set* ms, const float* feats, long T, float* scores, int use_fma,
                                       int clusters, int select, int iterations, float backoff, uint32_t* cluster_of,
                                       float* cluster_means, int* n_clusters) {
    BatchFloat s;
    int        rc = s.init(*ms);
    if (rc)
        return rc;
    DensityClusteringF dc;
    dc.build(s.means, s.padded, s.nDens, (unsigned)clusters, (unsigned)select, (unsigned)iterations, use_fma != 0);
    if ((unsigned)select > dc.nClusters)
        return -4; /* verify(nSelected_ <= nClusters_) */
    if (cluster_of)
        std::copy(dc.clusterOf.begin(), dc.clusterOf.end(), cluster_of);
    if (cluster_means)
        std::copy(dc.clusterMeans.begin(), dc.clusterMeans.end(), cluster_means);
    if (n_clusters)
        *n_clusters = (int)dc.nClusters;
    float*            x = alignedFloats(s.padded);
    std::vector<char> active;
    for (long t = 0; t < T; ++t) {
        std::memset(x, 0, sizeof(float) * s.padded);
        for (unsigned d = 0; d < s.dim; ++d)
            x[d] = feats[(size_t)t * s.dim + d] * s.isd[d];
        dc.select(active, x);
        for (unsigned m = 0; m < s.nMix; ++m) {
            float best = FLT_MAX;
            for (unsigned dns = s.offsets[m]; dns < s.offsets[m + 1]; ++dns) {
                if (!active[dc.clusterOf[dns]])
                    continue;
                const float sc = use_fma ? s.densityScore<true>(dns, x) : s.densityScore<false>(dns, x);
                if (sc < best) /* min_ps(best, sc) */
                    best = sc;
            }
            if (best < FLT_MAX)
                best *= 0.5;
            if (best == FLT_MAX)
                best = backoff;
            scores[(size_t)t * s.nMix + m] = best;
        }
    }
    free(x);
    return 0;
}

/* ---- test hooks for oracle/std_sort_restated.h: sort (key, index) pairs by key only, with the real std::sort and with
 * the restatement; the resulting index permutations must be identical */
extern "C" void orc_sort_pairs(const int32_t* keys, int n, int32_t* perm, int restated) {
    std::vector<std::pair<int32_t, int32_t>> v(n);
    for (int i = 0; i < n; ++i)
        v[i] = std::make_pair(keys[i], i);
    auto less = [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) { return a.first < b.first; };
    if (restated)
        stdsort::sort(v.data(), v.data() + n, less);
    else
        std::sort(v.begin(), v.end(), less);
    for (int i = 0; i < n; ++i)
        perm[i] = v[i].second;
}

extern "C" long orc_sort_heap_calls(void) {
    return stdsort::heapSortCalls();
}

/* An input that drives THIS library's std::sort into its depth limit (the heap sort branch), built with McIlroy's
 * adversary ("A Killer Adversary for Quicksort", 1999): the comparison freezes the values of undecided ("gas")
 * elements as late as possible, always making the pivot candidate the smallest remaining one.  keys [n] receives a
 * permutation of 0..n-1. */
extern "C" void orc_sort_killer(int n, int32_t* keys) {
    std::vector<int32_t> val(n, n); /* n = gas */
    std::vector<int32_t> idx(n);
    for (int i = 0; i < n; ++i)
        idx[i] = i;
    int32_t nsolid = 0, candidate = 0;
    const int32_t gas = n;
    std::sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) {
        if (val[x] == gas && val[y] == gas) {
            if (x == candidate)
                val[x] = nsolid++;
            else
                val[y] = nsolid++;
        }
        if (val[x] == gas)
            candidate = x;
        else if (val[y] == gas)
            candidate = y;
        return val[x] < val[y];
    });
    for (int i = 0; i < n; ++i)
        keys[i] = val[i] == gas ? nsolid++ : val[i];
}

/* Mm::BatchPreselectionIntFeatureScorer ("preselection-batch-int", src/Mm/BatchFeatureScorer.cc:514-577) with
 * Mm::DensityClustering<u8, s32>: like the float variant on the quantised means and features -- s32 distances
 * (unrolledVectorDistance<u8, s32>), cluster means truncated back to u8 (the f64 centroid is assigned to a u8),
 * no back-off score: a mixture without a selected density scores (f32)INT_MAX / scale.  The cluster choice uses the
 * real std::sort (ties!).  cluster_of [n_densities] is an optional output. */
extern "C" int orc_gmm_preselect_int(const orc_mixture_set* ms, const float* feats, long T, float* scores, int clusters,
                                     int select, int iterations, uint32_t* cluster_of, int restated_sort) {
    BatchInt s;
    int      rc = s.init(*ms);
    if (rc)
        return rc;
    const unsigned dim = s.padded, nDens = s.nDens, nClusters = std::min<unsigned>((unsigned)clusters, nDens);
    if ((unsigned)select > nClusters)
        return -4;
    auto distance = [dim](const uint8_t* a, const uint8_t* b) {
        int32_t score = 0;
        for (unsigned d = 0; d < dim; ++d) {
            int32_t df = (int32_t)a[d] - (int32_t)b[d];
            score += df * df;
        }
        return score;
    };
    std::vector<uint8_t>  clusterMeans((size_t)nClusters * dim);
    std::vector<unsigned> clusterOf(nDens, 0);
    std::set<unsigned>    used;
    srand(1);
    for (unsigned c = 0; c < nClusters; ++c) {
        unsigned pick = 0;
        do {
            pick = rand() % nDens;
        } while (used.count(pick));
        used.insert(pick);
        std::memcpy(&clusterMeans[(size_t)c * dim], s.means + (size_t)pick * dim, dim);
    }
    for (int it = 0; it < iterations; ++it) {
        std::vector<std::vector<unsigned>> assigned(nClusters);
        for (unsigned dns = 0; dns < nDens; ++dns) {
            int32_t  best = 2147483647;
            unsigned bc   = 0;
            for (unsigned c = 0; c < nClusters; ++c) {
                int32_t dist = distance(&clusterMeans[(size_t)c * dim], s.means + (size_t)dns * dim);
                if (dist < best) {
                    best = dist;
                    bc   = c;
                }
            }
            clusterOf[dns] = bc;
            assigned[bc].push_back(dns);
        }
        for (unsigned c = 0; c < nClusters; ++c) {
            if (assigned[c].empty())
                continue;
            std::vector<double> sums(dim, 0.0);
            for (unsigned a : assigned[c])
                for (unsigned d = 0; d < dim; ++d)
                    sums[d] += s.means[(size_t)a * dim + d];
            for (unsigned d = 0; d < dim; ++d)
                clusterMeans[(size_t)c * dim + d] = (uint8_t)(sums[d] / (double)assigned[c].size());
        }
    }
    if (cluster_of)
        std::copy(clusterOf.begin(), clusterOf.end(), cluster_of);
    std::vector<uint8_t>                     x(dim);
    std::vector<std::pair<int32_t, unsigned>> byDistance(nClusters);
    std::vector<char>                        active(nClusters);
    auto less = [](const std::pair<int32_t, unsigned>& a, const std::pair<int32_t, unsigned>& b) { return a.first < b.first; };
    for (long t = 0; t < T; ++t) {
        std::fill(x.begin(), x.end(), 0);
        for (unsigned d = 0; d < s.dim; ++d)
            x[d] = BatchInt::quantize(feats[(size_t)t * s.dim + d] * s.variance[d]);
        for (unsigned c = 0; c < nClusters; ++c)
            byDistance[c] = std::make_pair(distance(x.data(), &clusterMeans[(size_t)c * dim]), c);
        if (restated_sort)
            stdsort::sort(byDistance.data(), byDistance.data() + nClusters, less);
        else
            std::sort(byDistance.begin(), byDistance.end(), less);
        std::fill(active.begin(), active.end(), 0);
        for (unsigned i = 0; i < (unsigned)select; ++i)
            active[byDistance[i].second] = 1;
        for (unsigned m = 0; m < s.nMix; ++m) {
            int32_t best = 2147483647;
            for (unsigned dns = s.offsets[m]; dns < s.offsets[m + 1]; ++dns) {
                if (!active[clusterOf[dns]])
                    continue;
                int32_t tmp = distance(s.means + (size_t)dns * dim, x.data()) + s.consts[dns];
                if (tmp < best)
                    best = tmp;
            }
            scores[(size_t)t * s.nMix + m] = static_cast<float>(best) / s.scale_;
        }
    }
    return 0;
}

extern "C" int orc_gmm_batch_float_mt(const orc_mixture_set* ms, const float* feats, long T, float* scores,
                                      int use_fma, int n_threads) {
    BatchFloat s;
    int        rc = s.init(*ms);
    if (rc)
        return rc;
    if (n_threads < 1)
        n_threads = 1;
    std::vector<std::thread> pool;
    for (int i = 0; i < n_threads; ++i) {
        long a = T * i / n_threads, b = T * (i + 1) / n_threads;
        pool.emplace_back([&s, feats, scores, a, b, use_fma]() {
            if (use_fma)
                s.scoreFrames<true>(feats, a, b, scores);
            else
                s.scoreFrames<false>(feats, a, b, scores);
        });
    }
    for (auto& th : pool)
        th.join();
    return 0;
}

extern "C" int orc_gmm_diag_max(const orc_mixture_set* ms, float mixture_weight_scale, float gaussian_scale,
                                const float* feats, long T, float* scores, uint32_t* best, int use_fma) {
    DiagonalScorer s;
    s.init(*ms, mixture_weight_scale, gaussian_scale);
    if (use_fma)
        s.scoreMax<true>(feats, T, scores, best);
    else
        s.scoreMax<false>(feats, T, scores, best);
    return 0;
}

extern "C" int orc_gmm_diag_sum(const orc_mixture_set* ms, float mixture_weight_scale, float gaussian_scale,
                                const float* feats, long T, float* scores, uint32_t* best, int use_fma) {
    DiagonalScorer s;
    s.init(*ms, mixture_weight_scale, gaussian_scale);
    if (use_fma)
        s.scoreSum<true>(feats, T, scores, best);
    else
        s.scoreSum<false>(feats, T, scores, best);
    return 0;
}
