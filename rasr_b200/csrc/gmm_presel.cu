// gmm_presel.cu -- Mm::BatchPreselectionFloatFeatureScorer ("preselection-batch-float",
// src/Mm/BatchFeatureScorer.cc:257-315) with Mm::DensityClustering<f32, f32> (src/Mm/DensityClustering.{hh,cc,tcc}).
//
// The reference's CPU trick: cluster the density means once (k-means, 256 clusters), and per frame score only the
// densities of the 32 clusters nearest to the feature vector; a mixture left without a scored density gets the
// back-off score.  A GPU does not need the saving -- dense scoring is the fast path -- but a system tuned with this
// scorer expects ITS scores, so the approximation is reproduced: same clustering (same pseudo-random initialisation,
// same f32 distances and f64 centroid sums), same per-frame cluster choice, same per-density arithmetic and minimum.
//
//   host (create)           k-means on the scaled, padded means: restatement of initializeClusters / assignDensities /
//                           updateClusterMeans; rand() is restated too (glibc's additive-feedback generator) so that
//                           creating a scorer does not disturb the host program's random state
//   presel_select_kernel    one warp per frame: f32 distances to all clusters (sequential sum over the padded
//                           dimension like unrolledVectorDistance), then the `select` smallest by a radix select on the
//                           distance bits -> one bit per cluster.  The reference sorts (distance, cluster) pairs by
//                           distance only (std::sort, not stable): when more clusters sit exactly at the boundary
//                           distance than there are places left (duplicate centroids, for instance), which of them
//                           survive is decided by libstdc++'s introsort -- lane 0 then runs introsort.cuh on the
//                           frame's pairs, as in gmm_presel_int.cu.
//   presel_score_kernel     one block per frame: the densities whose cluster bit is set are collected in a shared list
//                           and scored by groups of 8 lanes with the lane arithmetic of fillScoreCacheTpl (two 4-lane
//                           accumulators over 8-dimension blocks, the constant first in lane 0, horizontal add);
//                           minimum per mixture by atomicMin on order-preserving keys, x0.5, back-off.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <set>
#include <vector>

#include "common.cuh"
#include "introsort.cuh"

namespace {

// glibc's rand() after srand(seed) (TYPE_3 additive feedback generator, r[i] = r[i-3] + r[i-31], output >> 1)
struct GlibcRand {
    std::vector<uint32_t> r;
    size_t                k;
    explicit GlibcRand(uint32_t seed) : r(34), k(0) {
        r[0] = seed ? seed : 1;
        for (int i = 1; i < 31; ++i) {
            const int64_t hi = (int32_t)r[i - 1] / 127773, lo = (int32_t)r[i - 1] % 127773;
            int64_t       w  = 16807 * lo - 2836 * hi;
            if (w < 0)
                w += 2147483647;
            r[i] = (uint32_t)w;
        }
        for (int i = 31; i < 34; ++i)
            r[i] = r[i - 31];
        for (int i = 34; i < 344; ++i)
            r.push_back(r[i - 31] + r[i - 3]);
    }
    int next() {
        const size_t i = r.size();
        r.push_back(r[i - 31] + r[i - 3]);
        return (int)(r.back() >> 1);
    }
};

struct PreselParams {
    const float*   feats;      // [T * dim]
    const float*   isd;        // [padded]
    const float*   means;      // [nDens * padded] scaled means
    const float*   consts;     // [nDens]
    const uint32_t* offsets;   // [nMix + 1]
    const uint8_t* clusterOf;  // [nDens]
    const uint32_t* densMix;   // [nDens] mixture of every density entry
    const float*   clusterMeans;  // [nClusters * padded]
    uint32_t*      active;     // [T * 8] one bit per cluster
    float*         scores;     // [T * nMix]
    long           T;
    int            dim, padded, nMix, nClusters, nSelected, fuse;
    float          backoff;
};

__device__ __forceinline__ float sq_acc(float df, float acc, bool fuse) {
    return fuse ? __fmaf_rn(df, df, acc) : __fadd_rn(acc, __fmul_rn(df, df));
}

constexpr int kSelWarps = 8;

struct DistCluster {  // std::pair<f32 distance, u32 cluster>; distances are >= 0, their bit patterns order like the values
    uint32_t distBits;
    uint32_t cluster;
};
struct ByDistance {
    __host__ __device__ __forceinline__ bool operator()(const DistCluster& a, const DistCluster& b) const {
        return a.distBits < b.distBits;
    }
};

// float offset of the (distance, cluster) pairs in the select kernel's dynamic shared memory (8-byte aligned)
__host__ __device__ __forceinline__ size_t presel_pairs_offset(int nClusters, int padded) {
    return ((size_t)nClusters * (padded + 1) + (size_t)kSelWarps * padded + 1) & ~(size_t)1;
}

// dynamic smem: cluster means [nClusters][padded + 1] | per warp: x [padded] | per warp: pairs [256]
__global__ void __launch_bounds__(kSelWarps * 32) presel_select_kernel(const PreselParams p) {
    extern __shared__ __align__(8) float smem[];
    const int stride = p.padded + 1;  // odd stride: lanes reading different clusters hit different banks
    float*    cm     = smem;
    float*    xs     = smem + (size_t)p.nClusters * stride + (threadIdx.x >> 5) * p.padded;
    DistCluster* pairs =
            reinterpret_cast<DistCluster*>(smem + presel_pairs_offset(p.nClusters, p.padded)) + (threadIdx.x >> 5) * 256;
    for (int i = threadIdx.x; i < p.nClusters * p.padded; i += blockDim.x)
        cm[(i / p.padded) * stride + i % p.padded] = p.clusterMeans[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (long t = (long)blockIdx.x * kSelWarps + (threadIdx.x >> 5); t < p.T; t += (long)gridDim.x * kSelWarps) {
        for (int d = lane; d < p.padded; d += 32)
            xs[d] = d < p.dim ? __fmul_rn(p.feats[t * p.dim + d], p.isd[d]) : 0.0f;
        __syncwarp();
        // distances: cluster c = lane + 32 k
        // (the eight distances of a lane advance together: eight independent accumulation chains, one read of the
        // feature element per dimension; every chain keeps the reference's sequential order over the dimensions)
        uint32_t     key[8];
        float        score[8];
        const float* m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            score[k] = 0.0f;
            m[k]     = cm + min(lane + 32 * k, p.nClusters - 1) * stride;
        }
        for (int d = 0; d < p.padded; ++d) {
            const float xd = xs[d];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                score[k] = sq_acc(__fsub_rn(xd, m[k][d]), score[k], p.fuse);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = lane + 32 * k;
            key[k]      = 0xffffffffu;
            if (c < p.nClusters) {
                key[k]   = __float_as_uint(score[k]);  // distances are >= 0: their bit patterns order like the values
                pairs[c] = DistCluster{key[k], (uint32_t)c};
            }
        }
        // radix select of the nSelected-th smallest key
        uint32_t prefix = 0, mask = 0;
        int      remaining = p.nSelected;
        for (int bit = 31; bit >= 0; --bit) {
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                cnt += ((key[k] & mask) == prefix && !((key[k] >> bit) & 1u) && lane + 32 * k < p.nClusters) ? 1 : 0;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (remaining > cnt) {
                remaining -= cnt;
                prefix |= 1u << bit;
            }
            mask |= 1u << bit;
        }
        int atBoundary = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            atBoundary += (lane + 32 * k < p.nClusters && key[k] == prefix) ? 1 : 0;
        atBoundary = __reduce_add_sync(0xffffffffu, atBoundary);
        if (atBoundary == remaining) {  // warp-uniform: everything up to the boundary distance is selected
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t word = __ballot_sync(0xffffffffu, lane + 32 * k < p.nClusters && key[k] <= prefix);
                if (lane == 0)
                    p.active[t * 8 + k] = word;
            }
        }
        else {  // more candidates at the boundary distance than places: the reference's sort decides
            __syncwarp();
            if (lane == 0) {
                rb::introsort::sort<18>(pairs, pairs + p.nClusters, ByDistance());
                uint32_t words[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int i = 0; i < p.nSelected; ++i)
                    words[pairs[i].cluster >> 5] |= 1u << (pairs[i].cluster & 31);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    p.active[t * 8 + k] = words[k];
            }
        }
        __syncwarp();
    }
}

// order-preserving map of f32 onto u32 (and back), so that the minimum over a mixture's scored densities can be taken
// with an integer atomicMin: the minimum of a set does not depend on the order its elements arrive in
__device__ __forceinline__ uint32_t presel_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float presel_unkey(uint32_t k) {
    return __uint_as_float(k ^ ((k & 0x80000000u) ? 0x80000000u : 0xffffffffu));
}

constexpr int kListCap = 2048;  // densities looked at per round (8 per thread)

// One block per frame.  Round by round, 2048 densities are tested against the frame's cluster bits and the active ones
// collected in a shared list; the list is then scored by groups of 8 lanes -- lane j of a group owns accumulator lane
// j of fillScoreCacheTpl's two 4-lane registers, so a group reads 32 contiguous bytes of the density row per
// 8-dimension block -- and the horizontal adds are the reference's two shuffle-adds.
// dynamic smem: best [nMix] keys
__global__ void __launch_bounds__(256) presel_score_kernel(const PreselParams p) {
    extern __shared__ uint32_t sBest[];
    __shared__ float    xs[128];
    __shared__ uint32_t act[8];
    __shared__ uint32_t list[kListCap];
    __shared__ int      count;
    const int nDens = (int)p.offsets[p.nMix];
    const int lane8 = threadIdx.x & 7, group = threadIdx.x >> 3;
    for (long t = blockIdx.x; t < p.T; t += gridDim.x) {
        __syncthreads();
        for (int d = threadIdx.x; d < p.padded; d += blockDim.x)
            xs[d] = d < p.dim ? __fmul_rn(p.feats[t * p.dim + d], p.isd[d]) : 0.0f;
        if (threadIdx.x < 8)
            act[threadIdx.x] = p.active[t * 8 + threadIdx.x];
        for (int m = threadIdx.x; m < p.nMix; m += blockDim.x)
            sBest[m] = presel_key(FLT_MAX);
        for (int base = 0; base < nDens; base += kListCap) {
            if (threadIdx.x == 0)
                count = 0;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kListCap / 256; ++j) {
                const int dns = base + j * 256 + threadIdx.x;
                if (dns < nDens) {
                    const uint32_t c = p.clusterOf[dns];
                    if ((act[c >> 5] >> (c & 31)) & 1u)
                        list[atomicAdd(&count, 1)] = (uint32_t)dns;
                }
            }
            __syncthreads();
            const int n = count;
            for (int i0 = 0; i0 < n; i0 += 32) {  // warp-uniform trip count: the shuffles below need every lane
                const int      i     = i0 + group;
                const bool     valid = i < n;
                const uint32_t dns   = list[valid ? i : 0];
                const float*   mu    = p.means + (size_t)dns * p.padded + lane8;
                float          acc   = lane8 == 0 ? p.consts[dns] : 0.0f;
                for (int d = 0; d < p.padded; d += 8)
                    acc = sq_acc(__fsub_rn(mu[d], xs[d + lane8]), acc, p.fuse);
                // s1 + s2; s1 += shuffle(1,0,3,2); s1 += shuffle(2,3,0,1): lane 0 = (v1 + v3) + (v0 + v2)
                float v = __fadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, 4, 8));
                v       = __fadd_rn(__shfl_down_sync(0xffffffffu, v, 2, 8), v);
                v       = __fadd_rn(__shfl_down_sync(0xffffffffu, v, 1, 8), v);
                if (valid && lane8 == 0)
                    atomicMin(&sBest[p.densMix[dns]], presel_key(v));
            }
        }
        __syncthreads();
        for (int m = threadIdx.x; m < p.nMix; m += blockDim.x) {
            float best = presel_unkey(sBest[m]);
            if (best < FLT_MAX)
                best = __fmul_rn(best, 0.5f);
            if (best == FLT_MAX)
                best = p.backoff;
            p.scores[t * p.nMix + m] = best;
        }
    }
}

}  // namespace

struct rb_gmm_presel {
    rb::DeviceInfo dev;
    int            dim = 0, padded = 0, nMix = 0, nDens = 0, nClusters = 0, nSelected = 0;
    bool           fuse = true;
    float          backoff = 40000.0f;
    std::vector<float>    isd, means, consts, clusterMeans;
    std::vector<uint32_t> offsets, clusterOf;
    rb::DevBuf<float>     dIsd, dMeans, dConsts, dClusterMeans;
    rb::DevBuf<uint32_t>  dOffsets, dActive, dDensMix;
    rb::DevBuf<uint8_t>   dClusterOf;
};

namespace {

// Mm::unrolledVectorDistance<f32, f32> (src/Mm/Utilities.hh:254-296): sequential score += df * df
float host_distance(const float* a, const float* b, int dim, bool fuse) {
    float score = 0;
    for (int d = 0; d < dim; ++d) {
        const float df = a[d] - b[d];
        if (fuse)
            score = std::fmaf(df, df, score);
        else {
            volatile float sq = df * df;  // keep the product rounded on its own
            score             = score + sq;
        }
    }
    return score;
}

// DensityClustering::build (.tcc:126-161) with initializeClusters :62-75, assignDensities :82-100,
// updateClusterMeans :103-123
void build_clustering(rb_gmm_presel* h, int clusters, int iterations) {
    const int dim = h->padded, nDens = h->nDens;
    h->nClusters  = std::min(clusters, nDens);  // DensityClusteringBase::init, .cc:52-56
    h->clusterMeans.assign((size_t)h->nClusters * dim, 0.0f);
    h->clusterOf.assign(nDens, 0);
    std::set<uint32_t> used;
    GlibcRand          rng(1);
    for (int c = 0; c < h->nClusters; ++c) {
        uint32_t pick = 0;
        do {
            pick = (uint32_t)(rng.next() % nDens);
        } while (used.count(pick));
        used.insert(pick);
        std::copy(h->means.begin() + (size_t)pick * dim, h->means.begin() + (size_t)(pick + 1) * dim,
                  h->clusterMeans.begin() + (size_t)c * dim);
    }
    for (int it = 0; it < iterations; ++it) {
        std::vector<std::vector<uint32_t>> assigned(h->nClusters);
        for (int dns = 0; dns < nDens; ++dns) {
            float best = FLT_MAX;
            int   bc   = 0;
            for (int c = 0; c < h->nClusters; ++c) {
                const float dist = host_distance(&h->clusterMeans[(size_t)c * dim], &h->means[(size_t)dns * dim], dim, h->fuse);
                if (dist < best) {
                    best = dist;
                    bc   = c;
                }
            }
            h->clusterOf[dns] = (uint32_t)bc;
            assigned[bc].push_back((uint32_t)dns);
        }
        for (int c = 0; c < h->nClusters; ++c) {
            if (assigned[c].empty())
                continue;
            std::vector<double> sums(dim, 0.0);
            for (uint32_t a : assigned[c])
                for (int d = 0; d < dim; ++d)
                    sums[d] += h->means[(size_t)a * dim + d];
            for (int d = 0; d < dim; ++d)
                h->clusterMeans[(size_t)c * dim + d] = (float)(sums[d] / (double)assigned[c].size());
        }
    }
}

int upload_clustering(rb_gmm_presel* h, cudaStream_t s) {
    std::vector<uint8_t> c8(h->clusterOf.begin(), h->clusterOf.end());
    RB_CHECK(h->dClusterOf.upload(c8.data(), c8.size(), s));
    RB_CHECK(h->dClusterMeans.upload(h->clusterMeans, s));
    RB_CUDA(cudaStreamSynchronize(s));
    return RB_OK;
}

}  // namespace

int rb_gmm_presel_configure(rb_gmm_presel* h, int clusters, int select, int iterations, float backoff, cudaStream_t s) {
    RB_REQUIRE(clusters >= 1 && clusters <= 256 && select >= 1 && iterations >= 0, "bad preselection parameters");
    build_clustering(h, clusters, iterations);
    RB_REQUIRE(select <= h->nClusters, "select-clusters %d exceeds the %d clusters", select, h->nClusters);
    h->nSelected = select;
    h->backoff   = backoff;
    return upload_clustering(h, s);
}

// model preparation = BatchFloatFeatureScorer::init (src/Mm/BatchFeatureScorer.cc:164-197)
int rb_gmm_presel_create(const rb_mixture_set* ms, bool fuse, const rb::DeviceInfo& dev, cudaStream_t stream,
                         rb_gmm_presel** out) {
    *out = nullptr;
    if (ms->n_covariances != 1) {
        rb::set_error("feature scorer supports only globally pooled covariance");
        return RB_ERR_UNSUPPORTED;
    }
    if (ms->dim > 120) {
        rb::set_error("preselection scorer supports feature dimension <= 120 (got %u)", ms->dim);
        return RB_ERR_UNSUPPORTED;
    }
    rb_gmm_presel* h = new (std::nothrow) rb_gmm_presel();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    h->dev    = dev;
    h->fuse   = fuse;
    h->dim    = (int)ms->dim;
    h->padded = (h->dim + 7) / 8 * 8;
    h->nMix   = (int)ms->n_mixtures;
    h->nDens  = (int)ms->mix_offsets[ms->n_mixtures];
    h->isd.assign(h->padded, 0.0f);
    double sumLog = 0;
    for (int d = 0; d < h->dim; ++d) {
        h->isd[d] = (float)1 / (float)std::sqrt(ms->variances[d]);
        sumLog += std::log(std::fabs(ms->variances[d]));
    }
    const float logNormFactor = (float)((double)h->dim * std::log((double)2 * M_PI) + sumLog);
    h->means.assign((size_t)h->nDens * h->padded, 0.0f);
    h->consts.assign(h->nDens, 0.0f);
    h->offsets.assign(ms->mix_offsets, ms->mix_offsets + ms->n_mixtures + 1);
    for (uint32_t e = 0; e < (uint32_t)h->nDens; ++e) {
        const uint32_t dns = ms->mix_density[e];
        if (ms->dens_cov[dns] != 0) {
            delete h;
            rb::set_error("feature scorer supports only globally pooled covariance");
            return RB_ERR_UNSUPPORTED;
        }
        const float* mu = ms->means + (size_t)ms->dens_mean[dns] * h->dim;
        for (int d = 0; d < h->dim; ++d)
            h->means[(size_t)e * h->padded + d] = mu[d] * h->isd[d];
        h->consts[e] = (float)(logNormFactor - 2 * ms->mix_log_weight[e]);
    }
    int rc = h->dIsd.upload(h->isd, stream);
    if (rc == RB_OK)
        rc = h->dMeans.upload(h->means, stream);
    if (rc == RB_OK)
        rc = h->dConsts.upload(h->consts, stream);
    if (rc == RB_OK)
        rc = h->dOffsets.upload(h->offsets, stream);
    std::vector<uint32_t> densMix(h->nDens);
    for (int m = 0; m < h->nMix; ++m)
        for (uint32_t e = h->offsets[m]; e < h->offsets[m + 1]; ++e)
            densMix[e] = (uint32_t)m;
    if (rc == RB_OK)
        rc = h->dDensMix.upload(densMix, stream);
    if (rc == RB_OK && cudaStreamSynchronize(stream) != cudaSuccess)
        rc = RB_ERR_CUDA;
    if (rc == RB_OK)  // DensityClustering.cc:20-34: clusters 256, select-clusters 32, iterations 5, backoff-score 40000
        rc = rb_gmm_presel_configure(h, 256, std::min(32, std::min(256, h->nDens)), 5, 40000.0f, stream);
    if (rc != RB_OK) {
        delete h;
        return rc;
    }
    *out = h;
    return RB_OK;
}

void rb_gmm_presel_destroy(rb_gmm_presel* h) {
    delete h;
}

int rb_gmm_presel_score(rb_gmm_presel* h, const float* dFeats, long T, float* dScores, cudaStream_t s) {
    RB_CHECK(h->dActive.reserve((size_t)T * 8));
    PreselParams p;
    p.feats        = dFeats;
    p.isd          = h->dIsd.p;
    p.means        = h->dMeans.p;
    p.consts       = h->dConsts.p;
    p.offsets      = h->dOffsets.p;
    p.clusterOf    = h->dClusterOf.p;
    p.densMix      = h->dDensMix.p;
    p.clusterMeans = h->dClusterMeans.p;
    p.active       = h->dActive.p;
    p.scores       = dScores;
    p.T            = T;
    p.dim          = h->dim;
    p.padded       = h->padded;
    p.nMix         = h->nMix;
    p.nClusters    = h->nClusters;
    p.nSelected    = h->nSelected;
    p.fuse         = h->fuse ? 1 : 0;
    p.backoff      = h->backoff;
    const size_t smem = presel_pairs_offset(h->nClusters, h->padded) * 4 + (size_t)kSelWarps * 256 * sizeof(DistCluster);
    RB_CUDA(cudaFuncSetAttribute(presel_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid1 = (int)std::min<long>((T + kSelWarps - 1) / kSelWarps, (long)h->dev.sm_count * 4);
    presel_select_kernel<<<grid1, kSelWarps * 32, smem, s>>>(p);
    RB_LAUNCH_CHECK();
    const size_t smemBest = (size_t)h->nMix * 4;
    RB_REQUIRE(smemBest + 12 * 1024 <= h->dev.smem_optin, "preselection scorer: %d mixtures do not fit", h->nMix);
    RB_CUDA(cudaFuncSetAttribute(presel_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBest));
    const int grid2 = (int)std::min<long>(T, (long)h->dev.sm_count * 16);
    presel_score_kernel<<<grid2, 256, smemBest, s>>>(p);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

// The cluster selection alone (presel_select_kernel): *active = [T * 8] words, one bit per cluster; plus the tables a
// consumer needs to turn it into candidate sets (gmm.cu: the refinement kernel of the exact batch-float route)
int rb_gmm_presel_select(rb_gmm_presel* h, const float* dFeats, long T, const uint32_t** active,
                         const uint8_t** clusterOf, const uint32_t** offsets, float* backoff, cudaStream_t s) {
    RB_CHECK(h->dActive.reserve((size_t)T * 8));
    PreselParams p;
    p.feats        = dFeats;
    p.isd          = h->dIsd.p;
    p.means        = h->dMeans.p;
    p.consts       = h->dConsts.p;
    p.offsets      = h->dOffsets.p;
    p.clusterOf    = h->dClusterOf.p;
    p.densMix      = h->dDensMix.p;
    p.clusterMeans = h->dClusterMeans.p;
    p.active       = h->dActive.p;
    p.scores       = nullptr;
    p.T            = T;
    p.dim          = h->dim;
    p.padded       = h->padded;
    p.nMix         = h->nMix;
    p.nClusters    = h->nClusters;
    p.nSelected    = h->nSelected;
    p.fuse         = h->fuse ? 1 : 0;
    p.backoff      = h->backoff;
    const size_t smem = presel_pairs_offset(h->nClusters, h->padded) * 4 + (size_t)kSelWarps * 256 * sizeof(DistCluster);
    RB_CUDA(cudaFuncSetAttribute(presel_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid1 = (int)std::min<long>((T + kSelWarps - 1) / kSelWarps, (long)h->dev.sm_count * 4);
    presel_select_kernel<<<grid1, kSelWarps * 32, smem, s>>>(p);
    RB_LAUNCH_CHECK();
    *active    = h->dActive.p;
    *clusterOf = h->dClusterOf.p;
    *offsets   = h->dOffsets.p;
    *backoff   = h->backoff;
    return RB_OK;
}

void rb_gmm_presel_clustering(const rb_gmm_presel* h, uint32_t* cluster_of, float* cluster_means, int* n_clusters) {
    if (cluster_of)
        std::copy(h->clusterOf.begin(), h->clusterOf.end(), cluster_of);
    if (cluster_means)
        std::copy(h->clusterMeans.begin(), h->clusterMeans.end(), cluster_means);
    if (n_clusters)
        *n_clusters = h->nClusters;
}

// test hook (host only): the first n values of rand() after srand(seed), as restated above
extern "C" void rb_test_glibc_rand(unsigned seed, int n, int* out) {
    GlibcRand r(seed);
    for (int i = 0; i < n; ++i)
        out[i] = r.next();
}
