// gmm_presel_int.cu -- Mm::BatchPreselectionIntFeatureScorer ("preselection-batch-int",
// src/Mm/BatchFeatureScorer.cc:514-577) with Mm::DensityClustering<u8, s32> (src/Mm/DensityClustering.{hh,cc,tcc}).
//
// The u8 / s32 sibling of gmm_presel.cu: the model is quantised like Mm::BatchIntFeatureScorer's
// (src/Mm/BatchFeatureScorer.cc:355-424, see gmm_int.cu), the quantised means are clustered once on the host (k-means
// in s32 distances, centroids truncated back to u8), per frame the `select` nearest clusters are chosen and only their
// densities scored.  There is no back-off score here: a mixture without a scored density gets (f32)INT_MAX / scale_.
//
// The cluster choice is the delicate part.  s32 distances between u8 vectors tie often, the reference sorts
// (distance, cluster) pairs by distance only with std::sort and keeps the first `select`: which of the clusters tied
// at the boundary survive is decided by libstdc++'s introsort.  introsort.cuh restates that algorithm; one lane per
// frame runs it on the frame's pairs in shared memory.
//
//   presel_int_select_kernel   one warp per frame: quantise the feature (also stored for the score kernel), s32
//                              distances to all clusters (lanes over clusters, 4 dimensions per dp4a), then a radix
//                              select of the `select`-th smallest distance d*.  Clusters nearer than d* are in, farther
//                              ones are out whatever the sort does; only when more clusters sit exactly at d* than
//                              there are places left does the permutation matter, and only then lane 0 runs the
//                              introsort.  One bit per selected cluster.
//   presel_int_score_kernel    one block per frame.  The densities are stored cluster by cluster (the clustering is
//                              static), so the block walks the selected clusters' rows only: a warp per cluster, a
//                              lane per density (dp4a on |m - x| bytes), minimum per mixture by atomicMin in shared
//                              memory (a minimum does not depend on the order), (f32)best / scale_ correctly rounded
#include <climits>
#include <cmath>
#include <cstring>
#include <set>
#include <vector>

#include "common.cuh"
#include "introsort.cuh"

namespace {

struct PreselIntParams {
    const float*    feats;         // [T * dim]
    const float*    variance;      // [dim] inverse standard deviation * quantisation scale
    // densities in cluster order (all densities of cluster 0, then of cluster 1, ...)
    const uint32_t* means;         // [nDens * words] u8 means, 4 per word, zero padded
    const int*      consts;        // [nDens]
    const uint32_t* densMix;       // [nDens] mixture of the density
    const uint32_t* clusterStart;  // [nClusters + 1] first row of every cluster
    const uint32_t* clusterMeans;  // [nClusters * words]
    uint32_t*       xq;            // [T * words] quantised features
    uint32_t*       active;        // [T * 8]
    float*          scores;        // [T * nMix]
    long            T;
    int             dim, words, nMix, nDens, nClusters, nSelected;
    float           scale;  // scale_ = 2 * quantisation scale^2
};

struct DistCluster {  // std::pair<s32 distance, u32 cluster>
    int      dist;
    uint32_t cluster;
};
static_assert(sizeof(DistCluster) == 8, "pair layout");

struct ByDistance {
    __host__ __device__ __forceinline__ bool operator()(const DistCluster& a, const DistCluster& b) const {
        return a.dist < b.dist;
    }
};

// Mm::quantize<f32, u8> (src/Mm/Utilities.hh:190-202): clip((int)round(x) + 128) to [0, 255]
__device__ __forceinline__ uint32_t quantize_u8_dev(float x) {
    const float r = roundf(x);
    // (int) of an out-of-range or NaN float is INT_MIN on the reference's x86 (cvttss2si)
    const int i = fabsf(r) < 2147483648.0f ? __float2int_rz(r) : INT_MIN;
    return (uint32_t)min(max(i + (i < INT_MAX - 128 ? 128 : 0), 0), 255);
}

// sum over 4 byte lanes of (a - b)^2
__device__ __forceinline__ int sq_diff4(uint32_t a, uint32_t b, int acc) {
    const uint32_t ad = __vabsdiffu4(a, b);
    return (int)__dp4a(ad, ad, (uint32_t)acc);
}

constexpr int kSelWarps = 8;

// dynamic smem (u32): cluster means [nClusters][rowStride] | per warp: x [words] | per warp: pairs [256] (8 B each)
__global__ void __launch_bounds__(kSelWarps * 32) presel_int_select_kernel(const PreselIntParams p) {
    extern __shared__ __align__(8) uint32_t smemU[];
    __shared__ uint32_t bits[kSelWarps][8];
    const int    rowStride = p.words | 1;  // odd stride: lanes reading different clusters hit different banks
    const int    warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t*    cm    = smemU;
    size_t       off   = ((size_t)p.nClusters * rowStride + 1) & ~(size_t)1;
    uint32_t*    xs    = smemU + off + (size_t)warp * p.words;
    off                = (off + (size_t)kSelWarps * p.words + 1) & ~(size_t)1;
    DistCluster* pairs = reinterpret_cast<DistCluster*>(smemU + off) + (size_t)warp * 256;
    for (int i = threadIdx.x; i < p.nClusters * p.words; i += blockDim.x)
        cm[(i / p.words) * rowStride + i % p.words] = p.clusterMeans[i];
    __syncthreads();
    for (long t = (long)blockIdx.x * kSelWarps + warp; t < p.T; t += (long)gridDim.x * kSelWarps) {
        // setFeature (src/Mm/BatchFeatureScorer.cc:418-424): u8 = quantize(f * variance), padding stays 0
        for (int w = lane; w < p.words; w += 32) {
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int d = w * 4 + j;
                if (d < p.dim)
                    packed |= quantize_u8_dev(__fmul_rn(p.feats[t * p.dim + d], p.variance[d])) << (8 * j);
            }
            xs[w]                 = packed;
            p.xq[t * p.words + w] = packed;
        }
        if (lane < 8)
            bits[warp][lane] = 0;
        __syncwarp();
        uint32_t key[8];  // cluster lane + 32 k
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = lane + 32 * k;
            key[k]      = 0xffffffffu;
            if (c < p.nClusters) {
                const uint32_t* m    = cm + c * rowStride;
                int             dist = 0;
                for (int w = 0; w < p.words; ++w)
                    dist = sq_diff4(xs[w], m[w], dist);
                key[k]   = (uint32_t)dist;  // < 2^24: at most 128 dimensions of at most 255^2
                pairs[c] = DistCluster{dist, (uint32_t)c};
            }
        }
        // radix select of the nSelected-th smallest distance
        uint32_t prefix = 0, mask = 0xff000000u;
        int      remaining = p.nSelected;
        for (int bit = 23; bit >= 0; --bit) {
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                cnt += ((key[k] & mask) == prefix && !((key[k] >> bit) & 1u)) ? 1 : 0;
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
            atBoundary += key[k] == prefix ? 1 : 0;
        atBoundary = __reduce_add_sync(0xffffffffu, atBoundary);
        if (atBoundary == remaining) {  // warp-uniform: every cluster at the boundary distance is selected
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t word = __ballot_sync(0xffffffffu, key[k] <= prefix);
                if (lane == 0)
                    bits[warp][k] = word;
            }
        }
        else {  // more candidates at the boundary than places: the reference's sort decides
            __syncwarp();
            if (lane == 0)
                rb::introsort::sort<18>(pairs, pairs + p.nClusters, ByDistance());
            __syncwarp();
            for (int i = lane; i < p.nSelected; i += 32) {
                const uint32_t c = pairs[i].cluster;
                atomicOr(&bits[warp][c >> 5], 1u << (c & 31));
            }
        }
        __syncwarp();
        if (lane < 8)
            p.active[t * 8 + lane] = bits[warp][lane];
        __syncwarp();
    }
}

// dynamic smem: best [nMix] (s32)
__global__ void __launch_bounds__(256) presel_int_score_kernel(const PreselIntParams p) {
    extern __shared__ int sBest[];
    __shared__ uint32_t xs[32];
    __shared__ uint32_t act[8];
    __shared__ uint32_t sel[256];  // the selected clusters, ascending
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long t = blockIdx.x; t < p.T; t += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < p.words)
            xs[threadIdx.x] = p.xq[t * p.words + threadIdx.x];
        if (threadIdx.x < 8)
            act[threadIdx.x] = p.active[t * 8 + threadIdx.x];
        for (int m = threadIdx.x; m < p.nMix; m += blockDim.x)
            sBest[m] = INT_MAX;
        __syncthreads();
        {  // thread c owns cluster c: its place in the list = number of selected clusters below it
            const uint32_t word = act[warp];
            if ((word >> lane) & 1u) {
                int pos = __popc(word & ((1u << lane) - 1u));
                for (int k = 0; k < warp; ++k)
                    pos += __popc(act[k]);
                sel[pos] = (uint32_t)threadIdx.x;
            }
        }
        int nSel = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            nSel += __popc(act[k]);
        __syncthreads();
        for (int j = warp; j < nSel; j += 8) {
            const uint32_t c   = sel[j];
            const int      end = (int)p.clusterStart[c + 1];
            for (int row = (int)p.clusterStart[c] + lane; row < end; row += 32) {
                const uint32_t* mu   = p.means + (size_t)row * p.words;
                int             dist = 0;
                for (int w = 0; w < p.words; w += 4) {  // words is a multiple of 4 (dimension padded to 16)
                    const uint4 v = *reinterpret_cast<const uint4*>(mu + w);
                    dist          = sq_diff4(v.x, xs[w], dist);
                    dist          = sq_diff4(v.y, xs[w + 1], dist);
                    dist          = sq_diff4(v.z, xs[w + 2], dist);
                    dist          = sq_diff4(v.w, xs[w + 3], dist);
                }
                atomicMin(&sBest[p.densMix[row]], dist + p.consts[row]);
            }
        }
        __syncthreads();
        for (int m = threadIdx.x; m < p.nMix; m += blockDim.x)
            p.scores[t * p.nMix + m] = __fdiv_rn(__int2float_rn(sBest[m]), p.scale);
    }
}

// glibc's rand() after srand(seed) (TYPE_3 additive feedback generator); see gmm_presel.cu
struct GlibcRandInt {
    std::vector<uint32_t> r;
    explicit GlibcRandInt(uint32_t seed) : r(34) {
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

unsigned char quantize_u8_host(float x) {
    const int v = (int)std::round(x) + 128;
    return (unsigned char)std::min(std::max(v, 0), 255);
}

// Mm::unrolledVectorDistance<u8, s32> (src/Mm/Utilities.hh:254-296)
int host_distance(const uint8_t* a, const uint8_t* b, int dim) {
    int score = 0;
    for (int d = 0; d < dim; ++d) {
        const int df = (int)a[d] - (int)b[d];
        score += df * df;
    }
    return score;
}

}  // namespace

struct rb_gmm_presel_int {
    rb::DeviceInfo dev;
    int            dim = 0, padded = 0, nMix = 0, nDens = 0, nClusters = 0, nSelected = 0;
    float          scale = 1.0f;
    std::vector<uint8_t>  means, clusterMeans;  // [nDens * padded] in mixture order, [nClusters * padded]
    std::vector<int>      consts;               // [nDens] in mixture order
    std::vector<uint32_t> densMix, clusterOf;   // [nDens] in mixture order
    rb::DevBuf<float>     dVariance;
    rb::DevBuf<uint32_t>  dMeans, dClusterMeans, dDensMix, dClusterStart, dXq, dActive;  // dMeans.. in cluster order
    rb::DevBuf<int>       dConsts;
};

namespace {

// DensityClustering<u8, s32>::build (.tcc:126-161): initializeClusters :62-75, assignDensities :82-100,
// updateClusterMeans :103-123 (f64 sums, the quotient assigned to a u8: truncation)
void build_clustering(rb_gmm_presel_int* h, int clusters, int iterations) {
    const int dim = h->padded, nDens = h->nDens;
    h->nClusters  = std::min(clusters, nDens);
    h->clusterMeans.assign((size_t)h->nClusters * dim, 0);
    h->clusterOf.assign(nDens, 0);
    std::set<uint32_t> used;
    GlibcRandInt       rng(1);
    for (int c = 0; c < h->nClusters; ++c) {
        uint32_t pick = 0;
        do {
            pick = (uint32_t)(rng.next() % nDens);
        } while (used.count(pick));
        used.insert(pick);
        std::memcpy(&h->clusterMeans[(size_t)c * dim], &h->means[(size_t)pick * dim], dim);
    }
    for (int it = 0; it < iterations; ++it) {
        std::vector<std::vector<uint32_t>> assigned(h->nClusters);
        for (int dns = 0; dns < nDens; ++dns) {
            int best = INT_MAX, bc = 0;
            for (int c = 0; c < h->nClusters; ++c) {
                const int dist = host_distance(&h->clusterMeans[(size_t)c * dim], &h->means[(size_t)dns * dim], dim);
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
                h->clusterMeans[(size_t)c * dim + d] = (uint8_t)(sums[d] / (double)assigned[c].size());
        }
    }
}

}  // namespace

int rb_gmm_presel_int_configure(rb_gmm_presel_int* h, int clusters, int select, int iterations, cudaStream_t s) {
    RB_REQUIRE(clusters >= 1 && clusters <= 256 && select >= 1 && iterations >= 0, "bad preselection parameters");
    RB_REQUIRE(select <= std::min(clusters, h->nDens), "select-clusters %d exceeds the %d clusters", select,
               std::min(clusters, h->nDens));
    build_clustering(h, clusters, iterations);
    h->nSelected = select;
    // the device copy of the model, cluster by cluster (within a cluster: mixture order)
    std::vector<uint32_t> start(h->nClusters + 1, 0), mix(h->nDens);
    for (int e = 0; e < h->nDens; ++e)
        ++start[h->clusterOf[e] + 1];
    for (int c = 0; c < h->nClusters; ++c)
        start[c + 1] += start[c];
    std::vector<uint32_t> fill(start.begin(), start.end() - 1);
    std::vector<uint8_t>  rows(h->means.size());
    std::vector<int>      consts(h->nDens);
    for (int e = 0; e < h->nDens; ++e) {
        const uint32_t row = fill[h->clusterOf[e]]++;
        std::memcpy(&rows[(size_t)row * h->padded], &h->means[(size_t)e * h->padded], h->padded);
        consts[row] = h->consts[e];
        mix[row]    = h->densMix[e];
    }
    RB_CHECK(h->dMeans.upload(reinterpret_cast<const uint32_t*>(rows.data()), rows.size() / 4, s));
    RB_CHECK(h->dConsts.upload(consts, s));
    RB_CHECK(h->dDensMix.upload(mix, s));
    RB_CHECK(h->dClusterStart.upload(start, s));
    RB_CHECK(h->dClusterMeans.upload(reinterpret_cast<const uint32_t*>(h->clusterMeans.data()),
                                     h->clusterMeans.size() / 4, s));
    RB_CUDA(cudaStreamSynchronize(s));
    return RB_OK;
}

// model preparation = BatchIntFeatureScorer::init (src/Mm/BatchFeatureScorer.cc:375-416), f32 / f64 mixing as there
int rb_gmm_presel_int_create(const rb_mixture_set* ms, const rb::DeviceInfo& dev, cudaStream_t stream,
                             rb_gmm_presel_int** out) {
    *out = nullptr;
    if (ms->n_covariances != 1) {
        rb::set_error("int feature scorer supports only globally pooled variance (got %u covariances)",
                      ms->n_covariances);
        return RB_ERR_UNSUPPORTED;
    }
    if (ms->dim > 128) {
        rb::set_error("int preselection scorer supports feature dimension <= 128 (got %u)", ms->dim);
        return RB_ERR_UNSUPPORTED;
    }
    if (ms->mix_offsets[ms->n_mixtures] == 0) {
        rb::set_error("int preselection scorer needs at least one density");
        return RB_ERR_INVALID;
    }
    rb_gmm_presel_int* h = new (std::nothrow) rb_gmm_presel_int();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    const unsigned D = ms->dim;
    h->dev    = dev;
    h->dim    = (int)D;
    h->padded = ((int)D + 15) / 16 * 16;
    h->nMix   = (int)ms->n_mixtures;
    h->nDens  = (int)ms->mix_offsets[ms->n_mixtures];
    std::vector<float> variance(D, 0.0f);
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
    for (unsigned d = 0; d < D; ++d)
        variance[d] = variance[d] * scale;
    double sumLog = 0;
    for (unsigned d = 0; d < D; ++d)
        sumLog += std::log(std::fabs((double)ms->variances[d]));
    const float logNorm       = (float)((double)D * std::log(2.0 * M_PI) + sumLog);
    const float logNormFactor = logNorm * scaleSquared;

    h->means.assign((size_t)h->nDens * h->padded, 0);
    h->consts.assign(h->nDens, 0);
    h->densMix.assign(h->nDens, 0);
    for (uint32_t m = 0; m < ms->n_mixtures; ++m) {
        for (uint32_t e = ms->mix_offsets[m]; e < ms->mix_offsets[m + 1]; ++e) {
            const uint32_t dns = ms->mix_density[e];
            if (ms->dens_cov[dns] != 0) {
                delete h;
                rb::set_error("density %u does not use covariance 0", dns);
                return RB_ERR_INVALID;
            }
            const float* mu = ms->means + (size_t)ms->dens_mean[dns] * D;
            for (unsigned d = 0; d < D; ++d)
                h->means[(size_t)e * h->padded + d] = quantize_u8_host(mu[d] * variance[d]);
            h->consts[e]  = (int)((double)logNormFactor - (double)h->scale * ms->mix_log_weight[e]);
            h->densMix[e] = m;
        }
    }
    int rc = h->dVariance.upload(variance, stream);
    if (rc == RB_OK && cudaStreamSynchronize(stream) != cudaSuccess) {
        rb::set_error("int preselection model upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = RB_ERR_CUDA;
    }
    if (rc == RB_OK)  // DensityClustering.cc:20-34: clusters 256, select-clusters 32, iterations 5
        rc = rb_gmm_presel_int_configure(h, 256, std::min(32, h->nDens), 5, stream);
    if (rc != RB_OK) {
        delete h;
        return rc;
    }
    *out = h;
    return RB_OK;
}

void rb_gmm_presel_int_destroy(rb_gmm_presel_int* h) {
    delete h;
}

int rb_gmm_presel_int_score(rb_gmm_presel_int* h, const float* dFeats, long T, float* dScores, cudaStream_t s) {
    const int words = h->padded / 4;
    RB_CHECK(h->dActive.reserve((size_t)T * 8));
    RB_CHECK(h->dXq.reserve((size_t)T * words));
    PreselIntParams p;
    p.feats        = dFeats;
    p.variance     = h->dVariance.p;
    p.means        = h->dMeans.p;
    p.consts       = h->dConsts.p;
    p.densMix      = h->dDensMix.p;
    p.clusterStart = h->dClusterStart.p;
    p.clusterMeans = h->dClusterMeans.p;
    p.xq           = h->dXq.p;
    p.active       = h->dActive.p;
    p.scores       = dScores;
    p.T            = T;
    p.dim          = h->dim;
    p.words        = words;
    p.nMix         = h->nMix;
    p.nDens        = h->nDens;
    p.nClusters    = h->nClusters;
    p.nSelected    = h->nSelected;
    p.scale        = h->scale;
    const size_t smem = ((size_t)h->nClusters * (words | 1) + (size_t)kSelWarps * words + 4) * 4 +
                        (size_t)kSelWarps * 256 * sizeof(DistCluster);
    RB_CUDA(cudaFuncSetAttribute(presel_int_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid1 = (int)std::min<long>((T + kSelWarps - 1) / kSelWarps, (long)h->dev.sm_count * 4);
    presel_int_select_kernel<<<grid1, kSelWarps * 32, smem, s>>>(p);
    RB_LAUNCH_CHECK();
    const size_t smemBest = (size_t)h->nMix * 4;
    RB_REQUIRE(smemBest + 4 * 1024 <= h->dev.smem_optin, "int preselection scorer: %d mixtures do not fit", h->nMix);
    RB_CUDA(cudaFuncSetAttribute(presel_int_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBest));
    const int grid2 = (int)std::min<long>(T, (long)h->dev.sm_count * 8);
    presel_int_score_kernel<<<grid2, 256, smemBest, s>>>(p);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

// cluster means as f32 values of the u8 centroids, [n_clusters * padded] with padded = dim rounded up to 16
void rb_gmm_presel_int_clustering(const rb_gmm_presel_int* h, uint32_t* cluster_of, float* cluster_means,
                                  int* n_clusters) {
    if (cluster_of)
        std::copy(h->clusterOf.begin(), h->clusterOf.end(), cluster_of);
    if (cluster_means)
        for (size_t i = 0; i < h->clusterMeans.size(); ++i)
            cluster_means[i] = (float)h->clusterMeans[i];
    if (n_clusters)
        *n_clusters = h->nClusters;
}

// test hook (host only): (key, index) pairs sorted by key only with introsort.cuh; perm = the index order
extern "C" void rb_test_introsort(const int* keys, int n, int* perm) {
    std::vector<DistCluster> v((size_t)std::max(n, 0));
    for (int i = 0; i < n; ++i)
        v[i] = DistCluster{keys[i], (uint32_t)i};
    rb::introsort::sort(v.data(), v.data() + v.size(), ByDistance());
    for (int i = 0; i < n; ++i)
        perm[i] = (int)v[i].cluster;
}
