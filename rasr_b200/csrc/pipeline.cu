// pipeline.cu -- fused audio -> emission scores (BASELINE config C3): the front-end kernels and the
// GMM scorer are chained on one stream, features stay in HBM/L2 and never visit the host.
// Replaces the recognizer's per-frame loop  DataSource::getData -> FeatureScorer::getScorer ->
// ContextScorer::score  (src/Speech/DataExtractor.cc:101-111, src/Speech/Recognizer.cc:271-281).
#include "internal.h"

namespace {
struct Scratch {
    rb::DevBuf<float>   samples, feats, post, scores;
    rb::DevBuf<int16_t> pcm;
    rb::CopyStreams     copy;
    rb::HostStager      stager;  // pageable result buffers (common.cuh)
    rb::PinnedBuf<unsigned char> hIn;  // pageable sample buffers: copied here by several cores, then DMA
};
// one scratch set per front-end handle (and calling thread), released when the handle is destroyed
std::vector<std::pair<const rb_frontend*, Scratch*>>& scratch_table() {
    static thread_local std::vector<std::pair<const rb_frontend*, Scratch*>> table;
    return table;
}
Scratch& scratch_for(const rb_frontend* fe) {
    auto& table = scratch_table();
    for (auto& e : table)
        if (e.first == fe)
            return *e.second;
    table.emplace_back(fe, new Scratch());
    return *table.back().second;
}
}  // namespace

// called by rb_frontend_destroy: the device buffers and copy streams of the pipelines that used this handle go with it
void rb_pipeline_forget(const rb_frontend* fe) {
    auto& table = scratch_table();
    for (size_t i = 0; i < table.size(); ++i)
        if (table[i].first == fe) {
            delete table[i].second;
            table.erase(table.begin() + i);
            return;
        }
}

extern "C" int rb_pipeline_score_dev(rb_frontend* fe, rb_gmm* gmm, const float* d_samples, const int64_t* offsets,
                                     int n_utt, float* d_feats, float* d_scores, void* stream) {
    RB_REQUIRE(fe && gmm && offsets && n_utt >= 0, "bad argument");
    RB_REQUIRE(rb_frontend_feat_dim(fe) == rb_gmm_dim(gmm), "front-end emits %d-dim features, mixture set expects %d",
               rb_frontend_feat_dim(fe), rb_gmm_dim(gmm));
    if (n_utt == 0)
        return RB_OK;
    const long T = rb_frontend_count_frames(fe, offsets, n_utt, nullptr);
    RB_REQUIRE(T >= 0, "bad offsets");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(d_samples && d_feats && d_scores, "NULL device buffer");
    cudaStream_t s = stream ? (cudaStream_t)stream : rb_frontend_stream(fe);
    RB_CHECK(rb_frontend_process_dev(fe, d_samples, offsets, n_utt, d_feats, s));
    RB_CHECK(rb_gmm_score_dev(gmm, d_feats, T, d_scores, nullptr, s));
    return RB_OK;
}

extern "C" int rb_pipeline_score_fanout_dev(rb_frontend* fe, rb_gmm* gmm, const float* d_samples, const int64_t* offsets,
                                            int n_utt, float* d_feats, int n_dst, float* const* d_dst, void* stream) {
    RB_REQUIRE(fe && gmm && offsets && n_utt >= 0, "bad argument");
    RB_REQUIRE(rb_frontend_feat_dim(fe) == rb_gmm_dim(gmm), "front-end emits %d-dim features, mixture set expects %d",
               rb_frontend_feat_dim(fe), rb_gmm_dim(gmm));
    if (n_utt == 0)
        return RB_OK;
    const long T = rb_frontend_count_frames(fe, offsets, n_utt, nullptr);
    RB_REQUIRE(T >= 0, "bad offsets");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(d_samples && d_feats && d_dst && n_dst >= 1, "NULL device buffer");
    cudaStream_t s = stream ? (cudaStream_t)stream : rb_frontend_stream(fe);
    RB_CHECK(rb_frontend_process_dev(fe, d_samples, offsets, n_utt, d_feats, s));
    RB_CHECK(rb_gmm_score_fanout_dev(gmm, d_feats, T, n_dst, d_dst, s));
    return RB_OK;
}

// Host-pointer entry point: the utterances are cut into slabs (whole utterances, ~8 per call); H2D of slab i+1,
// front-end + scoring of slab i and D2H of slab i-1 overlap on three streams.  PCIe is the end-to-end bound
// (640 B of samples in, 1 KB of scores out per frame).
namespace {
// samples: f32 mono (channels == 0) or interleaved s16 with `channels` channels of which `track` is used
// scores == nullptr with keepOnDevice: the scores stay in the scratch buffer (*dScoresOut) for a consumer on the device
int pipeline_score_host(rb_frontend* fe, rb_gmm* gmm, const void* samplesRaw, int channels, int track,
                        const int64_t* offsets, int n_utt, float* scores, float* feats, bool keepOnDevice = false,
                        float** dScoresOut = nullptr, std::vector<int64_t>* frameOffOut = nullptr) {
    RB_REQUIRE(fe && gmm && offsets && n_utt >= 0, "bad argument");
    if (n_utt == 0)
        return RB_OK;
    const float*   samples = channels ? nullptr : static_cast<const float*>(samplesRaw);
    const int16_t* pcm     = channels ? static_cast<const int16_t*>(samplesRaw) : nullptr;
    const int64_t base = offsets[0], nS = offsets[n_utt] - base;
    RB_REQUIRE(nS >= 0 && (samplesRaw || nS == 0), "bad sample buffer");
    std::vector<int64_t> rel(n_utt + 1), fo(n_utt + 1);
    for (int u = 0; u <= n_utt; ++u)
        rel[u] = offsets[u] - base;
    const long T = rb_frontend_count_frames(fe, rel.data(), n_utt, fo.data());
    if (T <= 0)
        return T == 0 ? RB_OK : RB_ERR_INVALID;
    RB_REQUIRE(scores != nullptr || keepOnDevice, "NULL score buffer");
    RB_CUDA(cudaSetDevice(rb_frontend_device(fe).ordinal));
    Scratch&     sc = scratch_for(fe);
    const int    D = rb_frontend_feat_dim(fe), M = rb_gmm_n_mixtures(gmm);
    if (frameOffOut)
        *frameOffOut = fo;
    cudaStream_t sK = rb_frontend_stream(fe);
    RB_CHECK(sc.samples.reserve((size_t)nS + 8));
    RB_CHECK(sc.feats.reserve((size_t)T * D));
    RB_CHECK(sc.scores.reserve((size_t)T * M));
    if (channels)
        RB_CHECK(sc.pcm.reserve((size_t)nS * channels));

    // slab boundaries (utterance indices).  The call is bound by the D2H stream of the scores, so the first slabs are
    // small (the D2H stream starts early) and grow geometrically to ~16384 frames
    std::vector<int> cut(1, 0);
    long             target = 2048, targetMax = 16384;
    if (const char* e = getenv("RB_PIPE_SLAB0"))  // experiments
        target = std::max(1, atoi(e));
    if (const char* e = getenv("RB_PIPE_SLABMAX"))
        targetMax = std::max(1, atoi(e));
    for (int u = 1; u <= n_utt; ++u)
        if (u == n_utt || fo[u] - fo[cut.back()] >= target) {
            cut.push_back(u);
            target = std::min<long>(2 * target, targetMax);
        }
    const int    nSlabs = (int)cut.size() - 1;
    {  // the scorer's scratch for the largest slab once, before the pipeline starts (a reallocation synchronises the device)
        long largest = 0;
        for (int i = 0; i < nSlabs; ++i)
            largest = std::max<long>(largest, fo[cut[i + 1]] - fo[cut[i]]);
        RB_CHECK(rb_gmm_reserve(gmm, largest));
    }
    RB_CHECK(sc.copy.ensure(2 * (size_t)nSlabs));
    cudaStream_t sIn = sc.copy.in, sOut = sc.copy.out;
    cudaEvent_t* evIn = sc.copy.pool.data();
    cudaEvent_t* evK  = sc.copy.pool.data() + nSlabs;
    int rc = RB_OK;
    // pageable result buffers go through the page-locked staging ring + worker threads (common.cuh HostStager)
    const bool stagedScores = scores && !rb_host_is_pinned(scores) && getenv("RB_NO_HOST_STAGER") == nullptr;
    const bool stagedFeats  = feats && !rb_host_is_pinned(feats) && getenv("RB_NO_HOST_STAGER") == nullptr;
    if (stagedScores || stagedFeats)
        RB_CHECK(sc.stager.ensure((size_t)4 << 20, 8, rb_frontend_device(fe).ordinal));
    const size_t sampleBytes = channels ? (size_t)channels * 2 : 4;
    const bool   stagedIn = nS * sampleBytes >= ((size_t)4 << 20) && !rb_host_is_pinned(samplesRaw) &&
                          getenv("RB_NO_HOST_STAGER") == nullptr;
    if (stagedIn)
        RB_CHECK(sc.hIn.reserve((size_t)nS * sampleBytes));
    for (int i = 0; i < nSlabs && rc == RB_OK; ++i) {
        const int     u0 = cut[i], u1 = cut[i + 1];
        // the aligned bulk copies of the front-end may read up to 3 samples before a slab's first utterance:
        // they are either the previous slab's (already ordered on sK) or padding, and never used
        const int64_t sA = rel[u0], sB = rel[u1];
        const int64_t fA = fo[u0], fB = fo[u1];
        if (sB > sA) {
            const void*  src   = channels ? (const void*)(pcm + (base + sA) * channels) : (const void*)(samples + base + sA);
            const size_t bytes = (size_t)(sB - sA) * sampleBytes;
            if (stagedIn) {  // (the staging area holds the whole call: no slot has to be waited for)
                rb::parallel_memcpy(sc.hIn.p + (size_t)sA * sampleBytes, src, bytes);
                src = sc.hIn.p + (size_t)sA * sampleBytes;
            }
            void* dst = channels ? (void*)(sc.pcm.p + sA * channels) : (void*)(sc.samples.p + sA);
            if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, sIn) != cudaSuccess)
                rc = RB_ERR_CUDA;
        }
        cudaEventRecord(evIn[i], sIn);
        cudaStreamWaitEvent(sK, evIn[i], 0);
        if (rc == RB_OK && channels)
            rc = rb_frontend_convert_s16_dev(fe, sc.pcm.p + sA * channels, sc.samples.p + sA, (long)(sB - sA), channels,
                                             track, sK);
        if (rc == RB_OK && fB > fA) {
            rb_frontend_set_upload_stream(fe, sIn);
            rc = rb_pipeline_score_dev(fe, gmm, sc.samples.p, rel.data() + u0, u1 - u0, sc.feats.p + fA * D,
                                       sc.scores.p + fA * M, sK);
            rb_frontend_set_upload_stream(fe, nullptr);
        }
        cudaEventRecord(evK[i], sK);
        cudaStreamWaitEvent(sOut, evK[i], 0);
        if (rc == RB_OK && fB > fA) {
            if (stagedScores)
                rc = sc.stager.d2h(scores + fA * M, sc.scores.p + fA * M, (size_t)(fB - fA) * M * 4, sOut);
            else if (scores && cudaMemcpyAsync(scores + fA * M, sc.scores.p + fA * M, (size_t)(fB - fA) * M * 4,
                                               cudaMemcpyDeviceToHost, sOut) != cudaSuccess)
                rc = RB_ERR_CUDA;
            if (rc == RB_OK && stagedFeats)
                rc = sc.stager.d2h(feats + fA * D, sc.feats.p + fA * D, (size_t)(fB - fA) * D * 4, sOut);
            else if (rc == RB_OK && feats &&
                     cudaMemcpyAsync(feats + fA * D, sc.feats.p + fA * D, (size_t)(fB - fA) * D * 4, cudaMemcpyDeviceToHost,
                                     sOut) != cudaSuccess)
                rc = RB_ERR_CUDA;
        }
    }
    const cudaError_t e1 = cudaStreamSynchronize(sIn), e2 = cudaStreamSynchronize(sK), e3 = cudaStreamSynchronize(sOut);
    if (stagedScores || stagedFeats) {
        const int rs = sc.stager.drain();
        if (rc == RB_OK)
            rc = rs;
    }
    if (rc == RB_ERR_CUDA && rb::get_error()[0] == 0)
        rb::set_error("asynchronous copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != RB_OK)
        return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        const cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
        rb::set_error("pipeline failed on the device: %s", cudaGetErrorString(e));
        return RB_ERR_CUDA;
    }
    if (dScoresOut)
        *dScoresOut = sc.scores.p;
    return RB_OK;
}
}  // namespace

// audio -> MFCC -> GMM scores -> LinearSearch (config C5): only the audio goes to the device and only the word
// sequences come back (rb_search_traceback / rb_search_traceback_all afterwards); the score matrix never leaves HBM
extern "C" int rb_pipeline_search(rb_frontend* fe, rb_gmm* gmm, rb_search* ls, const void* samples, int n_channels,
                                  int track, const int64_t* offsets, int n_utt) {
    RB_REQUIRE(fe && gmm && ls && offsets && n_utt >= 0, "bad argument");
    RB_REQUIRE(n_channels >= 0 && (n_channels == 0 || (track >= 0 && track < n_channels)), "track %d of %d channels",
               track, n_channels);
    float*               dScores = nullptr;
    std::vector<int64_t> fo;
    RB_CHECK(pipeline_score_host(fe, gmm, samples, n_channels, track, offsets, n_utt, nullptr, nullptr, true, &dScores,
                                 &fo));
    if (fo.empty())
        fo.assign((size_t)n_utt + 1, 0);
    return rb_search_decode_dev(ls, dScores, rb_gmm_n_mixtures(gmm), fo.data(), n_utt, rb_frontend_stream(fe));
}

extern "C" int rb_pipeline_score(rb_frontend* fe, rb_gmm* gmm, const float* samples, const int64_t* offsets, int n_utt,
                                 float* scores, float* feats) {
    return pipeline_score_host(fe, gmm, samples, 0, 0, offsets, n_utt, scores, feats);
}

extern "C" int rb_pipeline_score_s16(rb_frontend* fe, rb_gmm* gmm, const int16_t* samples, int n_channels, int track,
                                     const int64_t* offsets, int n_utt, float* scores, float* feats) {
    RB_REQUIRE(n_channels >= 1 && track >= 0 && track < n_channels, "track %d of %d channels", track, n_channels);
    return pipeline_score_host(fe, gmm, samples, n_channels, track, offsets, n_utt, scores, feats);
}

// =====================================================================================================
// audio -> MFCC -> (normalisation / splice / matrix) -> Nn scores: the C4 model fed from audio.  Same structure
// as the GMM pipeline above; the post-processing handle may be NULL (features go to the network as they are).
// =====================================================================================================
extern "C" int rb_pipeline_nn_score_dev(rb_frontend* fe, rb_postproc* pp, rb_nn* nn, const float* d_samples,
                                        const int64_t* offsets, int n_utt, float* d_feats, float* d_post,
                                        float* d_scores, void* stream) {
    RB_REQUIRE(fe && nn && offsets && n_utt >= 0, "bad argument");
    const int D = rb_frontend_feat_dim(fe), Dp = pp ? rb_postproc_dim_out(pp) : D;
    RB_REQUIRE(Dp == rb_nn_n_inputs(nn), "the network expects %d-dim input, the feature pipeline emits %d",
               rb_nn_n_inputs(nn), Dp);
    if (n_utt == 0)
        return RB_OK;
    std::vector<int64_t> fo(n_utt + 1);
    const long T = rb_frontend_count_frames(fe, offsets, n_utt, fo.data());
    RB_REQUIRE(T >= 0, "bad offsets");
    if (T == 0)
        return RB_OK;
    RB_REQUIRE(d_samples && d_feats && d_scores && (d_post || !pp), "NULL device buffer");
    cudaStream_t s = stream ? (cudaStream_t)stream : rb_frontend_stream(fe);
    RB_CHECK(rb_frontend_process_dev(fe, d_samples, offsets, n_utt, d_feats, s));
    const float* in = d_feats;
    if (pp) {
        RB_CHECK(rb_postproc_process_dev(pp, d_feats, fo.data(), n_utt, d_post, s));
        in = d_post;
    }
    return rb_nn_score_dev(nn, in, T, d_scores, s);
}

extern "C" int rb_pipeline_nn_score(rb_frontend* fe, rb_postproc* pp, rb_nn* nn, const float* samples,
                                    const int64_t* offsets, int n_utt, float* scores) {
    RB_REQUIRE(fe && nn && offsets && n_utt >= 0, "bad argument");
    if (n_utt == 0)
        return RB_OK;
    const int64_t base = offsets[0], nS = offsets[n_utt] - base;
    RB_REQUIRE(nS >= 0 && (samples || nS == 0), "bad sample buffer");
    std::vector<int64_t> rel(n_utt + 1), fo(n_utt + 1);
    for (int u = 0; u <= n_utt; ++u)
        rel[u] = offsets[u] - base;
    const long T = rb_frontend_count_frames(fe, rel.data(), n_utt, fo.data());
    if (T <= 0)
        return T == 0 ? RB_OK : RB_ERR_INVALID;
    RB_REQUIRE(scores != nullptr, "NULL score buffer");
    RB_CUDA(cudaSetDevice(rb_frontend_device(fe).ordinal));
    Scratch&     sc = scratch_for(fe);
    // rb_nn_score_dev writes [T x n_emissions] rows: the class count once a class mapping is active (which may exceed the
    // number of network outputs when classes are disregarded), else the number of outputs
    const int    D = rb_frontend_feat_dim(fe), Dp = pp ? rb_postproc_dim_out(pp) : D, M = rb_nn_n_emissions(nn);
    cudaStream_t sK = rb_frontend_stream(fe);
    // slabs of whole utterances, ~16384 frames each: the f32 score slab (48 KB per frame for 12k senones) is what
    // bounds the size, and its D2H copy is what bounds the call
    const long       target = 16384;
    std::vector<int> cut(1, 0);
    for (int u = 1; u <= n_utt; ++u)
        if (u == n_utt || fo[u] - fo[cut.back()] >= target)
            cut.push_back(u);
    const int nSlabs = (int)cut.size() - 1;
    long      maxSlab = 0;
    for (int i = 0; i < nSlabs; ++i)
        maxSlab = std::max<long>(maxSlab, (long)(fo[cut[i + 1]] - fo[cut[i]]));
    RB_CHECK(sc.samples.reserve((size_t)nS + 8));
    RB_CHECK(sc.feats.reserve((size_t)T * D));
    RB_CHECK(sc.post.reserve((size_t)(pp ? T : 1) * Dp));
    RB_CHECK(sc.scores.reserve((size_t)2 * maxSlab * M));  // two slabs in flight: scoring of i, D2H of i-1
    RB_CHECK(sc.copy.ensure(3 * (size_t)nSlabs));
    cudaStream_t sIn = sc.copy.in, sOut = sc.copy.out;
    cudaEvent_t* evIn  = sc.copy.pool.data();
    cudaEvent_t* evK   = sc.copy.pool.data() + nSlabs;
    cudaEvent_t* evOut = sc.copy.pool.data() + 2 * nSlabs;
    int rc = RB_OK;
    for (int i = 0; i < nSlabs && rc == RB_OK; ++i) {
        const int     u0 = cut[i], u1 = cut[i + 1];
        const int64_t sA = rel[u0], sB = rel[u1], fA = fo[u0], fB = fo[u1];
        float*        dScores = sc.scores.p + (size_t)(i & 1) * maxSlab * M;
        if (sB > sA && cudaMemcpyAsync(sc.samples.p + sA, samples + base + sA, (size_t)(sB - sA) * 4,
                                       cudaMemcpyHostToDevice, sIn) != cudaSuccess)
            rc = RB_ERR_CUDA;
        cudaEventRecord(evIn[i], sIn);
        cudaStreamWaitEvent(sK, evIn[i], 0);
        if (i >= 2)
            cudaStreamWaitEvent(sK, evOut[i - 2], 0);  // the score buffer of slab i-2 has been copied out
        if (rc == RB_OK && fB > fA) {
            rb_frontend_set_upload_stream(fe, sIn);
            rc = rb_pipeline_nn_score_dev(fe, pp, nn, sc.samples.p, rel.data() + u0, u1 - u0, sc.feats.p + fA * D,
                                          pp ? sc.post.p + fA * Dp : nullptr, dScores, sK);
            rb_frontend_set_upload_stream(fe, nullptr);
        }
        cudaEventRecord(evK[i], sK);
        cudaStreamWaitEvent(sOut, evK[i], 0);
        if (rc == RB_OK && fB > fA &&
            cudaMemcpyAsync(scores + fA * M, dScores, (size_t)(fB - fA) * M * 4, cudaMemcpyDeviceToHost, sOut) !=
                    cudaSuccess)
            rc = RB_ERR_CUDA;
        cudaEventRecord(evOut[i], sOut);
    }
    const cudaError_t e1 = cudaStreamSynchronize(sIn), e2 = cudaStreamSynchronize(sK), e3 = cudaStreamSynchronize(sOut);
    if (rc == RB_ERR_CUDA && rb::get_error()[0] == 0)
        rb::set_error("asynchronous copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != RB_OK)
        return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        const cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
        rb::set_error("pipeline failed on the device: %s", cudaGetErrorString(e));
        return RB_ERR_CUDA;
    }
    return RB_OK;
}
