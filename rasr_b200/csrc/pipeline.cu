// pipeline.cu -- fused audio -> emission scores (BASELINE config C3): the front-end kernels and the
// GMM scorer are chained on one stream, features stay in HBM/L2 and never visit the host.
// Replaces the recognizer's per-frame loop  DataSource::getData -> FeatureScorer::getScorer ->
// ContextScorer::score  (src/Speech/DataExtractor.cc:101-111, src/Speech/Recognizer.cc:271-281).
#include "internal.h"

namespace {
struct Scratch {
    rb::DevBuf<float> samples, feats, scores;
};
// one scratch set per front-end handle would be cleaner; pipelines are few, so key by handle
Scratch& scratch_for(const rb_frontend* fe) {
    static thread_local std::vector<std::pair<const rb_frontend*, Scratch*>> table;
    for (auto& e : table)
        if (e.first == fe)
            return *e.second;
    table.emplace_back(fe, new Scratch());
    return *table.back().second;
}
}  // namespace

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

extern "C" int rb_pipeline_score(rb_frontend* fe, rb_gmm* gmm, const float* samples, const int64_t* offsets, int n_utt,
                                 float* scores, float* feats) {
    RB_REQUIRE(fe && gmm && offsets && n_utt >= 0, "bad argument");
    if (n_utt == 0)
        return RB_OK;
    const int64_t base = offsets[0], nS = offsets[n_utt] - base;
    RB_REQUIRE(nS >= 0 && (samples || nS == 0), "bad sample buffer");
    std::vector<int64_t> rel(n_utt + 1);
    for (int u = 0; u <= n_utt; ++u)
        rel[u] = offsets[u] - base;
    const long T = rb_frontend_count_frames(fe, rel.data(), n_utt, nullptr);
    if (T <= 0)
        return T == 0 ? RB_OK : RB_ERR_INVALID;
    RB_REQUIRE(scores != nullptr, "NULL score buffer");
    RB_CUDA(cudaSetDevice(rb_frontend_device(fe).ordinal));
    Scratch&     sc = scratch_for(fe);
    const int    D = rb_frontend_feat_dim(fe), M = rb_gmm_n_mixtures(gmm);
    cudaStream_t s = rb_frontend_stream(fe);
    RB_CHECK(sc.samples.reserve((size_t)nS + 8));
    RB_CHECK(sc.feats.reserve((size_t)T * D));
    RB_CHECK(sc.scores.reserve((size_t)T * M));
    RB_CUDA(cudaMemcpyAsync(sc.samples.p, samples + base, (size_t)nS * 4, cudaMemcpyHostToDevice, s));
    RB_CHECK(rb_pipeline_score_dev(fe, gmm, sc.samples.p, rel.data(), n_utt, sc.feats.p, sc.scores.p, s));
    RB_CUDA(cudaMemcpyAsync(scores, sc.scores.p, (size_t)T * M * 4, cudaMemcpyDeviceToHost, s));
    if (feats)
        RB_CUDA(cudaMemcpyAsync(feats, sc.feats.p, (size_t)T * D * 4, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaStreamSynchronize(s));
    return RB_OK;
}
