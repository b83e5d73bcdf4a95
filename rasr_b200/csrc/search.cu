// search.cu -- the score consumer of config C5: Search::LinearSearch (time-synchronous Viterbi over the linear HMMs of
// all pronunciations, unigram LM, one book-keeping entry per time frame) fed from the dense score matrix on the device.
//
//   LinearSearch::feed            src/Search/LinearSearch.cc:233-379
//   LinearSearch::bookKeeping     src/Search/LinearSearch.cc:381-432
//   restart / Book                src/Search/LinearSearch.cc:145-152,224-231,487-493
//   getCurrentBestSentence        src/Search/LinearSearch.cc:438-468
//
// The reference calls  emissionScores->score(mixture)  through two virtual calls per HMM state and frame; here the
// scores of a whole segment already sit in HBM ([T x nEmissions], written by the scorers of this library) and one
// CTA walks one segment: a thread owns a word, keeps the recurrence of feed() for its states in place (descending
// state order, so the previous frame's values of states s, s-1, s-2 are still intact), then the warp 0 replays the
// book-keeping scan over the word ends IN WORD ORDER -- its acceptance test compares against (score - lm) + lm of the
// current best, which is not an associative minimum, so a tree reduction would not be bit-identical -- skipping 32
// words at a time when none of them beats the current threshold.  Segments are independent (one CTA each); the time
// loop is inherently sequential.  All arithmetic is the reference's f32 arithmetic in its order: word sequences, word
// end times and both scores are bit-identical to the CPU path.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int kThreads = 1024;

struct SearchParams {
    // lexicon
    const uint32_t* wordOff;    // [W+1]
    const uint32_t* stateEmis;  // [nStates]
    const uint32_t* stateTdp;   // [nStates]
    const float*    tdp;        // [nModels * 4]
    const float*    unigram;    // [W]
    uint32_t        W, nStates, entryModel, nModels;
    // input
    const float*   scores;    // [frames * nEmis]
    const int64_t* frameOff;  // [U+1]
    int            nEmis;
    // per segment state, stride nStates + W entries (index 0 of a word = entry hypothesis)
    float* hypScore;
    float* hypLm;
    int*   hypBkp;
    // per segment book (capacity = frames of the segment, laid out at frameOff[u])
    float* bookScore;
    float* bookLm;
    int*   bookWord;
    int*   bookBkp;
    int*   bookTime;
    int*   nBooks;  // [U]
    // word-end candidates of the current frame, per segment [W]
    float* endScore;
    int    useSmem;  // hypotheses, lexicon tables and word-end candidates of the segment live in shared memory
};

__global__ void __launch_bounds__(kThreads) linear_search_kernel(const SearchParams p) {
    const int      u      = blockIdx.x;
    const int64_t  f0     = p.frameOff[u];
    const int      T      = (int)(p.frameOff[u + 1] - f0);
    const size_t   stride = (size_t)p.nStates + p.W;
    float*          hs     = p.hypScore + (size_t)u * stride;
    float*          hl     = p.hypLm + (size_t)u * stride;
    int*            hb     = p.hypBkp + (size_t)u * stride;
    float*          es     = p.endScore + (size_t)u * p.W;
    const uint32_t* wordOff   = p.wordOff;
    const uint32_t* stateEmis = p.stateEmis;
    const uint32_t* stateTdp  = p.stateTdp;
    const float*    tdp       = p.tdp;
    const float*    unigram   = p.unigram;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    if (p.useSmem) {
        // the working set of a segment (a few hundred KB at most for a 1000-word lexicon) is latency critical: every
        // frame is a dependent step
        float* f = reinterpret_cast<float*>(smemRaw);
        hs       = f;
        hl       = hs + stride;
        hb       = reinterpret_cast<int*>(hl + stride);
        es       = reinterpret_cast<float*>(hb + stride);
        float*    sUni = es + p.W;
        uint32_t* sOff = reinterpret_cast<uint32_t*>(sUni + p.W);
        uint32_t* sEm  = sOff + p.W + 1;
        uint32_t* sTd  = sEm + p.nStates;
        for (uint32_t i = threadIdx.x; i < p.W; i += kThreads)
            sUni[i] = p.unigram[i];
        for (uint32_t i = threadIdx.x; i <= p.W; i += kThreads)
            sOff[i] = p.wordOff[i];
        for (uint32_t i = threadIdx.x; i < p.nStates; i += kThreads) {
            sEm[i] = p.stateEmis[i];
            sTd[i] = p.stateTdp[i];
        }
        wordOff   = sOff;
        stateEmis = sEm;
        stateTdp  = sTd;
        unigram   = sUni;
    }
    float*         bScore = p.bookScore + f0;
    float*         bLm    = p.bookLm + f0;
    int*           bWord  = p.bookWord + f0;
    int*           bBkp   = p.bookBkp + f0;
    int*           bTime  = p.bookTime + f0;
    __shared__ int   sLast;                  // index of the newest book entry, -1 if none
    __shared__ float sLastScore, sLastLm;
    __shared__ float sTdp[64 * 4];           // transition models (when there are at most 64)
    if (p.nModels <= 64) {
        for (uint32_t i = threadIdx.x; i < p.nModels * 4; i += kThreads)
            sTdp[i] = p.tdp[i];
        tdp = sTdp;
    }

    // restart(): every hypothesis at FLT_MAX
    for (size_t i = threadIdx.x; i < stride; i += kThreads) {
        hs[i] = FLT_MAX;
        hl[i] = 0.0f;
        hb[i] = -1;
    }
    if (threadIdx.x == 0) {
        sLast      = -1;
        sLastScore = 0.0f;
        sLastLm    = 0.0f;
    }
    __syncthreads();

    for (int t = 1; t <= T; ++t) {
        const float* sc   = p.scores + (size_t)(f0 + t - 1) * p.nEmis;
        const int    last = sLast;
        const float  lastScore = sLastScore, lastLm = sLastLm;
        for (uint32_t w = threadIdx.x; w < p.W; w += kThreads) {
            const uint32_t s0 = wordOff[w], S = wordOff[w + 1] - s0;
            float*         ws = hs + s0 + w;  // [0..S]
            float*         wl = hl + s0 + w;
            int*           wb = hb + s0 + w;
            // word start (:271-290): from the newest book entry, or from scratch
            float h0lm, h0s;
            if (last >= 0) {
                h0lm = __fadd_rn(unigram[w], lastLm);
                h0s  = lastScore;
            }
            else {
                h0lm = unigram[w];
                h0s  = 0.0f;
            }
            h0s = __fadd_rn(h0s, h0lm);
            // states in descending order: the old values of sta, sta-1, sta-2 are still in place
            for (uint32_t sta = S; sta >= 1; --sta) {
                float bestS = FLT_MAX, bestL = 0.0f;
                int   bestB = -1;
                for (uint32_t pre = sta >= 2 ? sta - 2 : 0; pre <= sta; ++pre) {
                    float    ps, pl;
                    int      pb;
                    uint32_t model;
                    if (pre == 0) {
                        ps    = h0s;
                        pl    = h0lm;
                        pb    = last;
                        model = p.entryModel;
                    }
                    else {
                        ps    = ws[pre];
                        pl    = wl[pre];
                        pb    = wb[pre];
                        model = stateTdp[s0 + pre - 1];
                    }
                    const float sco = __fadd_rn(ps, tdp[model * 4 + (sta - pre)]);
                    if (sco < bestS) {
                        bestS = sco;
                        bestL = pl;
                        bestB = pb;
                    }
                }
                ws[sta] = __fadd_rn(bestS, __ldg(sc + stateEmis[s0 + sta - 1]));
                wl[sta] = bestL;
                wb[sta] = bestB;
            }
            // word end candidate (:400-404)
            es[w] = __fadd_rn(ws[S], tdp[stateTdp[s0 + S - 1] * 4 + 3]);
        }
        __syncthreads();
        // book keeping (:381-432): sequential scan over the words, replayed by warp 0
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            float     nbScore = FLT_MAX, nbLm = 0.0f;
            int       nbWord = -1, nbBkp = -1;
            for (uint32_t base = 0; base < p.W; base += 32) {
                const uint32_t w   = base + lane;
                const float    cand = w < p.W ? es[w] : FLT_MAX;
                uint32_t       todo = 0xffffffffu;
                while (true) {
                    const float    thr  = __fadd_rn(nbScore, nbLm);
                    const uint32_t hits = __ballot_sync(0xffffffffu, w < p.W && cand < thr) & todo;
                    if (!hits)
                        break;
                    const int      first = __ffs(hits) - 1;  // lowest word of the chunk that beats the current best
                    const uint32_t ww    = base + first;
                    const uint32_t s0 = wordOff[ww], S = wordOff[ww + 1] - s0;
                    const float    tmpScore = __shfl_sync(0xffffffffu, cand, first);
                    const float    lmw = hl[s0 + ww + S];
                    nbScore = __fsub_rn(tmpScore, lmw);
                    nbLm    = lmw;
                    nbBkp   = hb[s0 + ww + S];
                    nbWord  = (int)ww;
                    todo    = first == 31 ? 0u : (0xffffffffu << (first + 1));  // words before it were already rejected
                }
            }
            if (lane == 0 && nbScore != FLT_MAX) {
                const int b = sLast + 1;  // entries are only ever appended: the newest is the last
                bScore[b]   = nbScore;
                bLm[b]      = nbLm;
                bWord[b]    = nbWord;
                bBkp[b]     = nbBkp;
                bTime[b]    = t;
                sLast       = b;
                sLastScore  = nbScore;
                sLastLm     = nbLm;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        p.nBooks[u] = sLast + 1;
}

}  // namespace

struct rb_search {
    rb::DeviceInfo dev;
    uint32_t       W = 0, nStates = 0, nModels = 0, entryModel = 0;
    cudaStream_t   stream = nullptr;
    rb::DevBuf<uint32_t> dWordOff, dStateEmis, dStateTdp;
    rb::DevBuf<float>    dTdp, dUnigram, dHypScore, dHypLm, dBookScore, dBookLm, dEndScore, dScores;
    rb::DevBuf<int>      dHypBkp, dBookWord, dBookBkp, dBookTime, dNBooks;
    rb::DevBuf<int64_t>  dFrameOff;
    // results of the last decode, on the host
    std::vector<int64_t> frameOff;
    std::vector<float>   bookScore, bookLm;
    std::vector<int>     bookWord, bookBkp, bookTime, nBooks;
    ~rb_search() {
        if (stream)
            cudaStreamDestroy(stream);
    }
};

extern "C" int rb_search_create(const rb_lexicon* lx, int device, rb_search** out) {
    RB_REQUIRE(lx && out, "NULL argument");
    *out = nullptr;
    RB_REQUIRE(lx->n_words >= 1 && lx->word_offsets && lx->state_emission && lx->state_tdp_model && lx->tdp && lx->unigram,
               "lexicon has NULL tables or no words");
    RB_REQUIRE(lx->n_models >= 1 && lx->entry_model < lx->n_models, "entry model %u of %u", lx->entry_model, lx->n_models);
    const uint32_t nStates = lx->word_offsets[lx->n_words];
    for (uint32_t w = 0; w < lx->n_words; ++w)
        RB_REQUIRE(lx->word_offsets[w + 1] > lx->word_offsets[w], "word %u has no HMM state", w);
    for (uint32_t s = 0; s < nStates; ++s)
        RB_REQUIRE(lx->state_tdp_model[s] < lx->n_models, "state %u refers to transition model %u >= %u", s,
                   lx->state_tdp_model[s], lx->n_models);
    rb_search* h = new (std::nothrow) rb_search();
    if (!h) {
        rb::set_error("out of host memory");
        return RB_ERR_NOMEM;
    }
    auto fail = [&](int code) {
        delete h;
        return code;
    };
    int rc = rb::use_device(device, &h->dev);
    if (rc != RB_OK)
        return fail(rc);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        rb::set_error("cudaStreamCreate failed");
        return fail(RB_ERR_CUDA);
    }
    h->W          = lx->n_words;
    h->nStates    = nStates;
    h->nModels    = lx->n_models;
    h->entryModel = lx->entry_model;
    if (h->dWordOff.upload(lx->word_offsets, lx->n_words + 1, h->stream) != RB_OK ||
        h->dStateEmis.upload(lx->state_emission, nStates, h->stream) != RB_OK ||
        h->dStateTdp.upload(lx->state_tdp_model, nStates, h->stream) != RB_OK ||
        h->dTdp.upload(lx->tdp, (size_t)lx->n_models * 4, h->stream) != RB_OK ||
        h->dUnigram.upload(lx->unigram, lx->n_words, h->stream) != RB_OK ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
        rb::set_error("lexicon upload failed");
        return fail(RB_ERR_CUDA);
    }
    *out = h;
    return RB_OK;
}

extern "C" void rb_search_destroy(rb_search* h) {
    if (!h)
        return;
    cudaSetDevice(h->dev.ordinal);
    delete h;
}

// d_scores [frames x n_emissions] on the device; results are kept in the handle (rb_search_traceback)
extern "C" int rb_search_decode_dev(rb_search* h, const float* d_scores, int n_emissions, const int64_t* frame_offsets,
                                    int n_utt, void* stream) {
    RB_REQUIRE(h && frame_offsets && n_utt >= 0 && n_emissions >= 1, "bad argument");
    h->frameOff.assign(frame_offsets, frame_offsets + n_utt + 1);
    const int64_t base = frame_offsets[0], T = frame_offsets[n_utt] - base;
    for (auto& f : h->frameOff)
        f -= base;
    h->nBooks.assign(n_utt, 0);
    if (n_utt == 0 || T <= 0)
        return RB_OK;
    RB_REQUIRE(d_scores != nullptr, "NULL score buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    cudaStream_t s      = stream ? (cudaStream_t)stream : h->stream;
    const size_t stride = (size_t)h->nStates + h->W;
    RB_CHECK(h->dHypScore.reserve(stride * n_utt));
    RB_CHECK(h->dHypLm.reserve(stride * n_utt));
    RB_CHECK(h->dHypBkp.reserve(stride * n_utt));
    RB_CHECK(h->dEndScore.reserve((size_t)h->W * n_utt));
    RB_CHECK(h->dBookScore.reserve((size_t)T));
    RB_CHECK(h->dBookLm.reserve((size_t)T));
    RB_CHECK(h->dBookWord.reserve((size_t)T));
    RB_CHECK(h->dBookBkp.reserve((size_t)T));
    RB_CHECK(h->dBookTime.reserve((size_t)T));
    RB_CHECK(h->dNBooks.reserve((size_t)n_utt));
    RB_CHECK(h->dFrameOff.reserve((size_t)n_utt + 1));
    RB_CUDA(cudaMemcpyAsync(h->dFrameOff.p, h->frameOff.data(), sizeof(int64_t) * (n_utt + 1), cudaMemcpyHostToDevice, s));
    SearchParams p;
    p.wordOff    = h->dWordOff.p;
    p.stateEmis  = h->dStateEmis.p;
    p.stateTdp   = h->dStateTdp.p;
    p.tdp        = h->dTdp.p;
    p.unigram    = h->dUnigram.p;
    p.W          = h->W;
    p.nStates    = h->nStates;
    p.entryModel = h->entryModel;
    p.nModels    = h->nModels;
    p.scores     = d_scores;
    p.frameOff   = h->dFrameOff.p;
    p.nEmis      = n_emissions;
    p.hypScore   = h->dHypScore.p;
    p.hypLm      = h->dHypLm.p;
    p.hypBkp     = h->dHypBkp.p;
    p.bookScore  = h->dBookScore.p;
    p.bookLm     = h->dBookLm.p;
    p.bookWord   = h->dBookWord.p;
    p.bookBkp    = h->dBookBkp.p;
    p.bookTime   = h->dBookTime.p;
    p.nBooks     = h->dNBooks.p;
    p.endScore   = h->dEndScore.p;
    const size_t smem = (stride * 3 + (size_t)h->W * 2 + (h->W + 1) + (size_t)h->nStates * 2) * 4;
    p.useSmem         = smem <= h->dev.smem_optin - 1024 ? 1 : 0;
    if (p.useSmem)
        RB_CUDA(cudaFuncSetAttribute(linear_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linear_search_kernel<<<n_utt, kThreads, p.useSmem ? smem : 0, s>>>(p);
    RB_LAUNCH_CHECK();
    h->bookScore.resize(T);
    h->bookLm.resize(T);
    h->bookWord.resize(T);
    h->bookBkp.resize(T);
    h->bookTime.resize(T);
    RB_CUDA(cudaMemcpyAsync(h->bookScore.data(), h->dBookScore.p, sizeof(float) * T, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaMemcpyAsync(h->bookLm.data(), h->dBookLm.p, sizeof(float) * T, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaMemcpyAsync(h->bookWord.data(), h->dBookWord.p, sizeof(int) * T, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaMemcpyAsync(h->bookBkp.data(), h->dBookBkp.p, sizeof(int) * T, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaMemcpyAsync(h->bookTime.data(), h->dBookTime.p, sizeof(int) * T, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaMemcpyAsync(h->nBooks.data(), h->dNBooks.p, sizeof(int) * n_utt, cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaStreamSynchronize(s));
    return RB_OK;
}

extern "C" int rb_search_decode(rb_search* h, const float* scores, int n_emissions, const int64_t* frame_offsets,
                                int n_utt) {
    RB_REQUIRE(h && frame_offsets && n_utt >= 0 && n_emissions >= 1, "bad argument");
    const int64_t base = frame_offsets[0], T = frame_offsets[n_utt] - base;
    if (n_utt == 0 || T <= 0)
        return rb_search_decode_dev(h, nullptr, n_emissions, frame_offsets, n_utt, nullptr);
    RB_REQUIRE(scores != nullptr, "NULL score buffer");
    RB_CUDA(cudaSetDevice(h->dev.ordinal));
    RB_CHECK(h->dScores.reserve((size_t)T * n_emissions));
    RB_CUDA(cudaMemcpyAsync(h->dScores.p, scores + base * n_emissions, (size_t)T * n_emissions * 4,
                            cudaMemcpyHostToDevice, h->stream));
    return rb_search_decode_dev(h, h->dScores.p, n_emissions, frame_offsets, n_utt, h->stream);
}

// getCurrentBestSentence of segment `utt` of the last decode: word ends in chronological order.  Any output pointer
// may be NULL; capacity = frames of the segment.  Returns the number of words, < 0 on error.
extern "C" long rb_search_traceback(const rb_search* h, int utt, uint32_t* words, int32_t* times, float* am_scores,
                                    float* lm_scores) {
    if (!h || utt < 0 || utt >= (int)h->nBooks.size()) {
        rb::set_error("no such segment in the last decode");
        return RB_ERR_INVALID;
    }
    const int64_t    f0 = h->frameOff[utt];
    std::vector<int> chain;
    for (int b = h->nBooks[utt] - 1; b >= 0; b = h->bookBkp[f0 + b])
        chain.push_back(b);
    long n = 0;
    for (auto it = chain.rbegin(); it != chain.rend(); ++it, ++n) {
        if (words)
            words[n] = (uint32_t)h->bookWord[f0 + *it];
        if (times)
            times[n] = h->bookTime[f0 + *it];
        if (am_scores)
            am_scores[n] = h->bookScore[f0 + *it];
        if (lm_scores)
            lm_scores[n] = h->bookLm[f0 + *it];
    }
    return n;
}
