/*
 * search_oracle.cc -- CPU restatement of the score consumer of config C5: Search::LinearSearch (time-synchronous
 * Viterbi over the linear HMMs of all pronunciations, unigram LM, one book-keeping entry per time frame).
 *   LinearSearch::feed            src/Search/LinearSearch.cc:233-379
 *   LinearSearch::bookKeeping     src/Search/LinearSearch.cc:381-432
 *   LinearSearch::restart / Book  src/Search/LinearSearch.cc:145-152,224-231,487-493
 *   getCurrentBestSentence        src/Search/LinearSearch.cc:438-468
 *   transition types              src/Am/TransitionModel.hh:32-37 (loop 0, forward 1, skip 2, exit 3)
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED: the reference has no test for Search, and LinearSearch.cc
 * does not compile into oracle/_ref without the Am / Lm / Bliss / Fsa stack; pinned by a literal Python transcription
 * of feed / bookKeeping only (tests/test_oracle_search.py).
 * What is restated is the arithmetic of feed / bookKeeping on a lexicon given as flat arrays (every pronunciation a
 * regular word, single-word recognition off, unigram scores precomputed as LinearSearch does: isUnigram() is always
 * true, :434-436); the Bliss lexicon, the Am model lookup and the lattice / history plumbing are not.
 */
#include "oracle.h"

#include <cfloat>
#include <vector>

namespace {
struct Book {
    float score, lmScore;
    int   word, bkp, time;
};
struct Hypo {
    float score, lmScore;
    int   bkp;
};
}  // namespace

/* scores [T x n_emissions] of ONE segment.  Outputs (chronological, capacity T each): words, times (1-based frame of
 * the word end), am (Book::score: without LM), lm (Book::lmScore).  Returns the number of words, < 0 on error. */
extern "C" long orc_linear_search(const orc_lexicon* lx, const float* scores, long T, int n_emissions, uint32_t* words,
                                  int32_t* times, float* am, float* lm) {
    const uint32_t W = lx->n_words;
    std::vector<std::vector<Hypo>> hyp(W);
    for (uint32_t w = 0; w < W; ++w) {
        const uint32_t S = lx->word_offsets[w + 1] - lx->word_offsets[w];
        if (S == 0)
            return -1;
        hyp[w].assign(S + 1, Hypo{FLT_MAX, 0.0f, -1}); /* WordPronunciationState::restart */
    }
    std::vector<Book> book;
    std::vector<Hypo> tmp;
    for (long t = 1; t <= T; ++t) {
        const float* sc = scores + (size_t)(t - 1) * n_emissions;
        for (uint32_t w = 0; w < W; ++w) {
            std::vector<Hypo>& h    = hyp[w];
            const uint32_t     s0   = lx->word_offsets[w];
            const int          last = book.empty() ? -1 : (int)book.size() - 1;
            h[0].bkp                = last;
            if (last >= 0) {
                h[0].lmScore = lx->unigram[w] + book[last].lmScore;
                h[0].score   = book[last].score;
            }
            else {
                h[0].lmScore = lx->unigram[w];
                h[0].score   = 0;
            }
            h[0].score += h[0].lmScore;
            tmp.resize(h.size());
            for (uint32_t sta = 1; sta < h.size(); ++sta) {
                tmp[sta].score   = FLT_MAX;
                tmp[sta].lmScore = 0;
                for (uint32_t pre = sta >= 2 ? sta - 2 : 0; pre <= sta; ++pre) {
                    const uint32_t model = pre != 0 ? lx->state_tdp_model[s0 + pre - 1] : lx->entry_model;
                    const float    sco   = h[pre].score + lx->tdp[model * 4 + (sta - pre)];
                    if (sco < tmp[sta].score) {
                        tmp[sta].score   = sco;
                        tmp[sta].bkp     = h[pre].bkp;
                        tmp[sta].lmScore = h[pre].lmScore;
                    }
                }
            }
            for (uint32_t sta = 1; sta < h.size(); ++sta) {
                h[sta].bkp     = tmp[sta].bkp;
                h[sta].score   = tmp[sta].score + sc[lx->state_emission[s0 + sta - 1]];
                h[sta].lmScore = tmp[sta].lmScore;
            }
        }
        Book nb{FLT_MAX, 0.0f, -1, -1, 0};
        for (uint32_t w = 0; w < W; ++w) { /* bookKeeping */
            const Hypo&    h     = hyp[w].back();
            const uint32_t model = lx->state_tdp_model[lx->word_offsets[w + 1] - 1];
            const float    tmpScore = h.score + lx->tdp[model * 4 + 3];
            if (tmpScore < nb.score + nb.lmScore) {
                nb.score   = tmpScore - h.lmScore;
                nb.lmScore = h.lmScore;
                nb.bkp     = h.bkp;
                nb.word    = (int)w;
                nb.time    = (int)t;
            }
        }
        if (nb.score != FLT_MAX)
            book.push_back(nb);
    }
    /* getCurrentBestSentence: follow the back pointers from the last book entry */
    std::vector<int> chain;
    for (int b = book.empty() ? -1 : (int)book.size() - 1; b >= 0; b = book[b].bkp)
        chain.push_back(b);
    long n = 0;
    for (auto it = chain.rbegin(); it != chain.rend(); ++it, ++n) {
        words[n] = (uint32_t)book[*it].word;
        times[n] = book[*it].time;
        am[n]    = book[*it].score;
        lm[n]    = book[*it].lmScore;
    }
    return n;
}
