/*
 * search_oracle.cc -- CPU restatement of the score consumer of config C5: Search::LinearSearch (time-synchronous
 * Viterbi over the linear HMMs of all pronunciations, unigram LM, one book-keeping entry per time frame).
 *   LinearSearch::feed            src/Search/LinearSearch.cc:233-379
 *   LinearSearch::bookKeeping     src/Search/LinearSearch.cc:381-432
 *   LinearSearch::restart / Book  src/Search/LinearSearch.cc:145-152,224-231,487-493
 *   getCurrentBestSentence        src/Search/LinearSearch.cc:438-468
 *   transition types              src/Am/TransitionModel.hh:32-37 (loop 0, forward 1, skip 2, exit 3)
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY PINNED to the reference's own object code: LinearSearch.cc with
 * the Bliss lexicon parser, the Am state / transition / tying models and the Lm scaling wrapper compiled from where
 * they lie into oracle/_ref/librasr_ref_search.so (oracle/refbuild/ref_search.cc), compared bit for bit on words,
 * end frames and both scores in tests/test_ref_search.py -- with single-word recognition on (the reference's
 * default, :26-30) and off; a literal Python transcription of feed / bookKeeping is the second pin
 * (tests/test_oracle_search.py).
 * What is restated is the arithmetic of feed / bookKeeping on a lexicon given as flat arrays (unigram scores
 * precomputed as LinearSearch does: isUnigram() is always true, :434-436); the Bliss lexicon, the Am model lookup
 * and the lattice / history plumbing are not.
 * Single-word recognition (:248-258, 355-370, 390-395, 479-485): an irregular pronunciation (silence, noise) gets a
 * second "irregular chain" entry right behind it; a second book records the best sequence of irregular words only;
 * a regular word may only start from that book once the main book's newest entry already contains a regular word,
 * so every sentence is  irregular* regular irregular*.
 */
#include "oracle.h"

#include <cfloat>
#include <vector>

namespace {
/* back pointers name an entry of either book: >= 0 the main book, <= -2 entry (-2 - bkp) of the irregular book */
struct Book {
    float score, lmScore;
    int   word, bkp, time;
    bool  hadRegularWord;
};
struct Hypo {
    float score, lmScore;
    int   bkp;
};
struct Entry {  /* one WordPronunciationState */
    uint32_t word;
    bool     regular, chain;
};
}  // namespace

/* scores [T x n_emissions] of ONE segment.  Outputs (chronological, capacity T each): words, times (1-based frame of
 * the word end), am (Book::score: without LM), lm (Book::lmScore).  Returns the number of words, < 0 on error. */
extern "C" long orc_linear_search(const orc_lexicon* lx, const float* scores, long T, int n_emissions, uint32_t* words,
                                  int32_t* times, float* am, float* lm) {
    const bool         single = lx->single_word != 0;
    std::vector<Entry> entries; /* addPronunciations, :472-487 */
    for (uint32_t w = 0; w < lx->n_words; ++w) {
        const bool regular = !lx->word_regular || lx->word_regular[w];
        entries.push_back(Entry{w, regular, false});
        if (single && !regular)
            entries.push_back(Entry{w, regular, true});
    }
    const size_t                   E = entries.size();
    std::vector<std::vector<Hypo>> hyp(E);
    for (size_t e = 0; e < E; ++e) {
        const uint32_t w = entries[e].word;
        const uint32_t S = lx->word_offsets[w + 1] - lx->word_offsets[w];
        if (S == 0)
            return -1;
        hyp[e].assign(S + 1, Hypo{FLT_MAX, 0.0f, -1}); /* WordPronunciationState::restart */
    }
    std::vector<Book> book, irregularBook;
    auto              entry_of = [&](int bkp) -> const Book& { return bkp >= 0 ? book[bkp] : irregularBook[-2 - bkp]; };
    std::vector<Hypo> tmp;
    /* bookKeeping (:381-432) */
    auto keep = [&](Book& nb, bool irregular, long t) {
        for (size_t e = 0; e < E; ++e) {
            const uint32_t w = entries[e].word;
            const Hypo&    h = hyp[e].back();
            if (irregular && entries[e].regular)
                continue;
            if (irregular && h.bkp != -1 && entry_of(h.bkp).hadRegularWord)
                continue;
            const uint32_t model    = lx->state_tdp_model[lx->word_offsets[w + 1] - 1];
            const float    tmpScore = h.score + lx->tdp[model * 4 + 3];
            if (tmpScore < nb.score + nb.lmScore) {
                nb.score          = tmpScore - h.lmScore;
                nb.lmScore        = h.lmScore;
                nb.bkp            = h.bkp;
                nb.word           = (int)w;
                nb.time           = (int)t;
                nb.hadRegularWord = entries[e].regular ? true : (h.bkp != -1 ? entry_of(h.bkp).hadRegularWord : false);
            }
        }
    };
    for (long t = 1; t <= T; ++t) {
        const float* sc = scores + (size_t)(t - 1) * n_emissions;
        for (size_t e = 0; e < E; ++e) {
            const uint32_t     w  = entries[e].word;
            std::vector<Hypo>& h  = hyp[e];
            const uint32_t     s0 = lx->word_offsets[w];
            int                last = book.empty() ? -1 : (int)book.size() - 1;
            if ((single && last >= 0 && book[last].hadRegularWord && entries[e].regular) || entries[e].chain)
                last = irregularBook.empty() ? -1 : -2 - ((int)irregularBook.size() - 1); /* :248-258 */
            h[0].bkp = last;
            if (last != -1) {
                h[0].lmScore = lx->unigram[w] + entry_of(last).lmScore;
                h[0].score   = entry_of(last).score;
            }
            else {
                h[0].lmScore = lx->unigram[w];
                h[0].score   = 0;
            }
            h[0].score += h[0].lmScore;
            tmp.resize(h.size(), Hypo{FLT_MAX, 0.0f, -1}); /* new elements: value-initialised in the reference, bkp = null */
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
        Book nb{FLT_MAX, 0.0f, -1, -1, 0, false};
        keep(nb, false, t);
        if (nb.score != FLT_MAX)
            book.push_back(nb);
        if (single) { /* :355-370 */
            Book ib{FLT_MAX, 0.0f, -1, -1, 0, false};
            keep(ib, true, t);
            if (ib.score != FLT_MAX)
                irregularBook.push_back(ib);
        }
    }
    /* getCurrentBestSentence: follow the back pointers from the last entry of the main book */
    std::vector<const Book*> chain;
    for (int b = book.empty() ? -1 : (int)book.size() - 1; b != -1; b = entry_of(b).bkp)
        chain.push_back(&entry_of(b));
    long n = 0;
    for (auto it = chain.rbegin(); it != chain.rend(); ++it, ++n) {
        words[n] = (uint32_t)(*it)->word;
        times[n] = (*it)->time;
        am[n]    = (*it)->score;
        lm[n]    = (*it)->lmScore;
    }
    return n;
}
