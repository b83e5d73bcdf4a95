/*
 * B200::LinearSearch -- Search::SearchAlgorithm adapter that forwards to the score consumer of librasr_b200.so
 * (the rb_search calls of include/rasr_b200.h): the drop-in for Search::LinearSearch (src/Search/LinearSearch.{hh,cc}).
 *
 * Written against the RASR headers (src/Search/Search.hh:39-150); compiled inside a RASR checkout.
 * setModelCombination() flattens lexicon + acoustic model + language model into the arrays of rb_lexicon the way
 * LinearSearch::Pronunciation / addPronunciations do (src/Search/LinearSearch.cc:32-96,472-487); feed() only collects
 * the score row of the frame; the first request for a result runs ONE rb_search_decode over the whole segment and
 * turns the book entries into the Traceback getCurrentBestSentence builds (:438-468).  Same parameter
 * ("single-word-recognition", default true) and same results, bit for bit.
 */
#ifndef _B200_LINEAR_SEARCH_HH
#define _B200_LINEAR_SEARCH_HH

#include <Core/Parameter.hh>
#include <Search/Search.hh>
#include <Speech/ModelCombination.hh>
#include <vector>

#include "B200HostBuffer.hh"
#include "rasr_b200.h"

namespace B200 {

class LinearSearch : public Search::SearchAlgorithm {
public:
    static const Core::ParameterBool paramSingleWordRecognition;  // as src/Search/LinearSearch.cc:26-30
    static const Core::ParameterInt  paramDevice;                 // CUDA ordinal

    LinearSearch(const Core::Configuration&);
    virtual ~LinearSearch();

    virtual bool setModelCombination(const Speech::ModelCombination& modelCombination);
    virtual bool setLanguageModel(Core::Ref<const Lm::ScaledLanguageModel>) {
        defect();
    }
    virtual void setGrammar(Fsa::ConstAutomatonRef);

    virtual void restart();
    virtual void feed(const Mm::FeatureScorer::Scorer&);
    virtual void getPartialSentence(Search::Traceback& result);
    virtual void getCurrentBestSentence(Search::Traceback& result) const;
    virtual Core::Ref<const Search::LatticeAdaptor> getCurrentWordLattice() const;
    virtual void resetStatistics();
    virtual void logStatistics() const;

    /** frames whose scores arrived as a whole row from a b200 feature scorer / through score(e) calls */
    u64 nDenseRows() const {
        return nDenseRows_;
    }
    u64 nScoreCalls() const {
        return nScoreCalls_;
    }

private:
    void decode() const;

    Bliss::LexiconRef                             lexicon_;
    Core::Ref<const Am::AcousticModel>            acousticModel_;
    Core::Ref<const Lm::ScaledLanguageModel>      lm_;
    bool                                          singleWordRecognition_;
    int                                           device_;
    rb_search*                                    handle_;
    std::vector<const Bliss::LemmaPronunciation*> pronunciations_;  // word number of the flat lexicon -> pronunciation
    u32                                           nEmissions_;
    HostBuffer                                    scores_;  // [time_ x nEmissions_] rows collected by feed()
    TimeframeIndex                                time_;
    u64                                           nDenseRows_, nScoreCalls_;
    // result of the last decode (getCurrentBestSentence is const, like the reference's)
    mutable TimeframeIndex      decodedTime_;
    mutable std::vector<u32>    words_;
    mutable std::vector<s32>    times_;
    mutable std::vector<Score>  am_, lmScores_;
};

}  // namespace B200

#endif
