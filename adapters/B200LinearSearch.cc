#include "B200LinearSearch.hh"

#include "B200FeatureScorer.hh"

#include <Am/ClassicStateModel.hh>
#include <Lattice/LatticeAdaptor.hh>
#include <Lm/ScaledLanguageModel.hh>
#include <Search/Traceback.hh>

using namespace B200;

const Core::ParameterBool LinearSearch::paramSingleWordRecognition(
        "single-word-recognition", "only recognize single words (as Search::LinearSearch)", true);
const Core::ParameterInt LinearSearch::paramDevice("device", "CUDA device ordinal", 0, 0);

LinearSearch::LinearSearch(const Core::Configuration& c)
        : Core::Component(c),
          Search::SearchAlgorithm(c),
          singleWordRecognition_(paramSingleWordRecognition(c)),
          device_(paramDevice(c)),
          handle_(0),
          nEmissions_(0),
          time_(0),
          nDenseRows_(0),
          nScoreCalls_(0),
          decodedTime_(0) {
    log("using B200 linear search") << (singleWordRecognition_ ? " (single-word recognition)" : "");
}

LinearSearch::~LinearSearch() {
    rb_search_destroy(handle_);
}

bool LinearSearch::setModelCombination(const Speech::ModelCombination& mc) {
    lexicon_       = mc.lexicon();
    acousticModel_ = mc.acousticModel();
    lm_            = mc.languageModel();
    require(lexicon_ && acousticModel_ && lm_);
    const Score          pronunciationScale = mc.pronunciationScale();
    const Am::Phonology& phonology          = *acousticModel_->phonology();
    verify(acousticModel_->silence() != Bliss::Phoneme::invalidId);

    // transition models by index (src/Am/TransitionModel.hh:72-78); the states name them by that index
    const int        nModels = acousticModel_->nStateTransitions();
    std::vector<f32> tdp(size_t(nModels) * 4);
    for (int m = 0; m < nModels; ++m)
        for (int k = 0; k < Am::StateTransitionModel::nTransitionTypes; ++k)
            tdp[m * 4 + k] = (*acousticModel_->stateTransition(m))[k];

    std::vector<u32> wordOffsets(1, 0), emission, model;
    std::vector<f32> unigram;
    std::vector<u8>  regular;
    pronunciations_.clear();
    Bliss::Lexicon::LemmaPronunciationIterator lp, lpEnd;
    for (Core::tie(lp, lpEnd) = lexicon_->lemmaPronunciations(); lp != lpEnd; ++lp) {
        const Bliss::Pronunciation& pron = *(*lp)->pronunciation();
        for (u32 i = 0; i < pron.length(); ++i) {
            // the allophone of phoneme i: its context as the phonology cuts it, word boundary flags
            s16 boundary = 0;
            if (i == 0)
                boundary |= Am::Allophone::isInitialPhone;
            if (i + 1 == pron.length())
                boundary |= Am::Allophone::isFinalPhone;
            const Am::Allophone* allophone =
                    acousticModel_->allophoneAlphabet()->allophone(Am::Allophone(phonology(pron, i), boundary));
            verify(allophone);
            const Am::ClassicHmmTopology* topology = acousticModel_->hmmTopology(pron[i]);
            verify(topology && topology->nPhoneStates() > 0 && topology->nSubStates() > 0);
            const bool isSilence = pron[i] == acousticModel_->silence();
            for (int state = 0; state < topology->nPhoneStates(); ++state) {
                const u32 e = acousticModel_->emissionIndex(
                        acousticModel_->allophoneStateAlphabet()->allophoneState(allophone, state));
                for (int rep = 0; rep < topology->nSubStates(); ++rep) {
                    emission.push_back(e);
                    model.push_back(isSilence ? Am::TransitionModel::silence : Am::TransitionModel::phone0 + rep);
                }
            }
        }
        wordOffsets.push_back(emission.size());
        // WordPronunciationState::unigramScore (:475-477)
        Score       score = 0;
        Lm::History history(lm_->startHistory());
        Lm::addLemmaPronunciationScore(lm_, *lp, pronunciationScale, lm_->scale(), history, score);
        unigram.push_back(score);
        // Pronunciation::isRegularWord (:86-96): an evaluation token sequence, and none of them empty
        std::pair<Bliss::Lemma::EvaluationTokenSequenceIterator, Bliss::Lemma::EvaluationTokenSequenceIterator> ev =
                (*lp)->lemma()->evaluationTokenSequences();
        bool isRegular = ev.first != ev.second;
        for (; ev.first != ev.second; ++ev.first)
            if (ev.first->isEpsilon())
                isRegular = false;
        regular.push_back(isRegular ? 1 : 0);
        pronunciations_.push_back(*lp);
    }
    nEmissions_ = acousticModel_->nEmissions();

    rb_lexicon lx;
    lx.n_words         = pronunciations_.size();
    lx.word_offsets    = wordOffsets.data();
    lx.state_emission  = emission.data();
    lx.state_tdp_model = model.data();
    lx.n_models        = nModels;
    lx.tdp             = tdp.data();
    lx.entry_model     = Am::TransitionModel::entryM1;
    lx.unigram         = unigram.data();
    lx.word_regular    = regular.data();
    lx.single_word     = singleWordRecognition_ ? 1 : 0;
    rb_search_destroy(handle_);
    handle_ = 0;
    if (rb_search_create(&lx, device_, &handle_) != RB_OK) {
        error("rb_search_create: %s", rb_last_error());
        return false;
    }
    restart();
    return true;
}

void LinearSearch::setGrammar(Fsa::ConstAutomatonRef) {
    error("B200 linear search: grammars are not supported (the LM is folded into the word entry scores)");
}

void LinearSearch::restart() {
    scores_.clear();
    time_ = decodedTime_ = 0;
    words_.clear();
    times_.clear();
    am_.clear();
    lmScores_.clear();
}

void LinearSearch::feed(const Mm::FeatureScorer::Scorer& emissionScores) {
    require(emissionScores);
    require(emissionScores->nEmissions() >= nEmissions_);
    const size_t at = scores_.size();
    scores_.resize(at + nEmissions_);
    f32* row = scores_.data() + at;
    if (const DenseScoreRow* dense = dynamic_cast<const DenseScoreRow*>(emissionScores.get())) {
        std::memcpy(row, dense->scoreRow(), nEmissions_ * sizeof(f32));  // a b200 feature scorer: the row is there already
        ++nDenseRows_;
    }
    else {
        for (u32 e = 0; e < nEmissions_; ++e)
            row[e] = emissionScores->score(e);
        nScoreCalls_ += nEmissions_;
    }
    ++time_;
}

void LinearSearch::decode() const {
    if (decodedTime_ == time_)
        return;
    require(handle_);
    const int64_t frameOffsets[2] = {0, int64_t(time_)};
    if (rb_search_decode(handle_, scores_.data(), nEmissions_, frameOffsets, 1) != RB_OK)
        criticalError("rb_search_decode: %s", rb_last_error());
    words_.resize(time_);
    times_.resize(time_);
    am_.resize(time_);
    lmScores_.resize(time_);
    const long n = rb_search_traceback(handle_, 0, words_.data(), times_.data(), am_.data(), lmScores_.data());
    if (n < 0)
        criticalError("rb_search_traceback: %s", rb_last_error());
    words_.resize(n);
    times_.resize(n);
    am_.resize(n);
    lmScores_.resize(n);
    decodedTime_ = time_;
}

void LinearSearch::getCurrentBestSentence(Search::Traceback& result) const {
    result.clear();
    if (time_ == 0)
        return;
    decode();
    if (words_.empty())
        return;
    typedef Search::TracebackItem Item;
    result.push_back(Item(0, 0, Search::ScoreVector(0, 0), Item::Transit()));
    for (size_t i = 0; i < words_.size(); ++i)
        result.push_back(Item(pronunciations_[words_[i]], times_[i], Search::ScoreVector(am_[i], lmScores_[i]), Item::Transit()));
    // the closing item carries the sentence end score for the history of the last two words (:446-453)
    Lm::History h(lm_->startHistory());
    if (words_.size() >= 2)
        Lm::extendHistoryByLemmaPronunciation(lm_, pronunciations_[words_[words_.size() - 2]], h);
    Lm::extendHistoryByLemmaPronunciation(lm_, pronunciations_[words_.back()], h);
    result.push_back(Item(0, time_, Search::ScoreVector(am_.back(), lmScores_.back() + lm_->sentenceEndScore(h)), Item::Transit()));
    log("returning %zu words", words_.size());
}

void LinearSearch::resetStatistics() {
    nDenseRows_ = nScoreCalls_ = 0;
}

void LinearSearch::logStatistics() const {
    log("b200 linear search: %llu frames taken as dense score rows, %llu score() calls", (unsigned long long)nDenseRows_,
        (unsigned long long)nScoreCalls_);
}

void LinearSearch::getPartialSentence(Search::Traceback& result) {
    getCurrentBestSentence(result);
    restart();
}

Core::Ref<const Search::LatticeAdaptor> LinearSearch::getCurrentWordLattice() const {
    return Core::ref(new Lattice::WordLatticeAdaptor);
}
