/*
 * ref_search.cc -- C entry points that drive the REFERENCE's own Search::LinearSearch (src/Search/LinearSearch.cc,
 * compiled from where it lies into oracle/_ref/librasr_ref_search.so together with the Bliss lexicon, the Am state
 * model / transition model / state tying, the Lm scaling wrapper and the Fsa library they need).  It exists to pin
 * oracle/search_oracle.cc to the reference's object code.  TEST INFRASTRUCTURE ONLY.
 *
 * What is the reference's own code on this path: the lexicon parser (Bliss), phonology / allophone alphabets /
 * HMM topologies (Am::ClassicStateModel), the transition model read from the configuration
 * (Am::ScaledTransitionModel -> GlobalTransitionModel), the lookup state tying (Am::LutStateTying), the language
 * model scaling (Lm::LanguageModelScaling), LinearSearch and its Traceback.
 * What is a stand-in written here (the reference's versions need the decision-tree, mixture-set and ONNX stacks):
 *   - TableAcousticModel: the Am::AcousticModel interface over those parts (Am::ClassicAcousticModel does the same
 *     assembly in src/Am/ClassicAcousticModel.cc:95-170 but also loads a mixture set)
 *   - TableLm: a unigram Lm::LanguageModel whose scores come from a table (the reference's ARPA reader is the
 *     only unigram model it ships)
 *   - the members of Speech::ModelCombination that LinearSearch touches (the four-reference constructor and the
 *     setters of src/Speech/ModelCombination.cc:92-122; that file itself needs Nn::Module -> onnxruntime)
 *   - Am::Module_ / ClassicAcousticModel::paramType: referenced by the state-tying factory, never executed here
 */
#include <Am/AcousticModel.hh>
#include <Am/ClassicAcousticModel.hh>
#include <Am/ClassicHmmTopologySet.hh>
#include <Am/ClassicStateModel.hh>
#include <Am/ClassicStateTying.hh>
#include <Am/Module.hh>
#include <Am/TransitionModel.hh>
#include <Bliss/Lexicon.hh>
#include <Core/Application.hh>
#include <Core/Configuration.hh>
#include <Lm/LanguageModel.hh>
#include <Lm/ScaledLanguageModel.hh>
#include <Mm/Feature.hh>
#include <Mm/FeatureScorer.hh>
#include <Search/LinearSearch.hh>
#include <Search/Traceback.hh>
#include <Speech/ModelCombination.hh>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

extern "C" int   ref_init(const char* log_file);
extern "C" void* ref_mm_feature_scorer(void* handle);  // ref_host.cc

// ---------------------------------------------------------------------------------------------------------
// stand-ins for symbols whose home translation units cannot be compiled here (see the header comment)
namespace Speech {
const ModelCombination::Mode ModelCombination::complete = 0x3;

ModelCombination::ModelCombination(const Core::Configuration& c, Bliss::LexiconRef lexicon,
                                   Core::Ref<Am::AcousticModel> am, Core::Ref<Lm::ScaledLanguageModel> lm)
        : Core::Component(c), Mc::Component(c), pronunciationScale_(0), labelScorers_(1) {
    // the parameter of src/Speech/ModelCombination.cc:31-32, read under the same name
    pronunciationScale_ = Core::ParameterFloat("pronunciation-scale", "", 0.0)(c);
    lexicon_            = lexicon;
    acousticModel_      = am;
    languageModel_      = lm;
    if (acousticModel_)
        acousticModel_->setParentScale(scale());
    if (languageModel_)
        languageModel_->setParentScale(scale());
}
ModelCombination::~ModelCombination() {}
void ModelCombination::distributeScaleUpdate(const Mc::ScaleUpdate&) {}
void ModelCombination::getDependencies(Core::DependencySet&) const {}
}  // namespace Speech

namespace Am {
Module_::Module_() {}
static const Core::Choice    standInTyingChoice("lookup", ClassicAcousticModel::lutTying, Core::Choice::endMark());
const Core::ParameterChoice ClassicAcousticModel::paramType("type", &standInTyingChoice, "", ClassicAcousticModel::lutTying);
}  // namespace Am

namespace {

int indexOfSymbol(const std::string& s) {  // "w12" / "p3" -> 12 / 3; anything else -> -1
    if (s.size() < 2 || (s[0] != 'w' && s[0] != 'p'))
        return -1;
    char* end = 0;
    long  v   = strtol(s.c_str() + 1, &end, 10);
    return (*end == 0) ? (int)v : -1;
}

// ---------------------------------------------------------------------------------------------------------
class TableAcousticModel : public Am::AcousticModel {
    Bliss::LexiconRef                        lexicon_;
    Core::Ref<const Am::ClassicStateModel>   stateModel_;
    Core::Ref<Am::ScaledTransitionModel>     transitions_;
    Core::Ref<const Am::ClassicStateTying>   tying_;
    Bliss::Phoneme::Id                       silence_;
    u32                                      nEmissions_;

public:
    TableAcousticModel(const Core::Configuration& c, Bliss::LexiconRef lexicon, const int32_t* emissionOf,
                       int statesPerPhone, int silenceEmission, u32 nEmissions)
            : Core::Component(c), Am::AcousticModel(c), lexicon_(lexicon), nEmissions_(nEmissions) {
        Core::Ref<const Bliss::PhonemeInventory> pi = lexicon->phonemeInventory();
        const Bliss::Phoneme* sil = pi->phoneme("si");
        silence_                  = sil ? sil->id() : Bliss::Phoneme::invalidId;
        // the same selections as src/Am/ClassicAcousticModel.cc:99-121
        Am::ClassicHmmTopologySetRef hmm(new Am::ClassicHmmTopologySet(select("hmm"), silence_));
        Bliss::PhonemeInventoryRef   piMutable(const_cast<Bliss::PhonemeInventory*>(pi.get()));
        Am::ConstPhonologyRef        phonology(new Am::Phonology(select("phonology"), piMutable));
        Am::ConstAllophoneAlphabetRef allophones(new Am::AllophoneAlphabet(select("allophones"), phonology, lexicon));
        Am::ConstAllophoneStateAlphabetRef states(
                new Am::AllophoneStateAlphabet(select("allophone-states"), allophones, hmm));
        stateModel_ = Core::ref(new Am::ClassicStateModel(phonology, allophones, states, hmm));

        // lookup table of the state tying, one line per allophone state of the lexicon, in the reference's own
        // symbol syntax; read back by Am::LutStateTying (src/Am/ClassicStateTying.cc:186-230)
        Core::Configuration tc(select("state-tying"));
        {
            std::ofstream lut(Am::ClassicStateTying::paramFilename(tc).c_str());
            const Am::AllophoneAlphabet::AllophoneList& list = allophones->allophones();
            for (size_t a = 0; a < list.size(); ++a) {
                const Am::Allophone* allo = list[a];
                const bool           isSil = allo->central() == silence_;
                const int            n     = isSil ? 1 : statesPerPhone;
                const int            ph    = isSil ? -1 : indexOfSymbol(pi->phoneme(allo->central())->symbol().str());
                for (int s = 0; s < n; ++s) {
                    const int e = isSil ? silenceEmission : emissionOf[ph * statesPerPhone + s];
                    lut << states->toString(states->allophoneState(allo, s)) << " " << e << "\n";
                }
            }
        }
        tying_       = Core::ref(new Am::LutStateTying(tc, stateModel_));
        transitions_ = Core::ref(new Am::ScaledTransitionModel(select("tdp"), stateModel_));
        transitions_->setParentScale(scale());
        transitions_->load();
    }

    virtual void load(Mode) {}
    virtual void getDependencies(Core::DependencySet&) const {}
    virtual Core::Ref<Am::TransducerBuilder> createTransducerBuilder() const {
        return Core::Ref<Am::TransducerBuilder>();
    }
    virtual Core::Ref<const Bliss::PhonemeInventory> phonemeInventory() const {
        return stateModel_->getPhonemeInventory();
    }
    virtual Core::Ref<const Am::AllophoneAlphabet> allophoneAlphabet() const {
        return stateModel_->getAllophoneAlphabet();
    }
    virtual Core::Ref<const Am::AllophoneStateAlphabet> allophoneStateAlphabet() const {
        return stateModel_->getAllophoneStateAlphabet();
    }
    virtual Core::Ref<const Am::Phonology> phonology() const {
        return stateModel_->getPhonology();
    }
    virtual Bliss::Phoneme::Id silence() const {
        return silence_;
    }
    virtual Am::AllophoneStateIndex silenceAllophoneStateIndex() const {
        return Fsa::InvalidLabelId;
    }
    virtual Am::AllophoneStateIndex blankAllophoneStateIndex() const {
        return Fsa::InvalidLabelId;
    }
    virtual Core::Ref<Mm::AbstractMixtureSet> mixtureSet() {
        return Core::Ref<Mm::AbstractMixtureSet>();
    }
    virtual Core::Ref<const Mm::ScaledFeatureScorer> featureScorer() {
        return Core::Ref<const Mm::ScaledFeatureScorer>();
    }
    virtual Core::Ref<Mm::ScaledFeatureScorer> mutableFeatureScorer() {
        return Core::Ref<Mm::ScaledFeatureScorer>();
    }
    virtual bool setFeatureScorer(Core::Ref<Mm::ScaledFeatureScorer>) {
        return false;
    }
    virtual EmissionIndex emissionIndex(Am::AllophoneState s) const {
        return tying_->classify(s);
    }
    virtual EmissionIndex emissionIndex(Am::AllophoneStateIndex i) const {
        return tying_->classifyIndex(i);
    }
    virtual EmissionIndex nEmissions() const {
        return nEmissions_;
    }
    virtual StateTransitionIndex nStateTransitions() const {
        return transitions_->nModels();
    }
    virtual const Am::StateTransitionModel* stateTransition(StateTransitionIndex i) const {
        return (*transitions_)[i];
    }
    virtual StateTransitionIndex stateTransitionIndex(Am::AllophoneState s, s8 sub) const {
        return transitions_->classify(s, sub);
    }
    virtual StateTransitionIndex stateTransitionIndex(Am::AllophoneStateIndex i, s8 sub) const {
        return transitions_->classifyIndex(i, sub);
    }
    virtual Core::Ref<const Am::ClassicHmmTopologySet> hmmTopologySet() const {
        return stateModel_->getHmmTopologySet();
    }
    virtual const Am::ClassicHmmTopology* hmmTopology(Bliss::Phoneme::Id p) const {
        return stateModel_->hmmTopologySet().get(p);
    }
    virtual bool isAcrossWordModelEnabled() const {
        return false;
    }
};

// ---------------------------------------------------------------------------------------------------------
class TableLm : public Lm::LanguageModel, private Lm::SingletonHistoryManager {
    std::vector<float> scores_;  // by the number in the syntactic token's symbol "w<k>"
    virtual std::string format(Lm::HistoryHandle) const {
        return "table";
    }

public:
    TableLm(const Core::Configuration& c, Bliss::LexiconRef l, const float* scores, int n)
            : Core::Component(c), Lm::LanguageModel(c, l), scores_(scores, scores + n) {
        historyManager_ = this;
    }
    virtual Lm::History startHistory() const {
        return history(0);
    }
    virtual Lm::History extendedHistory(const Lm::History& h, Lm::Token) const {
        return h;
    }
    virtual Lm::Score score(const Lm::History&, Lm::Token w) const {
        const int k = indexOfSymbol(w->symbol().str());
        return (k >= 0 && k < (int)scores_.size()) ? scores_[k] : 0.0f;
    }
    virtual Lm::Score sentenceEndScore(const Lm::History&) const {
        return 0;
    }
};

// one row of a dense score matrix as the scorer LinearSearch::feed pulls from
class RowScorer : public Mm::FeatureScorer::ContextScorer {
    const float* row_;
    u32          n_;

public:
    RowScorer(const float* row, u32 n) : row_(row), n_(n) {}
    virtual Mm::EmissionIndex nEmissions() const {
        return n_;
    }
    virtual Mm::Score score(Mm::EmissionIndex e) const {
        return row_[e];
    }
};

// an alternative Search::SearchAlgorithm to drive instead of the reference's (the adapter of adapters/B200LinearSearch.cc
// registers itself here when oracle/_ref/libb200_search_adapter.so is loaded)
typedef Search::SearchAlgorithm* (*SearchFactory)(const Core::Configuration&);
SearchFactory alternativeFactory = 0;

struct SearchHandle {
    Core::Configuration                 config;
    Bliss::LexiconRef                   lexicon;
    Core::Ref<TableAcousticModel>       am;
    Core::Ref<Lm::ScaledLanguageModel>  lm;
    Search::SearchAlgorithm*            search = 0;
    std::vector<int32_t>                order;  // lemma-pronunciation order of the lexicon, as word numbers
};

}  // namespace

extern "C" {

/* The configuration under `selection` supplies everything the reference reads itself (set with ref_config_set):
 *   <sel>.lexicon.file  Bliss XML: phonemes "p<k>" and "si" (= silence), lemmata with orthography "w<k>"
 *   <sel>.acoustic-model.hmm.states-per-phone / state-repetitions, .tdp.{entry-m1,silence,state-0,state-1}.{loop,
 *   forward,skip,exit}, .tdp.scale, .state-tying.file (scratch: written here, read by Am::LutStateTying)
 *   <sel>.lm.scale, <sel>.pronunciation-scale
 * emission_of[k * states_per_phone + s]: emission of state s of phoneme p<k>; silence_emission: of "si".
 * unigram[k]: unscaled LM score of the syntactic token w<k>. */
void ref_search_set_factory(void* factory) {
    alternativeFactory = (SearchFactory)factory;
}

/* alternative = 0: the reference's Search::LinearSearch; 1: the algorithm of ref_search_set_factory */
void* ref_search_create(const char* selection, const int32_t* emission_of, int states_per_phone,
                        int silence_emission, int n_emissions, const float* unigram, int n_unigram, int alternative) {
    ref_init(0);
    SearchHandle* h = new SearchHandle();
    h->config       = Core::Configuration(Core::Application::us()->getConfiguration(), selection ? selection : "search");
    h->lexicon      = Bliss::Lexicon::create(Core::Configuration(h->config, "lexicon"));
    if (!h->lexicon) {
        delete h;
        return 0;
    }
    h->am = Core::ref(new TableAcousticModel(Core::Configuration(h->config, "acoustic-model"), h->lexicon, emission_of,
                                             states_per_phone, silence_emission, (u32)n_emissions));
    Core::Configuration      lmc(h->config, "lm");
    Core::Ref<Lm::LanguageModel> table(new TableLm(lmc, h->lexicon, unigram, n_unigram));
    h->lm     = Core::ref(new Lm::LanguageModelScaling(lmc, table));
    if (alternative && !alternativeFactory) {
        delete h;
        return 0;
    }
    h->search = alternative ? alternativeFactory(Core::Configuration(h->config, "recognizer"))
                            : new Search::LinearSearch(Core::Configuration(h->config, "recognizer"));
    Speech::ModelCombination mc(h->config, h->lexicon, h->am, h->lm);
    if (!h->search->setModelCombination(mc)) {
        delete h->search;
        delete h;
        return 0;
    }
    Bliss::Lexicon::LemmaPronunciationIterator it, end;
    for (Core::tie(it, end) = h->lexicon->lemmaPronunciations(); it != end; ++it)
        h->order.push_back(indexOfSymbol((*it)->lemma()->preferredOrthographicForm().str()));
    return h;
}

void ref_search_destroy(void* handle) {
    SearchHandle* h = (SearchHandle*)handle;
    if (!h)
        return;
    delete h->search;
    delete h;
}

/* word numbers (the k of "w<k>", -1 for the silence lemma) in the order LinearSearch visits the pronunciations */
long ref_search_order(void* handle, int32_t* out, long capacity) {
    SearchHandle* h = (SearchHandle*)handle;
    for (size_t i = 0; i < h->order.size() && (long)i < capacity; ++i)
        out[i] = h->order[i];
    return (long)h->order.size();
}

/* flat view of what LinearSearch::Pronunciation (src/Search/LinearSearch.cc:32-84) derives from the models for
 * pronunciation i of ref_search_order: emission and transition-model index of every state */
long ref_search_states(void* handle, long i, int32_t* emission, int32_t* tdp_model, long capacity) {
    SearchHandle* h = (SearchHandle*)handle;
    Bliss::Lexicon::LemmaPronunciationIterator it, end;
    Core::tie(it, end) = h->lexicon->lemmaPronunciations();
    it += i;
    if (it >= end)
        return -1;
    Search::LinearSearch::Pronunciation p(*it, h->am.get());
    long                                n = 0;
    for (Search::LinearSearch::Pronunciation::MixtureVector::iterator m = p.mixtures().begin();
         m != p.mixtures().end(); ++m, ++n) {
        if (n >= capacity)
            continue;
        emission[n]  = (int32_t)m->mixture;
        tdp_model[n] = -1;
        for (int t = 0; t < h->am->nStateTransitions(); ++t)
            if (h->am->stateTransition(t) == m->stateTransitionModel)
                tdp_model[n] = t;
    }
    return n;
}

/* the transition scores LinearSearch reads: out[model * 4 + {loop, forward, skip, exit}], 5 models
 * (src/Am/TransitionModel.hh:72-78: entry-m1, entry-m2, silence, state-0, state-1) */
void ref_search_tdps(void* handle, float* out) {
    SearchHandle* h = (SearchHandle*)handle;
    for (int t = 0; t < 5; ++t)
        for (int k = 0; k < 4; ++k)
            out[t * 4 + k] = (*h->am->stateTransition(t))[k];
}

/* every item of the traceback as getCurrentBestSentence returns it (after ref_search_run): word number (-2 for the
 * items without a pronunciation), time, acoustic and lm score */
long ref_search_items(void* handle, int32_t* words, int32_t* times, float* am, float* lm, long capacity) {
    SearchHandle*     h = (SearchHandle*)handle;
    Search::Traceback tb;
    h->search->getCurrentBestSentence(tb);
    for (size_t i = 0; i < tb.size() && (long)i < capacity; ++i) {
        words[i] = tb[i].pronunciation ? indexOfSymbol(tb[i].pronunciation->lemma()->preferredOrthographicForm().str()) : -2;
        times[i] = (int32_t)tb[i].time;
        am[i]    = tb[i].score.acoustic;
        lm[i]    = tb[i].score.lm;
    }
    return (long)tb.size();
}

/* restart, feed every row of scores [T x n_emissions], getCurrentBestSentence.  Outputs (capacity each): the
 * traceback items that carry a pronunciation, chronological: word number, end frame, acoustic score, lm score.
 * final_scores[2]: acoustic and lm score of the closing item (lm includes the sentence-end score). */
long ref_search_run(void* handle, const float* scores, long T, int n_emissions, int32_t* words, int32_t* times,
                    float* am, float* lm, long capacity, float* final_scores) {
    SearchHandle* h = (SearchHandle*)handle;
    h->search->restart();
    for (long t = 0; t < T; ++t) {
        Mm::FeatureScorer::Scorer s(new RowScorer(scores + (size_t)t * n_emissions, (u32)n_emissions));
        h->search->feed(s);
    }
    Search::Traceback tb;
    h->search->getCurrentBestSentence(tb);
    long n = 0;
    for (size_t i = 0; i < tb.size(); ++i) {
        if (!tb[i].pronunciation) {
            if (i + 1 == tb.size() && final_scores) {
                final_scores[0] = tb[i].score.acoustic;
                final_scores[1] = tb[i].score.lm;
            }
            continue;
        }
        if (n < capacity) {
            words[n] = indexOfSymbol(tb[i].pronunciation->lemma()->preferredOrthographicForm().str());
            times[n] = (int32_t)tb[i].time;
            am[n]    = tb[i].score.acoustic;
            lm[n]    = tb[i].score.lm;
        }
        ++n;
    }
    return n;
}

/* The recognizer's loop (src/Speech/Recognizer.cc:271-281 processFeature, :197-205 finish) over one segment: the
 * feature scorer of ref_mm_create (mm_handle) turns every feature vector into a scorer object -- buffered scorers get
 * addFeature() until bufferFilled(), then getScorer() per further frame and flush() at the end -- and every scorer
 * is fed to the search.  Read the result with ref_search_items. */
long ref_search_run_features(void* handle, void* mm_handle, const float* feats, long T, int dim) {
    SearchHandle*            h  = (SearchHandle*)handle;
    const Mm::FeatureScorer& fs = *static_cast<const Mm::FeatureScorer*>(ref_mm_feature_scorer(mm_handle));
    h->search->restart();
    fs.reset();
    long fed = 0;
    for (long t = 0; t < T; ++t) {
        Core::Ref<const Mm::Feature> f(new Mm::Feature(Mm::FeatureVector(feats + (size_t)t * dim, feats + (size_t)(t + 1) * dim)));
        if (fs.isBuffered() && !fs.bufferFilled())
            fs.addFeature(f);
        else {
            h->search->feed(fs.getScorer(f));
            ++fed;
        }
    }
    while (fs.isBuffered() && !fs.bufferEmpty()) {
        h->search->feed(fs.flush());
        ++fed;
    }
    return fed;
}

}  // extern "C"
