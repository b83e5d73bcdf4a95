#include "B200FeatureScorer.hh"

#include <Mm/Feature.hh>
#include <Mm/GaussDensity.hh>
#include <Mm/Mixture.hh>

using namespace B200;

const Core::ParameterInt   FeatureScorer::paramDevice("device", "CUDA device ordinal", 0, 0);
const Core::ParameterInt   FeatureScorer::paramBufferSize("buffer-size", "frames buffered per dense scoring launch (default: the whole segment)", Core::Type<s32>::max, 1);
const Core::ParameterBool  FeatureScorer::paramContraction("fma-contraction", "reproduce a CPU build with -ffp-contract=fast (gcc default) instead of a strict one", true);
const Core::ParameterFloat FeatureScorer::paramMixtureWeightScale("mixture-weight-scale", "scale of the -log mixture weights (diagonal scorers)", 1.0);
const Core::ParameterFloat FeatureScorer::paramGaussianScale("gaussian-scale", "scale of the Gaussian exponent (diagonal scorers)", 1.0);

class FeatureScorer::ContextScorer : public Mm::FeatureScorer::ContextScorer {
public:
    ContextScorer(const FeatureScorer* parent, u32 segment, u32 frame)
            : parent_(parent), segment_(segment), frame_(frame) {}
    virtual Mm::EmissionIndex nEmissions() const {
        return parent_->nMixtures();
    }
    virtual Mm::Score score(Mm::EmissionIndex e) const {
        return parent_->score(segment_, frame_, e);
    }

private:
    const FeatureScorer* parent_;
    u32                  segment_, frame_;
};

FeatureScorer::FeatureScorer(const Core::Configuration& c, Core::Ref<const Mm::MixtureSet> ms, rb_gmm_mode mode)
        : Core::Component(c),
          Mm::FeatureScorer(c),
          handle_(0),
          nMixtures_(ms->nMixtures()),
          dimension_(ms->dimension()),
          bufferSize_(paramBufferSize(c)),
          nScored_(0),
          nextFrame_(0),
          segment_(0) {
    // flatten Mm::MixtureSet (src/Mm/MixtureSet.hh:123-201) into the rb_mixture_set view
    std::vector<u32> mixOffsets(1, 0), mixDensity, densMean(ms->nDensities()), densCov(ms->nDensities());
    std::vector<f64> mixLogWeight;
    for (Mm::MixtureIndex m = 0; m < ms->nMixtures(); ++m) {
        const Mm::Mixture& mixture = *ms->mixture(m);
        for (size_t d = 0; d < mixture.nDensities(); ++d) {
            mixDensity.push_back(mixture.densityIndex(d));
            mixLogWeight.push_back(mixture.logWeight(d));
        }
        mixOffsets.push_back(mixDensity.size());
    }
    for (Mm::DensityIndex d = 0; d < ms->nDensities(); ++d) {
        densMean[d] = ms->density(d)->meanIndex();
        densCov[d]  = ms->density(d)->covarianceIndex();
    }
    std::vector<f32> means(size_t(ms->nMeans()) * dimension_), variances(size_t(ms->nCovariances()) * dimension_);
    for (Mm::MeanIndex i = 0; i < ms->nMeans(); ++i)
        std::copy(ms->mean(i)->begin(), ms->mean(i)->end(), means.begin() + size_t(i) * dimension_);
    for (Mm::CovarianceIndex i = 0; i < ms->nCovariances(); ++i) {
        const Mm::DiagonalCovariance* cov = dynamic_cast<const Mm::DiagonalCovariance*>(ms->covariance(i));
        if (!cov)
            criticalError("b200 feature scorers support diagonal covariances only");
        std::copy(cov->diagonal().begin(), cov->diagonal().end(), variances.begin() + size_t(i) * dimension_);
    }
    rb_mixture_set view = {dimension_, nMixtures_, u32(ms->nDensities()), u32(ms->nMeans()), u32(ms->nCovariances()),
                           mixOffsets.data(), mixDensity.data(), mixLogWeight.data(), densMean.data(),
                           densCov.data(), means.data(), variances.data()};
    if (rb_gmm_create(&view, mode, paramMixtureWeightScale(c), paramGaussianScale(c), paramContraction(c),
                      paramDevice(c), &handle_) != RB_OK)
        criticalError("rasr_b200: %s", rb_last_error());
    log("b200 feature scorer: %d mixtures, %d densities, dimension %d on device %d",
        int(nMixtures_), int(ms->nDensities()), int(dimension_), int(paramDevice(c)));
}

FeatureScorer::~FeatureScorer() {
    rb_gmm_destroy(handle_);
}

void FeatureScorer::getFeatureDescription(Mm::FeatureDescription& description) const {
    description.mainStream().setValue(Mm::FeatureDescription::nameDimension, dimension_);
}

void FeatureScorer::reset() const {
    features_.clear();
    scores_.clear();
    nScored_ = nextFrame_ = 0;
    ++segment_;
}

void FeatureScorer::addFeature(const Mm::FeatureVector& f) const {
    require(!bufferFilled());
    if (f.size() != dimension_)
        criticalError("feature has dimension %zu, mixture set expects %d", f.size(), int(dimension_));
    features_.insert(features_.end(), f.begin(), f.end());
}

Mm::FeatureScorer::Scorer FeatureScorer::getScorer(const Mm::FeatureVector& f) const {
    if (f.size() != dimension_)
        criticalError("feature has dimension %zu, mixture set expects %d", f.size(), int(dimension_));
    features_.insert(features_.end(), f.begin(), f.end());
    return flush();
}

Mm::FeatureScorer::Scorer FeatureScorer::flush() const {
    require(!bufferEmpty());
    return Scorer(new ContextScorer(this, segment_, nextFrame_++));
}

Mm::FeatureScorer::Scorer FeatureScorer::getTimeIndexedScorer(u32 time) const {
    require_lt(time, features_.size() / dimension_);
    return Scorer(new ContextScorer(this, segment_, time));
}

void FeatureScorer::scoreBufferedFrames() const {
    const u32 T = features_.size() / dimension_;
    scores_.resize(size_t(T) * nMixtures_);
    if (rb_gmm_score(handle_, features_.data() + size_t(nScored_) * dimension_, T - nScored_,
                     scores_.data() + size_t(nScored_) * nMixtures_, 0) != RB_OK)
        criticalError("rasr_b200: %s", rb_last_error());
    nScored_ = T;
}

Mm::Score FeatureScorer::score(u32 segment, u32 frame, Mm::EmissionIndex e) const {
    require_eq(segment, segment_);  // a scorer must not be used after reset() (same rule as OnnxFeatureScorer)
    require_lt(e, nMixtures_);
    if (frame >= nScored_)
        scoreBufferedFrames();
    return scores_[size_t(frame) * nMixtures_ + e];
}
