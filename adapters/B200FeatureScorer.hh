/*
 * B200FeatureScorer -- Mm::FeatureScorer adapter that forwards to librasr_b200.so (include/rasr_b200.h).
 *
 * Written against the RASR headers (src/Mm/FeatureScorer.hh:28-167); compiled inside a RASR checkout
 * (see INTEGRATION.md), not in this repository.  It is a whole-segment buffered scorer in the style of
 * src/Onnx/OnnxFeatureScorer.{hh,cc}: addFeature() collects the segment, the first flush()/score() runs ONE
 * dense rb_gmm_score over all buffered frames, ContextScorer::score(e) is then a lookup into the T x nMix
 * matrix.  Selected with  feature-scorer-type = b200-batch-float | b200-diagonal-maximum | b200-diagonal-sum
 * | b200-batch-tensor | b200-batch-int.
 */
#ifndef _B200_FEATURE_SCORER_HH
#define _B200_FEATURE_SCORER_HH

#include <Core/Parameter.hh>
#include <Mm/FeatureScorer.hh>
#include <Mm/MixtureSet.hh>
#include <vector>

#include "rasr_b200.h"

namespace B200 {

class FeatureScorer : public Mm::FeatureScorer {
public:
    static const Core::ParameterInt   paramDevice;              // CUDA ordinal
    static const Core::ParameterInt   paramBufferSize;          // frames buffered before bufferFilled()
    static const Core::ParameterBool  paramContraction;         // FMA contraction like the default CPU build
    static const Core::ParameterFloat paramMixtureWeightScale;  // as Mm::GaussDiagonalMaximumFeatureScorer
    static const Core::ParameterFloat paramGaussianScale;

    FeatureScorer(const Core::Configuration& c, Core::Ref<const Mm::MixtureSet> mixtureSet, rb_gmm_mode mode);
    virtual ~FeatureScorer();

    virtual Mm::EmissionIndex nMixtures() const {
        return nMixtures_;
    }
    virtual void getFeatureDescription(Mm::FeatureDescription& description) const;

    // the scorer returned is the one of the OLDEST buffered frame (src/Mm/BatchFeatureScorer.hh:81-91)
    virtual Scorer getScorer(Core::Ref<const Mm::Feature> f) const {
        return getScorer(*f->mainStream());
    }
    virtual Scorer getScorer(const Mm::FeatureVector& f) const;

    virtual void reset() const;
    virtual void finalize() const {}
    virtual bool isBuffered() const {
        return true;
    }
    virtual void addFeature(const Mm::FeatureVector& f) const;
    virtual void addFeature(Core::Ref<const Mm::Feature> f) const {
        addFeature(*f->mainStream());
    }
    virtual Scorer flush() const;
    virtual bool   bufferFilled() const {
        return features_.size() / dimension_ - nextFrame_ >= bufferSize_;
    }
    virtual bool bufferEmpty() const {
        return nextFrame_ >= features_.size() / dimension_;
    }
    virtual u32 bufferSize() const {
        return bufferSize_;
    }
    virtual bool hasTimeIndexedCache() const {
        return true;
    }
    virtual Scorer getTimeIndexedScorer(u32 time) const;

private:
    class ContextScorer;
    friend class ContextScorer;

    Mm::Score score(u32 segment, u32 frame, Mm::EmissionIndex e) const;
    void      scoreBufferedFrames() const;

    rb_gmm*           handle_;
    Mm::EmissionIndex nMixtures_;
    u32               dimension_;
    u32               bufferSize_;
    // all methods of the interface are const => the state is mutable (src/Mm/BatchFeatureScorer.hh:164-166)
    mutable std::vector<f32> features_;  // T x D row-major, the segment so far
    mutable std::vector<f32> scores_;    // T x nMix row-major, rows [0, nScored_)
    mutable u32              nScored_;
    mutable u32              nextFrame_;  // oldest frame that has no scorer yet
    mutable u32              segment_;    // guards delayed score() calls across reset()
};

template<rb_gmm_mode mode>
class FeatureScorerOf : public FeatureScorer {
public:
    FeatureScorerOf(const Core::Configuration& c, Core::Ref<const Mm::MixtureSet> m)
            : Core::Component(c), FeatureScorer(c, m, mode) {}
};

}  // namespace B200

#endif
