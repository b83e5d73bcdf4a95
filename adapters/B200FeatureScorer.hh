/*
 * B200FeatureScorer -- Mm::FeatureScorer adapter that forwards to librasr_b200.so (include/rasr_b200.h).
 *
 * Written against the RASR headers (src/Mm/FeatureScorer.hh:28-167); compiled inside a RASR checkout
 * (see INTEGRATION.md), not in this repository.  It is a whole-segment buffered scorer in the style of
 * src/Onnx/OnnxFeatureScorer.{hh,cc}: addFeature() collects the segment, the first flush()/score() runs ONE
 * dense rb_gmm_score over all buffered frames, ContextScorer::score(e) is then a lookup into the T x nMix
 * matrix.  Selected with  feature-scorer-type = b200-batch-float | b200-diagonal-maximum | b200-diagonal-sum
 * | b200-batch-tensor | b200-batch-int | b200-preselection-batch-float | b200-preselection-batch-int | b200-SIMD-diagonal-maximum (GmmFeatureScorer -> the rb_gmm calls)  or  b200-nn-batch-feature-scorer
 * (NnFeatureScorer -> the rb_nn calls, the drop-in for src/Nn/BatchFeatureScorer.{hh,cc}).
 */
#ifndef _B200_FEATURE_SCORER_HH
#define _B200_FEATURE_SCORER_HH

#include <Core/Parameter.hh>
#include <Mm/FeatureScorer.hh>
#include <Mm/MixtureSet.hh>
#include <vector>

#include "B200HostBuffer.hh"
#include "rasr_b200.h"

namespace B200 {

/** Implemented by the ContextScorers the b200 feature scorers hand out: the whole row of emission scores of the frame
 *  at once.  A consumer that wants every score of a frame anyway (B200::LinearSearch::feed) takes the row instead of
 *  nEmissions() virtual score(e) calls. */
class DenseScoreRow {
public:
    virtual ~DenseScoreRow() {}
    /** nEmissions() scores of this scorer's frame; valid until the feature scorer is reset() */
    virtual const f32* scoreRow() const = 0;
};

/** Buffering half of the adapters: collects the segment, scores all frames not scored yet in one call of the
 *  subclass, answers ContextScorer::score(e) from the dense matrix. */
class FeatureScorer : public Mm::FeatureScorer {
public:
    static const Core::ParameterInt paramDevice;      // CUDA ordinal
    static const Core::ParameterInt paramBufferSize;  // frames buffered before bufferFilled()

    FeatureScorer(const Core::Configuration& c);
    virtual ~FeatureScorer() {}

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

    Mm::Score  score(u32 segment, u32 frame, Mm::EmissionIndex e) const;
    const f32* row(u32 segment, u32 frame) const;
    void       scoreBufferedFrames() const;

protected:
    /** scores [T x nMixtures_] of feats [T x dimension_], both row-major */
    virtual void scoreFrames(const f32* feats, u32 T, f32* scores) const = 0;

    Mm::EmissionIndex nMixtures_;
    u32               dimension_;
    u32               bufferSize_;

private:
    // all methods of the interface are const => the state is mutable (src/Mm/BatchFeatureScorer.hh:164-166)
    // page-locked (B200HostBuffer.hh): the library's slab pipeline copies from / into these at the PCIe rate
    mutable HostBuffer features_;  // T x D row-major, the segment so far
    mutable HostBuffer scores_;    // T x nMix row-major, rows [0, nScored_)
    mutable u32              nScored_;
    mutable u32              nextFrame_;  // oldest frame that has no scorer yet
    mutable u32              segment_;    // guards delayed score() calls across reset()
};

class GmmFeatureScorer : public FeatureScorer {
public:
    static const Core::ParameterBool  paramContraction;         // FMA contraction like the default CPU build
    static const Core::ParameterFloat paramMixtureWeightScale;  // as Mm::GaussDiagonalMaximumFeatureScorer
    static const Core::ParameterFloat paramGaussianScale;
    static const Core::ParameterInt   paramClusters;              // density-clustering.* of the preselection scorer
    static const Core::ParameterInt   paramSelectClusters;
    static const Core::ParameterInt   paramClusteringIterations;
    static const Core::ParameterFloat paramBackoffScore;

    GmmFeatureScorer(const Core::Configuration& c, Core::Ref<const Mm::MixtureSet> mixtureSet, rb_gmm_mode mode);
    virtual ~GmmFeatureScorer();

protected:
    virtual void scoreFrames(const f32* feats, u32 T, f32* scores) const;

private:
    rb_gmm* handle_;
};

template<rb_gmm_mode mode>
class FeatureScorerOf : public GmmFeatureScorer {
public:
    FeatureScorerOf(const Core::Configuration& c, Core::Ref<const Mm::MixtureSet> m)
            : Core::Component(c), GmmFeatureScorer(c, m, mode) {}
};

/** Feed-forward network scorer: score(e) = -(w_e.h + b_e - priori-scale * logprior_e), the top-layer softmax is not
 *  evaluated (src/Nn/BatchFeatureScorer.cc:45-171, LinearAndActivationLayer.cc:154-160).  The network is described
 *  by the reference's own configuration keys (neural-network.links, <layer>.layer-type / dimension-* / links,
 *  parameters-old: see B200NnNetwork.hh), so a configuration written for nn-batch-feature-scorer works unchanged;
 *  without a layer description the legacy `parameter-files` list (one Math::Matrix per layer, row = output unit,
 *  column 0 = bias, src/Nn/LinearLayer.cc:383-424) plus `hidden-activation` is used.  The prior comes from `prior-file` or, like
 *  Prior::setFromMixtureSet (src/Nn/Prior.cc:158-188), from the mixture weights.  Emission classes map to network
 *  outputs through the Nn::ClassLabelWrapper file (class-labels.load-from-file; disregarded classes score FLT_MAX,
 *  src/Nn/BatchFeatureScorer.cc:163-169), else one-to-one. */
class NnFeatureScorer : public FeatureScorer {
public:
    static const Core::ParameterStringVector paramParameterFiles;  // "parameters-old" of the layers, bottom to top
    static const Core::ParameterString       paramHiddenActivation;
    static const Core::ParameterString       paramPriorFile;
    static const Core::ParameterFloat        paramPrioriScale;
    static const Core::ParameterString       paramClassLabelFile;  // class-labels.load-from-file
    static const Core::ParameterBool         paramBf16;

    NnFeatureScorer(const Core::Configuration& c, Core::Ref<const Mm::MixtureSet> mixtureSet);
    virtual ~NnFeatureScorer();

protected:
    virtual void scoreFrames(const f32* feats, u32 T, f32* scores) const;

private:
    void readParameterFiles(const Core::Configuration& c, struct NnNetwork& net) const;

    rb_nn* handle_;
};

}  // namespace B200

#endif
