/*
 * B200Nodes -- the other Flow-side wrappers of the C ABI, next to B200MfccNode:
 *
 *   b200-neural-network-forward   drop-in for Nn::NeuralNetworkForwardNode ("neural-network-forward",
 *                                 src/Nn/NeuralNetworkForwardNode.{hh,cc}): vector-f32 features in, the top layer's
 *                                 output (softmax evaluated for linear+softmax) out, one packet per frame carrying the
 *                                 input packet's Timestamp.  Network and prior are configured with the reference's own
 *                                 keys (B200NnNetwork.hh; prior-file / priori-scale, src/Nn/Prior.cc:24-30).
 *   b200-feature-postprocessing   signal-normalization -> signal-vector-f32-sequence-concatenation ->
 *                                 signal-matrix-multiplication-f32 in one node (the wiring of
 *                                 processing.standard_system.flow:25-27 + lda.flow:11-19), every stage optional.
 *   b200-audio-feature-scorer     samples in, one dense score vector per frame out, in +log space like
 *                                 Speech::FeatureScorerNode (src/Speech/FeatureScorerNode.cc:95-111): MFCC chain and
 *                                 GMM scorer run back to back on the device (rb_pipeline_score), the features never
 *                                 leave HBM.  Mixture set and scorer are configured like the reference's node:
 *                                 selections "mixture-set" and "feature-scorer" (FeatureScorerNode.cc:27-29).
 *
 * All three are batch nodes in the style of NeuralNetworkForwardNode::work with dynamic-buffer=true
 * (src/Nn/NeuralNetworkForwardNode.cc:187-256): pull until EOS, one call into the engine for the segment, then one
 * packet per work() call.  Segment buffers live in page-locked memory (B200HostBuffer.hh).
 */
#ifndef _B200_NODES_HH
#define _B200_NODES_HH

#include <Flow/Node.hh>
#include <Flow/Vector.hh>
#include <Mm/MixtureSet.hh>
#include <vector>

#include "B200HostBuffer.hh"
#include "B200MfccNode.hh"
#include "rasr_b200.h"

namespace B200 {

/** flattens Mm::MixtureSet (src/Mm/MixtureSet.hh:123-201) into the arrays behind an rb_mixture_set view */
struct FlatMixtureSet {
    std::vector<u32> mixOffsets, mixDensity, densMean, densCov;
    std::vector<f64> mixLogWeight;
    std::vector<f32> means, variances;
    rb_mixture_set   view;
    /** false (with `what` set) if the set holds something other than diagonal Gaussians */
    bool assign(const Mm::MixtureSet& ms, std::string& what);
};

/** common part: buffers the vector-f32 packets of a segment, emits the rows of `out_` with the kept time stamps */
class SegmentNode : public Flow::SleeveNode {
public:
    SegmentNode(const Core::Configuration& c);
    virtual bool configure();
    virtual bool work(Flow::PortId p);

protected:
    /** in_ holds nIn_ rows of dimIn_ floats; fill out_ with nOut_ == nIn_ rows of dimOut_ floats */
    virtual bool processSegment() = 0;

    HostBuffer                   in_, out_;
    std::vector<Flow::Timestamp> times_;
    u32                          dimIn_, dimOut_;
    size_t                       nIn_, next_;
};

class NnForwardNode : public SegmentNode {
public:
    static const Core::ParameterString paramId;
    static const Core::ParameterInt    paramBufferSize;
    static const Core::ParameterBool   paramCheckValues;
    static const Core::ParameterBool   paramDynamicBuffer;
    static const Core::ParameterString paramPriorFile;
    static const Core::ParameterFloat  paramPrioriScale;
    static const Core::ParameterBool   paramBf16;
    static const Core::ParameterInt    paramDevice;

    static std::string filterName() {
        return "b200-neural-network-forward";
    }
    NnForwardNode(const Core::Configuration& c);
    virtual ~NnForwardNode();
    virtual bool setParameter(const std::string& name, const std::string& value);

protected:
    virtual bool processSegment();

private:
    rb_nn* handle_;
    bool   checkValues_;
};

class PostprocessingNode : public SegmentNode {
public:
    static const Core::ParameterString paramNormalizationType;    // none | mean | mean-and-variance
    static const Core::ParameterString paramNormalizationLength;  // frames or "infinite" (signal-normalization length)
    static const Core::ParameterString paramNormalizationRight;   // signal-normalization right
    static const Core::ParameterInt    paramWindowMaxSize;        // sequence concatenation max-size; 0: none
    static const Core::ParameterInt    paramWindowRight;          // sequence concatenation right
    static const Core::ParameterString paramMatrixFile;           // signal-matrix-multiplication-f32 file
    static const Core::ParameterBool   paramContraction;
    static const Core::ParameterInt    paramDevice;

    static std::string filterName() {
        return "b200-feature-postprocessing";
    }
    PostprocessingNode(const Core::Configuration& c);
    virtual ~PostprocessingNode();
    virtual bool setParameter(const std::string& name, const std::string& value);

protected:
    virtual bool processSegment();

private:
    rb_postproc_cfg  cfg_;
    std::string      normType_, normLength_, normRight_, matrixFile_;
    std::vector<f32> matrix_;
    rb_postproc*     handle_;
    u32              handleDim_;
};

class AudioScorerNode : public MfccNode {
public:
    static std::string filterName() {
        return "b200-audio-feature-scorer";
    }
    AudioScorerNode(const Core::Configuration& c);
    virtual ~AudioScorerNode();
    virtual bool work(Flow::PortId p);

private:
    bool ensureScorer();

    rb_gmm*             gmm_;
    u32                 nMixtures_;
    HostBuffer          samples_, scores_;
    std::vector<double> frameStart_, frameEnd_;
    double              segmentStart_;
    bool                haveStart_;
    size_t              nScoreFrames_, nextScore_;
};

}  // namespace B200

#endif
