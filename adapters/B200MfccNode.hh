/*
 * B200MfccNode -- Flow node "b200-mfcc": one node that replaces the whole network of
 * src/Tools/FeatureExtraction/share/mfcc.flow:8-34 (+ derivationWithRegression.flow:7-27):
 *
 *   samples --> [ b200-mfcc ] --> vector-f32 features (one packet per frame, Timestamp per frame)
 *
 * Written against the RASR headers (src/Flow/Node.hh:214-228); compiled inside a RASR checkout.
 * Batch node in the style of Nn::NeuralNetworkForwardNode::work (src/Nn/NeuralNetworkForwardNode.cc:187-256):
 * pull sample packets until EOS, rb_frontend_finish() runs the kernels over the segment, then one
 * Flow::Vector<f32> is emitted per work() call.
 */
#ifndef _B200_MFCC_NODE_HH
#define _B200_MFCC_NODE_HH

#include <Flow/Node.hh>
#include <Flow/Vector.hh>
#include <vector>

#include "rasr_b200.h"

namespace B200 {

class MfccNode : public Flow::SleeveNode {
public:
    static const Core::ParameterFloat paramAlpha;            // signal-preemphasis alpha
    static const Core::ParameterFloat paramLength;           // signal-window length
    static const Core::ParameterFloat paramShift;            // signal-window shift
    static const Core::ParameterFloat paramMaximumInputSize; // signal-real-fast-fourier-transform maximum-input-size
    static const Core::ParameterFloat paramFilterWidth;      // signal-filterbank filter-width
    static const Core::ParameterInt   paramNrOutputs;        // signal-cosine-transform nr-outputs
    static const Core::ParameterBool  paramDerivatives;      // append delta / delta-delta
    static const Core::ParameterInt   paramDevice;
    static const Core::Choice          choiceWindowType;          // signal-window type
    static const Core::ParameterChoice paramWindowType;
    static const Core::ParameterBool  paramDcDetection;           // signal-dc-detection in front of the chain
    static const Core::ParameterFloat paramMinDcLength;           // its parameters, named as in DcDetection.cc:231-241
    static const Core::ParameterFloat paramMaxDcIncrement;
    static const Core::ParameterFloat paramMinNonDcSegmentLength;
    static const Core::ParameterInt   paramMaximalOutputSize;

    static std::string filterName() {
        return "b200-mfcc";
    }
    MfccNode(const Core::Configuration& c);
    virtual ~MfccNode();

    virtual bool setParameter(const std::string& name, const std::string& value);
    virtual bool configure();
    virtual bool work(Flow::PortId p);

protected:  // b200-audio-feature-scorer (B200Nodes.hh) builds on this node
    bool ensureHandle();
    void computeSegment();

    rb_frontend_cfg     cfg_;
    rb_dc_cfg           dc_;
    bool                dcDetection_;
    rb_frontend*        handle_;
    bool                dirty_;      // parameters changed since the handle was created
    bool                segmentOpen_;
    std::vector<f32>    feats_;      // T x featDim of the segment being emitted
    std::vector<f64>    tStart_, tEnd_;
    size_t              nFrames_, nextFrame_;
    int                 featDim_;
};

}  // namespace B200

#endif
