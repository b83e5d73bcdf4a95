#include "B200MfccNode.hh"

using namespace B200;

const Core::ParameterFloat MfccNode::paramAlpha("alpha", "preemphasis weight", 1.0);
const Core::ParameterFloat MfccNode::paramLength("length", "window length in seconds", 0.025, 0.0);
const Core::ParameterFloat MfccNode::paramShift("shift", "window shift in seconds", 0.01, 0.0);
const Core::ParameterFloat MfccNode::paramMaximumInputSize("maximum-input-size", "FFT: longest input in seconds", 0.025, 0.0);
const Core::ParameterFloat MfccNode::paramFilterWidth("filter-width", "mel filter width", 268.258);
const Core::ParameterInt   MfccNode::paramNrOutputs("nr-outputs", "number of cepstral coefficients", 13, 1);
const Core::ParameterBool  MfccNode::paramDerivatives("derivatives", "append first and second order regression", true);
const Core::ParameterInt   MfccNode::paramDevice("device", "CUDA device ordinal", 0, 0);
// signal-window "type" (src/Signal/WindowFunction.cc:25-33)
const Core::Choice          MfccNode::choiceWindowType("hamming", RB_WINDOW_HAMMING, "rectangular", RB_WINDOW_RECTANGULAR, "hanning", RB_WINDOW_HANNING,
                                                        "periodic-hanning", RB_WINDOW_PERIODIC_HANNING, "bartlett", RB_WINDOW_BARTLETT,
                                                        "blackman", RB_WINDOW_BLACKMAN, "kaiser", RB_WINDOW_KAISER, Core::Choice::endMark());
const Core::ParameterChoice MfccNode::paramWindowType("window-type", &choiceWindowType, "type of window", RB_WINDOW_HAMMING);
// signal-dc-detection in front of the chain (src/Signal/DcDetection.cc:231-241; values of samples.flow:34-35)
const Core::ParameterBool  MfccNode::paramDcDetection("dc-detection", "discard DC stretches of the input like signal-dc-detection", false);
const Core::ParameterFloat MfccNode::paramMinDcLength("min-dc-length", "minimum length (in seconds) of DC necesseary for the decision", .0125, 0);
const Core::ParameterFloat MfccNode::paramMaxDcIncrement("max-dc-increment", "interval with less variation taken as DC, 0 disables DC detection", 0.9, 0);
const Core::ParameterFloat MfccNode::paramMinNonDcSegmentLength("min-non-dc-segment-length", "smaller segments (given in seconds) are discarded", .026, 0);
const Core::ParameterInt   MfccNode::paramMaximalOutputSize("maximal-output-size", "maximal size of output", 4096, 1);

MfccNode::MfccNode(const Core::Configuration& c)
        : Core::Component(c), Flow::SleeveNode(c), handle_(0), dirty_(true), segmentOpen_(false), nFrames_(0), nextFrame_(0), featDim_(0) {
    rb_frontend_default_cfg(&cfg_);
    cfg_.preemphasis_alpha = paramAlpha(c);
    cfg_.window_length_s   = paramLength(c);
    cfg_.window_shift_s    = paramShift(c);
    cfg_.fft_max_input_s   = paramMaximumInputSize(c);
    cfg_.filter_width      = paramFilterWidth(c);
    cfg_.n_cepstra         = paramNrOutputs(c);
    cfg_.derivatives       = paramDerivatives(c) ? 1 : 0;
    cfg_.device            = paramDevice(c);
    cfg_.window_type       = paramWindowType(c);
    dcDetection_                    = paramDcDetection(c);
    dc_.min_dc_length_s             = paramMinDcLength(c);
    dc_.max_dc_increment            = paramMaxDcIncrement(c);
    dc_.min_non_dc_segment_length_s = paramMinNonDcSegmentLength(c);
    dc_.maximal_output_size         = paramMaximalOutputSize(c);
}

MfccNode::~MfccNode() {
    rb_frontend_destroy(handle_);
}

bool MfccNode::setParameter(const std::string& name, const std::string& value) {
    if (paramAlpha.match(name))
        cfg_.preemphasis_alpha = paramAlpha(value);
    else if (paramLength.match(name))
        cfg_.window_length_s = paramLength(value);
    else if (paramShift.match(name))
        cfg_.window_shift_s = paramShift(value);
    else if (paramMaximumInputSize.match(name))
        cfg_.fft_max_input_s = paramMaximumInputSize(value);
    else if (paramFilterWidth.match(name))
        cfg_.filter_width = paramFilterWidth(value);
    else if (paramNrOutputs.match(name))
        cfg_.n_cepstra = paramNrOutputs(value);
    else if (paramDerivatives.match(name))
        cfg_.derivatives = paramDerivatives(value) ? 1 : 0;
    else if (paramDevice.match(name))
        cfg_.device = paramDevice(value);
    else if (paramWindowType.match(name))
        cfg_.window_type = paramWindowType(value);
    else if (paramDcDetection.match(name))
        dcDetection_ = paramDcDetection(value);
    else if (paramMinDcLength.match(name))
        dc_.min_dc_length_s = paramMinDcLength(value);
    else if (paramMaxDcIncrement.match(name))
        dc_.max_dc_increment = paramMaxDcIncrement(value);
    else if (paramMinNonDcSegmentLength.match(name))
        dc_.min_non_dc_segment_length_s = paramMinNonDcSegmentLength(value);
    else if (paramMaximalOutputSize.match(name))
        dc_.maximal_output_size = paramMaximalOutputSize(value);
    else
        return false;
    dirty_ = true;
    return true;
}

bool MfccNode::configure() {
    auto a = std::make_shared<Flow::Attributes>();
    getInputAttributes(0, *a);
    if (!configureDatatype(a, Flow::Vector<f32>::type()))
        return false;
    // attributes travel as text (src/Flow/Attributes.hh:104-113): same atof round trip as the reference nodes
    f64 sampleRate = atof(a->get("sample-rate").c_str());
    if (sampleRate <= 0.0)
        criticalError("Sample rate is not positive: %f", sampleRate);
    if (sampleRate != cfg_.sample_rate) {
        cfg_.sample_rate = sampleRate;
        dirty_           = true;
    }
    if (!ensureHandle())
        return false;
    // what the chain window -> ... -> cosine-transform leaves in the attributes: the window node adds "frame-shift"
    // (src/Signal/Window.cc:166), the FFT and filter-bank nodes rewrite "sample-rate" and the cosine transform finally
    // sets it to 1 (src/Signal/CosineTransform.cc:208); the regression / concat mergers keep both
    a->set("frame-shift", cfg_.window_shift_s);
    a->set("sample-rate", 1);
    a->set("datatype", Flow::Vector<f32>::type()->name());
    return putOutputAttributes(0, a);
}

bool MfccNode::ensureHandle() {
    if (handle_ && !dirty_)
        return true;
    rb_frontend_destroy(handle_);
    handle_ = 0;
    if (rb_frontend_create(&cfg_, &handle_) != RB_OK || rb_frontend_set_dc_detection(handle_, dcDetection_ ? &dc_ : 0) != RB_OK) {
        criticalError("rasr_b200: %s", rb_last_error());
        return false;
    }
    rb_frontend_geometry g;
    rb_frontend_get_geometry(handle_, &g);
    featDim_ = g.feat_dim;
    dirty_   = false;
    return true;
}

void MfccNode::computeSegment() {
    if (rb_frontend_finish(handle_) != RB_OK)
        criticalError("rasr_b200: %s", rb_last_error());
    nFrames_   = rb_frontend_nframes(handle_);
    nextFrame_ = 0;
    feats_.resize(nFrames_ * featDim_);
    tStart_.resize(nFrames_);
    tEnd_.resize(nFrames_);
    if (nFrames_ && rb_frontend_read(handle_, feats_.data(), tStart_.data(), tEnd_.data()) != RB_OK)
        criticalError("rasr_b200: %s", rb_last_error());
}

bool MfccNode::work(Flow::PortId p) {
    if (nextFrame_ >= nFrames_) {
        // nothing left to emit: pull the next segment
        Flow::DataPtr<Flow::Vector<f32>> in;
        if (!segmentOpen_) {
            rb_frontend_reset(handle_);
            segmentOpen_ = true;
        }
        bool any = false;
        while (getData(0, in)) {
            if (rb_frontend_push(handle_, in->data(), in->size(), in->startTime()) != RB_OK)
                criticalError("rasr_b200: %s", rb_last_error());
            any = true;
        }
        if (in == Flow::Data::ood())
            return putOod(p);
        // EOS: the segment is complete (state reset like src/Signal/Preemphasis.cc:110-118)
        segmentOpen_ = false;
        nFrames_ = nextFrame_ = 0;
        if (any)
            computeSegment();
        if (nFrames_ == 0)
            return putData(0, in.get());  // forward the sentinel
    }
    // one packet per work() call; ownership passes to putData (src/Flow/Node.hh:119-126) and the packet is
    // not referenced afterwards (Speech::Feature::take steals its storage, src/Speech/Feature.cc:19-25)
    Flow::Vector<f32>* out = new Flow::Vector<f32>(feats_.begin() + nextFrame_ * featDim_,
                                                   feats_.begin() + (nextFrame_ + 1) * featDim_);
    out->setStartTime(tStart_[nextFrame_]);
    out->setEndTime(tEnd_[nextFrame_]);
    ++nextFrame_;
    return putData(0, out);
}
