#include "B200Nodes.hh"

#include <Core/Application.hh>
#include <Math/Matrix.hh>
#include <Math/Module.hh>
#include <Math/Vector.hh>
#include <Mm/Module.hh>
#include <cmath>

#include "B200NnNetwork.hh"

using namespace B200;

// ---------------------------------------------------------------------------------------------------------------
bool FlatMixtureSet::assign(const Mm::MixtureSet& ms, std::string& what) {
    const u32 dim = ms.dimension();
    mixOffsets.assign(1, 0);
    mixDensity.clear();
    mixLogWeight.clear();
    for (Mm::MixtureIndex m = 0; m < ms.nMixtures(); ++m) {
        const Mm::Mixture& mixture = *ms.mixture(m);
        for (size_t d = 0; d < mixture.nDensities(); ++d) {
            mixDensity.push_back(mixture.densityIndex(d));
            mixLogWeight.push_back(mixture.logWeight(d));
        }
        mixOffsets.push_back(mixDensity.size());
    }
    densMean.resize(ms.nDensities());
    densCov.resize(ms.nDensities());
    for (Mm::DensityIndex d = 0; d < ms.nDensities(); ++d) {
        densMean[d] = ms.density(d)->meanIndex();
        densCov[d]  = ms.density(d)->covarianceIndex();
    }
    means.resize(size_t(ms.nMeans()) * dim);
    variances.resize(size_t(ms.nCovariances()) * dim);
    for (Mm::MeanIndex i = 0; i < ms.nMeans(); ++i)
        std::copy(ms.mean(i)->begin(), ms.mean(i)->end(), means.begin() + size_t(i) * dim);
    for (Mm::CovarianceIndex i = 0; i < ms.nCovariances(); ++i) {
        const Mm::DiagonalCovariance* cov = dynamic_cast<const Mm::DiagonalCovariance*>(ms.covariance(i));
        if (!cov) {
            what = "b200 feature scorers support diagonal covariances only";
            return false;
        }
        std::copy(cov->diagonal().begin(), cov->diagonal().end(), variances.begin() + size_t(i) * dim);
    }
    const rb_mixture_set v = {dim, u32(ms.nMixtures()), u32(ms.nDensities()), u32(ms.nMeans()), u32(ms.nCovariances()),
                              mixOffsets.data(), mixDensity.data(), mixLogWeight.data(), densMean.data(),
                              densCov.data(), means.data(), variances.data()};
    view = v;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
SegmentNode::SegmentNode(const Core::Configuration& c)
        : Core::Component(c), Flow::SleeveNode(c), dimIn_(0), dimOut_(0), nIn_(0), next_(0) {}

bool SegmentNode::configure() {
    auto a = std::make_shared<Flow::Attributes>();
    getInputAttributes(0, *a);
    if (!configureDatatype(a, Flow::Vector<f32>::type()))
        return false;
    a->set("datatype", Flow::Vector<f32>::type()->name());
    return putOutputAttributes(0, a);
}

bool SegmentNode::work(Flow::PortId p) {
    if (next_ >= nIn_) {
        // nothing left to emit: pull the next segment (to EOS), then one call into the engine
        in_.clear();
        times_.clear();
        nIn_ = next_ = 0;
        Flow::DataPtr<Flow::Vector<f32>> in;
        while (getData(0, in)) {
            if (times_.empty())
                dimIn_ = in->size();
            else if (in->size() != dimIn_)
                criticalError("feature of dimension %zu after features of dimension %d in one segment", in->size(), int(dimIn_));
            in_.append(in->data(), in->data() + in->size());
            times_.push_back(Flow::Timestamp(*in));
        }
        nIn_ = times_.size();
        if (nIn_ == 0)
            return putData(0, in.get());  // forward the sentinel (EOS / OOD)
        if (!processSegment())
            return false;
        require_eq(out_.size(), nIn_ * dimOut_);
    }
    // one packet per work() call; ownership passes to putData (src/Flow/Node.hh:119-126)
    Flow::Vector<f32>* out = new Flow::Vector<f32>(out_.data() + next_ * dimOut_, out_.data() + (next_ + 1) * dimOut_);
    out->setTimestamp(times_[next_]);
    ++next_;
    return putData(0, out);
}

// ---------------------------------------------------------------------------------------------------------------
// parameter names and defaults of Nn::NeuralNetworkForwardNode (src/Nn/NeuralNetworkForwardNode.cc:25-35) and
// Nn::Prior (src/Nn/Prior.cc:24-30)
const Core::ParameterString NnForwardNode::paramId("id", "Changing the id resets the caches for the recurrent connections.");
const Core::ParameterInt    NnForwardNode::paramBufferSize("buffer-size", "buffer size (accepted for compatibility: the whole segment is forwarded in one call)", 32);
const Core::ParameterBool   NnForwardNode::paramCheckValues("check-values", "check output of network for finiteness", false);
const Core::ParameterBool   NnForwardNode::paramDynamicBuffer("dynamic-buffer", "do not use fixed buffer size, but extend it until eos (always the case here)", false);
const Core::ParameterString NnForwardNode::paramPriorFile("prior-file", "", "");
const Core::ParameterFloat  NnForwardNode::paramPrioriScale("priori-scale", "scaling of the logarithmized state priori probability", 1.0);
const Core::ParameterBool   NnForwardNode::paramBf16("bf16", "bf16 operands with f32 accumulation on the tensor cores (false: f32 arithmetic)", true);
const Core::ParameterInt    NnForwardNode::paramDevice("device", "CUDA device ordinal", 0, 0);

NnForwardNode::NnForwardNode(const Core::Configuration& c)
        : Core::Component(c), SegmentNode(c), handle_(0), checkValues_(paramCheckValues(c)) {
    NnNetwork net;
    if (!net.read(*this, c))
        return;
    // the reference node removes the log prior from the bias of a linear+softmax top layer when a prior file is
    // given (NeuralNetworkForwardNode::initialize, :112-123): the forward output is softmax(W h + b - scale * logprior)
    const std::string priorFile = paramPriorFile(c);
    if (!priorFile.empty() && net.topIsLinearAndSoftmax) {
        Math::Vector<f32> prior;
        if (!Math::Module::instance().formats().read(priorFile, prior) || prior.size() != size_t(net.dims.back()))
            criticalError("failed to read a prior of dimension %d from '%s'", net.dims.back(), priorFile.c_str());
        const f32 scale = paramPrioriScale(c);
        for (size_t o = 0; o < prior.size(); ++o)
            net.biases.back()[o] -= scale * prior[o];
    }
    handle_ = net.create(*this, std::vector<f32>(), 0.0f, paramBf16(c), paramDevice(c));
    dimOut_ = net.dims.back();
    log("b200 neural network forward node: %d layers, %d inputs, %d outputs; the whole segment is one batch", net.nLayers(),
        net.dims.front(), net.dims.back());
}

NnForwardNode::~NnForwardNode() {
    rb_nn_destroy(handle_);
}

bool NnForwardNode::setParameter(const std::string& name, const std::string& value) {
    return true;  // "id" only resets recurrent state in the reference (:68-73); a feed-forward network has none
}

bool NnForwardNode::processSegment() {
    if (!handle_)
        return false;
    if (int(dimIn_) != rb_nn_n_inputs(handle_))
        criticalError("the network expects %d-dim input, the stream delivers %d", rb_nn_n_inputs(handle_), int(dimIn_));
    out_.resize(nIn_ * dimOut_);
    if (rb_nn_forward(handle_, in_.data(), long(nIn_), out_.data()) != RB_OK) {
        criticalError("rasr_b200: %s", rb_last_error());
        return false;
    }
    if (checkValues_)
        for (size_t i = 0; i < out_.size(); ++i)
            if (!std::isfinite(out_[i])) {
                error("non-finite output of neural network detected");
                break;
            }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
const Core::ParameterString PostprocessingNode::paramNormalizationType("normalization-type", "none, mean or mean-and-variance (signal-normalization type)", "none");
const Core::ParameterString PostprocessingNode::paramNormalizationLength("normalization-length", "length of the sliding window in frames, or infinite", "infinite");
const Core::ParameterString PostprocessingNode::paramNormalizationRight("normalization-right", "output point, or infinite", "infinite");
const Core::ParameterInt    PostprocessingNode::paramWindowMaxSize("window-max-size", "sequence concatenation: maximum length of sliding window; 0: none", 0, 0);
const Core::ParameterInt    PostprocessingNode::paramWindowRight("window-right", "sequence concatenation: position of zero time index from right", 0, 0);
const Core::ParameterString PostprocessingNode::paramMatrixFile("matrix-file", "matrix multiplication: name of matrix file to load; empty: none", "");
const Core::ParameterBool   PostprocessingNode::paramContraction("fma-contraction", "reproduce a CPU build with -ffp-contract=fast (gcc default) instead of a strict one", true);
const Core::ParameterInt    PostprocessingNode::paramDevice("device", "CUDA device ordinal", 0, 0);

PostprocessingNode::PostprocessingNode(const Core::Configuration& c)
        : Core::Component(c),
          SegmentNode(c),
          normType_(paramNormalizationType(c)),
          normLength_(paramNormalizationLength(c)),
          normRight_(paramNormalizationRight(c)),
          matrixFile_(paramMatrixFile(c)),
          handle_(0),
          handleDim_(0) {
    std::memset(&cfg_, 0, sizeof(cfg_));
    cfg_.splice_length = paramWindowMaxSize(c);
    cfg_.splice_right  = paramWindowRight(c);
    cfg_.contraction   = paramContraction(c) ? 1 : 0;
    cfg_.device        = paramDevice(c);
}

PostprocessingNode::~PostprocessingNode() {
    rb_postproc_destroy(handle_);
}

bool PostprocessingNode::setParameter(const std::string& name, const std::string& value) {
    if (paramNormalizationType.match(name))
        normType_ = value;
    else if (paramNormalizationLength.match(name))
        normLength_ = value;
    else if (paramNormalizationRight.match(name))
        normRight_ = value;
    else if (paramWindowMaxSize.match(name))
        cfg_.splice_length = paramWindowMaxSize(value);
    else if (paramWindowRight.match(name))
        cfg_.splice_right = paramWindowRight(value);
    else if (paramMatrixFile.match(name))
        matrixFile_ = value;
    else if (paramContraction.match(name))
        cfg_.contraction = paramContraction(value) ? 1 : 0;
    else
        return false;
    rb_postproc_destroy(handle_);
    handle_ = 0;
    return true;
}

bool PostprocessingNode::processSegment() {
    if (!handle_ || handleDim_ != dimIn_) {
        rb_postproc_destroy(handle_);
        handle_ = 0;
        if (normType_ == "none" || normType_.empty())
            cfg_.norm_type = 0;
        else if (normType_ == "mean")
            cfg_.norm_type = 1;
        else if (normType_ == "mean-and-variance")
            cfg_.norm_type = 2;
        else
            criticalError("normalization-type '%s': none, mean or mean-and-variance are supported", normType_.c_str());
        cfg_.norm_length = normLength_ == "infinite" ? -1 : atol(normLength_.c_str());
        cfg_.norm_right  = normRight_ == "infinite" ? -1 : atol(normRight_.c_str());
        cfg_.matrix      = 0;
        if (!matrixFile_.empty()) {
            Math::Matrix<f32> m;
            if (!Math::Module::instance().formats().read(matrixFile_, m))
                criticalError("Failed to read matrix from file '%s'.", matrixFile_.c_str());
            cfg_.matrix_rows = m.nRows();
            cfg_.matrix_cols = m.nColumns();
            matrix_.resize(size_t(m.nRows()) * m.nColumns());
            for (u32 r = 0; r < m.nRows(); ++r)
                for (u32 k = 0; k < m.nColumns(); ++k)
                    matrix_[size_t(r) * m.nColumns() + k] = m[r][k];
            cfg_.matrix = matrix_.data();
        }
        if (rb_postproc_create(&cfg_, int(dimIn_), &handle_) != RB_OK) {
            criticalError("rasr_b200: %s", rb_last_error());
            return false;
        }
        handleDim_ = dimIn_;
        dimOut_    = rb_postproc_dim_out(handle_);
    }
    const int64_t offsets[2] = {0, int64_t(nIn_)};
    out_.resize(nIn_ * dimOut_);
    if (rb_postproc_process(handle_, in_.data(), offsets, 1, out_.data()) != RB_OK) {
        criticalError("rasr_b200: %s", rb_last_error());
        return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
AudioScorerNode::AudioScorerNode(const Core::Configuration& c)
        : Core::Component(c), MfccNode(c), gmm_(0), nMixtures_(0), segmentStart_(0), haveStart_(false), nScoreFrames_(0), nextScore_(0) {}

AudioScorerNode::~AudioScorerNode() {
    rb_gmm_destroy(gmm_);
}

// mixture set and scorer type exactly as Speech::FeatureScorerNode finds them (src/Speech/FeatureScorerNode.cc:27-29):
// selection "mixture-set" (file, Mm::Module::readMixtureSet) and selection "feature-scorer" (feature-scorer-type)
bool AudioScorerNode::ensureScorer() {
    if (gmm_)
        return true;
    Core::Ref<Mm::MixtureSet> ms = Mm::Module::instance().readMixtureSet(select("mixture-set"));
    if (!ms) {
        criticalError("b200-audio-feature-scorer: cannot read the mixture set");
        return false;
    }
    static const Core::Choice          choiceMode("batch-diagonal-maximum-float", RB_GMM_BATCH_FLOAT, "b200-batch-float", RB_GMM_BATCH_FLOAT,
                                                  "diagonal-maximum", RB_GMM_DIAG_MAX, "b200-diagonal-maximum", RB_GMM_DIAG_MAX,
                                                  "b200-diagonal-sum", RB_GMM_DIAG_SUM, "b200-batch-tensor", RB_GMM_BATCH_TENSOR,
                                                  "batch-diagonal-maximum-int", RB_GMM_BATCH_INT, "b200-batch-int", RB_GMM_BATCH_INT,
                                                  "preselection-batch-float", RB_GMM_BATCH_PRESELECT,
                                                  "b200-preselection-batch-float", RB_GMM_BATCH_PRESELECT,
                                                  "preselection-batch-int", RB_GMM_BATCH_PRESELECT_INT,
                                                  "b200-preselection-batch-int", RB_GMM_BATCH_PRESELECT_INT, Core::Choice::endMark());
    static const Core::ParameterChoice paramMode("feature-scorer-type", &choiceMode, "type of feature scorer", RB_GMM_DIAG_MAX);
    static const Core::ParameterBool   paramContraction("fma-contraction", "reproduce a CPU build with -ffp-contract=fast (gcc default) instead of a strict one", true);
    const Core::Configuration fsc = select("feature-scorer");
    FlatMixtureSet            flat;
    std::string               what;
    if (!flat.assign(*ms, what)) {
        criticalError("%s", what.c_str());
        return false;
    }
    if (rb_gmm_create(&flat.view, paramMode(fsc), 1.0f, 1.0f, paramContraction(fsc) ? 1 : 0, cfg_.device, &gmm_) != RB_OK) {
        criticalError("rasr_b200: %s", rb_last_error());
        return false;
    }
    nMixtures_ = ms->nMixtures();
    return true;
}

bool AudioScorerNode::work(Flow::PortId p) {
    if (nextScore_ >= nScoreFrames_) {
        samples_.clear();
        haveStart_    = false;
        nScoreFrames_ = nextScore_ = 0;
        Flow::DataPtr<Flow::Vector<f32>> in;
        while (getData(0, in)) {
            if (!haveStart_) {
                segmentStart_ = in->startTime();
                haveStart_    = true;
            }
            samples_.append(in->data(), in->data() + in->size());
        }
        if (!haveStart_ || samples_.size() == 0)
            return putData(0, in.get());  // forward the sentinel
        if (!ensureHandle() || !ensureScorer())
            return false;
        const long T = rb_frontend_timestamps(handle_, long(samples_.size()), segmentStart_, 0, 0);
        frameStart_.resize(T);
        frameEnd_.resize(T);
        rb_frontend_timestamps(handle_, long(samples_.size()), segmentStart_, frameStart_.data(), frameEnd_.data());
        scores_.resize(size_t(T) * nMixtures_);
        const int64_t offsets[2] = {0, int64_t(samples_.size())};
        // MFCC chain and scorer back to back on the device; only the samples go in and only the scores come back
        if (rb_pipeline_score(handle_, gmm_, samples_.data(), offsets, 1, scores_.data(), 0) != RB_OK) {
            criticalError("rasr_b200: %s", rb_last_error());
            return false;
        }
        nScoreFrames_ = T;
        if (T == 0)
            return putData(0, in.get());
    }
    // a FeatureScorer returns scores in -log space, this node (like Speech::FeatureScorerNode::putData, :95-111)
    // emits +log
    Flow::Vector<f32>* out = new Flow::Vector<f32>(nMixtures_);
    const f32*         row = scores_.data() + nextScore_ * nMixtures_;
    for (u32 e = 0; e < nMixtures_; ++e)
        (*out)[e] = -row[e];
    out->setStartTime(frameStart_[nextScore_]);
    out->setEndTime(frameEnd_[nextScore_]);
    ++nextScore_;
    return putData(0, out);
}
