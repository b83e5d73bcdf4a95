#include "B200FeatureScorer.hh"
#include "B200NnNetwork.hh"

#include <Math/Matrix.hh>
#include <Math/Module.hh>
#include <Math/Vector.hh>
#include <Mm/Feature.hh>
#include <Mm/GaussDensity.hh>
#include <Mm/Mixture.hh>
#include <cmath>
#include <numeric>

using namespace B200;

const Core::ParameterInt   FeatureScorer::paramDevice("device", "CUDA device ordinal", 0, 0);
const Core::ParameterInt   FeatureScorer::paramBufferSize("buffer-size", "frames buffered per dense scoring launch (default: the whole segment)", Core::Type<s32>::max, 1);
const Core::ParameterBool  GmmFeatureScorer::paramContraction("fma-contraction", "reproduce a CPU build with -ffp-contract=fast (gcc default) instead of a strict one", true);
const Core::ParameterFloat GmmFeatureScorer::paramMixtureWeightScale("mixture-weight-scale", "scale of the -log mixture weights (diagonal scorers)", 1.0);
const Core::ParameterInt   GmmFeatureScorer::paramClusters("clusters", "number of density clusters to build for density preselection", 256, 1, 256);
const Core::ParameterInt   GmmFeatureScorer::paramSelectClusters("select-clusters", "number of clusters to select in density preselection", 32, 1, 256);
const Core::ParameterInt   GmmFeatureScorer::paramClusteringIterations("iterations", "number of clustering iterations", 5);
const Core::ParameterFloat GmmFeatureScorer::paramBackoffScore("backoff-score", "score used if no cluster is selected", 40000);
const Core::ParameterFloat GmmFeatureScorer::paramGaussianScale("gaussian-scale", "scale of the Gaussian exponent (diagonal scorers)", 1.0);

class FeatureScorer::ContextScorer : public Mm::FeatureScorer::ContextScorer, public DenseScoreRow {
public:
    ContextScorer(const FeatureScorer* parent, u32 segment, u32 frame)
            : parent_(parent), segment_(segment), frame_(frame) {}
    virtual Mm::EmissionIndex nEmissions() const {
        return parent_->nMixtures();
    }
    virtual Mm::Score score(Mm::EmissionIndex e) const {
        return parent_->score(segment_, frame_, e);
    }
    virtual const f32* scoreRow() const {
        return parent_->row(segment_, frame_);
    }

private:
    const FeatureScorer* parent_;
    u32                  segment_, frame_;
};

FeatureScorer::FeatureScorer(const Core::Configuration& c)
        : Core::Component(c),
          Mm::FeatureScorer(c),
          nMixtures_(0),
          dimension_(0),
          bufferSize_(paramBufferSize(c)),
          nScored_(0),
          nextFrame_(0),
          segment_(0) {}

GmmFeatureScorer::GmmFeatureScorer(const Core::Configuration& c, Core::Ref<const Mm::MixtureSet> ms, rb_gmm_mode mode)
        : Core::Component(c),
          FeatureScorer(c),
          handle_(0) {
    nMixtures_ = ms->nMixtures();
    dimension_ = ms->dimension();
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
    if (mode == RB_GMM_BATCH_PRESELECT || mode == RB_GMM_BATCH_PRESELECT_INT) {  // the int variant ignores the back-off score
        // the parameters of Mm::DensityClusteringBase under the same selection (src/Mm/BatchFeatureScorer.cc:262,
        // src/Mm/DensityClustering.cc:20-34)
        const Core::Configuration dc(c, "density-clustering");
        if (rb_gmm_configure_preselection(handle_, paramClusters(dc), paramSelectClusters(dc), paramClusteringIterations(dc),
                                          paramBackoffScore(dc)) != RB_OK)
            criticalError("rasr_b200: %s", rb_last_error());
    }
    log("b200 feature scorer: %d mixtures, %d densities, dimension %d on device %d",
        int(nMixtures_), int(ms->nDensities()), int(dimension_), int(paramDevice(c)));
}

GmmFeatureScorer::~GmmFeatureScorer() {
    rb_gmm_destroy(handle_);
}

void GmmFeatureScorer::scoreFrames(const f32* feats, u32 T, f32* scores) const {
    if (rb_gmm_score(handle_, feats, T, scores, 0) != RB_OK)
        criticalError("rasr_b200: %s", rb_last_error());
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
    features_.append(f.data(), f.data() + f.size());
}

Mm::FeatureScorer::Scorer FeatureScorer::getScorer(const Mm::FeatureVector& f) const {
    if (f.size() != dimension_)
        criticalError("feature has dimension %zu, mixture set expects %d", f.size(), int(dimension_));
    features_.append(f.data(), f.data() + f.size());
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
    scoreFrames(features_.data() + size_t(nScored_) * dimension_, T - nScored_,
                scores_.data() + size_t(nScored_) * nMixtures_);
    nScored_ = T;
}

Mm::Score FeatureScorer::score(u32 segment, u32 frame, Mm::EmissionIndex e) const {
    require_eq(segment, segment_);  // a scorer must not be used after reset() (same rule as OnnxFeatureScorer)
    require_lt(e, nMixtures_);
    if (frame >= nScored_)
        scoreBufferedFrames();
    return scores_[size_t(frame) * nMixtures_ + e];
}

const f32* FeatureScorer::row(u32 segment, u32 frame) const {
    require_eq(segment, segment_);
    if (frame >= nScored_)
        scoreBufferedFrames();
    return scores_.data() + size_t(frame) * nMixtures_;
}

// ---- Nn -------------------------------------------------------------------------------------------------------------

const Core::ParameterStringVector NnFeatureScorer::paramParameterFiles("parameter-files", "parameter matrix file of every layer, bottom to top (the layers' parameters-old)", ",");
const Core::ParameterString       NnFeatureScorer::paramHiddenActivation("hidden-activation", "activation of the hidden layers: sigmoid, rectified, tanh or linear", "sigmoid");
const Core::ParameterString       NnFeatureScorer::paramPriorFile("prior-file", "log prior vector file; empty: estimate from the mixture weights", "");
const Core::ParameterFloat        NnFeatureScorer::paramPrioriScale("priori-scale", "scaling of the logarithmized state priori probability", 1.0);
const Core::ParameterString       NnFeatureScorer::paramClassLabelFile("load-from-file", "class-labels: network-output-to-class-index mapping (Math::Vector<s32>, -1 = disregarded class)", "");
const Core::ParameterBool         NnFeatureScorer::paramBf16("bf16", "bf16 operands with f32 accumulation on the tensor cores (false: f32 arithmetic)", true);

void NnFeatureScorer::readParameterFiles(const Core::Configuration& c, NnNetwork& net) const {
    const std::vector<std::string> files = paramParameterFiles(c);
    if (files.empty())
        criticalError("b200-nn-batch-feature-scorer: no parameter-files given");
    const std::string hidden = paramHiddenActivation(c);
    int               hiddenAct;
    if (hidden == "sigmoid")
        hiddenAct = RB_ACT_SIGMOID;
    else if (hidden == "rectified" || hidden == "relu")
        hiddenAct = RB_ACT_RELU;
    else if (hidden == "tanh")
        hiddenAct = RB_ACT_TANH;
    else if (hidden == "linear")
        hiddenAct = RB_ACT_LINEAR;
    else
        criticalError("unknown hidden-activation '%s'", hidden.c_str());

    // one Math::Matrix per layer: row = output unit, column 0 = bias, the others the weights of that unit
    // (LinearLayer::setParameters, src/Nn/LinearLayer.cc:383-424); rb_nn_create takes weights[out][in] row-major
    const int nLayers = files.size();
    net.dims.assign(nLayers + 1, 0);
    net.acts.assign(nLayers, hiddenAct);
    net.weights.resize(nLayers);
    net.biases.resize(nLayers);
    std::vector<int>&              dims    = net.dims;
    std::vector<std::vector<f32>>& weights = net.weights;
    std::vector<std::vector<f32>>& biases  = net.biases;
    for (int l = 0; l < nLayers; ++l) {
        Math::Matrix<f32> parameters;
        log("reading parameter file ") << files[l] << " for layer " << l;
        if (!Math::Module::instance().formats().read(files[l], parameters))
            criticalError("failed to read parameter file '%s'", files[l].c_str());
        if (parameters.nColumns() < 2)
            criticalError("parameter file '%s' has no weight columns", files[l].c_str());
        const u32 out = parameters.nRows(), in = parameters.nColumns() - 1;
        if (l == 0)
            dims[0] = in;
        else if (u32(dims[l]) != in)
            criticalError("dimension mismatch: (parameter file vs. layer-dimension) %d vs. %d", int(in), dims[l]);
        dims[l + 1] = out;
        weights[l].resize(size_t(out) * in);
        biases[l].resize(out);
        for (u32 r = 0; r < out; ++r) {
            biases[l][r] = parameters[r][0];
            for (u32 k = 0; k < in; ++k)
                weights[l][size_t(r) * in + k] = parameters[r][k + 1];
        }
    }
}

NnFeatureScorer::NnFeatureScorer(const Core::Configuration& c, Core::Ref<const Mm::MixtureSet> ms)
        : Core::Component(c),
          FeatureScorer(c),
          handle_(0) {
    nMixtures_ = ms->nMixtures();
    // the network: the reference's own configuration keys (neural-network.links, <layer>.layer-type, parameters-old ...,
    // B200NnNetwork.hh) as Nn::BatchFeatureScorer reads them through Nn::NeuralNetwork (src/Nn/BatchFeatureScorer.cc:58-67),
    // or -- for set-ups without a layer description -- an explicit list of parameter files
    NnNetwork net;
    if (NnNetwork::configured(c)) {
        if (!net.read(*this, c))
            return;
        if (!net.topIsLinearAndSoftmax)
            criticalError("output layer must be of type 'linear+softmax'");
    }
    else
        readParameterFiles(c, net);
    const int         nLayers = net.nLayers();
    std::vector<int>& dims    = net.dims;
    net.acts[nLayers - 1]     = RB_ACT_SOFTMAX;  // only the scores of the output layer are computed
    dimension_        = dims[0];
    // Nn::ClassLabelWrapper (src/Nn/ClassLabelWrapper.cc:33-100): emission class -> network output from the file
    // the reference's wrapper saves (class-labels.load-from-file), else the identity
    std::vector<s32> classToOutput;
    const std::string labelFile = paramClassLabelFile(Core::Configuration(c, "class-labels"));
    if (!labelFile.empty()) {
        Math::Vector<s32> mapping;
        if (!Math::Module::instance().formats().read(labelFile, mapping) || mapping.size() != nMixtures_)
            criticalError("failed to read a mapping of %d classes from '%s'", int(nMixtures_), labelFile.c_str());
        classToOutput.assign(mapping.begin(), mapping.end());
    }
    else if (u32(dims[nLayers]) != nMixtures_)
        criticalError("no one-to-one correspondence between network outputs (%d) and classes (%d)!", dims[nLayers],
                      int(nMixtures_));

    // log prior per network output: from file, or relative mixture weight mass of the class mapped to the output
    // (Prior::setFromMixtureSet, src/Nn/Prior.cc:158-188)
    const u32        nOut = dims[nLayers];
    std::vector<f32> logPrior(nOut, 0.0f);
    if (!paramPriorFile(c).empty()) {
        Math::Vector<f32> priors;
        if (!Math::Module::instance().formats().read(paramPriorFile(c), priors) || priors.size() != nOut)
            criticalError("failed to read a prior of dimension %d from '%s'", int(nOut), paramPriorFile(c).c_str());
        std::copy(priors.begin(), priors.end(), logPrior.begin());
    }
    else {
        for (Mm::MixtureIndex m = 0; m < ms->nMixtures(); ++m) {
            const s32 o = classToOutput.empty() ? s32(m) : classToOutput[m];
            if (o < 0)
                continue;
            f32 mass = 0;
            for (size_t d = 0; d < ms->mixture(m)->nDensities(); ++d)
                mass += ms->mixture(m)->weight(d);
            logPrior[o] = mass;
        }
        const f32 observationWeight = std::accumulate(logPrior.begin(), logPrior.end(), 0.0);
        for (u32 o = 0; o < nOut; ++o)
            logPrior[o] = std::log(logPrior[o] / observationWeight);
    }
    handle_ = net.create(*this, logPrior, paramPrioriScale(c), paramBf16(c), paramDevice(c));
    if (!handle_)
        return;
    if (!classToOutput.empty() && rb_nn_set_class_mapping(handle_, classToOutput.size(), classToOutput.data()) != RB_OK)
        criticalError("rasr_b200: %s", rb_last_error());
    log("b200 nn feature scorer: %d layers, %d inputs, %d outputs on device %d", nLayers, dims[0], dims[nLayers],
        int(paramDevice(c)));
}

NnFeatureScorer::~NnFeatureScorer() {
    rb_nn_destroy(handle_);
}

void NnFeatureScorer::scoreFrames(const f32* feats, u32 T, f32* scores) const {
    if (rb_nn_score(handle_, feats, T, scores) != RB_OK)
        criticalError("rasr_b200: %s", rb_last_error());
}
