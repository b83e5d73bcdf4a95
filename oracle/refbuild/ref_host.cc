/*
 * ref_host.cc -- C entry points around the REFERENCE's own object code (oracle/_ref/librasr_ref*.so).
 *
 * TEST INFRASTRUCTURE ONLY.  This file contains no reference code.  It is compiled together with translation
 * units of /root/reference/src/{Core,Flow,Math,Signal,Mm,...} -- from where they lie, recipe in
 * oracle/refbuild/Makefile -- and drives them the way RASR's tools do:
 *
 *   - ref_flow_*   builds a Flow::Network with the reference's own NetworkParser from a .flow file (the reference's
 *                  mfcc.flow / derivationWithRegression.flow are included by the top-level files under
 *                  oracle/refbuild/flows/), feeds it samples through a source node defined here (the audio file
 *                  readers of src/Audio need libsndfile & co.) and pulls packets from its output port, as
 *                  Speech::DataSource::getData does (src/Speech/DataSource.cc).
 *   - ref_mm_*     builds an Mm::MixtureSet through its public interface, creates a feature scorer with
 *                  Mm::Module's own factory (src/Mm/Module.cc:83-105) and replays the recognizer's buffered call
 *                  protocol (src/Speech/Recognizer.cc:271-281,197-205) to obtain every emission's score per frame.
 *
 * The tests use it to pin the restatement in oracle/*.cc (and, through fixtures generated from it, the CUDA path)
 * to the reference's real behaviour; bench.py --impl reference times it as the CPU baseline ("kind": "reference").
 */
#include <Core/Application.hh>
#include <Core/Configuration.hh>
#include <Flow/Attributes.hh>
#include <Flow/Network.hh>
#include <Flow/Node.hh>
#include <Flow/Registry.hh>
#include <Flow/Vector.hh>
#include <Mm/AssigningFeatureScorer.hh>
#include <Mm/FeatureScorer.hh>
#include <Mm/GaussDiagonalMaximumFeatureScorer.hh>
#include <Mm/MixtureSet.hh>
#include <Mm/Module.hh>
#include <Signal/Module.hh>
#include <Flow/Module.hh>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------------------
// the one Core::Application instance every Core::Component reports to (logging goes to /dev/null by default)
class RefApplication : public Core::Application {
public:
    RefApplication() {
        setTitle("ref-host");
        setDefaultLoadConfigurationFile(false);
        setDefaultOutputXmlHeader(false);
    }
    virtual int main(const std::vector<std::string>&) {
        return 0;
    }
    Core::Configuration& configuration() {
        return config;
    }
    void startLogging() {
        openLogging();
    }
};

RefApplication* app       = 0;
std::string     lastError = "";

// ---------------------------------------------------------------------------------------------------------
// source node of the flow networks: hands out the samples set by ref_flow_run in packets of `block-size`
// samples with contiguous time stamps, then EOS -- what an audio input node delivers after
// generic-vector-s16-demultiplex + generic-convert-vector-s16-to-vector-f32 (samples.flow:13-18)
class SampleSourceNode : public Flow::SourceNode {
public:
    static std::string filterName() {
        return "ref-sample-source";
    }
    SampleSourceNode(const Core::Configuration& c)
            : Core::Component(c), Flow::SourceNode(c), blockSize_(4096) {
        current() = this;
    }
    static SampleSourceNode*& current() {
        static SampleSourceNode* n = 0;
        return n;
    }
    virtual bool setParameter(const std::string& name, const std::string& value) {
        if (name == "block-size")
            blockSize_ = atoi(value.c_str());
        else
            return false;
        return true;
    }
    virtual bool configure() {
        std::shared_ptr<Flow::Attributes> a(new Flow::Attributes());
        a->set("datatype", Flow::Vector<f32>::type()->name());
        a->set("sample-rate", sampleRate_);
        a->set("track-count", 1);
        return putOutputAttributes(0, a);
    }
    virtual bool work(Flow::PortId out) {
        if (pos_ >= n_)
            return putEos(out);
        const long            m = std::min<long>(blockSize_, n_ - pos_);
        Flow::Vector<f32>*    v = new Flow::Vector<f32>(samples_ + pos_, samples_ + pos_ + m);
        v->setStartTime(startTime_ + f64(pos_) / sampleRate_);
        v->setEndTime(startTime_ + f64(pos_ + m) / sampleRate_);
        pos_ += m;
        return putData(out, v);
    }
    void set(const float* samples, long n, double sampleRate, double startTime) {
        samples_    = samples;
        n_          = n;
        pos_        = 0;
        sampleRate_ = sampleRate;
        startTime_  = startTime;
    }

private:
    int          blockSize_;
    const float* samples_    = 0;
    long         n_          = 0, pos_ = 0;
    double       sampleRate_ = 16000, startTime_ = 0;
};

struct FlowHandle {
    Core::Configuration config;
    Flow::Network*      net = 0;
    SampleSourceNode*   src = 0;
    ~FlowHandle() {
        delete net;
    }
};

struct MmHandle {
    Core::Configuration                 config;
    Core::Ref<Mm::MixtureSet>           ms;
    Core::Ref<Mm::FeatureScorer>        fs;
    const Mm::AssigningFeatureScorer*   assigning = 0;
};

}  // namespace

extern "C" {

struct ref_mixture_set {  // same layout as rb_mixture_set (include/rasr_b200.h) and oracle.h's orc_mixture_set
    uint32_t        dim, n_mixtures, n_densities, n_means, n_covariances;
    const uint32_t* mix_offsets;
    const uint32_t* mix_density;
    const double*   mix_log_weight;
    const uint32_t* dens_mean;
    const uint32_t* dens_cov;
    const float*    means;
    const float*    variances;
};

const char* ref_last_error() {
    return lastError.c_str();
}

/* creates the application object and registers the flow filters; log_file = NULL: /dev/null */
int ref_init(const char* log_file);
void ref_register_nn();

int ref_init(const char* log_file) {
    if (app)
        return 0;
    app = new RefApplication();
    Core::Configuration& c = app->configuration();
    const std::string    log = log_file ? log_file : "/dev/null";
    c.set("*.log.channel", log);
    c.set("*.warning.channel", log);
    c.set("*.error.channel", log_file ? log : "stderr");
    c.set("*.statistics.channel", log);
    c.set("*.dot.channel", "nil");
    c.set("*.encoding", "UTF-8");
    c.set("*.on-error", "ignore");
    app->startLogging();
    // the reference's own registrations: generic-* filters (src/Flow/Module.cc), signal-* filters
    // (src/Signal/Module.cc:83-120), feature scorers (src/Mm/Module.cc:83-105)
    INIT_MODULE(Flow);
    INIT_MODULE(Signal);
    INIT_MODULE(Mm);
    ref_register_nn();  // neural-network-forward, nn-batch-feature-scorer ... (ref_nn.cc, as src/Nn/Module.cc:100-115)
    Flow::Registry::instance().registerFilter<SampleSourceNode>();
    return 0;
}

/* a resource of the global configuration, e.g. ("*.density-clustering.clusters", "64") */
void ref_config_set(const char* name, const char* value) {
    ref_init(0);
    app->configuration().set(name, value);
}

/* --------------------------------------------------------------------------------------------------------- */
/* Flow networks */

void* ref_flow_create(const char* flow_file, const char* selection) {
    ref_init(0);
    FlowHandle* h = new FlowHandle();
    h->config     = Core::Configuration(app->configuration(), selection ? selection : "flow");
    h->net        = new Flow::Network(h->config, false);
    SampleSourceNode::current() = 0;
    h->net->buildFromFile(flow_file);
    if (h->net->hasFatalErrors()) {
        lastError = std::string("cannot build the network from ") + flow_file;
        delete h;
        return 0;
    }
    h->src = SampleSourceNode::current();
    return h;
}

int ref_flow_set_parameter(void* handle, const char* name, const char* value) {
    FlowHandle* h = static_cast<FlowHandle*>(handle);
    return h->net->setParameter(name, value) ? 0 : -1;
}

void ref_flow_destroy(void* handle) {
    delete static_cast<FlowHandle*>(handle);
}

/* runs one segment through the network and pulls Flow::Vector<f32> packets from output port `port` until EOS.
 * feats [capacity * dim] / t_start / t_end [capacity] may be NULL (count only); *dim receives the packet size
 * (ragged packets: the maximum; rows are zero padded).  Returns the number of packets, < 0 on error. */
long ref_flow_run(void* handle, const char* port, const float* samples, long n, double sample_rate,
                  double start_time, float* feats, long capacity, int row_stride, double* t_start, double* t_end,
                  int* dim, int* sizes) {
    FlowHandle* h = static_cast<FlowHandle*>(handle);
    if (!h->src) {
        lastError = "the network has no ref-sample-source node";
        return -1;
    }
    h->src->set(samples, n, sample_rate, start_time);
    h->net->reset();
    h->net->configureAll();
    const Flow::PortId out = h->net->getOutput(port);
    if (out == Flow::IllegalPortId) {
        lastError = std::string("no output port ") + port;
        return -1;
    }
    h->net->activateOutput(out);
    long                              count = 0;
    int                               width = 0;
    Flow::DataPtr<Flow::Vector<f32>>  d;
    while (h->net->getData(out, d)) {
        width = std::max<int>(width, (int)d->size());
        if (count < capacity) {
            if (feats) {
                const int m = std::min<int>((int)d->size(), row_stride);
                std::memset(feats + (size_t)count * row_stride, 0, sizeof(float) * row_stride);
                std::memcpy(feats + (size_t)count * row_stride, d->data(), sizeof(float) * m);
            }
            if (t_start)
                t_start[count] = d->startTime();
            if (t_end)
                t_end[count] = d->endTime();
            if (sizes)
                sizes[count] = (int)d->size();
        }
        ++count;
    }
    if (dim)
        *dim = width;
    return count;
}

/* attribute of an output port after configuration (text, as it travels in the reference) */
int ref_flow_get_attribute(void* handle, const char* port, const char* name, char* value, int capacity) {
    FlowHandle*        h   = static_cast<FlowHandle*>(handle);
    const Flow::PortId out = h->net->getOutput(port);
    if (out == Flow::IllegalPortId)
        return -1;
    const std::string v = h->net->getAttribute(out, name);
    std::snprintf(value, capacity, "%s", v.c_str());
    return (int)v.size();
}

/* --------------------------------------------------------------------------------------------------------- */
/* Mm feature scorers */

void* ref_mm_create(const ref_mixture_set* m, const char* scorer_type, const char* selection) {
    ref_init(0);
    MmHandle* h = new MmHandle();
    h->config   = Core::Configuration(app->configuration(), selection ? selection : "feature-scorer");
    Mm::MixtureSet* ms = new Mm::MixtureSet(m->dim);
    for (uint32_t i = 0; i < m->n_means; ++i)
        ms->addMean(new Mm::Mean(std::vector<Mm::MeanType>(m->means + (size_t)i * m->dim,
                                                            m->means + (size_t)(i + 1) * m->dim)));
    for (uint32_t i = 0; i < m->n_covariances; ++i)
        ms->addCovariance(new Mm::DiagonalCovariance(std::vector<Mm::VarianceType>(
                m->variances + (size_t)i * m->dim, m->variances + (size_t)(i + 1) * m->dim)));
    for (uint32_t i = 0; i < m->n_densities; ++i)
        ms->addDensity(new Mm::GaussDensity(m->dens_mean[i], m->dens_cov[i]));
    for (uint32_t i = 0; i < m->n_mixtures; ++i) {
        Mm::Mixture* mix = new Mm::Mixture();
        for (uint32_t e = m->mix_offsets[i]; e < m->mix_offsets[i + 1]; ++e)
            mix->addLogDensity(m->mix_density[e], m->mix_log_weight[e]);
        ms->addMixture(mix);
    }
    h->ms = Core::Ref<Mm::MixtureSet>(ms);
    h->config.set(h->config.getSelection() + ".feature-scorer-type", scorer_type);
    if (std::string(scorer_type) == "diagonal-sum") {
        // defined in the reference but not registered in the factory (src/Mm/Module.cc:83-105)
        h->fs = Core::Ref<Mm::FeatureScorer>(new Mm::GaussDiagonalSumFeatureScorer(h->config, h->ms));
    }
    else
        h->fs = Mm::Module::instance().createFeatureScorer(h->config, h->ms);
    if (!h->fs) {
        lastError = std::string("cannot create feature scorer ") + scorer_type;
        delete h;
        return 0;
    }
    h->assigning = dynamic_cast<const Mm::AssigningFeatureScorer*>(h->fs.get());
    return h;
}

void ref_mm_destroy(void* handle) {
    delete static_cast<MmHandle*>(handle);
}

/* the scorer object itself (an Mm::FeatureScorer*), for hosts that feed it to other reference code (ref_search.cc) */
void* ref_mm_feature_scorer(void* handle) {
    return static_cast<MmHandle*>(handle)->fs.get();
}

int ref_mm_n_mixtures(void* handle) {
    return static_cast<MmHandle*>(handle)->fs->nMixtures();
}

/* all emission scores of T frames, scores [T * nMixtures]; best_density (optional, assigning scorers only).
 * The call sequence is the recognizer's (src/Speech/Recognizer.cc:271-281 processFeature, :197-205 finish):
 * buffered scorers get addFeature() until bufferFilled(), then one getScorer() per further frame, flush() at the
 * end; the scorer handed out always belongs to the oldest buffered frame. */
int ref_mm_score(void* handle, const float* feats, long T, float* scores, uint32_t* best_density) {
    MmHandle*                 h  = static_cast<MmHandle*>(handle);
    const Mm::FeatureScorer&  fs = *h->fs;
    const int                 D  = h->ms->dimension();
    const Mm::EmissionIndex   M  = fs.nMixtures();
    long                      t_out = 0;
    auto emit = [&](Mm::FeatureScorer::Scorer s) {
        float* row = scores + (size_t)t_out * M;
        for (Mm::EmissionIndex e = 0; e < M; ++e)
            row[e] = s->score(e);
        ++t_out;
    };
    if (best_density && h->assigning) {
        // assigning scorers (diagonal-maximum, diagonal-sum) are not buffered: one scorer per frame
        for (long t = 0; t < T; ++t) {
            Mm::AssigningFeatureScorer::AssigningScorer a = h->assigning->getAssigningScorer(
                    Mm::FeatureVector(feats + (size_t)t * D, feats + (size_t)(t + 1) * D));
            for (Mm::EmissionIndex e = 0; e < M; ++e) {
                scores[(size_t)t * M + e]       = a->score(e);
                best_density[(size_t)t * M + e] = a->bestDensity(e);
            }
        }
        return 0;
    }
    fs.reset();
    for (long t = 0; t < T; ++t) {
        Core::Ref<const Mm::Feature> f(new Mm::Feature(Mm::FeatureVector(feats + (size_t)t * D, feats + (size_t)(t + 1) * D)));
        if (fs.isBuffered() && !fs.bufferFilled())
            fs.addFeature(f);
        else
            emit(fs.getScorer(f));
    }
    while (fs.isBuffered() && !fs.bufferEmpty())
        emit(fs.flush());
    fs.finalize();
    return t_out == T ? 0 : -1;
}

}  // extern "C"
