/*
 * ref_nn.cc -- registers the reference's legacy Nn components in the oracle build exactly as src/Nn/Module.cc:100-115
 * does (same filter name, same feature-scorer ids and names).  Nn::Module itself cannot be compiled here: its header
 * pulls in the ONNX runtime (src/Onnx/Value.hh).  TEST INFRASTRUCTURE ONLY; contains no reference code.
 */
#include <Flow/Registry.hh>
#include <Mm/FeatureScorerFactory.hh>
#include <Mm/Module.hh>
#include <Nn/BatchFeatureScorer.hh>
#include <Nn/FeatureScorer.hh>
#include <Nn/NeuralNetworkForwardNode.hh>
#include <Nn/Statistics.hh>

#include <cstdio>
#include <cstdlib>

// src/Nn/Statistics.cc includes Nn/Module.hh (ONNX runtime headers) and is therefore not part of the oracle build; the one
// member src/Nn/Prior.cc refers to (Prior::setFromClassCounts, a training-side path nothing here calls) resolves to this
namespace Nn {
template<>
u32 Statistics<f32>::classCount(u32) const {
    std::fprintf(stderr, "oracle/_ref: Nn::Statistics is not part of this build\n");
    std::abort();
}
template<>
u32 Statistics<f64>::classCount(u32) const {
    std::fprintf(stderr, "oracle/_ref: Nn::Statistics is not part of this build\n");
    std::abort();
}
}  // namespace Nn

extern "C" void ref_register_nn() {
    static bool done = false;
    if (done)
        return;
    done = true;
    Flow::Registry::instance().registerFilter<Nn::NeuralNetworkForwardNode>();
    Mm::FeatureScorerFactory* f = Mm::Module::instance().featureScorerFactory();
    f->registerFeatureScorer<Nn::OnDemandFeatureScorer, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x300, "nn-on-demand-hybrid");
    f->registerFeatureScorer<Nn::FullFeatureScorer, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x301, "nn-full-hybrid");
    f->registerFeatureScorer<Nn::BatchFeatureScorer, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x303, "nn-batch-feature-scorer");
}
