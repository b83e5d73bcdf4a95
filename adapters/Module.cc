#include "Module.hh"

#include <Flow/Registry.hh>
#include <Mm/FeatureScorerFactory.hh>
#include <Mm/Module.hh>

#include "B200FeatureScorer.hh"
#include "B200MfccNode.hh"
#include "B200Nodes.hh"

namespace B200 {

Module_::Module_() {
    // feature-scorer-type ids in use: 0..19 Mm (src/Mm/Module.hh:47-69), 0x300.. Nn, 0x400.. Onnx
    Mm::FeatureScorerFactory* f = Mm::Module::instance().featureScorerFactory();
    f->registerFeatureScorer<FeatureScorerOf<RB_GMM_BATCH_FLOAT>, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x500 + 0, "b200-batch-float");
    f->registerFeatureScorer<FeatureScorerOf<RB_GMM_DIAG_MAX>, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x500 + 1, "b200-diagonal-maximum");
    f->registerFeatureScorer<FeatureScorerOf<RB_GMM_DIAG_SUM>, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x500 + 2, "b200-diagonal-sum");
    f->registerFeatureScorer<FeatureScorerOf<RB_GMM_BATCH_TENSOR>, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x500 + 3, "b200-batch-tensor");
    f->registerFeatureScorer<FeatureScorerOf<RB_GMM_BATCH_INT>, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x500 + 4, "b200-batch-int");
    f->registerFeatureScorer<FeatureScorerOf<RB_GMM_BATCH_PRESELECT>, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x500 + 6, "b200-preselection-batch-float");
    f->registerFeatureScorer<FeatureScorerOf<RB_GMM_BATCH_PRESELECT_INT>, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x500 + 7, "b200-preselection-batch-int");
    f->registerFeatureScorer<FeatureScorerOf<RB_GMM_SIMD_DIAG_MAX>, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x500 + 8, "b200-SIMD-diagonal-maximum");
    f->registerFeatureScorer<NnFeatureScorer, Mm::MixtureSet, Mm::AbstractMixtureSetLoader>(0x500 + 5, "b200-nn-batch-feature-scorer");
    Flow::Registry::instance().registerFilter<MfccNode>();
    Flow::Registry::instance().registerFilter<NnForwardNode>();
    Flow::Registry::instance().registerFilter<PostprocessingNode>();
    Flow::Registry::instance().registerFilter<AudioScorerNode>();
}

}  // namespace B200
