/*
 * ref_io.cc -- the REFERENCE's own readers and writers of its on-disk model formats, behind C entry points
 * (part of oracle/_ref/librasr_ref*.so; TEST INFRASTRUCTURE ONLY, contains no reference code).
 *
 * Used to pin rasr_b200/io.py: files written here by the reference's code are what the Python
 * readers are checked against (tests/test_ref_io.py, fixtures under tests/golden/ made by tests/golden/make_golden.py),
 * and files written by the Python writers are read back through the reference's readers.
 *
 *   mixture text files   Mm::Module_::writeMixtureSet / readMixtureSet        src/Mm/Module.cc:151-181,328-338
 *   accumulator files    Mm::MixtureSetEstimator (magic "MIXSET"), written after a Viterbi accumulation over the given
 *                        frames and read back through MixtureSetReader's default reader, which ESTIMATES the mixture set
 *                        from the accumulators (src/Mm/MixtureSetReader.hh:108-121, .cc:49-72,
 *                        src/Mm/AbstractMixtureSetEstimator.cc:299-338,433-478)
 *   Math::Matrix / Math::Vector files in "bin:" and "xml:" form              src/Math/Module.cc:26-40
 */
#include <Core/Application.hh>
#include <Core/Configuration.hh>
#include <Math/Matrix.hh>
#include <Math/Module.hh>
#include <Math/Vector.hh>
#include <Mm/AbstractMixtureSetEstimator.hh>
#include <Mm/AssigningFeatureScorer.hh>
#include <Mm/GaussDiagonalMaximumFeatureScorer.hh>
#include <Mm/MixtureSet.hh>
#include <Mm/Module.hh>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

struct ref_mixture_set {  // as in ref_host.cc
    uint32_t        dim, n_mixtures, n_densities, n_means, n_covariances;
    const uint32_t* mix_offsets;
    const uint32_t* mix_density;
    const double*   mix_log_weight;
    const uint32_t* dens_mean;
    const uint32_t* dens_cov;
    const float*    means;
    const float*    variances;
};
int  ref_init(const char* log_file);
void ref_config_set(const char* name, const char* value);
}

namespace {

Core::Ref<Mm::MixtureSet> build(const ref_mixture_set* m) {
    Mm::MixtureSet* ms = new Mm::MixtureSet(m->dim);
    for (uint32_t i = 0; i < m->n_means; ++i)
        ms->addMean(new Mm::Mean(std::vector<Mm::MeanType>(m->means + (size_t)i * m->dim, m->means + (size_t)(i + 1) * m->dim)));
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
    return Core::Ref<Mm::MixtureSet>(ms);
}

struct Loaded {
    Core::Ref<Mm::MixtureSet> ms;
};

}  // namespace

extern "C" {

/* text file with `precision` significant digits (the reference's default is 6) */
int ref_mm_write_text(const ref_mixture_set* m, const char* filename, int precision) {
    ref_init(0);
    Core::Ref<Mm::MixtureSet> ms = build(m);
    return Mm::Module::instance().writeMixtureSet(filename, *ms, precision) ? 0 : -1;
}

/* Viterbi accumulation of T frames (frame t belongs to mixture mix_of_frame[t]; the density is chosen by the
 * reference's diagonal-maximum scorer of the same mixture set, as in training), accumulators written to filename */
int ref_mm_write_estimator(const ref_mixture_set* m, const float* feats, long T, const uint32_t* mix_of_frame,
                           const char* filename) {
    ref_init(0);
    Core::Ref<Mm::MixtureSet> ms = build(m);
    Core::Configuration       config(Core::Application::us()->getConfiguration(), "mixture-set-estimator");
    Mm::AbstractMixtureSetEstimator* est = Mm::Module::instance().createMixtureSetEstimator(config);
    if (!est)
        return -1;
    est->setTopology(ms);
    Core::Configuration fsConfig(Core::Application::us()->getConfiguration(), "estimator-feature-scorer");
    Core::Ref<const Mm::AssigningFeatureScorer> fs(new Mm::GaussDiagonalMaximumFeatureScorer(fsConfig, ms));
    est->setAssigningFeatureScorer(fs);
    for (long t = 0; t < T; ++t) {
        Mm::Feature::VectorRef v(new Mm::Feature::Vector(Mm::FeatureVector(feats + (size_t)t * m->dim, feats + (size_t)(t + 1) * m->dim)));
        est->accumulate(mix_of_frame[t], v);
    }
    const bool ok = Mm::Module::instance().writeMixtureSetEstimator(filename, *est);
    delete est;
    return ok ? 0 : -1;
}

/* Mm::Module_::readMixtureSet: text files by extension .pms / .gz, anything else through the estimator reader */
void* ref_mm_load(const char* filename, const char* selection) {
    ref_init(0);
    Core::Configuration       config(Core::Application::us()->getConfiguration(), selection ? selection : "mixture-set");
    Core::Ref<Mm::MixtureSet> ms = Mm::Module::instance().readMixtureSet(filename, config);
    if (!ms)
        return 0;
    Loaded* l = new Loaded();
    l->ms     = ms;
    return l;
}

void ref_mm_unload(void* handle) {
    delete static_cast<Loaded*>(handle);
}

/* sizes: dim, n_mixtures, n_densities, n_means, n_covariances, n_entries */
void ref_mm_sizes(void* handle, uint32_t* sizes) {
    const Mm::MixtureSet& ms = *static_cast<Loaded*>(handle)->ms;
    uint32_t              n  = 0;
    for (Mm::MixtureIndex i = 0; i < ms.nMixtures(); ++i)
        n += ms.mixture(i)->nDensities();
    sizes[0] = ms.dimension();
    sizes[1] = ms.nMixtures();
    sizes[2] = ms.nDensities();
    sizes[3] = ms.nMeans();
    sizes[4] = ms.nCovariances();
    sizes[5] = n;
}

void ref_mm_dump(void* handle, uint32_t* mix_offsets, uint32_t* mix_density, double* mix_log_weight, uint32_t* dens_mean,
                 uint32_t* dens_cov, float* means, float* variances) {
    const Mm::MixtureSet& ms = *static_cast<Loaded*>(handle)->ms;
    const uint32_t        D  = ms.dimension();
    uint32_t              n  = 0;
    for (Mm::MixtureIndex i = 0; i < ms.nMixtures(); ++i) {
        const Mm::Mixture* mix = ms.mixture(i);
        mix_offsets[i]         = n;
        for (Mm::DensityIndex d = 0; d < mix->nDensities(); ++d, ++n) {
            mix_density[n]    = mix->densityIndex(d);
            mix_log_weight[n] = mix->logWeight(d);
        }
    }
    mix_offsets[ms.nMixtures()] = n;
    for (Mm::DensityIndex i = 0; i < ms.nDensities(); ++i) {
        dens_mean[i] = ms.density(i)->meanIndex();
        dens_cov[i]  = ms.density(i)->covarianceIndex();
    }
    for (Mm::MeanIndex i = 0; i < ms.nMeans(); ++i)
        std::memcpy(means + (size_t)i * D, &(*ms.mean(i))[0], sizeof(float) * D);
    for (Mm::CovarianceIndex i = 0; i < ms.nCovariances(); ++i)
        std::memcpy(variances + (size_t)i * D, &ms.covariance(i)->diagonal()[0], sizeof(float) * D);
}

/* Math::Matrix<f32> / Math::Vector<f32> through Math::Module's format set; filename may carry "bin:" / "xml:" */
int ref_math_write_matrix(const char* filename, int rows, int cols, const float* data) {
    ref_init(0);
    Math::Matrix<f32> m(rows, cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c)
            m[r][c] = data[(size_t)r * cols + c];
    return Math::Module::instance().formats().write(filename, m, 20) ? 0 : -1;
}

/* rows / cols receive the shape; data (may be NULL) has room for capacity floats */
int ref_math_read_matrix(const char* filename, int* rows, int* cols, float* data, long capacity) {
    ref_init(0);
    Math::Matrix<f32> m;
    if (!Math::Module::instance().formats().read(filename, m))
        return -1;
    *rows = (int)m.nRows();
    *cols = (int)m.nColumns();
    if (data && (long)m.nRows() * (long)m.nColumns() <= capacity)
        for (size_t r = 0; r < m.nRows(); ++r)
            for (size_t c = 0; c < m.nColumns(); ++c)
                data[r * m.nColumns() + c] = m[r][c];
    return 0;
}

int ref_math_write_vector(const char* filename, int n, const float* data) {
    ref_init(0);
    Math::Vector<f32> v(std::vector<f32>(data, data + n));
    return Math::Module::instance().formats().write(filename, v, 20) ? 0 : -1;
}

int ref_math_read_vector(const char* filename, int* n, float* data, long capacity) {
    ref_init(0);
    Math::Vector<f32> v;
    if (!Math::Module::instance().formats().read(filename, v))
        return -1;
    *n = (int)v.size();
    if (data && (long)v.size() <= capacity)
        std::memcpy(data, &v[0], sizeof(float) * v.size());
    return 0;
}

}  // extern "C"
