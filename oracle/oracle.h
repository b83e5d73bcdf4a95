/*
 * oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a CPU restatement of the reference's
 * (rwth-i6/rasr) algorithms for the acoustic front-end and the emission scorers.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product (rasr_b200/) never links or calls anything in oracle/.
 *
 * Every function names the reference file:line it follows (paths relative to the
 * reference checkout).  Parity status: PINNED to the reference's own object code.  The reference has
 * no unit tests for this path (only the Nn layer vectors of src/Test/Nn_*.cc, which are used too), so
 * 129 of its translation units (Core, Flow, Math, Mc, Mm, Nn, Signal) are compiled from where they lie
 * into oracle/_ref/librasr_ref{,_native}.so (oracle/refbuild/Makefile) and driven the way RASR's tools
 * drive them (ref_host.cc, ref_nn.cc, ref_io.cc):
 *   - front-end (pre-emphasis, framing / flush / time stamps, window, FFT, amplitude, filter bank, log,
 *     DCT, delay / regression / concat, DC detection), normalisation / splice / matrix nodes: bit for bit
 *     against Flow networks the reference's NetworkParser builds from its own .flow files;
 *   - all Mm scorers (batch-float, int, unrolled int, both preselection scorers, diagonal max / sum):
 *     bit for bit against scorers made by Mm::Module's factory, read through the recognizer's protocol;
 *   - Nn batch scorer and neural-network-forward node: against the reference's own, same configuration;
 *   (tests/test_ref_parity.py; fixtures made by that object code: tests/golden/ref_*.npz, ref_io/)
 *   - Search::LinearSearch (continuous and single-word recognition): bit for bit against the reference's own
 *     LinearSearch, compiled with the Bliss / Am / Lm / Fsa units it needs into oracle/_ref/librasr_ref_search.so
 *     and fed from a lexicon file and the reference's configuration (ref_search.cc, tests/test_ref_search.py,
 *     fixture tests/golden/ref_search.npz).
 */
#ifndef RASR_ORACLE_H
#define RASR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- front-end */

typedef struct {
    double sample_rate;       /* "sample-rate" attribute of the samples stream, e.g. 16000 */
    double window_length_s;   /* signal-window length=  (mfcc.flow:12-13), 0.025 */
    double window_shift_s;    /* signal-window shift=   , 0.01 */
    double fft_max_input_s;   /* signal-real-fast-fourier-transform maximum-input-size= , 0.025 */
    double filter_width;      /* signal-filterbank filter-width= (mel), 268.258 */
    float  preemphasis_alpha; /* signal-preemphasis alpha= , 1.0 */
    int    n_cepstra;         /* signal-cosine-transform nr-outputs= */
    int    derivatives;       /* 0: static cepstra only; 1: static || delta || delta-delta */
    int    use_fma;           /* 1: floating-point contraction as gcc -O2 -march=native does for the
                                 reference's default build; 0: every operation rounded separately */
    int    window_type;       /* 0 hamming (default), 1 rectangular, 2 hanning, 3 periodic-hanning, 4 bartlett,
                                 5 blackman (src/Signal/WindowFunction.cc:25-33,62-132) */
} orc_frontend_cfg;

typedef struct {
    int win_length;  /* samples */
    int win_shift;   /* samples */
    int fft_length;  /* points */
    int n_bins;      /* fft_length/2+1 */
    int n_filters;
    int n_weights;   /* total non-zero filterbank taps */
    int feat_dim;    /* n_cepstra * (derivatives ? 3 : 1) */
} orc_frontend_geometry;

int orc_frontend_get_geometry(const orc_frontend_cfg* cfg, orc_frontend_geometry* g);

/* number of frames the window node emits for n samples in one segment */
long orc_frontend_nframes(const orc_frontend_cfg* cfg, long n_samples);

/* Tables: window[win_length]; fb_start/fb_end[n_filters]; fb_weights dense [n_filters * n_bins]
 * (zero outside [start,end)); dct [n_cepstra * n_filters].  Any pointer may be NULL. */
int orc_frontend_tables(const orc_frontend_cfg* cfg, float* window, int* fb_start, int* fb_end,
                        float* fb_weights, float* dct);

/* Whole pipeline on one segment, fed to the network in packets of `chunk` samples (<=0: one packet).
 * feats [T * feat_dim]; t_start/t_end [T] are the timestamps of the final packets.
 * Optional per-stage dumps (NULL to skip): spectrum [T*(fft_length+2)], amplitude [T*n_bins],
 * fbank [T*n_filters] (before log), cepstra [T*n_cepstra].  Returns T or <0. */
long orc_mfcc(const orc_frontend_cfg* cfg, const float* samples, long n_samples, long chunk,
              float* feats, double* t_start, double* t_end, float* spectrum, float* amplitude,
              float* fbank, float* cepstra);

/* signal-dc-detection (src/Signal/DcDetection.{hh,cc}) between the audio source and the MFCC chain, as wired in
 * src/Tools/FeatureExtraction/share/samples.flow:34-37.  feats/t_start/t_end have room for `capacity` frames (-3 if
 * that is too small); run_* receive the kept sample runs [begin, end) and their start times. */
typedef struct orc_dc_cfg {
    double min_dc_length_s;
    float  max_dc_increment;
    double min_non_dc_segment_length_s;
    int    maximal_output_size;
} orc_dc_cfg;
long orc_mfcc_dc(const orc_frontend_cfg* cfg, const orc_dc_cfg* dc, const float* samples, long n_samples, long chunk,
                 float* feats, double* t_start, double* t_end, long capacity, long* run_begin, long* run_end,
                 double* run_start, long run_capacity, long* n_runs);

/* in-place real FFT of `n` (power of two) f32 values in the packed layout of
 * Math::FastFourierTransform::transformReal (restatement). */
void orc_fft_real_packed(float* v, int n);

/* ---------------------------------------------------------------- GMM scorers */

typedef struct {
    uint32_t        dim;
    uint32_t        n_mixtures;
    uint32_t        n_densities;
    uint32_t        n_means;
    uint32_t        n_covariances;
    const uint32_t* mix_offsets;    /* [n_mixtures+1] into mix_density / mix_log_weight */
    const uint32_t* mix_density;    /* density index of each mixture entry */
    const double*   mix_log_weight; /* natural-log weight of each mixture entry (Mm::Weight = f64) */
    const uint32_t* dens_mean;      /* [n_densities] mean index */
    const uint32_t* dens_cov;       /* [n_densities] covariance index */
    const float*    means;          /* [n_means * dim] */
    const float*    variances;      /* [n_covariances * dim] diagonal */
} orc_mixture_set;

/* Mm::BatchFloatFeatureScorer, all mixtures x all frames. scores [T * n_mixtures]. */
int orc_gmm_batch_float(const orc_mixture_set* ms, const float* feats, long T, float* scores, int use_fma);
/* Mm::GaussDiagonalMaximumFeatureScorer. best [T * n_mixtures] density-in-mixture index (may be NULL). */
int orc_gmm_diag_max(const orc_mixture_set* ms, float mixture_weight_scale, float gaussian_scale,
                     const float* feats, long T, float* scores, uint32_t* best, int use_fma);
/* Mm::GaussDiagonalSumFeatureScorer (log-sum-exp). */
int orc_gmm_diag_sum(const orc_mixture_set* ms, float mixture_weight_scale, float gaussian_scale,
                     const float* feats, long T, float* scores, uint32_t* best, int use_fma);
/* multi-threaded driver over frame slices (the reference's only parallel mode is independent
 * processes over corpus partitions; threads over frame ranges are the same thing for a dense scorer) */
/* Mm::BatchIntFeatureScorer (u8-quantised means / features, s32 distances; src/Mm/BatchFeatureScorer.cc:321-510) */
int orc_gmm_batch_int(const orc_mixture_set* ms, const float* feats, long T, float* scores, int n_threads);
/* Mm::SimdGaussDiagonalMaximumFeatureScorer ("SIMD-diagonal-maximum"): u8-quantised, one quantised feature vector
 * per covariance, best density-in-mixture index (optional) */
int orc_gmm_simd_diag_max(const orc_mixture_set* ms, const float* feats, long T, float* scores, uint32_t* best, int n_threads);
int orc_gmm_simd_model(const orc_mixture_set* ms, uint8_t* means, int32_t* consts, float* isd, float* scaling_squared);
int orc_gmm_batch_int_model(const orc_mixture_set* ms, uint8_t* means, int32_t* consts, float* variance, float* scale,
                            int* padded);
/* Mm::BatchPreselectionFloatFeatureScorer ("preselection-batch-float", src/Mm/BatchFeatureScorer.cc:257-315 with
 * Mm::DensityClustering<f32, f32>): k-means clusters of the density means, only the densities of the `select`
 * clusters nearest to the frame are scored, mixtures left without one get `backoff`. */
int orc_gmm_preselect_float(const orc_mixture_set* ms, const float* feats, long T, float* scores, int use_fma,
                            int clusters, int select, int iterations, float backoff, uint32_t* cluster_of,
                            float* cluster_means, int* n_clusters);
/* Mm::BatchPreselectionIntFeatureScorer ("preselection-batch-int", src/Mm/BatchFeatureScorer.cc:514-577).  Checker for
 * a kernel that is not built yet: restated_sort != 0 picks the clusters with oracle/std_sort_restated.h
 * instead of std::sort -- the two must agree. */
int orc_gmm_preselect_int(const orc_mixture_set* ms, const float* feats, long T, float* scores, int clusters, int select,
                          int iterations, uint32_t* cluster_of, int restated_sort);
/* (key, index) pairs sorted by key only: index permutation of std::sort (restated = 0) / of the restatement */
void orc_sort_pairs(const int32_t* keys, int n, int32_t* perm, int restated);
/* number of times the restatement fell back to heap sort so far; an adversarial input (McIlroy) for this std::sort */
long orc_sort_heap_calls(void);
void orc_sort_killer(int n, int32_t* keys);
int orc_gmm_batch_float_mt(const orc_mixture_set* ms, const float* feats, long T, float* scores, int use_fma,
                           int n_threads);

/* ---------------------------------------------------------------- Nn scorer */

enum { ORC_ACT_LINEAR = 0, ORC_ACT_SIGMOID = 1, ORC_ACT_RELU = 2, ORC_ACT_SOFTMAX = 3, ORC_ACT_TANH = 4 };
enum { ORC_NN_F32 = 0, ORC_NN_F64ACC = 1, ORC_NN_BF16 = 2 };

/* Feed-forward network.  dims[n_layers+1]; weights[l] is in x out column-major exactly as
 * Nn::LinearLayer keeps it (element (i,o) at [o*in+i]); bias[l] is [out].  x is [T * dims[0]]
 * (a frame is a contiguous column of the reference's dim x T matrix); out is [T * dims[n_layers]]. */
int orc_nn_forward(int n_layers, const int* dims, const int* act, const float* const* weights,
                   const float* const* bias, const float* x, long T, float* out, int mode);

/* Nn::BatchFeatureScorer: softmax of the top layer not evaluated, scaled log-prior removed from the
 * output bias, score = -activation.  log_prior may be NULL (scale 0). */
int orc_nn_scores(int n_layers, const int* dims, const int* act, const float* const* weights,
                  const float* const* bias, const float* log_prior, float prior_scale, const float* x,
                  long T, float* scores, int mode);

/* double-precision variants used to check the reference's own f64 unit-test vectors */
int orc_nn_forward_f64(int n_layers, const int* dims, const int* act, const double* const* weights,
                       const double* const* bias, const double* x, long T, double* out);

const char* orc_version(void);

/* feature post-processing between front-end and scorer (postproc_oracle.cc): signal-normalization
 * (type 1 mean, 2 mean-and-variance; length / right < 0 = "infinite"), sequence concatenation, matrix multiplication */
int orc_normalize(int type, long length, long right, const float* feats, const long* frame_offsets, int n_utt, int dim,
                  float* out, int use_fma);
int orc_splice(int length, int right, const float* feats, const long* frame_offsets, int n_utt, int dim, float* out);
int orc_matmul(const float* M, int rows, int cols, const float* x, long T, float* y, int use_fma);

/* Search::LinearSearch on a flat lexicon (search_oracle.cc) */
typedef struct {
    uint32_t        n_words;
    const uint32_t* word_offsets;    /* [n_words+1] into the state arrays */
    const uint32_t* state_emission;  /* emission (mixture) index of every HMM state */
    const uint32_t* state_tdp_model; /* transition model of every state */
    uint32_t        n_models;
    const float*    tdp;             /* [n_models * 4]: loop, forward, skip, exit */
    uint32_t        entry_model;     /* Am::TransitionModel::entryM1 */
    const float*    unigram;         /* [n_words] LM score of entering the word */
    const uint8_t*  word_regular;    /* [n_words] Pronunciation::isRegularWord (1) or not (silence, noise); NULL = all regular */
    int32_t         single_word;     /* LinearSearch "single-word-recognition" (reference default: on) */
} orc_lexicon;
long orc_linear_search(const orc_lexicon* lx, const float* scores, long T, int n_emissions, uint32_t* words,
                       int32_t* times, float* am, float* lm);

#ifdef __cplusplus
}
#endif
#endif
