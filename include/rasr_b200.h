/*
 * rasr_b200.h -- C ABI of librasr_b200.so: the B200-native acoustic front-end and emission-score
 * engine that sits behind RASR's Flow::Node and Mm::FeatureScorer interfaces.
 *
 * The reference (rwth-i6/rasr) has no FFI / plugin ABI of its own: extensions are C++ classes that
 * register themselves into Flow::Registry and Mm::FeatureScorerFactory at link time
 * (src/Flow/Registry.hh:49-51, src/Mm/FeatureScorerFactory.hh:54-66).  The entry points below are
 * what those two adapter classes (adapters/B200MfccNode.cc, adapters/B200FeatureScorer.cc, see
 * INTEGRATION.md) bind; each one names the reference interface it replaces.  All paths are
 * relative to the reference checkout.
 *
 * Conventions
 *   - plain C: opaque handles, plain pointers and sizes, no C++/torch types.
 *   - every call returns RB_OK (0) or a negative rb_status; rb_last_error() gives the text of the
 *     last failure on the calling thread.  Nothing throws, nothing aborts (the reference's
 *     "no exceptions" rule, doc/architecture.rst:168; the adapter maps != 0 to criticalError()).
 *   - host buffers are caller-owned; the library never frees or keeps caller memory.
 *   - "*_dev" variants take DEVICE pointers on the handle's device and a cudaStream_t passed as
 *     void* (NULL = the handle's own stream); they only enqueue work.
 *   - a handle is bound to one CUDA device and is not thread-safe (the Flow pull graph and the
 *     recognizer are single-threaded: src/Flow/Network.cc:507-538).
 *   - there is NO CPU fallback: without a usable CUDA device every create call fails with
 *     RB_ERR_NO_DEVICE.
 */
#ifndef RASR_B200_H
#define RASR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RB_OK              = 0,
    RB_ERR_INVALID     = -1, /* bad argument / configuration */
    RB_ERR_NO_DEVICE   = -2, /* no CUDA device, or device is not sm_100 */
    RB_ERR_CUDA        = -3, /* a CUDA runtime call or kernel failed */
    RB_ERR_UNSUPPORTED = -4, /* valid in the reference but outside this engine (documented) */
    RB_ERR_STATE       = -5, /* call order violates the node / scorer protocol */
    RB_ERR_NOMEM       = -6
} rb_status;

const char* rb_last_error(void);
const char* rb_version(void);
/* number of usable sm_100 devices (0 when there is none; never fails) */
int rb_device_count(void);
/* kernels launched by this library since load (all handles, all threads) */
uint64_t rb_launch_count(void);

/* -------------------------------------------------------------------------------------
 * Page-locked host memory.  Every host-buffer entry point below (rb_frontend_process*, rb_gmm_score, rb_nn_score,
 * rb_pipeline_*) accepts any host pointer, but only from page-locked memory do the copies run asynchronously at the
 * PCIe rate and overlap with the kernels (pageable memory is staged by the driver, every copy synchronous).  The
 * adapters keep the buffered segment -- the reference's `std::vector<f32>` feature / score buffers, e.g.
 * src/Mm/BatchFeatureScorer.hh:164-166, src/Nn/BatchFeatureScorer.hh -- in memory from rb_host_alloc; a host that
 * cannot move its buffers registers them in place.
 * ------------------------------------------------------------------------------------- */
int  rb_host_alloc(size_t bytes, void** out); /* cudaHostAlloc, usable from every device */
void rb_host_free(void* p);
int  rb_host_register(void* p, size_t bytes); /* page-lock an existing allocation in place */
int  rb_host_unregister(void* p);
int  rb_host_is_pinned(const void* p);        /* 1: page-locked, 0: pageable / unknown */

/* =====================================================================================
 * Front-end: the Flow network of src/Tools/FeatureExtraction/share/mfcc.flow:8-34
 *   signal-preemphasis -> signal-window (hamming) -> signal-real-fast-fourier-transform ->
 *   signal-vector-alternating-complex-f32-amplitude -> signal-filterbank (mel) ->
 *   generic-vector-f32-log -> signal-cosine-transform
 * followed by derivationWithRegression.flow:7-27 (signal-delay / signal-regression x2 /
 * generic-vector-f32-concat) when derivatives != 0.
 * ===================================================================================== */

typedef struct rb_frontend rb_frontend;

typedef struct {
    double sample_rate;       /* "sample-rate" attribute of the incoming samples stream */
    double window_length_s;   /* signal-window length=   (src/Signal/Window.cc:69-82) */
    double window_shift_s;    /* signal-window shift= */
    double fft_max_input_s;   /* signal-real-fast-fourier-transform maximum-input-size=
                                 (src/Signal/FastFourierTransform.hh:298-308) */
    double filter_width;      /* signal-filterbank filter-width= in mel (mfcc.flow:24) */
    float  preemphasis_alpha; /* signal-preemphasis alpha= (src/Signal/Preemphasis.cc:79) */
    int    n_cepstra;         /* signal-cosine-transform nr-outputs= */
    int    derivatives;       /* 0: static cepstra; 1: static || delta || delta-delta */
    int    device;            /* CUDA device ordinal */
    int    window_type;       /* rb_window_type; signal-window "type" (src/Signal/WindowFunction.cc:25-33) */
} rb_frontend_cfg;

/* window functions of src/Signal/WindowFunction.cc:62-132; 0 = the default of the reference and of mfcc.flow */
enum rb_window_type {
    RB_WINDOW_HAMMING = 0,
    RB_WINDOW_RECTANGULAR = 1,
    RB_WINDOW_HANNING = 2,
    RB_WINDOW_PERIODIC_HANNING = 3,
    RB_WINDOW_BARTLETT = 4,
    RB_WINDOW_BLACKMAN = 5,
    RB_WINDOW_KAISER = 6 /* beta = 0 as WindowFunction::create builds it (no node parameter sets beta): all ones */
};

typedef struct {
    int win_length; /* samples */
    int win_shift;  /* samples */
    int fft_length; /* points */
    int n_bins;     /* fft_length/2+1 */
    int n_filters;
    int n_weights;  /* non-zero filterbank taps */
    int feat_dim;   /* n_cepstra * (derivatives ? 3 : 1) */
} rb_frontend_geometry;

/* fills *cfg with the mfcc.flow defaults for 16 kHz audio (13 cepstra + derivatives) */
void rb_frontend_default_cfg(rb_frontend_cfg* cfg);

/* node construction + configure(): builds window / twiddle / filterbank / DCT tables on the host in
 * f64 exactly as the reference's init code does and uploads them.
 * replaces: PreemphasisNode, WindowNode, FastFourierTransformNode, FilterBankNode::init
 * (src/Signal/Filterbank.cc:765-797), CosineTransformNode::configure (src/Signal/CosineTransform.cc:199-211) */
int rb_frontend_create(const rb_frontend_cfg* cfg, rb_frontend** out);
void rb_frontend_destroy(rb_frontend* h);
int rb_frontend_get_geometry(const rb_frontend* h, rb_frontend_geometry* g);
/* host copies of the tables the kernels use (any pointer may be NULL):
 * window[win_length], fb_start/fb_end[n_filters], fb_weights dense [n_filters*n_bins], dct[n_cepstra*n_filters] */
int rb_frontend_get_tables(const rb_frontend* h, float* window, int* fb_start, int* fb_end, float* fb_weights,
                           float* dct);
/* frames the window node emits for one segment of n samples (WindowBuffer get/flush protocol,
 * src/Signal/WindowBuffer.cc:50-126, flush-all=false) */
long rb_frontend_nframes_for(const rb_frontend* h, long n_samples);
/* Timestamp (src/Flow/Timestamp.hh:39-44) of every packet the chain emits for a segment of n_samples samples whose first
 * sample is at start_time: start accumulated by repeated += shift / sample rate as WindowBuffer does
 * (src/Signal/WindowBuffer.cc:94), merged over the regression window when derivatives are on
 * (src/Flow/Merger.hh:79-101).  t_start / t_end [frames] may be NULL; returns the number of frames.  Host only. */
long rb_frontend_timestamps(const rb_frontend* h, long n_samples, double start_time, double* t_start, double* t_end);

/* --- streaming protocol = what Flow::Node::work() sees (src/Signal/SlidingAlgorithmNode.hh:60-80):
 * packets of samples arrive in time order; on the EOS sentinel the segment is computed. */
int  rb_frontend_reset(rb_frontend* h); /* EOS / new segment: state reset (src/Signal/Preemphasis.cc:96-106) */
int  rb_frontend_push(rb_frontend* h, const float* samples, long n, double start_time);
int  rb_frontend_finish(rb_frontend* h); /* runs the kernels on everything pushed since reset */
long rb_frontend_nframes(const rb_frontend* h);
/* feats [T*feat_dim] row-major (a frame = one Flow::Vector<f32>); t_start/t_end [T] = Timestamp of
 * each emitted packet (src/Flow/Timestamp.hh:39-44); any pointer may be NULL */
int rb_frontend_read(rb_frontend* h, float* feats, double* t_start, double* t_end);

/* --- batch of independent segments (utterances) in one call: samples of utterance u are
 * samples[offsets[u] .. offsets[u+1]).  frame_offsets [n_utt+1] receives the frame prefix sums;
 * feats [frame_offsets[n_utt] * feat_dim].  Use rb_frontend_count_frames first to size buffers. */
long rb_frontend_count_frames(const rb_frontend* h, const int64_t* offsets, int n_utt, int64_t* frame_offsets);
int  rb_frontend_process(rb_frontend* h, const float* samples, const int64_t* offsets, int n_utt, float* feats,
                         double* t_start, double* t_end);
/* the same fed with 16-bit PCM as the audio nodes deliver it: interleaved s16 frames of n_channels channels, of which
 * `track` is demultiplexed and converted (generic-vector-s16-demultiplex + generic-convert-vector-s16-to-vector-f32,
 * src/Tools/FeatureExtraction/share/samples.flow:13-18); offsets count sample frames.  Halves the bytes that cross
 * PCIe.  signal-dc-detection (samples.flow:35-37) is not applied. */
int  rb_frontend_process_s16(rb_frontend* h, const int16_t* samples, int n_channels, int track, const int64_t* offsets,
                             int n_utt, float* feats, double* t_start, double* t_end);
/* device variant: d_samples / d_feats are device pointers, offsets stays on the host.
 * d_stage (optional, device, [total_frames*n_cepstra] floats) receives the static cepstra. */
int rb_frontend_process_dev(rb_frontend* h, const float* d_samples, const int64_t* offsets, int n_utt,
                            float* d_feats, void* stream);
/* per-stage dumps of the last process/finish call for parity tests (host pointers, any may be NULL):
 * amplitude [T*n_bins], fbank [T*n_filters] (before log), cepstra [T*n_cepstra].  Only filled when
 * rb_frontend_set_debug(h, 1) was called before processing. */
int rb_frontend_set_debug(rb_frontend* h, int on);
int rb_frontend_read_stages(rb_frontend* h, float* amplitude, float* fbank, float* cepstra);

/* -------------------------------------------------------------------------------------
 * signal-dc-detection in front of the MFCC chain (src/Signal/DcDetection.{hh,cc}; wired by
 * src/Tools/FeatureExtraction/share/samples.flow:34-37): stretches of at least min_dc_length_s whose samples stay
 * within max_dc_increment of the last non-DC sample are discarded, as are non-DC segments shorter than
 * min_non_dc_segment_length_s.  Every kept run of samples is framed on its own (pre-emphasis and window restart
 * at the gap, Preemphasis.cc:54-55, SlidingAlgorithmNode.hh:83-99) and starts at its own time; the derivative
 * window runs across the runs of an utterance (the delay node ignores time stamps, src/Signal/Delay.cc:137-160).
 * ------------------------------------------------------------------------------------- */
typedef struct rb_dc_cfg {
    double min_dc_length_s;             /* DcDetectionNode::paramMinDcLength,  DcDetection.cc:231-232 */
    float  max_dc_increment;            /* paramMaxDcIncrement :234-235; 0: every sample is non-DC */
    double min_non_dc_segment_length_s; /* paramMinNonDcSegmentLength :237-238 */
    int    maximal_output_size;         /* paramMaximalOutputSize :240-241 (only shapes the accumulated start times) */
} rb_dc_cfg;
void rb_dc_default_cfg(rb_dc_cfg* cfg); /* the values samples.flow sets: .0125, 0.9, .026; 4096 */
/* upper bound of the frames rb_frontend_process_dc can produce for these utterance lengths */
long rb_frontend_dc_max_frames(const rb_frontend* h, const rb_dc_cfg* dc, const int64_t* offsets, int n_utt);
/* like rb_frontend_process, but the frame counts depend on the data: feats/t_start/t_end have room for `capacity`
 * frames (RB_ERR_INVALID if that is too small), frame_offsets [n_utt + 1] receives the frame range of every
 * utterance. */
int rb_frontend_process_dc(rb_frontend* h, const rb_dc_cfg* dc, const float* samples, const int64_t* offsets, int n_utt,
                           float* feats, long capacity, int64_t* frame_offsets, double* t_start, double* t_end);
/* streaming interface (rb_frontend_push / finish / read): run the detector over the pushed samples at
 * rb_frontend_finish; NULL switches it off again.  Start times continue from the start time of the first packet. */
int rb_frontend_set_dc_detection(rb_frontend* h, const rb_dc_cfg* dc);
/* the kept sample runs of the last rb_frontend_process_dc call: utterance, [begin, end) in the caller's buffer,
 * start time relative to the utterance.  Any pointer may be NULL.  *sequential_path != 0: the input was not
 * "equal to or an increment away from its predecessor" everywhere and the reference chain was replayed sample by
 * sample.  Returns the number of runs. */
long rb_frontend_dc_runs(const rb_frontend* h, int64_t* run_utt, int64_t* run_begin, int64_t* run_end,
                         double* run_start, long capacity, int* sequential_path);

/* =====================================================================================
 * GMM emission scorer behind Mm::FeatureScorer (src/Mm/FeatureScorer.hh:28-167)
 * ===================================================================================== */

typedef struct rb_gmm rb_gmm;

/* flat view of Mm::MixtureSet (src/Mm/MixtureSet.hh:123-201): mixtures -> (density, log weight),
 * density -> (mean, covariance), means f32[dim], diagonal variances f32[dim] */
typedef struct {
    uint32_t        dim;
    uint32_t        n_mixtures;
    uint32_t        n_densities;
    uint32_t        n_means;
    uint32_t        n_covariances;
    const uint32_t* mix_offsets;    /* [n_mixtures+1] into mix_density / mix_log_weight */
    const uint32_t* mix_density;    /* density index of each mixture entry */
    const double*   mix_log_weight; /* natural-log weight of each entry (Mm::Weight is f64) */
    const uint32_t* dens_mean;      /* [n_densities] */
    const uint32_t* dens_cov;       /* [n_densities] */
    const float*    means;          /* [n_means*dim] */
    const float*    variances;      /* [n_covariances*dim] */
} rb_mixture_set;

typedef enum {
    /* Mm::BatchFloatFeatureScorer (src/Mm/BatchFeatureScorer.cc:164-253): one pooled covariance,
     * maximum approximation, arithmetic in the reference's SSE lane order (bit-identical scores) */
    RB_GMM_BATCH_FLOAT = 0,
    /* Mm::GaussDiagonalMaximumFeatureScorer (src/Mm/GaussDiagonalMaximumFeatureScorer.cc:116-233):
     * per-density covariance, maximum approximation, returns the best density */
    RB_GMM_DIAG_MAX = 1,
    /* Mm::GaussDiagonalSumFeatureScorer (same file :239-290): log-sum-exp over the mixture */
    RB_GMM_DIAG_SUM = 2,
    /* tensor-core formulation of RB_GMM_BATCH_FLOAT (split-precision GEMM |x|^2 - 2 x.mu + |mu|^2 on
     * tcgen05, f32 accumulate); scores within 1e-4 relative of the reference, not bit-identical */
    RB_GMM_BATCH_TENSOR = 3,
    /* Mm::BatchIntFeatureScorer / BatchUnrolledIntFeatureScorer ("batch-diagonal-maximum-int" / "-fast",
     * src/Mm/BatchFeatureScorer.cc:321-510, 581-650): means and features quantised to u8, s32 distances;
     * integer arithmetic on the tensor cores (IMMA), bit-identical scores */
    RB_GMM_BATCH_INT = 4,
    /* Mm::BatchPreselectionFloatFeatureScorer ("preselection-batch-float", src/Mm/BatchFeatureScorer.cc:257-315) with
     * Mm::DensityClustering<f32, f32> (src/Mm/DensityClustering.{hh,cc,tcc}): the density means are clustered once
     * (k-means, the reference's pseudo-random initialisation), per frame only the densities of the clusters nearest
     * to the feature are scored, mixtures left without one get the back-off score.  Reproduces the reference's
     * approximation (same clustering, same per-frame cluster choice, same scores); defaults as in
     * DensityClustering.cc:20-34: 256 clusters, 32 selected, 5 iterations, back-off 40000. */
    RB_GMM_BATCH_PRESELECT = 5,
    /* Mm::BatchPreselectionIntFeatureScorer ("preselection-batch-int", src/Mm/BatchFeatureScorer.cc:514-577) with
     * Mm::DensityClustering<u8, s32>: the same scheme on the u8 model of RB_GMM_BATCH_INT -- s32 distances, centroids
     * truncated to u8, no back-off score (a mixture without a scored density gets (f32)INT_MAX / scale).  Equal
     * distances at the selection boundary are resolved like the reference's std::sort (libstdc++ introsort restated,
     * rasr_b200/csrc/introsort.cuh): bit-identical scores. */
    RB_GMM_BATCH_PRESELECT_INT = 6,
    /* Mm::SimdGaussDiagonalMaximumFeatureScorer ("SIMD-diagonal-maximum", src/Mm/SimdFeatureScorer.{hh,cc}): u8-quantised
     * means, one u8-quantised copy of the feature vector per covariance, s32 distances, first best density reported;
     * scores and densities bit-identical (feature dimension <= 64).  mixture_weight_scale / gaussian_scale must be 1. */
    RB_GMM_SIMD_DIAG_MAX = 7
} rb_gmm_mode;

/* contraction: 1 = fused multiply-add where the reference's default build (gcc -O2 -march=native,
 * -ffp-contract=fast) fuses; 0 = every operation rounded separately (strict build) */
int  rb_gmm_create(const rb_mixture_set* ms, int mode, float mixture_weight_scale, float gaussian_scale,
                   int contraction, int device, rb_gmm** out);
void rb_gmm_destroy(rb_gmm* h);
/* RB_GMM_BATCH_PRESELECT / _PRESELECT_INT only: rebuild the clustering with other parameters
 * (density-clustering.clusters, .select-clusters, .iterations, .backoff-score; the int variant ignores the back-off
 * score) / read it back: cluster_of_density [n_densities] in mixture order, cluster_means [n_clusters * padded
 * dimension] (padded = dim rounded up to 8; int variant: rounded up to 16, the u8 centroids as f32); any pointer may
 * be NULL */
int  rb_gmm_configure_preselection(rb_gmm* h, int clusters, int select, int iterations, float backoff_score);
int  rb_gmm_get_clustering(const rb_gmm* h, uint32_t* cluster_of_density, float* cluster_means, int* n_clusters);
int  rb_gmm_n_mixtures(const rb_gmm* h);
int  rb_gmm_dim(const rb_gmm* h);
/* dense scoring of T frames against ALL mixtures: scores [T*n_mixtures] row-major,
 * scores[t][m] = -log p(x_t | m) (natural log, f32; Mm::Score, src/Mm/Types.hh:26).
 * best_density (optional, modes DIAG_*): index within the mixture of the best density.
 * replaces: fillScoreCacheTpl (src/Mm/BatchFeatureScorer.cc:207-253) for every emission and frame,
 * i.e. what Speech::FeatureScorerNode materialises per frame (src/Speech/FeatureScorerNode.cc:95-111) */
int rb_gmm_score(rb_gmm* h, const float* feats, long T, float* scores, uint32_t* best_density);
int rb_gmm_score_dev(rb_gmm* h, const float* d_feats, long T, float* d_scores, uint32_t* d_best_density,
                     void* stream);
/* the same scores into n_dst (1..16) destinations at once, row for row: d_dst[k] [T * n_mixtures] may lie in the HBM of
 * another GPU of the box (rb_comm_window_ptr) -- the all-gather of the score matrix fused into the scorer.  The exact
 * route of RB_GMM_BATCH_FLOAT stores to every destination from its last kernel (64 contiguous bytes per frame and store
 * instruction, crossing NVLink while the kernel is still computing); the other modes copy the finished matrix. */
int rb_gmm_score_fanout_dev(rb_gmm* h, const float* d_feats, long T, int n_dst, float* const* d_dst, void* stream);
/* measurement hook (bench.py): RB_GMM_BATCH_FLOAT scores its batches in three kernels (operand split,
 * tensor-core screening of the candidate densities, exact evaluation of the candidates); with timing on, CUDA events
 * bracket them on the launching stream and rb_gmm_get_timing returns their durations of the last such call
 * (ms3 = split, screen, refine; waits for the call; RB_ERR_STATE if the last calls took the single direct kernel) */
int rb_gmm_set_timing(rb_gmm* h, int on);
int rb_gmm_get_timing(rb_gmm* h, float* ms3);

/* =====================================================================================
 * Legacy Nn feed-forward scorer (src/Nn/BatchFeatureScorer.cc:45-171, src/Nn/NeuralNetwork.cc:313-425)
 * ===================================================================================== */

typedef struct rb_nn rb_nn;

enum { RB_ACT_LINEAR = 0, RB_ACT_SIGMOID = 1, RB_ACT_RELU = 2, RB_ACT_SOFTMAX = 3, RB_ACT_TANH = 4 };
/* F32: f32 operands and accumulation on CUDA cores (parity with cblas_sgemm to 1e-4);
 * BF16: bf16 operands on tcgen05 tensor cores, f32 accumulation in TMEM (the fast path) */
enum { RB_NN_F32 = 0, RB_NN_BF16 = 1 };

/* dims[n_layers+1]; weights[l] is the memory of the reference's in x out column-major matrix
 * (element (i,o) at [o*in+i], src/Nn/LinearLayer.cc:402-420); bias[l] is [out] (may be NULL);
 * log_prior [dims[n_layers]] or NULL (src/Nn/Prior.cc:159-188, removed from the output bias scaled by
 * prior_scale: src/Nn/LinearLayer.cc:499-519). */
int  rb_nn_create(int n_layers, const int* dims, const int* act, const float* const* weights,
                  const float* const* bias, const float* log_prior, float prior_scale, int precision,
                  int device, rb_nn** out);
void rb_nn_destroy(rb_nn* h);
int  rb_nn_n_outputs(const rb_nn* h);
int  rb_nn_n_inputs(const rb_nn* h);
/* Nn::ClassLabelWrapper (src/Nn/ClassLabelWrapper.cc:57-100): class_to_output [n_classes] maps an emission class to a
 * network output, -1 = disregarded class.  rb_nn_score[_dev] then writes [T * n_classes] scores, FLT_MAX for the
 * disregarded classes (src/Nn/BatchFeatureScorer.cc:163-169); rb_nn_forward is not affected.  NULL removes the
 * mapping.  rb_nn_n_emissions = n_classes, or the number of outputs without a mapping. */
int  rb_nn_set_class_mapping(rb_nn* h, int n_classes, const int32_t* class_to_output);
int  rb_nn_n_emissions(const rb_nn* h);
/* scores [T*n_out]: score = -(w.h + b - prior_scale*log_prior), top-layer softmax NOT evaluated
 * (src/Nn/BatchFeatureScorer.cc:64,148-171) */
int rb_nn_score(rb_nn* h, const float* feats, long T, float* scores);
int rb_nn_score_dev(rb_nn* h, const float* d_feats, long T, float* d_scores, void* stream);
/* plain forward pass incl. the top-layer activation = neural-network-forward Flow node
 * (src/Nn/NeuralNetworkForwardNode.cc:187-256) */
int rb_nn_forward(rb_nn* h, const float* feats, long T, float* out);
int rb_nn_forward_dev(rb_nn* h, const float* d_feats, long T, float* d_out, void* stream);

/* =====================================================================================
 * Feature post-processing between the front-end and the scorers: the nodes of
 * src/Tools/FeatureExtraction/share/processing.standard_system.flow:25-27 and lda.flow:11-19
 *   signal-normalization (type mean | mean-and-variance, length / right)   src/Signal/Normalization.cc:41-190
 *   signal-vector-f32-sequence-concatenation (max-size, right)             src/Signal/VectorSequenceConcatenation.hh:89-103
 *   signal-matrix-multiplication-f32 (file = rows x cols matrix)           src/Signal/MatrixMult.hh
 * chained on the device in this order; every stage is optional.
 * ===================================================================================== */

typedef struct rb_postproc rb_postproc;

typedef struct {
    int          norm_type;     /* 0: none, 1: mean, 2: mean-and-variance */
    long         norm_length;   /* sliding window in frames; < 0 = "infinite" (whole segment) */
    long         norm_right;    /* output point counted from the newest frame; < 0 = "infinite" */
    int          splice_length; /* max-size of the concatenation window; 0: none */
    int          splice_right;  /* frames to the right of the present frame */
    int          matrix_rows;   /* output dimension of the matrix multiplication */
    int          matrix_cols;   /* must equal the (spliced) input dimension */
    const float* matrix;        /* row-major rows x cols (Math::Matrix<f32>); NULL: none; copied by create */
    int          contraction;   /* as rb_gmm_create */
    int          device;
} rb_postproc_cfg;

int  rb_postproc_create(const rb_postproc_cfg* cfg, int dim_in, rb_postproc** out);
void rb_postproc_destroy(rb_postproc* h);
int  rb_postproc_dim_out(const rb_postproc* h);
/* feats [frame_offsets[n_utt] * dim_in], segments = [frame_offsets[u], frame_offsets[u+1]) (statistics and
 * windows never cross a segment boundary: the nodes reset at EOS); out [frames * dim_out].  in and out must not
 * alias.  The device variant only enqueues work on `stream` (frame_offsets stays on the host). */
int rb_postproc_process(rb_postproc* h, const float* feats, const int64_t* frame_offsets, int n_utt, float* out);
int rb_postproc_process_dev(rb_postproc* h, const float* d_feats, const int64_t* frame_offsets, int n_utt,
                            float* d_out, void* stream);

/* =====================================================================================
 * Fused audio -> scores pipeline (config C3): front-end and GMM scorer chained on the device,
 * features never leave HBM.  scores [total_frames * n_mixtures].
 * ===================================================================================== */
int rb_pipeline_score(rb_frontend* fe, rb_gmm* gmm, const float* samples, const int64_t* offsets, int n_utt,
                      float* scores, float* feats /* optional host copy */);
/* the same fed with interleaved 16-bit PCM (see rb_frontend_process_s16) */
int rb_pipeline_score_s16(rb_frontend* fe, rb_gmm* gmm, const int16_t* samples, int n_channels, int track,
                          const int64_t* offsets, int n_utt, float* scores, float* feats /* optional host copy */);
int rb_pipeline_score_dev(rb_frontend* fe, rb_gmm* gmm, const float* d_samples, const int64_t* offsets, int n_utt,
                          float* d_feats, float* d_scores, void* stream);

/* rb_pipeline_score_dev with the scores fanned out to n_dst destinations (rb_gmm_score_fanout_dev) */
int rb_pipeline_score_fanout_dev(rb_frontend* fe, rb_gmm* gmm, const float* d_samples, const int64_t* offsets, int n_utt,
                                 float* d_feats, int n_dst, float* const* d_dst, void* stream);

/* audio -> MFCC -> (rb_postproc, may be NULL) -> Nn scores (config C4 fed from audio).  scores [frames *
 * rb_nn_n_emissions(nn)] (= n_classes with a class mapping, else the number of outputs);
 * the device variant needs d_feats [frames * feat_dim] and, with a post-processor, d_post [frames * dim_out]. */
int rb_pipeline_nn_score(rb_frontend* fe, rb_postproc* pp, rb_nn* nn, const float* samples, const int64_t* offsets,
                         int n_utt, float* scores);
int rb_pipeline_nn_score_dev(rb_frontend* fe, rb_postproc* pp, rb_nn* nn, const float* d_samples,
                             const int64_t* offsets, int n_utt, float* d_feats, float* d_post, float* d_scores,
                             void* stream);

/* =====================================================================================
 * Score consumer (config C5): Search::LinearSearch -- time-synchronous Viterbi over the linear HMMs of
 * all pronunciations with a unigram LM and one book-keeping entry per frame (src/Search/LinearSearch.cc:233-432) --
 * fed from the dense score matrix the scorers above leave on the device, instead of two virtual calls per HMM state
 * and frame (emissionScores->score(mixture), :346).  The lexicon is passed as flat arrays (what LinearSearch builds
 * from Bliss::Lexicon + Am::AcousticModel in setModelCombination).  Both modes of the reference are implemented:
 * continuous recognition and single-word recognition (its default) with irregular words (silence, noise) around
 * the one regular word.
 * ===================================================================================== */

typedef struct rb_search rb_search;

typedef struct {
    uint32_t        n_words;
    const uint32_t* word_offsets;    /* [n_words+1] into the state arrays (Pronunciation::mixtures_) */
    const uint32_t* state_emission;  /* MixtureItem::mixture of every HMM state */
    const uint32_t* state_tdp_model; /* MixtureItem::stateTransitionModel, as an index into tdp */
    uint32_t        n_models;
    const float*    tdp;             /* [n_models * 4]: loop, forward, skip, exit (src/Am/TransitionModel.hh:32-37) */
    uint32_t        entry_model;     /* Am::TransitionModel::entryM1 */
    const float*    unigram;         /* [n_words] WordPronunciationState::unigramScore */
    const uint8_t*  word_regular;    /* [n_words] Pronunciation::isRegularWord (src/Search/LinearSearch.cc:86-96): 0 for
                                        silence / noise lemmata (empty evaluation sequence); NULL = every word regular */
    int32_t         single_word;     /* the recognizer's "single-word-recognition" (:26-30; the reference's default is
                                        true): sentences are irregular* regular irregular*.  0 = continuous recognition */
} rb_lexicon;

int  rb_search_create(const rb_lexicon* lx, int device, rb_search** out);
void rb_search_destroy(rb_search* h);
/* restart() + feed() for every frame of every segment (segments are independent, one CTA each); scores
 * [frames * n_emissions]; results stay in the handle until the next decode */
int rb_search_decode(rb_search* h, const float* scores, int n_emissions, const int64_t* frame_offsets, int n_utt);
int rb_search_decode_dev(rb_search* h, const float* d_scores, int n_emissions, const int64_t* frame_offsets, int n_utt,
                         void* stream);
/* getCurrentBestSentence of segment utt: word ends in chronological order (word index, 1-based end frame,
 * Book::score without LM, Book::lmScore); capacity = frames of the segment; returns the number of words or < 0 */
long rb_search_traceback(const rb_search* h, int utt, uint32_t* words, int32_t* times, float* am_scores,
                         float* lm_scores);
/* all segments of the last decode at once: word_offsets [n_utt + 1] receives prefix counts into the flat arrays
 * (room for `capacity` words each, any may be NULL; total frames of the decode is always enough).  Returns the
 * total number of words, < 0 on error. */
long rb_search_traceback_all(const rb_search* h, int64_t* word_offsets, uint32_t* words, int32_t* times,
                             float* am_scores, float* lm_scores, long capacity);
/* config C5 in one call: audio -> MFCC -> GMM scores -> LinearSearch.  samples: interleaved 16-bit PCM with
 * n_channels >= 1 (track selects the channel), or f32 when n_channels == 0.  The score matrix stays on the device;
 * results are read with rb_search_traceback / rb_search_traceback_all as after rb_search_decode. */
int rb_pipeline_search(rb_frontend* fe, rb_gmm* gmm, rb_search* ls, const void* samples, int n_channels, int track,
                       const int64_t* offsets, int n_utt);

/* =====================================================================================
 * Exchange of the score matrix between the GPUs of one box.  The reference has no collective: its only
 * parallelism is N independent processes over corpus partitions (src/Bliss/CorpusDescription.cc:173-180,482-491).  This
 * serves the one case where a single decoder rank consumes the frames of every shard: the rows
 * [row_offsets[r], row_offsets[r+1]) of the gathered matrix [row_offsets[world] x row_len] f32 come from rank r.
 * One process per GPU; the processes exchange two small byte strings through whatever the host has (MPI,
 * torch.distributed, files) -- the library has no transport of its own.
 *   rb_comm_create              on every rank (selects `device`)
 *   rb_comm_window_alloc        the window = this rank's copy of the gathered matrix in HBM; `handle`
 *                               [RB_COMM_HANDLE_BYTES] is what the peers need to map it
 *   rb_comm_window_attach       handles [world * RB_COMM_HANDLE_BYTES] in rank order: maps every peer's window (CUDA IPC;
 *                               RB_ERR_UNSUPPORTED without peer access)
 *   rb_comm_window_ptr          device pointer of rank `peer`'s window in THIS process: passing
 *                               (float*)ptr + row_offsets[rank] * row_len as d_scores to rb_gmm_score_dev /
 *                               rb_pipeline_score_dev / rb_nn_score_dev makes the scorer's epilogue store straight into
 *                               the consumer's HBM over NVLink (compute and transfer fused, no local copy)
 *                               (rb_gmm_score_fanout_dev / rb_pipeline_score_fanout_dev take one such pointer per rank: the
 *                               all-gather fused into the scorer)
 *   rb_comm_gather_scores_dev   root < 0: all-gather, else gather to `root`.  RB_COMM_P2P: one kernel of 128-bit loads
 *                               and NVLink stores into the peers' windows (skips the local copy when d_send already is
 *                               this rank's slice of its own window); RB_COMM_NCCL: ncclAllGather (equal shards) or
 *                               grouped in-place ncclBroadcasts (unequal shards; no padding, no staging copy).  Only
 *                               enqueues on `stream` (NULL: the communicator's own).
 *   rb_comm_barrier_dev         device-side barrier of all ranks on `stream`: once it has passed, everything the peers
 *                               stored into this rank's window before THEIR barrier is visible to the kernels enqueued
 *                               behind it (needed after P2P stores; also call it before a window is overwritten).  A
 *                               peer that never arrives traps the kernel after ~10 s instead of hanging the GPU.
 *   rb_comm_nccl_unique_id      rank 0 creates the id [RB_COMM_ID_BYTES], the host distributes it,
 *   rb_comm_nccl_init           every rank joins (collective).  libnccl.so.2 is resolved at run time.
 * ===================================================================================== */
typedef struct rb_comm rb_comm;
#define RB_COMM_HANDLE_BYTES 64
#define RB_COMM_ID_BYTES 128
enum { RB_COMM_P2P = 0, RB_COMM_NCCL = 1 };

int  rb_comm_create(int world, int rank, int device, rb_comm** out);
void rb_comm_destroy(rb_comm* c);
int  rb_comm_world(const rb_comm* c);
int  rb_comm_rank(const rb_comm* c);
int  rb_comm_window_alloc(rb_comm* c, size_t bytes, void** d_window, void* handle);
int  rb_comm_window_attach(rb_comm* c, const void* handles);
int  rb_comm_window_ptr(const rb_comm* c, int peer, void** d_ptr);
int  rb_comm_gather_scores_dev(rb_comm* c, const float* d_send, const int64_t* row_offsets, int row_len, int root,
                               int transport, void* stream);
/* P2P push of a row range: rows [first_row, first_row + n_rows) of the gathered matrix, read from d_send, go to `root`
 * (or to every rank, root < 0).  Lets the host pipeline the exchange slab by slab behind the scorer (second stream). */
int  rb_comm_push_rows_dev(rb_comm* c, const float* d_send, int64_t first_row, int64_t n_rows, int row_len, int root,
                           void* stream);
int  rb_comm_barrier_dev(rb_comm* c, void* stream);
int  rb_comm_nccl_unique_id(void* id);
int  rb_comm_nccl_init(rb_comm* c, const void* id);
int  rb_comm_nccl_version(void); /* e.g. 22809; 0 when libnccl.so.2 cannot be loaded */

/* =====================================================================================
 * Test hook: one bf16 tcgen05 GEMM  D[M x N] = A[M x K] * B[N x K]^T (+bias, activation),
 * A/B f32 on the host, rounded to bf16 on the device.  Used by tests/ only.
 * ===================================================================================== */
/* host-only test hook: the first n values of glibc's rand() after srand(seed), as restated for the clustering */
void rb_test_glibc_rand(unsigned seed, int n, int* out);
/* test hook (host only): (key, index) pairs sorted by key only with csrc/introsort.cuh, the restatement of libstdc++'s
 * std::sort behind RB_GMM_BATCH_PRESELECT_INT; perm [n] = the index order */
void rb_test_introsort(const int* keys, int n, int* perm);
int rb_test_gemm_bf16(const float* a, const float* b, const float* bias, int M, int N, int K, int act,
                      float* d, int device);
/* times `iters` launches of the GEMM (variant 0: 128x256 tile; other values are rejected) on device-resident
 * operands with the bf16 hidden-layer epilogue; *ms_per_iter from CUDA events.  Used by scripts/ only. */
int rb_test_gemm_bench(int M, int N, int K, int variant, int iters, float* ms_per_iter, int device);

#ifdef __cplusplus
}
#endif
#endif
