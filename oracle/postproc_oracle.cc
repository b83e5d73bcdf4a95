/*
 * postproc_oracle.cc -- CPU restatement of the feature post-processing nodes that sit between the MFCC front-end and
 * the scorers in every real RASR system:
 *   signal-normalization (mean / mean-and-variance)      src/Signal/Normalization.cc:41-190, SlidingWindow.hh:397-470
 *   signal-vector-f32-sequence-concatenation             src/Signal/VectorSequenceConcatenation.hh:89-103
 *   signal-matrix-multiplication-f32                     src/Signal/MatrixMult.hh, src/Math/Matrix.hh:487-494,
 *                                                         src/Math/Vector.hh:95-101
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY PINNED to the reference's own nodes (src/Signal compiled into
 * oracle/_ref, wired by its NetworkParser from oracle/refbuild/flows/postproc_chain.flow): bit for bit over the window / length /
 * output-point combinations of tests/test_ref_parity.py.  The restatement follows the cited lines including the f32 / f64
 * mixing and the update order of the running sums.
 * Build with -ffp-contract=off; use_fma selects the contraction the reference's default build applies.
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

/* One segment, one dimension at a time would hide the reference's loop structure; keep it frame-major as written:
 * Normalization::update (:41-60): add -> removed -> updateStatistics(add, removed) -> if the window has an element
 * at the output point: normalize (finalize if the statistics changed, then apply).  flush (:62-64) moves the window
 * WITHOUT touching the statistics.  With length L and output point R (SlidingWindow::init :397-409) frame i leaves
 * the window when frame i+L is added, frame i-R is emitted when frame i is added, the last R frames at flush. */
template<bool Fuse>
void normalize_segment(int type, long L, long R, const float* x, long T, int D, float* out) {
    std::vector<double> sum(D, 0.0), sumSq(D, 0.0);
    std::vector<float>  mean(D, 0.0f), sd(D, 1.0f);
    double              w       = 0;
    bool                changed = true;
    auto finalize = [&]() {
        if (w > 0) {
            for (int d = 0; d < D; ++d)
                mean[d] = (float)(sum[d] / w); /* MeanNormalization::finalize :123-129 */
            if (type == 2)
                for (int d = 0; d < D; ++d) { /* MeanAndVarianceNormalization::finalize :166-180 */
                    sd[d] = (float)std::sqrt((sumSq[d] - sum[d] * sum[d] / w) / w);
                    if (sd[d] == 0)
                        sd[d] = 1.0f;
                }
        }
        changed = false;
    };
    auto apply = [&](long t) {
        if (changed)
            finalize();
        for (int d = 0; d < D; ++d) {
            float v = x[t * D + d] - mean[d]; /* :131-133 */
            if (type == 2)
                v = v / sd[d]; /* :182-185 */
            out[t * D + d] = v;
        }
    };
    long emitted = 0;
    for (long i = 0; i < T; ++i) {
        for (int d = 0; d < D; ++d) { /* add */
            const double a = (double)x[i * D + d];
            sum[d] += a;
            if (type == 2)
                sumSq[d] = Fuse ? std::fma(a, a, sumSq[d]) : sumSq[d] + a * a;
        }
        if (i >= L) { /* the element pushed out by this add */
            for (int d = 0; d < D; ++d) {
                const double r = (double)x[(i - L) * D + d];
                sum[d] -= r;
                if (type == 2)
                    sumSq[d] = Fuse ? std::fma(-r, r, sumSq[d]) : sumSq[d] - r * r;
            }
        }
        w += 1;
        if (i >= L)
            w -= 1;
        changed = true;
        if (i >= R) {
            apply(i - R);
            emitted = i - R + 1;
        }
    }
    for (long t = emitted; t < T; ++t) /* flush */
        apply(t);
}

}  // namespace

/* type: 1 mean, 2 mean-and-variance; length/right < 0 = "infinite" (whole segment) */
extern "C" int orc_normalize(int type, long length, long right, const float* feats, const long* frame_offsets,
                             int n_utt, int dim, float* out, int use_fma) {
    if (type != 1 && type != 2)
        return -1;
    const long INF = 2147483647L;
    long       L = length < 0 ? INF : length, R = right < 0 ? INF : right;
    if (L >= INF && R >= INF)
        --R; /* SlidingWindow::init special case :399-401 */
    if (L <= R)
        return -2; /* "Cannot initialize with parameters ..." */
    for (int u = 0; u < n_utt; ++u) {
        const long a = frame_offsets[u], T = frame_offsets[u + 1] - a;
        if (use_fma)
            normalize_segment<true>(type, L, R, feats + a * dim, T, dim, out + a * dim);
        else
            normalize_segment<false>(type, L, R, feats + a * dim, T, dim, out + a * dim);
    }
    return 0;
}

/* VectorSequenceConcatenation::putData :89-103 over a DelayNode window (max-size = length, right), margin policy
 * copy (the nearest existing frame stands in for a missing one), oldest frame first */
extern "C" int orc_splice(int length, int right, const float* feats, const long* frame_offsets, int n_utt, int dim,
                          float* out) {
    if (length <= right || right < 0)
        return -2;
    const int past = length - right - 1;
    for (int u = 0; u < n_utt; ++u) {
        const long a = frame_offsets[u], T = frame_offsets[u + 1] - a;
        for (long t = 0; t < T; ++t)
            for (int rel = -past; rel <= right; ++rel) {
                const long s = std::min(std::max(t + rel, 0L), T - 1);
                std::copy(feats + (a + s) * dim, feats + (a + s + 1) * dim,
                          out + ((a + t) * length + (rel + past)) * dim);
            }
    }
    return 0;
}

/* y = M x per frame: Math::Matrix::operator*(Vector) -> Vector::operator*(Vector), sequential f32 dot */
extern "C" int orc_matmul(const float* M, int rows, int cols, const float* x, long T, float* y, int use_fma) {
    for (long t = 0; t < T; ++t)
        for (int n = 0; n < rows; ++n) {
            float r = 0.0f;
            for (int i = 0; i < cols; ++i)
                r = use_fma ? std::fmaf(M[(size_t)n * cols + i], x[t * cols + i], r)
                            : r + M[(size_t)n * cols + i] * x[t * cols + i];
            y[t * rows + n] = r;
        }
    return 0;
}
