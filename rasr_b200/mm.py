"""Host-side mirror of Mm::MixtureSet and the buffered Mm::FeatureScorer protocol
(src/Mm/FeatureScorer.hh:28-167, src/Mm/BatchFeatureScorer.hh:34-199).

The reference scorers are lazy: `ContextScorer::score(e)` fills a small cache for emission e on
demand.  Here the whole segment is scored densely on the device when the first scorer is asked for a
score, and `score(e)` is an O(1) lookup into the T x nMix matrix -- the caller protocol
(`reset / addFeature / getScorer / flush / bufferFilled / bufferEmpty`, src/Speech/Recognizer.cc:271-281,
197-205) and the rule that `getScorer(f)` answers for the OLDEST buffered frame
(src/Mm/BatchFeatureScorer.hh:81-91) are preserved.
"""
import ctypes as C

import numpy as np

from . import capi


class MixtureSet:
    """Flat arrays of a mixture set; `from_dict` accepts the layout produced by rasr_b200.synth."""

    def __init__(self, dim, mix_offsets, mix_density, mix_log_weight, dens_mean, dens_cov, means, variances):
        self.dim = int(dim)
        self.mix_offsets = np.ascontiguousarray(mix_offsets, np.uint32)
        self.mix_density = np.ascontiguousarray(mix_density, np.uint32)
        self.mix_log_weight = np.ascontiguousarray(mix_log_weight, np.float64)
        self.dens_mean = np.ascontiguousarray(dens_mean, np.uint32)
        self.dens_cov = np.ascontiguousarray(dens_cov, np.uint32)
        self.means = np.ascontiguousarray(means, np.float32).reshape(-1, self.dim)
        self.variances = np.ascontiguousarray(variances, np.float32).reshape(-1, self.dim)

    @classmethod
    def from_dict(cls, d):
        return cls(**d)

    @classmethod
    def read(cls, path):
        """Load a mixture file the way Mm::Module_::readMixtureSet does: text (".pms" / ".gz",
        doc/file_formats/mixture_file.rst) or, for any other name (".mix"), the accumulator file of the last training
        iteration, estimated on the fly -- rasr_b200.io"""
        from . import io as rio

        return cls(**rio.read_mixture_file(path))

    @property
    def n_mixtures(self):
        return self.mix_offsets.size - 1

    def c_struct(self):
        P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        return capi.MixtureSetC(self.dim, self.n_mixtures, self.dens_mean.size, self.means.shape[0],
                                self.variances.shape[0], P(self.mix_offsets, C.c_uint32),
                                P(self.mix_density, C.c_uint32), P(self.mix_log_weight, C.c_double),
                                P(self.dens_mean, C.c_uint32), P(self.dens_cov, C.c_uint32),
                                P(self.means, C.c_float), P(self.variances, C.c_float))


class GmmScorer:
    """RAII wrapper of rb_gmm_*: dense scoring of frames against every mixture."""

    MODES = {"batch-float": capi.GMM_BATCH_FLOAT, "diagonal-maximum": capi.GMM_DIAG_MAX,
             "diagonal-sum": capi.GMM_DIAG_SUM, "batch-tensor": capi.GMM_BATCH_TENSOR,
             "batch-int": capi.GMM_BATCH_INT, "preselection-batch-float": capi.GMM_BATCH_PRESELECT,
             "preselection-batch-int": capi.GMM_BATCH_PRESELECT_INT, "SIMD-diagonal-maximum": capi.GMM_SIMD_DIAG_MAX}

    def __init__(self, mixture_set, mode="batch-float", mixture_weight_scale=1.0, gaussian_scale=1.0,
                 contraction=True, device=0):
        if isinstance(mixture_set, dict):
            mixture_set = MixtureSet.from_dict(mixture_set)
        self.mixture_set = mixture_set
        self.mode = self.MODES[mode] if isinstance(mode, str) else int(mode)
        self._h = C.c_void_p()
        cs = mixture_set.c_struct()
        capi.check(capi.lib().rb_gmm_create(C.byref(cs), self.mode, mixture_weight_scale, gaussian_scale,
                                            int(contraction), device, C.byref(self._h)))
        self.n_mixtures = mixture_set.n_mixtures
        self.dim = mixture_set.dim

    def close(self):
        if getattr(self, "_h", None) and capi is not None:  # capi is None while the interpreter shuts down
            capi.lib().rb_gmm_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def configure_preselection(self, clusters=256, select=32, iterations=5, backoff_score=40000.0):
        """density-clustering parameters of the preselection scorer (src/Mm/DensityClustering.cc:20-34)"""
        capi.check(capi.lib().rb_gmm_configure_preselection(self._h, int(clusters), int(select), int(iterations),
                                                            float(backoff_score)))

    def clustering(self):
        """(cluster index of every density in mixture order, cluster means [n_clusters, padded dim])"""
        n_dens = int(self.mixture_set.mix_offsets[-1])
        padded = (self.dim + 15) // 16 * 16 if self.mode == capi.GMM_BATCH_PRESELECT_INT else (self.dim + 7) // 8 * 8
        cluster_of, means, n = np.zeros(n_dens, np.uint32), np.zeros((256, padded), np.float32), C.c_int(0)
        capi.check(capi.lib().rb_gmm_get_clustering(self._h, capi.ptr(cluster_of), capi.ptr(means), C.byref(n)))
        return cluster_of, means[:n.value]

    def score(self, feats, want_density=False, out=None):
        """Host buffers (numpy, or pinned torch CPU tensors through their data_ptr)."""
        if isinstance(feats, np.ndarray):
            feats = np.ascontiguousarray(feats, np.float32)
        T = int(feats.shape[0])
        scores = out if out is not None else np.zeros((T, self.n_mixtures), np.float32)
        best = np.zeros((T, self.n_mixtures), np.uint32) if want_density else None
        capi.check(capi.lib().rb_gmm_score(self._h, capi.ptr(feats), T, capi.ptr(scores), capi.ptr(best)))
        return (scores, best) if want_density else scores

    def score_fanout_dev(self, d_feats, T, dsts, stream=None):
        """the same scores into every destination of `dsts` (device addresses, e.g. ScoreExchange.targets())"""
        arr = (C.c_void_p * len(dsts))(*[capi.ptr(d).value for d in dsts])
        capi.check(capi.lib().rb_gmm_score_fanout_dev(self._h, capi.ptr(d_feats), int(T), len(dsts), arr,
                                                      capi.ptr(stream)))

    def set_timing(self, on=True):
        capi.check(capi.lib().rb_gmm_set_timing(self._h, int(on)))

    def timing(self):
        """(split, screen, refine) ms of the last timed two-pass call, or None if the direct kernel served it"""
        ms = np.zeros(3, np.float32)
        rc = capi.lib().rb_gmm_get_timing(self._h, capi.ptr(ms))
        if rc == -5:
            return None
        capi.check(rc)
        return [float(v) for v in ms]

    def score_dev(self, d_feats, T, d_scores, d_best=None, stream=None):
        capi.check(capi.lib().rb_gmm_score_dev(self._h, capi.ptr(d_feats), int(T), capi.ptr(d_scores),
                                               capi.ptr(d_best), capi.ptr(stream)))


class ContextScorer:
    """Mm::FeatureScorer::ContextScorer: nEmissions() / score(e) for one frame; may outlive later
    getScorer() calls (delayed scoring)."""

    def __init__(self, parent, t):
        self._parent, self._t = parent, t

    def n_emissions(self):
        return self._parent.n_mixtures()

    def score(self, e):
        return float(self._parent._row(self._t)[e])

    def scores(self):
        """Dense row (what Speech::FeatureScorerNode materialises: out[i] = -score(i),
        src/Speech/FeatureScorerNode.cc:95-111 -- sign left to the caller)."""
        return self._parent._row(self._t)


class BatchFeatureScorer:
    """Whole-segment buffered feature scorer over a dense device-side score matrix.

    Protocol mirror: isBuffered()=True; addFeature() until bufferFilled(); getScorer(f) pushes f and
    returns the scorer of the oldest unscored frame; flush() drains; reset() starts a new segment.
    `buffer_size` is what bufferFilled() reports against; the reference default is 4
    (src/Mm/BatchFeatureScorer.cc:28-29), a whole-segment scorer uses a huge value so the recognizer
    buffers everything until the segment ends (template: src/Onnx/OnnxFeatureScorer.cc:117-186).
    """

    def __init__(self, scorer, buffer_size=1 << 30):
        self._scorer = scorer
        self._buffer_size = int(buffer_size)
        self.reset()

    def n_mixtures(self):
        return self._scorer.n_mixtures

    def dimension(self):
        return self._scorer.dim

    def is_buffered(self):
        return True

    def buffer_size(self):
        return self._buffer_size

    def reset(self):
        self._feats = []
        self._next = 0          # index of the oldest frame without a scorer
        self._scores = None     # dense matrix of the frames [0, _scored)
        self._scored = 0

    def finalize(self):
        pass

    def buffer_filled(self):
        return len(self._feats) - self._next >= self._buffer_size

    def buffer_empty(self):
        return self._next >= len(self._feats)

    def add_feature(self, f):
        f = np.asarray(f, np.float32)
        if f.shape != (self._scorer.dim,):
            raise capi.RasrB200Error(-1, "feature has dimension %s, mixture set expects %d" % (f.shape, self._scorer.dim))
        if self.buffer_filled():
            raise capi.RasrB200Error(-5, "addFeature on a filled buffer (require(!bufferFilled()))")
        self._feats.append(f)

    def get_scorer(self, f):
        """Push f, return the scorer of the oldest buffered frame."""
        f = np.asarray(f, np.float32)
        if f.shape != (self._scorer.dim,):
            raise capi.RasrB200Error(-1, "feature has dimension %s, mixture set expects %d" % (f.shape, self._scorer.dim))
        self._feats.append(f)
        return self._pop()

    def flush(self):
        if self.buffer_empty():
            raise capi.RasrB200Error(-5, "flush on an empty buffer (require(!bufferEmpty()))")
        return self._pop()

    def _pop(self):
        t = self._next
        self._next += 1
        return ContextScorer(self, t)

    def _row(self, t):
        if t >= self._scored:
            # score everything buffered so far in one dense launch
            block = np.stack(self._feats[self._scored:])
            s = self._scorer.score(block)
            self._scores = s if self._scores is None else np.concatenate([self._scores, s])
            self._scored = len(self._feats)
        return self._scores[t]
