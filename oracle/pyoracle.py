"""ctypes binding of the CPU oracle (oracle/liboracle.so, oracle/_ref/libref_fft*.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package rasr_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ACT = {"linear": 0, "sigmoid": 1, "relu": 2, "rectified": 2, "softmax": 3, "tanh": 4}
NN_F32, NN_F64ACC, NN_BF16 = 0, 1, 2


def build(ref=True):
    """(Re)build liboracle.so and, when the reference checkout is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    if ref and os.path.isdir("/root/reference/src/Math"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_ref"])


WINDOW_TYPES = {"hamming": 0, "rectangular": 1, "hanning": 2, "periodic-hanning": 3, "bartlett": 4, "blackman": 5, "kaiser": 6}


class FrontendCfg(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_double),
        ("window_length_s", C.c_double),
        ("window_shift_s", C.c_double),
        ("fft_max_input_s", C.c_double),
        ("filter_width", C.c_double),
        ("preemphasis_alpha", C.c_float),
        ("n_cepstra", C.c_int),
        ("derivatives", C.c_int),
        ("use_fma", C.c_int),
        ("window_type", C.c_int),
    ]


class FrontendGeometry(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("win_length", "win_shift", "fft_length", "n_bins", "n_filters", "n_weights", "feat_dim")]


class MixtureSetC(C.Structure):
    _fields_ = [
        ("dim", C.c_uint32),
        ("n_mixtures", C.c_uint32),
        ("n_densities", C.c_uint32),
        ("n_means", C.c_uint32),
        ("n_covariances", C.c_uint32),
        ("mix_offsets", C.POINTER(C.c_uint32)),
        ("mix_density", C.POINTER(C.c_uint32)),
        ("mix_log_weight", C.POINTER(C.c_double)),
        ("dens_mean", C.POINTER(C.c_uint32)),
        ("dens_cov", C.POINTER(C.c_uint32)),
        ("means", C.POINTER(C.c_float)),
        ("variances", C.POINTER(C.c_float)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _lib = C.CDLL(path)
        _lib.orc_frontend_nframes.restype = C.c_long
        _lib.orc_frontend_nframes.argtypes = [C.POINTER(FrontendCfg), C.c_long]
        _lib.orc_mfcc.restype = C.c_long
        _lib.orc_version.restype = C.c_char_p
    return _lib


def ref_fft(native=False):
    """The reference's own FFT object code (None when oracle/_ref has not been built)."""
    path = os.path.join(_HERE, "_ref", "libref_fft_native.so" if native else "libref_fft.so")
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def frontend_cfg(sample_rate=16000.0, window_length_s=0.025, window_shift_s=0.01, fft_max_input_s=0.025,
                 filter_width=268.258, alpha=1.0, n_cepstra=13, derivatives=True, use_fma=True, window_type="hamming"):
    return FrontendCfg(sample_rate, window_length_s, window_shift_s, fft_max_input_s, filter_width, alpha,
                       n_cepstra, int(derivatives), int(use_fma), WINDOW_TYPES[window_type])


def geometry(cfg):
    g = FrontendGeometry()
    rc = lib().orc_frontend_get_geometry(C.byref(cfg), C.byref(g))
    if rc:
        raise ValueError("invalid front-end configuration (%d)" % rc)
    return g


def nframes(cfg, n):
    return int(lib().orc_frontend_nframes(C.byref(cfg), n))


def tables(cfg):
    g = geometry(cfg)
    window = np.zeros(g.win_length, np.float32)
    start = np.zeros(g.n_filters, np.int32)
    end = np.zeros(g.n_filters, np.int32)
    w = np.zeros((g.n_filters, g.n_bins), np.float32)
    dct = np.zeros((cfg.n_cepstra, g.n_filters), np.float32)
    lib().orc_frontend_tables(C.byref(cfg), _p(window, C.c_float), _p(start, C.c_int), _p(end, C.c_int),
                              _p(w, C.c_float), _p(dct, C.c_float))
    return dict(window=window, fb_start=start, fb_end=end, fb_weights=w, dct=dct)


def mfcc(cfg, samples, chunk=0, stages=False):
    samples = np.ascontiguousarray(samples, np.float32)
    g = geometry(cfg)
    T = nframes(cfg, samples.size)
    feats = np.zeros((T, g.feat_dim), np.float32)
    ts = np.zeros(T, np.float64)
    te = np.zeros(T, np.float64)
    spec = amp = fb = cep = None
    if stages:
        spec = np.zeros((T, g.fft_length + 2), np.float32)
        amp = np.zeros((T, g.n_bins), np.float32)
        fb = np.zeros((T, g.n_filters), np.float32)
        cep = np.zeros((T, cfg.n_cepstra), np.float32)
    n = lib().orc_mfcc(C.byref(cfg), _p(samples, C.c_float), C.c_long(samples.size), C.c_long(chunk),
                       _p(feats, C.c_float), _p(ts, C.c_double), _p(te, C.c_double), _p(spec, C.c_float),
                       _p(amp, C.c_float), _p(fb, C.c_float), _p(cep, C.c_float))
    if n != T:
        raise RuntimeError("oracle frame count %d != closed form %d" % (n, T))
    out = dict(feats=feats, t_start=ts, t_end=te)
    if stages:
        out.update(spectrum=spec, amplitude=amp, fbank=fb, cepstra=cep)
    return out


class DcCfg(C.Structure):
    _fields_ = [("min_dc_length_s", C.c_double), ("max_dc_increment", C.c_float),
                ("min_non_dc_segment_length_s", C.c_double), ("maximal_output_size", C.c_int)]


def dc_cfg(min_dc_length_s=0.0125, max_dc_increment=0.9, min_non_dc_segment_length_s=0.026, maximal_output_size=4096):
    """parameters of signal-dc-detection as samples.flow:34-35 sets them (node defaults: .0125 / 0.9 / .02 / 4096)"""
    return DcCfg(min_dc_length_s, max_dc_increment, min_non_dc_segment_length_s, maximal_output_size)


def mfcc_dc(cfg, dc, samples, chunk=0):
    """samples -> signal-dc-detection -> MFCC chain.  Returns feats, t_start, t_end and the kept sample runs
    (begin, end, start time)."""
    samples = np.ascontiguousarray(samples, np.float32)
    g = geometry(cfg)
    cap = samples.size // max(1, g.win_shift) + samples.size // 64 + 8
    feats = np.zeros((cap, g.feat_dim), np.float32)
    ts = np.zeros(cap, np.float64)
    te = np.zeros(cap, np.float64)
    rcap = samples.size // 8 + 2
    rb, re = np.zeros(rcap, np.int64), np.zeros(rcap, np.int64)
    rs = np.zeros(rcap, np.float64)
    nr = C.c_long(0)
    L = lib()
    L.orc_mfcc_dc.restype = C.c_long
    T = L.orc_mfcc_dc(C.byref(cfg), C.byref(dc), _p(samples, C.c_float), C.c_long(samples.size), C.c_long(chunk),
                      _p(feats, C.c_float), _p(ts, C.c_double), _p(te, C.c_double), C.c_long(cap), _p(rb, C.c_long),
                      _p(re, C.c_long), _p(rs, C.c_double), C.c_long(rcap), C.byref(nr))
    if T < 0:
        raise RuntimeError("orc_mfcc_dc failed: %d" % T)
    n = nr.value
    return dict(feats=feats[:T].copy(), t_start=ts[:T].copy(), t_end=te[:T].copy(), run_begin=rb[:n].copy(),
                run_end=re[:n].copy(), run_start=rs[:n].copy())


def fft_real_packed(v):
    v = np.ascontiguousarray(v, np.float32).copy()
    lib().orc_fft_real_packed(_p(v, C.c_float), C.c_int(v.size))
    return v


class MixtureSet:
    """Holds the numpy arrays alive behind an orc_mixture_set."""

    def __init__(self, dim, mix_offsets, mix_density, mix_log_weight, dens_mean, dens_cov, means, variances):
        self.a = dict(
            mix_offsets=np.ascontiguousarray(mix_offsets, np.uint32),
            mix_density=np.ascontiguousarray(mix_density, np.uint32),
            mix_log_weight=np.ascontiguousarray(mix_log_weight, np.float64),
            dens_mean=np.ascontiguousarray(dens_mean, np.uint32),
            dens_cov=np.ascontiguousarray(dens_cov, np.uint32),
            means=np.ascontiguousarray(means, np.float32).reshape(-1, dim),
            variances=np.ascontiguousarray(variances, np.float32).reshape(-1, dim),
        )
        a = self.a
        self.c = MixtureSetC(dim, a["mix_offsets"].size - 1, a["dens_mean"].size, a["means"].shape[0],
                             a["variances"].shape[0], _p(a["mix_offsets"], C.c_uint32),
                             _p(a["mix_density"], C.c_uint32), _p(a["mix_log_weight"], C.c_double),
                             _p(a["dens_mean"], C.c_uint32), _p(a["dens_cov"], C.c_uint32),
                             _p(a["means"], C.c_float), _p(a["variances"], C.c_float))
        self.dim = dim
        self.n_mixtures = a["mix_offsets"].size - 1


def gmm_batch_float(ms, feats, use_fma=True, threads=1):
    feats = np.ascontiguousarray(feats, np.float32)
    T = feats.shape[0]
    scores = np.zeros((T, ms.n_mixtures), np.float32)
    if threads > 1:
        rc = lib().orc_gmm_batch_float_mt(C.byref(ms.c), _p(feats, C.c_float), C.c_long(T), _p(scores, C.c_float),
                                          int(use_fma), int(threads))
    else:
        rc = lib().orc_gmm_batch_float(C.byref(ms.c), _p(feats, C.c_float), C.c_long(T), _p(scores, C.c_float),
                                       int(use_fma))
    if rc:
        raise RuntimeError("orc_gmm_batch_float failed: %d" % rc)
    return scores


def gmm_preselect_float(ms, feats, use_fma=True, clusters=256, select=32, iterations=5, backoff=40000.0):
    """(scores, cluster index of every density, cluster means [n_clusters, padded])"""
    feats = np.ascontiguousarray(feats, np.float32)
    T = feats.shape[0]
    scores = np.zeros((T, ms.n_mixtures), np.float32)
    n_dens = int(ms.c.mix_offsets[ms.n_mixtures])
    padded = (ms.dim + 7) // 8 * 8
    cluster_of = np.zeros(n_dens, np.uint32)
    means = np.zeros((clusters, padded), np.float32)
    n = C.c_int(0)
    rc = lib().orc_gmm_preselect_float(C.byref(ms.c), _p(feats, C.c_float), C.c_long(T), _p(scores, C.c_float),
                                       int(use_fma), int(clusters), int(select), int(iterations), C.c_float(backoff),
                                       _p(cluster_of, C.c_uint32), _p(means, C.c_float), C.byref(n))
    if rc:
        raise RuntimeError("orc_gmm_preselect_float failed: %d" % rc)
    return scores, cluster_of, means[:n.value]


def gmm_preselect_int(ms, feats, clusters=256, select=32, iterations=5, restated_sort=False):
    feats = np.ascontiguousarray(feats, np.float32)
    T = feats.shape[0]
    scores = np.zeros((T, ms.n_mixtures), np.float32)
    cluster_of = np.zeros(int(ms.c.mix_offsets[ms.n_mixtures]), np.uint32)
    rc = lib().orc_gmm_preselect_int(C.byref(ms.c), _p(feats, C.c_float), C.c_long(T), _p(scores, C.c_float),
                                     int(clusters), int(select), int(iterations), _p(cluster_of, C.c_uint32),
                                     int(restated_sort))
    if rc:
        raise RuntimeError("orc_gmm_preselect_int failed: %d" % rc)
    return scores, cluster_of


def sort_pairs(keys, restated):
    keys = np.ascontiguousarray(keys, np.int32)
    perm = np.zeros(keys.size, np.int32)
    lib().orc_sort_pairs(_p(keys, C.c_int32), int(keys.size), _p(perm, C.c_int32), int(restated))
    return perm


def sort_heap_calls():
    lib().orc_sort_heap_calls.restype = C.c_long
    return int(lib().orc_sort_heap_calls())


def sort_killer(n):
    keys = np.zeros(n, np.int32)
    lib().orc_sort_killer(int(n), _p(keys, C.c_int32))
    return keys


def gmm_batch_int(ms, feats, threads=1):
    """Mm::BatchIntFeatureScorer ("batch-diagonal-maximum-int"): dense scores [T x nMix]."""
    feats = np.ascontiguousarray(feats, np.float32)
    T = feats.shape[0]
    scores = np.zeros((T, ms.n_mixtures), np.float32)
    rc = lib().orc_gmm_batch_int(C.byref(ms.c), _p(feats, C.c_float), C.c_long(T), _p(scores, C.c_float), int(threads))
    if rc:
        raise RuntimeError("orc_gmm_batch_int failed: %d" % rc)
    return scores


def gmm_simd_diag_max(ms, feats, threads=1, want_best=False):
    """Mm::SimdGaussDiagonalMaximumFeatureScorer over all mixtures; optionally the best density-in-mixture indices"""
    feats = np.ascontiguousarray(feats, np.float32)
    T = feats.shape[0]
    scores = np.zeros((T, ms.n_mixtures), np.float32)
    best = np.zeros((T, ms.n_mixtures), np.uint32) if want_best else None
    rc = lib().orc_gmm_simd_diag_max(C.byref(ms.c), _p(feats, C.c_float), C.c_long(T), _p(scores, C.c_float),
                                     _p(best, C.c_uint32), int(threads))
    if rc:
        raise RuntimeError("orc_gmm_simd_diag_max failed: %d" % rc)
    return (scores, best) if want_best else scores


def gmm_batch_int_model(ms):
    """The quantised model: dict(means u8 [nDens x padded], consts s32 [nDens], variance f32 [padded], scale)."""
    n_dens = int(ms.a["mix_offsets"][-1])
    padded = (ms.dim + 15) // 16 * 16
    means = np.zeros((n_dens, padded), np.uint8)
    consts = np.zeros(n_dens, np.int32)
    variance = np.zeros(padded, np.float32)
    scale = C.c_float()
    pad = C.c_int()
    rc = lib().orc_gmm_batch_int_model(C.byref(ms.c), _p(means, C.c_uint8), _p(consts, C.c_int32),
                                       _p(variance, C.c_float), C.byref(scale), C.byref(pad))
    if rc:
        raise RuntimeError("orc_gmm_batch_int_model failed: %d" % rc)
    return dict(means=means, consts=consts, variance=variance, scale=scale.value, padded=pad.value)


def _gmm_diag(fn, ms, feats, mixture_weight_scale, gaussian_scale, use_fma):
    feats = np.ascontiguousarray(feats, np.float32)
    T = feats.shape[0]
    scores = np.zeros((T, ms.n_mixtures), np.float32)
    best = np.zeros((T, ms.n_mixtures), np.uint32)
    rc = fn(C.byref(ms.c), C.c_float(mixture_weight_scale), C.c_float(gaussian_scale), _p(feats, C.c_float),
            C.c_long(T), _p(scores, C.c_float), _p(best, C.c_uint32), int(use_fma))
    if rc:
        raise RuntimeError("oracle gmm scorer failed: %d" % rc)
    return scores, best


def gmm_diag_max(ms, feats, mixture_weight_scale=1.0, gaussian_scale=1.0, use_fma=True):
    return _gmm_diag(lib().orc_gmm_diag_max, ms, feats, mixture_weight_scale, gaussian_scale, use_fma)


def gmm_diag_sum(ms, feats, mixture_weight_scale=1.0, gaussian_scale=1.0, use_fma=True):
    return _gmm_diag(lib().orc_gmm_diag_sum, ms, feats, mixture_weight_scale, gaussian_scale, use_fma)


def _offsets(frame_offsets, T):
    fo = np.ascontiguousarray(frame_offsets if frame_offsets is not None else [0, T], np.int64)
    return fo, fo.size - 1


def normalize(feats, frame_offsets=None, kind="mean", length=-1, right=-1, use_fma=True):
    """signal-normalization over segments given by frame_offsets (default: one segment); length/right < 0 = infinite."""
    feats = np.ascontiguousarray(feats, np.float32)
    fo, n = _offsets(frame_offsets, feats.shape[0])
    out = np.zeros_like(feats)
    rc = lib().orc_normalize({"mean": 1, "mean-and-variance": 2}[kind], C.c_long(length), C.c_long(right),
                             _p(feats, C.c_float), _p(fo, C.c_long), n, feats.shape[1], _p(out, C.c_float), int(use_fma))
    if rc:
        raise RuntimeError("orc_normalize failed: %d" % rc)
    return out


def splice(feats, length, right, frame_offsets=None):
    feats = np.ascontiguousarray(feats, np.float32)
    fo, n = _offsets(frame_offsets, feats.shape[0])
    out = np.zeros((feats.shape[0], length * feats.shape[1]), np.float32)
    rc = lib().orc_splice(int(length), int(right), _p(feats, C.c_float), _p(fo, C.c_long), n, feats.shape[1],
                          _p(out, C.c_float))
    if rc:
        raise RuntimeError("orc_splice failed: %d" % rc)
    return out


def matmul(M, x, use_fma=True):
    M = np.ascontiguousarray(M, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    y = np.zeros((x.shape[0], M.shape[0]), np.float32)
    lib().orc_matmul(_p(M, C.c_float), M.shape[0], M.shape[1], _p(x, C.c_float), C.c_long(x.shape[0]),
                     _p(y, C.c_float), int(use_fma))
    return y


class LexiconC(C.Structure):
    _fields_ = [("n_words", C.c_uint32), ("word_offsets", C.POINTER(C.c_uint32)),
                ("state_emission", C.POINTER(C.c_uint32)), ("state_tdp_model", C.POINTER(C.c_uint32)),
                ("n_models", C.c_uint32), ("tdp", C.POINTER(C.c_float)), ("entry_model", C.c_uint32),
                ("unigram", C.POINTER(C.c_float)), ("word_regular", C.POINTER(C.c_uint8)), ("single_word", C.c_int32)]


def linear_search(lex, scores):
    """Search::LinearSearch over one segment; lex: dict(word_offsets, state_emission, state_tdp_model, tdp, entry_model,
    unigram).  Returns dict(words, times, am, lm) in chronological order."""
    a = dict(word_offsets=np.ascontiguousarray(lex["word_offsets"], np.uint32),
             state_emission=np.ascontiguousarray(lex["state_emission"], np.uint32),
             state_tdp_model=np.ascontiguousarray(lex["state_tdp_model"], np.uint32),
             tdp=np.ascontiguousarray(lex["tdp"], np.float32).reshape(-1, 4),
             unigram=np.ascontiguousarray(lex["unigram"], np.float32))
    reg = lex.get("word_regular")
    if reg is not None:
        reg = np.ascontiguousarray(reg, np.uint8)
    c = LexiconC(a["word_offsets"].size - 1, _p(a["word_offsets"], C.c_uint32), _p(a["state_emission"], C.c_uint32),
                 _p(a["state_tdp_model"], C.c_uint32), a["tdp"].shape[0], _p(a["tdp"], C.c_float),
                 int(lex["entry_model"]), _p(a["unigram"], C.c_float),
                 _p(reg, C.c_uint8) if reg is not None else None, int(bool(lex.get("single_word", False))))
    scores = np.ascontiguousarray(scores, np.float32)
    T = scores.shape[0]
    words, times = np.zeros(max(T, 1), np.uint32), np.zeros(max(T, 1), np.int32)
    am, lm = np.zeros(max(T, 1), np.float32), np.zeros(max(T, 1), np.float32)
    fn = lib().orc_linear_search
    fn.restype = C.c_long
    n = fn(C.byref(c), _p(scores, C.c_float), C.c_long(T), int(scores.shape[1]), _p(words, C.c_uint32),
           _p(times, C.c_int32), _p(am, C.c_float), _p(lm, C.c_float))
    if n < 0:
        raise RuntimeError("orc_linear_search failed: %d" % n)
    return dict(words=words[:n], times=times[:n], am=am[:n], lm=lm[:n])


def _nn_args(dims, acts, weights, biases, dtype, ctype):
    n = len(weights)
    dims_a = np.asarray(dims, np.int32)
    act_a = np.asarray([ACT[a] if isinstance(a, str) else a for a in acts], np.int32)
    ws = [np.ascontiguousarray(w, dtype) for w in weights]
    bs = [np.ascontiguousarray(b, dtype) for b in biases]
    wp = (C.POINTER(ctype) * n)(*[_p(w, ctype) for w in ws])
    bp = (C.POINTER(ctype) * n)(*[_p(b, ctype) for b in bs])
    return n, dims_a, act_a, ws, bs, wp, bp


def nn_forward(dims, acts, weights, biases, x, mode=NN_F32):
    """weights[l]: array of shape (out, in) == the reference's in x out column-major storage."""
    n, dims_a, act_a, ws, bs, wp, bp = _nn_args(dims, acts, weights, biases, np.float32, C.c_float)
    x = np.ascontiguousarray(x, np.float32)
    T = x.shape[0]
    out = np.zeros((T, dims[-1]), np.float32)
    rc = lib().orc_nn_forward(n, _p(dims_a, C.c_int), _p(act_a, C.c_int), wp, bp, _p(x, C.c_float), C.c_long(T),
                              _p(out, C.c_float), int(mode))
    if rc:
        raise RuntimeError("orc_nn_forward failed: %d" % rc)
    return out


def nn_forward_f64(dims, acts, weights, biases, x):
    n, dims_a, act_a, ws, bs, wp, bp = _nn_args(dims, acts, weights, biases, np.float64, C.c_double)
    x = np.ascontiguousarray(x, np.float64)
    T = x.shape[0]
    out = np.zeros((T, dims[-1]), np.float64)
    rc = lib().orc_nn_forward_f64(n, _p(dims_a, C.c_int), _p(act_a, C.c_int), wp, bp, _p(x, C.c_double),
                                  C.c_long(T), _p(out, C.c_double))
    if rc:
        raise RuntimeError("orc_nn_forward_f64 failed: %d" % rc)
    return out


def nn_scores(dims, acts, weights, biases, log_prior, prior_scale, x, mode=NN_F32):
    n, dims_a, act_a, ws, bs, wp, bp = _nn_args(dims, acts, weights, biases, np.float32, C.c_float)
    x = np.ascontiguousarray(x, np.float32)
    lp = np.ascontiguousarray(log_prior, np.float32) if log_prior is not None else None
    T = x.shape[0]
    out = np.zeros((T, dims[-1]), np.float32)
    rc = lib().orc_nn_scores(n, _p(dims_a, C.c_int), _p(act_a, C.c_int), wp, bp, _p(lp, C.c_float),
                             C.c_float(prior_scale), _p(x, C.c_float), C.c_long(T), _p(out, C.c_float), int(mode))
    if rc:
        raise RuntimeError("orc_nn_scores failed: %d" % rc)
    return out
