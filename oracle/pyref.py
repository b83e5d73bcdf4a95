"""ctypes binding of oracle/_ref/librasr_ref*.so: the REFERENCE's own object code (Core, Flow, Math, Signal, Mm
translation units compiled from /root/reference by oracle/refbuild/Makefile) behind the C entry points of
oracle/refbuild/ref_host.cc.

TEST INFRASTRUCTURE ONLY: importable from tests/, tests/golden/make_golden.py and bench.py's cpu_baseline /
--impl reference legs.  The product package rasr_b200 never imports this module.

Two variants, both loadable in one process (each library binds its own symbols, -Bsymbolic / -fno-gnu-unique):
`native=False` = every operation rounded separately, `native=True` = gcc's default FMA contraction, what the
reference's own -march=native build does.  The adapters (libb200_adapters.so) resolve against the native variant.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REFBUILD = os.path.join(_HERE, "refbuild")
REFERENCE = "/root/reference"
FLOW_SHARE = os.path.join(REFERENCE, "src", "Tools", "FeatureExtraction", "share")

_libs = {}


def path(native=False):
    return os.path.join(_HERE, "_ref", "librasr_ref_native.so" if native else "librasr_ref.so")


def available(native=False):
    return os.path.exists(path(native))


def build():
    """Compile the reference's sources from where they lie (needs /root/reference; a few minutes the first time)."""
    if not os.path.isdir(os.path.join(REFERENCE, "src", "Mm")):
        raise RuntimeError("the reference checkout is not present on this host")
    subprocess.check_call(["make", "-s", "-j%d" % (os.cpu_count() or 4), "-C", _REFBUILD, "all"])


def lib(native=False, log_file=None):
    native = bool(native)
    if native in _libs:
        return _libs[native]
    if not available(native):
        build()
    # the native variant is loaded globally: libb200_adapters.so resolves the RASR symbols against it
    L = C.CDLL(path(native), mode=C.RTLD_GLOBAL if native else C.RTLD_LOCAL)
    L.ref_last_error.restype = C.c_char_p
    L.ref_flow_create.restype = C.c_void_p
    L.ref_flow_run.restype = C.c_long
    L.ref_mm_create.restype = C.c_void_p
    L.ref_init(log_file.encode() if log_file else None)
    # the reference's own .flow files are found through this path when a network names them as a filter; on a host
    # without the reference checkout (the GPU box) only self-contained networks can be built
    L.ref_config_set(b"*.network-file-path", FLOW_SHARE.encode())
    _libs[native] = L
    return L


_adapters = None


def load_adapters():
    """Load oracle/_ref/libb200_adapters.so (adapters/*.cc compiled against the reference's headers) into the
    reference host and run INIT_MODULE(B200): afterwards the reference's Flow registry knows the filter "b200-mfcc" and
    its Mm factory the "b200-*" feature scorers, both forwarding to librasr_b200.so."""
    global _adapters
    L = lib(native=True)
    if _adapters is None:
        p = os.path.join(_HERE, "_ref", "libb200_adapters.so")
        if not os.path.exists(p):
            build()
        _adapters = C.CDLL(p, mode=C.RTLD_GLOBAL)
        _adapters.b200_adapters_register()
    return L


def config_set(name, value, native=False):
    """A resource of the reference's global Core::Configuration, e.g. ("*.density-clustering.clusters", "64")."""
    lib(native).ref_config_set(name.encode(), str(value).encode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# the values src/Tools/FeatureExtraction/share/mfcc.flow:8-34 and samples.flow:34-35 fix, as the text they carry there
CHAIN_DEFAULTS = {"block-size": 4096, "alpha": "1.00", "window-type": "hamming", "shift": ".01", "length": "0.025",
                  "maximum-input-size": "0.025", "filter-width": "268.258", "nr-cepstrum-coefficients": 13}
DC_DEFAULTS = {"min-dc-length": ".0125", "max-dc-increment": "0.9", "min-non-dc-segment-length": ".026",
               "maximal-output-size": 4096}


def chain_parameters(dc=False, **overrides):
    """Parameters of mfcc_chain_plain.flow / mfcc_chain_dc.flow / b200_mfcc.flow; keyword names with '_' for '-'."""
    p = dict(CHAIN_DEFAULTS)
    if dc:
        p.update(DC_DEFAULTS)
    p.update({k.replace("_", "-"): v for k, v in overrides.items()})
    return p


class FlowNetwork:
    """A Flow::Network built by the reference's NetworkParser from `flow_file` (relative names: oracle/refbuild/flows)."""

    def __init__(self, flow_file="mfcc_derivatives.flow", parameters=None, native=False, selection="flow"):
        self._L = lib(native)
        if not os.path.isabs(flow_file):
            flow_file = os.path.join(_REFBUILD, "flows", flow_file)
        self._h = self._L.ref_flow_create(flow_file.encode(), selection.encode())
        if not self._h:
            raise RuntimeError("ref_flow_create: %s" % self._L.ref_last_error().decode())
        for k, v in (parameters or {}).items():
            if self._L.ref_flow_set_parameter(C.c_void_p(self._h), k.encode(), str(v).encode()) != 0:
                raise RuntimeError("network has no parameter %s" % k)

    def run(self, samples, port="features", sample_rate=16000.0, start_time=0.0, capacity=None, width=None):
        """One segment through the network; returns dict(feats [T x dim], t_start, t_end, sizes).  The segment is pulled
        twice (count, then fill) unless `width` (the packet size) is given: networks with side effects -- a cache node
        that writes what passes through -- must be run once."""
        samples = np.ascontiguousarray(samples, np.float32)
        cap = int(capacity if capacity is not None else samples.size // 16 + 64)
        dim = C.c_int(0)
        if width is not None:
            feats = np.zeros((cap, int(width)), np.float32)
            ts, te, sizes = np.zeros(cap, np.float64), np.zeros(cap, np.float64), np.zeros(cap, np.int32)
            n = self._L.ref_flow_run(C.c_void_p(self._h), port.encode(), _p(samples), C.c_long(samples.size),
                                     C.c_double(sample_rate), C.c_double(start_time), _p(feats), C.c_long(cap),
                                     int(width), _p(ts), _p(te), C.byref(dim), _p(sizes))
            if n < 0 or n > cap or dim.value > int(width):
                raise RuntimeError("ref_flow_run: %s" % (self._L.ref_last_error().decode() or "capacity / width too small"))
            return dict(feats=feats[:n], t_start=ts[:n], t_end=te[:n], sizes=sizes[:n])
        # first pass counts packets and finds their width
        n = self._L.ref_flow_run(C.c_void_p(self._h), port.encode(), _p(samples), C.c_long(samples.size),
                                 C.c_double(sample_rate), C.c_double(start_time), None, C.c_long(0), 0, None, None,
                                 C.byref(dim), None)
        if n < 0:
            raise RuntimeError("ref_flow_run: %s" % self._L.ref_last_error().decode())
        cap, d = int(n), max(1, dim.value)
        feats = np.zeros((cap, d), np.float32)
        ts, te = np.zeros(cap, np.float64), np.zeros(cap, np.float64)
        sizes = np.zeros(cap, np.int32)
        n2 = self._L.ref_flow_run(C.c_void_p(self._h), port.encode(), _p(samples), C.c_long(samples.size),
                                  C.c_double(sample_rate), C.c_double(start_time), _p(feats), C.c_long(cap), d,
                                  _p(ts), _p(te), C.byref(dim), _p(sizes))
        if n2 != n:
            raise RuntimeError("the network produced %d packets on the second pass, %d on the first" % (n2, n))
        return dict(feats=feats, t_start=ts, t_end=te, sizes=sizes)

    def set_parameter(self, name, value):
        if self._L.ref_flow_set_parameter(C.c_void_p(self._h), name.encode(), str(value).encode()) != 0:
            raise RuntimeError("network has no parameter %s" % name)

    def attribute(self, port, name):
        buf = C.create_string_buffer(256)
        if self._L.ref_flow_get_attribute(C.c_void_p(self._h), port.encode(), name.encode(), buf, 256) < 0:
            raise RuntimeError("no output port %s" % port)
        return buf.value.decode()

    def close(self):
        if getattr(self, "_h", None):
            self._L.ref_flow_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        self.close()


class FeatureScorer:
    """An Mm::FeatureScorer created by the reference's own factory (`scorer_type` = the value of feature-scorer-type,
    src/Mm/Module.cc:83-105; "diagonal-sum" instantiates the unregistered GaussDiagonalSumFeatureScorer).  `ms` is an
    oracle.pyoracle.MixtureSet (same C layout).  `config`: resources below the scorer's selection, e.g.
    {"density-clustering.clusters": 64, "buffer-size": 4}."""

    def __init__(self, ms, scorer_type, config=None, native=False, selection=None, library=None):
        # library: another self-contained copy of the reference's object code (search_lib()) instead of lib(native)
        self._L = library or lib(native)
        self._L.ref_mm_create.restype = C.c_void_p
        FeatureScorer._count = getattr(FeatureScorer, "_count", 0) + 1
        sel = selection or "feature-scorer-%d" % FeatureScorer._count
        for k, v in (config or {}).items():
            self._L.ref_config_set(("*.%s.%s" % (sel, k)).encode(), str(v).encode())
        self._ms = ms
        self._h = self._L.ref_mm_create(C.byref(ms.c), scorer_type.encode(), sel.encode())
        if not self._h:
            raise RuntimeError("ref_mm_create: %s" % self._L.ref_last_error().decode())
        self.n_mixtures = int(self._L.ref_mm_n_mixtures(C.c_void_p(self._h)))

    def score(self, feats, want_best=False):
        feats = np.ascontiguousarray(feats, np.float32)
        T = feats.shape[0]
        scores = np.zeros((T, self.n_mixtures), np.float32)
        best = np.zeros((T, self.n_mixtures), np.uint32) if want_best else None
        rc = self._L.ref_mm_score(C.c_void_p(self._h), _p(feats), C.c_long(T), _p(scores), _p(best))
        if rc:
            raise RuntimeError("ref_mm_score failed (%d)" % rc)
        return (scores, best) if want_best else scores

    def close(self):
        if getattr(self, "_h", None):
            self._L.ref_mm_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        self.close()


# ---------------------------------------------------------------------------------------------------------------
# the reference's own readers / writers of its model files (oracle/refbuild/ref_io.cc)
def _io(native=False):
    L = lib(native)
    L.ref_mm_load.restype = C.c_void_p
    return L


def write_mixture_text(ms, path, precision=6, native=False):
    """Mm::Module_::writeMixtureSet (text; `ms` is an oracle.pyoracle.MixtureSet)"""
    if _io(native).ref_mm_write_text(C.byref(ms.c), str(path).encode(), int(precision)) != 0:
        raise RuntimeError("the reference could not write %s" % path)


def write_mixture_estimator(ms, feats, mix_of_frame, path, native=False):
    """Viterbi accumulation of the frames on `ms` (frame t -> mixture mix_of_frame[t], density chosen by the reference's
    diagonal-maximum scorer), accumulators written by Mm::Module_::writeMixtureSetEstimator"""
    feats = np.ascontiguousarray(feats, np.float32)
    mof = np.ascontiguousarray(mix_of_frame, np.uint32)
    if _io(native).ref_mm_write_estimator(C.byref(ms.c), _p(feats), C.c_long(feats.shape[0]), _p(mof),
                                          str(path).encode()) != 0:
        raise RuntimeError("the reference could not write %s" % path)


def read_mixture_file(path, native=False):
    """Mm::Module_::readMixtureSet -> dict in the layout of rasr_b200.io.read_mixture_set"""
    L = _io(native)
    h = L.ref_mm_load(str(path).encode(), None)
    if not h:
        raise RuntimeError("the reference could not read %s" % path)
    h = C.c_void_p(h)
    sizes = np.zeros(6, np.uint32)
    L.ref_mm_sizes(h, _p(sizes))
    dim, n_mix, n_dns, n_mean, n_cov, n_ent = (int(v) for v in sizes)
    out = dict(dim=dim, mix_offsets=np.zeros(n_mix + 1, np.uint32), mix_density=np.zeros(n_ent, np.uint32),
               mix_log_weight=np.zeros(n_ent, np.float64), dens_mean=np.zeros(n_dns, np.uint32),
               dens_cov=np.zeros(n_dns, np.uint32), means=np.zeros((n_mean, dim), np.float32),
               variances=np.zeros((n_cov, dim), np.float32))
    L.ref_mm_dump(h, _p(out["mix_offsets"]), _p(out["mix_density"]), _p(out["mix_log_weight"]), _p(out["dens_mean"]),
                  _p(out["dens_cov"]), _p(out["means"]), _p(out["variances"]))
    L.ref_mm_unload(h)
    return out


def write_matrix(filename, m, native=False):
    m = np.ascontiguousarray(m, np.float32)
    if _io(native).ref_math_write_matrix(str(filename).encode(), m.shape[0], m.shape[1], _p(m)) != 0:
        raise RuntimeError("the reference could not write %s" % filename)


def read_matrix(filename, native=False):
    L, r, c = _io(native), C.c_int(0), C.c_int(0)
    if L.ref_math_read_matrix(str(filename).encode(), C.byref(r), C.byref(c), None, C.c_long(0)) != 0:
        raise RuntimeError("the reference could not read %s" % filename)
    out = np.zeros((r.value, c.value), np.float32)
    L.ref_math_read_matrix(str(filename).encode(), C.byref(r), C.byref(c), _p(out), C.c_long(out.size))
    return out


def write_vector(filename, v, native=False):
    v = np.ascontiguousarray(v, np.float32)
    if _io(native).ref_math_write_vector(str(filename).encode(), v.size, _p(v)) != 0:
        raise RuntimeError("the reference could not write %s" % filename)


def read_vector(filename, native=False):
    L, n = _io(native), C.c_int(0)
    if L.ref_math_read_vector(str(filename).encode(), C.byref(n), None, C.c_long(0)) != 0:
        raise RuntimeError("the reference could not read %s" % filename)
    out = np.zeros(n.value, np.float32)
    L.ref_math_read_vector(str(filename).encode(), C.byref(n), _p(out), C.c_long(out.size))
    return out


# ---------------------------------------------------------------------------------------------------------
# Search::LinearSearch (oracle/_ref/librasr_ref_search.so, oracle/refbuild/ref_search.cc)

_search = None
TDP_MODELS = ("entry-m1", "entry-m2", "silence", "state-0", "state-1")  # src/Am/TransitionModel.hh:72-78
TDP_TYPES = ("loop", "forward", "skip", "exit")                         # src/Am/TransitionModel.hh:32-37


def search_path():
    return os.path.join(_HERE, "_ref", "librasr_ref_search.so")


def search_lib(global_scope=False):
    """global_scope: needed before load_search_adapter() -- the adapter library resolves its RASR symbols (and the
    inline statics of the reference's headers) against this library; do that in a process of its own, never next
    to load_adapters()."""
    global _search
    if _search is None:
        if not os.path.exists(search_path()):
            build()
        # self-contained (its own copy of the strict objects and its own application object / configuration)
        S = C.CDLL(search_path(), mode=C.RTLD_GLOBAL if global_scope else C.RTLD_LOCAL)
        S.ref_last_error.restype = C.c_char_p
        S.ref_init(None)
        S.ref_search_create.restype = C.c_void_p
        for f in ("ref_search_order", "ref_search_states", "ref_search_run", "ref_search_items"):
            getattr(S, f).restype = C.c_long
        _search = S
    return _search


_search_adapter = None


def load_search_adapter():
    """oracle/_ref/libb200_search_adapter.so: adapters/B200LinearSearch.cc compiled against the reference's headers;
    afterwards LinearSearch(..., adapter=True) drives it through the Search::SearchAlgorithm interface."""
    global _search_adapter
    if _search_adapter is None:
        if _search is not None:
            raise RuntimeError("load_search_adapter() must come before the first use of the search library")
        S = search_lib(global_scope=True)
        p = os.path.join(_HERE, "_ref", "libb200_search_adapter.so")
        if not os.path.exists(p):
            build()
        _search_adapter = C.CDLL(p, mode=C.RTLD_GLOBAL)
        _search_adapter.b200_search_adapter_register()
        # the other adapters (feature scorers, Flow nodes) into the same host: INIT_MODULE(B200) registers them with the
        # Mm / Flow factories of the search library, which carries the same reference objects as librasr_ref.so
        a = C.CDLL(os.path.join(_HERE, "_ref", "libb200_adapters.so"), mode=C.RTLD_GLOBAL)
        a.b200_adapters_register()
        _search_adapter._others = a
    return _search_adapter


def write_lexicon(path, n_phonemes, words, silence=True, silence_first=False, irregular=()):
    """A Bliss lexicon file: context-independent phonemes p0..p<n-1> (+ "si"), one lemma "w<k>" per entry of
    `words` (each a list of phoneme numbers, or a list of such lists = several pronunciations), optionally the
    special lemma "silence" (no syntactic token, empty evaluation sequence: an irregular word).  Lemmata whose number
    is in `irregular` keep their syntactic token (LM score) but get an empty evaluation sequence (noise words)."""
    sil = ('  <lemma special="silence"><orth>[SILENCE]</orth><phon>si</phon><synt/><eval/></lemma>\n'
           if silence else "")
    with open(path, "w") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n<lexicon>\n  <phoneme-inventory>\n')
        for k in range(n_phonemes):
            f.write("    <phoneme><symbol>p%d</symbol><variation>none</variation></phoneme>\n" % k)
        # always in the inventory: LinearSearch verifies that the acoustic model knows a silence phoneme
        f.write("    <phoneme><symbol>si</symbol><variation>none</variation></phoneme>\n")
        f.write("  </phoneme-inventory>\n")
        if silence_first:
            f.write(sil)
        for k, prons in enumerate(words):
            if len(prons) and not isinstance(prons[0], (list, tuple)):
                prons = [prons]
            f.write("  <lemma><orth>w%d</orth>" % k)
            for p in prons:
                f.write("<phon>%s</phon>" % " ".join("p%d" % q for q in p))
            f.write("<eval/></lemma>\n" if k in irregular else "</lemma>\n")
        if not silence_first:
            f.write(sil)
        f.write("</lexicon>\n")


class LinearSearch:
    """The reference's Search::LinearSearch over a lexicon file, a (phoneme, state) -> emission table, transition
    scores tdp[model][type] (models in TDP_MODELS order) and unscaled unigram scores; everything else is read by
    the reference's own classes from its configuration."""
    _count = 0

    def __init__(self, lexicon_file, emission_of, silence_emission, n_emissions, tdp, unigram, states_per_phone=3,
                 state_repetitions=1, lm_scale=1.0, tdp_scale=1.0, pronunciation_scale=0.0, single_word=False,
                 scratch_dir="/tmp", adapter=False):
        S = search_lib()
        LinearSearch._count += 1
        sel = "search-%d-%d" % (os.getpid(), LinearSearch._count)
        put = lambda k, v: S.ref_config_set(("*.%s.%s" % (sel, k)).encode(), str(v).encode())
        put("lexicon.file", lexicon_file)
        put("acoustic-model.hmm.states-per-phone", states_per_phone)
        put("acoustic-model.hmm.state-repetitions", state_repetitions)
        put("acoustic-model.state-tying.type", "lookup")
        put("acoustic-model.state-tying.file", os.path.join(scratch_dir, sel + ".tying"))
        put("acoustic-model.tdp.scale", repr(float(tdp_scale)))
        tdp = np.asarray(tdp, np.float32).reshape(len(TDP_MODELS), len(TDP_TYPES))
        for m, model in enumerate(TDP_MODELS):
            for t, typ in enumerate(TDP_TYPES):
                v = float(tdp[m, t])
                put("acoustic-model.tdp.%s.%s" % (model, typ), "infinity" if np.isinf(v) else repr(v))
        # the reference's default is true (src/Search/LinearSearch.cc:26-30)
        put("recognizer.single-word-recognition", "true" if single_word else "false")
        put("lm.scale", repr(float(lm_scale)))
        put("pronunciation-scale", repr(float(pronunciation_scale)))
        self._emis = np.ascontiguousarray(emission_of, np.int32)
        self._uni = np.ascontiguousarray(unigram, np.float32)
        self.n_emissions = int(n_emissions)
        self._S = S
        self._h = S.ref_search_create(sel.encode(), _p(self._emis), int(states_per_phone), int(silence_emission),
                                      self.n_emissions, _p(self._uni), int(self._uni.size), int(bool(adapter)))
        if not self._h:
            raise RuntimeError("reference LinearSearch could not be set up: %s" % S.ref_last_error().decode())
        self._h = C.c_void_p(self._h)

    def order(self):
        """word number of every lemma pronunciation in the order LinearSearch visits them (-1: silence)"""
        n = self._S.ref_search_order(self._h, None, C.c_long(0))
        out = np.empty(n, np.int32)
        self._S.ref_search_order(self._h, _p(out), C.c_long(n))
        return out

    def states(self, i, capacity=256):
        """(emission, transition model) of every state of pronunciation i, as LinearSearch derives them"""
        e = np.empty(capacity, np.int32)
        m = np.empty(capacity, np.int32)
        n = self._S.ref_search_states(self._h, C.c_long(i), _p(e), _p(m), C.c_long(capacity))
        if n < 0 or n > capacity:
            raise RuntimeError("ref_search_states(%d) -> %d" % (i, n))
        return e[:n].copy(), m[:n].copy()

    def tdps(self):
        out = np.empty((len(TDP_MODELS), len(TDP_TYPES)), np.float32)
        self._S.ref_search_tdps(self._h, _p(out))
        return out

    def run(self, scores):
        scores = np.ascontiguousarray(scores, np.float32)
        T = scores.shape[0]
        assert scores.shape[1] == self.n_emissions
        cap = max(T, 1)
        words, times = np.empty(cap, np.int32), np.empty(cap, np.int32)
        am, lm = np.empty(cap, np.float32), np.empty(cap, np.float32)
        fin = np.zeros(2, np.float32)
        n = self._S.ref_search_run(self._h, _p(scores), C.c_long(T), self.n_emissions, _p(words), _p(times), _p(am),
                                   _p(lm), C.c_long(cap), _p(fin))
        return words[:n].copy(), times[:n].copy(), am[:n].copy(), lm[:n].copy(), fin

    def run_features(self, feature_scorer, feats):
        """the recognizer's loop: every feature vector through `feature_scorer` (a FeatureScorer made with
        library=search_lib()), every scorer object it hands out into the search; then items()"""
        feats = np.ascontiguousarray(feats, np.float32)
        self._S.ref_search_run_features.restype = C.c_long
        n = self._S.ref_search_run_features(self._h, C.c_void_p(feature_scorer._h), _p(feats), C.c_long(feats.shape[0]),
                                            int(feats.shape[1]))
        if n != feats.shape[0]:
            raise RuntimeError("ref_search_run_features fed %d of %d frames" % (n, feats.shape[0]))

    def items(self, capacity=4096):
        """every item of getCurrentBestSentence after run(): (word or -2, time, acoustic, lm) arrays"""
        w, t = np.empty(capacity, np.int32), np.empty(capacity, np.int32)
        a, l = np.empty(capacity, np.float32), np.empty(capacity, np.float32)
        n = self._S.ref_search_items(self._h, _p(w), _p(t), _p(a), _p(l), C.c_long(capacity))
        return w[:n].copy(), t[:n].copy(), a[:n].copy(), l[:n].copy()

    def close(self):
        if self._h:
            self._S.ref_search_destroy(self._h)
            self._h = None


def flat_lexicon(words, emission_of, silence_emission, tdp, unigram, states_per_phone=3, state_repetitions=1,
                 silence=True, silence_first=False, lm_scale=1.0, tdp_scale=1.0, single_word=False, irregular=()):
    """The flat-array form (oracle.h orc_lexicon / rb_lexicon) of what write_lexicon + LinearSearch describe, derived
    HERE from the rules of src/Search/LinearSearch.cc:32-84,472-480 -- tests compare it with what the reference's own
    objects hand out (LinearSearch.order / states / tdps).  Rows of `word` give the word number of every flat entry
    (-1 = silence)."""
    F = np.float32
    emission_of = np.asarray(emission_of, np.int64).reshape(-1, states_per_phone)
    entries = []  # (word number, [phonemes] or None for silence)
    for k, prons in enumerate(words):
        if len(prons) and not isinstance(prons[0], (list, tuple)):
            prons = [prons]
        entries += [(k, list(p)) for p in prons]
    sil = [(-1, None)] if silence else []
    entries = sil + entries if silence_first else entries + sil
    offs, emis, model, uni, word = [0], [], [], [], []
    for k, p in entries:
        if p is None:
            emis.append(silence_emission)
            model.append(2)  # TransitionModel::silence
            uni.append(F(0))
        else:
            for ph in p:
                for a in range(states_per_phone):
                    for b in range(state_repetitions):
                        emis.append(int(emission_of[ph, a]))
                        model.append(3 + b)  # phone0 + b
            uni.append(F(F(lm_scale) * F(unigram[k])))
        offs.append(len(emis))
        word.append(k)
    fmax = np.finfo(np.float32).max
    t = np.asarray(tdp, np.float32).reshape(5, 4)
    with np.errstate(over="ignore", invalid="ignore"):
        t = np.clip(F(tdp_scale) * t, -fmax, fmax).astype(np.float32)  # StateTransitionModel::load: scale * value, Core::clip
    t[0, 0] = t[1, 0] = fmax                                           # TransitionModel::correct(): entry loops forbidden
    return dict(word_offsets=np.asarray(offs, np.uint32), state_emission=np.asarray(emis, np.uint32),
                state_tdp_model=np.asarray(model, np.uint32), tdp=t, entry_model=0,
                unigram=np.asarray(uni, np.float32), word=np.asarray(word, np.int32),
                word_regular=np.asarray([k >= 0 and k not in irregular for k in word], np.uint8),
                single_word=bool(single_word))
