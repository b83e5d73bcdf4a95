"""The oracle restatement against the REFERENCE's own object code (oracle/_ref/librasr_ref*.so = the reference's Core /
Flow / Math / Signal / Mm translation units compiled from /root/reference by oracle/refbuild/Makefile, driven through
oracle/refbuild/ref_host.cc): the reference's Flow::NetworkParser builds the network from the reference's own mfcc.flow /
derivationWithRegression.flow, the reference's Mm factory creates the feature scorers.

With every operation rounded separately (-ffp-contract=off) the restatement must agree BIT FOR BIT: features, time
stamps, scores, best densities, clusterings.  With gcc's default contraction (what the reference's -march=native build
does) the SSE scorers still agree bit for bit with the oracle's contraction model; the scalar paths (front-end,
diagonal scorers) differ by a few ulp, bounded here.

Runs on the CPU; skipped only where neither the built libraries nor the reference checkout exist."""
import os

import numpy as np
import pytest

from rasr_b200 import io, synth
from tests.helpers_nn import NN_FLOW, nn_files

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref(oracle):
    from oracle import pyref

    if not (pyref.available(False) and pyref.available(True)) and not os.path.isdir(pyref.REFERENCE):
        pytest.skip("oracle/_ref is not built and the reference checkout is absent")
    pyref.lib(False)
    pyref.lib(True)
    return pyref


def same(a, b):
    """identical values; NaN (log10 of an all-zero frame gives -inf, the DCT then inf - inf) equals NaN"""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


HAVE_REFERENCE_FLOWS = os.path.isdir("/root/reference/src/Tools/FeatureExtraction/share")


# ---------------------------------------------------------------------------------------------------------------
# front-end

@pytest.mark.skipif(not HAVE_REFERENCE_FLOWS, reason="needs the reference's own .flow files")
def test_c1_utterance_through_the_references_own_flow_files(ref, oracle):
    """BASELINE config C1: one 10 s utterance through mfcc.flow + derivationWithRegression.flow as the reference ships
    them: 999 frames, the last window 320 samples long, 39 dimensions, all bit-identical to the oracle."""
    x = synth.utterance(160000)
    net = ref.FlowNetwork("mfcc_derivatives.flow", {"nr-cepstrum-coefficients": 13, "block-size": 4096})
    r = net.run(x)
    w = oracle.mfcc(oracle.frontend_cfg(use_fma=False), x, stages=True)
    assert r["feats"].shape == (999, 39)
    assert np.array_equal(r["feats"], w["feats"])
    assert np.array_equal(r["t_start"], w["t_start"]) and np.array_equal(r["t_end"], w["t_end"])
    assert np.array_equal(net.run(x, port="cepstra")["feats"], w["cepstra"])
    # what the chain leaves in the attributes (text): the adapters publish the same
    assert net.attribute("features", "sample-rate") == "1"
    assert net.attribute("features", "frame-shift") == "0.01"
    assert net.attribute("features", "datatype") == "vector-f32"
    # the self-contained network (same nodes, parameters instead of literals) produces identical packets
    c = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters()).run(x)
    assert np.array_equal(c["feats"], r["feats"]) and np.array_equal(c["t_start"], r["t_start"])
    assert np.array_equal(c["t_end"], r["t_end"])


def test_every_stage_of_the_chain(ref, oracle):
    x = synth.utterance(48240, seed=7)
    net = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters())
    w = oracle.mfcc(oracle.frontend_cfg(use_fma=False), x, stages=True)
    T = w["feats"].shape[0]
    fr = net.run(x, port="frames")
    assert fr["feats"].shape[0] == T and fr["sizes"][-1] == 400 and fr["sizes"][0] == 400
    for port, key in (("spectrum", "spectrum"), ("amplitude", "amplitude"), ("filterbank", "fbank"),
                      ("cepstra", "cepstra"), ("features", "feats")):
        r = net.run(x, port=port)
        assert r["feats"].shape == w[key].shape, port
        assert np.array_equal(r["feats"], w[key]), port
    # the windowed frames themselves: pre-emphasis, framing and the Hamming table
    tab = oracle.tables(oracle.frontend_cfg())
    pre = x.copy()
    pre[1:] = x[1:] - x[:-1]
    pre[0] = x[0] - x[0]
    assert np.array_equal(fr["feats"][3], pre[480:880] * tab["window"])


@pytest.mark.parametrize("n", [1, 159, 160, 399, 400, 401, 559, 560, 561, 719, 720, 721, 800, 1040, 16000, 16001])
def test_frame_count_and_tail_at_boundary_lengths(ref, oracle, n):
    """WindowBuffer get / flush protocol (src/Signal/WindowBuffer.cc:50-126): frame count, length of the short last
    window and the f64 time stamps for segment lengths around the window and shift multiples."""
    x = synth.utterance(n, seed=n)
    net = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters(block_size=333))
    cfg = oracle.frontend_cfg(use_fma=False)
    r = net.run(x, port="cepstra")
    assert r["feats"].shape[0] == oracle.nframes(cfg, n)
    if n:
        w = oracle.mfcc(cfg, x, stages=True)
        ws = oracle.mfcc(oracle.frontend_cfg(use_fma=False, derivatives=False), x)  # 25 ms time stamps of the statics
        assert same(r["feats"], w["cepstra"]) and same(r["feats"], ws["feats"])
        assert np.array_equal(r["t_start"], ws["t_start"]) and np.array_equal(r["t_end"], ws["t_end"])
        f = net.run(x)  # static || delta || delta-delta: time stamps merged over the 5-frame regression window
        assert same(f["feats"], w["feats"])
        assert np.array_equal(f["t_start"], w["t_start"]) and np.array_equal(f["t_end"], w["t_end"])


@pytest.mark.parametrize("block", [1, 160, 1000, 100000])
def test_packet_size_does_not_matter(ref, oracle, block):
    x = synth.utterance(20000, seed=3)
    r = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters(block_size=block)).run(x)
    w = oracle.mfcc(oracle.frontend_cfg(use_fma=False), x)
    assert np.array_equal(r["feats"], w["feats"]) and np.array_equal(r["t_end"], w["t_end"])


@pytest.mark.parametrize("sr", [8000.0, 11025.0, 22050.0, 44100.0])
def test_other_sample_rates(ref, oracle, sr):
    """window / FFT length derivation, Hz-per-bin text round trip and filter-bank geometry away from 16 kHz"""
    x = synth.utterance(int(sr * 0.7), seed=int(sr))
    r = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters())
    cfg = oracle.frontend_cfg(sample_rate=sr, use_fma=False)
    w = oracle.mfcc(cfg, x, stages=True)
    for port, key in (("amplitude", "amplitude"), ("filterbank", "fbank"), ("features", "feats")):
        got = r.run(x, port=port, sample_rate=sr)
        assert got["feats"].shape == w[key].shape, (port, got["feats"].shape, w[key].shape)
        assert np.array_equal(got["feats"], w[key]), port
    got = r.run(x, sample_rate=sr)
    assert np.array_equal(got["t_start"], w["t_start"]) and np.array_equal(got["t_end"], w["t_end"])


@pytest.mark.parametrize("wtype", ["hamming", "rectangular", "hanning", "bartlett", "blackman", "kaiser"])
def test_window_functions(ref, oracle, wtype):
    x = synth.utterance(8000, seed=11)
    r = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters(window_type=wtype)).run(x)
    w = oracle.mfcc(oracle.frontend_cfg(use_fma=False, window_type=wtype), x)
    assert np.array_equal(r["feats"], w["feats"])


@pytest.mark.parametrize("kw", [dict(alpha="0.97"), dict(nr_cepstrum_coefficients=16), dict(shift=".005", length=".02"),
                                dict(filter_width="200"), dict(alpha="0", length=".032", maximum_input_size=".032")])
def test_other_chain_parameters(ref, oracle, kw):
    x = synth.utterance(12000, seed=5)
    r = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters(**kw)).run(x)
    o = dict(alpha=float(kw.get("alpha", 1.0)), n_cepstra=int(kw.get("nr_cepstrum_coefficients", 13)),
             window_shift_s=float(kw.get("shift", 0.01)), window_length_s=float(kw.get("length", 0.025)),
             filter_width=float(kw.get("filter_width", 268.258)),
             fft_max_input_s=float(kw.get("maximum_input_size", 0.025)))
    w = oracle.mfcc(oracle.frontend_cfg(use_fma=False, **o), x)
    assert r["feats"].shape == w["feats"].shape
    assert np.array_equal(r["feats"], w["feats"])
    assert np.array_equal(r["t_start"], w["t_start"]) and np.array_equal(r["t_end"], w["t_end"])


def test_two_segments_through_one_network(ref, oracle):
    """state is reset at EOS (src/Signal/Preemphasis.cc:96-106): the second segment is processed like a first one"""
    net = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters())
    for seed, n, t0 in ((1, 9000, 0.0), (2, 7777, 12.5)):
        x = synth.utterance(n, seed=seed)
        r = net.run(x, start_time=t0)
        w = oracle.mfcc(oracle.frontend_cfg(use_fma=False), x)
        assert np.array_equal(r["feats"], w["feats"])
        assert np.allclose(r["t_start"] - t0, w["t_start"], atol=1e-9)


def test_dc_detection_in_front_of_the_chain(ref, oracle):
    """signal-dc-detection (src/Signal/DcDetection.cc) as samples.flow:34-37 wires it: kept runs, their start times,
    the frames and the time stamps after the gaps"""
    x = synth.utterance(40000, seed=9)
    x[5000:9000] = x[4999]
    x[15000:15300] = 7.0
    x[20000:20250] = -3.0  # 250 samples + neighbours: around the minimum DC length of 200
    x[30000:30190] = 1.0
    net = ref.FlowNetwork("mfcc_chain_dc.flow", ref.chain_parameters(dc=True))
    r = net.run(x)
    w = oracle.mfcc_dc(oracle.frontend_cfg(use_fma=False), oracle.dc_cfg(), x)
    assert r["feats"].shape == w["feats"].shape and r["feats"].shape[0] < oracle.nframes(oracle.frontend_cfg(), x.size)
    assert np.array_equal(r["feats"], w["feats"])
    assert np.array_equal(r["t_start"], w["t_start"]) and np.array_equal(r["t_end"], w["t_end"])
    # the sample stream the detector lets through: total length and the start time of every kept run
    s = net.run(x, port="samples")
    assert int(s["sizes"].sum()) == int((w["run_end"] - w["run_begin"]).sum())
    assert set(np.round(w["run_start"], 12)) <= set(np.round(s["t_start"], 12))


def test_native_build_is_within_a_few_ulp(ref, oracle):
    """gcc's own contraction choices (the reference's default -march=native build) against the strict build"""
    x = synth.utterance(32000, seed=13)
    a = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters(), native=False).run(x)
    b = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters(), native=True).run(x)
    assert np.array_equal(a["t_start"], b["t_start"]) and np.array_equal(a["t_end"], b["t_end"])
    scale = np.sqrt(np.mean(a["feats"] ** 2, axis=0))
    assert (np.abs(a["feats"] - b["feats"]) / scale).max() < 1e-5  # of the rms of the dimension (10x below the 1e-4 tolerance)


# ---------------------------------------------------------------------------------------------------------------
# post-processing nodes and the Math::Matrix file formats (8f-3)

@pytest.mark.parametrize("form", ["xml", "bin"])
@pytest.mark.parametrize("kind,length,right", [("mean-and-variance", "infinite", "infinite"), ("mean", 51, 25),
                                               ("mean-and-variance", 21, 0), ("mean", 5, 4)])
def test_normalization_splice_and_matrix_nodes(ref, oracle, tmp_path, form, kind, length, right):
    """signal-normalization -> signal-vector-f32-sequence-concatenation -> signal-matrix-multiplication-f32 as lda.flow
    wires them; the matrix file is written by rasr_b200.io and read by the reference's own Math::Matrix reader"""
    f = synth.features(300, 13, seed=5)
    M = np.random.default_rng(3).standard_normal((20, 65)).astype(np.float32)
    path = "%s:%s" % (form, tmp_path / ("lda." + form))
    io.write_matrix(path, M)
    P = {"block-size": 13, "norm-type": kind, "norm-length": length, "norm-right": right, "splice-length": 5,
         "splice-right": 2, "matrix-file": path}
    net = ref.FlowNetwork("postproc_chain.flow", P)
    L = -1 if length == "infinite" else length
    R = -1 if right == "infinite" else right
    wn = oracle.normalize(f, kind=kind, length=L, right=R, use_fma=False)
    ws = oracle.splice(wn, 5, 2)
    wp = oracle.matmul(M, ws, use_fma=False)
    assert np.array_equal(net.run(f.reshape(-1), port="normalized", sample_rate=1300.0)["feats"], wn)
    assert np.array_equal(net.run(f.reshape(-1), port="spliced", sample_rate=1300.0)["feats"], ws)
    assert np.array_equal(net.run(f.reshape(-1), port="projected", sample_rate=1300.0)["feats"], wp)


# ---------------------------------------------------------------------------------------------------------------
# Mm feature scorers (rows a10-a14, f4) through the recognizer's buffered call protocol

SCORERS = [("batch-diagonal-maximum-float", lambda o, ms, f, fma: o.gmm_batch_float(ms, f, use_fma=fma)),
           ("batch-diagonal-maximum-int", lambda o, ms, f, fma: o.gmm_batch_int(ms, f)),
           ("batch-diagonal-maximum-fast", lambda o, ms, f, fma: o.gmm_batch_int(ms, f)),
           ("preselection-batch-float", lambda o, ms, f, fma: o.gmm_preselect_float(ms, f, use_fma=fma)[0]),
           ("preselection-batch-int", lambda o, ms, f, fma: o.gmm_preselect_int(ms, f)[0])]


@pytest.mark.parametrize("native", [False, True])
@pytest.mark.parametrize("name,fn", SCORERS, ids=[s[0] for s in SCORERS])
def test_c2_model_batch_scorers(ref, oracle, name, fn, native):
    """BASELINE config C2's model (39 dims, 4096 densities, 256 mixtures), 400 frames: every score bit-identical, in the
    strict build and in the one with gcc's default contraction"""
    msd = synth.mixture_set()
    ms = oracle.MixtureSet(**msd)
    f = synth.features(400, 39)
    got = ref.FeatureScorer(ms, name, native=native).score(f)
    want = fn(oracle, ms, f, native)
    assert got.shape == (400, 256)
    assert np.array_equal(got, want), "%d of %d scores differ" % ((got != want).sum(), got.size)


@pytest.mark.parametrize("name", ["diagonal-maximum", "diagonal-sum"])
def test_c2_model_diagonal_scorers(ref, oracle, name):
    msd = synth.mixture_set()
    ms = oracle.MixtureSet(**msd)
    f = synth.features(300, 39)
    fn = oracle.gmm_diag_max if name == "diagonal-maximum" else oracle.gmm_diag_sum
    got, best = ref.FeatureScorer(ms, name, native=False).score(f, want_best=True)
    want, wbest = fn(ms, f, use_fma=False)
    assert np.array_equal(got, want) and np.array_equal(best, wbest)
    # gcc's contraction moves single scores by an ulp or two, never the best density of a clear winner
    gotn, bestn = ref.FeatureScorer(ms, name, native=True).score(f, want_best=True)
    assert (np.abs(gotn - want) / np.abs(want)).max() < 1e-6
    assert (bestn != wbest).mean() < 1e-3


@pytest.mark.parametrize("native", [False, True])
def test_c2_model_simd_diagonal_maximum(ref, oracle, native):
    """"SIMD-diagonal-maximum" (src/Mm/SimdFeatureScorer.cc): scores and best densities bit-identical in both builds
    (integer distances; the f32 / f64 mixing of init() has no multiply-add to contract).  The reference's run-time
    code generator is compiled with -DPROC_x86_64 (oracle/refbuild/Makefile): without it the scorer crashes."""
    msd = synth.mixture_set()
    ms = oracle.MixtureSet(**msd)
    f = synth.features(400, 39)
    got, best = ref.FeatureScorer(ms, "SIMD-diagonal-maximum", native=native).score(f, want_best=True)
    want, wbest = oracle.gmm_simd_diag_max(ms, f, want_best=True)
    assert np.array_equal(got, want) and np.array_equal(best, wbest)
    # a quantised version of diagonal-maximum: same best density almost everywhere, scores within 10 %
    exact, ebest = oracle.gmm_diag_max(ms, f, use_fma=False)
    assert (np.abs(got - exact) / exact).max() < 0.1 and (best == ebest).mean() > 0.9


def _ragged_model(dim, seed):
    rng = np.random.default_rng(seed)
    sizes = [0, 1, 3, 16, 2, 0, 7, 33, 1, 5]
    n_dens = sum(sizes)
    n_cov = 3
    return dict(dim=dim, mix_offsets=np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32),
                mix_density=rng.permutation(n_dens).astype(np.uint32),
                mix_log_weight=np.log(rng.uniform(0.05, 1.0, n_dens)),
                dens_mean=rng.permutation(n_dens).astype(np.uint32),
                dens_cov=rng.integers(0, n_cov, n_dens).astype(np.uint32),
                means=rng.standard_normal((n_dens, dim)).astype(np.float32),
                variances=rng.uniform(0.3, 3.0, (n_cov, dim)).astype(np.float32))


@pytest.mark.parametrize("dim", [1, 7, 8, 9, 40, 48])
def test_ragged_models_and_other_dimensions(ref, oracle, dim):
    """empty / single-density / large mixtures, permuted density and mean tables, several covariances (diagonal
    scorers) or one pooled covariance (batch scorers), dimensions around the 8- and 16-element padding"""
    msd = _ragged_model(dim, seed=dim)
    f = synth.features(64, dim, seed=dim + 1)
    ms = oracle.MixtureSet(**msd)
    for name, fn in (("diagonal-maximum", oracle.gmm_diag_max), ("diagonal-sum", oracle.gmm_diag_sum)):
        got, best = ref.FeatureScorer(ms, name).score(f, want_best=True)
        want, wbest = fn(ms, f, use_fma=False)
        nonempty = np.diff(msd["mix_offsets"]) > 0
        assert np.array_equal(got[:, nonempty], want[:, nonempty]), name
        assert np.array_equal(best[:, nonempty], wbest[:, nonempty]), name
    # one quantised feature vector per covariance; features far outside the range of the means saturate at 0 / 255
    for feats in (f, 4.0 * f):
        got, best = ref.FeatureScorer(ms, "SIMD-diagonal-maximum").score(feats, want_best=True)
        want, wbest = oracle.gmm_simd_diag_max(ms, feats, want_best=True)
        assert np.array_equal(got[:, nonempty], want[:, nonempty]) and np.array_equal(best[:, nonempty], wbest[:, nonempty])
    pooled = dict(msd, dens_cov=np.zeros_like(msd["dens_cov"]), variances=msd["variances"][:1])
    ms1 = oracle.MixtureSet(**pooled)
    for name, fn in SCORERS[:3]:
        if name.endswith("-fast") and not 33 <= dim <= 48:
            # BatchUnrolledIntFeatureScorer::fillScoreCache steps through the means with the fixed stride 48
            # (src/Mm/BatchFeatureScorer.cc:581,632) although init() lays them out with paddedDimension_: for padded
            # dimensions other than 48 the reference reads the wrong rows (and past the table) -- undefined, not a target
            continue
        got = ref.FeatureScorer(ms1, name).score(f)
        want = fn(oracle, ms1, f, False)
        assert np.array_equal(got, want), name


@pytest.mark.parametrize("clusters,select,iterations", [(64, 8, 5), (16, 16, 2), (256, 40, 1), (7, 3, 9)])
def test_preselection_parameters(ref, oracle, clusters, select, iterations):
    """density-clustering.* resources (src/Mm/DensityClustering.cc:20-34) reach the reference scorer through its own
    configuration; clustering (rand()-seeded k-means) and cluster choice must agree for every parameter set"""
    msd = synth.mixture_set(n_mixtures=32)
    ms = oracle.MixtureSet(**msd)
    f = synth.features(200, 39, seed=4)
    cfg = {"density-clustering.clusters": clusters, "density-clustering.select-clusters": select,
           "density-clustering.iterations": iterations}
    got = ref.FeatureScorer(ms, "preselection-batch-float", dict(cfg, **{"density-clustering.backoff-score": 1234.5})).score(f)
    want = oracle.gmm_preselect_float(ms, f, use_fma=False, clusters=clusters, select=select, iterations=iterations,
                                      backoff=1234.5)[0]
    assert np.array_equal(got, want)
    got = ref.FeatureScorer(ms, "preselection-batch-int", cfg).score(f)
    want = oracle.gmm_preselect_int(ms, f, clusters=clusters, select=select, iterations=iterations)[0]
    assert np.array_equal(got, want)


def test_int_preselection_with_tied_distances(ref, oracle):
    """densities that share four means: most cluster distances tie and the reference's std::sort decides which clusters
    survive -- the oracle's restated introsort must make the same choice as the reference's object code"""
    msd = synth.mixture_set(n_mixtures=64)
    rng = np.random.default_rng(8)
    base = rng.standard_normal((4, 39)).astype(np.float32)
    msd["means"] = base[rng.integers(0, 4, msd["means"].shape[0])]
    ms = oracle.MixtureSet(**msd)
    f = synth.features(300, 39, seed=6)
    cfg = {"density-clustering.clusters": 64, "density-clustering.select-clusters": 8}
    got = ref.FeatureScorer(ms, "preselection-batch-int", cfg).score(f)
    for restated in (False, True):
        want = oracle.gmm_preselect_int(ms, f, clusters=64, select=8, restated_sort=restated)[0]
        assert np.array_equal(got, want), restated


@pytest.mark.parametrize("buffer_size", [1, 4, 13])
def test_buffer_size_does_not_change_scores(ref, oracle, buffer_size):
    """BatchFeatureScorerBase ring buffer (src/Mm/BatchFeatureScorer.hh:34-199): the scorer handed out belongs to the
    oldest buffered frame whatever the buffer size"""
    msd = synth.mixture_set(n_mixtures=16)
    ms = oracle.MixtureSet(**msd)
    f = synth.features(50, 39, seed=2)
    got = ref.FeatureScorer(ms, "batch-diagonal-maximum-float", {"buffer-size": buffer_size}).score(f)
    assert np.array_equal(got, oracle.gmm_batch_float(ms, f, use_fma=False))


# ---------------------------------------------------------------------------------------------------------------
# legacy Nn module (rows a15-a17): the reference's own Nn::BatchFeatureScorer and neural-network-forward node, configured
# with the reference's keys, parameter / prior files written by rasr_b200.io (row f3) -- against the oracle's f32 path.
# The BLAS below the reference is the plain-loop stand-in (oracle/refbuild/miniblas.cc), sequential f32 accumulation.

@pytest.mark.parametrize("hidden", ["sigmoid", "relu", "tanh"])
def test_nn_batch_feature_scorer(ref, oracle, tmp_path, hidden):
    net = synth.network(dims=(20, 32, 24, 16), hidden=hidden, seed=3)
    cfg = nn_files(tmp_path, net, hidden)
    io.write_vector("xml:" + str(tmp_path / "prior.xml"), net["log_prior"])
    cfg.update({"prior-file": "xml:" + str(tmp_path / "prior.xml"), "priori-scale": 0.7, "buffer-size": 8})
    ms = oracle.MixtureSet(**synth.mixture_set(dim=20, n_mixtures=16, densities_per_mixture=1))
    x = synth.features(37, 20, seed=9, scale=1.0)
    got = ref.FeatureScorer(ms, "nn-batch-feature-scorer", cfg).score(x)
    want = oracle.nn_scores(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 0.7, x,
                            mode=oracle.NN_F32)
    assert got.shape == (37, 16)
    if hidden == "tanh":  # tanh of libm vs the oracle's formulation: an ulp
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-6
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("name", ["nn-full-hybrid"])
def test_nn_frame_at_a_time_scorers(ref, oracle, tmp_path, name):
    """the reference's unbuffered Nn scorer (src/Nn/FeatureScorer.cc:187-253: the whole network per frame) computes the
    function of nn-batch-feature-scorer: -(w_e.h + b_e - priori-scale * logprior_e), softmax not evaluated.  One oracle
    (and one CUDA path, the b200-nn-batch-feature-scorer adapter) therefore stands for both.
    (nn-on-demand-hybrid, :97-170, is meant to compute the same per requested output unit, but it cannot be run: its
    forwardHiddenLayers() calls finishComputation() on the activation that LinearAndSoftmaxLayer::getScore then hands
    to CudaMatrix::dotWithColumn, whose precondition X.isComputing_ aborts the process -- observed with the reference's
    object code, src/Math/CudaMatrix.hh:1052.)"""
    net = synth.network(dims=(20, 32, 24, 16), hidden="sigmoid", seed=3)
    cfg = nn_files(tmp_path, net, "sigmoid")
    io.write_vector("xml:" + str(tmp_path / "prior.xml"), net["log_prior"])
    cfg.update({"prior-file": "xml:" + str(tmp_path / "prior.xml"), "priori-scale": 0.7})
    ms = oracle.MixtureSet(**synth.mixture_set(dim=20, n_mixtures=16, densities_per_mixture=1))
    x = synth.features(23, 20, seed=9, scale=1.0)
    got = ref.FeatureScorer(ms, name, cfg).score(x)
    want = oracle.nn_scores(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 0.7, x,
                            mode=oracle.NN_F32)
    batch = ref.FeatureScorer(ms, "nn-batch-feature-scorer", dict(cfg, **{"buffer-size": 8})).score(x)
    assert got.shape == (23, 16)
    # matrix-vector instead of matrix-matrix products below the same layers: summation order may differ by an ulp
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6
    assert np.abs(got - batch).max() / np.abs(batch).max() < 2e-6


def test_nn_prior_from_mixture_weights(ref, oracle, tmp_path):
    """without a prior file the prior is the relative mixture-weight mass of each class (Prior::setFromMixtureSet,
    src/Nn/Prior.cc:158-188)"""
    net = synth.network(dims=(12, 16, 8), hidden="sigmoid", seed=5)
    cfg = nn_files(tmp_path, net, "sigmoid")
    msd = synth.ragged_mixture_set(dim=12, sizes=(1, 3, 2, 5, 1, 4, 2, 2), seed=4)
    # linear weights (not normalised per mixture) so that the classes carry different mass
    rng = np.random.default_rng(1)
    msd["mix_log_weight"] = np.log(rng.uniform(0.1, 2.0, msd["mix_log_weight"].size))
    ms = oracle.MixtureSet(**msd)
    x = synth.features(20, 12, seed=10, scale=1.0)
    got = ref.FeatureScorer(ms, "nn-batch-feature-scorer", cfg).score(x)
    mass = np.array([np.exp(msd["mix_log_weight"][a:b]).astype(np.float32).sum(dtype=np.float32)
                     for a, b in zip(msd["mix_offsets"][:-1], msd["mix_offsets"][1:])], np.float32)
    log_prior = np.log(mass / mass.sum(dtype=np.float32))
    want = oracle.nn_scores(net["dims"], net["acts"], net["weights"], net["biases"], log_prior, 1.0, x, mode=oracle.NN_F32)
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6


def test_neural_network_forward_node(ref, oracle, tmp_path):
    """the reference's neural-network-forward Flow node (src/Nn/NeuralNetworkForwardNode.cc): softmax output, one packet
    per input packet with its time stamp, buffer size irrelevant"""
    net = synth.network(dims=(20, 32, 16), hidden="sigmoid", seed=6)
    cfg = nn_files(tmp_path, net, "sigmoid")
    for k, v in cfg.items():
        ref.config_set("*.nnflow.nn." + k, v)
    ref.config_set("*.nnflow.nn.buffer-size", 7)
    (tmp_path / "nn.flow").write_text(NN_FLOW % "neural-network-forward")
    x = synth.features(45, 20, seed=12, scale=1.0)
    r = ref.FlowNetwork(str(tmp_path / "nn.flow"), {"block-size": 20}, selection="nnflow").run(x.reshape(-1), sample_rate=2000.0)
    want = oracle.nn_forward(net["dims"], net["acts"], net["weights"], net["biases"], x, mode=oracle.NN_F32)
    assert r["feats"].shape == want.shape
    assert np.abs(r["feats"] - want).max() < 1e-6 and np.allclose(r["feats"].sum(axis=1), 1.0, atol=1e-5)
    # NeuralNetworkForwardNode::putNextFeature (src/Nn/NeuralNetworkForwardNode.cc:171) means to copy the input packet's
    # time stamp, but its `cond ? aggregateBuffer_[i] : featureBuffer_[i]` mixes two DataPtr types, which only convert
    # to each other through bool: every output packet of the reference carries the time stamp [1, 1].  The adapter
    # (b200-neural-network-forward) carries the input packet's time stamp, as the statement intends.
    assert np.array_equal(r["t_start"], np.ones(45)) and np.array_equal(r["t_end"], np.ones(45))
