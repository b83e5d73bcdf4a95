"""The CUDA engine (through the C ABI) against the committed fixtures under tests/golden/ -- outputs of the REFERENCE's
own object code (tests/golden/make_golden.py), not of the oracle restatement: MFCC features from a Flow network built by
the reference's NetworkParser from its own mfcc.flow, scores from feature scorers made by its Mm factory.

Bar: scores of the max-approximation scorers, best-density indices, frame counts and f64 time stamps bit-identical;
float features within 1e-4 (north_star), measured against the rms of each feature dimension."""
import os

import numpy as np
import pytest

from rasr_b200 import flow, mm, postproc, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-4


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def dim_err(got, want):
    """largest deviation in units of the rms of the reference's values in that column"""
    scale = np.sqrt(np.mean(want.astype(np.float64) ** 2, axis=0))
    return float((np.abs(got.astype(np.float64) - want) / scale).max())


def test_c1_utterance_against_the_references_flow_network(diag):
    """BASELINE config C1: 10 s, 999 frames x 39"""
    g = load("ref_mfcc_c1.npz")
    x = synth.utterance(int(g["n_samples"]), int(g["seed"]))
    r = flow.FrontEnd().process(x, stages=True)
    assert r["feats"].shape == g["feats_strict"].shape == (999, 39)
    assert np.array_equal(r["t_start"], g["t_start"]) and np.array_equal(r["t_end"], g["t_end"])
    e_strict, e_native = dim_err(r["feats"], g["feats_strict"]), dim_err(r["feats"], g["feats_native"])
    e_cep = dim_err(r["cepstra"], g["cepstra_strict"])
    diag("golden_ref_mfcc_c1", feats_vs_strict=e_strict, feats_vs_native=e_native, cepstra_vs_strict=e_cep)
    assert e_strict < RTOL and e_native < RTOL and e_cep < RTOL
    # static cepstra only: 25 ms time stamps
    s = flow.FrontEnd(derivatives=False).process(x)
    assert np.array_equal(s["t_start"], g["cepstra_t_start"]) and np.array_equal(s["t_end"], g["cepstra_t_end"])
    assert dim_err(s["feats"], g["cepstra_strict"]) < RTOL


def test_stages_against_the_references_nodes(diag):
    g = load("ref_mfcc_stages.npz")
    x = synth.utterance(int(g["n_samples"]), int(g["seed"]))
    r = flow.FrontEnd().process(x, stages=True)
    amp = float(np.abs(r["amplitude"] - g["amplitude"]).max() / g["amplitude"].max())
    fb = float((np.abs(r["fbank"] - g["filterbank"]) / g["filterbank"]).max())
    cep, feats = dim_err(r["cepstra"], g["cepstra"]), dim_err(r["feats"], g["features"])
    diag("golden_ref_mfcc_stages", amplitude_fullscale=amp, fbank_rel=fb, cepstra=cep, feats=feats)
    assert amp < 1e-5 and fb < RTOL and cep < RTOL and feats < RTOL


def test_fft_vectors_of_the_references_translation_unit(diag):
    """tests/golden/fft512_reference.npz: 512-point real transforms computed by the reference's own
    src/Math/FastFourierTransform.cc.  The engine's spectrum is not exposed, its amplitude stage is: frame = the 400
    samples as they are (alpha 0, rectangular window) -> |X[k]| / sample rate."""
    g = load("fft512_reference.npz")
    fe = flow.FrontEnd(alpha=0.0, window_type="rectangular")
    for x, y in zip(g["x"], g["y"]):
        r = fe.process(x[:400], stages=True)
        re, im = y[0::2].astype(np.float64), y[1::2].astype(np.float64)
        # packed real transform (src/Signal/FastFourierTransform.cc:88-94): y[0] = X[0], y[1] = X[N/2], then (re, im)
        want = np.empty(257)
        want[0], want[256] = abs(re[0]), abs(im[0])
        want[1:256] = np.hypot(re[1:], im[1:])
        want /= 16000.0
        got = r["amplitude"][0]
        assert np.abs(got - want).max() / want.max() < 1e-6


def test_dc_detection_against_the_references_node():
    g = load("ref_mfcc_dc.npz")
    r = flow.FrontEnd().process_dc(g["samples"].astype(np.float32))
    assert r["feats"].shape == g["feats"].shape
    assert np.array_equal(r["t_start"], g["t_start"]) and np.array_equal(r["t_end"], g["t_end"])
    assert dim_err(r["feats"], g["feats"]) < RTOL


MODES = [("batch-float", "batch-diagonal-maximum-float"), ("batch-int", "batch-diagonal-maximum-int"),
         ("preselection-batch-float", "preselection-batch-float"), ("preselection-batch-int", "preselection-batch-int")]


def _check_scorers(g, msd, f, diag, tag, presel=None, batch=True):
    gms = mm.MixtureSet.from_dict(msd)
    if batch:
        for contraction, variant in ((False, "strict"), (True, "native")):
            for mode, name in MODES:
                sc = mm.GmmScorer(gms, mode, contraction=contraction)
                if presel and mode.startswith("preselection"):
                    sc.configure_preselection(*presel)
                got, want = sc.score(f), g["%s/%s" % (name, variant)]
                assert np.array_equal(got, want), "%s %s: %d of %d scores differ" % (
                    mode, variant, (got != want).sum(), got.size)
    mx, mb = mm.GmmScorer(gms, "diagonal-maximum", contraction=False).score(f, want_density=True)
    assert np.array_equal(mx, g["diagonal-maximum/strict"]) and np.array_equal(mb, g["diagonal-maximum/strict/best"])
    sm, sb = mm.GmmScorer(gms, "diagonal-sum", contraction=False).score(f, want_density=True)
    ws = g["diagonal-sum/strict"]
    ok = np.isfinite(ws) & (np.abs(ws) < 1e30)
    rel = float((np.abs(sm[ok] - ws[ok]) / np.abs(ws[ok])).max())
    diag("golden_ref_gmm_" + tag, diag_sum_rel=rel)
    assert rel < 1e-6  # expf / logf of CUDA vs glibc
    assert np.array_equal(sb, g["diagonal-sum/strict/best"])
    # the contracted variants against gcc's own contraction choices: an ulp or two, same winners
    mxn, mbn = mm.GmmScorer(gms, "diagonal-maximum", contraction=True).score(f, want_density=True)
    wn = g["diagonal-maximum/native"]
    ok = np.isfinite(wn) & (np.abs(wn) < 1e30)
    assert float((np.abs(mxn[ok] - wn[ok]) / np.abs(wn[ok])).max()) < 1e-6
    assert (mbn != g["diagonal-maximum/native/best"]).mean() < 1e-3


def test_c2_model_all_scorers(diag):
    """BASELINE config C2's model, first 96 frames: reference object code vs CUDA, bit for bit"""
    g = load("ref_gmm_c2.npz")
    _check_scorers(g, synth.mixture_set(), synth.features(100000, 39)[:int(g["n_frames"])], diag, "c2")


def test_ragged_model_all_scorers(diag):
    g = load("ref_gmm_ragged.npz")
    _check_scorers(g, synth.ragged_mixture_set(dim=39, n_covariances=1), synth.features(64, 39, seed=5), diag, "ragged",
                   presel=(16, 4, 5))


def test_ragged_model_three_covariances(diag):
    g = load("ref_gmm_ragged_3cov.npz")
    _check_scorers(g, synth.ragged_mixture_set(dim=24, n_covariances=3, seed=11), synth.features(64, 24, seed=6), diag,
                   "ragged_3cov", batch=False)


@pytest.mark.parametrize("tag", ["c2", "ragged", "ragged_3cov"])
def test_simd_diagonal_maximum(simd_golden_cases, tag):
    """the reference's "SIMD-diagonal-maximum" scorer (object code, tests/golden/make_golden.py) vs RB_GMM_SIMD_DIAG_MAX:
    scores through the IMMA kernel (pooled covariance) and, with densities, through the DP4A kernel -- bit for bit"""
    g = load("ref_gmm_simd.npz")
    msd, f = simd_golden_cases[tag]()
    sc = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "SIMD-diagonal-maximum")
    nonempty = np.diff(msd["mix_offsets"]) > 0
    want, wbest = g["%s/SIMD-diagonal-maximum/native" % tag], g["%s/SIMD-diagonal-maximum/native/best" % tag]
    assert np.array_equal(sc.score(f)[:, nonempty], want[:, nonempty])
    s, b = sc.score(f, want_density=True)
    assert np.array_equal(s[:, nonempty], want[:, nonempty]) and np.array_equal(b[:, nonempty], wbest[:, nonempty])


def test_postprocessing_against_the_references_nodes():
    g = load("ref_postproc.npz")
    f = synth.features(300, 13, seed=5)
    for kind, length, right in (("mean-and-variance", "infinite", "infinite"), ("mean", 51, 25)):
        key = "%s/%s/" % (kind, length)
        n = postproc.PostProcessor(13, kind, length=length, right=right, contraction=False).process(f)
        assert np.array_equal(n, g[key + "normalized"])
        s = postproc.PostProcessor(13, kind, length=length, right=right, splice=(5, 2), contraction=False).process(f)
        assert np.array_equal(s, g[key + "spliced"])
        p = postproc.PostProcessor(13, kind, length=length, right=right, splice=(5, 2), matrix=g["matrix"],
                                   contraction=False).process(f)
        assert np.array_equal(p, g[key + "projected"])
