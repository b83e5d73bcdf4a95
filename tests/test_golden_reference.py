"""The oracle restatement against the committed fixtures under tests/golden/ -- outputs of the REFERENCE's own object
code (tests/golden/make_golden.py: Flow networks built by the reference's NetworkParser from its own mfcc.flow, scorers
made by its Mm factory).  Unlike tests/test_ref_parity.py this needs neither the reference checkout nor oracle/_ref, so
it also runs on the GPU box: the oracle that the -m gpu tests check the CUDA path against is itself pinned there."""
import os

import numpy as np
import pytest

from rasr_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def test_c1_mfcc_strict_is_bit_identical(oracle):
    g = load("ref_mfcc_c1.npz")
    x = synth.utterance(int(g["n_samples"]), int(g["seed"]))
    w = oracle.mfcc(oracle.frontend_cfg(use_fma=False), x, stages=True)
    assert g["feats_strict"].shape == (999, 39)
    assert np.array_equal(w["feats"], g["feats_strict"])
    assert np.array_equal(w["cepstra"], g["cepstra_strict"])
    assert np.array_equal(w["t_start"], g["t_start"]) and np.array_equal(w["t_end"], g["t_end"])
    ws = oracle.mfcc(oracle.frontend_cfg(use_fma=False, derivatives=False), x)
    assert np.array_equal(ws["t_start"], g["cepstra_t_start"]) and np.array_equal(ws["t_end"], g["cepstra_t_end"])
    # the oracle's contraction model against gcc's own choices (native build): a few ulp
    wf = oracle.mfcc(oracle.frontend_cfg(use_fma=True), x)
    scale = np.sqrt(np.mean(g["feats_native"] ** 2, axis=0))
    assert (np.abs(wf["feats"] - g["feats_native"]) / scale).max() < 2e-5


def test_every_stage(oracle):
    g = load("ref_mfcc_stages.npz")
    x = synth.utterance(int(g["n_samples"]), int(g["seed"]))
    w = oracle.mfcc(oracle.frontend_cfg(use_fma=False), x, stages=True)
    for port, key in (("spectrum", "spectrum"), ("amplitude", "amplitude"), ("filterbank", "fbank"),
                      ("cepstra", "cepstra"), ("features", "feats")):
        assert np.array_equal(w[key], g[port]), port


def test_dc_detection(oracle):
    g = load("ref_mfcc_dc.npz")
    w = oracle.mfcc_dc(oracle.frontend_cfg(use_fma=False), oracle.dc_cfg(), g["samples"].astype(np.float32))
    assert np.array_equal(w["feats"], g["feats"])
    assert np.array_equal(w["t_start"], g["t_start"]) and np.array_equal(w["t_end"], g["t_end"])


def test_fft_translation_unit(oracle):
    g = load("fft512_reference.npz")
    for x, y in zip(g["x"], g["y"]):
        assert np.array_equal(oracle.fft_real_packed(x), y)


def _check_scorers(oracle, g, ms, f, presel=None):
    kw = presel or {}
    for tag, fma in (("strict", False), ("native", True)):
        if "batch-diagonal-maximum-float/" + tag in g:
            assert np.array_equal(oracle.gmm_batch_float(ms, f, use_fma=fma), g["batch-diagonal-maximum-float/" + tag])
            assert np.array_equal(oracle.gmm_batch_int(ms, f), g["batch-diagonal-maximum-int/" + tag])
            assert np.array_equal(oracle.gmm_preselect_float(ms, f, use_fma=fma, **kw)[0], g["preselection-batch-float/" + tag])
            assert np.array_equal(oracle.gmm_preselect_int(ms, f, **kw)[0], g["preselection-batch-int/" + tag])
    for name, fn in (("diagonal-maximum", oracle.gmm_diag_max), ("diagonal-sum", oracle.gmm_diag_sum)):
        s, b = fn(ms, f, use_fma=False)
        assert np.array_equal(s, g[name + "/strict"]) and np.array_equal(b, g[name + "/strict/best"]), name
        sn, bn = fn(ms, f, use_fma=True)  # contraction model vs gcc's choices: an ulp or two
        assert (np.abs(sn - g[name + "/native"]) / np.abs(g[name + "/native"])).max() < 1e-6
        assert (bn != g[name + "/native/best"]).mean() < 1e-3


def test_c2_model_all_scorers(oracle):
    g = load("ref_gmm_c2.npz")
    ms = oracle.MixtureSet(**synth.mixture_set())
    f = synth.features(100000, 39)[:int(g["n_frames"])]
    _check_scorers(oracle, g, ms, f)


def test_ragged_model_all_scorers(oracle):
    g = load("ref_gmm_ragged.npz")
    ms = oracle.MixtureSet(**synth.ragged_mixture_set(dim=39, n_covariances=1))
    _check_scorers(oracle, g, ms, synth.features(64, 39, seed=5), presel=dict(clusters=16, select=4))


def test_ragged_model_three_covariances(oracle):
    g = load("ref_gmm_ragged_3cov.npz")
    ms = oracle.MixtureSet(**synth.ragged_mixture_set(dim=24, n_covariances=3, seed=11))
    _check_scorers(oracle, g, ms, synth.features(64, 24, seed=6))


@pytest.mark.parametrize("tag", ["c2", "ragged", "ragged_3cov"])
def test_simd_diagonal_maximum(oracle, simd_golden_cases, tag):
    """"SIMD-diagonal-maximum": integer arithmetic, so the strict and the contracted build of the reference agree and
    the oracle reproduces both, best densities included"""
    g = load("ref_gmm_simd.npz")
    msd, f = simd_golden_cases[tag]()
    s, b = oracle.gmm_simd_diag_max(oracle.MixtureSet(**msd), f, want_best=True)
    nonempty = np.diff(msd["mix_offsets"]) > 0
    for build in ("strict", "native"):
        key = "%s/SIMD-diagonal-maximum/%s" % (tag, build)
        assert np.array_equal(s[:, nonempty], g[key][:, nonempty])
        assert np.array_equal(b[:, nonempty], g[key + "/best"][:, nonempty])


def test_postprocessing(oracle):
    g = load("ref_postproc.npz")
    f = synth.features(300, 13, seed=5)
    for kind, length in (("mean-and-variance", "infinite"), ("mean", 51)):
        L, R = (-1, -1) if length == "infinite" else (51, 25)
        wn = oracle.normalize(f, kind=kind, length=L, right=R, use_fma=False)
        ws = oracle.splice(wn, 5, 2)
        assert np.array_equal(wn, g["%s/%s/normalized" % (kind, length)])
        assert np.array_equal(ws, g["%s/%s/spliced" % (kind, length)])
        assert np.array_equal(oracle.matmul(g["matrix"], ws, use_fma=False), g["%s/%s/projected" % (kind, length)])
