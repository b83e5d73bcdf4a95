"""GPU parity of the GMM scorers against the CPU oracle, through the C ABI.

Bar (BASELINE.json north_star): density / mixture / frame indices bit-exact, log-likelihoods within
1e-4 relative.  RB_GMM_BATCH_FLOAT and RB_GMM_DIAG_MAX follow the reference's accumulation order, so
they are checked for BIT-IDENTICAL scores; RB_GMM_DIAG_SUM differs only through expf/logf (libm vs CUDA).
"""
import os

import numpy as np
import pytest

from rasr_b200 import capi, mm, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-4  # north_star tolerance for float log-likelihoods


def both(oracle, msd):
    return oracle.MixtureSet(**msd), mm.MixtureSet.from_dict(msd)


@pytest.fixture(params=["default", "direct"], autouse=True)
def route(request, monkeypatch):
    """every case twice: the default routing (tensor-core screening + exact refinement wherever the model allows it, at
    any batch size) and the direct-form kernels alone (RB_GMM_EXACT=0: what ineligible models run on)"""
    if request.param == "direct":
        monkeypatch.setenv("RB_GMM_EXACT", "0")
    return request.param


@pytest.mark.parametrize("contraction", [True, False])
def test_batch_float_c2_shape_bit_exact(oracle, diag, contraction):
    """C2 geometry (39-dim, 256 mixtures x 16 densities) at a size the oracle finishes in seconds."""
    msd = synth.mixture_set()
    oms, gms = both(oracle, msd)
    f = synth.features(3000, 39)
    want = oracle.gmm_batch_float(oms, f, use_fma=contraction, threads=8)
    got = mm.GmmScorer(gms, "batch-float", contraction=contraction).score(f)
    diag("gmm_batch_c2", contraction=contraction, max_abs=np.abs(got - want).max(),
         n_diff=int((got != want).sum()), total=got.size)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dim", [1, 7, 8, 13, 33, 39, 40, 45, 64])
def test_batch_float_dimensions(oracle, dim):
    msd = synth.mixture_set(dim=dim, n_mixtures=12, densities_per_mixture=5, seed=dim)
    oms, gms = both(oracle, msd)
    f = synth.features(700, dim, seed=dim)
    assert np.array_equal(mm.GmmScorer(gms).score(f), oracle.gmm_batch_float(oms, f))


@pytest.mark.parametrize("T", [1, 2, 255, 256, 257, 511, 512, 513, 1025])
def test_batch_float_ragged_frame_counts(oracle, T):
    msd = synth.mixture_set(dim=39, n_mixtures=8, densities_per_mixture=16, seed=3)
    oms, gms = both(oracle, msd)
    f = synth.features(T, 39, seed=T)
    got = mm.GmmScorer(gms).score(f)
    assert got.shape == (T, 8)
    assert np.array_equal(got, oracle.gmm_batch_float(oms, f))


def test_empty_input(oracle):
    gms = mm.MixtureSet.from_dict(synth.mixture_set(dim=39, n_mixtures=8, densities_per_mixture=4))
    got = mm.GmmScorer(gms).score(np.zeros((0, 39), np.float32))
    assert got.shape == (0, 8)


def test_ragged_mixtures(oracle, diag):
    """Mixtures of unequal size sharing densities out of order (the reference-made fixture of the same model is
    checked in tests/test_gpu_golden_reference.py)."""
    msd = synth.ragged_mixture_set(dim=39, n_covariances=1)
    oms, gms = both(oracle, msd)
    f = synth.features(64, 39, seed=5)
    got = mm.GmmScorer(gms, "batch-float").score(f)
    assert np.array_equal(got, oracle.gmm_batch_float(oms, f))
    mx, mb = mm.GmmScorer(gms, "diagonal-maximum").score(f, want_density=True)
    wx, wb = oracle.gmm_diag_max(oms, f)
    assert np.array_equal(mx, wx) and np.array_equal(mb, wb)
    sm, sb = mm.GmmScorer(gms, "diagonal-sum").score(f, want_density=True)
    ws, wsb = oracle.gmm_diag_sum(oms, f)
    diag("gmm_sum_ragged", max_rel=(np.abs(sm - ws) / np.abs(ws)).max())
    np.testing.assert_allclose(sm, ws, rtol=RTOL)
    assert np.array_equal(sb, wsb)


def test_mixture_with_no_density_scores_flt_max(oracle):
    msd = synth.ragged_mixture_set(dim=16, sizes=(3, 0, 5, 1), seed=2)
    oms, gms = both(oracle, msd)
    f = synth.features(40, 16, seed=1)
    want = oracle.gmm_batch_float(oms, f)
    got = mm.GmmScorer(gms).score(f)
    assert np.array_equal(got, want)
    assert np.all(got[:, 1] == np.finfo(np.float32).max)


@pytest.mark.parametrize("ncov", [1, 3])
@pytest.mark.parametrize("dim", [39, 40, 13, 7, 2])
def test_diagonal_maximum_scores_and_density_index_bit_exact(oracle, diag, dim, ncov):
    msd = synth.ragged_mixture_set(dim=dim, sizes=(1, 3, 16, 7, 32, 2, 9, 4), seed=dim + ncov, n_covariances=ncov)
    oms, gms = both(oracle, msd)
    f = synth.features(900, dim, seed=9)
    want, wbest = oracle.gmm_diag_max(oms, f)
    got, best = mm.GmmScorer(gms, "diagonal-maximum").score(f, want_density=True)
    diag("gmm_diag_max", dim=dim, ncov=ncov, n_score_diff=int((got != want).sum()),
         n_index_diff=int((best != wbest).sum()))
    assert np.array_equal(best, wbest)
    assert np.array_equal(got, want)


def test_diagonal_maximum_scales(oracle):
    msd = synth.ragged_mixture_set(dim=39, n_covariances=2, seed=4)
    oms, gms = both(oracle, msd)
    f = synth.features(300, 39, seed=2)
    want, wbest = oracle.gmm_diag_max(oms, f, mixture_weight_scale=0.7, gaussian_scale=1.3)
    got, best = mm.GmmScorer(gms, "diagonal-maximum", mixture_weight_scale=0.7, gaussian_scale=1.3).score(
        f, want_density=True)
    assert np.array_equal(best, wbest)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dim", [39, 13, 6])
def test_diagonal_sum_log_sum_exp(oracle, diag, dim):
    msd = synth.ragged_mixture_set(dim=dim, sizes=(1, 3, 16, 7, 32, 2), seed=dim, n_covariances=2)
    oms, gms = both(oracle, msd)
    f = synth.features(900, dim, seed=11)
    want, wbest = oracle.gmm_diag_sum(oms, f)
    got, best = mm.GmmScorer(gms, "diagonal-sum").score(f, want_density=True)
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
    diag("gmm_diag_sum", dim=dim, max_rel=rel.max(), n_index_diff=int((best != wbest).sum()))
    assert np.array_equal(best, wbest)
    assert rel.max() < RTOL


def test_batch_float_rejects_multiple_covariances():
    gms = mm.MixtureSet.from_dict(synth.ragged_mixture_set(dim=8, n_covariances=2))
    with pytest.raises(capi.RasrB200Error) as e:
        mm.GmmScorer(gms, "batch-float")
    assert e.value.status == -4


def test_full_size_c2_properties(oracle, diag):
    """BASELINE config C2 at full size (100k frames x 256 mixtures): checked through size-independent
    properties -- a strided sample of frames against the oracle (bit-exact), invariance under frame
    permutation, and max-approximation <= every single-density score of the mixture."""
    msd = synth.mixture_set()
    oms, gms = both(oracle, msd)
    T = 100000
    f = synth.features(T, 39)
    scorer = mm.GmmScorer(gms)
    got = scorer.score(f)
    idx = np.arange(0, T, 97)
    want = oracle.gmm_batch_float(oms, f[idx], threads=8)
    n_diff = int((got[idx] != want).sum())
    diag("gmm_c2_full", sampled=idx.size, n_diff=n_diff)
    assert n_diff == 0
    perm = np.random.default_rng(0).permutation(T)
    assert np.array_equal(scorer.score(f[perm]), got[perm])
    assert np.isfinite(got).all()


def test_buffered_feature_scorer_over_the_device(oracle):
    """The Mm::FeatureScorer protocol mirror on top of the dense device scores."""
    msd = synth.mixture_set(dim=39, n_mixtures=16, densities_per_mixture=8, seed=6)
    oms, gms = both(oracle, msd)
    f = synth.features(50, 39, seed=6)
    want = oracle.gmm_batch_float(oms, f)
    fs = mm.BatchFeatureScorer(mm.GmmScorer(gms))
    for x in f:
        fs.add_feature(x)
    t = 0
    while not fs.buffer_empty():
        s = fs.flush()
        assert s.n_emissions() == 16
        assert s.score(3) == want[t, 3]
        assert np.array_equal(s.scores(), want[t])
        t += 1
    assert t == 50
