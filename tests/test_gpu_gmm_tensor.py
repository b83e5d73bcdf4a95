"""RB_GMM_BATCH_TENSOR: the split-precision tcgen05 formulation of the pooled-covariance GMM scorer.
Not bit-identical by construction; the bar is the north_star tolerance of 1e-4 relative."""
import numpy as np
import pytest

from rasr_b200 import capi, mm, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel(got, want):
    return float((np.abs(got.astype(np.float64) - want) / np.abs(want)).max())


def test_c2_shape(oracle, diag):
    msd = synth.mixture_set()
    f = synth.features(3000, 39)
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=8)
    got = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-tensor").score(f)
    e = rel(got, want)
    diag("gmm_tensor_c2", max_rel=e)
    assert e < RTOL


@pytest.mark.parametrize("dpm", [8, 16, 32, 1, 5, 64, 256])
def test_uniform_mixture_sizes(oracle, diag, dpm):
    msd = synth.mixture_set(dim=39, n_mixtures=20, densities_per_mixture=dpm, seed=dpm)
    f = synth.features(777, 39, seed=dpm)
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=4)
    got = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-tensor").score(f)
    e = rel(got, want)
    diag("gmm_tensor_uniform", dpm=dpm, max_rel=e)
    assert e < RTOL


@pytest.mark.parametrize("dim", [1, 7, 13, 39, 40, 45, 64])
def test_dimensions_and_ragged_mixtures(oracle, diag, dim):
    msd = synth.ragged_mixture_set(dim=dim, sizes=(1, 3, 16, 7, 32, 2, 200, 100, 9, 64, 255), seed=dim)
    f = synth.features(500, dim, seed=dim)
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=4)
    got = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-tensor").score(f)
    e = rel(got, want)
    diag("gmm_tensor_ragged", dim=dim, max_rel=e)
    assert e < RTOL


def test_offset_features_stay_accurate(oracle, diag):
    """Raw (un-normalised) cepstra sit far from the origin: centring on the model mean keeps the
    expanded quadratic form well conditioned."""
    msd = synth.mixture_set(dim=39, n_mixtures=32, densities_per_mixture=16, seed=9)
    off = np.linspace(-40, 60, 39).astype(np.float32)
    msd["means"] = msd["means"] + off
    f = synth.features(1000, 39, seed=9) + off
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=4)
    got = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-tensor").score(f)
    e = rel(got, want)
    diag("gmm_tensor_offset", max_rel=e)
    assert e < RTOL


def test_model_with_more_column_groups_than_sms(oracle, diag):
    """300 mixtures x 256 densities = 300 column blocks = 150 resident groups > 148 CTAs (several rounds)."""
    msd = synth.mixture_set(dim=8, n_mixtures=300, densities_per_mixture=256, seed=21)
    f = synth.features(300, 8, seed=21)
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=8)
    got = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-tensor").score(f)
    e = rel(got, want)
    diag("gmm_tensor_big_model", max_rel=e)
    assert e < RTOL


def test_frame_count_edges(oracle):
    msd = synth.mixture_set(dim=39, n_mixtures=16, densities_per_mixture=16, seed=2)
    sc = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-tensor")
    oms = oracle.MixtureSet(**msd)
    for T in (1, 127, 128, 129, 1000):
        f = synth.features(T, 39, seed=T)
        assert rel(sc.score(f), oracle.gmm_batch_float(oms, f)) < RTOL


def test_unsupported_shapes_fail_loudly():
    with pytest.raises(capi.RasrB200Error) as e:
        mm.GmmScorer(mm.MixtureSet.from_dict(synth.ragged_mixture_set(dim=8, sizes=(3, 0, 2))), "batch-tensor")
    assert e.value.status == -4
    with pytest.raises(capi.RasrB200Error) as e:
        mm.GmmScorer(mm.MixtureSet.from_dict(synth.ragged_mixture_set(dim=8, n_covariances=2)), "batch-tensor")
    assert e.value.status == -4


def test_full_size_c2_against_exact_kernel(diag):
    """100k frames: tensor scores vs the bit-exact CUDA-core kernel (itself pinned to the oracle)."""
    msd = synth.mixture_set()
    gms = mm.MixtureSet.from_dict(msd)
    f = synth.features(100000, 39)
    exact = mm.GmmScorer(gms, "batch-float").score(f)
    got = mm.GmmScorer(gms, "batch-tensor").score(f)
    e = rel(got, exact)
    diag("gmm_tensor_c2_full", max_rel=e)
    assert e < RTOL
