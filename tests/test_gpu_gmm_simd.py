"""GPU parity of RB_GMM_SIMD_DIAG_MAX (Mm::SimdGaussDiagonalMaximumFeatureScorer, "SIMD-diagonal-maximum") against the
CPU oracle, through the C ABI.  Integer path: scores AND best-density indices must be BIT-IDENTICAL.  The oracle is
pinned to the reference's own scorer object code in tests/test_ref_parity.py (test_c2_model_simd_diagonal_maximum,
test_ragged_models_and_other_dimensions)."""
import numpy as np
import pytest

from rasr_b200 import capi, mm, synth

pytestmark = pytest.mark.gpu


def both(oracle, msd):
    return oracle.MixtureSet(**msd), mm.MixtureSet.from_dict(msd)


def check(oracle, msd, f, threads=4):
    oms, gms = both(oracle, msd)
    want, wbest = oracle.gmm_simd_diag_max(oms, f, threads=threads, want_best=True)
    sc = mm.GmmScorer(gms, "SIMD-diagonal-maximum")
    got, best = sc.score(f, want_density=True)
    assert np.array_equal(got, want), "%d of %d scores differ" % ((got != want).sum(), got.size)
    assert np.array_equal(best, wbest)
    assert np.array_equal(sc.score(f), want)  # without the density output
    return got


def test_c2_shape_bit_exact(oracle, diag):
    """C2 geometry (39-dim, 256 mixtures x 16 densities, pooled covariance): quantised features in registers"""
    msd = synth.mixture_set()
    got = check(oracle, msd, synth.features(3000, 39), threads=8)
    diag("gmm_simd_c2", frames=3000, mean_score=float(got.mean()))


@pytest.mark.parametrize("dim", [1, 7, 16, 17, 33, 39, 48, 64])
def test_dimensions(oracle, dim):
    msd = synth.mixture_set(dim=dim, n_mixtures=12, densities_per_mixture=5, seed=dim)
    check(oracle, msd, synth.features(700, dim, seed=dim))


@pytest.mark.parametrize("T", [1, 2, 255, 256, 257, 511, 512, 513, 1500])
def test_ragged_frame_counts(oracle, T):
    msd = synth.mixture_set(dim=39, n_mixtures=8, densities_per_mixture=16, seed=3)
    check(oracle, msd, synth.features(T, 39, seed=T))


@pytest.mark.parametrize("sizes", [(1, 3, 16, 7, 32, 2), (3, 0, 5, 1), (9, 8, 7, 24, 0, 0, 1, 40), (64,) * 5, (2000, 3)])
def test_ragged_and_empty_mixtures(oracle, sizes):
    """mixtures of unequal size, mixtures without densities (INT_MAX and density 0xffffffff as in the reference), a
    mixture that fills a shared-memory group on its own, mixture counts that are not multiples of 4"""
    msd = synth.ragged_mixture_set(dim=39, sizes=sizes, seed=len(sizes))
    check(oracle, msd, synth.features(300, 39, seed=8))


@pytest.mark.parametrize("n_cov", [2, 3, 7, 40])
def test_several_covariances(oracle, n_cov):
    """one quantised copy of the feature vector per covariance; permuted density / mean tables"""
    rng = np.random.default_rng(n_cov)
    sizes = [5, 1, 0, 16, 9, 33, 2, 7]
    n_dens = sum(sizes)
    dim = 24
    msd = dict(dim=dim, mix_offsets=np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32),
               mix_density=rng.permutation(n_dens).astype(np.uint32),
               mix_log_weight=np.log(rng.uniform(0.05, 1.0, n_dens)),
               dens_mean=rng.permutation(n_dens).astype(np.uint32),
               dens_cov=rng.integers(0, n_cov, n_dens).astype(np.uint32),
               means=rng.standard_normal((n_dens, dim)).astype(np.float32),
               variances=rng.uniform(0.3, 3.0, (n_cov, dim)).astype(np.float32))
    check(oracle, msd, synth.features(900, dim, seed=n_cov + 1))


def test_clipping_and_ties(oracle):
    """features far outside the quantisation interval clip at 0 / 255 (and +-1e30, NaN go through the reference's
    float -> int conversion); repeated densities tie exactly: the first one must be reported"""
    msd = synth.mixture_set(dim=39, n_mixtures=16, densities_per_mixture=8, seed=21)
    msd["dens_mean"] = (msd["dens_mean"] // 2 * 2).astype(np.uint32)  # pairs of densities share a mean
    lw = msd["mix_log_weight"].reshape(-1, 2)
    lw[:, 1] = lw[:, 0]
    msd["mix_log_weight"] = lw.reshape(-1)
    f = synth.features(256, 39, seed=1, scale=40.0)
    f[0, :] = 1e30
    f[1, :] = -1e30
    f[2, 3] = np.nan
    check(oracle, msd, f)


def test_rejects_what_it_does_not_cover():
    msd = synth.mixture_set(dim=65, n_mixtures=4, densities_per_mixture=2, seed=1)
    with pytest.raises(capi.RasrB200Error):
        mm.GmmScorer(mm.MixtureSet.from_dict(msd), "SIMD-diagonal-maximum")
