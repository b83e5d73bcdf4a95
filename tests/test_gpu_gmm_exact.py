"""RB_GMM_BATCH_FLOAT on large batches: tensor-core screening + exact evaluation of the surviving densities
(gmm_tensor.cu EpiGmmScreen + gmm.cu gmm_refine_kernel) must give the SAME BITS as the direct kernel and
as the CPU oracle of Mm::BatchFloatFeatureScorer (src/Mm/BatchFeatureScorer.cc:164-253) -- on ordinary data, on models
built to produce ties and near-ties inside a mixture, on ragged mixtures and on frames with non-finite values."""
import os

import numpy as np
import pytest

from rasr_b200 import mm, synth

pytestmark = pytest.mark.gpu


def scorer(msd, route, contraction=True, mode="batch-float"):
    """route: 'two-pass' (every call, whatever its size), 'direct' (the SIMT kernel only)"""
    old = os.environ.get("RB_GMM_EXACT")
    os.environ["RB_GMM_EXACT"] = "2" if route == "two-pass" else "0"
    try:
        return mm.GmmScorer(mm.MixtureSet.from_dict(msd), mode, contraction=contraction)
    finally:
        if old is None:
            del os.environ["RB_GMM_EXACT"]
        else:
            os.environ["RB_GMM_EXACT"] = old


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def ragged_set(dim, sizes, seed):
    """mixtures of the given sizes over their own densities, one pooled covariance"""
    rng = np.random.default_rng(seed)
    n = int(sum(sizes))
    offs = np.zeros(len(sizes) + 1, np.uint32)
    offs[1:] = np.cumsum(sizes)
    return dict(dim=dim, mix_offsets=offs, mix_density=rng.permutation(n).astype(np.uint32),
                mix_log_weight=np.concatenate([np.log(rng.dirichlet(np.ones(s))) for s in sizes]).astype(np.float64),
                dens_mean=rng.permutation(n).astype(np.uint32), dens_cov=np.zeros(n, np.uint32),
                means=rng.standard_normal((n, dim)).astype(np.float32),
                variances=rng.uniform(0.5, 2.0, (1, dim)).astype(np.float32))


@pytest.mark.parametrize("contraction", [True, False])
def test_c2_shape_equals_oracle_and_direct_kernel(oracle, diag, contraction):
    msd = synth.mixture_set()
    f = synth.features(6000, 39, seed=11)
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, use_fma=contraction, threads=8)
    two = scorer(msd, "two-pass", contraction).score(f)
    one = scorer(msd, "direct", contraction).score(f)
    diag("gmm_exact_c2", contraction=contraction, n_diff_oracle=int((bits(two) != bits(want)).sum()),
         n_diff_direct=int((bits(two) != bits(one)).sum()), total=two.size)
    assert np.array_equal(bits(one), bits(want))
    assert np.array_equal(bits(two), bits(want))


@pytest.mark.parametrize("dpm,dim", [(8, 39), (32, 39), (16, 26), (8, 32), (32, 30)])
def test_other_uniform_mixture_sizes(oracle, dpm, dim):
    """8 / 16 / 32 densities per mixture; 68 mixtures: the model's last 256-column block is only partly filled"""
    msd = synth.mixture_set(dim=dim, n_mixtures=68, densities_per_mixture=dpm, seed=dpm + dim)
    f = synth.features(3000, dim, seed=dpm)
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=8)
    assert np.array_equal(bits(scorer(msd, "two-pass").score(f)), bits(want))


@pytest.mark.parametrize("dim", [1, 7, 8, 13, 33, 40, 45, 64])
def test_dimensions(oracle, dim):
    msd = synth.mixture_set(dim=dim, n_mixtures=12, densities_per_mixture=16, seed=dim)
    f = synth.features(1500, dim, seed=dim)
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=8)
    assert np.array_equal(bits(scorer(msd, "two-pass").score(f)), bits(want))


def test_ragged_mixtures(oracle):
    """unequal mixture sizes 1..32 (the sequential screening epilogue; mixtures straddle 32-column chunks)"""
    rng = np.random.default_rng(5)
    sizes = [int(s) for s in rng.integers(1, 33, 40)]
    sizes[3], sizes[17] = 32, 1
    msd = ragged_set(39, sizes, 9)
    f = synth.features(2500, 39, seed=3)
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=8)
    assert np.array_equal(bits(scorer(msd, "two-pass").score(f)), bits(want))


def test_ties_and_near_ties_inside_a_mixture(oracle, diag):
    """densities that share a mean (exact ties), differ by one ulp in one component, or differ only in their weight by
    1e-7: the winner in the reference's arithmetic is decided far below the accuracy of the screening product, so the
    candidate sets must carry every one of them"""
    msd = synth.mixture_set(dim=39, n_mixtures=32, densities_per_mixture=16, seed=21)
    means = msd["means"].reshape(-1, 39)
    lw = msd["mix_log_weight"]
    for m in range(32):
        base = m * 16
        means[base + 1] = means[base]                      # identical mean, different weight
        means[base + 2] = means[base]
        means[base + 2, 5] = np.nextafter(means[base, 5], np.float32(10))  # one ulp away
        means[base + 3] = means[base]
        lw[base + 3] = lw[base] + 1e-7                     # identical mean, weights 1e-7 apart
        means[base + 4] = means[base] + np.float32(1e-6)   # a whisker away in every component
        lw[base + 1] = lw[base]                            # exact tie
    f = synth.features(4000, 39, seed=8)
    # frames sitting exactly on a mean and exactly between two means
    f[0] = means[0] * np.sqrt(msd["variances"][0])
    f[1] = 0.5 * (means[16] + means[17]) * np.sqrt(msd["variances"][0])
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=8)
    got = scorer(msd, "two-pass").score(f)
    diag("gmm_exact_ties", n_diff=int((bits(got) != bits(want)).sum()), total=got.size)
    assert np.array_equal(bits(got), bits(want))


def test_non_finite_and_huge_frames_follow_the_direct_kernel():
    """frames the screening cannot trust (NaN, Inf, values beyond the fp16 range of the split operands) are scored over
    every density in the reference's order: bit patterns equal the direct kernel's, NaNs included"""
    msd = synth.mixture_set(dim=39, n_mixtures=16, densities_per_mixture=16, seed=4)
    f = synth.features(3000, 39, seed=6)
    f[5, 3] = np.nan
    f[6, :] = np.inf
    f[7, 0] = -np.inf
    f[8, 10] = 3.0e38
    f[9, :] = 1.0e6
    f[10, 38] = 7.0e4
    f[2999, 0] = np.nan
    one = scorer(msd, "direct").score(f)
    two = scorer(msd, "two-pass").score(f)
    assert np.array_equal(bits(one), bits(two))
    assert np.isfinite(two[11:2999]).all()


def test_small_calls_and_host_slabs(oracle):
    """the host-buffer call cuts a segment into slabs of growing size, and calls of any size (one frame included) take
    the two-pass route -- one matrix, the oracle's bits"""
    msd = synth.mixture_set(dim=39, n_mixtures=64, densities_per_mixture=16, seed=12)
    f = synth.features(9000, 39, seed=13)
    want = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f, threads=8)
    sc = mm.GmmScorer(mm.MixtureSet.from_dict(msd))
    assert np.array_equal(bits(sc.score(f)), bits(want))
    for T in (1, 255, 2047, 2048, 2049):
        assert np.array_equal(bits(sc.score(f[:T])), bits(want[:T]))


def test_models_the_route_does_not_cover_fall_back():
    """an empty mixture, a mixture of more than 32 densities, a mixture count that is not a multiple of 4: the direct
    kernel serves them (same call, same scores as before)"""
    for sizes in ([4, 0, 3, 5], [40, 3, 3, 3], [3, 4, 5]):
        msd = ragged_set(20, sizes, 3)
        f = synth.features(2500, 20, seed=2)
        assert np.array_equal(bits(scorer(msd, "two-pass").score(f)), bits(scorer(msd, "direct").score(f)))


def test_screening_error_has_margin(diag):
    """The candidate threshold assumes the split-precision product is within 2^-17 Q of the exact cross term, Q = |xc|^2 +
    |mu_c|^2 + |c + |mu_c|^2| (gmm_tensor.cu).  Measured here through RB_GMM_BATCH_TENSOR (score = 0.5 (|xc|^2 + min
    acc)): the distance to the exact scorer's result, over the smallest Q of the mixture, must stay below 2^-19."""
    msd = synth.mixture_set()
    f = synth.features(20000, 39, seed=31)
    exact = scorer(msd, "direct").score(f).astype(np.float64)
    approx = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-tensor").score(f).astype(np.float64)
    isd = 1.0 / np.sqrt(msd["variances"][0].astype(np.float64))
    mu = msd["means"].astype(np.float64) * isd
    centre = mu.mean(axis=0)
    xc2 = (((f.astype(np.float64) * isd) - centre) ** 2).sum(axis=1)            # [T]
    mu2 = ((mu - centre) ** 2).sum(axis=1)                                      # [densities]
    log_norm = 39 * np.log(2 * np.pi) + np.log(msd["variances"][0].astype(np.float64)).sum()
    c = log_norm - 2 * msd["mix_log_weight"]
    qmu = (mu2 + np.abs(c + mu2)).reshape(256, 16).min(axis=1)                  # [mixtures]
    ratio = 2 * np.abs(approx - exact) / (xc2[:, None] + qmu[None, :])
    diag("gmm_screen_error", max_ratio_log2=float(np.log2(ratio.max())), mean_ratio_log2=float(np.log2(ratio.mean())))
    assert ratio.max() < 2.0 ** -19


def test_full_c2_batch_equals_direct_kernel(diag):
    """BASELINE config C2 at full size: all 25.6 M scores of the two routes agree bit for bit"""
    msd = synth.mixture_set()
    f = synth.features(100000, 39, seed=2024)
    one = scorer(msd, "direct").score(f)
    two = mm.GmmScorer(mm.MixtureSet.from_dict(msd)).score(f)
    n_diff = int((bits(one) != bits(two)).sum())
    diag("gmm_exact_full_c2", n_diff=n_diff, total=one.size)
    assert n_diff == 0


# ---- RB_GMM_DIAG_MAX (Mm::GaussDiagonalMaximumFeatureScorer, per-density covariance, best density reported) through the
# ---- same screening + refinement route: scores AND density indices must be the direct kernel's and the oracle's
def diag_scorer(msd, route, contraction=True):
    return scorer(msd, route, contraction, mode="diagonal-maximum")


@pytest.mark.parametrize("contraction", [True, False])
@pytest.mark.parametrize("n_cov", [1, 3])
def test_diag_max_c2_shape(oracle, diag, contraction, n_cov):
    msd = synth.mixture_set(n_covariances=n_cov)
    f = synth.features(5000, 39, seed=17)
    want, want_b = oracle.gmm_diag_max(oracle.MixtureSet(**msd), f, use_fma=contraction)
    two, two_b = diag_scorer(msd, "two-pass", contraction).score(f, want_density=True)
    one, one_b = diag_scorer(msd, "direct", contraction).score(f, want_density=True)
    diag("gmm_exact_diag_c2", contraction=contraction, n_cov=n_cov, n_diff=int((bits(two) != bits(want)).sum()),
         n_diff_idx=int((two_b != want_b).sum()), total=two.size)
    assert np.array_equal(bits(one), bits(want)) and np.array_equal(one_b, want_b)
    assert np.array_equal(bits(two), bits(want)) and np.array_equal(two_b, want_b)
    # without the index output
    assert np.array_equal(bits(diag_scorer(msd, "two-pass", contraction).score(f)), bits(want))


@pytest.mark.parametrize("dim", [2, 7, 8, 13, 24, 33, 40])
def test_diag_max_dimensions_and_scales(oracle, dim):
    """tail dimensions (dim % 4 != 0) are summed sequentially in the reference; mixture-weight and Gaussian scales"""
    msd = synth.mixture_set(dim=dim, n_mixtures=12, densities_per_mixture=16, seed=dim, n_covariances=2)
    f = synth.features(2500, dim, seed=dim)
    want, want_b = oracle.gmm_diag_max(oracle.MixtureSet(**msd), f, mixture_weight_scale=0.7, gaussian_scale=1.3)
    old = os.environ.get("RB_GMM_EXACT")
    os.environ["RB_GMM_EXACT"] = "2"
    try:
        sc = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "diagonal-maximum", mixture_weight_scale=0.7, gaussian_scale=1.3)
    finally:
        if old is None:
            del os.environ["RB_GMM_EXACT"]
        else:
            os.environ["RB_GMM_EXACT"] = old
    got, got_b = sc.score(f, want_density=True)
    assert np.array_equal(bits(got), bits(want)) and np.array_equal(got_b, want_b)


def test_diag_max_ragged_ties_and_non_finite(oracle):
    rng = np.random.default_rng(11)
    sizes = [int(s) for s in rng.integers(1, 33, 24)]
    msd = ragged_set(39, sizes, 4)
    msd["variances"] = rng.uniform(0.3, 3.0, (4, 39)).astype(np.float32)
    msd["dens_cov"] = rng.integers(0, 4, len(msd["dens_cov"])).astype(np.uint32)
    means = msd["means"].reshape(-1, 39)
    # duplicate densities inside the first mixtures: exact ties decided by the reference's order of comparison
    offs = msd["mix_offsets"]
    for m in range(len(sizes)):
        if sizes[m] >= 3:
            d0, d1, d2 = (int(msd["mix_density"][offs[m] + i]) for i in range(3))
            means[msd["dens_mean"][d1]] = means[msd["dens_mean"][d0]]
            msd["dens_cov"][d1] = msd["dens_cov"][d0]
            msd["mix_log_weight"][offs[m] + 1] = msd["mix_log_weight"][offs[m]]
            means[msd["dens_mean"][d2]] = means[msd["dens_mean"][d0]] + np.float32(1e-6)
    f = synth.features(3000, 39, seed=9)
    f[3, 2] = np.nan
    f[4, :] = np.inf
    f[5, 7] = 1.0e5
    f[6, :] = 300.0
    one, one_b = diag_scorer(msd, "direct").score(f, want_density=True)
    two, two_b = diag_scorer(msd, "two-pass").score(f, want_density=True)
    assert np.array_equal(bits(one), bits(two)) and np.array_equal(one_b, two_b)
    ok = np.ones(3000, bool)
    ok[3:7] = False
    want, want_b = oracle.gmm_diag_max(oracle.MixtureSet(**msd), f[ok])
    assert np.array_equal(bits(two[ok]), bits(want)) and np.array_equal(two_b[ok], want_b)


def test_diag_max_full_size_equals_direct_kernel(diag):
    msd = synth.mixture_set(n_covariances=2)
    f = synth.features(100000, 39, seed=5)
    one, one_b = diag_scorer(msd, "direct").score(f, want_density=True)
    two, two_b = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "diagonal-maximum").score(f, want_density=True)
    n_diff, n_idx = int((bits(one) != bits(two)).sum()), int((one_b != two_b).sum())
    diag("gmm_exact_diag_full", n_diff=n_diff, n_diff_idx=n_idx, total=one.size)
    assert n_diff == 0 and n_idx == 0
