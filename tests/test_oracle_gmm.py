"""Oracle pins for the GMM scorers: closed-form KATs and an independent float64 model.  CPU only.
The reference has no unit test for any Mm scorer: parity unpinned by the
reference, pinned here by derivation from the cited source."""
import os

import numpy as np
import pytest

from rasr_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def f64_scores(msd, feats, mode):
    """-log p(x|m) under max approximation ('max') or exact mixture likelihood ('sum'), in float64."""
    dim = msd["dim"]
    mu = msd["means"].astype(np.float64)[msd["dens_mean"]]
    var = msd["variances"].astype(np.float64)[msd["dens_cov"]]
    x = feats.astype(np.float64)
    lognorm = dim * np.log(2 * np.pi) + np.log(var).sum(1)
    nm = msd["mix_offsets"].size - 1
    out = np.zeros((x.shape[0], nm))
    best = np.zeros((x.shape[0], nm), np.int64)
    for m in range(nm):
        e = slice(msd["mix_offsets"][m], msd["mix_offsets"][m + 1])
        d = msd["mix_density"][e]
        dist = (((x[:, None, :] - mu[d][None]) ** 2) / var[d][None]).sum(2)
        s = 0.5 * (dist + lognorm[d][None] - 2 * msd["mix_log_weight"][e][None])
        best[:, m] = s.argmin(1)
        if mode == "max":
            out[:, m] = s.min(1)
        else:
            mn = s.min(1)
            out[:, m] = mn - np.log(np.exp(mn[:, None] - s).sum(1))
    return out, best


def test_single_density_closed_form(oracle):
    """One mixture, one density, unit variance, weight 1: score = 0.5*(D*ln(2 pi) + |x-mu|^2)."""
    D = 8
    msd = dict(dim=D, mix_offsets=[0, 1], mix_density=[0], mix_log_weight=[0.0], dens_mean=[0], dens_cov=[0],
               means=np.arange(D, dtype=np.float32)[None], variances=np.ones((1, D), np.float32))
    ms = oracle.MixtureSet(**msd)
    x = np.zeros((1, D), np.float32)
    want = 0.5 * (D * np.log(2 * np.pi) + float((np.arange(D) ** 2).sum()))
    for fn in (oracle.gmm_batch_float, lambda m, f: oracle.gmm_diag_max(m, f)[0], lambda m, f: oracle.gmm_diag_sum(m, f)[0]):
        assert fn(ms, x)[0, 0] == pytest.approx(want, rel=1e-6)


@pytest.mark.parametrize("dim", [39, 40, 13, 7])
def test_batch_float_matches_f64(oracle, dim):
    msd = synth.mixture_set(dim=dim, n_mixtures=24, densities_per_mixture=16, seed=11)
    ms = oracle.MixtureSet(**msd)
    f = synth.features(200, dim, seed=3)
    want, _ = f64_scores(msd, f, "max")
    for fma in (True, False):
        got = oracle.gmm_batch_float(ms, f, use_fma=fma)
        np.testing.assert_allclose(got, want, rtol=3e-6)
    a = oracle.gmm_batch_float(ms, f, use_fma=True, threads=4)
    assert np.array_equal(a, oracle.gmm_batch_float(ms, f, use_fma=True))


@pytest.mark.parametrize("ncov", [1, 3])
def test_diag_max_and_sum_match_f64(oracle, ncov):
    msd = synth.ragged_mixture_set(dim=39, n_covariances=ncov, seed=21)
    ms = oracle.MixtureSet(**msd)
    f = synth.features(150, 39, seed=4)
    want, wbest = f64_scores(msd, f, "max")
    got, best = oracle.gmm_diag_max(ms, f)
    np.testing.assert_allclose(got, want, rtol=3e-6)
    # argmin may legitimately differ only on f32 near-ties
    bad = best != wbest
    if bad.any():
        t, m = np.nonzero(bad)
        assert np.all(np.abs(want[t, m] - got[t, m]) <= 3e-6 * np.abs(want[t, m]))
    swant, _ = f64_scores(msd, f, "sum")
    sgot, sbest = oracle.gmm_diag_sum(ms, f)
    np.testing.assert_allclose(sgot, swant, rtol=5e-6, atol=1e-5)
    assert np.array_equal(sbest, best) or (sbest != best).mean() < 1e-3
    # log-sum-exp is never worse than the max approximation
    assert np.all(sgot <= got + 1e-4)


def test_pooled_batch_equals_diag_max_to_rounding(oracle):
    """With one pooled covariance the two max scorers compute the same quantity in different order."""
    msd = synth.mixture_set(dim=39, n_mixtures=16, densities_per_mixture=16, seed=5)
    ms = oracle.MixtureSet(**msd)
    f = synth.features(100, 39, seed=6)
    a = oracle.gmm_batch_float(ms, f)
    b, _ = oracle.gmm_diag_max(ms, f)
    np.testing.assert_allclose(a, b, rtol=5e-6)


def test_batch_float_rejects_multiple_covariances(oracle):
    msd = synth.ragged_mixture_set(n_covariances=2)
    ms = oracle.MixtureSet(**msd)
    with pytest.raises(RuntimeError):
        oracle.gmm_batch_float(ms, synth.features(2, 39))


def test_scales(oracle):
    msd = synth.ragged_mixture_set(dim=39, n_covariances=2, seed=8)
    ms = oracle.MixtureSet(**msd)
    f = synth.features(20, 39, seed=9)
    # gaussian-scale g multiplies distance and log-norm by g (sqrt applied to isd, squared on the norm)
    s1, _ = oracle.gmm_diag_max(ms, f, mixture_weight_scale=1.0, gaussian_scale=1.0)
    zero_w = dict(msd)
    zero_w["mix_log_weight"] = np.zeros_like(msd["mix_log_weight"])
    msz = oracle.MixtureSet(**zero_w)
    a, _ = oracle.gmm_diag_max(msz, f, 1.0, 1.0)
    b, _ = oracle.gmm_diag_max(msz, f, 1.0, 4.0)
    np.testing.assert_allclose(b, 4.0 * a, rtol=2e-6)
    assert s1.shape == a.shape


# ---------------------------------------------------------------- Mm::BatchIntFeatureScorer (a12)

def int_model_numpy(msd):
    """Independent restatement of BatchIntFeatureScorer::init / quantizationScale in numpy, keeping the f32 / f64
    mixing of src/Mm/BatchFeatureScorer.cc:355-416 (every product below is a single rounding)."""
    f32 = np.float32
    dim = msd["dim"]
    isd = (f32(1) / np.sqrt(msd["variances"][0].astype(np.float32)).astype(np.float32)).astype(np.float32)
    divided = (msd["means"][msd["dens_mean"]].astype(np.float32) * isd[None]).astype(np.float32)
    interval = f32(2) * max(abs(f32(divided.min())), abs(f32(divided.max())))
    scale = f32(np.float64(f32(255)) / (1.25 * np.float64(interval)))
    scale_sq = f32(scale * scale)
    scale2 = f32(2.0 * np.float64(scale_sq))
    variance = (isd * scale).astype(np.float32)
    lognorm = f32(dim * np.log(2 * np.pi) + np.log(np.abs(msd["variances"][0].astype(np.float64))).sum())
    lognorm_factor = f32(lognorm * scale_sq)

    def quant(v):
        r = np.where(v >= 0, np.floor(v.astype(np.float64) + 0.5), -np.floor(-v.astype(np.float64) + 0.5))  # roundf
        return np.clip(r.astype(np.int64) + 128, 0, 255).astype(np.uint8)

    d = msd["mix_density"]
    means = quant((msd["means"][msd["dens_mean"][d]].astype(np.float32) * variance[None]).astype(np.float32))
    consts = np.trunc(np.float64(lognorm_factor) - np.float64(scale2) * msd["mix_log_weight"]).astype(np.int32)
    return dict(means=means, consts=consts, variance=variance, scale=scale2, quant=quant)


def int_scores_numpy(msd, feats):
    m = int_model_numpy(msd)
    xq = m["quant"]((feats.astype(np.float32) * m["variance"][None]).astype(np.float32)).astype(np.int64)
    nm = msd["mix_offsets"].size - 1
    out = np.zeros((feats.shape[0], nm), np.float32)
    for k in range(nm):
        e = slice(msd["mix_offsets"][k], msd["mix_offsets"][k + 1])
        if e.start == e.stop:
            out[:, k] = np.float32(2147483647) / np.float32(m["scale"])
            continue
        dist = ((xq[:, None, :] - m["means"][e].astype(np.int64)[None]) ** 2).sum(2) + m["consts"][e][None]
        out[:, k] = dist.min(1).astype(np.int32).astype(np.float32) / np.float32(m["scale"])
    return out


@pytest.mark.parametrize("dim", [39, 16, 7, 48])
def test_batch_int_matches_numpy_restatement(oracle, dim):
    msd = synth.mixture_set(dim=dim, n_mixtures=24, densities_per_mixture=16, seed=11)
    ms = oracle.MixtureSet(**msd)
    f = synth.features(200, dim, seed=3)
    model = oracle.gmm_batch_int_model(ms)
    ref = int_model_numpy(msd)
    assert model["padded"] == (dim + 15) // 16 * 16
    assert model["scale"] == ref["scale"]
    assert np.array_equal(model["variance"][:dim], ref["variance"])
    assert np.array_equal(model["means"][:, :dim], ref["means"]) and not model["means"][:, dim:].any()
    assert np.array_equal(model["consts"], ref["consts"])
    got = oracle.gmm_batch_int(ms, f)
    assert np.array_equal(got, int_scores_numpy(msd, f))  # integer path: bit-exact
    assert np.array_equal(got, oracle.gmm_batch_int(ms, f, threads=4))


def test_batch_int_quantisation_kat(oracle):
    """quantize() = clip(round-half-away(x) + 128, 0, 255) (src/Mm/Utilities.hh:190-202): one density at the
    origin, unit variance => scale = 255 / (1.25 * 2 * max|mu|); features far away clip at 0 / 255."""
    D = 4
    msd = dict(dim=D, mix_offsets=[0, 1], mix_density=[0], mix_log_weight=[0.0], dens_mean=[0], dens_cov=[0],
               means=np.array([[2.0, -2.0, 0.5, 0.0]], np.float32), variances=np.ones((1, D), np.float32))
    ms = oracle.MixtureSet(**msd)
    m = oracle.gmm_batch_int_model(ms)
    scale = np.float32(255.0 / (1.25 * 4.0))                       # 51
    assert m["scale"] == np.float32(2.0 * scale * scale)
    assert list(m["means"][0, :D]) == [230, 26, 154, 128]           # round(2*51)+128, round(-102)+128, round(25.5)+128
    x = np.array([[100.0, -100.0, 0.0, 0.0]], np.float32)           # clips to 255 / 0
    got = oracle.gmm_batch_int(ms, x)[0, 0]
    dist = (255 - 230) ** 2 + (0 - 26) ** 2 + (128 - 154) ** 2 + 0
    c = int(np.trunc(np.float64(np.float32(np.float32(D * np.log(2 * np.pi)) * np.float32(scale * scale)))))
    assert got == np.float32(dist + c) / m["scale"]


def test_batch_int_close_to_float_scorer(oracle):
    """The int scorer approximates the float one: same argmin mixtures mostly, scores within quantisation noise."""
    msd = synth.mixture_set(dim=39, n_mixtures=32, densities_per_mixture=16, seed=5)
    ms = oracle.MixtureSet(**msd)
    f = synth.features(300, 39, seed=9)
    a, b = oracle.gmm_batch_int(ms, f), oracle.gmm_batch_float(ms, f)
    assert np.median(np.abs(a - b) / b) < 0.02


def test_batch_int_ragged_and_empty(oracle):
    msd = synth.mixture_set(dim=39, n_mixtures=5, densities_per_mixture=3, seed=2)
    msd["mix_offsets"] = np.array([0, 1, 1, 6, 7, 15], np.uint32)  # sizes 1, 0, 5, 1, 8
    ms = oracle.MixtureSet(**msd)
    f = synth.features(50, 39, seed=4)
    got = oracle.gmm_batch_int(ms, f)
    assert np.array_equal(got, int_scores_numpy(msd, f))
    assert np.all(got[:, 1] == np.float32(2147483647) / np.float32(oracle.gmm_batch_int_model(ms)["scale"]))


# ---- density preselection ("preselection-batch-float") ---------------------------------------------------------------

def test_glibc_rand_restatement_matches_libc():
    """the clustering is initialised from rand() after srand(1) (src/Mm/DensityClustering.tcc:62-75); the library
    restates glibc's generator instead of calling it (no side effect on the host program's random state)"""
    import ctypes
    from rasr_b200 import capi
    libc = ctypes.CDLL(None)
    for seed in (1, 2, 12345):
        libc.srand(seed)
        want = [libc.rand() for _ in range(1000)]
        got = np.zeros(1000, np.int32)
        capi.lib().rb_test_glibc_rand(seed, 1000, capi.ptr(got))
        assert list(got) == want, seed


def test_preselection_oracle_properties(oracle):
    msd = synth.mixture_set()
    ms = oracle.MixtureSet(**msd)
    f = synth.features(300, 39)
    full = oracle.gmm_batch_float(ms, f)
    sc, cluster_of, means = oracle.gmm_preselect_float(ms, f)
    assert means.shape == (256, 40) and cluster_of.max() < 256
    hit = sc != np.float32(40000)
    # a preselected score is the minimum over a subset of the densities: never below the full minimum, and equal to
    # it whenever the best density's cluster was selected
    assert (sc[hit] >= full[hit]).all() and (sc == full).mean() > 0.3 and (~hit).mean() < 0.2
    # selecting every cluster is the plain batch scorer
    assert np.array_equal(oracle.gmm_preselect_float(ms, f, select=256)[0], full)
    # one cluster, one iteration: the cluster mean is the f64 mean of all scaled density means
    _, c1, m1 = oracle.gmm_preselect_float(ms, f[:2], clusters=1, select=1, iterations=1)
    isd = (1 / np.sqrt(msd["variances"][0].astype(np.float32))).astype(np.float32)
    scaled = (msd["means"].astype(np.float32) * isd).astype(np.float32)
    assert (c1 == 0).all() and np.array_equal(m1[0, :39], scaled.astype(np.float64).mean(0).astype(np.float32))


def test_std_sort_restatement_matches_libstdcxx(oracle):
    """oracle/std_sort_restated.h against the real std::sort on (key, index) pairs compared by key only: with ties
    the resulting index permutation is whatever introsort does, and the restatement has to do the same
    (Mm::DensityClustering::selectClusters, src/Mm/DensityClustering.tcc:164-189)"""
    rng = np.random.default_rng(5)
    cases = [np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(16, np.int32), np.zeros(17, np.int32),
             np.zeros(256, np.int32), np.arange(300, dtype=np.int32), np.arange(300, dtype=np.int32)[::-1],
             np.repeat(np.arange(8, dtype=np.int32), 40)]
    for n in (2, 15, 16, 17, 33, 100, 256, 1000, 4097):
        for hi in (2, 5, 50, 1 << 30):
            cases.append(rng.integers(0, hi, n).astype(np.int32))
    # organ-pipe and sawtooth inputs drive introsort towards its depth limit (heap sort branch)
    cases.append(np.concatenate([np.arange(2000), np.arange(2000)[::-1]]).astype(np.int32))
    cases.append((np.arange(5000) % 7).astype(np.int32))
    # McIlroy's adversary against this libstdc++: quadratic partitioning, so the depth limit (heap sort) is reached
    for n in (200, 256, 3000):
        killer = oracle.sort_killer(n)
        assert np.array_equal(np.sort(killer), np.arange(n))
        cases.append(killer)
        cases.append((killer // 3).astype(np.int32))  # the same shape with ties
    from rasr_b200 import capi
    heap_before = oracle.sort_heap_calls()
    for k in cases:
        k = np.ascontiguousarray(k)
        want, got = oracle.sort_pairs(k, False), oracle.sort_pairs(k, True)
        assert np.array_equal(want, got), (k.size, k[:8])
        # the library's own restatement (rasr_b200/csrc/introsort.cuh, host instantiation of the code the
        # preselection kernel runs per frame)
        dev = np.zeros(k.size, np.int32)
        capi.lib().rb_test_introsort(capi.ptr(k), int(k.size), capi.ptr(dev))
        assert np.array_equal(want, dev), (k.size, k[:8])
        assert np.array_equal(np.sort(want), np.arange(k.size)) and (np.diff(k[want]) >= 0).all()
    assert oracle.sort_heap_calls() > heap_before, "no case reached the heap sort branch"


def test_preselection_int_oracle_properties(oracle):
    """Mm::BatchPreselectionIntFeatureScorer (src/Mm/BatchFeatureScorer.cc:514-577): subset minimum of the int scorer"""
    msd = synth.mixture_set()
    ms = oracle.MixtureSet(**msd)
    f = synth.features(200, 39)
    full = oracle.gmm_batch_int(ms, f)
    sc, cluster_of = oracle.gmm_preselect_int(ms, f)
    none = np.float32(2147483647) / np.float32(oracle.gmm_batch_int_model(ms)["scale"])
    hit = sc != none
    assert cluster_of.max() < 256 and (~hit).mean() < 0.3
    assert (sc[hit] >= full[hit]).all() and (sc == full).mean() > 0.3
    # selecting every cluster is the plain int scorer
    assert np.array_equal(oracle.gmm_preselect_int(ms, f, select=256)[0], full)
    # the restated sort picks the same clusters as std::sort, ties included (few clusters, coarse u8 distances)
    for clusters, select in ((256, 32), (64, 8), (16, 3)):
        a = oracle.gmm_preselect_int(ms, f, clusters=clusters, select=select, restated_sort=False)
        b = oracle.gmm_preselect_int(ms, f, clusters=clusters, select=select, restated_sort=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_std_sort_restatements_property(oracle):
    """random sizes and key ranges (few distinct keys = many ties): std::sort, the oracle's restatement and the
    library's introsort.cuh produce the same permutation"""
    from hypothesis import given, settings, strategies as st
    from rasr_b200 import capi

    @settings(max_examples=300, deadline=None)
    @given(st.integers(0, 700), st.sampled_from([1, 2, 3, 7, 40, 1000, 1 << 30]), st.integers(0, 2 ** 32 - 1),
           st.sampled_from(["random", "sorted", "reversed", "sawtooth"]))
    def check(n, hi, seed, shape):
        k = np.random.default_rng(seed).integers(0, hi, n).astype(np.int32)
        if shape == "sorted":
            k.sort()
        elif shape == "reversed":
            k = np.ascontiguousarray(np.sort(k)[::-1])
        elif shape == "sawtooth":
            k = np.ascontiguousarray(np.concatenate([np.sort(k[: n // 2]), np.sort(k[n // 2:])]))
        want = oracle.sort_pairs(k, False)
        lib = np.zeros(n, np.int32)
        capi.lib().rb_test_introsort(capi.ptr(k), int(n), capi.ptr(lib))
        assert np.array_equal(want, oracle.sort_pairs(k, True)) and np.array_equal(want, lib)

    check()


