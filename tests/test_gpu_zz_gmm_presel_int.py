"""GPU parity of RB_GMM_BATCH_PRESELECT_INT (Mm::BatchPreselectionIntFeatureScorer, "preselection-batch-int") against
the CPU oracle, through the C ABI.  Integer work: the clustering, the per-frame cluster choice -- ties at the selection
boundary resolved like std::sort does -- and every score must be BIT-IDENTICAL.  (File name: runs after the other
scorers' tests.)"""
import numpy as np
import pytest

from rasr_b200 import capi, mm, synth

pytestmark = pytest.mark.gpu


def both(oracle, msd):
    return oracle.MixtureSet(**msd), mm.MixtureSet.from_dict(msd)


def test_c2_shape_bit_exact(oracle, diag):
    msd = synth.mixture_set()
    oms, gms = both(oracle, msd)
    f = synth.features(2000, 39)
    want, cl = oracle.gmm_preselect_int(oms, f)
    sc = mm.GmmScorer(gms, "preselection-batch-int")
    got_cl, got_means = sc.clustering()
    assert np.array_equal(got_cl, cl) and got_means.shape == (256, 48)
    assert np.array_equal(got_means, np.floor(got_means)) and got_means.min() >= 0 and got_means.max() <= 255
    got = sc.score(f)
    none = np.float32(2147483647) / np.float32(oracle.gmm_batch_int_model(oms)["scale"])
    diag("gmm_presel_int_c2", n_diff=int((got != want).sum()), total=got.size, none_frac=float((want == none).mean()))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("clusters,select,iterations", [(64, 8, 2), (256, 256, 5), (7, 1, 0), (200, 31, 3), (16, 3, 5),
                                                        (256, 17, 1)])
def test_other_clustering_parameters_and_ties(oracle, clusters, select, iterations):
    """few clusters and coarse u8 distances: equal distances at the selection boundary are frequent"""
    msd = synth.mixture_set(dim=39, n_mixtures=64, densities_per_mixture=16, seed=5)
    oms, gms = both(oracle, msd)
    f = synth.features(600, 39, seed=6)
    want, cl = oracle.gmm_preselect_int(oms, f, clusters=clusters, select=select, iterations=iterations)
    sc = mm.GmmScorer(gms, "preselection-batch-int")
    sc.configure_preselection(clusters, select, iterations)
    assert np.array_equal(sc.clustering()[0], cl)
    assert np.array_equal(sc.score(f), want)
    if select == clusters:  # no preselection at all: the plain int scorer
        assert np.array_equal(want, mm.GmmScorer(gms, "batch-int").score(f))


@pytest.mark.parametrize("dim,n_mix,per_mix", [(9, 5, 3), (16, 40, 2), (45, 30, 9), (64, 3, 1)])
def test_small_models_and_dimensions(oracle, dim, n_mix, per_mix):
    msd = synth.mixture_set(dim=dim, n_mixtures=n_mix, densities_per_mixture=per_mix, seed=dim)
    oms, gms = both(oracle, msd)
    f = synth.features(300, dim, seed=dim)
    n_dens = n_mix * per_mix
    want, cl = oracle.gmm_preselect_int(oms, f, select=min(32, n_dens))
    sc = mm.GmmScorer(gms, "preselection-batch-int")
    assert sc.clustering()[1].shape[0] == min(256, n_dens)
    assert np.array_equal(sc.clustering()[0], cl) and np.array_equal(sc.score(f), want)


def test_constant_features_all_distances_tie(oracle):
    """identical density means: every cluster is at the same distance, the choice is purely the sort's permutation"""
    msd = synth.mixture_set(dim=12, n_mixtures=40, densities_per_mixture=8, seed=3)
    msd["means"] = np.repeat(msd["means"][:4], (msd["means"].shape[0] + 3) // 4, axis=0)[:msd["means"].shape[0]].copy()
    oms, gms = both(oracle, msd)
    f = synth.features(200, 12, seed=4)
    want, cl = oracle.gmm_preselect_int(oms, f, clusters=100, select=13)
    sc = mm.GmmScorer(gms, "preselection-batch-int")
    sc.configure_preselection(100, 13, 5)
    assert np.array_equal(sc.clustering()[0], cl) and np.array_equal(sc.score(f), want)


def test_full_size_on_the_device(oracle, diag):
    """BASELINE C2 at full size (100k frames) through the device-pointer entry point; sampled frames bit-identical"""
    import torch

    msd = synth.mixture_set()
    oms, gms = both(oracle, msd)
    T = 100000
    f = synth.features(T, 39)
    sc = mm.GmmScorer(gms, "preselection-batch-int")
    d_in = torch.from_numpy(f).cuda()
    d_out = torch.empty((T, 256), dtype=torch.float32, device="cuda")
    sc.score_dev(d_in, T, d_out)
    torch.cuda.synchronize()
    sc.score_dev(d_in, T, d_out)  # a second call on the handle's own stream (timings: bench.py --workload ...)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    idx = np.random.default_rng(0).choice(T, 1500, replace=False)
    idx.sort()
    want = oracle.gmm_preselect_int(oms, f[idx])[0]
    diag("gmm_presel_int_100k", n_diff=int((got[idx] != want).sum()))
    assert np.array_equal(got[idx], want)


def test_rejects_bad_parameters():
    gms = mm.MixtureSet.from_dict(synth.mixture_set(dim=9, n_mixtures=4, densities_per_mixture=2))
    sc = mm.GmmScorer(gms, "preselection-batch-int")
    with pytest.raises(capi.RasrB200Error):
        sc.configure_preselection(clusters=4, select=5)
    with pytest.raises(capi.RasrB200Error):
        sc.configure_preselection(clusters=300, select=5)


@pytest.mark.parametrize("contraction", [True, False])
def test_float_variant_duplicate_means_tie_like_std_sort(oracle, contraction):
    """RB_GMM_BATCH_PRESELECT (float): densities sharing four means give duplicate centroids, hence exactly equal f32
    distances around the selection boundary; the choice among them follows the reference's std::sort"""
    msd = synth.mixture_set(dim=12, n_mixtures=40, densities_per_mixture=8, seed=3)
    msd["means"] = np.repeat(msd["means"][:4], (msd["means"].shape[0] + 3) // 4, axis=0)[:msd["means"].shape[0]].copy()
    oms, gms = both(oracle, msd)
    f = synth.features(200, 12, seed=4)
    want, cl, means = oracle.gmm_preselect_float(oms, f, use_fma=contraction, clusters=100, select=13)
    sc = mm.GmmScorer(gms, "preselection-batch-float", contraction=contraction)
    sc.configure_preselection(100, 13, 5)
    got_cl, got_means = sc.clustering()
    assert np.array_equal(got_cl, cl) and np.array_equal(got_means, means)
    assert np.array_equal(sc.score(f), want)


