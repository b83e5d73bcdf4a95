"""GPU parity of RB_GMM_BATCH_PRESELECT (Mm::BatchPreselectionFloatFeatureScorer, "preselection-batch-float") against
the CPU oracle, through the C ABI.  The clustering (integer assignment, f32 means) and every score -- including which
mixtures fall back to the back-off score -- must be BIT-IDENTICAL: the scorer is an approximation, and what is
reproduced is the reference's approximation."""
import numpy as np
import pytest

from rasr_b200 import capi, mm, synth

pytestmark = pytest.mark.gpu


def both(oracle, msd):
    return oracle.MixtureSet(**msd), mm.MixtureSet.from_dict(msd)


@pytest.fixture(params=["refinement", "direct"], autouse=True)
def route(request, monkeypatch):
    """both scoring routes: the candidate sets handed to the refinement kernel of the exact batch-float scorer (models it
    covers: mixture count a multiple of 4, at most 32 densities per mixture), and presel_score_kernel (everything else;
    RB_GMM_PRESEL_DIRECT forces it)"""
    if request.param == "direct":
        monkeypatch.setenv("RB_GMM_PRESEL_DIRECT", "1")
    return request.param


@pytest.mark.parametrize("contraction", [True, False])
def test_c2_shape_bit_exact(oracle, diag, contraction):
    msd = synth.mixture_set()
    oms, gms = both(oracle, msd)
    f = synth.features(2000, 39)
    want, cl, means = oracle.gmm_preselect_float(oms, f, use_fma=contraction)
    sc = mm.GmmScorer(gms, "preselection-batch-float", contraction=contraction)
    got_cl, got_means = sc.clustering()
    assert np.array_equal(got_cl, cl) and np.array_equal(got_means, means)
    got = sc.score(f)
    diag("gmm_presel_c2", contraction=contraction, n_diff=int((got != want).sum()), total=got.size,
         backoff_frac=float((want == 40000).mean()))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("clusters,select,iterations,backoff", [(64, 8, 2, 123.5), (256, 256, 5, 40000.0), (7, 1, 0, 1.0),
                                                                (200, 31, 3, 1e6)])
def test_other_clustering_parameters(oracle, clusters, select, iterations, backoff):
    msd = synth.mixture_set(dim=39, n_mixtures=64, densities_per_mixture=16, seed=5)
    oms, gms = both(oracle, msd)
    f = synth.features(600, 39, seed=6)
    want, cl, means = oracle.gmm_preselect_float(oms, f, clusters=clusters, select=select, iterations=iterations,
                                                 backoff=backoff)
    sc = mm.GmmScorer(gms, "preselection-batch-float")
    sc.configure_preselection(clusters, select, iterations, backoff)
    got_cl, got_means = sc.clustering()
    assert np.array_equal(got_cl, cl) and np.array_equal(got_means, means)
    assert np.array_equal(sc.score(f), want)
    if select == clusters:  # no preselection at all: the plain batch scorer
        assert np.array_equal(want, mm.GmmScorer(gms, "batch-float").score(f))


@pytest.mark.parametrize("dim,n_mix,per_mix", [(9, 5, 3), (16, 40, 2), (45, 30, 9), (64, 3, 1)])
def test_small_models_and_dimensions(oracle, dim, n_mix, per_mix):
    """fewer densities than clusters: the number of clusters shrinks to the number of densities
    (DensityClusteringBase::init, src/Mm/DensityClustering.cc:52-56)"""
    msd = synth.mixture_set(dim=dim, n_mixtures=n_mix, densities_per_mixture=per_mix, seed=dim)
    oms, gms = both(oracle, msd)
    f = synth.features(300, dim, seed=dim)
    n_dens = n_mix * per_mix
    select = min(32, n_dens)
    want, cl, means = oracle.gmm_preselect_float(oms, f, select=select)
    sc = mm.GmmScorer(gms, "preselection-batch-float")
    assert sc.clustering()[1].shape[0] == min(256, n_dens)
    assert np.array_equal(sc.clustering()[0], cl) and np.array_equal(sc.score(f), want)


def test_full_size_on_the_device(oracle, diag):
    """BASELINE C2 at full size (100k frames) through the device-pointer entry point; sampled frames are bit-identical
    to the oracle"""
    import torch

    msd = synth.mixture_set()
    oms, gms = both(oracle, msd)
    T = 100000
    f = synth.features(T, 39)
    sc = mm.GmmScorer(gms, "preselection-batch-float")
    d_in = torch.from_numpy(f).cuda()
    d_out = torch.empty((T, 256), dtype=torch.float32, device="cuda")
    sc.score_dev(d_in, T, d_out)
    torch.cuda.synchronize()
    sc.score_dev(d_in, T, d_out)  # a second call on the handle's own stream (timings: bench.py --workload ...)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    idx = np.random.default_rng(0).choice(T, 1500, replace=False)
    idx.sort()
    want = oracle.gmm_preselect_float(oms, f[idx])[0]
    diag("gmm_presel_100k", n_diff=int((got[idx] != want).sum()))
    assert np.array_equal(got[idx], want) and np.isfinite(got).all()


def test_rejects_bad_parameters():
    gms = mm.MixtureSet.from_dict(synth.mixture_set(dim=9, n_mixtures=4, densities_per_mixture=2))
    sc = mm.GmmScorer(gms, "preselection-batch-float")
    with pytest.raises(capi.RasrB200Error):
        sc.configure_preselection(clusters=4, select=5)      # verify(nSelected_ <= nClusters_)
    with pytest.raises(capi.RasrB200Error):
        sc.configure_preselection(clusters=300, select=5)    # clusters: 1..256 (ClusterIndex is u8)
    with pytest.raises(capi.RasrB200Error):
        mm.GmmScorer(gms, "batch-float").configure_preselection()
