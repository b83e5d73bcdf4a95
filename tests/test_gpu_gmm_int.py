"""GPU parity of RB_GMM_BATCH_INT (Mm::BatchIntFeatureScorer, "batch-diagonal-maximum-int") against the CPU
oracle, through the C ABI.  Integer path: every score must be BIT-IDENTICAL (north_star: bit-exact for integer
work); the oracle itself is pinned by an independent numpy restatement in tests/test_oracle_gmm.py."""
import numpy as np
import pytest

from rasr_b200 import capi, mm, synth

pytestmark = pytest.mark.gpu


def both(oracle, msd):
    return oracle.MixtureSet(**msd), mm.MixtureSet.from_dict(msd)


def test_batch_int_c2_shape_bit_exact(oracle, diag):
    """C2 geometry (39-dim, 256 mixtures x 16 densities) at a size the oracle finishes in seconds."""
    msd = synth.mixture_set()
    oms, gms = both(oracle, msd)
    f = synth.features(3000, 39)
    want = oracle.gmm_batch_int(oms, f, threads=8)
    got = mm.GmmScorer(gms, "batch-int").score(f)
    diag("gmm_int_c2", n_diff=int((got != want).sum()), total=got.size, max_abs=np.abs(got - want).max())
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dim", [1, 7, 16, 17, 33, 39, 48, 64])
def test_batch_int_dimensions(oracle, dim):
    msd = synth.mixture_set(dim=dim, n_mixtures=12, densities_per_mixture=5, seed=dim)
    oms, gms = both(oracle, msd)
    f = synth.features(700, dim, seed=dim)
    assert np.array_equal(mm.GmmScorer(gms, "batch-int").score(f), oracle.gmm_batch_int(oms, f))


@pytest.mark.parametrize("T", [1, 2, 63, 64, 65, 511, 512, 513, 1500])
def test_batch_int_ragged_frame_counts(oracle, T):
    msd = synth.mixture_set(dim=39, n_mixtures=8, densities_per_mixture=16, seed=3)
    oms, gms = both(oracle, msd)
    f = synth.features(T, 39, seed=T)
    got = mm.GmmScorer(gms, "batch-int").score(f)
    assert got.shape == (T, 8)
    assert np.array_equal(got, oracle.gmm_batch_int(oms, f))


@pytest.mark.parametrize("sizes", [(1, 3, 16, 7, 32, 2), (3, 0, 5, 1), (9, 8, 7, 24, 0, 0, 1, 40), (64,) * 5])
def test_batch_int_ragged_and_empty_mixtures(oracle, sizes):
    """Mixtures of unequal size (padded to 8-column tiles on the device), mixtures without densities score
    (f32)INT_MAX / scale_ as in the reference, mixture counts that are not multiples of 4 (scalar stores)."""
    msd = synth.ragged_mixture_set(dim=39, sizes=sizes, seed=len(sizes))
    oms, gms = both(oracle, msd)
    f = synth.features(300, 39, seed=8)
    assert np.array_equal(mm.GmmScorer(gms, "batch-int").score(f), oracle.gmm_batch_int(oms, f))


def test_batch_int_clipping(oracle):
    """Features far outside the quantisation interval clip at 0 / 255 exactly as quantize() does."""
    msd = synth.mixture_set(dim=39, n_mixtures=16, densities_per_mixture=8, seed=21)
    oms, gms = both(oracle, msd)
    f = synth.features(256, 39, seed=1, scale=40.0)
    f[0, :] = 1e30
    f[1, :] = -1e30
    f[2, ::2] = 0.0
    assert np.array_equal(mm.GmmScorer(gms, "batch-int").score(f), oracle.gmm_batch_int(oms, f))


def test_batch_int_device_pointers_and_full_size(oracle, diag):
    """BASELINE C2 at full size (100k frames) through the device-pointer entry point; 4096 sampled frames are
    compared with the oracle, all rows must be finite and no better than the per-row oracle minimum allows."""
    import torch

    msd = synth.mixture_set()
    oms, gms = both(oracle, msd)
    T = 100000
    f = synth.features(T, 39)
    sc = mm.GmmScorer(gms, "batch-int")
    d_in = torch.from_numpy(f).cuda()
    d_out = torch.empty((T, 256), dtype=torch.float32, device="cuda")
    sc.score_dev(d_in, T, d_out)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    idx = np.random.default_rng(0).choice(T, 4096, replace=False)
    idx.sort()
    want = oracle.gmm_batch_int(oms, f[idx], threads=8)
    diag("gmm_int_100k", n_diff=int((got[idx] != want).sum()), total=want.size)
    assert np.array_equal(got[idx], want)
    assert np.isfinite(got).all()
    # a second call reuses the handle and its staging: results are deterministic
    sc.score_dev(d_in, T, d_out)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), got)


def test_batch_int_rejects_what_the_reference_rejects():
    msd = synth.ragged_mixture_set(dim=16, n_covariances=2)
    with pytest.raises(capi.RasrB200Error):
        mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-int")
    msd = synth.mixture_set(dim=39, n_mixtures=4, densities_per_mixture=4)
    sc = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-int")
    with pytest.raises(capi.RasrB200Error):
        sc.score(np.zeros((4, 39), np.float32), want_density=True)
