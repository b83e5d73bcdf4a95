"""GPU parity of the feature post-processing stages (signal-normalization, sequence concatenation, matrix
multiplication; rb_postproc_*) against the CPU oracle, through the C ABI.  All three follow the reference's
operation order (running f64 sums, sequential f32 dot products), so results must be BIT-IDENTICAL."""
import numpy as np
import pytest

from rasr_b200 import capi, postproc

pytestmark = pytest.mark.gpu


def segments(rng, n_utt, lo, hi):
    lens = rng.integers(lo, hi, n_utt)
    fo = np.zeros(n_utt + 1, np.int64)
    fo[1:] = np.cumsum(lens)
    return fo


@pytest.mark.parametrize("kind", ["mean", "mean-and-variance"])
@pytest.mark.parametrize("length,right", [("infinite", "infinite"), (201, 100), (5, 2), (7, 0), (9, 8)])
@pytest.mark.parametrize("fma", [True, False])
def test_normalization_bit_exact(oracle, kind, length, right, fma):
    rng = np.random.default_rng(3)
    fo = segments(rng, 9, 1, 400)
    x = (rng.standard_normal((int(fo[-1]), 39)) * 4 + 2).astype(np.float32)
    L = -1 if length == "infinite" else length
    R = -1 if right == "infinite" else right
    want = oracle.normalize(x, fo, kind, L, R, use_fma=fma)
    got = postproc.PostProcessor(39, kind, length, right, contraction=fma).process(x, fo)
    assert np.array_equal(got, want)


def test_constant_dimension_and_single_frame(oracle):
    """standard deviation 0 -> 1 (src/Signal/Normalization.cc:169-173), one-frame segments"""
    x = np.random.default_rng(1).standard_normal((50, 13)).astype(np.float32)
    x[:, 4] = 2.5
    fo = [0, 1, 2, 50]
    want = oracle.normalize(x, fo, "mean-and-variance")
    got = postproc.PostProcessor(13, "mean-and-variance").process(x, fo)
    assert np.array_equal(got, want) and np.all(got[:, 4] == 0)


@pytest.mark.parametrize("length,right", [(11, 5), (5, 2), (3, 0), (2, 1), (1, 0)])
def test_splice_bit_exact(oracle, length, right):
    rng = np.random.default_rng(length)
    fo = segments(rng, 7, 1, 150)
    x = rng.standard_normal((int(fo[-1]), 39)).astype(np.float32)
    got = postproc.PostProcessor(39, splice=(length, right)).process(x, fo)
    assert got.shape == (x.shape[0], 39 * length)
    assert np.array_equal(got, oracle.splice(x, length, right, fo))


@pytest.mark.parametrize("fma", [True, False])
@pytest.mark.parametrize("rows,cols", [(45, 429), (1, 7), (64, 64), (130, 33)])
def test_matrix_multiplication_bit_exact(oracle, rows, cols, fma):
    rng = np.random.default_rng(rows)
    M = rng.standard_normal((rows, cols)).astype(np.float32)
    x = rng.standard_normal((333, cols)).astype(np.float32)
    got = postproc.PostProcessor(cols, matrix=M, contraction=fma).process(x)
    assert np.array_equal(got, oracle.matmul(M, x, use_fma=fma))


def test_full_chain_cmvn_splice_lda(oracle, diag):
    """processing.standard_system.flow + lda.flow: segment CMVN -> 11-frame window -> 429 x 45 matrix, on 125
    utterances x 1000 frames of 39-dim features (one C3 shard), device-pointer entry point."""
    import torch

    rng = np.random.default_rng(9)
    fo = np.arange(126, dtype=np.int64) * 1000
    x = (rng.standard_normal((125000, 39)) * 3 + 1).astype(np.float32)
    M = (rng.standard_normal((45, 429)) / 20).astype(np.float32)
    pp = postproc.PostProcessor(39, "mean-and-variance", splice=(11, 5), matrix=M)
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty((125000, 45), dtype=torch.float32, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    pp.process_dev(d_in, fo, d_out)
    torch.cuda.synchronize()
    ev[0].record()
    pp.process_dev(d_in, fo, d_out)
    ev[1].record()
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    sel = slice(17000, 21000)  # 4 utterances checked against the oracle
    n = oracle.normalize(x[sel], fo[17:22] - 17000, "mean-and-variance")
    want = oracle.matmul(M, oracle.splice(n, 11, 5, fo[17:22] - 17000))
    diag("postproc_chain", ms=ev[0].elapsed_time(ev[1]), frames=125000, n_diff=int((got[sel] != want).sum()))
    assert np.array_equal(got[sel], want)
    assert np.isfinite(got).all()


def test_rejects_bad_configuration():
    with pytest.raises(capi.RasrB200Error):
        postproc.PostProcessor(39, "mean", 5, 5)  # length <= right
    with pytest.raises(capi.RasrB200Error):
        postproc.PostProcessor(39, splice=(3, 3))
    with pytest.raises(capi.RasrB200Error):
        postproc.PostProcessor(39, splice=(11, 5), matrix=np.zeros((45, 39), np.float32))  # needs 429 columns
