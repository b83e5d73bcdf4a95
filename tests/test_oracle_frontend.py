"""Oracle pins for the front-end: the reference's own FFT object code (oracle/_ref) and self-derived
known-answer vectors.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from rasr_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_geometry_16k(oracle):
    cfg = oracle.frontend_cfg()
    g = oracle.geometry(cfg)
    assert (g.win_length, g.win_shift, g.fft_length, g.n_bins) == (400, 160, 512, 257)
    assert g.n_filters == 20
    assert g.n_weights == 479
    assert g.feat_dim == 39
    t = oracle.tables(cfg)
    assert (t["fb_start"][0], t["fb_end"][0]) == (0, 7)
    assert (t["fb_start"][19], t["fb_end"][19]) == (197, 257)
    assert int((t["fb_weights"] != 0).sum()) <= 479
    assert np.all(t["fb_weights"] >= 0)


def test_frame_count_and_tail(oracle):
    cfg = oracle.frontend_cfg()
    assert oracle.nframes(cfg, 160000) == 999
    assert oracle.nframes(cfg, 0) == 0
    assert oracle.nframes(cfg, 1) == 1
    assert oracle.nframes(cfg, 400) == 1
    assert oracle.nframes(cfg, 401) == 2
    assert oracle.nframes(cfg, 560) == 2
    assert oracle.nframes(cfg, 561) == 3
    x = synth.utterance(160000)
    r = oracle.mfcc(cfg, x)
    assert r["feats"].shape == (999, 39)
    # static frame t covers [t*S, t*S+len)/sr; the concat packet spans the 5-frame delay window
    assert r["t_start"][0] == 0.0 and r["t_start"][10] == pytest.approx(0.08, abs=1e-12)
    assert r["t_end"][998] == pytest.approx(10.0, abs=1e-9)
    assert np.isfinite(r["feats"]).all()


@pytest.mark.parametrize("chunk", [1, 7, 160, 400, 4096, 100000])
def test_chunking_invariance(oracle, chunk):
    """Results do not depend on how the audio node packetises the samples."""
    cfg = oracle.frontend_cfg()
    x = synth.utterance(16000 if chunk > 7 else 2000)
    a = oracle.mfcc(cfg, x, chunk=0)
    b = oracle.mfcc(cfg, x, chunk=chunk)
    assert np.array_equal(a["feats"], b["feats"])
    assert np.array_equal(a["t_start"], b["t_start"]) and np.array_equal(a["t_end"], b["t_end"])


def test_preemphasis_first_sample_and_window(oracle):
    cfg = oracle.frontend_cfg(derivatives=False)
    x = synth.utterance(1600)
    r = oracle.mfcc(cfg, x, stages=True)
    w = oracle.tables(cfg)["window"]
    assert w[0] == np.float32(0.54 - 0.46) and w[0] == w[399]
    assert abs(w[199] - 1.0) < 1e-4
    # spectrum of frame 0 == FFT of window * diff(x) with first sample 0, in the reference's packing
    pre = np.empty(400, np.float32)
    pre[0] = 0.0
    pre[1:] = x[1:400] - x[:399]
    v = np.zeros(512, np.float32)
    v[:400] = w * pre
    ref = np.fft.rfft(v.astype(np.float64)) / 16000.0
    got = r["spectrum"][0].reshape(257, 2)
    # sign convention e^{+i theta}: imaginary parts are the negatives of numpy's
    scale = np.abs(ref).max()
    assert np.abs(got[:, 0] - ref.real).max() / scale < 2e-6
    assert np.abs(got[:, 1] + ref.imag).max() / scale < 2e-6
    assert got[0, 1] == 0.0 and got[256, 1] == 0.0


def test_fft_impulse_kat(oracle):
    """Impulse at n=1, N=8: 1 -1 .7071 .7071 0 1 -.7071 .7071 in packed layout."""
    v = np.zeros(8, np.float32)
    v[1] = 1.0
    out = oracle.fft_real_packed(v)
    s = np.float32(np.sqrt(0.5))
    np.testing.assert_allclose(out, [1, -1, s, s, 0, 1, -s, s], atol=1e-7)


@pytest.mark.parametrize("n", [8, 64, 512, 1024])
def test_fft_restatement_equals_reference_object_code(oracle, n):
    """The restated FFT is bit-identical to the reference's TU compiled with strict rounding and at
    most 1 f32-ulp-of-max away from the same TU compiled with the reference's default flags."""
    ref = oracle.ref_fft(native=False)
    if ref is None:
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    rng = np.random.default_rng(n)
    for trial in range(20):
        x = (rng.standard_normal(n) * 1000).astype(np.float32)
        if trial == 0:
            x[n // 2:] = 0  # zero padding as in the pipeline
        mine = oracle.fft_real_packed(x)
        theirs = x.copy()
        ref.ref_fft_transform_real(theirs.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(n))
        assert np.array_equal(mine, theirs)
        nat = oracle.ref_fft(native=True)
        native = x.copy()
        nat.ref_fft_transform_real(native.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(n))
        assert np.abs(native - mine).max() <= np.abs(mine).max() * 2.0 ** -22


def test_regression_coefficients(oracle):
    """Delta = [-2 -1 0 1 2]/10, delta-delta = [2 -1 -2 -1 2]/7 on the static window (edges replicated)."""
    cfg = oracle.frontend_cfg()
    x = synth.utterance(8000)
    r = oracle.mfcc(cfg, x, stages=True)
    c = r["cepstra"].astype(np.float64)
    T = c.shape[0]
    idx = np.clip(np.arange(T)[:, None] + np.arange(-2, 3)[None, :], 0, T - 1)
    win = c[idx]  # T x 5 x 13
    d = (win * (np.array([-2, -1, 0, 1, 2]) / 10.0)[None, :, None]).sum(1)
    dd = (win * (np.array([2, -1, -2, -1, 2]) / 7.0)[None, :, None]).sum(1)
    f = r["feats"]
    assert np.array_equal(f[:, :13], r["cepstra"])
    np.testing.assert_allclose(f[:, 13:26], d, rtol=0, atol=2e-5)
    np.testing.assert_allclose(f[:, 26:], dd, rtol=0, atol=4e-5)


def test_mfcc_against_float64_model(oracle):
    """Independent numpy float64 model of the whole chain agrees with the restatement to f32 accuracy."""
    cfg = oracle.frontend_cfg(derivatives=False)
    x = synth.utterance(32000)
    r = oracle.mfcc(cfg, x, stages=True)
    t = oracle.tables(cfg)
    T = r["feats"].shape[0]
    pre = np.concatenate([[0.0], np.diff(x.astype(np.float64))])
    frames = np.zeros((T, 512))
    for i in range(T):
        seg = pre[i * 160:i * 160 + 400]
        frames[i, :seg.size] = seg * t["window"][:seg.size]
    amp = np.abs(np.fft.rfft(frames, axis=1)) / 16000.0
    fb = amp @ t["fb_weights"].astype(np.float64).T
    cep = np.log10(fb) @ t["dct"].astype(np.float64).T
    np.testing.assert_allclose(r["amplitude"], amp, rtol=0, atol=amp.max() * 1e-6)
    np.testing.assert_allclose(r["fbank"], fb, rtol=2e-5)
    np.testing.assert_allclose(r["cepstra"], cep, rtol=0, atol=2e-4)


def test_fma_variant_is_close(oracle):
    x = synth.utterance(16000)
    a = oracle.mfcc(oracle.frontend_cfg(use_fma=True), x)["feats"]
    b = oracle.mfcc(oracle.frontend_cfg(use_fma=False), x)["feats"]
    assert np.abs(a - b).max() < 1e-4


@pytest.mark.parametrize("wt", ["hamming", "rectangular", "hanning", "periodic-hanning", "bartlett", "blackman", "kaiser"])
def test_window_functions(oracle, wt):
    """the window types of src/Signal/WindowFunction.cc:62-132 against their textbook definitions (numpy, f64) and
    the properties the reference's fill order implies"""
    L = 400
    w = oracle.tables(oracle.frontend_cfg(window_type=wt))["window"]
    n = np.arange(L, dtype=np.float64)
    # kaiser: WindowFunction::create builds it with beta = 0 and nothing sets beta: I0(0) / I0(0) = 1 everywhere
    want = {"hamming": 0.54 - 0.46 * np.cos(2 * np.pi * n / (L - 1)), "rectangular": np.ones(L), "kaiser": np.kaiser(L, 0.0),
            "hanning": 0.5 - 0.5 * np.cos(2 * np.pi * n / (L - 1)), "periodic-hanning": 0.5 - 0.5 * np.cos(2 * np.pi * n / L),
            "bartlett": 1 - np.abs(2 * n / (L - 1) - 1),
            "blackman": 0.42 - 0.5 * np.cos(2 * np.pi * n / (L - 1)) + 0.08 * np.cos(4 * np.pi * n / (L - 1))}[wt]
    assert w.dtype == np.float32 and w.shape == (L,)
    assert np.abs(w - want).max() < 2e-7
    if wt != "periodic-hanning":
        assert np.array_equal(w, w[::-1])            # filled symmetrically, bit for bit
    else:
        assert np.array_equal(w[1:], w[1:][::-1]) and w[0] == 0.0 and w[L // 2] == 1.0
    if wt in ("hanning", "bartlett"):
        assert w[0] == 0.0 and w[-1] == 0.0
