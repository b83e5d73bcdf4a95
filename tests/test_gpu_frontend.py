"""GPU parity of the MFCC front-end against the CPU oracle, through the C ABI.

Frame counts and timestamps are bit-exact (integer / host f64 work); float stages are compared at the
north_star tolerance of 1e-4 relative -- relative to the scale of each stage, because cepstra and
derivatives cross zero."""
import os

import numpy as np
import pytest

from rasr_b200 import flow, synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-4


def rel_err(got, want):
    scale = np.maximum(np.abs(want), np.sqrt(np.mean(want.astype(np.float64) ** 2)))
    return float((np.abs(got.astype(np.float64) - want) / scale).max())


def test_geometry_and_tables_match_the_oracle(oracle):
    fe = flow.FrontEnd()
    g, og = fe.geometry, oracle.geometry(oracle.frontend_cfg())
    for k in ("win_length", "win_shift", "fft_length", "n_bins", "n_filters", "n_weights", "feat_dim"):
        assert getattr(g, k) == getattr(og, k), k
    t, ot = fe.tables(), oracle.tables(oracle.frontend_cfg())
    for k in t:
        assert np.array_equal(t[k], ot[k]), k


@pytest.mark.parametrize("sr,nc", [(8000.0, 12), (16000.0, 16), (22050.0, 13), (44100.0, 20)])
def test_tables_other_sample_rates(oracle, sr, nc):
    fe = flow.FrontEnd(sample_rate=sr, n_cepstra=nc)
    ocfg = oracle.frontend_cfg(sample_rate=sr, n_cepstra=nc)
    t, ot = fe.tables(), oracle.tables(ocfg)
    for k in t:
        assert np.array_equal(t[k], ot[k]), k
    assert fe.geometry.fft_length == oracle.geometry(ocfg).fft_length


def test_c1_utterance_stage_by_stage(oracle, diag):
    """BASELINE config C1: one 10 s 16 kHz utterance, 999 frames, last frame 320 samples."""
    x = synth.utterance(160000)
    fe = flow.FrontEnd()
    r = fe.process(x, stages=True)
    o = oracle.mfcc(oracle.frontend_cfg(), x, stages=True)
    assert r["feats"].shape == (999, 39)
    assert np.array_equal(r["t_start"], o["t_start"]) and np.array_equal(r["t_end"], o["t_end"])
    errs = {}
    for k in ("amplitude", "fbank", "cepstra", "feats"):
        errs[k] = rel_err(r[k], o[k])
    # amplitude is compared against the spectrum's full scale (individual bins can be ~0)
    errs["amplitude_fullscale"] = float(np.abs(r["amplitude"] - o["amplitude"]).max() / o["amplitude"].max())
    errs["fbank_rel"] = float((np.abs(r["fbank"] - o["fbank"]) / o["fbank"]).max())
    diag("frontend_c1", **errs)
    assert errs["amplitude_fullscale"] < 1e-5
    assert errs["fbank_rel"] < RTOL
    assert errs["cepstra"] < RTOL
    assert errs["feats"] < RTOL
    assert np.array_equal(r["feats"][:, :13], r["cepstra"])


@pytest.mark.parametrize("n", [1, 2, 3, 160, 399, 400, 401, 560, 561, 1000, 5119, 5120, 5121, 5361])
def test_short_and_boundary_lengths(oracle, n):
    """Frame count / short last frame / tile boundaries (32 frames = 5120 samples of shift)."""
    x = synth.utterance(n, seed=n)
    fe = flow.FrontEnd()
    r = fe.process(x)
    o = oracle.mfcc(oracle.frontend_cfg(), x)
    assert r["feats"].shape == o["feats"].shape
    assert fe.nframes_for(n) == oracle.nframes(oracle.frontend_cfg(), n)
    assert np.array_equal(r["t_start"], o["t_start"]) and np.array_equal(r["t_end"], o["t_end"])
    finite = np.isfinite(o["feats"])
    assert np.array_equal(np.isfinite(r["feats"]), finite)
    if finite.any():
        assert rel_err(r["feats"][finite], o["feats"][finite]) < RTOL


def test_empty_segment():
    fe = flow.FrontEnd()
    r = fe.process(np.zeros(0, np.float32))
    assert r["feats"].shape == (0, 39)
    assert fe.nframes_for(0) == 0


def test_batch_of_utterances_equals_one_by_one(oracle):
    """Utterances are independent units: no halo crosses a segment boundary."""
    lens = [16000, 401, 7, 48000, 5121, 1, 160240]
    parts = [synth.utterance(n, seed=100 + i) for i, n in enumerate(lens)]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    fe = flow.FrontEnd()
    r = fe.process(np.concatenate(parts), offs)
    fo = r["frame_offsets"]
    assert fo[-1] == r["feats"].shape[0]
    for i, p in enumerate(parts):
        single = fe.process(p)
        assert np.array_equal(r["feats"][fo[i]:fo[i + 1]], single["feats"], equal_nan=True), i
        assert np.array_equal(r["t_start"][fo[i]:fo[i + 1]], single["t_start"])
        o = oracle.mfcc(oracle.frontend_cfg(), p)
        finite = np.isfinite(o["feats"])
        assert np.array_equal(np.isfinite(single["feats"]), finite)
        if finite.any():
            assert rel_err(single["feats"][finite], o["feats"][finite]) < RTOL


@pytest.mark.parametrize("chunk", [1, 160, 4096, 100000])
def test_streaming_protocol_is_packet_size_invariant(oracle, chunk):
    """Flow::Node view: packets of any size, EOS computes the segment."""
    x = synth.utterance(4000 if chunk == 1 else 24000, seed=77)
    fe = flow.FrontEnd()
    whole = fe.process(x)
    fe.reset()
    for a in range(0, x.size, chunk):
        fe.push(x[a:a + chunk], a / 16000.0)
    feats, ts, te = fe.finish()
    assert np.array_equal(feats, whole["feats"])
    assert np.array_equal(ts, whole["t_start"]) and np.array_equal(te, whole["t_end"])


def test_mfcc_node_emits_one_packet_per_frame(oracle):
    x = synth.utterance(8000, seed=5)
    node = flow.MfccNode()
    assert node.set_parameter("nr-outputs", "13")
    assert node.configure({"datatype": "vector-f32", "sample-rate": "16000"})
    assert node.output_attributes["frame-shift"] == "0.01"
    for a in range(0, x.size, 1024):
        node.put(flow.Packet(x[a:a + 1024], a / 16000.0, (a + 1024) / 16000.0))
    node.put(flow.EOS)
    o = oracle.mfcc(oracle.frontend_cfg(), x)
    n = 0
    while True:
        p = node.work()
        if p is flow.EOS:
            break
        assert p.start == o["t_start"][n] and p.end == o["t_end"][n]
        assert rel_err(p.data, o["feats"][n]) < 1e-3
        n += 1
    assert n == o["feats"].shape[0]


@pytest.mark.parametrize("kw", [dict(alpha=0.97), dict(derivatives=False), dict(n_cepstra=16),
                                dict(sample_rate=8000.0), dict(window_shift=0.005),
                                dict(window_length=0.02, window_shift=0.03)])
def test_other_configurations(oracle, diag, kw):
    x = synth.utterance(20000, seed=8)
    fe = flow.FrontEnd(**kw)
    okw = dict(kw)
    for a, b in (("window_length", "window_length_s"), ("window_shift", "window_shift_s")):
        if a in okw:
            okw[b] = okw.pop(a)
    o = oracle.mfcc(oracle.frontend_cfg(**okw), x)
    r = fe.process(x)
    assert r["feats"].shape == o["feats"].shape
    assert np.array_equal(r["t_start"], o["t_start"]) and np.array_equal(r["t_end"], o["t_end"])
    e = rel_err(r["feats"], o["feats"])
    diag("frontend_cfg", cfg=str(kw), err=e)
    assert e < RTOL


def test_s16_pcm_input_equals_float_input():
    """rb_frontend_process_s16: demultiplex + s16 -> f32 on the device (samples.flow:13-18) is a plain value
    conversion, so the features are bit-identical to feeding the converted samples."""
    samples, offs = synth.corpus(4, n_samples=20240)
    assert np.array_equal(samples, np.round(samples)) and np.abs(samples).max() < 32768
    fe = flow.FrontEnd()
    want = fe.process(samples, offs)
    mono = fe.process_s16(samples.astype(np.int16), offs)
    assert np.array_equal(mono["feats"], want["feats"]) and np.array_equal(mono["t_start"], want["t_start"])
    # track 1 of a 3-channel interleaved stream
    pcm = np.zeros((samples.size, 3), np.int16)
    pcm[:, 0] = 7
    pcm[:, 1] = samples.astype(np.int16)
    pcm[:, 2] = -samples.astype(np.int16)
    multi = fe.process_s16(pcm, offs, n_channels=3, track=1)
    assert np.array_equal(multi["feats"], want["feats"])
    with pytest.raises(Exception):
        fe.process_s16(pcm, offs, n_channels=3, track=3)


@pytest.mark.parametrize("wt", ["rectangular", "hanning", "periodic-hanning", "bartlett", "blackman", "kaiser"])
def test_other_window_types(oracle, wt):
    """signal-window type= (src/Signal/WindowFunction.cc:25-33): the table is bit-identical to the oracle's, the
    features follow at the front-end's tolerance"""
    x = synth.utterance(20000, seed=14)
    fe = flow.FrontEnd(window_type=wt)
    ocfg = oracle.frontend_cfg(window_type=wt)
    assert np.array_equal(fe.tables()["window"], oracle.tables(ocfg)["window"])
    r, o = fe.process(x), oracle.mfcc(ocfg, x)
    assert r["feats"].shape == o["feats"].shape and rel_err(r["feats"], o["feats"]) < RTOL
    assert not np.array_equal(r["feats"], flow.FrontEnd().process(x)["feats"])
    node = flow.MfccNode()
    assert node.set_parameter("window-type", wt) and node.configure({"sample-rate": "16000"})


def test_unknown_window_type_is_rejected():
    from rasr_b200 import capi
    with pytest.raises(capi.RasrB200Error):
        flow.FrontEnd(window_type="gaussian")
