"""GPU parity of signal-dc-detection + MFCC front-end (rb_frontend_process_dc) against the CPU oracle.  The kept sample
runs, their start times, frame counts and timestamps are bit-exact; features at the 1e-4 tolerance of the front-end."""
import numpy as np
import pytest

from rasr_b200 import capi, flow, synth
from tests.test_gpu_frontend import RTOL, rel_err
from tests.test_oracle_dc import audio_with_plateaus

pytestmark = pytest.mark.gpu


def check_against_oracle(oracle, fe, x, offs, r, dc=None, cfg=None):
    fo = r["frame_offsets"]
    k = 0
    for u in range(len(offs) - 1):
        o = oracle.mfcc_dc(oracle.frontend_cfg(**(cfg or {})), oracle.dc_cfg(**(dc or {})), x[offs[u]:offs[u + 1]])
        n = len(o["run_begin"])
        sel = slice(k, k + n)
        assert list(r["runs"]["utt"][sel]) == [u] * n
        assert list(r["runs"]["begin"][sel] - offs[u]) == list(o["run_begin"]), u
        assert list(r["runs"]["end"][sel] - offs[u]) == list(o["run_end"]), u
        assert np.array_equal(r["runs"]["start"][sel], o["run_start"]), u
        k += n
        T = o["feats"].shape[0]
        assert fo[u + 1] - fo[u] == T, u
        f = slice(int(fo[u]), int(fo[u + 1]))
        assert np.array_equal(r["t_start"][f], o["t_start"]) and np.array_equal(r["t_end"][f], o["t_end"]), u
        if T:
            # a frame of nothing but equal samples is all zero after pre-emphasis: log10(0) = -inf and NaN cepstra,
            # on both sides (only seen when the detector is switched off)
            ok = np.isfinite(o["feats"]).all(axis=1)
            assert np.array_equal(np.isfinite(r["feats"][f]).all(axis=1), ok), u
            assert rel_err(r["feats"][f][ok], o["feats"][ok]) < RTOL, u
    assert k == len(r["runs"]["utt"])


def test_batch_with_dc_stretches(oracle, diag):
    lens = [24000, 160, 31000, 5000, 1, 40000, 0, 27000]
    xs = [audio_with_plateaus(n, 40 + i) if n > 2000 else synth.utterance(max(n, 1), seed=40 + i)[:n]
          for i, n in enumerate(lens)]
    xs[3] = np.full(5000, 3.0, np.float32)  # nothing but DC: the utterance yields no frame
    x = np.concatenate(xs)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    fe = flow.FrontEnd()
    r = fe.process_dc(x, offs)
    assert not r["runs"]["sequential_path"]  # integer-valued audio: the flag formulation applies
    check_against_oracle(oracle, fe, x, offs, r)
    assert r["frame_offsets"][4] == r["frame_offsets"][3]
    diag("frontend_dc", runs=int(len(r["runs"]["utt"])), frames=int(r["frame_offsets"][-1]))


def test_sequential_path_equals_flag_path(oracle, monkeypatch):
    x = audio_with_plateaus(50000, 3)
    fe = flow.FrontEnd()
    a = fe.process_dc(x)
    monkeypatch.setenv("RB_DC_SEQUENTIAL", "1")
    b = fe.process_dc(x)
    assert b["runs"]["sequential_path"] and not a["runs"]["sequential_path"]
    for k in ("begin", "end", "start"):
        assert np.array_equal(a["runs"][k], b["runs"][k])
    assert np.array_equal(a["feats"], b["feats"]) and np.array_equal(a["t_start"], b["t_start"])


def test_float_input_replays_the_reference_chain(oracle):
    """samples that drift by less than the increment per step: the last non-DC sample is NOT the previous sample, the
    flags cannot be computed from neighbours; the library notices and restates the chain sample by sample"""
    rng = np.random.default_rng(5)
    x = audio_with_plateaus(30000, 9) + rng.uniform(-0.2, 0.2, 30000).astype(np.float32)
    x[8000:9000] = x[8000] + np.linspace(0, 3.0, 1000, dtype=np.float32)  # a slow ramp: DC by pieces
    fe = flow.FrontEnd()
    r = fe.process_dc(x)
    assert r["runs"]["sequential_path"]
    check_against_oracle(oracle, fe, x, np.array([0, x.size], np.int64), r)


@pytest.mark.parametrize("dc", [dict(min_dc_length_s=0.002, min_non_dc_segment_length_s=0.001, maximal_output_size=100),
                                dict(max_dc_increment=0.0), dict(min_dc_length_s=0.0)])
def test_other_parameters(oracle, dc):
    x = audio_with_plateaus(20000, 11, n_plateaus=12)
    fe = flow.FrontEnd()
    r = fe.process_dc(x, dc=dc)
    check_against_oracle(oracle, fe, x, np.array([0, x.size], np.int64), r, dc)


def test_without_dc_equals_plain_processing():
    x, offs = synth.corpus(4, n_samples=12240)
    fe = flow.FrontEnd()
    a, b = fe.process(x, offs), fe.process_dc(x, offs)
    assert np.array_equal(a["feats"], b["feats"]) and np.array_equal(a["frame_offsets"], b["frame_offsets"])
    assert np.array_equal(a["t_start"], b["t_start"]) and np.array_equal(a["t_end"], b["t_end"])


def test_capacity_is_checked():
    import ctypes as C
    x = synth.utterance(16000)
    fe = flow.FrontEnd()
    cfg = capi.DcCfg()
    capi.lib().rb_dc_default_cfg(C.byref(cfg))
    offs, fo = np.array([0, x.size], np.int64), np.zeros(2, np.int64)
    feats = np.zeros((10, 39), np.float32)
    rc = capi.lib().rb_frontend_process_dc(fe.handle, C.byref(cfg), capi.ptr(x), capi.ptr(offs), 1, capi.ptr(feats), 10,
                                           capi.ptr(fo), None, None)
    assert rc == -1 and b"99 are needed" in capi.lib().rb_last_error()


def test_flow_node_with_dc_detection(oracle):
    """the Flow-node mirror (what adapters/B200MfccNode.cc does): packets with a start time in, dc-detection="true";
    the start times of the runs continue from the first packet's start time like DcDetection::put (:105-108)"""
    x = audio_with_plateaus(30000, 17)
    node = flow.MfccNode()
    assert node.set_parameter("dc-detection", "true") and node.set_parameter("min-dc-length", "0.0125")
    assert node.configure({"sample-rate": "16000", "datatype": "vector-f32"})
    t0 = 1.25
    for a in range(0, x.size, 4000):
        node.put(flow.Packet(x[a:a + 4000], t0 + a / 16000.0, t0 + min(a + 4000, x.size) / 16000.0))
    node.put(flow.EOS)
    out = []
    while True:
        p = node.work()
        if p is flow.EOS:
            break
        out.append(p)
    o = oracle.mfcc_dc(oracle.frontend_cfg(), oracle.dc_cfg(), x)
    assert len(out) == o["feats"].shape[0] and len(o["run_begin"]) > 1
    # the oracle entry starts at time 0: shifted times agree up to the rounding of the additions
    assert np.allclose([p.start for p in out], t0 + o["t_start"], rtol=0, atol=1e-9)
    assert np.allclose([p.end for p in out], t0 + o["t_end"], rtol=0, atol=1e-9)
    assert rel_err(np.stack([p.data for p in out]), o["feats"]) < RTOL


@pytest.mark.parametrize("kw", [dict(sample_rate=8000.0), dict(window_shift_s=0.005), dict(derivatives=0)])
def test_other_front_end_geometries(oracle, kw):
    """the generic front-end kernel (any geometry but the 512-point one) and other frame shifts also take the kept
    runs as segments; min-dc-length etc. are given in seconds and follow the sample rate"""
    x = audio_with_plateaus(30000, 23, n_plateaus=8)
    fkw = dict(kw)
    if "window_shift_s" in fkw:
        fkw["window_shift"] = fkw.pop("window_shift_s")
    if "derivatives" in fkw:
        fkw["derivatives"] = bool(fkw["derivatives"])
    fe = flow.FrontEnd(**fkw)
    r = fe.process_dc(x)
    check_against_oracle(oracle, fe, x, np.array([0, x.size], np.int64), r, cfg=kw)
