"""signal-dc-detection (src/Signal/DcDetection.{hh,cc}) in the oracle: known-answer cases derived by hand from the
reference's state machine, and the closed formulation the CUDA path uses (flags by neighbour difference + run
lengths), checked against the sequential restatement on integer-valued audio."""
import numpy as np
import pytest

from rasr_b200 import synth


def runs_from_flags(x, min_dc=200, min_seg=416, cut=4096, sr=16000.0, inc=0.9):
    """The event-driven formulation of rasr_b200/csrc/frontend.cu (dc_runs_kernel) in numpy/python: valid when every
    sample equals its predecessor or differs from it by >= inc.  Returns [(begin, end, start_time)]."""
    n = len(x)
    nd = np.ones(n, bool)
    nd[1:] = np.abs(x[1:] - x[:-1]) >= np.float32(inc)
    runs, state = [], dict(seg=0, time=0.0, last_end=-1)

    def emit(b, non_dc, dc):
        state["seg"] += non_dc
        if state["seg"] >= min_seg:
            if runs and state["last_end"] == b:
                runs[-1][1] = b + non_dc
            else:
                runs.append([b, b + non_dc, state["time"]])
            state["last_end"] = b + non_dc
        if dc > 0:
            state["seg"] = 0
        state["time"] += float(non_dc + dc) / sr

    def next_flag(pos, want):
        idx = np.flatnonzero(nd[pos:] == want)
        return pos + int(idx[0]) if idx.size else n

    if n == 0:
        return []
    b, pos = 0, 1
    while True:
        z0 = next_flag(pos, False)
        while True:
            q = max(b + cut, pos)
            if q >= z0:
                break
            emit(b, q - b, 0)
            b, pos = q, q + 1
        if z0 == n:
            emit(b, n - b, 0)
            break
        z1 = next_flag(z0, True)
        dc = z1 - z0
        if z1 == n:
            emit(b, z0 - b, dc) if dc >= min_dc else emit(b, n - b, 0)
            break
        if dc >= min_dc:
            emit(b, z0 - b, dc)
            b = z1
        elif z1 - b >= cut:
            emit(b, z1 - b, 0)
            b = z1
        pos = z1 + 1
    return [tuple(r) for r in runs]


def audio_with_plateaus(n, seed, n_plateaus=6):
    rng = np.random.default_rng(seed)
    x = synth.utterance(n, seed=seed).copy()
    for _ in range(n_plateaus):
        a = int(rng.integers(0, n - 1))
        ln = int(rng.choice([3, 150, 199, 200, 201, 400, 1500, 6000]))
        x[a:a + ln] = x[a]
    return x


def test_signal_without_dc_is_untouched(oracle):
    cfg, x = oracle.frontend_cfg(), synth.utterance(16240)
    a, b = oracle.mfcc(cfg, x), oracle.mfcc_dc(cfg, oracle.dc_cfg(), x, chunk=777)
    assert np.array_equal(a["feats"], b["feats"]) and np.array_equal(a["t_start"], b["t_start"])
    assert np.array_equal(a["t_end"], b["t_end"])
    assert list(b["run_begin"]) == [0] and list(b["run_end"]) == [16240]


def test_known_answers(oracle):
    """400 equal samples from 5000 on: sample 5000 is the last non-DC sample, 5001..5399 (399 >= 200) are discarded,
    5400 restarts everything at 5400 / 16000 s; frames = frames(5001) + frames(10840) = 30 + 67"""
    cfg, dc = oracle.frontend_cfg(), oracle.dc_cfg()
    x = synth.utterance(16240).copy()
    x[5000:5400] = x[5000]
    r = oracle.mfcc_dc(cfg, dc, x)
    assert list(r["run_begin"]) == [0, 5400] and list(r["run_end"]) == [5001, 16240]
    assert r["run_start"][1] == 0.3375 and r["feats"].shape == (97, 39)
    # static features of the second run are those of the run as an utterance of its own; the derivatives are not
    # (the delay window runs across the gap)
    alone = oracle.mfcc(cfg, x[5400:])
    assert np.array_equal(r["feats"][30:, :13], alone["feats"][:, :13])
    assert not np.array_equal(r["feats"][30, 13:], alone["feats"][0, 13:])
    assert np.array_equal(r["feats"][35:, 13:], alone["feats"][5:, 13:])
    assert r["t_start"][32] == 0.3375 and r["t_start"][31] < 0.3375  # frame 31 reaches back into the first run
    # a 199-sample plateau is kept
    y = synth.utterance(16240).copy()
    y[5000:5200] = y[5000]
    assert np.array_equal(oracle.mfcc_dc(cfg, dc, y)["feats"], oracle.mfcc(cfg, y)["feats"])
    # a non-DC island of 300 samples (< 416) between two DC stretches is dropped, one of 416 survives
    for island, kept in ((300, False), (416, True)):
        z = synth.utterance(16240).copy()
        z[3000:3500] = z[3000]
        z[3500 + island - 1:3500 + island + 500] = z[3500 + island - 1]
        rr = oracle.mfcc_dc(cfg, dc, z)
        assert (len(rr["run_begin"]) == 3) == kept, (island, rr["run_begin"], rr["run_end"])
    # all-DC input: one reference sample followed by DC only -> nothing survives
    assert oracle.mfcc_dc(cfg, dc, np.full(5000, 7.0, np.float32))["feats"].shape[0] == 0


@pytest.mark.parametrize("seed", range(6))
def test_flag_formulation_equals_the_sequential_state_machine(oracle, seed):
    cfg, dc = oracle.frontend_cfg(), oracle.dc_cfg()
    x = audio_with_plateaus(24000 + 1000 * seed, seed)
    r = oracle.mfcc_dc(cfg, dc, x, chunk=[0, 160, 4096, 1000, 333, 50000][seed])
    want = list(zip(r["run_begin"], r["run_end"], r["run_start"]))
    assert runs_from_flags(x) == want
