"""Oracle pins for the feature post-processing nodes (signal-normalization, sequence concatenation, matrix
multiplication).  The reference has no unit test for them: parity unpinned by the reference, pinned here by a
LITERAL Python transcription of the SlidingWindow / Normalization call protocol (deque, indexOfPresent, add /
flushOut / removeOldest as in src/Signal/SlidingWindow.hh:397-470, src/Signal/Normalization.cc:41-190) against
which the oracle's closed-form restatement must agree bit for bit.  CPU only."""
from collections import deque

import numpy as np
import pytest

INF = 2147483647


class SlidingWindowSim:
    """src/Signal/SlidingWindow.hh: elements are frame indices; front = latest."""

    def __init__(self, max_size, right):
        if max_size >= INF and right >= INF:
            right -= 1
        assert max_size > right
        self.max_future, self.max_past = right, max_size - right - 1
        self.d = deque()
        self.present = self.max_future
        self.removed = None

    def _remove_oldest(self):
        s = len(self.d) - (self.present + 1)
        self.removed = self.d.pop() if s > self.max_past else None

    def add(self, i):
        assert self.present == self.max_future
        self.d.appendleft(i)
        self._remove_oldest()

    def flush_out(self):
        if len(self.d) < self.present:
            self.present = len(self.d)
        if self.present >= -self.max_past:
            self.present -= 1
            self._remove_oldest()
        else:
            self.removed = None

    def out(self):
        return self.d[self.present] if 0 <= self.present < len(self.d) else None


def normalization_sim(x, kind, length, right):
    """Normalization::update / flush with MeanNormalization / MeanAndVarianceNormalization statistics (no FMA)."""
    T, D = x.shape
    win = SlidingWindowSim(INF if length < 0 else length, INF if right < 0 else right)
    s, sq, w, changed = np.zeros(D, np.float64), np.zeros(D, np.float64), 0.0, True
    mean, sd = np.zeros(D, np.float32), np.ones(D, np.float32)
    out = np.zeros_like(x)
    n_out = 0

    def emit():
        nonlocal changed, mean, sd, n_out
        t = win.out()
        if t is None:
            return False
        if changed:
            if w > 0:
                mean = (s / w).astype(np.float32)
                if kind == "mean-and-variance":
                    sd = np.sqrt((sq - s * s / w) / w).astype(np.float32)
                    sd[sd == 0] = 1.0
            changed = False
        v = x[t] - mean
        out[t] = v / sd if kind == "mean-and-variance" else v
        n_out += 1
        return True

    for i in range(T):
        win.add(i)
        a = x[i].astype(np.float64)
        s += a
        sq += a * a
        w += 1
        if win.removed is not None:
            r = x[win.removed].astype(np.float64)
            s -= r
            sq -= r * r
            w -= 1
        changed = True
        emit()
    while True:
        win.flush_out()
        if not emit():
            break
    assert n_out == T
    return out


@pytest.mark.parametrize("kind", ["mean", "mean-and-variance"])
@pytest.mark.parametrize("length,right", [(-1, -1), (5, 2), (7, 0), (9, 8), (201, 100), (3, 1)])
@pytest.mark.parametrize("T", [1, 2, 6, 40, 333])
def test_normalize_matches_literal_simulation(oracle, kind, length, right, T):
    rng = np.random.default_rng(T * 31 + length)
    x = (rng.standard_normal((T, 5)) * 3 + 1).astype(np.float32)
    want = normalization_sim(x, kind, length, right)
    got = oracle.normalize(x, kind=kind, length=length, right=right, use_fma=False)
    assert np.array_equal(got, want)


def test_normalize_segments_are_independent(oracle):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((90, 13)).astype(np.float32)
    fo = [0, 30, 31, 90]
    got = oracle.normalize(x, fo, "mean-and-variance")
    for a, b in zip(fo[:-1], fo[1:]):
        assert np.array_equal(got[a:b], oracle.normalize(x[a:b], kind="mean-and-variance"))
    # whole-segment mean normalisation: the output has (nearly) zero mean, unit variance
    assert abs(got[31:].mean(0)).max() < 1e-6 and abs(got[31:].std(0) - 1).max() < 1e-5
    assert np.all(got[30] == 0)  # a one-frame segment: x - mean = 0, sd = 0 -> 1


def test_splice_edges_replicate_and_order(oracle):
    x = np.arange(12, dtype=np.float32).reshape(6, 2)
    got = oracle.splice(x, 5, 2)  # past 2, future 2
    assert got.shape == (6, 10)
    assert list(got[0]) == [0, 1, 0, 1, 0, 1, 2, 3, 4, 5]          # t-2, t-1 replicated from frame 0; oldest first
    assert list(got[3]) == [2, 3, 4, 5, 6, 7, 8, 9, 10, 11]
    assert list(got[5]) == [6, 7, 8, 9, 10, 11, 10, 11, 10, 11]
    got = oracle.splice(x, 3, 0)  # causal window: right = 0
    assert list(got[1]) == [0, 1, 0, 1, 2, 3]
    two = oracle.splice(x, 5, 2, [0, 2, 6])  # segments do not leak into each other
    assert np.array_equal(two[:2], oracle.splice(x[:2], 5, 2)) and np.array_equal(two[2:], oracle.splice(x[2:], 5, 2))


def test_matmul_sequential_dot(oracle):
    rng = np.random.default_rng(1)
    M = rng.standard_normal((7, 33)).astype(np.float32)
    x = rng.standard_normal((20, 33)).astype(np.float32)
    got = oracle.matmul(M, x, use_fma=False)
    want = np.zeros((20, 7), np.float32)
    for t in range(20):
        for n in range(7):
            r = np.float32(0)
            for i in range(33):
                r = np.float32(r + np.float32(M[n, i] * x[t, i]))
            want[t, n] = r
    assert np.array_equal(got, want)
    np.testing.assert_allclose(oracle.matmul(M, x, use_fma=True), x.astype(np.float64) @ M.T.astype(np.float64), rtol=2e-5, atol=2e-5)
