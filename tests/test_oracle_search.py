"""Oracle pin for the score consumer (Search::LinearSearch).  The reference has no unit test for Search: parity
unpinned by the reference, pinned here by a LITERAL Python transcription of LinearSearch::feed / bookKeeping /
getCurrentBestSentence (objects, back pointers, the hypTmp vector that is reused across words exactly as in
src/Search/LinearSearch.cc:233-468) with numpy float32 arithmetic, which the flat-array oracle must match bit for
bit.  CPU only."""
import numpy as np
import pytest

from rasr_b200 import synth

F = np.float32
FLT_MAX = np.finfo(np.float32).max


class Book:
    def __init__(self):
        self.score, self.lmScore, self.word, self.bkp, self.time = FLT_MAX, F(0), None, None, 0


class Hypo:
    def __init__(self):
        self.score, self.lmScore, self.bkp, self.mixture = FLT_MAX, F(0), None, None


def linear_search_literal(lex, scores):
    W = len(lex["word_offsets"]) - 1
    tdp = np.asarray(lex["tdp"], np.float32).reshape(-1, 4)
    words = []
    for w in range(W):
        a, b = int(lex["word_offsets"][w]), int(lex["word_offsets"][w + 1])
        hyp = [Hypo() for _ in range(b - a + 1)]
        for i in range(a, b):
            hyp[i - a + 1].mixture = (int(lex["state_emission"][i]), int(lex["state_tdp_model"][i]))
        words.append(hyp)
    book = []
    hypTmp = []
    with np.errstate(over="ignore"):
        for t in range(1, scores.shape[0] + 1):
            for w, hyp in enumerate(words):
                last = book[-1] if book else None
                hyp[0].bkp = last
                if last is not None:
                    hyp[0].lmScore = F(F(lex["unigram"][w]) + last.lmScore)
                    hyp[0].score = last.score
                else:
                    hyp[0].lmScore = F(lex["unigram"][w])
                    hyp[0].score = F(0)
                hyp[0].score = F(hyp[0].score + hyp[0].lmScore)
                while len(hypTmp) < len(hyp):  # resize keeps the old elements
                    hypTmp.append(Hypo())
                del hypTmp[len(hyp):]
                for sta in range(1, len(hyp)):
                    hypTmp[sta].score, hypTmp[sta].lmScore = FLT_MAX, F(0)
                    for pre in range(sta - 2 if sta >= 2 else 0, sta + 1):
                        model = hyp[pre].mixture[1] if pre != 0 else int(lex["entry_model"])
                        sco = F(hyp[pre].score + tdp[model, sta - pre])
                        if sco < hypTmp[sta].score:
                            hypTmp[sta].score, hypTmp[sta].bkp, hypTmp[sta].lmScore = sco, hyp[pre].bkp, hyp[pre].lmScore
                for sta in range(1, len(hyp)):
                    hyp[sta].bkp = hypTmp[sta].bkp
                    hyp[sta].score = F(hypTmp[sta].score + scores[t - 1, hyp[sta].mixture[0]])
                    hyp[sta].lmScore = hypTmp[sta].lmScore
            nb = Book()
            for w, hyp in enumerate(words):
                h = hyp[-1]
                tmp = F(h.score + tdp[h.mixture[1], 3])
                if tmp < F(nb.score + nb.lmScore):
                    nb.score, nb.lmScore, nb.bkp, nb.word, nb.time = F(tmp - h.lmScore), h.lmScore, h.bkp, w, t
            if nb.score != FLT_MAX:
                book.append(nb)
    out = []
    b = book[-1] if book else None
    while b is not None:
        out.append((b.word, b.time, b.score, b.lmScore))
        b = b.bkp
    out.reverse()
    return out


@pytest.mark.parametrize("n_words,T,seed", [(1, 5, 0), (7, 40, 1), (40, 90, 2), (33, 2, 3), (64, 130, 4)])
def test_flat_oracle_matches_literal_transcription(oracle, n_words, T, seed):
    lex = synth.lexicon(n_words, 32, seed=seed)
    rng = np.random.default_rng(seed)
    scores = (rng.random((T, 32)) * 25 + 2).astype(np.float32)
    want = linear_search_literal(lex, scores)
    got = oracle.linear_search(lex, scores)
    assert [w for w, _, _, _ in want] == list(got["words"])
    assert [t for _, t, _, _ in want] == list(got["times"])
    assert np.array_equal(np.array([s for _, _, s, _ in want], np.float32), got["am"])
    assert np.array_equal(np.array([s for _, _, _, s in want], np.float32), got["lm"])


def test_forced_path_kat(oracle):
    """Two words of two states; the scores make word 1 then word 0 the only cheap path: the traceback must name them
    with the right end frames and Book::score = sum of emissions + transition scores (LM kept apart)."""
    lex = dict(word_offsets=[0, 2, 4], state_emission=[0, 1, 2, 3], state_tdp_model=[0, 0, 0, 0],
               tdp=np.array([[3, 0, 30, 0], [1e30, 0, 30, 0]], np.float32), entry_model=1,
               unigram=np.array([0.5, 0.25], np.float32))
    big = 100.0
    sc = np.full((4, 4), big, np.float32)
    sc[0, 2] = sc[1, 3] = sc[2, 0] = sc[3, 1] = 1.0  # word 1: states e2, e3 at frames 1, 2; word 0: e0, e1 at frames 3, 4
    r = oracle.linear_search(lex, sc)
    assert list(r["words"]) == [1, 0] and list(r["times"]) == [2, 4]
    assert r["am"][0] == np.float32(2.0) and r["lm"][0] == np.float32(0.25)
    assert r["am"][1] == np.float32(4.0) and r["lm"][1] == np.float32(0.75)


@pytest.mark.parametrize("name", ["continuous", "continuous_scaled", "single_word", "single_word_noise", "single_word_ties"])
def test_oracle_reproduces_reference_golden(oracle, reference_search_golden, name):
    c = reference_search_golden[name]
    fo, ro = c["frame_offsets"], c["result_offsets"]
    for u in range(fo.size - 1):
        got = oracle.linear_search(c, c["scores"][fo[u]:fo[u + 1]])
        a, b = int(ro[u]), int(ro[u + 1])
        assert np.array_equal(got["words"], c["words"][a:b]) and np.array_equal(got["times"], c["times"][a:b])
        assert np.array_equal(got["am"], c["am"][a:b]) and np.array_equal(got["lm"], c["lm"][a:b])


def test_single_word_with_a_long_irregular_word_first(oracle):
    """the scratch hypotheses of feed() (hypTmp, src/Search/LinearSearch.cc:238,303) start with a null back pointer; the
    irregular book-keeping scan reads the back pointer of word ends that were never reached in the first frames"""
    lex = synth.lexicon(50, 64, seed=3)
    lex["word_regular"] = (np.arange(50) % 7 != 0).astype(np.uint8)  # word 0: irregular, several states
    lex["single_word"] = True
    rng = np.random.default_rng(0)
    scores = (rng.random((120, 64)) * 25 + 2).astype(np.float32)
    r = oracle.linear_search(lex, scores)
    assert len(r["words"]) >= 1 and int(lex["word_regular"][r["words"]].sum()) <= 1
