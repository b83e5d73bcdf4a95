"""Pins oracle/search_oracle.cc to the REFERENCE's own Search::LinearSearch object code
(oracle/_ref/librasr_ref_search.so: src/Search/LinearSearch.cc with the Bliss lexicon parser, Am::ClassicStateModel,
Am::ScaledTransitionModel, Am::LutStateTying and Lm::LanguageModelScaling compiled from where they lie; the stand-ins
are listed in oracle/refbuild/ref_search.cc).  The reference reads a lexicon FILE and its configuration; the flat
arrays the oracle and the CUDA search take are derived from the same description by pyref.flat_lexicon and checked
against what the reference's objects hand out, then words / end frames / acoustic and LM scores of the traceback
must agree bit for bit.  CPU only."""
import os

import numpy as np
import pytest

from oracle import pyref

pytestmark = pytest.mark.skipif(not (os.path.exists(pyref.search_path()) or os.path.isdir(pyref.REFERENCE)),
                                reason="oracle/_ref/librasr_ref_search.so not built and no reference checkout")

INF = np.inf
TDP_DEFAULT = np.array([[INF, 0.0, 3.0, 0.0], [INF, 0.0, 3.0, 0.0], [0.7, 0.7, INF, 20.0],
                        [3.0, 0.0, 30.0, 0.0], [2.5, 0.25, 28.0, 0.5]], np.float32)


def make_case(tmp_path, n_words, n_phonemes, n_emissions, seed, P=3, R=1, silence=True, silence_first=False,
              lm_scale=1.0, tdp_scale=1.0, tdp=TDP_DEFAULT, multi=False, duplicates=0, max_len=4, grid=False,
              single_word=False, irregular=()):
    rng = np.random.default_rng(seed)
    words = []
    for k in range(n_words):
        mk = lambda: [int(p) for p in rng.integers(0, n_phonemes, rng.integers(1, max_len + 1))]
        first = mk()
        # a lemma's pronunciations must differ (the lexicon parser rejects duplicates): the variant is one phoneme longer
        words.append([first, first + mk()[:1]] if multi and k % 3 == 0 else first)
    for d in range(duplicates):  # same pronunciation and LM score under a second lemma: the first one must win ties
        words.append(words[d])
    emission_of = rng.integers(0, n_emissions - 1, (n_phonemes, P)).astype(np.int32)
    p = rng.dirichlet(np.ones(n_words))
    unigram = (-np.log(p)).astype(np.float32)
    if grid:
        unigram = (np.round(unigram * 2) / 2).astype(np.float32)
    unigram = np.concatenate([unigram, unigram[:duplicates]])
    lex_file = str(tmp_path / ("lexicon_%d.xml" % seed))
    pyref.write_lexicon(lex_file, n_phonemes, words, silence=silence, silence_first=silence_first, irregular=irregular)
    kw = dict(states_per_phone=P, state_repetitions=R, lm_scale=lm_scale, tdp_scale=tdp_scale, single_word=single_word)
    ref = pyref.LinearSearch(lex_file, emission_of, n_emissions - 1, n_emissions, tdp, unigram,
                             scratch_dir=str(tmp_path), **kw)
    flat = pyref.flat_lexicon(words, emission_of, n_emissions - 1, tdp, unigram, silence=silence,
                              silence_first=silence_first, irregular=irregular, **kw)
    return ref, flat


def check_flat(ref, flat):
    assert np.array_equal(ref.order(), flat["word"])
    assert np.array_equal(ref.tdps(), flat["tdp"])
    for i in range(flat["word"].size):
        e, m = ref.states(i)
        a, b = int(flat["word_offsets"][i]), int(flat["word_offsets"][i + 1])
        assert np.array_equal(e, flat["state_emission"][a:b].astype(np.int32)), i
        assert np.array_equal(m, flat["state_tdp_model"][a:b].astype(np.int32)), i


def check_run(oracle, ref, flat, scores):
    words, times, am, lm, fin = ref.run(scores)
    got = oracle.linear_search(flat, scores)
    assert np.array_equal(flat["word"][got["words"]], words)
    assert np.array_equal(got["times"], times)
    assert np.array_equal(got["am"], am) and np.array_equal(got["lm"], lm)
    if len(words):
        assert fin[0] == am[-1] and fin[1] == lm[-1]  # closing item: last book entry + sentence end score (0 here)
    if flat["single_word"]:
        assert int(flat["word_regular"][got["words"]].sum()) <= 1
    return words


CASES = [
    # n_words, n_phonemes, n_emissions, T, seed, kwargs
    (1, 2, 8, 12, 0, {}),
    (7, 5, 16, 60, 1, {}),
    (40, 12, 64, 150, 2, dict(R=2)),
    (25, 8, 32, 90, 3, dict(P=1, R=1, silence_first=True)),
    (30, 10, 48, 120, 4, dict(P=2, R=2, lm_scale=12.5, tdp_scale=0.75)),
    (20, 6, 32, 80, 5, dict(silence=False)),
    (24, 9, 40, 100, 6, dict(multi=True)),
    (16, 4, 24, 70, 7, dict(duplicates=5)),
    (64, 20, 128, 300, 8, dict(R=2, max_len=6)),
    (12, 6, 20, 2, 9, {}),      # shorter than most words: few or no book entries
    (5, 3, 10, 1, 10, dict(P=1)),
    # single-word recognition, the reference's default: irregular* regular irregular*
    (7, 5, 16, 60, 11, dict(single_word=True)),
    (40, 12, 64, 150, 12, dict(single_word=True, R=2)),
    (25, 8, 32, 90, 13, dict(single_word=True, P=1, silence_first=True)),
    (30, 10, 48, 200, 14, dict(single_word=True, P=2, R=2, lm_scale=12.5, tdp_scale=0.75)),
    (20, 6, 32, 80, 15, dict(single_word=True, silence=False)),   # no irregular word at all
    (16, 4, 24, 70, 16, dict(single_word=True, duplicates=5, multi=True)),
    (12, 6, 20, 2, 17, dict(single_word=True)),
    (30, 8, 40, 160, 18, dict(single_word=True, irregular=(0, 7, 8, 29))),   # noise words besides silence
    (30, 8, 40, 160, 19, dict(single_word=False, irregular=(0, 7, 8, 29))),  # the flag is inert without the mode
    (10, 4, 16, 120, 20, dict(single_word=True, irregular=tuple(range(10)))),  # nothing but irregular words
    # the first pronunciations are long irregular words, no silence: word ends that are unreachable in the first frames
    # are looked at by the irregular book keeping with a null back pointer
    (20, 6, 32, 80, 22, dict(single_word=True, irregular=(0, 1), silence=False, max_len=5)),
]


@pytest.mark.parametrize("n_words,n_phonemes,n_emissions,T,seed,kw", CASES)
def test_oracle_matches_reference_linear_search(oracle, tmp_path, n_words, n_phonemes, n_emissions, T, seed, kw):
    ref, flat = make_case(tmp_path, n_words, n_phonemes, n_emissions, seed, **kw)
    try:
        check_flat(ref, flat)
        rng = np.random.default_rng(100 + seed)
        for rep in range(3):  # the same search object re-used: restart() between segments
            scores = (rng.random((T, n_emissions)) * 25 + 2).astype(np.float32)
            check_run(oracle, ref, flat, scores)
    finally:
        ref.close()


@pytest.mark.parametrize("single_word", [False, True])
def test_ties_resolve_to_the_first_pronunciation(oracle, tmp_path, single_word):
    """Scores, transition and LM scores on a coarse grid (multiples of 0.5, exact in float32) make equal path scores
    common, and six lemmata repeat earlier ones: both sides must keep the first of equals (strict < in feed and
    bookKeeping, src/Search/LinearSearch.cc:321,405)."""
    tdp = np.array([[INF, 0, 1, 0], [INF, 0, 1, 0], [0.5, 0.5, INF, 1], [1, 0, 2, 0], [1, 0, 2, 0]], np.float32)
    ref, flat = make_case(tmp_path, 12, 3, 6, 21, P=1, R=1, tdp=tdp, duplicates=6, max_len=2, grid=True,
                          single_word=single_word, irregular=(2, 5) if single_word else ())
    try:
        check_flat(ref, flat)
        rng = np.random.default_rng(5)
        seen = set()
        for rep in range(6):
            scores = (rng.integers(2, 9, (80, 6)) * 0.5).astype(np.float32)
            seen.update(int(w) for w in check_run(oracle, ref, flat, scores))
        assert not (seen & set(range(12, 18)))  # a repeated lemma never wins against its first occurrence
    finally:
        ref.close()
