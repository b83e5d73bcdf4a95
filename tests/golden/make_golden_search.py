"""Tracebacks of the REFERENCE's own Search::LinearSearch (oracle/_ref/librasr_ref_search.so: src/Search/LinearSearch.cc
driven by oracle/refbuild/ref_search.cc from a Bliss lexicon file and the reference's configuration) ->
tests/golden/ref_search.npz.  Every case stores the flat lexicon (pyref.flat_lexicon, checked here against what the
reference's objects hand out), the score matrix and the reference's words / end frames / scores.  The CUDA search
(tests/test_gpu_search.py) and the oracle (tests/test_oracle_search.py) must reproduce them bit for bit.  Run from
the repo root in the build container (needs /root/reference):

    python tests/golden/make_golden_search.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_search.npz")
INF = np.inf
TDP = np.array([[INF, 0.0, 3.0, 0.0], [INF, 0.0, 3.0, 0.0], [0.7, 0.7, INF, 20.0],
                [3.0, 0.0, 30.0, 0.0], [2.5, 0.25, 28.0, 0.5]], np.float32)
# name: n_words, n_phonemes, n_emissions, frames of the segments, seed, options
CASES = {
    "continuous": (40, 12, 64, (150, 90), 2, dict(R=2)),
    "continuous_scaled": (30, 10, 48, (120,), 4, dict(P=2, R=2, lm_scale=12.5, tdp_scale=0.75)),
    "single_word": (40, 12, 64, (150, 90, 3), 12, dict(single_word=True, R=2)),
    "single_word_noise": (30, 8, 40, (160, 60), 18, dict(single_word=True, irregular=(0, 7, 8, 29), silence_first=True)),
    "single_word_ties": (18, 3, 6, (80, 80, 80), 21, dict(single_word=True, irregular=(2, 5), P=1, grid=True, max_len=2)),
}


def build_case(tmp, n_words, n_phonemes, n_emissions, seed, P=3, R=1, silence=True, silence_first=False, lm_scale=1.0,
               tdp_scale=1.0, single_word=False, irregular=(), grid=False, max_len=4):
    rng = np.random.default_rng(seed)
    words = [[int(p) for p in rng.integers(0, n_phonemes, rng.integers(1, max_len + 1))] for _ in range(n_words)]
    emission_of = rng.integers(0, n_emissions - 1, (n_phonemes, P)).astype(np.int32)
    unigram = (-np.log(rng.dirichlet(np.ones(n_words)))).astype(np.float32)
    tdp = TDP
    if grid:  # everything on multiples of 0.5: exact ties between predecessors and between word ends
        unigram = (np.round(unigram * 2) / 2).astype(np.float32)
        tdp = np.array([[INF, 0, 1, 0], [INF, 0, 1, 0], [0.5, 0.5, INF, 1], [1, 0, 2, 0], [1, 0, 2, 0]], np.float32)
    lex_file = os.path.join(tmp, "lexicon_%d.xml" % seed)
    pyref.write_lexicon(lex_file, n_phonemes, words, silence=silence, silence_first=silence_first, irregular=irregular)
    kw = dict(states_per_phone=P, state_repetitions=R, lm_scale=lm_scale, tdp_scale=tdp_scale, single_word=single_word)
    ref = pyref.LinearSearch(lex_file, emission_of, n_emissions - 1, n_emissions, tdp, unigram, scratch_dir=tmp, **kw)
    flat = pyref.flat_lexicon(words, emission_of, n_emissions - 1, tdp, unigram, silence=silence,
                              silence_first=silence_first, irregular=irregular, **kw)
    assert np.array_equal(ref.order(), flat["word"]) and np.array_equal(ref.tdps(), flat["tdp"])
    for i in range(flat["word"].size):
        e, m = ref.states(i)
        a, b = int(flat["word_offsets"][i]), int(flat["word_offsets"][i + 1])
        assert np.array_equal(e, flat["state_emission"][a:b]) and np.array_equal(m, flat["state_tdp_model"][a:b])
    return ref, flat, rng, grid


def main():
    pyref.build()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (n_words, n_phonemes, n_emissions, frames, seed, kw) in CASES.items():
            ref, flat, rng, grid = build_case(tmp, n_words, n_phonemes, n_emissions, seed, **kw)
            T = int(sum(frames))
            if grid:
                scores = (rng.integers(2, 9, (T, n_emissions)) * 0.5).astype(np.float32)
            else:
                scores = (rng.random((T, n_emissions)) * 25 + 2).astype(np.float32)
            fo = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
            for k in ("word_offsets", "state_emission", "state_tdp_model", "tdp", "unigram", "word", "word_regular"):
                out["%s/%s" % (name, k)] = flat[k]
            out["%s/single_word" % name] = np.array(int(flat["single_word"]), np.int32)
            out["%s/scores" % name] = scores
            out["%s/frame_offsets" % name] = fo
            wo, res = [0], [[], [], [], []]
            by_word = {int(k): i for i, k in enumerate(flat["word"])}  # lemma number -> flat entry (first pronunciation)
            for u in range(len(frames)):
                words, times, am, lm, _ = ref.run(scores[fo[u]:fo[u + 1]])
                res[0] += [by_word[int(w)] for w in words]
                res[1] += list(times)
                res[2] += list(am)
                res[3] += list(lm)
                wo.append(len(res[0]))
            out["%s/result_offsets" % name] = np.asarray(wo, np.int64)
            out["%s/words" % name] = np.asarray(res[0], np.uint32)
            out["%s/times" % name] = np.asarray(res[1], np.int32)
            out["%s/am" % name] = np.asarray(res[2], np.float32)
            out["%s/lm" % name] = np.asarray(res[3], np.float32)
            ref.close()
            print(name, "segments", len(frames), "words", wo[-1])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
