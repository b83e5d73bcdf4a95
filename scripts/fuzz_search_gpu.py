"""One-off fuzz of the CUDA search against the oracle: random lexica (word count, state counts, irregular words, both
recognition modes, quantised scores for ties), random segment lengths.  Usage on the GPU box: python scripts/fuzz_search_gpu.py [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as o  # noqa: E402
from rasr_b200 import search, synth  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(2024)
bad = 0
for it in range(n_cases):
    n_words = int(rng.choice([1, 2, 5, 17, 60, 200, 700, 1500]))
    n_emis = int(rng.choice([8, 64, 256]))
    lo = int(rng.integers(1, 4))
    hi = lo + int(rng.integers(0, 9))
    lex = synth.lexicon(n_words, n_emis, min_states=lo, max_states=hi, seed=int(rng.integers(1 << 30)))
    single = bool(rng.integers(0, 2))
    n_irr = int(rng.integers(0, min(n_words, 30) + 1))
    reg = np.ones(n_words, np.uint8)
    reg[rng.choice(n_words, n_irr, replace=False)] = 0
    lex["word_regular"] = reg
    lex["single_word"] = single
    quantised = bool(rng.integers(0, 2))
    fo = np.concatenate([[0], np.cumsum(rng.integers(1, 120, int(rng.integers(1, 5))))]).astype(np.int64)
    T = int(fo[-1])
    if quantised:
        lex["unigram"] = (np.round(lex["unigram"] * 2) / 2).astype(np.float32)
        scores = rng.integers(1, 6, (T, n_emis)).astype(np.float32)
    else:
        scores = (rng.random((T, n_emis)) * 25 + 2).astype(np.float32)
    env = rng.choice(["", "RB_SEARCH_PER_WORD", "RB_SEARCH_FORCE_SCAN", "RB_SEARCH_FORCE_SCAN=2"])
    for k in ("RB_SEARCH_PER_WORD", "RB_SEARCH_FORCE_SCAN"):
        os.environ.pop(k, None)
    if env:
        k, _, v = env.partition("=")
        os.environ[k] = v or "1"
    got = search.LinearSearch(lex).decode(scores, fo)
    for u in range(fo.size - 1):
        want = o.linear_search(lex, scores[fo[u]:fo[u + 1]])
        if not all(np.array_equal(got[u][k], want[k]) for k in ("words", "times", "am", "lm")):
            bad += 1
            print("MISMATCH case", it, "segment", u, dict(n_words=n_words, n_emis=n_emis, lo=lo, hi=hi, single=single,
                                                          n_irr=n_irr, quantised=quantised, env=str(env)), flush=True)
            break
print("fuzz done: %d cases, %d mismatches" % (n_cases, bad))
