"""Run in a process of its own by tests/test_gpu_zzz_reference_host.py: the Search::SearchAlgorithm adapter
(adapters/B200LinearSearch.cc, compiled against the reference's headers) next to the reference's own
Search::LinearSearch, both set up by the reference's code from the same lexicon file, acoustic model parts and language
model (oracle/refbuild/ref_search.cc) and fed the same scorer objects frame by frame.  Every item of the two
tracebacks must agree bit for bit.  Prints one JSON line."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import pyref  # noqa: E402

pyref.load_search_adapter()
from make_golden_search import CASES, TDP, INF  # noqa: E402


def main():
    report = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (n_words, n_phonemes, n_emissions, frames, seed, kw) in CASES.items():
            kw = dict(kw)
            rng = np.random.default_rng(seed)
            P, R, max_len = kw.pop("P", 3), kw.pop("R", 1), kw.pop("max_len", 4)
            grid, irregular = kw.pop("grid", False), kw.pop("irregular", ())
            silence_first = kw.pop("silence_first", False)
            words = [[int(p) for p in rng.integers(0, n_phonemes, rng.integers(1, max_len + 1))] for _ in range(n_words)]
            if name == "continuous":  # several pronunciations per lemma, too
                words = [[w, w + [0]] if k % 4 == 0 else w for k, w in enumerate(words)]
            emission_of = rng.integers(0, n_emissions - 1, (n_phonemes, P)).astype(np.int32)
            unigram = (-np.log(rng.dirichlet(np.ones(n_words)))).astype(np.float32)
            tdp = TDP
            if grid:
                unigram = (np.round(unigram * 2) / 2).astype(np.float32)
                tdp = np.array([[INF, 0, 1, 0], [INF, 0, 1, 0], [0.5, 0.5, INF, 1], [1, 0, 2, 0], [1, 0, 2, 0]], np.float32)
            lex_file = os.path.join(tmp, name + ".xml")
            pyref.write_lexicon(lex_file, n_phonemes, words, silence_first=silence_first, irregular=irregular)
            args = (lex_file, emission_of, n_emissions - 1, n_emissions, tdp, unigram)
            opts = dict(states_per_phone=P, state_repetitions=R, scratch_dir=tmp, **kw)
            ref = pyref.LinearSearch(*args, **opts)
            b200 = pyref.LinearSearch(*args, adapter=True, **opts)
            n_items = 0
            for T in frames:
                if grid:
                    scores = (rng.integers(2, 9, (T, n_emissions)) * 0.5).astype(np.float32)
                else:
                    scores = (rng.random((T, n_emissions)) * 25 + 2).astype(np.float32)
                ref.run(scores)
                b200.run(scores)
                want, got = ref.items(), b200.items()
                for a, b in zip(want, got):
                    if not np.array_equal(a, b):
                        print(json.dumps(dict(ok=False, case=name, want=[x.tolist() for x in want],
                                              got=[x.tolist() for x in got])))
                        return 1
                n_items += len(want[0])
            report[name] = n_items
            ref.close()
            b200.close()
        report["whole_path"] = whole_path(tmp)
        if report["whole_path"] < 0:
            return 1
    print(json.dumps(dict(ok=True, items=report)))
    return 0


def whole_path(tmp):
    """The recognizer's loop inside the reference host, twice: the reference's batch-diagonal-maximum-float scorer
    feeding the reference's LinearSearch, and the adapter b200-batch-float (strict arithmetic) feeding the adapter
    B200::LinearSearch -- which recognises the b200 scorer objects and takes whole score rows from them.  Feature
    vectors in, tracebacks out; every item identical."""
    from oracle import pyoracle as o
    from rasr_b200 import synth
    S = pyref.search_lib()
    n_mix, n_phonemes, P = 48, 12, 3
    msd = synth.mixture_set(dim=39, n_mixtures=n_mix, densities_per_mixture=8, seed=17)
    ms = o.MixtureSet(**msd)
    rng = np.random.default_rng(23)
    words = [[int(p) for p in rng.integers(0, n_phonemes, rng.integers(1, 5))] for _ in range(60)]
    emission_of = rng.integers(0, n_mix - 1, (n_phonemes, P)).astype(np.int32)
    unigram = (-np.log(rng.dirichlet(np.ones(60)))).astype(np.float32)
    lex_file = os.path.join(tmp, "whole.xml")
    pyref.write_lexicon(lex_file, n_phonemes, words, irregular=(3, 11))
    total = 0
    for single_word in (False, True):
        opts = dict(states_per_phone=P, scratch_dir=tmp, single_word=single_word)
        args = (lex_file, emission_of, n_mix - 1, n_mix, TDP, unigram)
        theirs = (pyref.FeatureScorer(ms, "batch-diagonal-maximum-float", library=S), pyref.LinearSearch(*args, **opts))
        ours = (pyref.FeatureScorer(ms, "b200-batch-float", {"fma-contraction": "false"}, library=S),
                pyref.LinearSearch(*args, adapter=True, **opts))
        for T in (300, 1, 77):
            feats = synth.features(T, 39, seed=100 + T)
            theirs[1].run_features(theirs[0], feats)
            ours[1].run_features(ours[0], feats)
            want, got = theirs[1].items(), ours[1].items()
            for a, b in zip(want, got):
                if not np.array_equal(a, b):
                    print(json.dumps(dict(ok=False, case="whole_path", single_word=single_word, T=T,
                                          want=[x.tolist() for x in want], got=[x.tolist() for x in got])))
                    return -1
            total += len(want[0])
        # the adapter search recognised the adapter scorer's objects: all 378 frames came as dense rows
        import ctypes as C
        stats = (C.c_ulonglong * 2)()
        pyref.load_search_adapter().b200_search_adapter_statistics(stats)
        if stats[0] != 300 + 1 + 77 or stats[1] != 0:
            print(json.dumps(dict(ok=False, case="whole_path", dense_rows=int(stats[0]), score_calls=int(stats[1]))))
            return -1
    return total


if __name__ == "__main__":
    sys.exit(main())
