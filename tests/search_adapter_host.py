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
    print(json.dumps(dict(ok=True, items=report)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
