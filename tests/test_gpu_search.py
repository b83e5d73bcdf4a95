"""GPU parity of the score consumer (rb_search_*: Search::LinearSearch over the dense score matrix) against the CPU
oracle, through the C ABI.  Word sequences, word-end frames (indices) and both scores must be BIT-IDENTICAL."""
import numpy as np
import pytest

from rasr_b200 import capi, mm, search, synth

pytestmark = pytest.mark.gpu


def same(a, b):
    return (list(a["words"]) == list(b["words"]) and list(a["times"]) == list(b["times"]) and
            np.array_equal(a["am"], b["am"]) and np.array_equal(a["lm"], b["lm"]))


@pytest.mark.parametrize("n_words,n_emis,T,seed", [(1, 8, 5, 0), (40, 32, 90, 2), (33, 32, 2, 3), (700, 256, 300, 5),
                                                   (1500, 256, 64, 6)])
def test_linear_search_bit_exact(oracle, n_words, n_emis, T, seed):
    lex = synth.lexicon(n_words, n_emis, seed=seed)
    rng = np.random.default_rng(seed)
    scores = (rng.random((T, n_emis)) * 25 + 2).astype(np.float32)
    got = search.LinearSearch(lex).decode(scores)
    assert len(got) == 1 and same(got[0], oracle.linear_search(lex, scores))


@pytest.mark.parametrize("n_words,min_s,max_s,env", [(2100, 3, 12, None), (300, 1, 3, None), (90, 2, 5, None),
                                                     (700, 3, 12, "RB_SEARCH_PER_WORD"), (300, 1, 3, "RB_SEARCH_PER_WORD"),
                                                     (700, 3, 12, "RB_SEARCH_FORCE_SCAN"), (700, 3, 12, "RB_SEARCH_FORCE_SCAN=2"),
                                                     (1200, 3, 9, "RB_SEARCH_NPT=16")])
def test_kernel_variants_bit_exact(oracle, monkeypatch, n_words, min_s, max_s, env):
    """the register-resident kernel at 2, 4, 8 and 16 states per thread, words of one and two states (the entry
    hypothesis is both predecessors), its book keeping forced onto the sequential replay, and the per-word kernel that
    serves lexicons the register kernel does not fit; FORCE_SCAN = never the unique-minimum shortcut (1: the scan over
    the record-breaking warps by every warp, 2: the warp-0 replay behind a second barrier)"""
    if env:
        monkeypatch.setenv(*(env.split("=") if "=" in env else (env, "1")))
    lex = synth.lexicon(n_words, 64, min_states=min_s, max_states=max_s, seed=21)
    rng = np.random.default_rng(21)
    fo = np.array([0, 120, 121, 200], np.int64)
    scores = (rng.random((200, 64)) * 25 + 2).astype(np.float32)
    got = search.LinearSearch(lex).decode(scores, fo)
    for u in range(3):
        assert same(got[u], oracle.linear_search(lex, scores[fo[u]:fo[u + 1]])), u


def irregular_lexicon(n_words, n_emis, seed, n_irregular, min_states=3, max_states=12):
    """synth.lexicon with a one-state silence-like word in front and n_irregular noise words spread over the lexicon"""
    lex = synth.lexicon(n_words, n_emis, min_states=min_states, max_states=max_states, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    reg = np.ones(n_words, np.uint8)
    reg[rng.choice(n_words, n_irregular, replace=False)] = 0
    # silence: one state on the silence-like transition model, no LM score, irregular
    lex["word_offsets"] = np.concatenate([[0], lex["word_offsets"] + 1]).astype(np.uint32)
    lex["state_emission"] = np.concatenate([[n_emis - 1], lex["state_emission"]]).astype(np.uint32)
    lex["state_tdp_model"] = np.concatenate([[2], lex["state_tdp_model"]]).astype(np.uint32)
    lex["unigram"] = np.concatenate([[0.0], lex["unigram"]]).astype(np.float32)
    lex["word_regular"] = np.concatenate([[0], reg]).astype(np.uint8)
    lex["single_word"] = True
    return lex


@pytest.mark.parametrize("n_words,n_irr,min_s,max_s,quantised", [(60, 3, 3, 12, False), (700, 9, 3, 12, False),
                                                                 (2100, 20, 3, 12, False), (300, 40, 1, 3, True),
                                                                 (50, 50, 2, 5, True)])
@pytest.mark.parametrize("env", [None, "RB_SEARCH_PER_WORD", "RB_SEARCH_FORCE_SCAN", "RB_SEARCH_FORCE_SCAN=2"])
def test_single_word_recognition_bit_exact(oracle, monkeypatch, n_words, n_irr, min_s, max_s, quantised, env):
    """the recognizer's default mode (src/Search/LinearSearch.cc:26-30): a second book of irregular-only sequences,
    irregular-chain entries, regular words restricted to one per sentence; lexicons held in shared memory and (2100
    words) in global memory; quantised scores make ties in both book-keeping scans; on the register-resident kernel
    (irregular word ends mirrored into slots, second scan replayed by every warp) and on the per-word kernel"""
    if env:
        monkeypatch.setenv(*(env.split("=") if "=" in env else (env, "1")))
    lex = irregular_lexicon(n_words, 64, 31, n_irr, min_s, max_s)
    rng = np.random.default_rng(31)
    fo = np.array([0, 150, 151, 260], np.int64)
    if quantised:
        lex["unigram"] = np.round(lex["unigram"] * 2) / 2
        scores = rng.integers(1, 6, (260, 64)).astype(np.float32)
    else:
        scores = (rng.random((260, 64)) * 25 + 2).astype(np.float32)
    ls = search.LinearSearch(lex)
    got = ls.decode(scores, fo)
    for u in range(3):
        want = oracle.linear_search(lex, scores[fo[u]:fo[u + 1]])
        assert same(got[u], want), u
        assert int(lex["word_regular"][want["words"]].sum()) <= 1
        assert same(ls.traceback(u), got[u])
    # the same lexicon in continuous mode: the flags are inert
    lex["single_word"] = False
    got = search.LinearSearch(lex).decode(scores, fo)
    for u in range(3):
        assert same(got[u], oracle.linear_search(lex, scores[fo[u]:fo[u + 1]])), u


@pytest.mark.parametrize("name", ["continuous", "continuous_scaled", "single_word", "single_word_noise", "single_word_ties"])
@pytest.mark.parametrize("env", [None, "RB_SEARCH_PER_WORD"])
def test_reference_golden(monkeypatch, reference_search_golden, name, env):
    """tracebacks written by the reference's own LinearSearch object code (tests/golden/make_golden_search.py)"""
    if env:
        monkeypatch.setenv(env, "1")
    c = reference_search_golden[name]
    got = search.LinearSearch(c).decode(c["scores"], c["frame_offsets"])
    ro = c["result_offsets"]
    for u in range(ro.size - 1):
        a, b = int(ro[u]), int(ro[u + 1])
        want = dict(words=c["words"][a:b], times=c["times"][a:b], am=c["am"][a:b], lm=c["lm"][a:b])
        assert same(got[u], want), (name, u)


def test_segments_are_independent_and_ties_resolve_like_the_reference(oracle):
    """several segments of different length in one call; quantised scores create exact ties between predecessors and
    between word ends, which must resolve as in the reference (first predecessor / first word in lexicon order)"""
    lex = synth.lexicon(120, 64, seed=9)
    lex["unigram"] = np.round(lex["unigram"] * 2) / 2
    rng = np.random.default_rng(9)
    fo = np.array([0, 37, 38, 180, 400], np.int64)
    scores = rng.integers(1, 6, (400, 64)).astype(np.float32)
    ls = search.LinearSearch(lex)
    got = ls.decode(scores, fo)
    for u in range(4):
        assert same(got[u], oracle.linear_search(lex, scores[fo[u]:fo[u + 1]])), u
        assert same(ls.traceback(u), got[u])  # the per-segment entry point agrees with the all-at-once one


def test_scores_from_the_gmm_scorer_on_device(oracle, diag):
    """config C5 in small: GMM scores stay on the device and feed the search; 24 segments x 250 frames"""
    import torch

    msd = synth.mixture_set()
    gmm = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-int")
    lex = synth.lexicon(1000, 256, seed=11)
    n_utt, T = 24, 250
    f = synth.features(n_utt * T, 39, seed=12)
    fo = np.arange(n_utt + 1, dtype=np.int64) * T
    d_in = torch.from_numpy(f).cuda()
    d_sc = torch.empty((n_utt * T, 256), dtype=torch.float32, device="cuda")
    gmm.score_dev(d_in, n_utt * T, d_sc)
    ls = search.LinearSearch(lex)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    got = ls.decode_dev(d_sc, 256, fo)
    ev[1].record()
    torch.cuda.synchronize()
    host_scores = d_sc.cpu().numpy()
    for u in (0, 7, 23):
        assert same(got[u], oracle.linear_search(lex, host_scores[fo[u]:fo[u + 1]])), u
    diag("search_c5_small", ms=ev[0].elapsed_time(ev[1]), frames=n_utt * T, words=int(sum(len(g["words"]) for g in got)))
    assert all(len(g["words"]) > 0 for g in got)


def test_rejects_bad_lexicon():
    lex = synth.lexicon(5, 8)
    bad = dict(lex)
    bad["word_offsets"] = np.array([0, 2, 2, 5, 7, 9], np.uint32)  # a word without states
    with pytest.raises(capi.RasrB200Error):
        search.LinearSearch(bad)
    bad = dict(lex)
    bad["entry_model"] = 9
    with pytest.raises(capi.RasrB200Error):
        search.LinearSearch(bad)


def test_c5_shard_at_full_size(oracle, diag):
    """BASELINE config C5, one GPU's shard (125 segments x 1000 frames, 1000-word lexicon): audio in, word sequences
    out, scores never leave the device.  Sampled segments equal the oracle's LinearSearch on the same score rows, and
    decoding a segment alone gives the same result as decoding it in the batch"""
    import torch

    from rasr_b200 import flow, pipeline

    samples, offs = synth.corpus(125, n_samples=160240, seed0=3000)
    fe = flow.FrontEnd()
    gmm = mm.GmmScorer(mm.MixtureSet.from_dict(synth.mixture_set()))
    lex = synth.lexicon(1000, 256)
    ls = search.LinearSearch(lex)
    fo = fe.count_frames(offs)
    T = int(fo[-1])
    d_samples = torch.from_numpy(samples).cuda()
    d_feats = torch.empty((T, 39), dtype=torch.float32, device="cuda")
    d_scores = torch.empty((T, 256), dtype=torch.float32, device="cuda")
    pipeline.score_utterances_dev(fe, gmm, d_samples, offs, d_feats, d_scores)
    torch.cuda.synchronize()  # the scorer ran on the front-end's stream, the search runs on its own
    got = ls.decode_dev(d_scores, 256, fo)
    assert len(got) == 125 and all(len(g["words"]) > 0 for g in got)
    for u in (0, 63, 124):
        rows = d_scores[int(fo[u]):int(fo[u + 1])].cpu().numpy()
        assert same(got[u], oracle.linear_search(lex, rows)), u
        assert same(got[u], ls.decode(rows)[0]), u
    # the one-call host entry point (16-bit PCM in, word sequences out) gives the same sequences
    one = pipeline.search_utterances(fe, gmm, ls, samples.astype(np.int16), offs, pcm_channels=1)
    assert len(one) == 125 and all(same(a, b) for a, b in zip(one, got))
    diag("search_c5_full", frames=T, words=int(sum(len(g["words"]) for g in got)))
