"""Fused audio -> scores (config C3 shape at a small size) through the C ABI."""
import numpy as np
import pytest

from rasr_b200 import flow, mm, pipeline, synth

pytestmark = pytest.mark.gpu


def test_pipeline_equals_frontend_then_scorer(oracle, diag):
    samples, offs = synth.corpus(6, n_samples=16240)
    fe = flow.FrontEnd()
    msd = synth.mixture_set()
    gmm = mm.GmmScorer(mm.MixtureSet.from_dict(msd))
    scores, feats, fo = pipeline.score_utterances(fe, gmm, samples, offs, want_feats=True)
    r = fe.process(samples, offs)
    assert np.array_equal(feats, r["feats"])
    assert np.array_equal(scores, gmm.score(r["feats"]))
    # end to end against the oracle: features within tolerance => scores within tolerance
    oms = oracle.MixtureSet(**msd)
    T0 = int(fo[1])
    ofeats = oracle.mfcc(oracle.frontend_cfg(), samples[:offs[1]])["feats"]
    want = oracle.gmm_batch_float(oms, ofeats)
    rel = np.abs(scores[:T0] - want) / np.abs(want)
    diag("pipeline_c3_small", max_rel=rel.max())
    assert rel.max() < 1e-4
    # scoring the GPU features with the oracle is bit-identical (isolates the scorer)
    assert np.array_equal(scores[:T0], oracle.gmm_batch_float(oms, feats[:T0]))
    # 16-bit PCM input (the synthetic audio is integer valued): identical scores
    s16, _ = pipeline.score_utterances(fe, gmm, samples.astype(np.int16), offs, pcm_channels=1)
    assert np.array_equal(s16, scores)


def test_launch_counter_moves():
    from rasr_b200 import capi

    before = capi.launch_count()
    flow.FrontEnd().process(synth.utterance(4000))
    assert capi.launch_count() >= before + 2


def test_audio_to_nn_scores_pipeline(oracle, diag):
    """audio -> MFCC -> segment CMVN -> 11-frame window (429 dims) -> Nn scores in one call equals the stages run one
    by one through their own entry points, and the oracle chain within the bf16 tolerance of test_gpu_nn.py."""
    from rasr_b200 import nn, postproc

    samples, offs = synth.corpus(5, n_samples=24240)
    fe = flow.FrontEnd()
    pp = postproc.PostProcessor(39, "mean-and-variance", splice=(11, 5))
    net = synth.network(dims=(429, 512, 1000), seed=3)
    sc = nn.NnScorer(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 1.0, "bf16")
    scores, fo = pipeline.nn_score_utterances(fe, pp, sc, samples, offs)
    r = fe.process(samples, offs)
    x = pp.process(r["feats"], r["frame_offsets"])
    assert x.shape[1] == 429
    assert np.array_equal(scores, sc.score(x))
    # oracle chain on the first utterance
    T0 = int(fo[1])
    of = oracle.mfcc(oracle.frontend_cfg(), samples[:offs[1]])["feats"]
    ox = oracle.splice(oracle.normalize(of, kind="mean-and-variance"), 11, 5)
    want = oracle.nn_scores(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 1.0, ox,
                            mode=oracle.NN_BF16)
    err = np.abs(scores[:T0] - want).max() / np.abs(want).max()
    diag("pipeline_audio_to_nn", err=err)
    assert err < 3e-3
    # without a post-processor the network must take the raw feature dimension
    net39 = synth.network(dims=(39, 64, 100), seed=4)
    sc39 = nn.NnScorer(net39["dims"], net39["acts"], net39["weights"], net39["biases"], net39["log_prior"], 1.0, "bf16")
    s2, _ = pipeline.nn_score_utterances(fe, None, sc39, samples, offs)
    assert np.array_equal(s2, sc39.score(r["feats"]))
    with pytest.raises(Exception):
        pipeline.nn_score_utterances(fe, None, sc, samples, offs)  # 39-dim features into a 429-dim network


def test_audio_to_nn_scores_with_disregarded_classes():
    """a class mapping with -1 entries makes the score rows wider than the network output (n_classes > n_outputs):
    the one-call pipeline must size its slabs and the host stride by the class count.  More than one slab (> 16384
    frames) so the second slab's offset is exercised too."""
    from rasr_b200 import nn

    samples, offs = synth.corpus(9, n_samples=320240)  # 9 x 2001 frames -> two slabs
    fe = flow.FrontEnd()
    net = synth.network(dims=(39, 64, 40), seed=5)
    sc = nn.NnScorer(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 1.0, "bf16")
    rng = np.random.default_rng(11)
    mapping = np.full(57, -1, np.int32)
    mapping[rng.permutation(57)[:40]] = rng.permutation(40).astype(np.int32)
    sc.set_class_mapping(mapping)
    assert sc.n_emissions == 57 and sc.n_outputs == 40
    scores, fo = pipeline.nn_score_utterances(fe, None, sc, samples, offs)
    assert scores.shape == (int(fo[-1]), 57) and int(fo[-1]) > 16384
    r = fe.process(samples, offs)
    want = sc.score(r["feats"])
    assert np.array_equal(scores, want)
    assert (scores[:, mapping < 0] == np.finfo(np.float32).max).all()
    assert (scores[:, mapping >= 0] < 1e30).all()


def test_features_through_a_feature_cache(tmp_path):
    """what a two-pass RASR setup does: the extraction run writes the features to a cache archive (one entry per
    segment, timestamps included), the recognition run reads them back and scores them"""
    from rasr_b200 import cache
    samples, offs = synth.corpus(3, n_samples=8240)
    fe = flow.FrontEnd()
    r = fe.process(samples, offs)
    fo = r["frame_offsets"]
    with cache.FileArchive(tmp_path / "mfcc.cache", "w") as a:
        for u in range(3):
            sl = slice(int(fo[u]), int(fo[u + 1]))
            cache.write_features(a, "corpus/rec/%d" % u, r["feats"][sl], np.stack([r["t_start"][sl], r["t_end"][sl]], 1),
                                 {"datatype": "vector-f32", "sample-rate": "100"}, compress=(u == 1))
    gmm = mm.GmmScorer(mm.MixtureSet.from_dict(synth.mixture_set()))
    want = gmm.score(r["feats"])
    with cache.FileArchive(tmp_path / "mfcc.cache") as a:
        for u in range(3):
            f, t, atts = cache.read_features(a, "corpus/rec/%d" % u)
            sl = slice(int(fo[u]), int(fo[u + 1]))
            assert np.array_equal(f, r["feats"][sl]) and np.array_equal(t[:, 0], r["t_start"][sl])
            assert np.array_equal(gmm.score(f), want[sl])


def test_c3_shard_at_full_size(oracle, diag):
    """BASELINE config C3, one GPU's shard (125 utterances x 1000 frames): checked through size-independent properties --
    every utterance's rows equal the rows of the same utterance processed alone (sharding / batching cannot change a
    result), 16-bit PCM input equals f32 input, the scores of sampled utterances equal the oracle scorer on the GPU
    features bit for bit and the oracle pipeline within tolerance"""
    samples, offs = synth.corpus(125, n_samples=160240, seed0=3000)
    fe = flow.FrontEnd()
    msd = synth.mixture_set()
    gmm = mm.GmmScorer(mm.MixtureSet.from_dict(msd))
    scores, feats, fo = pipeline.score_utterances(fe, gmm, samples, offs, want_feats=True)
    assert scores.shape == (125000, 256) and np.isfinite(scores).all()
    s16, _ = pipeline.score_utterances(fe, gmm, samples.astype(np.int16), offs, pcm_channels=1)
    assert np.array_equal(s16, scores)
    oms = oracle.MixtureSet(**msd)
    for u in (0, 57, 124):
        a, b = int(offs[u]), int(offs[u + 1])
        alone, f_alone, _ = pipeline.score_utterances(fe, gmm, samples[a:b], np.array([0, b - a], np.int64), want_feats=True)
        sl = slice(int(fo[u]), int(fo[u + 1]))
        assert np.array_equal(feats[sl], f_alone) and np.array_equal(scores[sl], alone), u
        assert np.array_equal(scores[sl], oracle.gmm_batch_float(oms, feats[sl])), u
    want = oracle.gmm_batch_float(oms, oracle.mfcc(oracle.frontend_cfg(), samples[offs[57]:offs[58]])["feats"])
    sl = slice(int(fo[57]), int(fo[58]))
    rel = np.abs(scores[sl] - want) / np.abs(want)
    diag("pipeline_c3_full", max_rel=float(rel.max()), frames=int(scores.shape[0]))
    assert rel.max() < 1e-4
