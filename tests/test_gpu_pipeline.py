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


def test_launch_counter_moves():
    from rasr_b200 import capi

    before = capi.launch_count()
    flow.FrontEnd().process(synth.utterance(4000))
    assert capi.launch_count() >= before + 2
