"""The RASR-side adapters (adapters/*.cc) EXECUTED: compiled against the reference's headers, linked against the
reference's own object code (oracle/_ref/librasr_ref_native.so = its Core / Flow / Math / Signal / Mm sources) and
loaded into that host next to librasr_b200.so.  After INIT_MODULE(B200)

  * the reference's Flow::NetworkParser builds a network whose only processing node is the adapter "b200-mfcc"
    (oracle/refbuild/flows/b200_mfcc.flow) and Flow::Network::getData pulls its packets -- compared with the packets of
    the network made of the reference's own nodes (mfcc_chain_*.flow: same nodes and links as mfcc.flow +
    derivationWithRegression.flow) in the same process;
  * the reference's Mm::FeatureScorerFactory creates "b200-*" feature scorers from an Mm::MixtureSet, and the
    recognizer's buffered call protocol (src/Speech/Recognizer.cc:271-281,197-205, replayed by ref_host.cc) reads every
    emission score through Mm::FeatureScorer::ContextScorer::score -- compared with the reference's own scorers fed the
    same MixtureSet object.

Nothing here goes through the Python mirror of the interfaces: the call stack is reference code -> adapter -> C ABI ->
CUDA kernels."""
import os

import numpy as np
import pytest

from rasr_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-4


@pytest.fixture(scope="module")
def ref():
    from oracle import pyref

    need = [pyref.path(True), os.path.join(ROOT, "oracle", "_ref", "libb200_adapters.so")]
    missing = [p for p in need if not os.path.exists(p)]
    if missing and not os.path.isdir(pyref.REFERENCE):
        pytest.fail("%s missing: build them where the reference checkout is (python __graft_entry__.py)" % missing)
    pyref.load_adapters()
    return pyref


@pytest.fixture(scope="module")
def oms():
    from oracle import pyoracle  # only for the C layout of the mixture set handed to the reference

    return pyoracle


def dim_err(got, want):
    scale = np.sqrt(np.mean(want.astype(np.float64) ** 2, axis=0))
    return float((np.abs(got.astype(np.float64) - want) / scale).max())


def adapter_parameters(ref, dc=False, **kw):
    p = ref.chain_parameters(dc=True, **kw)
    p.update({"derivatives": "true", "dc-detection": "true" if dc else "false"})
    return p


def test_mfcc_node_in_a_network_built_by_the_reference(ref, diag):
    """C1: the 10 s utterance.  Same number of packets, identical f64 time stamps, features within 1e-4; the attributes
    the adapter publishes equal what the reference's chain leaves."""
    x = synth.utterance(160000)
    theirs = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters(), native=True)
    ours = ref.FlowNetwork("b200_mfcc.flow", adapter_parameters(ref), native=True)
    a, b = theirs.run(x), ours.run(x)
    assert a["feats"].shape == b["feats"].shape == (999, 39)
    assert np.array_equal(a["t_start"], b["t_start"]) and np.array_equal(a["t_end"], b["t_end"])
    err = dim_err(b["feats"], a["feats"])
    diag("adapter_mfcc_node_c1", err=err)
    assert err < RTOL
    for name in ("sample-rate", "frame-shift", "datatype"):
        assert ours.attribute("features", name) == theirs.attribute("features", name), name
    # a second and third segment through the same node instance: state reset at EOS, new start time, other packet size
    for seed, n, t0 in ((5, 12345, 3.25), (6, 401, 100.0)):
        y = synth.utterance(n, seed=seed)
        a, b = theirs.run(y, start_time=t0), ours.run(y, start_time=t0)
        assert a["feats"].shape == b["feats"].shape
        assert np.array_equal(a["t_start"], b["t_start"]) and np.array_equal(a["t_end"], b["t_end"])
        assert dim_err(b["feats"], a["feats"]) < RTOL


@pytest.mark.parametrize("kw", [dict(block_size=160), dict(alpha="0.97", nr_cepstrum_coefficients=16),
                                dict(window_type="hanning", shift=".005", length=".02")])
def test_mfcc_node_parameters_reach_the_engine(ref, kw):
    x = synth.utterance(20000, seed=21)
    a = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters(**kw), native=True).run(x)
    b = ref.FlowNetwork("b200_mfcc.flow", adapter_parameters(ref, **kw), native=True).run(x)
    assert a["feats"].shape == b["feats"].shape
    assert np.array_equal(a["t_start"], b["t_start"]) and np.array_equal(a["t_end"], b["t_end"])
    assert dim_err(b["feats"], a["feats"]) < RTOL


def test_mfcc_node_with_dc_detection(ref):
    x = synth.utterance(40000, seed=9)
    x[5000:9000] = x[4999]
    x[15000:15300] = 7.0
    a = ref.FlowNetwork("mfcc_chain_dc.flow", ref.chain_parameters(dc=True), native=True).run(x)
    b = ref.FlowNetwork("b200_mfcc.flow", adapter_parameters(ref, dc=True), native=True).run(x)
    assert a["feats"].shape == b["feats"].shape and a["feats"].shape[0] < 249
    assert np.array_equal(a["t_start"], b["t_start"]) and np.array_equal(a["t_end"], b["t_end"])
    assert dim_err(b["feats"], a["feats"]) < RTOL


PAIRS = [("batch-diagonal-maximum-float", "b200-batch-float"), ("batch-diagonal-maximum-int", "b200-batch-int"),
         ("SIMD-diagonal-maximum", "b200-SIMD-diagonal-maximum"),
         ("preselection-batch-float", "b200-preselection-batch-float"),
         ("preselection-batch-int", "b200-preselection-batch-int")]


@pytest.mark.parametrize("theirs,ours", PAIRS, ids=[p[1] for p in PAIRS])
def test_scorers_from_the_references_factory_bit_identical(ref, oms, theirs, ours):
    """C2's model, 600 frames: the reference's scorer and the adapter, both created by Mm::Module's factory from the
    same MixtureSet, both read through ContextScorer::score under the recognizer's protocol: identical bits"""
    ms = oms.MixtureSet(**synth.mixture_set())
    f = synth.features(600, 39, seed=31)
    a = ref.FeatureScorer(ms, theirs, native=True).score(f)
    b = ref.FeatureScorer(ms, ours, native=True).score(f)
    assert a.shape == b.shape == (600, 256)
    assert np.array_equal(a, b), "%d of %d scores differ" % ((a != b).sum(), a.size)


def test_strict_variant_and_preselection_resources(ref, oms):
    """fma-contraction=false (the adapter's resource) against the reference's strict build; density-clustering.* resources
    travel through the reference's own configuration to both scorers"""
    ms = oms.MixtureSet(**synth.mixture_set(n_mixtures=32))
    f = synth.features(300, 39, seed=32)
    a = ref.FeatureScorer(ms, "batch-diagonal-maximum-float", native=False).score(f)
    b = ref.FeatureScorer(ms, "b200-batch-float", {"fma-contraction": "false"}, native=True).score(f)
    assert np.array_equal(a, b)
    cfg = {"density-clustering.clusters": 64, "density-clustering.select-clusters": 8,
           "density-clustering.iterations": 3, "density-clustering.backoff-score": 777.0}
    for theirs, ours in PAIRS[3:]:
        a = ref.FeatureScorer(ms, theirs, cfg, native=True).score(f)
        b = ref.FeatureScorer(ms, ours, cfg, native=True).score(f)
        assert np.array_equal(a, b), ours


def test_diagonal_scorers_from_the_references_factory(ref, oms, diag):
    ms = oms.MixtureSet(**synth.ragged_mixture_set(dim=24, n_covariances=3, seed=11))
    f = synth.features(200, 24, seed=33)
    a = ref.FeatureScorer(ms, "diagonal-maximum", native=False).score(f)
    b = ref.FeatureScorer(ms, "b200-diagonal-maximum", {"fma-contraction": "false"}, native=True).score(f)
    assert np.array_equal(a, b)
    s = ref.FeatureScorer(ms, "diagonal-sum", native=False).score(f)
    t = ref.FeatureScorer(ms, "b200-diagonal-sum", {"fma-contraction": "false"}, native=True).score(f)
    rel = float((np.abs(s - t) / np.abs(s)).max())
    diag("adapter_diag_sum", rel=rel)
    assert rel < 1e-6
    c = ref.FeatureScorer(ms, "b200-diagonal-maximum", native=True).score(f)
    assert float((np.abs(c - a) / np.abs(a)).max()) < 1e-6  # contracted variant: an ulp or two
    # the quantised scorer with one quantised feature vector per covariance: integer arithmetic, identical bits
    q = ref.FeatureScorer(ms, "SIMD-diagonal-maximum", native=True).score(f)
    r = ref.FeatureScorer(ms, "b200-SIMD-diagonal-maximum", native=True).score(f)
    assert np.array_equal(q, r)


def test_tensor_mode_through_the_adapter(ref, oms):
    ms = oms.MixtureSet(**synth.mixture_set(n_mixtures=64))
    f = synth.features(512, 39, seed=34)
    a = ref.FeatureScorer(ms, "batch-diagonal-maximum-float", native=True).score(f)
    b = ref.FeatureScorer(ms, "b200-batch-tensor", native=True).score(f)
    assert float((np.abs(a - b) / np.abs(a)).max()) < RTOL


@pytest.mark.parametrize("buffer_size", [1, 5, 1000])
def test_adapter_buffer_sizes(ref, oms, buffer_size):
    """bufferFilled() / flush() protocol of the adapter for small, odd and larger-than-segment buffers"""
    ms = oms.MixtureSet(**synth.mixture_set(n_mixtures=16))
    f = synth.features(77, 39, seed=35)
    a = ref.FeatureScorer(ms, "batch-diagonal-maximum-float", native=True).score(f)
    b = ref.FeatureScorer(ms, "b200-batch-float", {"buffer-size": buffer_size}, native=True).score(f)
    assert np.array_equal(a, b)


# ---------------------------------------------------------------------------------------------------------------
# the other adapters: Nn scorer and forward node configured with the REFERENCE's keys, post-processing node, and the
# fused audio -> scores node -- each against the reference's own component in the same process

from tests.helpers_nn import NN_FLOW, nn_files  # noqa: E402  (same configuration / file helpers as the CPU suite)
from rasr_b200 import io  # noqa: E402


@pytest.mark.parametrize("hidden", ["sigmoid", "relu"])
def test_nn_scorer_reads_the_references_configuration(ref, oms, tmp_path, diag, hidden):
    """b200-nn-batch-feature-scorer given exactly the configuration of nn-batch-feature-scorer (neural-network.links,
    layer-type, dimension-*, parameters-old, prior-file, priori-scale): f32 path within 1e-5, bf16 path within the stated
    bf16 tolerance (2e-3 of scale; bf16 cannot meet 1e-4 against f32 sgemm)"""
    net = synth.network(dims=(45, 96, 64, 40), hidden=hidden, seed=3)
    cfg = nn_files(tmp_path, net, hidden)
    io.write_vector("xml:" + str(tmp_path / "prior.xml"), net["log_prior"])
    cfg.update({"prior-file": "xml:" + str(tmp_path / "prior.xml"), "priori-scale": 0.7})
    ms = oms.MixtureSet(**synth.mixture_set(dim=45, n_mixtures=40, densities_per_mixture=1))
    x = synth.features(333, 45, seed=9, scale=1.0)
    want = ref.FeatureScorer(ms, "nn-batch-feature-scorer", cfg, native=True).score(x)
    f32 = ref.FeatureScorer(ms, "b200-nn-batch-feature-scorer", dict(cfg, bf16="false"), native=True).score(x)
    bf16 = ref.FeatureScorer(ms, "b200-nn-batch-feature-scorer", cfg, native=True).score(x)
    scale = np.abs(want).max()
    e32, e16 = float(np.abs(f32 - want).max() / scale), float(np.abs(bf16 - want).max() / scale)
    diag("adapter_nn_scorer_" + hidden, f32=e32, bf16=e16)
    assert e32 < 1e-5 and e16 < 2e-3


def test_nn_forward_node_against_the_references_node(ref, tmp_path, diag):
    net = synth.network(dims=(45, 96, 40), hidden="sigmoid", seed=6)
    cfg = nn_files(tmp_path, net, "sigmoid")
    io.write_vector("xml:" + str(tmp_path / "prior.xml"), net["log_prior"])
    cfg.update({"prior-file": "xml:" + str(tmp_path / "prior.xml"), "priori-scale": 0.5})
    out = {}
    for sel, filt in (("nnref", "neural-network-forward"), ("nnb200", "b200-neural-network-forward")):
        for k, v in cfg.items():
            ref.config_set("*.%s.nn.%s" % (sel, k), v, native=True)
        ref.config_set("*.%s.nn.bf16" % sel, "false", native=True)
        (tmp_path / (sel + ".flow")).write_text(NN_FLOW % filt)
        x = synth.features(200, 45, seed=12, scale=1.0)
        out[sel] = ref.FlowNetwork(str(tmp_path / (sel + ".flow")), {"block-size": 45}, native=True,
                                   selection=sel).run(x.reshape(-1), sample_rate=4500.0)
    a, b = out["nnref"], out["nnb200"]
    assert a["feats"].shape == b["feats"].shape == (200, 40)
    err = float(np.abs(a["feats"] - b["feats"]).max())
    diag("adapter_nn_forward_node", err=err)
    assert err < 1e-6  # posteriors (softmax, prior removed from the bias), f32 path
    # the adapter carries the input packet's time stamps (the reference emits [1, 1]: see tests/test_ref_parity.py)
    assert np.allclose(b["t_start"], np.arange(200) * 0.01) and np.allclose(b["t_end"], (np.arange(200) + 1) * 0.01)


POSTPROC_FLOW = """<?xml version="1.0" encoding="ISO-8859-1"?>
<network name="network">
  <out name="projected"/>
  <param name="block-size"/>
  <param name="norm-type"/> <param name="norm-length"/> <param name="norm-right"/>
  <param name="splice-length"/> <param name="splice-right"/> <param name="matrix-file"/>
  <node name="source" filter="ref-sample-source" block-size="$(block-size)"/>
  <node name="post" filter="b200-feature-postprocessing" normalization-type="$(norm-type)"
        normalization-length="$(norm-length)" normalization-right="$(norm-right)" window-max-size="$(splice-length)"
        window-right="$(splice-right)" matrix-file="$(matrix-file)" fma-contraction="false"/>
  <link from="source" to="post"/>
  <link from="post" to="network:projected"/>
</network>
"""


@pytest.mark.parametrize("kind,length,right", [("mean-and-variance", "infinite", "infinite"), ("mean", 51, 25)])
def test_postprocessing_node_against_the_references_nodes(ref, tmp_path, kind, length, right):
    """one node instead of signal-normalization -> sequence concatenation -> matrix multiplication: identical bits and
    identical time stamps (strict arithmetic on both sides)"""
    f = synth.features(300, 13, seed=5)
    M = np.random.default_rng(3).standard_normal((20, 65)).astype(np.float32)
    path = "bin:%s" % (tmp_path / "lda.bin")
    io.write_matrix(path, M)
    P = {"block-size": 13, "norm-type": kind, "norm-length": length, "norm-right": right, "splice-length": 5,
         "splice-right": 2, "matrix-file": path}
    (tmp_path / "post.flow").write_text(POSTPROC_FLOW)
    a = ref.FlowNetwork("postproc_chain.flow", P, native=True).run(f.reshape(-1), port="projected", sample_rate=1300.0)
    b = ref.FlowNetwork(str(tmp_path / "post.flow"), P, native=True).run(f.reshape(-1), port="projected", sample_rate=1300.0)
    assert a["feats"].shape == b["feats"].shape == (300, 20)
    # the native reference build contracts the matrix product; the strict comparison is in the golden test
    assert np.abs(a["feats"] - b["feats"]).max() / np.abs(a["feats"]).max() < 1e-6
    assert np.array_equal(a["t_start"], b["t_start"]) and np.array_equal(a["t_end"], b["t_end"])


AUDIO_FLOW = """<?xml version="1.0" encoding="ISO-8859-1"?>
<network name="network">
  <out name="scores"/>
  <param name="block-size"/>
  <node name="source" filter="ref-sample-source" block-size="$(block-size)"/>
  <node name="scorer" filter="b200-audio-feature-scorer"/>
  <link from="source" to="scorer"/>
  <link from="scorer" to="network:scores"/>
</network>
"""


def test_audio_to_scores_node(ref, oms, tmp_path, diag):
    """b200-audio-feature-scorer: samples in, +log scores out, features never leave the device.  Against the reference's
    own chain: its MFCC network, then its batch scorer on those features, negated as Speech::FeatureScorerNode does.
    The mixture set travels as a text file (.pms) written by rasr_b200.io and read by the reference's MixtureSetReader."""
    msd = synth.mixture_set(n_mixtures=48)
    io.write_mixture_set(str(tmp_path / "model.pms"), msd)
    ref.config_set("*.audioflow.scorer.mixture-set.file", str(tmp_path / "model.pms"), native=True)
    ref.config_set("*.audioflow.scorer.feature-scorer.feature-scorer-type", "batch-diagonal-maximum-float", native=True)
    (tmp_path / "audio.flow").write_text(AUDIO_FLOW)
    x = synth.utterance(48240, seed=41)
    got = ref.FlowNetwork(str(tmp_path / "audio.flow"), {"block-size": 4096}, native=True, selection="audioflow").run(x, port="scores")
    feats = ref.FlowNetwork("mfcc_chain_plain.flow", ref.chain_parameters(), native=True).run(x)
    want = -ref.FeatureScorer(oms.MixtureSet(**msd), "batch-diagonal-maximum-float", native=True).score(feats["feats"])
    assert got["feats"].shape == want.shape == (300, 48)
    assert np.array_equal(got["t_start"], feats["t_start"]) and np.array_equal(got["t_end"], feats["t_end"])
    rel = float((np.abs(got["feats"] - want) / np.abs(want)).max())
    diag("adapter_audio_scorer_node", rel=rel)
    assert rel < RTOL


def test_linear_search_adapter_next_to_the_reference_search(diag):
    """adapters/B200LinearSearch.cc (Search::SearchAlgorithm over rb_search_*) and the reference's own
    Search::LinearSearch, both set up by the reference's lexicon parser / acoustic model parts / LM scaling from one
    lexicon file and fed the same scorer objects: every traceback item bit-identical, continuous and single-word
    recognition.  In a process of its own (tests/search_adapter_host.py): the search host is a second copy of the
    reference's object code and must not share a symbol scope with the one the tests above loaded."""
    import json
    import subprocess
    import sys

    from oracle import pyref
    need = [pyref.search_path(), os.path.join(ROOT, "oracle", "_ref", "libb200_search_adapter.so")]
    missing = [p for p in need if not os.path.exists(p)]
    if missing and not os.path.isdir(pyref.REFERENCE):
        pytest.fail("%s missing: build them where the reference checkout is (python __graft_entry__.py)" % missing)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "search_adapter_host.py")], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["ok"], out
    assert set(out["items"]) == {"continuous", "continuous_scaled", "single_word", "single_word_noise", "single_word_ties",
                                 "whole_path"}
    # whole_path: feature vectors -> b200-batch-float (adapter) -> B200::LinearSearch (adapter, dense score rows) against
    # the reference's scorer feeding the reference's search, through the recognizer's own loop
    assert out["items"]["whole_path"] > 0
    diag("search_adapter_in_reference_host", **{k: int(v) for k, v in out["items"].items()})
