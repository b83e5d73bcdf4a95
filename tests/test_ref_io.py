"""On-disk formats against files WRITTEN BY THE REFERENCE's own code: tests/golden/ref_io/ holds mixture
text files (Mm::Module_::writeMixtureSet), an accumulator file of a Viterbi pass (Mm::MixtureSetEstimator, the `.mix`
files of a trained system), Math::Matrix / Math::Vector files in both formats and Flow caches written by the reference's
generic-cache node; expected.npz is what the reference's own readers returned for them (tests/golden/make_golden_io.py).
Where oracle/_ref is built, files written by rasr_b200/io.py are also read back through the reference's readers."""
import os

import numpy as np
import pytest

from rasr_b200 import cache, io as rio, mm

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_io")
KEYS = ("dim", "mix_offsets", "mix_density", "mix_log_weight", "dens_mean", "dens_cov", "means", "variances")


@pytest.fixture(scope="module")
def exp():
    return np.load(os.path.join(HERE, "expected.npz"))


def same_mixture_set(got, exp, prefix):
    for k in KEYS:
        a, b = np.asarray(got[k]), exp["%s/%s" % (prefix, k)]
        assert a.shape == b.shape and np.array_equal(a, b), (prefix, k)


@pytest.mark.parametrize("name", ["ragged_p6.pms", "ragged_p9.pms", "ragged_p9.pms.gz"])
def test_mixture_text_files_written_by_the_reference(exp, name):
    same_mixture_set(rio.read_mixture_set(os.path.join(HERE, name)), exp, name)
    same_mixture_set(rio.read_mixture_file(os.path.join(HERE, name)), exp, name)


def test_accumulator_file_is_estimated_like_the_reference(exp):
    """`.mix`: accumulators -> mixture set (means, pooled variance, normalised log weights, renumbering); densities with
    fewer than minimum-observation-weight = 5 frames are dropped (two of the 24 here)"""
    got = rio.read_mixture_file(os.path.join(HERE, "viterbi.mix"))
    same_mixture_set(got, exp, "viterbi.mix")
    sizes = np.diff(got["mix_offsets"])
    assert list(sizes) == [4, 3, 4, 4, 3, 1]
    assert got["means"].shape == (19, 13) and got["variances"].shape == (1, 13)
    ms = mm.MixtureSet.read(os.path.join(HERE, "viterbi.mix"))
    assert ms.n_mixtures == 6 and ms.dim == 13


def test_accumulator_reader_rejects_other_files(tmp_path):
    p = tmp_path / "x.mix"
    p.write_bytes(b"NOTMIX\0\0" + b"\0" * 64)
    with pytest.raises(ValueError, match="MIXSET"):
        rio.read_mixture_estimator(p)


@pytest.mark.parametrize("fmt", ["bin", "xml"])
def test_matrix_and_vector_files_written_by_the_reference(exp, fmt):
    m = rio.read_matrix("%s:%s" % (fmt, os.path.join(HERE, "matrix." + fmt)))
    v = rio.read_vector("%s:%s" % (fmt, os.path.join(HERE, "vector." + fmt)))
    assert np.array_equal(m, exp["matrix"]) and np.array_equal(v, exp["vector"])


@pytest.mark.parametrize("name", ["plain.cache", "gather7_gz.cache", "dir.cache"])
def test_flow_caches_written_by_the_reference(exp, name):
    """file archive (one chunk per segment; chunks of 8 packets, gzip members), directory archive"""
    ar = cache.open_archive(os.path.join(HERE, name))
    assert sorted(ar.names()) == ["corpus/rec1/seg1", "corpus/rec1/seg1.attribs", "corpus/rec1/seg2",
                                  "corpus/rec1/seg2.attribs"]
    for seg in ("seg1", "seg2"):
        feats, times, atts = cache.read_features(ar, "corpus/rec1/" + seg)
        assert np.array_equal(feats, exp["%s/%s/feats" % (name, seg)])
        assert np.array_equal(times, exp["%s/%s/times" % (name, seg)])
        assert atts["datatype"] == "vector-f32" and atts["sample-rate"] == "1" and atts["frame-shift"] == "0.01"
    ar.close()


# ---- the other direction: what io.py / cache.py write, read by the reference (needs oracle/_ref)
def _pyref():
    from oracle import pyref

    if not pyref.available():
        pytest.skip("oracle/_ref is not built on this host")
    return pyref


def test_reference_reads_what_io_py_writes(tmp_path):
    from rasr_b200 import synth

    pyref = _pyref()
    msd = synth.ragged_mixture_set(dim=9, sizes=(2, 5, 1, 8), seed=3)
    p = str(tmp_path / "w.pms")
    rio.write_mixture_set(p, msd)
    got = pyref.read_mixture_file(p)
    for k in KEYS:
        assert np.array_equal(np.asarray(got[k]), np.asarray(msd[k])), k
    rng = np.random.default_rng(2)
    m = rng.standard_normal((4, 6)).astype(np.float32)
    for fmt in ("bin", "xml"):
        f = "%s:%s" % (fmt, tmp_path / ("m." + fmt))
        rio.write_matrix(f, m)
        assert np.array_equal(pyref.read_matrix(f), m)
        f = "%s:%s" % (fmt, tmp_path / ("v." + fmt))
        rio.write_vector(f, m[2])
        assert np.array_equal(pyref.read_vector(f), m[2])


def test_reference_reads_a_cache_cache_py_wrote(tmp_path):
    """a generic-cache node of the reference serves a segment out of an archive written by cache.py"""
    from rasr_b200 import synth

    pyref = _pyref()
    path = str(tmp_path / "mine.cache")
    feats = np.random.default_rng(4).standard_normal((37, 39)).astype(np.float32)
    times = np.stack([0.01 * np.arange(37), 0.01 * np.arange(37) + 0.025], axis=1)
    with cache.open_archive(path, "w") as ar:
        cache.write_features(ar, "corpus/rec1/seg1", feats, times, attributes={"datatype": "vector-f32"}, gather=5)
    P = pyref.chain_parameters()
    P.update({"path": path, "gather": 4294967295, "compress": "false", "id": "corpus/rec1/seg1"})
    net = pyref.FlowNetwork("mfcc_cache.flow", P)
    r = net.run(synth.utterance(400, 1), width=39)  # the samples are not looked at: the segment is cached
    net.close()
    assert np.array_equal(r["feats"], feats)
    assert np.array_equal(r["t_start"], times[:, 0]) and np.array_equal(r["t_end"], times[:, 1])
