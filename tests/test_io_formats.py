"""Mixture text file ("PMS") reader / writer against the format description of the reference
(doc/file_formats/mixture_file.rst:24-40) and the reader's semantics (src/Mm/MixtureSet.cc:168-208).  CPU only."""
import gzip

import numpy as np
import pytest

from rasr_b200 import io as rio
from rasr_b200 import synth

HAND_WRITTEN = """#Version: 1.0
#CovarianceType: DiagonalCovariance
2 2 3 3 1
2 0 0.25 1 0.75
1 2 1
0 0 1 0
2 0
2 1.5 -2.25
2 0.1 0.3 2 0 1e-1
  2 4 0.5 3 2
"""


def test_reads_hand_written_version_1_file(tmp_path):
    p = tmp_path / "m.pms"
    p.write_text(HAND_WRITTEN)
    ms = rio.read_mixture_set(p)
    assert ms["dim"] == 2 and list(ms["mix_offsets"]) == [0, 2, 3]
    assert list(ms["mix_density"]) == [0, 1, 2]
    np.testing.assert_array_equal(ms["mix_log_weight"], np.log([0.25, 0.75, 1.0]))  # version < 2.0: linear weights
    assert list(ms["dens_mean"]) == [0, 1, 2] and list(ms["dens_cov"]) == [0, 0, 0]
    # tokens may break across lines anywhere: mean 2 is "2 0.1 0.3", mean ... (whitespace-separated stream)
    np.testing.assert_array_equal(ms["means"], np.array([[1.5, -2.25], [0.1, 0.3], [0, 0.1]], np.float32))
    np.testing.assert_array_equal(ms["variances"], np.array([[2.0, 6.0]], np.float32))  # v * w


@pytest.mark.parametrize("gz", [False, True])
def test_round_trip_is_bit_exact(tmp_path, gz):
    msd = synth.ragged_mixture_set(dim=39, sizes=(1, 3, 16, 0, 7), seed=4)
    p = tmp_path / ("m.pms.gz" if gz else "m.pms")
    rio.write_mixture_set(p, msd)
    if gz:
        assert gzip.open(p).readline().startswith(b"#Version: 2.0")
    back = rio.read_mixture_set(p)
    for k in ("mix_offsets", "mix_density", "mix_log_weight", "dens_mean", "dens_cov", "means", "variances"):
        assert np.array_equal(np.asarray(back[k]), np.asarray(msd[k])), k
    assert back["dim"] == msd["dim"]


def test_f32_fields_are_parsed_with_a_single_rounding(tmp_path):
    # 1 + 2^-24 + 2^-54: the f64 nearest to it is the tie 1 + 2^-24, which then rounds to 1.0 (ties-to-even); the
    # reference's `istream >> float` rounds once and gives the next f32 above 1
    s = "1.0000000596046448286"
    p = tmp_path / "m.pms"
    p.write_text("#Version: 2.0\n#CovarianceType: DiagonalCovariance\n1 1 1 1 1\n1 0 0\n0 0\n1 %s\n1 1 1\n" % s)
    assert rio.read_mixture_set(p)["means"][0, 0] == np.nextafter(np.float32(1), np.float32(2))


def test_rejects_foreign_files(tmp_path):
    p = tmp_path / "x.pms"
    p.write_text("#Version: 3.0\n#CovarianceType: DiagonalCovariance\n1 0 0 0 0\n")
    with pytest.raises(ValueError):
        rio.read_mixture_set(p)
    p.write_text("#Version: 2.0\n#CovarianceType: FullCovariance\n1 0 0 0 0\n")
    with pytest.raises(ValueError):
        rio.read_mixture_set(p)
    p.write_text("#Version: 2.0\n#CovarianceType: DiagonalCovariance\n2 1 1 1 1\n1 0 0\n0 0\n2 1\n")
    with pytest.raises(ValueError):
        rio.read_mixture_set(p)


def test_loaded_model_scores_like_the_in_memory_one(oracle, tmp_path):
    msd = synth.mixture_set(dim=13, n_mixtures=6, densities_per_mixture=4, seed=8)
    p = tmp_path / "m.pms.gz"
    rio.write_mixture_set(p, msd)
    f = synth.features(20, 13, seed=1)
    a = oracle.gmm_batch_float(oracle.MixtureSet(**msd), f)
    b = oracle.gmm_batch_float(oracle.MixtureSet(**rio.read_mixture_set(p)), f)
    assert np.array_equal(a, b)


# ---- Math::Matrix / Math::Vector files (Nn layer parameters, priors) ------------------------------------------------

def test_reads_hand_written_matrix_and_vector_xml(tmp_path):
    """the layout of src/Core/MatrixParser.hh:76-112 / VectorParser.hh:78-103: attributes, row after row, line breaks
    without meaning, optional size"""
    p = tmp_path / "m.xml"
    p.write_text('<?xml version="1.0" encoding="ISO-8859-1"?>\n<matrix-f32 nRows="2" nColumns="3">\n'
                 "1 2.5\n-3e-1 4 5\n6.25e+00\n</matrix-f32>\n")
    m = rio.read_matrix(str(p))
    assert m.dtype == np.float32 and np.array_equal(m, np.array([[1, 2.5, -0.3], [4, 5, 6.25]], np.float32))
    assert np.array_equal(rio.read_matrix("xml:" + str(p)), m)
    v = tmp_path / "v.xml"
    v.write_text('<vector-f32 size="3"> 0.1 0.2 0.3 </vector-f32>')
    assert np.array_equal(rio.read_vector(str(v)), np.array([0.1, 0.2, 0.3], np.float32))
    v.write_text("<vector-f64> 0.1 0.2 </vector-f64>")
    assert np.array_equal(rio.read_vector(str(v), np.float64), np.array([0.1, 0.2]))
    v.write_text('<vector-f32 size="4"> 0.1 0.2 0.3 </vector-f32>')
    with pytest.raises(ValueError):
        rio.read_vector(str(v))
    p.write_text('<matrix-f32 nRows="2" nColumns="3"> 1 2 3 4 5 </matrix-f32>')
    with pytest.raises(ValueError):
        rio.read_matrix(str(p))
    p.write_text('<matrix-f32 nRows="2"> 1 2 </matrix-f32>')
    with pytest.raises(ValueError):
        rio.read_matrix(str(p))


def test_reads_hand_packed_binary_matrix(tmp_path):
    """u32 nRows, u32 nColumns, u32 #row vectors, per row u32 size + little-endian data (src/Math/Matrix.hh:560-575)"""
    import struct
    p = tmp_path / "m.bin"
    p.write_bytes(struct.pack("<III", 2, 2, 2) + struct.pack("<Iff", 2, 1.5, -2.0) + struct.pack("<Iff", 2, 0.25, 8.0))
    assert np.array_equal(rio.read_matrix("bin:" + str(p)), np.array([[1.5, -2.0], [0.25, 8.0]], np.float32))
    p.write_bytes(struct.pack("<III", 2, 2, 2) + struct.pack("<Iff", 2, 1.5, -2.0) + struct.pack("<Ifff", 3, 0, 0, 0))
    with pytest.raises(ValueError):
        rio.read_matrix("bin:" + str(p))


@pytest.mark.parametrize("fmt", ["bin:", "xml:", ""])
def test_matrix_and_vector_round_trip_bit_exact(tmp_path, fmt):
    rng = np.random.default_rng(5)
    m = (rng.standard_normal((7, 5)) * 10.0 ** rng.integers(-20, 20, (7, 5))).astype(np.float32)
    v = rng.standard_normal(11).astype(np.float32)
    rio.write_matrix(fmt + str(tmp_path / "m"), m)
    rio.write_vector(fmt + str(tmp_path / "v"), v)
    assert np.array_equal(rio.read_matrix(fmt + str(tmp_path / "m")), m)
    assert np.array_equal(rio.read_vector(fmt + str(tmp_path / "v")), v)
    d = rng.standard_normal((3, 4))
    rio.write_matrix(fmt + str(tmp_path / "d"), d)
    assert np.array_equal(rio.read_matrix(fmt + str(tmp_path / "d"), np.float64), d)


def test_layer_parameter_files_follow_set_parameters(tmp_path):
    """row = output unit, column 0 = bias, columns 1.. = weights (src/Nn/LinearLayer.cc:383-424)"""
    from rasr_b200 import nn
    param = np.arange(12, dtype=np.float32).reshape(3, 4)
    rio.write_matrix("bin:" + str(tmp_path / "l0"), param)
    w, b = nn.parameters_from_matrix(rio.read_matrix("bin:" + str(tmp_path / "l0")))
    assert np.array_equal(b, [0, 4, 8]) and np.array_equal(w, [[1, 2, 3], [5, 6, 7], [9, 10, 11]])


# ---- feature caches: Core::FileArchive + Flow cache streams ---------------------------------------------------------

def _hand_packed_archive(path, with_table):
    """bytes laid out by hand from the format comment of src/Core/FileArchive.cc:26-80"""
    import struct
    name = b"corpus/rec1/seg1"
    vec = lambda v, t0, t1: struct.pack("<I", len(v)) + struct.pack("<%df" % len(v), *v) + struct.pack("<dd", t0, t1)
    stream = (struct.pack("<I", 10) + b"vector-f32" + struct.pack("<I", 2) +
              vec([1.0, 2.0, 3.0], 0.0, 0.025) + vec([4.0, 5.0, 6.0], 0.01, 0.035))
    body = HEAD = b"SP_ARC1\0" + (b"\1" if with_table else b"\0")
    body += struct.pack("<II", 0xAA55AA55, len(name)) + name
    pos = len(body)
    body += struct.pack("<III", len(stream), 0, 0) + stream + struct.pack("<I", 0x55AA55AA)
    if with_table:
        table = len(body)
        body += struct.pack("<I", 1) + struct.pack("<I", len(name)) + name + struct.pack("<QII", pos, len(stream), 0)
        holes = len(body)
        body += struct.pack("<I", 0) + struct.pack("<QQ", holes, table)
    path.write_bytes(body)


@pytest.mark.parametrize("with_table", [True, False])
def test_reads_hand_packed_feature_cache(tmp_path, with_table):
    from rasr_b200 import cache
    _hand_packed_archive(tmp_path / "f.cache", with_table)
    with cache.FileArchive(tmp_path / "f.cache") as a:
        assert a.names() == ["corpus/rec1/seg1"]
        feats, times, atts = cache.read_features(a, "corpus/rec1/seg1")
    assert feats.dtype == np.float32 and np.array_equal(feats, [[1, 2, 3], [4, 5, 6]])
    assert np.array_equal(times, [[0.0, 0.025], [0.01, 0.035]]) and atts == {}


@pytest.mark.parametrize("compress,gather", [(False, 0xFFFFFFFF), (True, 0xFFFFFFFF), (True, 3)])
def test_feature_cache_round_trip(tmp_path, compress, gather):
    import gzip as gz
    from rasr_b200 import cache
    rng = np.random.default_rng(8)
    segs = {}
    with cache.FileArchive(tmp_path / "d" / "f.cache", "w") as a:
        for i, T in enumerate([17, 1, 40]):
            f = rng.standard_normal((T, 39)).astype(np.float32)
            t = np.stack([np.arange(T) * 0.01, np.arange(T) * 0.01 + 0.025], 1)
            segs["c/r/s%d" % i] = (f, t)
            cache.write_features(a, "c/r/s%d" % i, f, t, {"sample-rate": "100", "datatype": "vector-f32"},
                                 gather=gather, compress=compress)
        with pytest.raises(cache.ArchiveError):
            a.write("c/r/s0", b"again")  # allow-overwrite is off
    for strip_table in (False, True):
        if strip_table:  # a writer that died before the table was written: flag 0 -> the reader scans
            raw = bytearray((tmp_path / "d" / "f.cache").read_bytes())
            raw[8] = 0
            (tmp_path / "d" / "f.cache").write_bytes(bytes(raw))
        with cache.FileArchive(tmp_path / "d" / "f.cache") as a:
            assert sorted(a.names()) == sorted(list(segs) + [s + ".attribs" for s in segs])
            for s, (f, t) in segs.items():
                got, times, atts = cache.read_features(a, s)
                assert np.array_equal(got, f) and np.array_equal(times, t)
                assert atts == {"sample-rate": "100", "datatype": "vector-f32"}
            if compress:  # stored entries are complete gzip members
                pos, size, comp = a.files["c/r/s2"]
                raw = (tmp_path / "d" / "f.cache").read_bytes()[pos + 12:pos + 12 + comp]
                assert comp > 0 and len(gz.decompress(raw)) == size


def test_overwrite_leaves_a_hole_that_readers_skip(tmp_path):
    from rasr_b200 import cache
    with cache.FileArchive(tmp_path / "a", "w", allow_overwrite=True) as a:
        a.write("x", b"1111")
        a.write("y", b"22")
        a.write("x", b"333333")   # x is not the last entry: its old place becomes a hole
        a.write("x", b"4")        # now it is the last one: the archive shrinks
    for flag in (1, 0):
        raw = bytearray((tmp_path / "a").read_bytes())
        raw[8] = flag
        (tmp_path / "a").write_bytes(bytes(raw))
        with cache.FileArchive(tmp_path / "a") as a:
            assert sorted(a.names()) == ["x", "y"] and a.read("x") == b"4" and a.read("y") == b"22"
            assert len(a.holes) == 1
    with pytest.raises(cache.ArchiveError):
        cache.FileArchive(tmp_path / "missing")
    (tmp_path / "junk").write_bytes(b"not an archive")
    with pytest.raises(cache.ArchiveError):
        cache.FileArchive(tmp_path / "junk")


def test_directory_and_bundle_archives(tmp_path):
    """the other two archive kinds (src/Core/DirectoryArchive.cc, BundleArchive.cc) and the type detection of
    Archive::create: a feature cache split over a file archive and a directory archive, read through a bundle"""
    import gzip as gz
    from rasr_b200 import cache
    rng = np.random.default_rng(3)
    data = {}
    with cache.open_archive(tmp_path / "part1.cache", "w") as a:
        assert isinstance(a, cache.FileArchive)
        for s in ("c/r/a", "c/r/b"):
            data[s] = rng.standard_normal((9, 5)).astype(np.float32)
            cache.write_features(a, s, data[s], np.zeros((9, 2)), {"sample-rate": "100"})
    with cache.open_archive(str(tmp_path / "part2") + "/", "w") as a:
        assert isinstance(a, cache.DirectoryArchive)
        for s, comp in (("c/r/c", False), ("c/q/d", True)):
            data[s] = rng.standard_normal((4, 5)).astype(np.float32)
            cache.write_features(a, s, data[s], np.ones((4, 2)), compress=comp)
    assert (tmp_path / "part2" / "c" / "r" / "c").is_file()          # entries are plain files ...
    raw = (tmp_path / "part2" / "c" / "q" / "d").read_bytes()
    assert raw[:2] == b"\x1f\x8b" and len(gz.decompress(raw)) == int.from_bytes(raw[-4:], "little")  # ... or gzip files
    (tmp_path / "all.bundle").write_text("%s\n%s\n" % (tmp_path / "part1.cache", tmp_path / "part2"))
    for _ in range(2):  # second pass goes through the cached index
        with cache.open_archive(tmp_path / "all.bundle") as b:
            assert isinstance(b, cache.BundleArchive) and set(data) <= set(b.names())
            for s, f in data.items():
                got, _, _ = cache.read_features(b, s)
                assert np.array_equal(got, f)
            with pytest.raises(cache.ArchiveError):
                b.write("x", b"y")
    idx = gz.open(str(tmp_path / "all.bundle") + ".idx.gz", "rt").read().split()
    assert idx[0] == "2" and idx[1] == str(tmp_path / "part1.cache") and "c/q/d" in idx
    (tmp_path / "junk.bin").write_bytes(b"something else")
    with pytest.raises(cache.ArchiveError):
        cache.open_archive(tmp_path / "junk.bin")


def test_file_archive_against_a_dict_model(tmp_path):
    """random sequences of writes and overwrites (some compressed, some entries later removed by overwriting): after
    reopening -- through the info table and through the entry scan -- the archive holds exactly what a dict holds"""
    from hypothesis import given, settings, strategies as st
    from rasr_b200 import cache
    names = st.sampled_from(["a", "b", "seg/1", "seg/2", "corpus/rec/long-name-0001", "x.attribs"])
    ops = st.lists(st.tuples(names, st.binary(min_size=0, max_size=300), st.booleans()), min_size=1, max_size=25)
    counter = [0]

    @settings(max_examples=40, deadline=None)
    @given(ops)
    def run(sequence):
        counter[0] += 1
        path = tmp_path / ("m%d.cache" % counter[0])
        model = {}
        with cache.FileArchive(path, "w", allow_overwrite=True) as a:
            for name, data, comp in sequence:
                a.write(name, data, compress=comp)
                model[name] = data
        for flag in (1, 0):
            raw = bytearray(path.read_bytes())
            raw[8] = flag
            path.write_bytes(bytes(raw))
            with cache.FileArchive(path) as a:
                assert sorted(a.names()) == sorted(model)
                for name, data in model.items():
                    assert a.read(name) == data
        # appending to an existing archive keeps what is there
        with cache.FileArchive(path, "w", allow_overwrite=True) as a:
            a.write("late", b"entry")
        with cache.FileArchive(path) as a:
            assert a.read("late") == b"entry" and all(a.read(n) == d for n, d in model.items())

    run()


def test_matrix_formats_round_trip_any_finite_value(tmp_path):
    """property: every finite f32 / f64 (denormals, signed zeros, extremes) survives both formats bit for bit -- the
    xml writer prints 20 significant digits like formats().write(filename, parameters, 20)"""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp
    counter = [0]

    @settings(max_examples=30, deadline=None)
    @given(hnp.arrays(np.float32, hnp.array_shapes(min_dims=2, max_dims=2, max_side=6),
                      elements=st.floats(width=32, allow_nan=False, allow_infinity=False)),
           hnp.arrays(np.float64, st.integers(0, 9), elements=st.floats(allow_nan=False, allow_infinity=False)))
    def run(m, v):
        counter[0] += 1
        for fmt in ("bin:", "xml:"):
            pm, pv = fmt + str(tmp_path / ("m%d" % counter[0])), fmt + str(tmp_path / ("v%d" % counter[0]))
            rio.write_matrix(pm, m)
            rio.write_vector(pv, v)
            gm, gv = rio.read_matrix(pm), rio.read_vector(pv, np.float64)
            assert gm.shape == m.shape and gm.tobytes() == m.tobytes()
            assert gv.shape == v.shape and gv.tobytes() == v.tobytes()

    run()
