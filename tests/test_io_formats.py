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
