"""The C-ABI library loads without a GPU and exports every symbol include/rasr_b200.h declares.
No compute calls here (CPU-only suite)."""
import ctypes as C
import os
import re

import pytest

from rasr_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rasr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.LIB_PATH):
        capi.build()
    return capi.lib()


def test_header_and_binding_agree():
    assert declared_functions() == sorted(capi.SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in declared_functions() if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_counters(lib):
    assert capi.version().startswith("rasr_b200")
    assert capi.launch_count() >= 0
    assert capi.device_count() >= 0


def test_fails_loudly_without_a_device(lib):
    """The product path must not fall back to anything when there is no sm_100 device."""
    if capi.device_count() > 0:
        pytest.skip("a device is present")
    from rasr_b200 import flow, mm, synth

    with pytest.raises(capi.RasrB200Error) as e:
        flow.FrontEnd()
    assert e.value.status == -2
    with pytest.raises(capi.RasrB200Error) as e:
        mm.GmmScorer(synth.mixture_set(dim=8, n_mixtures=4, densities_per_mixture=2))
    assert e.value.status == -2


def test_configuration_errors_are_reported_before_device_errors(lib):
    from rasr_b200 import flow

    with pytest.raises(capi.RasrB200Error) as e:
        flow.FrontEnd(window_length=0.1, fft_max_input=0.01)  # window longer than the FFT
    assert e.value.status == -1
    with pytest.raises(capi.RasrB200Error) as e:
        flow.FrontEnd(sample_rate=0.0)
    assert e.value.status == -1


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under rasr_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "rasr_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                hit = re.search(r"(from|import)\s+oracle|liboracle|pyoracle|pyref|librasr_ref|oracle\.h|\borc_[a-z]|\bref_(mm|flow)_", text)
                assert hit is None, (os.path.join(dirpath, fn), hit.group(0))


def test_scorer_mode_constants_match_the_header():
    """rb_gmm_mode in include/rasr_b200.h against the Python mirror (capi.GMM_*) and the reference's scorer names"""
    from rasr_b200 import mm
    text = open(os.path.join(ROOT, "include", "rasr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    enum = dict((k, int(v)) for k, v in re.findall(r"\b(RB_GMM_[A-Z_]+)\s*=\s*(\d+)", text))
    assert enum == {"RB_" + k: getattr(capi, k) for k in dir(capi) if k.startswith("GMM_")}
    assert sorted(mm.GmmScorer.MODES.values()) == sorted(enum.values())
    assert mm.GmmScorer.MODES["preselection-batch-int"] == enum["RB_GMM_BATCH_PRESELECT_INT"]
    assert mm.GmmScorer.MODES["SIMD-diagonal-maximum"] == enum["RB_GMM_SIMD_DIAG_MAX"]  # name as in src/Mm/Module.cc:88
