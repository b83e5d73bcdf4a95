import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _no_nvidia_device():
    """True only when the host has no NVIDIA device at all.  With a device present the gpu tests always run, so a
    missing or broken librasr_b200.so fails loudly instead of hiding behind a skip."""
    return not (os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0") or os.path.exists("/dev/dxg"))


def pytest_collection_modifyitems(config, items):
    if not _no_nvidia_device():
        return
    skip = pytest.mark.skip(reason="no NVIDIA device on this host (the engine has no CPU path)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand, checker only."""
    from oracle import pyoracle

    pyoracle.build(ref=True)
    return pyoracle


@pytest.fixture(scope="session")
def diag():
    """Append one JSON line per measurement to gpurun_out/diag.jsonl (comes back from the GPU box)."""
    path = os.path.join(ROOT, "gpurun_out")
    os.makedirs(path, exist_ok=True)
    f = open(os.path.join(path, "diag.jsonl"), "a")

    def report(name, **kv):
        kv = {k: (v.item() if hasattr(v, "item") else v) for k, v in kv.items()}
        f.write(json.dumps(dict(test=name, **kv)) + "\n")
        f.flush()

    yield report
    f.close()


@pytest.fixture(scope="session")
def reference_search_golden():
    """tests/golden/ref_search.npz: tracebacks of the reference's own LinearSearch (tests/golden/make_golden_search.py)"""
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_search.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    for c in cases.values():
        c["entry_model"] = 0
        c["single_word"] = bool(c["single_word"])
    return cases


@pytest.fixture(scope="session")
def simd_golden_cases():
    """models and frames behind tests/golden/ref_gmm_simd.npz (tests/golden/make_golden.py simd_scorer)"""
    from rasr_b200 import synth
    return {"c2": lambda: (synth.mixture_set(), synth.features(100000, 39)[:96]),
            "ragged": lambda: (synth.ragged_mixture_set(dim=39, n_covariances=1), synth.features(64, 39, seed=5)),
            "ragged_3cov": lambda: (synth.ragged_mixture_set(dim=24, n_covariances=3, seed=11),
                                    synth.features(64, 24, seed=6))}
