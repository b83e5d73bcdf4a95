"""The RASR-side adapters (adapters/*.cc) are header-checked against the reference's own headers whenever a
reference checkout is present (it is in the build container, not on the GPU box); see INTEGRATION.md."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RASR_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "Mm")), reason="no reference checkout")
def test_adapters_compile_against_reference_headers():
    r = subprocess.run(["bash", os.path.join(ROOT, "adapters", "check_syntax.sh"), REF], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_adapters_only_call_declared_abi():
    """every rb_* identifier the adapters use is declared in include/rasr_b200.h"""
    import re
    hdr = open(os.path.join(ROOT, "include", "rasr_b200.h")).read()
    declared = set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", hdr)) | set(re.findall(r"\b(rb_[a-z0-9_]+)\b", hdr))
    used = set()
    for f in os.listdir(os.path.join(ROOT, "adapters")):
        if f.endswith((".cc", ".hh")):
            used |= set(re.findall(r"\b(rb_[a-z0-9_]+)\b", open(os.path.join(ROOT, "adapters", f)).read()))
    assert used and used <= declared, sorted(used - declared)
