"""Model and cache FILES written by the reference's own code (oracle/_ref/librasr_ref.so, oracle/refbuild/ref_io.cc and
the generic-cache node of the reference's Flow library) -> tests/golden/ref_io/.  tests/test_ref_io.py reads them with
rasr_b200/io.py and rasr_b200/cache.py and compares with what the reference's readers returned for the same files
(expected.npz).  Run from the repo root in the build container (needs /root/reference):

    python tests/golden/make_golden_io.py
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402  (MixtureSet: the C layout shared with the reference shim)
from oracle import pyref  # noqa: E402
from rasr_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_io")


def estimator_case(seed=3, T=3000):
    """6 mixtures x 4 densities; two densities see fewer than 5 frames (the reader must drop them as the reference's
    estimate() does), one mixture is fed through a single density"""
    msd = synth.mixture_set(dim=13, n_mixtures=6, densities_per_mixture=4, seed=seed)
    rng = np.random.default_rng(seed)
    mof = rng.integers(0, 6, T).astype(np.uint32)
    k = rng.integers(0, 4, T)
    k[(mof == 1) & (k == 2)] = 0           # density 2 of mixture 1: no observation at all
    rare = np.flatnonzero((mof == 4) & (k == 3))
    k[rare[3:]] = 1                        # density 3 of mixture 4: three observations (< minimum-observation-weight)
    k[mof == 5] = 2                        # mixture 5: everything on one density
    feats = (msd["means"].reshape(6, 4, 13)[mof, k] + 0.3 * rng.standard_normal((T, 13))).astype(np.float32)
    return msd, feats, mof


def main():
    o.build(ref=True)
    pyref.build()
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    exp = {}
    # mixture text files: the reference's default 6 digits, and 9 digits (exact f32 round trip)
    msd = synth.ragged_mixture_set(dim=13, sizes=(1, 3, 16, 7, 32, 2), seed=7)
    oms = o.MixtureSet(**msd)
    for name, prec in (("ragged_p6.pms", 6), ("ragged_p9.pms", 9), ("ragged_p9.pms.gz", 9)):
        pyref.write_mixture_text(oms, os.path.join(OUT, name), precision=prec)
        for k, v in pyref.read_mixture_file(os.path.join(OUT, name)).items():
            exp["%s/%s" % (name, k)] = np.asarray(v)
    # accumulator file of a Viterbi pass (what a trained system's .mix file is)
    msd, feats, mof = estimator_case()
    pyref.write_mixture_estimator(o.MixtureSet(**msd), feats, mof, os.path.join(OUT, "viterbi.mix"))
    for k, v in pyref.read_mixture_file(os.path.join(OUT, "viterbi.mix")).items():
        exp["viterbi.mix/%s" % k] = np.asarray(v)
    # Math::Matrix / Math::Vector
    rng = np.random.default_rng(11)
    m = rng.standard_normal((5, 7)).astype(np.float32)
    m[0, 0], m[1, 1], m[2, 2] = 0.0, np.float32(1e-30), np.float32(-3.4e38)
    for q in ("bin", "xml"):
        pyref.write_matrix("%s:%s" % (q, os.path.join(OUT, "matrix." + q)), m)
        pyref.write_vector("%s:%s" % (q, os.path.join(OUT, "vector." + q)), m[1])
    exp["matrix"], exp["vector"] = m, m[1]
    # Flow caches written by the reference's generic-cache node: file archive (plain / gathered + compressed), directory
    for name, gather, compress in (("plain.cache", 4294967295, "false"), ("gather7_gz.cache", 7, "true"),
                                   ("dir.cache/", 4294967295, "false")):
        P = pyref.chain_parameters()
        P.update({"path": os.path.join(OUT, name), "gather": gather, "compress": compress, "id": "corpus/rec1/seg1"})
        net = pyref.FlowNetwork("mfcc_cache.flow", P)
        r1 = net.run(synth.utterance(8240, 5), width=39)
        net.set_parameter("id", "corpus/rec1/seg2")
        r2 = net.run(synth.utterance(4240, 6), width=39, start_time=1.5)
        net.close()
        for seg, r in (("seg1", r1), ("seg2", r2)):
            exp["%s/%s/feats" % (name.rstrip("/"), seg)] = r["feats"]
            exp["%s/%s/times" % (name.rstrip("/"), seg)] = np.stack([r["t_start"], r["t_end"]], axis=1)
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **exp)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
