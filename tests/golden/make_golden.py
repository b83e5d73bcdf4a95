"""Regenerates the committed golden fixtures FROM THE REFERENCE'S OWN OBJECT CODE (oracle/_ref/librasr_ref*.so: the
reference's Core / Flow / Math / Signal / Mm sources compiled from /root/reference, see oracle/refbuild/Makefile):
the MFCC features come out of a Flow::Network that the reference's NetworkParser builds from the reference's own
mfcc.flow + derivationWithRegression.flow, the scores out of feature scorers made by the reference's Mm factory.
Nothing in these files was computed by the oracle restatement or by the CUDA engine.

Run from the repo root, in the build container (needs /root/reference):  python tests/golden/make_golden.py

"strict" = -ffp-contract=off (every operation rounded separately), "native" = gcc's default contraction with FMA
available, i.e. the reference's default -march=native build.  The fixtures travel to the GPU box, where
/root/reference does not exist; tests/test_oracle_*.py pin the oracle to them, tests/test_gpu_*.py the CUDA path."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402  (only MixtureSet: the C layout shared with the reference shim)
from oracle import pyref  # noqa: E402
from rasr_b200 import io, synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = "reference object code (oracle/_ref/librasr_ref.so = strict, librasr_ref_native.so = native)"


def scorers(ms, f, names, config=None):
    out = {}
    for name in names:
        for native in (False, True):
            key = "%s/%s" % (name, "native" if native else "strict")
            sc = pyref.FeatureScorer(ms, name, config, native=native)
            if name.startswith("diagonal") or name.startswith("SIMD"):
                out[key], out[key + "/best"] = sc.score(f, want_best=True)
            else:
                out[key] = sc.score(f)
    return out


def simd_scorer():
    """4c. "SIMD-diagonal-maximum" (scores and best densities) on the three models above -> ref_gmm_simd.npz
    (python tests/golden/make_golden.py simd writes this file alone)"""
    d = dict(source=SOURCE)
    cases = (("c2", synth.mixture_set(), synth.features(100000, 39)[:96]),
             ("ragged", synth.ragged_mixture_set(dim=39, n_covariances=1), synth.features(64, 39, seed=5)),
             ("ragged_3cov", synth.ragged_mixture_set(dim=24, n_covariances=3, seed=11), synth.features(64, 24, seed=6)))
    for tag, msd, f in cases:
        for k, v in scorers(o.MixtureSet(**msd), f, ["SIMD-diagonal-maximum"]).items():
            d["%s/%s" % (tag, k)] = v
    np.savez_compressed(os.path.join(HERE, "ref_gmm_simd.npz"), **d)


def main():
    o.build(ref=True)
    pyref.build()
    if sys.argv[1:] == ["simd"]:
        return simd_scorer()
    # 1. BASELINE config C1: the 10 s utterance through the reference's own flow files, 999 x 39
    n, seed = 160000, 1234
    x = synth.utterance(n, seed)
    P = {"nr-cepstrum-coefficients": 13, "block-size": 4096}
    d = dict(n_samples=n, seed=seed, source=SOURCE)
    for native in (False, True):
        net = pyref.FlowNetwork("mfcc_derivatives.flow", P, native=native)
        r, c = net.run(x), net.run(x, port="cepstra")
        tag = "native" if native else "strict"
        d.update({"feats_" + tag: r["feats"], "cepstra_" + tag: c["feats"]})
        d.update(t_start=r["t_start"], t_end=r["t_end"], cepstra_t_start=c["t_start"], cepstra_t_end=c["t_end"])
    np.savez_compressed(os.path.join(HERE, "ref_mfcc_c1.npz"), **d)
    # 1b. every stage of a short utterance (the self-contained chain: same nodes as mfcc.flow)
    n2, seed2 = 8240, 77
    x2 = synth.utterance(n2, seed2)
    net = pyref.FlowNetwork("mfcc_chain_plain.flow", pyref.chain_parameters())
    d = dict(n_samples=n2, seed=seed2, source=SOURCE)
    for port in ("frames", "spectrum", "amplitude", "filterbank", "cepstra", "features"):
        d[port] = net.run(x2, port=port)["feats"]
    np.savez_compressed(os.path.join(HERE, "ref_mfcc_stages.npz"), **d)
    # 1c. signal-dc-detection in front of the chain
    x3 = synth.utterance(40000, seed=9)
    x3[5000:9000] = x3[4999]
    x3[15000:15300] = 7.0
    x3[20000:20250] = -3.0
    r = pyref.FlowNetwork("mfcc_chain_dc.flow", pyref.chain_parameters(dc=True)).run(x3)
    np.savez_compressed(os.path.join(HERE, "ref_mfcc_dc.npz"), samples=x3.astype(np.int16), feats=r["feats"],
                        t_start=r["t_start"], t_end=r["t_end"], source=SOURCE)
    # 2. FFT vectors of the reference's FFT translation unit alone (strict build)
    ref = o.ref_fft(native=False)
    rng = np.random.default_rng(99)
    xf = (rng.standard_normal((8, 512)) * 3000).astype(np.float32)
    xf[:, 400:] = 0
    y = xf.copy()
    for row in y:
        ref.ref_fft_transform_real(row.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(512))
    np.savez_compressed(os.path.join(HERE, "fft512_reference.npz"), x=xf, y=y)
    # 3. BASELINE config C2's model (39 dims, 4096 densities, 256 mixtures), the first 96 of its 100 000 frames
    ms = o.MixtureSet(**synth.mixture_set())
    f = synth.features(100000, 39)[:96]
    names = ["batch-diagonal-maximum-float", "batch-diagonal-maximum-int", "preselection-batch-float",
             "preselection-batch-int", "diagonal-maximum", "diagonal-sum"]
    np.savez_compressed(os.path.join(HERE, "ref_gmm_c2.npz"), n_frames=96, source=SOURCE, **scorers(ms, f, names))
    # 4. a small ragged model (mixtures of unequal size sharing densities out of order); few clusters: contested ties
    ms = o.MixtureSet(**synth.ragged_mixture_set(dim=39, n_covariances=1))
    f = synth.features(64, 39, seed=5)
    cfg = {"density-clustering.clusters": 16, "density-clustering.select-clusters": 4}
    d = scorers(ms, f, ["batch-diagonal-maximum-float", "batch-diagonal-maximum-int", "diagonal-maximum", "diagonal-sum"])
    d.update(scorers(ms, f, ["preselection-batch-float", "preselection-batch-int"], cfg))
    np.savez_compressed(os.path.join(HERE, "ref_gmm_ragged.npz"), source=SOURCE, **d)
    # 4b. several covariances (diagonal scorers only)
    ms = o.MixtureSet(**synth.ragged_mixture_set(dim=24, n_covariances=3, seed=11))
    f = synth.features(64, 24, seed=6)
    np.savez_compressed(os.path.join(HERE, "ref_gmm_ragged_3cov.npz"), source=SOURCE,
                        **scorers(ms, f, ["diagonal-maximum", "diagonal-sum"]))
    simd_scorer()
    # 5. post-processing: signal-normalization -> sequence concatenation -> matrix multiplication (lda.flow wiring)
    ff = synth.features(300, 13, seed=5)
    M = np.random.default_rng(3).standard_normal((20, 65)).astype(np.float32)
    path = "bin:/tmp/make_golden_lda.bin"
    io.write_matrix(path, M)
    d = dict(source=SOURCE, matrix=M)
    for kind, length, right in (("mean-and-variance", "infinite", "infinite"), ("mean", 51, 25)):
        net = pyref.FlowNetwork("postproc_chain.flow", {"block-size": 13, "norm-type": kind, "norm-length": length,
                                                         "norm-right": right, "splice-length": 5, "splice-right": 2,
                                                         "matrix-file": path})
        for port in ("normalized", "spliced", "projected"):
            d["%s/%s/%s" % (kind, length, port)] = net.run(ff.reshape(-1), port=port, sample_rate=1300.0)["feats"]
    np.savez_compressed(os.path.join(HERE, "ref_postproc.npz"), **d)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
