"""Regenerates the committed golden fixtures from the CPU oracle (and, for the FFT, from the
reference's own object code in oracle/_ref).  Run from the repo root:  python tests/golden/make_golden.py

The fixtures pin the oracle against silent drift and travel to the GPU box, where /root/reference
does not exist."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402
from rasr_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    o.build(ref=True)
    # 1. MFCC+derivatives of a 2 s C1-style utterance
    n, seed = 32000, 1234
    r = o.mfcc(o.frontend_cfg(), synth.utterance(n, seed))
    np.savez_compressed(os.path.join(HERE, "mfcc_c1_2s.npz"), n_samples=n, seed=seed, feats=r["feats"],
                        t_start=r["t_start"], t_end=r["t_end"])
    # 2. FFT vectors produced by the REFERENCE's object code (strict build)
    ref = o.ref_fft(native=False)
    if ref is not None:
        rng = np.random.default_rng(99)
        x = (rng.standard_normal((8, 512)) * 3000).astype(np.float32)
        x[:, 400:] = 0
        y = x.copy()
        for row in y:
            ref.ref_fft_transform_real(row.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(512))
        np.savez_compressed(os.path.join(HERE, "fft512_reference.npz"), x=x, y=y)
    # 3. GMM scores of a small ragged model (all three scorers)
    msd = synth.ragged_mixture_set(dim=39, n_covariances=1)
    ms = o.MixtureSet(**msd)
    f = synth.features(64, 39, seed=5)
    batch = o.gmm_batch_float(ms, f)
    mx, mb = o.gmm_diag_max(ms, f)
    sm, sb = o.gmm_diag_sum(ms, f)
    np.savez_compressed(os.path.join(HERE, "gmm_ragged.npz"), batch=batch, max=mx, max_best=mb, sum=sm, sum_best=sb)
    # 4. the quantised and the density-preselection scorers on the same ragged model (few clusters: contested ties)
    sel_f, cl_f, means_f = o.gmm_preselect_float(ms, f, clusters=16, select=4)
    sel_i, cl_i = o.gmm_preselect_int(ms, f, clusters=16, select=4)
    np.savez_compressed(os.path.join(HERE, "gmm_ragged_int_presel.npz"), int=o.gmm_batch_int(ms, f), presel_float=sel_f,
                        presel_float_cluster_of=cl_f, presel_float_means=means_f, presel_int=sel_i,
                        presel_int_cluster_of=cl_i)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
