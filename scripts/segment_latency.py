"""Device-side time of ONE segment-sized call (1000 frames = a 10 s utterance) of every scorer mode, the Nn scorer and the
front-end: what a RASR adapter pays per segment besides the PCIe copies.  python scripts/segment_latency.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rasr_b200 import flow, mm, nn, synth  # noqa: E402

st = torch.cuda.Stream()
torch.cuda.set_stream(st)
sp = st.cuda_stream


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


msd = synth.mixture_set()
for T in (300, 1000, 3000):
    f = torch.from_numpy(synth.features(T, 39)).cuda()
    out = torch.empty((T, 256), dtype=torch.float32, device="cuda")
    row = ["T=%d" % T]
    for mode in ("batch-float", "diagonal-maximum", "batch-int", "SIMD-diagonal-maximum", "batch-tensor",
                 "preselection-batch-float", "preselection-batch-int"):
        sc = mm.GmmScorer(mm.MixtureSet.from_dict(msd), mode)
        row.append("%s %.0f us" % (mode, timed(lambda: sc.score_dev(f, T, out, None, sp))))
    print(" | ".join(row), flush=True)
net = synth.network()
scn = nn.NnScorer(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 1.0, "bf16")
for T in (300, 1000, 3000):
    x = torch.from_numpy(synth.features(T, 429, seed=3, scale=1.0)).cuda()
    o = torch.empty((T, 12000), dtype=torch.float32, device="cuda")
    print("Nn T=%d: %.0f us" % (T, timed(lambda: scn.score_dev(x, T, o, sp))), flush=True)
fe = flow.FrontEnd()
for secs in (3, 10, 30):
    s, offs = synth.corpus(1, n_samples=16000 * secs + 240, seed0=5)
    d = torch.from_numpy(s).cuda()
    T = int(fe.count_frames(offs)[-1])
    feats = torch.empty((T, 39), dtype=torch.float32, device="cuda")
    print("front-end %d s (%d frames): %.0f us" % (secs, T, timed(lambda: fe.process_dev(d, offs, feats, sp))), flush=True)
