#!/bin/bash
cat > /tmp/tr.py <<'PY'
import os, time, numpy as np, torch
from rasr_b200 import flow, synth
samples_h, offs = synth.corpus(125, n_samples=160240, seed0=3000)
fe = flow.FrontEnd()
T = int(fe.count_frames(offs)[-1])
h_pcm = torch.from_numpy(samples_h.astype(np.int16)).pin_memory()
h_feats = torch.empty((T, 39), dtype=torch.float32).pin_memory()
for _ in range(3): fe.process_s16(h_pcm, offs, timestamps=False, out=h_feats)
os.environ["RB_TRACE"] = "1"
t = time.perf_counter(); fe.process_s16(h_pcm, offs, timestamps=False, out=h_feats); print("call %.3f ms" % ((time.perf_counter() - t) * 1e3))
PY
PYTHONPATH=$PWD python /tmp/tr.py 2>&1 | tail -14
RB_PIPE_SLAB0=2048 RB_PIPE_SLABMAX=16384 PYTHONPATH=$PWD timeout 300 python /tmp/pe.py 2>&1 | grep " ms"
