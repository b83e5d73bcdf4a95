#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/pe.py <<'PY'
import os, sys, time, numpy as np, torch
from rasr_b200 import flow, mm, pipeline, synth
samples_h, offs = synth.corpus(125, n_samples=160240, seed0=3000)
fe = flow.FrontEnd()
T = int(fe.count_frames(offs)[-1])
h_pcm = torch.from_numpy(samples_h.astype(np.int16)).pin_memory()
h_f32 = torch.from_numpy(samples_h).pin_memory()
h_scores = torch.empty((T, 256), dtype=torch.float32).pin_memory()
h_feats = torch.empty((T, 39), dtype=torch.float32).pin_memory()
ms = mm.MixtureSet.from_dict(synth.mixture_set())
def timeit(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
for mode in ("batch-float", "batch-tensor"):
    sc = mm.GmmScorer(ms, mode)
    f = synth.features(T, 39, seed=1)
    hf = torch.from_numpy(f).pin_memory()
    print(mode, "gmm-only e2e %.3f ms" % timeit(lambda: sc.score(hf, out=h_scores)))
    print(mode, "pipeline s16 %.3f ms" % timeit(lambda: pipeline.score_utterances(fe, sc, h_pcm, offs, out=h_scores, pcm_channels=1)))
    print(mode, "pipeline f32 %.3f ms" % timeit(lambda: pipeline.score_utterances(fe, sc, h_f32, offs, out=h_scores)))
print("frontend s16 %.3f ms" % timeit(lambda: fe.process_s16(h_pcm, offs, timestamps=False, out=h_feats)))
PY
for cfg in "2048 16384" "2048 32768" "8192 8192" "16384 65536" "1024 4096"; do
  set -- $cfg
  echo "== slab0 $1 slabmax $2"
  RB_PIPE_SLAB0=$1 RB_PIPE_SLABMAX=$2 PYTHONPATH=$PWD timeout 300 python /tmp/pe.py 2>&1 | grep " ms"
done
