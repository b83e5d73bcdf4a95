#!/usr/bin/env python
"""bench.py -- acoustic frames scored/sec on B200 (BASELINE.json metric), one JSON line per run.

Headline workload (N=1) = BASELINE config C2: GMM FeatureScorer, 39-dim MFCC, 4096 diagonal Gaussians in
256 mixtures, 100 000 frames.  A "step" = one dense scoring pass of the hot path over that batch.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU scorer on the host cores

  value     frames/s with inputs resident in HBM (device pointers, rb_gmm_score_dev)
  e2e       frames/s through the host-buffer C-ABI call rb_gmm_score, caller buffers in page-locked memory from
            rb_host_alloc (what adapters/B200FeatureScorer does); e2e.pageable = the same call on pageable buffers;
            e2e.pcie_ceiling = a bare concurrent H2D + D2H of the same bytes on all ranks at once
  roofline  dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference arm (below) on a bounded sample, run in the same job
  workloads the other BASELINE configs, each with value / e2e / roofline / clocks:
            C3 (audio -> MFCC -> GMM scores), C4 (Nn 429 -> 6 x 2048 -> 12000, bf16), C5 (audio -> ... -> LinearSearch)

Reference arm: Mm::BatchFloatFeatureScorer of the reference itself -- its sources compiled into
oracle/_ref/librasr_ref_native.so (oracle/refbuild/Makefile), created by its own Mm factory and driven through the
recognizer's buffered call protocol -- one process per host core over frame partitions (the reference's only model of
parallelism, src/Bliss/CorpusDescription.cc:173-180).  `kind` is "port" (the oracle restatement) only where
oracle/_ref could not be built.

Single workloads for profiling: --workload gmm | gmm-diag | gmm-tensor | gmm-int | gmm-presel | gmm-presel-int |
frontend | pipeline | pipeline-nn | pipeline-search | nn   (prints that workload's line alone)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "acoustic_frames_scored_per_sec"
UNIT = "frames/s"
C2 = dict(dim=39, n_mixtures=256, densities_per_mixture=16, frames=100000)
C2_WORKLOAD = "C2: GMM FeatureScorer (batch-float), 39-dim, 4096 densities / 256 mixtures, 100000 frames per GPU"
ALGO_BYTES_PER_FRAME = {"gmm": 39 * 4 + 256 * 4, "frontend": 160 * 4 + 39 * 4, "pipeline": 160 * 4 + 256 * 4,
                        "nn": 429 * 4 + 12000 * 4}
NN_FLOP_PER_FRAME = 2 * (429 * 2048 + 5 * 2048 * 2048 + 2048 * 12000)
GMM_WORKLOADS = {"gmm": "batch-float", "gmm-diag": "diagonal-maximum", "gmm-tensor": "batch-tensor",
                 "gmm-int": "batch-int", "gmm-presel": "preselection-batch-float",
                 "gmm-presel-int": "preselection-batch-int"}
ALL_WORKLOADS = list(GMM_WORKLOADS) + ["frontend", "pipeline", "pipeline-nn", "pipeline-search", "nn"]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))),
                    source="measured")
    return dict(hbm=6650.0, bf16=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (one process for the whole run; every
    workload reads the rows between its own mark() and section())."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_ready(self, timeout=3.0):
        """nvidia-smi needs a moment to start: block until its first row has arrived"""
        t = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.005)

    def mark(self):
        """rows from here on count (start of a timed region)"""
        self.first = len(self.rows)

    def n_since_mark(self):
        return len(self.rows) - self.first

    def section(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    samples=len(sm), reasons=sorted(reasons))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()


# --------------------------------------------------------------------------------------------
# Reference arm / CPU baseline: the reference's own Mm::BatchFloatFeatureScorer on the host cores
# --------------------------------------------------------------------------------------------

def _reference_kind():
    from oracle import pyref

    if pyref.available(native=True):
        return "reference"
    if os.path.isdir(os.path.join(pyref.REFERENCE, "src", "Mm")):
        try:
            pyref.build()
            return "reference"
        except Exception:  # noqa: BLE001 -- a broken toolchain must not take the bench down
            pass
    return "port"


def _reference_worker(kind, lo, hi, n_frames, go, done, stop, result):
    """One process = one RASR job on a corpus partition: frames [lo, hi) of every step's sample."""
    from oracle import pyoracle as o
    from rasr_b200 import synth

    ms = o.MixtureSet(**synth.mixture_set())
    feats = synth.features(C2["frames"], 39)
    if kind == "reference":
        from oracle import pyref

        scorer = pyref.FeatureScorer(ms, "batch-diagonal-maximum-float", native=True)
        score = scorer.score
    else:
        o.build(ref=False)
        score = lambda f: o.gmm_batch_float(ms, f, threads=1)  # noqa: E731
    score(feats[:64])
    while True:
        go.wait()
        if stop.value:
            return
        n = n_frames.value
        a, b = lo * n // 1000000, hi * n // 1000000  # lo / hi are in millionths of the sample
        t = time.perf_counter()
        if b > a:
            s = score(feats[a:b])
            result.value = float(s[0, 0])
        done.wait()
        del t


class ReferencePool:
    """P worker processes, each with its own scorer object; step(n) scores the first n C2 frames, split evenly."""

    def __init__(self, kind, procs):
        import multiprocessing as mp

        ctx = mp.get_context("fork")
        self.procs = procs
        self.go, self.done = ctx.Barrier(procs + 1), ctx.Barrier(procs + 1)
        self.n, self.stop = ctx.Value("l", 0), ctx.Value("i", 0)
        self.result = ctx.Value("d", 0.0)
        self.workers = []
        for p in range(procs):
            lo, hi = p * 1000000 // procs, (p + 1) * 1000000 // procs
            w = ctx.Process(target=_reference_worker, args=(kind, lo, hi, self.n, self.go, self.done, self.stop, self.result),
                            daemon=True)
            w.start()
            self.workers.append(w)

    def step(self, n):
        self.n.value = int(n)
        self.go.wait()
        t = time.perf_counter()
        self.done.wait()
        return time.perf_counter() - t

    def close(self):
        self.stop.value = 1
        self.go.wait()
        for w in self.workers:
            w.join(timeout=5)


def reference_measurement(steps, warmup, seconds_per_step, one_thread_seconds=3.0):
    """(mean seconds per step, frames per step, dict for the JSON line) of the reference's CPU scorer on C2."""
    kind = _reference_kind()
    cores = len(os.sched_getaffinity(0))
    pool = ReferencePool(kind, cores)
    try:
        pool.step(64 * cores)  # page in, first-touch
        probe = 256 * cores
        rate = probe / pool.step(probe)
        n = int(min(C2["frames"], max(1024, rate * seconds_per_step)))
        for _ in range(warmup):
            pool.step(n)
        times = [pool.step(n) for _ in range(steps)]
    finally:
        pool.close()
    one = ReferencePool(kind, 1)
    try:
        one.step(64)
        r1 = 512 / one.step(512)
        n1 = int(max(512, min(C2["frames"], r1 * one_thread_seconds / 3)))
        one_best = max(n1 / one.step(n1) for _ in range(3))
    finally:
        one.close()
    mean = float(np.mean(times))
    what = ("Mm::BatchFloatFeatureScorer compiled from the reference's sources (oracle/_ref/librasr_ref_native.so: "
            "gnu++20 -O2 -msse3, AVX2/FMA code generation as -march=native gives on this class of host), created by its "
            "own factory, recognizer call protocol"
            if kind == "reference" else "oracle port of Mm::BatchFloatFeatureScorer (oracle/_ref not built)")
    info = dict(value=n / mean, unit=UNIT, cores=cores, kind=kind,
                sample="first %d of the 100000 C2 frames per step, all 256 mixtures, %d processes over frame "
                       "partitions, mean of %d steps after %d warm-up steps" % (n, cores, steps, warmup),
                best=n / min(times), one_thread=one_best, implementation=what)
    return mean, n, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = 150.0 / max(1, args.steps + args.warmup)  # the whole run ends within a few minutes
    mean, n, info = reference_measurement(args.steps, args.warmup, args.cpu_seconds or min(budget, 10.0))
    line = dict(metric=METRIC, value=info["value"], unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=mean * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", config=dict(workload=C2_WORKLOAD, frames_per_step=n),
                cpu_baseline=info,
                e2e=dict(value=info["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def cpu_baseline_same_job():
    """The reference arm on a bounded sample, run as a child of this job (its own process tree: the workers fork
    before anything touches CUDA).  One number for `cpu_baseline` and for the reference arm."""
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup",
                          "1", "--cpu-seconds", "4"], capture_output=True, text=True, timeout=600)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    return dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="port", sample="reference arm failed: " + out.stderr[-300:])


def cpu_baseline_frontend(seconds=10.0):
    """Oracle port of the Flow MFCC chain (mfcc.flow + derivationWithRegression.flow), one utterance per task on all host
    cores -- the reference's own model of parallelism (one process per corpus partition)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import pyoracle as o
    from rasr_b200 import synth

    o.build(ref=False)
    cores = os.cpu_count() or 1
    cfg = o.frontend_cfg()
    utt = [synth.utterance(160240, seed=3000 + u) for u in range(cores)]
    t = time.perf_counter()
    o.mfcc(cfg, utt[0])
    per_utt = time.perf_counter() - t
    rounds = int(max(1, min(400, seconds / max(per_utt, 1e-3))))
    with ThreadPoolExecutor(cores) as pool:  # the ctypes call releases the GIL
        t = time.perf_counter()
        frames = sum(r["feats"].shape[0] for _ in range(rounds) for r in pool.map(lambda x: o.mfcc(cfg, x), utt))
        dt = time.perf_counter() - t
    return dict(value=frames / dt, unit=UNIT, cores=cores, kind="port",
                sample="%d utterances of 1000 frames (%d per core), one utterance per thread" % (rounds * cores, rounds))


def cpu_baseline_nn(seconds=10.0):
    """The reference's Nn CPU path is cblas_sgemm per layer (src/Math/Blas.hh:410-421) with whatever BLAS the system
    has; here numpy's OpenBLAS sgemm on all host cores, whole-segment batches (the favourable case for the CPU), f32."""
    from rasr_b200 import synth

    cores = os.cpu_count() or 1
    net = synth.network()
    ws = [np.ascontiguousarray(w.T) for w in net["weights"]]

    def forward(x):
        h = x
        for l, (w, b) in enumerate(zip(ws, net["biases"])):
            h = h @ w + b
            if l + 1 < len(ws):
                h = np.maximum(h, 0.0, out=h) if net["acts"][l] in ("relu", "rectified") else 1.0 / (1.0 + np.exp(-h))
        return -(h - net["log_prior"])

    x = synth.features(2048, 429, seed=4, scale=1.0)
    forward(x[:256])
    t = time.perf_counter()
    forward(x)
    rate = x.shape[0] / (time.perf_counter() - t)
    n = int(max(2048, min(65536, rate * seconds)))
    x = synth.features(n, 429, seed=5, scale=1.0)
    t = time.perf_counter()
    for a in range(0, n, 4096):
        forward(x[a:a + 4096])
    dt = time.perf_counter() - t
    return dict(value=n / dt, unit=UNIT, cores=cores, kind="port",
                sample="%d frames in batches of 4096 through numpy / OpenBLAS sgemm (f32), all cores" % n)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(torch, index):
    """One process per GPU, like `numactl --cpunodebind` in front of a RASR job: run this rank (and first-touch its pinned
    host buffers) on the NUMA node its GPU hangs off.  With 8 ranks streaming scores to the host at once the PCIe
    traffic otherwise crosses the socket interconnect.  Returns the node, or None if it cannot be determined."""
    try:
        p = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError):
        return None


class stdout_to_stderr:
    """fd-level redirect: libraries that print to stdout (NCCL's version banner at communicator creation) must not get
    in front of the one JSON line this program owes its caller"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class Bench:
    """Set-up and timing of one workload on this rank's GPU."""

    R = 4  # rotating buffer sets so that no step finds its data in L2

    def __init__(self, args, torch, dist, rank, local_rank, world, sampler):
        self.args, self.torch, self.dist = args, torch, dist
        self.rank, self.local_rank, self.world, self.sampler = rank, local_rank, world, sampler
        self.dev = torch.device("cuda", local_rank)
        # a non-default stream: the C ABI treats a NULL stream as "the handle's own stream", and the
        # events below must sit on the stream the kernels are launched on
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.sptr = self.stream.cuda_stream
        assert self.sptr != 0
        self.peaks = measured_peaks()

    def barrier(self):
        # drain the device first: the library's own exchange (rb_comm_*: NCCL on a second communicator, spin barriers)
        # must never be in flight together with torch's NCCL barrier
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def pinned(self, a):
        """copy of a numpy array in page-locked memory from the library's own allocator (rb_host_alloc)"""
        from rasr_b200 import capi

        h = capi.host_empty(a.shape, a.dtype)
        h[...] = a
        return h

    # ---------------- workload set-up: every rank owns its own shard (weak scaling, no collective)
    def setup(self, wl, frames=0):
        torch, dev, sptr, R, rank, local_rank = self.torch, self.dev, self.sptr, self.R, self.rank, self.local_rank
        from rasr_b200 import capi, flow, mm, nn, pipeline, synth

        w = dict(name=wl, e2e_pageable=None)
        if wl in GMM_WORKLOADS:
            T = frames or C2["frames"]
            msd = synth.mixture_set()
            mode = GMM_WORKLOADS[wl]
            scorer = mm.GmmScorer(mm.MixtureSet.from_dict(msd), mode, device=local_rank)
            feats_h = synth.features(T, 39, seed=2024 + rank)
            d_in = [torch.from_numpy(feats_h).to(dev) for _ in range(R)]
            d_out = [torch.empty((T, 256), dtype=torch.float32, device=dev) for _ in range(R)]
            w["step"] = lambda i: scorer.score_dev(d_in[i % R], T, d_out[i % R], None, sptr)
            h_in, h_out = self.pinned(feats_h), capi.host_empty((T, 256), np.float32)
            w["e2e"] = lambda: scorer.score(h_in, out=h_out)
            p_out = np.empty((T, 256), np.float32)
            w["e2e_pageable"] = lambda: scorer.score(feats_h, out=p_out)
            w.update(scorer=scorer, h2d=T * 39 * 4, d2h=T * 256 * 4, units=T, algo_bytes=ALGO_BYTES_PER_FRAME["gmm"] * T, bound="hbm",
                     dtype="u8/s32" if wl in ("gmm-int", "gmm-presel-int") else "f32",
                     workload=C2_WORKLOAD.replace("batch-float", mode).replace("100000", str(T)),
                     keep=(scorer, d_in, d_out, msd))
        elif wl in ("frontend", "pipeline"):
            n_utt = max(1, frames // 1000) if frames else 125
            samples_h, offs = synth.corpus(n_utt, n_samples=160240, seed0=3000 + 1000 * rank)
            fe = flow.FrontEnd(device=local_rank)
            T = int(fe.count_frames(offs)[-1])
            d_samples = [torch.from_numpy(samples_h).to(dev) for _ in range(R)]
            d_feats = [torch.empty((T, 39), dtype=torch.float32, device=dev) for _ in range(R)]
            # end to end the audio arrives as 16-bit PCM (what the audio nodes deliver): 2 bytes per sample over PCIe
            pcm = samples_h.astype(np.int16)
            h_pcm = self.pinned(pcm)
            if wl == "frontend":
                w["step"] = lambda i: fe.process_dev(d_samples[i % R], offs, d_feats[i % R], sptr)
                h_feats, p_feats = capi.host_empty((T, 39), np.float32), np.empty((T, 39), np.float32)
                w["e2e"] = lambda: fe.process_s16(h_pcm, offs, timestamps=False, out=h_feats)
                w["e2e_pageable"] = lambda: fe.process_s16(pcm, offs, timestamps=False, out=p_feats)
                w.update(h2d=samples_h.size * 2, d2h=T * 39 * 4, keep=(fe, d_samples, d_feats))
            else:
                scorer = mm.GmmScorer(mm.MixtureSet.from_dict(synth.mixture_set()), device=local_rank)
                d_scores = [torch.empty((T, 256), dtype=torch.float32, device=dev) for _ in range(R)]
                w["step"] = lambda i: pipeline.score_utterances_dev(fe, scorer, d_samples[i % R], offs, d_feats[i % R],
                                                                     d_scores[i % R], sptr)
                h_scores, p_scores = capi.host_empty((T, 256), np.float32), np.empty((T, 256), np.float32)
                w["e2e"] = lambda: pipeline.score_utterances(fe, scorer, h_pcm, offs, out=h_scores, pcm_channels=1)
                w["e2e_pageable"] = lambda: pipeline.score_utterances(fe, scorer, pcm, offs, out=p_scores, pcm_channels=1)
                w.update(h2d=samples_h.size * 2, d2h=T * 256 * 4, keep=(fe, scorer, d_samples, d_feats, d_scores))
            w.update(units=T, algo_bytes=ALGO_BYTES_PER_FRAME[wl] * T, bound="hbm", dtype="f32",
                     workload="C3 shard: %s on %d utterances x 1000 frames per GPU (%d frames)" % (
                         "audio -> MFCC -> GMM scores (batch-float)" if wl == "pipeline" else "MFCC front-end", n_utt, T))
        elif wl == "pipeline-search":
            # C5: audio -> MFCC -> GMM scores -> Search::LinearSearch (1000-word synthetic lexicon), scores never leave HBM
            from rasr_b200 import search
            n_utt = max(1, (frames or 125000) // 1000)
            samples_h, offs = synth.corpus(n_utt, n_samples=160240, seed0=3000 + 1000 * rank)
            fe = flow.FrontEnd(device=local_rank)
            scorer = mm.GmmScorer(mm.MixtureSet.from_dict(synth.mixture_set()), device=local_rank)
            ls = search.LinearSearch(synth.lexicon(1000, 256), device=local_rank)
            fo = fe.count_frames(offs)
            T = int(fo[-1])
            d_samples = [torch.from_numpy(samples_h).to(dev) for _ in range(R)]
            d_feats = torch.empty((T, 39), dtype=torch.float32, device=dev)
            d_scores = torch.empty((T, 256), dtype=torch.float32, device=dev)

            def step(i):
                pipeline.score_utterances_dev(fe, scorer, d_samples[i % R], offs, d_feats, d_scores, sptr)
                ls.decode_dev(d_scores, 256, fo, sptr, want_result=False)

            pcm = samples_h.astype(np.int16)
            h_pcm = self.pinned(pcm)
            # 16-bit PCM in, word sequences out (rb_pipeline_search): slab-pipelined H2D of the audio, conversion, the
            # three stages, tracebacks to the host
            w["e2e"] = lambda: pipeline.search_utterances(fe, scorer, ls, h_pcm, offs, pcm_channels=1)
            w["e2e_pageable"] = lambda: pipeline.search_utterances(fe, scorer, ls, pcm, offs, pcm_channels=1)

            def search_alone():
                """the search kernel by itself on the resident score matrix: continuous recognition (the lexicon
                above) and the recognizer's default single-word recognition (the same words + a silence word, 2 %
                of the words irregular)"""
                lex = synth.lexicon(1000, 256)
                rs = np.random.default_rng(5)
                single = dict(lex, single_word=True)
                single["word_offsets"] = np.concatenate([[0], lex["word_offsets"] + 1]).astype(np.uint32)
                single["state_emission"] = np.concatenate([[255], lex["state_emission"]]).astype(np.uint32)
                single["state_tdp_model"] = np.concatenate([[2], lex["state_tdp_model"]]).astype(np.uint32)
                single["unigram"] = np.concatenate([[0.0], lex["unigram"]]).astype(np.float32)
                reg = np.ones(1001, np.uint8)
                reg[0] = 0
                reg[1 + rs.choice(1000, 20, replace=False)] = 0
                single["word_regular"] = reg
                out = {}
                for name, obj in (("continuous", ls), ("single_word", search.LinearSearch(single, device=local_rank))):
                    for _ in range(2):
                        obj.decode_dev(d_scores, 256, fo, sptr, want_result=False)
                    tev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    tev[0].record(self.stream)
                    for _ in range(5):
                        obj.decode_dev(d_scores, 256, fo, sptr, want_result=False)
                    tev[1].record(self.stream)
                    torch.cuda.synchronize()
                    ms = tev[0].elapsed_time(tev[1]) / 5
                    out[name] = dict(ms=ms, us_per_frame_and_segment=ms * 1e3 / 1000.0,
                                     frames_per_s=T / (ms * 1e-3))
                return out

            w["post"] = ("search_kernel_alone", search_alone)
            w.update(step=step, h2d=samples_h.size * 2, d2h=T * 20, units=T, algo_bytes=ALGO_BYTES_PER_FRAME["pipeline"] * T,
                     bound="hbm", dtype="f32", keep=(fe, scorer, ls, d_samples, d_feats, d_scores),
                     workload="C5: audio -> MFCC -> GMM scores -> LinearSearch (1000 words), %d utterances x 1000 frames "
                              "per GPU" % n_utt)
        elif wl == "pipeline-nn":
            # C4 fed from audio: MFCC -> segment CMVN -> 11-frame window -> 6 x 2048 -> 12000 senone scores
            from rasr_b200 import postproc
            n_utt = max(1, (frames or 37000) // 1000)
            samples_h, offs = synth.corpus(n_utt, n_samples=160240, seed0=3000 + 1000 * rank)
            fe = flow.FrontEnd(device=local_rank)
            pp = postproc.PostProcessor(39, "mean-and-variance", splice=(11, 5), device=local_rank)
            net = synth.network()
            sc = nn.NnScorer(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 1.0, "bf16",
                             device=local_rank)
            T = int(fe.count_frames(offs)[-1])
            d_samples = [torch.from_numpy(samples_h).to(dev) for _ in range(2)]
            d_feats = torch.empty((T, 39), dtype=torch.float32, device=dev)
            d_post = torch.empty((T, 429), dtype=torch.float32, device=dev)
            d_out = [torch.empty((T, 12000), dtype=torch.float32, device=dev) for _ in range(2)]
            w["step"] = lambda i: pipeline.nn_score_utterances_dev(fe, pp, sc, d_samples[i % 2], offs, d_feats, d_post,
                                                                    d_out[i % 2], sptr)
            h_samples, h_out = self.pinned(samples_h), capi.host_empty((T, 12000), np.float32)
            w["e2e"] = lambda: pipeline.nn_score_utterances(fe, pp, sc, h_samples, offs, out=h_out)
            w.update(h2d=samples_h.size * 4, d2h=T * 12000 * 4, units=T, algo_bytes=(160 * 4 + 12000 * 4) * T,
                     bound="tensor", dtype="bf16", keep=(fe, pp, sc, d_samples, d_feats, d_post, d_out),
                     workload="C4 from audio: MFCC -> CMVN -> 11-frame window -> Nn 429 -> 6x2048 -> 12000, %d utterances "
                              "x 1000 frames per GPU" % n_utt)
        else:  # nn
            T = frames or 65536
            net = synth.network()
            sc = nn.NnScorer(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 1.0, "bf16",
                             device=local_rank)
            x_h = synth.features(T, 429, seed=4096 + rank, scale=1.0)
            d_in = [torch.from_numpy(x_h).to(dev) for _ in range(2)]
            d_out = [torch.empty((T, 12000), dtype=torch.float32, device=dev) for _ in range(2)]
            w["step"] = lambda i: sc.score_dev(d_in[i % 2], T, d_out[i % 2], sptr)
            h_in, h_out = self.pinned(x_h), capi.host_empty((T, 12000), np.float32)
            w["e2e"] = lambda: sc.score(h_in, out=h_out)
            w.update(h2d=T * 429 * 4, d2h=T * 12000 * 4, units=T, algo_bytes=ALGO_BYTES_PER_FRAME["nn"] * T,
                     bound="tensor", dtype="bf16", keep=(sc, d_in, d_out),
                     workload="C4: Nn 429 -> 6x2048 -> 12000 senones, bf16 tcgen05, %d frames per GPU" % T)
        return w

    # ---------------- a bare concurrent H2D + D2H of the workload's bytes on every rank at once: what the host-buffer
    # call could reach if the kernels were free and the copies perfectly overlapped
    def pcie_ceiling(self, h2d, d2h, reps=5):
        torch = self.torch
        src = torch.empty(max(1, h2d), dtype=torch.uint8).pin_memory()
        dst = torch.empty(max(1, d2h), dtype=torch.uint8).pin_memory()
        d_a = torch.empty(max(1, h2d), dtype=torch.uint8, device=self.dev)
        d_b = torch.empty(max(1, d2h), dtype=torch.uint8, device=self.dev)
        s1, s2 = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)

        def once():
            with torch.cuda.stream(s1):
                d_a.copy_(src, non_blocking=True)
            with torch.cuda.stream(s2):
                dst.copy_(d_b, non_blocking=True)
            s1.synchronize()
            s2.synchronize()

        once()
        self.barrier()
        t = time.perf_counter()
        for _ in range(reps):
            once()
        self.barrier()
        return self.max_over_ranks((time.perf_counter() - t) * 1e3 / reps)

    def timed_wall(self, fn, n):
        for _ in range(2):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        self.barrier()
        return self.max_over_ranks((time.perf_counter() - t0) * 1e3 / n)

    # ---------------- timing of one workload: K steps between two events on the launching stream, max over ranks
    def measure(self, w, steps, warmup, with_ceiling=True):
        torch, stream, sampler = self.torch, self.stream, self.sampler
        from rasr_b200 import capi

        step, units, world = w["step"], w["units"], self.world
        for i in range(warmup):
            step(i)
        self.barrier()
        sampler.wait_ready()
        for i in range(min(warmup, 2)):  # the GPU idled while nvidia-smi started: back to steady state
            step(i)
        self.barrier()
        sampler.mark()
        launches0 = capi.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        ev_all = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        t_wall = time.perf_counter()
        ev_all[0].record(stream)
        for i in range(steps):
            ev[i][0].record(stream)
            step(i)
            ev[i][1].record(stream)
        ev_all[1].record(stream)
        self.barrier()
        t_wall = time.perf_counter() - t_wall
        launches = capi.launch_count() - launches0
        # dominant-kernel time: per-step event pairs on the launching stream (one step = one scoring launch)
        dev_ms = float(np.sum([a.elapsed_time(b) for a, b in ev])) / steps
        # step time: one event pair around EXACTLY K steps (inter-step gaps included), max over ranks
        wall_ms = self.max_over_ranks(t_wall * 1e3 / steps)
        ms_per_step = self.max_over_ranks(ev_all[0].elapsed_time(ev_all[1]) / steps)
        value = units * world / (ms_per_step * 1e-3)

        # end to end through the host-buffer C-ABI call
        n_e2e = max(3, min(steps, 10))
        e2e_ms = self.timed_wall(w["e2e"], n_e2e)
        e2e = dict(value=units * world / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=int(w["h2d"]),
                   d2h_bytes_per_step=int(w["d2h"]), ms_per_step=e2e_ms,
                   api="host-buffer C-ABI call, caller buffers page-locked with rb_host_alloc (as adapters/ do)")
        if w.get("e2e_pageable"):
            p_ms = self.timed_wall(w["e2e_pageable"], max(2, n_e2e // 2))
            e2e["pageable"] = dict(value=units * world / (p_ms * 1e-3), ms_per_step=p_ms,
                                   note="the same call on pageable caller buffers: the driver stages every copy")
        if with_ceiling:
            c_ms = self.pcie_ceiling(w["h2d"], w["d2h"])
            e2e["pcie_ceiling"] = dict(value=units * world / (c_ms * 1e-3), ms_per_step=c_ms,
                                       note="bare concurrent H2D + D2H of the same bytes from / to page-locked memory on "
                                            "all %d ranks at once, no kernels" % world)
            e2e["frac_of_ceiling"] = e2e["value"] / e2e["pcie_ceiling"]["value"]
        # clocks are sampled every 25 ms from the start of the device-timed steps to the end of the end-to-end steps; a
        # short run (the C2 steps take 1.2 ms each) ends before nvidia-smi has reported three times, so the same steps
        # keep running, untimed, until it has (the extension is stated in the JSON line)
        t_ext = time.perf_counter()
        while sampler.proc and sampler.n_since_mark() < 3 and time.perf_counter() - t_ext < 1.0:
            for i in range(4):
                step(i)
            torch.cuda.synchronize()
        ext_ms = (time.perf_counter() - t_ext) * 1e3
        clocks = sampler.section()
        if ext_ms > 1.0:
            clocks["untimed_load_extension_ms"] = round(ext_ms, 1)

        peaks, wl = self.peaks, w["name"]
        if w["bound"] == "hbm":
            achieved = w["algo_bytes"] / (dev_ms * 1e-3) / 1e9
            roof = dict(bound="hbm", achieved=achieved, peak=peaks["hbm"], unit="GB/s", frac=achieved / peaks["hbm"],
                        traffic=None, peak_source=peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs)")
            if wl == "gmm" and w.get("scorer") is not None:
                # the exact route is three kernels per step: live CUDA-event durations of each (rb_gmm_set_timing)
                sc, d_in, d_out = w["scorer"], w["keep"][1], w["keep"][2]
                sc.set_timing(True)
                parts = []
                for i in range(5):
                    sc.score_dev(d_in[i % self.R], units, d_out[i % self.R], None, self.sptr)
                    t3 = sc.timing()
                    if t3:
                        parts.append(t3)
                sc.set_timing(False)
                if parts:
                    k = np.mean(np.asarray(parts), axis=0)
                    roof["kernels_ms"] = {"gmm_split_features_kernel": float(k[0]),
                                          "gmm_tensor_kernel<EpiGmmScreen>": float(k[1]),
                                          "gmm_refine_kernel": float(k[2])}
                    roof["dominant_kernel"] = "gmm_refine_kernel"
                    roof["note"] = ("exact batch-float route = operand split + tcgen05 screening + exact refinement; "
                                    "achieved = algorithmic bytes of the step / step time; the refinement is bound by "
                                    "shared-memory wavefronts (per-lane row gathers), the screening by tcgen05.ld")
        else:
            achieved = NN_FLOP_PER_FRAME * units / (dev_ms * 1e-3) / 1e12
            roof = dict(bound="tensor", achieved=achieved, peak=peaks["bf16"], unit="TFLOP/s",
                        frac=achieved / peaks["bf16"], traffic=None,
                        peak_source=peaks["source"] + " (bf16_tflops_sustained)")
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            roof["traffic"] = json.load(open(tpath)).get(wl)  # dram bytes per launch (ncu, profiles/)
            roof["algorithmic_bytes"] = int(w["algo_bytes"])
        res = dict(value=value, unit=UNIT, ms_per_step=ms_per_step, steps=steps, warmup=warmup, dtype=w["dtype"],
                   config=dict(workload=w["workload"], wall_ms_per_step=wall_ms), e2e=e2e, gpu_launches=int(launches),
                   clocks=clocks, roofline=roof, dev_ms=dev_ms)
        if w.get("post"):
            res[w["post"][0]] = w["post"][1]()
        return res

    # other formulations of the C2 scorer, timed the same way on the same buffers and reported beside the headline:
    #   batch-tensor        RB_GMM_BATCH_TENSOR, 1e-4 relative instead of bit-identical
    #   batch-float-direct  the single direct-form kernel (RB_GMM_EXACT=0), what models
    #                       the screening does not cover run on; FP32-issue bound
    #   diagonal-maximum    RASR's default scorer (per-density covariance, best density), exact, same route
    def gmm_variants(self, w, steps, warmup):
        torch, stream = self.torch, self.stream
        from rasr_b200 import mm

        scorer, d_in, d_out, msd = w["keep"]
        T, R = w["units"], self.R
        out = {}
        for name in ("batch-tensor", "batch-float-direct", "diagonal-maximum", "batch-int", "SIMD-diagonal-maximum",
                     "preselection-batch-float", "preselection-batch-int"):
            if name == "batch-float-direct":
                os.environ["RB_GMM_EXACT"] = "0"
            try:
                vs = mm.GmmScorer(mm.MixtureSet.from_dict(msd), name.replace("-direct", ""), device=self.local_rank)
            finally:
                os.environ.pop("RB_GMM_EXACT", None)
            for i in range(warmup):
                vs.score_dev(d_in[i % R], T, d_out[i % R], None, self.sptr)
            self.barrier()
            tev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            tev[0].record(stream)
            for i in range(steps):
                vs.score_dev(d_in[i % R], T, d_out[i % R], None, self.sptr)
            tev[1].record(stream)
            self.barrier()
            tms = self.max_over_ranks(tev[0].elapsed_time(tev[1]) / steps)
            gbs = w["algo_bytes"] / (tms * 1e-3) / 1e9
            d = dict(value=T * self.world / (tms * 1e-3), unit=UNIT, ms_per_step=tms,
                     roofline=dict(bound="hbm", achieved=gbs, peak=self.peaks["hbm"], unit="GB/s", frac=gbs / self.peaks["hbm"]))
            if name == "batch-tensor":
                d.update(dtype="f16x3 split operands, f32 accumulate",
                         parity="<= 1e-4 relative to the reference scores (tests/test_gpu_gmm_tensor.py)")
            elif name.startswith("preselection"):
                d.update(dtype="f32" if name.endswith("float") else "u8 x u8 -> s32",
                         parity="bit-identical, clustering and back-off decisions included (tests/test_gpu_gmm_presel.py, "
                                "tests/test_gpu_zz_gmm_presel_int.py)",
                         note="the reference's density-preselection approximations (256 clusters, 32 selected), reproduced "
                              "as they are; the float variant runs through the refinement kernel of the exact route")
            elif name == "batch-int":
                d.update(dtype="u8 x u8 -> s32 (IMMA)", parity="bit-identical (tests/test_gpu_gmm_int.py)",
                         note="Mm::BatchIntFeatureScorer, the reference's quantised batch scorer")
            elif name == "SIMD-diagonal-maximum":
                d.update(dtype="u8 x u8 -> s32 (DP4A)",
                         parity="scores and best-density indices bit-identical (tests/test_gpu_gmm_simd.py)",
                         note="Mm::SimdGaussDiagonalMaximumFeatureScorer, the reference's quantised diagonal scorer "
                              "(crashes in the reference's own x86-64 cmake build, oracle/refbuild/Makefile)")
            elif name == "diagonal-maximum":
                d.update(dtype="f32", parity="scores and best-density indices bit-identical (tests/test_gpu_gmm_exact.py)",
                         note="Mm::GaussDiagonalMaximumFeatureScorer (per-density covariance) on the same model, through "
                              "the screening + refinement route")
            else:
                # FP32 peak: measured with scripts/micro/fp32_rate.cu (profiles/fp32_peak.json) -- a dependent
                # FADD2 -> FFMA2 stream, the instruction pair this kernel is made of, sustains 115 of the nominal 128
                # lane-FMAs per clock and SM
                ppath = os.path.join(ROOT, "profiles", "fp32_peak.json")
                lanes = json.load(open(ppath))["fadd2_ffma2_pairs"] if os.path.exists(ppath) else 128.0
                fp32_ceiling = 148 * lanes * 1965.0e6 / (2 * 39 * 4096)
                d.update(dtype="f32", parity="bit-identical (tests/test_gpu_gmm.py)",
                         fp32_alu=dict(ceiling_frames_per_s=fp32_ceiling, frac=(T / (tms * 1e-3)) / fp32_ceiling,
                                       lane_fma_per_clk_per_sm=lanes, nominal_lane_fma_per_clk_per_sm=128.0,
                                       peak_source="measured (profiles/fp32_peak.json, scripts/micro/fp32_rate.cu)"
                                       if os.path.exists(ppath) else "nominal",
                                       note="direct form: sub + fma per (frame, density, dim) on 148 SMs at 1965 MHz"))
            out[name] = d
            del vs
        return out

    # ---------------- north_star: "allgather of the score matrix only where a single decoder rank consumes all frames".
    # C3 shard per rank (125 utterances x 1000 frames x 256 scores = 128 MB), gathered into every rank's window (all-gather)
    # or into rank 0's (gather): the engine's own exchange over NVLink peer memory (push kernel; scorer epilogue storing
    # straight into the consumer's HBM) next to NCCL on the same buffers.  Device-timed, max over ranks.
    def gather_variants(self, steps, warmup):
        torch, dev, sptr, world, rank = self.torch, self.dev, self.sptr, self.world, self.rank
        from rasr_b200 import comm, flow, mm, pipeline, synth

        n_utt = 125
        samples_h, offs = synth.corpus(n_utt, n_samples=160240, seed0=3000 + 1000 * rank)
        fe = flow.FrontEnd(device=self.local_rank)
        scorer = mm.GmmScorer(mm.MixtureSet.from_dict(synth.mixture_set()), device=self.local_rank)
        T = int(fe.count_frames(offs)[-1])
        row_offsets = np.arange(world + 1, dtype=np.int64) * T
        with stdout_to_stderr():
            ex = comm.ScoreExchange(world, rank, self.local_rank, row_offsets, 256, comm.torch_exchange(self.dist), nccl=True)
        d_samples = torch.from_numpy(samples_h).to(dev)
        d_feats = torch.empty((T, 39), dtype=torch.float32, device=dev)
        d_local = torch.empty((T, 256), dtype=torch.float32, device=dev)
        own = ex.target(rank)
        shard_bytes = T * 256 * 4

        def timed(fn):
            for _ in range(warmup):
                fn()
            self.barrier()
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(self.stream)
            for _ in range(steps):
                fn()
            ev[1].record(self.stream)
            self.barrier()
            return self.max_over_ranks(ev[0].elapsed_time(ev[1]) / steps)

        def score(out):
            pipeline.score_utterances_dev(fe, scorer, d_samples, offs, d_feats, out, sptr)

        def v_compute():
            score(d_local)

        def v_p2p():
            ex.gather(d_local, -1, "p2p", sptr)
            ex.barrier(sptr)

        def v_nccl():
            ex.gather(own, -1, "nccl", sptr)

        def v_step_p2p():
            score(own)
            ex.gather(own, -1, "p2p", sptr)
            ex.barrier(sptr)

        def v_step_nccl():
            score(own)
            ex.gather(own, -1, "nccl", sptr)

        # the all-gather pipelined behind the scorer: the shard is scored in 5 slabs of whole utterances into this
        # rank's own window, every finished slab is pushed to the peers from a second stream while the next one is
        # being scored; only the last slab's push is exposed
        side = torch.cuda.Stream(device=dev)
        fo = fe.count_frames(offs)
        cuts = [n_utt * k // 5 for k in range(6)]
        sample_pos = [int(offs[c]) for c in cuts]
        slab_ev = [torch.cuda.Event() for _ in range(5)]
        done_ev = torch.cuda.Event()
        row0 = int(row_offsets[rank])

        def v_step_p2p_slabs():
            for k in range(5):
                u0, u1 = cuts[k], cuts[k + 1]
                f0, f1 = int(fo[u0]), int(fo[u1])
                pipeline.score_utterances_dev(fe, scorer, d_samples[sample_pos[k]:], offs[u0:u1 + 1] - offs[u0],
                                              d_feats[f0:], own + f0 * 1024, sptr)
                slab_ev[k].record(self.stream)
                side.wait_event(slab_ev[k])
                ex.push_rows(own + f0 * 1024, row0 + f0, f1 - f0, -1, side.cuda_stream)
            done_ev.record(side)
            self.stream.wait_event(done_ev)
            ex.barrier(sptr)

        def v_step_fused_root():
            score(ex.target(0))
            ex.barrier(sptr)

        fan = ex.targets()

        def v_step_fused_allgather():
            pipeline.score_utterances_fanout_dev(fe, scorer, d_samples, offs, d_feats, fan, sptr)
            ex.barrier(sptr)

        def v_step_push_root():
            score(d_local)
            ex.gather(d_local, 0, "p2p", sptr)
            ex.barrier(sptr)

        out = dict(shard_bytes=shard_bytes, gathered_bytes=shard_bytes * world, frames_per_rank=T,
                   nccl_version=comm.nccl_version(),
                   note="ms per step, CUDA events on the launching stream, max over ranks; GB/s = bytes arriving in ONE "
                        "rank's window per second ((world-1) shards for the all-gathers and for the root of the gathers)")
        t_compute = timed(v_compute)
        out["compute_only_ms"] = t_compute
        arriving = shard_bytes * (world - 1)
        for name, fn, with_compute in (("allgather_p2p", v_p2p, False), ("allgather_nccl", v_nccl, False),
                                       ("step_allgather_p2p", v_step_p2p, True), ("step_allgather_nccl", v_step_nccl, True),
                                       ("step_allgather_p2p_5_slabs_overlapped", v_step_p2p_slabs, True),
                                       ("step_allgather_fused_epilogue", v_step_fused_allgather, True),
                                       ("step_gather_root0_fused_epilogue", v_step_fused_root, True),
                                       ("step_gather_root0_push", v_step_push_root, True)):
            ms = timed(fn)
            d = dict(ms=ms)
            if with_compute:
                d["frames_per_s"] = T * world / (ms * 1e-3)
                d["exposed_exchange_ms"] = ms - t_compute
            else:
                d["gbs_in_per_rank"] = arriving / (ms * 1e-3) / 1e9
            out[name] = d
        ex.close()
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + ALL_WORKLOADS,
                    help="all (default): the C2 headline plus C3 / C4 / C5 under 'workloads'; else that workload alone")
    ap.add_argument("--frames", type=int, default=0, help="override the frame count of the workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", action="store_true",
                    help="N > 1: add the score exchange (own NVLink push / fused epilogue vs NCCL) under 'gather'")
    ap.add_argument("--cpu-seconds", type=float, default=0.0, help="reference arm: CPU seconds per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    sampler = ClockSampler(local_rank)
    sampler.start()
    b = Bench(args, torch, dist, rank, local_rank, world, sampler)

    head = "gmm" if args.workload == "all" else args.workload
    w = b.setup(head, args.frames)
    m = b.measure(w, args.steps, args.warmup)
    variants = b.gmm_variants(w, args.steps, args.warmup) if head == "gmm" else None
    line = dict(metric=METRIC, value=m["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=m["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype=m["dtype"],
                data="synthetic",
                config=dict(workload=w["workload"], sharding="independent frame/utterance shards per GPU, no collective",
                            l2="rotating %d input/output buffer sets (> 126 MB L2) between timed steps" % b.R,
                            wall_ms_per_step=m["config"]["wall_ms_per_step"], host_numa_node=numa),
                e2e=m["e2e"], gpu_launches=m["gpu_launches"], clocks=m["clocks"], roofline=m["roofline"])
    if variants:
        line["variants"] = variants
    if "search_kernel_alone" in m:
        line["search_kernel_alone"] = m["search_kernel_alone"]
    del w

    # ---------------- the other BASELINE configs (C3, C4, C5) in the same driver-run record
    if args.workload == "all":
        extra = {}
        for key, wl, frames in (("C3", "pipeline", 0), ("C4", "nn", 32768), ("C5", "pipeline-search", 0)):
            torch.cuda.empty_cache()
            we = b.setup(wl, frames)
            me = b.measure(we, max(3, min(args.steps, 5)), 3)
            me.pop("dev_ms")
            me["config"]["name"] = wl
            extra[key] = me
            del we
        line["workloads"] = extra
    if world > 1 and (args.gather or args.workload == "all"):
        torch.cuda.empty_cache()
        # the scoring numbers above stand whatever happens here: if the exchange does not come back (a peer that cannot
        # be mapped, a collective that never completes) every rank gives up after 90 s, rank 0 prints the line without it
        finished = threading.Event()

        def give_up():
            if not finished.is_set():
                line["gather"] = dict(unavailable="the score exchange did not finish within 90 s")
                if rank == 0:
                    print(json.dumps(line), flush=True)
                os._exit(0)

        watchdog = threading.Timer(90.0, give_up)
        watchdog.daemon = True
        watchdog.start()
        try:
            line["gather"] = b.gather_variants(max(3, min(args.steps, 10)), 3)
        except Exception as e:  # noqa: BLE001 -- e.g. no peer access on this box: reported, the scoring numbers stand
            line["gather"] = dict(unavailable=str(e)[:300])
        finished.set()
        watchdog.cancel()
    sampler.stop()

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and head in ("gmm", "frontend", "nn"):
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline uses every host core, not one NUMA node
            line["cpu_baseline"] = {"gmm": cpu_baseline_same_job, "frontend": cpu_baseline_frontend,
                                    "nn": cpu_baseline_nn}[head]()
        print(json.dumps(line), flush=True)
    if world > 1:
        if isinstance(line.get("gather"), dict) and "unavailable" in line["gather"]:
            os._exit(0)  # the device context may be gone (trapped barrier): do not wait for a collective teardown
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
