#!/usr/bin/env python
"""bench.py -- acoustic frames scored/sec on B200 (BASELINE.json metric), one JSON line per run.

Default workload (N=1) = BASELINE config C2: GMM FeatureScorer, 39-dim MFCC, 4096 diagonal Gaussians in
256 mixtures, 100 000 frames.  A "step" = one dense scoring pass of the hot path over that batch.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

  value   frames/s with inputs resident in HBM (device pointers, rb_gmm_score_dev)
  e2e     frames/s through the host-buffer C-ABI call rb_gmm_score (pinned host memory, H2D + D2H inside)
  roofline  dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the oracle restatement of Mm::BatchFloatFeatureScorer on the box's host cores

Other workloads for profiling (not the contract line): --workload frontend | pipeline | nn | gmm-diag
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
ALGO_BYTES_PER_FRAME = {"gmm": 39 * 4 + 256 * 4, "frontend": 160 * 4 + 39 * 4, "pipeline": 160 * 4 + 256 * 4,
                        "nn": 429 * 4 + 12000 * 4}
NN_FLOP_PER_FRAME = 2 * (429 * 2048 + 5 * 2048 * 2048 + 2048 * 12000)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))),
                    source="measured")
    return dict(hbm=6650.0, bf16=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

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
        """rows from here on count (start of the timed region)"""
        self.first = len(self.rows)

    def n_since_mark(self):
        return len(self.rows) - getattr(self, "first", 0)

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[getattr(self, "first", 0):]:
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


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle restatement of the reference's CPU scorer
# --------------------------------------------------------------------------------------------

def cpu_baseline_gmm(seconds=12.0):
    """Oracle port of Mm::BatchFloatFeatureScorer on all host cores over a bounded sample of C2."""
    from oracle import pyoracle as o
    from rasr_b200 import synth

    o.build(ref=False)
    cores = os.cpu_count() or 1
    ms = o.MixtureSet(**synth.mixture_set())
    probe = synth.features(64 * cores, 39)
    o.gmm_batch_float(ms, probe[:cores * 8], threads=cores)  # warm-up
    t = time.perf_counter()
    o.gmm_batch_float(ms, probe, threads=cores)
    rate = probe.shape[0] / (time.perf_counter() - t)
    n = int(min(C2["frames"], max(1024, rate * seconds)))
    f = synth.features(C2["frames"], 39)[:n]
    t = time.perf_counter()
    o.gmm_batch_float(ms, f, threads=cores)
    dt = time.perf_counter() - t
    return dict(value=n / dt, unit=UNIT, cores=cores, kind="port",
                sample="first %d of the 100000 C2 frames, all 256 mixtures, %d threads over frame ranges" % (n, cores))


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as o
    from rasr_b200 import synth

    o.build(ref=False)
    cores = os.cpu_count() or 1
    ms = o.MixtureSet(**synth.mixture_set())
    f_all = synth.features(C2["frames"], 39)
    probe = f_all[:64 * cores]
    o.gmm_batch_float(ms, probe, threads=cores)
    t = time.perf_counter()
    o.gmm_batch_float(ms, probe, threads=cores)
    rate = probe.shape[0] / (time.perf_counter() - t)
    budget = 150.0 / max(1, args.steps + args.warmup)  # the whole run ends within a few minutes
    n = int(min(C2["frames"], max(1024, rate * min(budget, 20.0))))
    f = f_all[:n]
    for _ in range(args.warmup):
        o.gmm_batch_float(ms, f, threads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        o.gmm_batch_float(ms, f, threads=cores)
    dt = (time.perf_counter() - t) / args.steps
    value = n / dt
    sample = "first %d of the 100000 C2 frames per step, %d threads" % (n, cores)
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload="C2: GMM FeatureScorer, 39-dim, 4096 densities / 256 mixtures, 100k frames",
                            note="reference cannot be built here (no libxml2/boost/BLAS); this is the oracle port of "
                                 "Mm::BatchFloatFeatureScorer (SSE lane order) on the host cores"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gmm", choices=["gmm", "gmm-diag", "gmm-tensor", "gmm-int", "gmm-presel", "gmm-presel-int", "frontend", "pipeline", "pipeline-nn", "pipeline-search", "nn"])
    ap.add_argument("--frames", type=int, default=0, help="override the frame count of the workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from rasr_b200 import capi, flow, mm, nn, pipeline, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # a non-default stream: the C ABI treats a NULL stream as "the handle's own stream", and the
    # events below must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    peaks = measured_peaks()
    wl = args.workload
    R = 4  # rotating buffer sets so that no step finds its data in L2

    # ---------------- workload set-up: every rank owns its own shard (weak scaling, no collective)
    e2e_fn = None
    if wl in ("gmm", "gmm-diag", "gmm-tensor", "gmm-int", "gmm-presel", "gmm-presel-int"):
        T = args.frames or C2["frames"]
        msd = synth.mixture_set()
        mode = {"gmm": "batch-float", "gmm-diag": "diagonal-maximum", "gmm-tensor": "batch-tensor",
                "gmm-int": "batch-int", "gmm-presel": "preselection-batch-float",
                "gmm-presel-int": "preselection-batch-int"}[wl]
        scorer = mm.GmmScorer(mm.MixtureSet.from_dict(msd), mode, device=local_rank)
        feats_h = synth.features(T, 39, seed=2024 + rank)
        d_in = [torch.from_numpy(feats_h).to(dev) for _ in range(R)]
        d_out = [torch.empty((T, 256), dtype=torch.float32, device=dev) for _ in range(R)]

        def step(i):
            scorer.score_dev(d_in[i % R], T, d_out[i % R], None, sptr)

        h_in = torch.from_numpy(feats_h).pin_memory()
        h_out = torch.empty((T, 256), dtype=torch.float32).pin_memory()

        def e2e_fn():
            scorer.score(h_in, out=h_out)

        h2d, d2h = T * 39 * 4, T * 256 * 4
        units = T
        workload = "C2: GMM FeatureScorer (%s), 39-dim, 4096 densities / 256 mixtures, %d frames per GPU" % (mode, T)
        algo_bytes = ALGO_BYTES_PER_FRAME["gmm"] * T
        bound, dtype = "hbm", ("u8/s32" if wl in ("gmm-int", "gmm-presel-int") else "f32")
    elif wl in ("frontend", "pipeline"):
        n_utt = 125
        if args.frames:
            n_utt = max(1, args.frames // 1000)
        samples_h, offs = synth.corpus(n_utt, n_samples=160240, seed0=3000 + 1000 * rank)
        fe = flow.FrontEnd(device=local_rank)
        fo = fe.count_frames(offs)
        T = int(fo[-1])
        d_samples = [torch.from_numpy(samples_h).to(dev) for _ in range(R)]
        d_feats = [torch.empty((T, 39), dtype=torch.float32, device=dev) for _ in range(R)]
        h_samples = torch.from_numpy(samples_h).pin_memory()
        if wl == "frontend":
            def step(i):
                fe.process_dev(d_samples[i % R], offs, d_feats[i % R], sptr)

            h_feats = torch.empty((T, 39), dtype=torch.float32).pin_memory()

            # end to end the audio arrives as 16-bit PCM (what the audio nodes deliver): 2 bytes per sample over PCIe
            h_pcm = torch.from_numpy(samples_h.astype(np.int16)).pin_memory()

            def e2e_fn():
                fe.process_s16(h_pcm, offs, timestamps=False, out=h_feats)

            h2d, d2h = samples_h.size * 2, T * 39 * 4
        else:
            scorer = mm.GmmScorer(mm.MixtureSet.from_dict(synth.mixture_set()), device=local_rank)
            d_scores = [torch.empty((T, 256), dtype=torch.float32, device=dev) for _ in range(R)]

            def step(i):
                pipeline.score_utterances_dev(fe, scorer, d_samples[i % R], offs, d_feats[i % R], d_scores[i % R], sptr)

            h_scores = torch.empty((T, 256), dtype=torch.float32).pin_memory()

            h_pcm = torch.from_numpy(samples_h.astype(np.int16)).pin_memory()  # 16-bit PCM over PCIe

            def e2e_fn():
                pipeline.score_utterances(fe, scorer, h_pcm, offs, out=h_scores, pcm_channels=1)

            h2d, d2h = samples_h.size * 2, T * 256 * 4
        units = T
        workload = "C3 shard: %s on %d utterances x 1000 frames per GPU (%d frames)" % (wl, n_utt, T)
        algo_bytes = ALGO_BYTES_PER_FRAME[wl] * T
        bound, dtype = "hbm", "f32"
    elif wl == "pipeline-search":
        # C5: audio -> MFCC -> GMM scores -> Search::LinearSearch (1000-word synthetic lexicon), scores never leave HBM
        from rasr_b200 import search
        n_utt = max(1, (args.frames or 125000) // 1000)
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

        h_pcm = torch.from_numpy(samples_h.astype(np.int16)).pin_memory()

        def e2e_fn():
            # 16-bit PCM in, word sequences out (rb_pipeline_search): slab-pipelined H2D of the audio, conversion, the
            # three stages, tracebacks to the host
            return pipeline.search_utterances(fe, scorer, ls, h_pcm, offs, pcm_channels=1)

        h2d, d2h = samples_h.size * 2, T * 20
        units = T
        workload = "C5: MFCC -> GMM scores -> LinearSearch (1000 words), %d utterances x 1000 frames per GPU" % n_utt
        algo_bytes = ALGO_BYTES_PER_FRAME["pipeline"] * T
        bound, dtype = "hbm", "f32"
    elif wl == "pipeline-nn":
        # C4 fed from audio: MFCC -> segment CMVN -> 11-frame window -> 6 x 2048 -> 12000 senone scores
        from rasr_b200 import postproc
        n_utt = max(1, (args.frames or 37000) // 1000)
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

        def step(i):
            pipeline.nn_score_utterances_dev(fe, pp, sc, d_samples[i % 2], offs, d_feats, d_post, d_out[i % 2], sptr)

        h_samples = torch.from_numpy(samples_h).pin_memory()
        h_out = torch.empty((T, 12000), dtype=torch.float32).pin_memory()

        def e2e_fn():
            pipeline.nn_score_utterances(fe, pp, sc, h_samples, offs, out=h_out)

        h2d, d2h = samples_h.size * 4, T * 12000 * 4
        units = T
        workload = "C4 from audio: MFCC -> CMVN -> 11-frame window -> Nn 429 -> 6x2048 -> 12000, %d utterances x 1000 frames per GPU" % n_utt
        algo_bytes = (160 * 4 + 12000 * 4) * T
        bound, dtype = "tensor", "bf16"
    else:  # nn
        T = args.frames or 65536
        net = synth.network()
        sc = nn.NnScorer(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 1.0, "bf16",
                         device=local_rank)
        x_h = synth.features(T, 429, seed=4096 + rank, scale=1.0)
        d_in = [torch.from_numpy(x_h).to(dev) for _ in range(2)]
        d_out = [torch.empty((T, 12000), dtype=torch.float32, device=dev) for _ in range(2)]

        def step(i):
            sc.score_dev(d_in[i % 2], T, d_out[i % 2], sptr)

        h_in = torch.from_numpy(x_h).pin_memory()
        h_out = torch.empty((T, 12000), dtype=torch.float32).pin_memory()

        def e2e_fn():
            sc.score(h_in, out=h_out)

        h2d, d2h = T * 429 * 4, T * 12000 * 4
        units = T
        workload = "C4: Nn 429 -> 6x2048 -> 12000 senones, bf16 tcgen05, %d frames per GPU" % T
        algo_bytes = ALGO_BYTES_PER_FRAME["nn"] * T
        bound, dtype = "tensor", "bf16"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler.wait_ready()
    for i in range(min(args.warmup, 2)):  # the GPU idled while nvidia-smi started: back to steady state
        step(i)
    barrier()
    sampler.mark()
    launches0 = capi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev_all = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    t_wall = time.perf_counter()
    ev_all[0].record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        step(i)
        ev[i][1].record(stream)
    ev_all[1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = capi.launch_count() - launches0
    # dominant-kernel time: per-step event pairs on the launching stream (one step = one scoring launch)
    dev_ms = float(np.sum([a.elapsed_time(b) for a, b in ev])) / args.steps
    # step time: one event pair around EXACTLY K steps (inter-step gaps included), max over ranks
    ms = torch.tensor([t_wall * 1e3 / args.steps], dtype=torch.float64, device=dev)
    kms = torch.tensor([ev_all[0].elapsed_time(ev_all[1]) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_per_step = float(kms.item())
    value = units * world / (ms_per_step * 1e-3)

    # ---------------- end to end through the host-buffer C-ABI call
    for _ in range(2):
        e2e_fn()
    barrier()
    n_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_fn()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / n_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = units * world / (float(e2e_ms.item()) * 1e-3)
    # clocks are sampled every 25 ms from the start of the device-timed steps to the end of the end-to-end steps; a
    # short run (the C2 steps take 1.2 ms each) ends before nvidia-smi has reported three times, so the same steps keep
    # running, untimed, until it has (the extension is stated in the JSON line)
    t_ext = time.perf_counter()
    while sampler.proc and sampler.n_since_mark() < 3 and time.perf_counter() - t_ext < 1.0:
        for i in range(4):
            step(i)
        torch.cuda.synchronize()
    ext_ms = (time.perf_counter() - t_ext) * 1e3
    clocks = sampler.stop()
    if ext_ms > 1.0:
        clocks["untimed_load_extension_ms"] = round(ext_ms, 1)

    # ---------------- the tensor-core formulation of the same scorer (RB_GMM_BATCH_TENSOR, 1e-4 relative instead of
    # bit-identical), timed the same way and reported beside the headline as "variants"
    variants = None
    if wl == "gmm":
        tscorer = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-tensor", device=local_rank)
        for i in range(args.warmup):
            tscorer.score_dev(d_in[i % R], T, d_out[i % R], None, sptr)
        barrier()
        tev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        tev[0].record(stream)
        for i in range(args.steps):
            tscorer.score_dev(d_in[i % R], T, d_out[i % R], None, sptr)
        tev[1].record(stream)
        barrier()
        tms = torch.tensor([tev[0].elapsed_time(tev[1]) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        tms = float(tms.item())
        variants = {"batch-tensor": dict(
            value=units * world / (tms * 1e-3), unit=UNIT, ms_per_step=tms, dtype="f16x3 split operands, f32 accumulate",
            parity="<= 1e-4 relative to the reference scores (tests/test_gpu_gmm_tensor.py)",
            roofline=dict(bound="hbm", achieved=algo_bytes / (tms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s",
                          frac=algo_bytes / (tms * 1e-3) / 1e9 / peaks["hbm"]))}

    if rank == 0:
        if bound == "hbm":
            achieved = algo_bytes / (dev_ms * 1e-3) / 1e9
            roof = dict(bound="hbm", achieved=achieved, peak=peaks["hbm"], unit="GB/s", frac=achieved / peaks["hbm"],
                        traffic=None, peak_source=peaks["source"] + " (MEASURED_PEAKS.json hbm_gbs)")
            if wl.startswith("gmm"):
                sm_mhz = clocks.get("sm_mhz") or 1965.0
                # CUDA-core ceiling of the reference-order arithmetic: sub + fma per (frame, density, dim)
                fp32_ceiling = 148 * 128 * sm_mhz * 1e6 / (2 * 39 * 4096)
                roof["fp32_alu"] = dict(ceiling_frames_per_s=fp32_ceiling, frac=(units / (dev_ms * 1e-3)) / fp32_ceiling,
                                        note="direct-form GMM is FP32-issue bound: 2 FMA-pipe ops per "
                                             "(frame,density,dim) at the sampled SM clock")
        else:
            achieved = NN_FLOP_PER_FRAME * units / (dev_ms * 1e-3) / 1e12
            roof = dict(bound="tensor", achieved=achieved, peak=peaks["bf16"], unit="TFLOP/s",
                        frac=achieved / peaks["bf16"], traffic=None,
                        peak_source=peaks["source"] + " (bf16_tflops_sustained)")
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            roof["traffic"] = json.load(open(tpath)).get(wl)  # dram bytes per launch (ncu, profiles/)
            roof["algorithmic_bytes"] = int(algo_bytes)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=dtype,
                    data="synthetic",
                    config=dict(workload=workload, sharding="independent frame/utterance shards per GPU, no collective",
                                l2="rotating %d input/output buffer sets (> 126 MB L2) between timed steps" % R,
                                wall_ms_per_step=float(ms.item())),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                             ms_per_step=float(e2e_ms.item()), api="host-buffer C-ABI call, pinned host memory"),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof)
        line["config"]["host_numa_node"] = numa
        if variants:
            line["variants"] = variants
        if world == 1 and not args.no_cpu_baseline and wl in ("gmm", "frontend", "nn"):
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline uses every host core, not one NUMA node
            line["cpu_baseline"] = {"gmm": cpu_baseline_gmm, "frontend": cpu_baseline_frontend,
                                    "nn": cpu_baseline_nn}[wl]()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
