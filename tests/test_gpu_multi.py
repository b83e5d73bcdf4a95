"""Score exchange between two GPUs of one box (rb_comm_*): one process per GPU, windows mapped over CUDA
IPC, push kernel / fused scorer epilogue / NCCL.  Needs two devices (gpurun --gpus 2); skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest

from rasr_b200 import capi

pytestmark = pytest.mark.gpu

WORLD = 2
ROWS = [700, 1300]   # unequal shards
M = 256


def _two_devices():
    try:
        return capi.device_count() >= WORLD
    except Exception:  # noqa: BLE001
        return False


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, q):
    try:
        import torch
        import torch.distributed as dist

        from oracle import pyoracle as o
        from rasr_b200 import comm, mm, synth

        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", rank=rank, world_size=WORLD)
        dev = torch.device("cuda", rank)
        offs = np.concatenate([[0], np.cumsum(ROWS)]).astype(np.int64)
        ex = comm.ScoreExchange(WORLD, rank, rank, offs, M, comm.torch_exchange(dist), nccl=True)
        win = ex.window(torch)
        stream = torch.cuda.Stream(device=dev)
        sp = stream.cuda_stream
        res = {}

        def expected(r):
            g = torch.Generator().manual_seed(100 + r)
            return torch.rand((ROWS[r], M), generator=g)

        want = torch.cat([expected(r) for r in range(WORLD)])
        local = expected(rank).to(dev)

        # 1. all-gather, push kernel over peer memory
        win.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        ex.gather(local, root=-1, transport="p2p", stream=sp)
        ex.barrier(sp)
        stream.synchronize()
        res["allgather_p2p"] = bool(torch.equal(win.cpu(), want))
        dist.barrier()

        # 2. gather to rank 1 only
        win.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        ex.gather(local, root=1, transport="p2p", stream=sp)
        ex.barrier(sp)
        stream.synchronize()
        got = win.cpu()
        res["gather_root1"] = bool(torch.equal(got, want)) if rank == 1 else bool((got == 0).all())
        dist.barrier()

        # 3. NCCL, unequal shards (grouped in-place broadcasts)
        win.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        ex.gather(local, root=-1, transport="nccl", stream=sp)
        stream.synchronize()
        res["allgather_nccl"] = bool(torch.equal(win.cpu(), want))
        dist.barrier()

        # 4. fused: the GMM scorer's epilogue stores straight into rank 0's window over NVLink
        msd = synth.mixture_set()
        feats = synth.features(int(offs[-1]), 39, seed=77)
        scorer = mm.GmmScorer(mm.MixtureSet.from_dict(msd), device=rank)
        mine = torch.from_numpy(feats[offs[rank]:offs[rank + 1]]).to(dev)
        win.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        scorer.score_dev(mine, ROWS[rank], ex.target(0), None, sp)
        ex.barrier(sp)
        stream.synchronize()
        if rank == 0:
            ref = o.gmm_batch_float(o.MixtureSet(**msd), feats, threads=4)
            res["fused_scorer"] = bool(np.array_equal(win.cpu().numpy(), ref))
        else:
            res["fused_scorer"] = bool((win.cpu() == 0).all())
        dist.barrier()

        # 5. the local copy is skipped when the scorer already wrote into its own window slice
        win.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        scorer.score_dev(mine, ROWS[rank], ex.target(rank), None, sp)
        ex.gather(ex.target(rank), root=-1, transport="p2p", stream=sp)
        ex.barrier(sp)
        stream.synchronize()
        ref = o.gmm_batch_float(o.MixtureSet(**msd), feats, threads=4)
        res["inplace_allgather"] = bool(np.array_equal(win.cpu().numpy(), ref))
        dist.barrier()

        # 5a. the all-gather fused into the scorer: its last kernel stores every row into both windows
        win.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        big = synth.features(6000, 39, seed=78)          # large enough for the screening + refinement route
        offs2 = np.array([0, 2500, 6000], np.int64)
        ex2 = comm.ScoreExchange(WORLD, rank, rank, offs2, M, comm.torch_exchange(dist))
        mine2 = torch.from_numpy(big[offs2[rank]:offs2[rank + 1]]).to(dev)
        scorer.score_fanout_dev(mine2, int(offs2[rank + 1] - offs2[rank]), ex2.targets(), sp)
        ex2.barrier(sp)
        stream.synchronize()
        ref2 = o.gmm_batch_float(o.MixtureSet(**msd), big, threads=4)
        res["fused_allgather"] = bool(np.array_equal(ex2.window(torch).cpu().numpy(), ref2))
        # ... and a mode without a fused store path (copies of the finished matrix)
        tscorer = mm.GmmScorer(mm.MixtureSet.from_dict(msd), "batch-int", device=rank)
        single = torch.empty((int(offs2[-1]), M), dtype=torch.float32, device=dev)
        tscorer.score_dev(torch.from_numpy(big).to(dev), int(offs2[-1]), single, None, sp)
        dist.barrier()
        tscorer.score_fanout_dev(mine2, int(offs2[rank + 1] - offs2[rank]), ex2.targets(), sp)
        ex2.barrier(sp)
        stream.synchronize()
        res["fanout_by_copy"] = bool(torch.equal(ex2.window(torch), single))
        dist.barrier()
        ex2.close()

        # 5c. the Nn scorer's TMA-store epilogue writing its shard straight into rank 1's window
        from rasr_b200 import nn as rnn
        net = synth.network(dims=(429, 256, 512), seed=1)
        xs = synth.features(300 + 200, 429, seed=2, scale=1.0)
        offs3 = np.array([0, 300, 500], np.int64)
        ex3 = comm.ScoreExchange(WORLD, rank, rank, offs3, 512, comm.torch_exchange(dist))
        sc3 = rnn.NnScorer(net["dims"], net["acts"], net["weights"], net["biases"], net["log_prior"], 1.0, "bf16", device=rank)
        whole = torch.empty((500, 512), dtype=torch.float32, device=dev)
        sc3.score_dev(torch.from_numpy(xs).to(dev), 500, whole, sp)
        ex3.window(torch).zero_()
        torch.cuda.synchronize()
        dist.barrier()
        part = torch.from_numpy(xs[offs3[rank]:offs3[rank + 1]]).to(dev)
        sc3.score_dev(part, int(offs3[rank + 1] - offs3[rank]), ex3.target(1), sp)
        ex3.barrier(sp)
        stream.synchronize()
        if rank == 1:
            res["nn_fused_gather"] = bool(torch.equal(ex3.window(torch), whole))
        else:
            res["nn_fused_gather"] = bool((ex3.window(torch) == 0).all())
        dist.barrier()
        ex3.close()

        # 5b. row-range pushes (the slab-pipelined exchange): two halves of the shard, the second to rank 0 only
        win.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        half = ROWS[rank] // 2
        r0 = int(offs[rank])
        ex.push_rows(local, r0, half, -1, sp)
        ex.push_rows(local[half:], r0 + half, ROWS[rank] - half, 0, sp)
        ex.barrier(sp)
        stream.synchronize()
        exp = want.clone()
        if rank != 0:
            for r in range(WORLD):
                if r != rank:
                    exp[int(offs[r]) + ROWS[r] // 2:int(offs[r + 1])] = 0
            exp[r0 + half:int(offs[rank + 1])] = 0
        res["push_rows"] = bool(torch.equal(win.cpu(), exp))
        dist.barrier()

        # 6. argument checks
        try:
            bad = offs.copy()
            bad[-1] += 1
            capi.check(capi.lib().rb_comm_gather_scores_dev(ex.handle, capi.ptr(local), capi.ptr(bad), M, -1, 0, None))
            res["rejects_overflow"] = False
        except capi.RasrB200Error as e:
            res["rejects_overflow"] = e.status == -1
        res["nccl_version"] = comm.nccl_version()
        ex.close()
        dist.destroy_process_group()
        q.put((rank, res))
    except Exception as e:  # noqa: BLE001
        import traceback

        q.put((rank, dict(error="%s\n%s" % (e, traceback.format_exc()))))


@pytest.mark.skipif(not _two_devices(), reason="needs two sm_100 devices on one box")
def test_score_exchange_two_ranks(oracle, diag):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank in range(WORLD):
        r = res[rank]
        assert "error" not in r, r.get("error")
        diag("score_exchange_rank%d" % rank, **r)
        for key in ("allgather_p2p", "gather_root1", "allgather_nccl", "fused_scorer", "inplace_allgather", "fused_allgather", "fanout_by_copy", "nn_fused_gather",
                    "push_rows", "rejects_overflow"):
            assert r[key] is True, (rank, key)
        assert r["nccl_version"] >= 20000


def test_single_rank_window_is_local():
    """world = 1: no IPC, the gather degenerates to a device copy"""
    import torch

    from rasr_b200 import comm

    offs = np.array([0, 100], np.int64)
    ex = comm.ScoreExchange(1, 0, 0, offs, 8)
    x = torch.rand((100, 8), device="cuda:0")
    s = torch.cuda.Stream()
    ex.gather(x, transport="p2p", stream=s.cuda_stream)
    ex.barrier(s.cuda_stream)
    s.synchronize()
    assert torch.equal(ex.window(torch), x)
    assert ex.target(0) == ex.window_ptr
