"""Fused audio -> emission scores (BASELINE config C3) and utterance sharding across ranks.

The reference's only parallel mode is N independent processes over corpus partitions
(src/Bliss/CorpusDescription.cc:173-180,482-491); `partition()` reproduces that with a
longest-first greedy bin-pack so ranks finish together.  There is no collective in the compute
path; `gather_scores()` is the optional all-gather for a single decoder rank.
"""
import numpy as np

from . import capi


def score_utterances(frontend, gmm, samples, offsets, want_feats=False, out=None, pcm_channels=0, track=0):
    """Host buffers in (numpy, or pinned torch CPU tensors), scores [total_frames x n_mixtures] out (`out` if
    given); features stay on the device.  pcm_channels > 0: `samples` is interleaved 16-bit PCM."""
    if isinstance(samples, np.ndarray) or not hasattr(samples, "data_ptr"):
        samples = np.ascontiguousarray(samples, np.int16 if pcm_channels else np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    fo = frontend.count_frames(offsets)
    T = int(fo[-1])
    scores = out if out is not None else np.zeros((T, gmm.n_mixtures), np.float32)
    feats = np.zeros((T, frontend.feat_dim), np.float32) if want_feats else None
    if pcm_channels:
        capi.check(capi.lib().rb_pipeline_score_s16(frontend.handle, gmm.handle, capi.ptr(samples), int(pcm_channels),
                                                    int(track), capi.ptr(offsets), offsets.size - 1, capi.ptr(scores),
                                                    capi.ptr(feats)))
    else:
        capi.check(capi.lib().rb_pipeline_score(frontend.handle, gmm.handle, capi.ptr(samples), capi.ptr(offsets),
                                                offsets.size - 1, capi.ptr(scores), capi.ptr(feats)))
    return (scores, feats, fo) if want_feats else (scores, fo)


def search_utterances(frontend, gmm, searcher, samples, offsets, pcm_channels=0, track=0, want_result=True):
    """Config C5 in one call: host audio in (f32, or interleaved 16-bit PCM with pcm_channels > 0), one traceback dict
    per segment out; features and scores never leave the device."""
    if isinstance(samples, np.ndarray) or not hasattr(samples, "data_ptr"):
        samples = np.ascontiguousarray(samples, np.int16 if pcm_channels else np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    capi.check(capi.lib().rb_pipeline_search(frontend.handle, gmm.handle, searcher.handle, capi.ptr(samples),
                                             int(pcm_channels), int(track), capi.ptr(offsets), offsets.size - 1))
    searcher._fo = frontend.count_frames(offsets)
    return searcher._result() if want_result else None


def score_utterances_dev(frontend, gmm, d_samples, offsets, d_feats, d_scores, stream=None):
    """Device buffers; only enqueues.  stream=None means the FRONT-END handle's own stream (every handle owns one
    non-blocking stream): pass one stream to all `*_dev` calls that feed each other, or synchronise between them."""
    offsets = np.ascontiguousarray(offsets, np.int64)
    capi.check(capi.lib().rb_pipeline_score_dev(frontend.handle, gmm.handle, capi.ptr(d_samples), capi.ptr(offsets),
                                                offsets.size - 1, capi.ptr(d_feats), capi.ptr(d_scores),
                                                capi.ptr(stream)))


def score_utterances_fanout_dev(frontend, gmm, d_samples, offsets, d_feats, dsts, stream=None):
    """score_utterances_dev with the scores stored into every destination of `dsts` (device addresses: this rank's rows
    in the windows of the GPUs that need them, rasr_b200.comm.ScoreExchange.targets()) -- the exchange fused into the
    scorer's last kernel"""
    offsets = np.ascontiguousarray(offsets, np.int64)
    import ctypes as C

    arr = (C.c_void_p * len(dsts))(*[capi.ptr(d).value for d in dsts])
    capi.check(capi.lib().rb_pipeline_score_fanout_dev(frontend.handle, gmm.handle, capi.ptr(d_samples), capi.ptr(offsets),
                                                       offsets.size - 1, capi.ptr(d_feats), len(dsts), arr,
                                                       capi.ptr(stream)))


def nn_score_utterances(frontend, postproc, nn, samples, offsets, out=None):
    """audio -> MFCC -> post-processing (may be None) -> Nn scores; host buffers in, [total_frames x n_emissions] out
    (n_emissions = the class count of an active class mapping, else the number of network outputs)."""
    if isinstance(samples, np.ndarray) or not hasattr(samples, "data_ptr"):
        samples = np.ascontiguousarray(samples, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    fo = frontend.count_frames(offsets)
    T = int(fo[-1])
    scores = out if out is not None else np.zeros((T, nn.n_emissions), np.float32)
    capi.check(capi.lib().rb_pipeline_nn_score(frontend.handle, postproc.handle if postproc else None, nn.handle,
                                               capi.ptr(samples), capi.ptr(offsets), offsets.size - 1,
                                               capi.ptr(scores)))
    return scores, fo


def nn_score_utterances_dev(frontend, postproc, nn, d_samples, offsets, d_feats, d_post, d_scores, stream=None):
    offsets = np.ascontiguousarray(offsets, np.int64)
    capi.check(capi.lib().rb_pipeline_nn_score_dev(frontend.handle, postproc.handle if postproc else None, nn.handle,
                                                   capi.ptr(d_samples), capi.ptr(offsets), offsets.size - 1,
                                                   capi.ptr(d_feats), capi.ptr(d_post), capi.ptr(d_scores),
                                                   capi.ptr(stream)))


def partition(lengths, world_size):
    """Assign utterances to ranks: longest first onto the least loaded rank.  Returns a list of index
    arrays (one per rank, each sorted ascending so a rank walks the corpus in order)."""
    lengths = np.asarray(lengths, np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, np.int64)
    parts = [[] for _ in range(world_size)]
    for u in order:
        r = int(np.argmin(load))
        parts[r].append(int(u))
        load[r] += int(lengths[u])
    return [np.array(sorted(p), np.int64) for p in parts]


def gather_scores(local_scores, dist, group=None, out=None):
    """All-gather of the per-rank score slabs through torch.distributed (the library baseline; the engine's own exchange
    is rasr_b200.comm.ScoreExchange over NVLink peer memory).  Shards may differ in length: every rank's rows are
    broadcast in place into one preallocated matrix -- no padding, no staging copy.  `local_scores` is a torch tensor
    [T_r x M] on this rank's device; returns the list of per-rank views of the gathered matrix (`out` if given)."""
    import torch

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([local_scores.shape[0]], dtype=torch.int64, device=local_scores.device)
    counts = torch.zeros(world, dtype=torch.int64, device=local_scores.device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = [int(c) for c in counts.tolist()]
    total, M = sum(counts), local_scores.shape[1]
    if out is None:
        out = torch.empty((total, M), dtype=local_scores.dtype, device=local_scores.device)
    offs = np.concatenate([[0], np.cumsum(counts)])
    views = [out[int(offs[r]):int(offs[r + 1])] for r in range(world)]
    if views[rank].data_ptr() != local_scores.data_ptr():
        views[rank].copy_(local_scores)
    if len(set(counts)) == 1 and total:
        dist.all_gather_into_tensor(out, views[rank], group=group)
    else:
        work = [dist.broadcast(views[r], src=dist.get_global_rank(group, r) if group is not None else r, group=group,
                               async_op=True) for r in range(world) if counts[r]]
        for wk in work:
            wk.wait()
    return views
