"""Host-side logic that needs no GPU: the buffered FeatureScorer protocol, the Flow node parameter
plumbing, utterance partitioning, and the world_size-2 gather over gloo."""
import os
import socket

import numpy as np
import pytest

from rasr_b200 import capi, flow, mm, pipeline


class FakeDenseScorer:
    """Stands in for GmmScorer: score = sum of the feature + mixture index."""
    n_mixtures, dim = 5, 3

    def __init__(self):
        self.calls = 0

    def score(self, feats):
        self.calls += 1
        return feats.sum(1, keepdims=True) + np.arange(self.n_mixtures, dtype=np.float32)[None]


def test_buffered_scorer_protocol_matches_recognizer_loop():
    """Speech::Recognizer::processFeature / flush loop (src/Speech/Recognizer.cc:271-281,197-205)."""
    dense = FakeDenseScorer()
    fs = mm.BatchFeatureScorer(dense, buffer_size=4)
    assert fs.is_buffered() and fs.buffer_empty() and not fs.buffer_filled()
    feats = np.arange(30, dtype=np.float32).reshape(10, 3)
    got = []
    for f in feats:
        if fs.is_buffered() and not fs.buffer_filled():
            fs.add_feature(f)
        else:
            got.append(fs.get_scorer(f))
    while not fs.buffer_empty():
        got.append(fs.flush())
    assert len(got) == 10
    # the t-th scorer answers for the t-th frame (oldest buffered), not for the frame just pushed
    for t, s in enumerate(got):
        assert s.n_emissions() == 5
        assert s.score(2) == pytest.approx(feats[t].sum() + 2)
    with pytest.raises(capi.RasrB200Error):
        fs.flush()
    fs.reset()
    assert fs.buffer_empty()


def test_whole_segment_scorer_scores_once():
    dense = FakeDenseScorer()
    fs = mm.BatchFeatureScorer(dense)
    feats = np.ones((7, 3), np.float32)
    for f in feats:
        assert not fs.buffer_filled()
        fs.add_feature(f)
    scorers = []
    while not fs.buffer_empty():
        scorers.append(fs.flush())
    assert [s.score(0) for s in scorers] == [3.0] * 7
    assert dense.calls == 1
    with pytest.raises(capi.RasrB200Error):
        fs.add_feature(np.ones(4, np.float32))


def test_mfcc_node_parameters():
    node = flow.MfccNode()
    assert flow.MfccNode.filter_name() == "b200-mfcc"
    assert node.set_parameter("alpha", "0.97")
    assert node.set_parameter("nr-outputs", "16")
    assert not node.set_parameter("no-such-parameter", "1")
    assert not node.configure({"datatype": "vector-f32"})  # no sample-rate attribute
    assert not node.configure({"datatype": "vector-s16", "sample-rate": "16000"})
    with pytest.raises(capi.RasrB200Error):
        node.put(flow.EOS)  # used before a successful configure()
    # the attributes the replaced chain leaves: "sample-rate" 1 (src/Signal/CosineTransform.cc:208), "frame-shift" =
    # the window node's shift parameter (src/Signal/Window.cc:166), as text
    assert flow.MfccNode.output_attributes_for(0.01) == {"datatype": "vector-f32", "sample-rate": "1",
                                                         "frame-shift": "0.01"}
    src = open(os.path.join(os.path.dirname(__file__), "..", "adapters", "B200MfccNode.cc")).read()
    assert 'a->set("sample-rate", 1);' in src and 'a->set("frame-shift", cfg_.window_shift_s);' in src


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    lengths = rng.integers(1000, 200000, 1000)
    parts = pipeline.partition(lengths, 8)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(1000))
    loads = np.array([lengths[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01
    assert all(np.all(np.diff(p) > 0) for p in parts)
    assert [p.tolist() for p in pipeline.partition([5, 1, 1], 2)] == [[0], [1, 2]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [7, 3, 5, 9, 2]
    mine = pipeline.partition(lengths, world)[rank]
    local = torch.cat([torch.full((lengths[u], 4), float(u)) for u in mine]) if len(mine) else torch.zeros((0, 4))
    slabs = pipeline.gather_scores(local, dist)
    q.put((rank, [s.shape[0] for s in slabs], float(sum(s.sum() for s in slabs))))
    dist.destroy_process_group()


def test_gather_scores_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = sum(l * 4 * u for u, l in enumerate([7, 3, 5, 9, 2]))
    for rank, sizes, s in res:
        assert sum(sizes) == 26
        assert s == pytest.approx(total)


def test_score_exchange_fails_loudly_without_a_device():
    """rb_comm_* has no CPU path either: creating a communicator without an sm_100 device is RB_ERR_NO_DEVICE, bad
    world / rank arguments are RB_ERR_INVALID before any device is touched"""
    import ctypes as C

    from rasr_b200 import capi, comm

    L = capi.lib()
    h = C.c_void_p()
    assert L.rb_comm_create(0, 0, 0, C.byref(h)) == -1
    assert L.rb_comm_create(4, 4, 0, C.byref(h)) == -1
    assert L.rb_comm_create(17, 0, 0, C.byref(h)) == -1
    with pytest.raises(ValueError):
        comm.ScoreExchange(2, 0, 0, [0, 5], 8)  # row_offsets must have world + 1 entries
    if capi.device_count() == 0:
        assert L.rb_comm_create(2, 0, 0, C.byref(h)) == -2
        with pytest.raises(capi.RasrB200Error) as e:
            comm.ScoreExchange(1, 0, 0, [0, 5], 8)
        assert e.value.status == -2
