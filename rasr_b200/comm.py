"""Score exchange between the GPUs of one box (include/rasr_b200.h, rb_comm_*): host side.

One process per GPU.  Every rank owns a *window* -- its copy of the gathered score matrix in HBM -- that is mapped into
all processes; a shard reaches the consumer either because the scorer writes there directly over NVLink
(`ScoreExchange.target(root)` as the scorer's output) or by the push kernel / NCCL (`gather`).  The byte strings the
ranks must swap (IPC handles, the NCCL id) travel through `exchange`, any callable that returns every rank's bytes in
rank order -- `torch_exchange(dist)` uses torch.distributed (gloo or nccl).

The reference has no counterpart: its only parallelism is N independent processes over corpus partitions
(src/Bliss/CorpusDescription.cc:173-180).
"""
import ctypes as C

import numpy as np

from . import capi


def torch_exchange(dist, group=None):
    """all-gather of one small byte string per rank through torch.distributed"""

    def exchange(payload):
        out = [None] * dist.get_world_size(group)
        dist.all_gather_object(out, bytes(payload), group=group)
        return out

    return exchange


class _DevArray:
    """__cuda_array_interface__ view of raw device memory (for torch.as_tensor / cupy.asarray)"""

    def __init__(self, ptr, shape, typestr="<f4", owner=None):
        self.__cuda_array_interface__ = dict(shape=tuple(int(s) for s in shape), typestr=typestr, data=(int(ptr), False),
                                             version=2)
        self._owner = owner


class ScoreExchange:
    def __init__(self, world, rank, device, row_offsets, row_len, exchange=None, nccl=False):
        self.world, self.rank, self.device, self.row_len = int(world), int(rank), int(device), int(row_len)
        self.row_offsets = np.ascontiguousarray(row_offsets, np.int64)
        if self.row_offsets.shape != (self.world + 1,) or self.row_offsets[0] != 0:
            raise ValueError("row_offsets must be [world + 1] prefix sums starting at 0")
        self.rows = int(self.row_offsets[-1])
        L = capi.lib()
        self.handle = C.c_void_p()
        capi.check(L.rb_comm_create(self.world, self.rank, self.device, C.byref(self.handle)))
        win, mine = C.c_void_p(), (C.c_char * capi.COMM_HANDLE_BYTES)()
        capi.check(L.rb_comm_window_alloc(self.handle, max(1, self.rows * self.row_len * 4), C.byref(win), mine))
        self.window_ptr = int(win.value)
        if self.world > 1:
            if exchange is None:
                raise ValueError("world > 1 needs an `exchange` callable (e.g. torch_exchange(dist))")
            allh = b"".join(exchange(bytes(mine)))
            if len(allh) != self.world * capi.COMM_HANDLE_BYTES:
                raise ValueError("exchange() must return one %d-byte handle per rank" % capi.COMM_HANDLE_BYTES)
            capi.check(L.rb_comm_window_attach(self.handle, allh))
        self.nccl = False
        if nccl and self.world > 1:
            uid = (C.c_char * capi.COMM_ID_BYTES)()
            if self.rank == 0:
                capi.check(L.rb_comm_nccl_unique_id(uid))
            uid0 = exchange(bytes(uid))[0]
            capi.check(L.rb_comm_nccl_init(self.handle, uid0))
            self.nccl = True

    def close(self):
        if self.handle:
            capi.lib().rb_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---- addresses
    def peer_window(self, peer):
        p = C.c_void_p()
        capi.check(capi.lib().rb_comm_window_ptr(self.handle, int(peer), C.byref(p)))
        return int(p.value)

    def target(self, peer):
        """device address (int) of THIS rank's rows inside rank `peer`'s window: hand it to a scorer as its output buffer
        and the scores are stored over NVLink as they are computed"""
        return self.peer_window(peer) + int(self.row_offsets[self.rank]) * self.row_len * 4

    def targets(self, root=-1):
        """this rank's rows in the window of `root`, or (root < 0) of every rank, own window first: the destination list
        of rb_gmm_score_fanout_dev / rb_pipeline_score_fanout_dev"""
        if root >= 0:
            return [self.target(root)]
        return [self.target(self.rank)] + [self.target(r) for r in range(self.world) if r != self.rank]

    @property
    def my_rows(self):
        return int(self.row_offsets[self.rank + 1] - self.row_offsets[self.rank])

    def window(self, torch):
        """the local window (the gathered matrix) as a torch tensor [rows x row_len] on this rank's device"""
        return torch.as_tensor(_DevArray(self.window_ptr, (self.rows, self.row_len), owner=self),
                               device=torch.device("cuda", self.device))

    # ---- operations (enqueue only)
    def gather(self, d_send, root=-1, transport="p2p", stream=None):
        t = dict(p2p=capi.COMM_P2P, nccl=capi.COMM_NCCL)[transport]
        capi.check(capi.lib().rb_comm_gather_scores_dev(self.handle, capi.ptr(d_send), capi.ptr(self.row_offsets),
                                                        self.row_len, int(root), t, capi.ptr(stream)))

    def push_rows(self, d_send, first_row, n_rows, root=-1, stream=None):
        """rows [first_row, first_row + n_rows) of the gathered matrix, read from d_send, to `root` or to every rank"""
        capi.check(capi.lib().rb_comm_push_rows_dev(self.handle, capi.ptr(d_send), int(first_row), int(n_rows),
                                                    self.row_len, int(root), capi.ptr(stream)))

    def barrier(self, stream=None):
        capi.check(capi.lib().rb_comm_barrier_dev(self.handle, capi.ptr(stream)))


def nccl_version():
    return int(capi.lib().rb_comm_nccl_version())
