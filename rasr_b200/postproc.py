"""Host-side mirror of the feature post-processing nodes between the MFCC front-end and the scorers
(src/Tools/FeatureExtraction/share/processing.standard_system.flow:25-27, lda.flow:11-19):
signal-normalization -> signal-vector-f32-sequence-concatenation -> signal-matrix-multiplication-f32,
each optional, chained on the device (rb_postproc_*)."""
import ctypes as C

import numpy as np

from . import capi

NORM = {None: 0, "none": 0, "mean": 1, "mean-and-variance": 2}


class PostProcessor:
    """normalization: None | "mean" | "mean-and-variance" with (length, right) in frames, "infinite" or < 0 for the
    whole segment (the reference's length="infinite" right="infinite"); splice=(max_size, right) or None;
    matrix: rows x cols array (rows = output dimension) or None."""

    def __init__(self, dim_in, normalization=None, length="infinite", right="infinite", splice=None, matrix=None,
                 contraction=True, device=0):
        inf = lambda v: -1 if v in ("infinite", None) else int(v)
        cfg = capi.PostprocCfg()
        cfg.norm_type = NORM[normalization]
        cfg.norm_length, cfg.norm_right = inf(length), inf(right)
        cfg.splice_length, cfg.splice_right = (int(splice[0]), int(splice[1])) if splice else (0, 0)
        self._matrix = None
        if matrix is not None:
            self._matrix = np.ascontiguousarray(matrix, np.float32)
            cfg.matrix_rows, cfg.matrix_cols = self._matrix.shape
            cfg.matrix = self._matrix.ctypes.data_as(C.POINTER(C.c_float))
        cfg.contraction, cfg.device = int(contraction), int(device)
        self.dim_in = int(dim_in)
        self._h = C.c_void_p()
        capi.check(capi.lib().rb_postproc_create(C.byref(cfg), self.dim_in, C.byref(self._h)))
        self.dim_out = int(capi.lib().rb_postproc_dim_out(self._h))

    def close(self):
        if getattr(self, "_h", None) and capi is not None:  # capi is None while the interpreter shuts down
            capi.lib().rb_postproc_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def process(self, feats, frame_offsets=None, out=None):
        """Host buffers; segments given by frame_offsets (default: one segment)."""
        if isinstance(feats, np.ndarray) or not hasattr(feats, "data_ptr"):
            feats = np.ascontiguousarray(feats, np.float32)
        T = int(feats.shape[0])
        fo = np.ascontiguousarray(frame_offsets if frame_offsets is not None else [0, T], np.int64)
        res = out if out is not None else np.zeros((T, self.dim_out), np.float32)
        capi.check(capi.lib().rb_postproc_process(self._h, capi.ptr(feats), capi.ptr(fo), fo.size - 1, capi.ptr(res)))
        return res

    def process_dev(self, d_feats, frame_offsets, d_out, stream=None):
        fo = np.ascontiguousarray(frame_offsets, np.int64)
        capi.check(capi.lib().rb_postproc_process_dev(self._h, capi.ptr(d_feats), capi.ptr(fo), fo.size - 1,
                                                      capi.ptr(d_out), capi.ptr(stream)))
