"""Host-side mirror of the legacy Nn feed-forward scorer (src/Nn/BatchFeatureScorer.cc:45-171) and of
the neural-network-forward Flow node (src/Nn/NeuralNetworkForwardNode.cc:140-256)."""
import ctypes as C

import numpy as np

from . import capi


def parameters_from_matrix(param):
    """Split the reference's per-layer parameter matrix: row = output unit, column 0 = bias,
    columns 1.. = weights (src/Nn/LinearLayer.cc:383-424).  Returns (weights[out,in], bias[out])."""
    param = np.asarray(param, np.float32)
    return np.ascontiguousarray(param[:, 1:]), np.ascontiguousarray(param[:, 0])


class NnScorer:
    def __init__(self, dims, acts, weights, biases, log_prior=None, prior_scale=1.0, precision="bf16", device=0):
        n = len(weights)
        assert len(dims) == n + 1 and len(acts) == n
        self.dims = [int(d) for d in dims]
        self._dims = np.asarray(self.dims, np.int32)
        self._acts = np.asarray([capi.ACT[a] if isinstance(a, str) else int(a) for a in acts], np.int32)
        self._w = [np.ascontiguousarray(w, np.float32) for w in weights]
        self._b = [np.ascontiguousarray(b, np.float32) if b is not None else None for b in biases]
        for l in range(n):
            if self._w[l].shape != (self.dims[l + 1], self.dims[l]):
                raise capi.RasrB200Error(-1, "weights[%d] has shape %s, expected (out=%d, in=%d)"
                                         % (l, self._w[l].shape, self.dims[l + 1], self.dims[l]))
        wp = (C.c_void_p * n)(*[w.ctypes.data for w in self._w])
        bp = (C.c_void_p * n)(*[(b.ctypes.data if b is not None else None) for b in self._b])
        lp = np.ascontiguousarray(log_prior, np.float32) if log_prior is not None else None
        prec = {"f32": capi.NN_F32, "bf16": capi.NN_BF16}[precision] if isinstance(precision, str) else precision
        self._h = C.c_void_p()
        capi.check(capi.lib().rb_nn_create(n, capi.ptr(self._dims), capi.ptr(self._acts), wp, bp, capi.ptr(lp),
                                           float(prior_scale if lp is not None else 0.0), prec, device,
                                           C.byref(self._h)))
        self.n_inputs, self.n_outputs = self.dims[0], self.dims[-1]
        self.n_emissions = self.n_outputs

    @classmethod
    def from_files(cls, layer_files, acts, prior_file=None, prior_scale=1.0, **kw):
        """Build the scorer from the reference's per-layer parameter files (one Math::Matrix per layer, row = output
        unit, column 0 = bias: LinearLayer::loadNetworkParameters / setParameters, src/Nn/LinearLayer.cc:219-237,
        383-424) and its log-prior vector file (Prior::read, src/Nn/Prior.cc:216-228).  File names may carry the
        "bin:" / "xml:" qualifier."""
        from . import io
        ws, bs = [], []
        for f in layer_files:
            w, b = parameters_from_matrix(io.read_matrix(f))
            ws.append(w)
            bs.append(b)
        for l in range(1, len(ws)):
            if ws[l].shape[1] != ws[l - 1].shape[0]:
                raise capi.RasrB200Error(-1, "dimension mismatch: (parameter file vs. layer-dimension) %d vs. %d"
                                         % (ws[l].shape[1], ws[l - 1].shape[0]))
        dims = [ws[0].shape[1]] + [w.shape[0] for w in ws]
        prior = io.read_vector(prior_file) if prior_file else None
        return cls(dims, acts, ws, bs, prior, prior_scale, **kw)

    def set_class_mapping(self, class_to_output):
        """Nn::ClassLabelWrapper: emission class -> network output, -1 = disregarded class (score FLT_MAX); None removes
        the mapping.  Affects score() / score_dev() only."""
        if class_to_output is None:
            capi.check(capi.lib().rb_nn_set_class_mapping(self._h, 0, None))
            self.n_emissions = self.n_outputs
            return
        m = np.ascontiguousarray(class_to_output, np.int32)
        capi.check(capi.lib().rb_nn_set_class_mapping(self._h, int(m.size), capi.ptr(m)))
        self.n_emissions = int(m.size)

    def close(self):
        if getattr(self, "_h", None) and capi is not None:  # capi is None while the interpreter shuts down
            capi.lib().rb_nn_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def score(self, feats, out=None):
        """-(w.h + b - scale*logprior); the top-layer softmax is not evaluated."""
        if isinstance(feats, np.ndarray):
            feats = np.ascontiguousarray(feats, np.float32)
        T = int(feats.shape[0])
        scores = out if out is not None else np.zeros((T, self.n_emissions), np.float32)
        capi.check(capi.lib().rb_nn_score(self._h, capi.ptr(feats), T, capi.ptr(scores)))
        return scores

    def forward(self, feats, out=None):
        if isinstance(feats, np.ndarray):
            feats = np.ascontiguousarray(feats, np.float32)
        T = int(feats.shape[0])
        res = out if out is not None else np.zeros((T, self.n_outputs), np.float32)
        capi.check(capi.lib().rb_nn_forward(self._h, capi.ptr(feats), T, capi.ptr(res)))
        return res

    def score_dev(self, d_feats, T, d_scores, stream=None):
        capi.check(capi.lib().rb_nn_score_dev(self._h, capi.ptr(d_feats), int(T), capi.ptr(d_scores),
                                              capi.ptr(stream)))

    def forward_dev(self, d_feats, T, d_out, stream=None):
        capi.check(capi.lib().rb_nn_forward_dev(self._h, capi.ptr(d_feats), int(T), capi.ptr(d_out),
                                                capi.ptr(stream)))


def test_gemm_bf16(a, b, bias=None, act="linear", device=0):
    """One tcgen05 GEMM through the C ABI test hook: act(a @ b.T + bias), operands rounded to bf16."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    M, K = a.shape
    N = b.shape[0]
    d = np.zeros((M, N), np.float32)
    bias = np.ascontiguousarray(bias, np.float32) if bias is not None else None
    capi.check(capi.lib().rb_test_gemm_bf16(capi.ptr(a), capi.ptr(b), capi.ptr(bias), M, N, K, capi.ACT[act],
                                            capi.ptr(d), device))
    return d
