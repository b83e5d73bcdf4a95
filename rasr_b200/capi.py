"""ctypes binding of the C ABI declared in include/rasr_b200.h (rasr_b200/lib/librasr_b200.so).

This is the only way the Python side reaches the kernels: the same entry points the RASR-side C++
adapters bind.  There is no CPU fallback -- if the library is missing or no sm_100
device is present every call fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RASR_B200_LIB") or os.path.join(_HERE, "lib", "librasr_b200.so")  # override: experiments

RB_OK = 0
STATUS = {0: "RB_OK", -1: "RB_ERR_INVALID", -2: "RB_ERR_NO_DEVICE", -3: "RB_ERR_CUDA", -4: "RB_ERR_UNSUPPORTED",
          -5: "RB_ERR_STATE", -6: "RB_ERR_NOMEM"}

GMM_BATCH_FLOAT, GMM_DIAG_MAX, GMM_DIAG_SUM, GMM_BATCH_TENSOR, GMM_BATCH_INT, GMM_BATCH_PRESELECT = 0, 1, 2, 3, 4, 5
GMM_BATCH_PRESELECT_INT = 6
GMM_SIMD_DIAG_MAX = 7
ACT = {"linear": 0, "sigmoid": 1, "relu": 2, "rectified": 2, "softmax": 3, "tanh": 4}
NN_F32, NN_BF16 = 0, 1

# every symbol include/rasr_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "rb_last_error", "rb_version", "rb_device_count", "rb_launch_count",
    "rb_host_alloc", "rb_host_free", "rb_host_register", "rb_host_unregister", "rb_host_is_pinned",
    "rb_frontend_default_cfg", "rb_frontend_create", "rb_frontend_destroy", "rb_frontend_get_geometry",
    "rb_frontend_get_tables", "rb_frontend_nframes_for", "rb_frontend_timestamps", "rb_frontend_reset", "rb_frontend_push",
    "rb_frontend_finish", "rb_frontend_nframes", "rb_frontend_read", "rb_frontend_count_frames",
    "rb_frontend_process", "rb_frontend_process_s16", "rb_frontend_process_dev", "rb_frontend_set_debug", "rb_frontend_read_stages",
    "rb_dc_default_cfg", "rb_frontend_dc_max_frames", "rb_frontend_process_dc", "rb_frontend_dc_runs", "rb_frontend_set_dc_detection",
    "rb_gmm_configure_preselection", "rb_gmm_get_clustering", "rb_test_glibc_rand", "rb_test_introsort",
    "rb_gmm_create", "rb_gmm_destroy", "rb_gmm_set_timing", "rb_gmm_get_timing", "rb_gmm_score_fanout_dev", "rb_pipeline_score_fanout_dev", "rb_gmm_n_mixtures", "rb_gmm_dim", "rb_gmm_score", "rb_gmm_score_dev",
    "rb_nn_create", "rb_nn_destroy", "rb_nn_n_outputs", "rb_nn_n_inputs", "rb_nn_set_class_mapping", "rb_nn_n_emissions", "rb_nn_score", "rb_nn_score_dev",
    "rb_nn_forward", "rb_nn_forward_dev", "rb_pipeline_score", "rb_pipeline_score_s16", "rb_pipeline_score_dev", "rb_test_gemm_bf16", "rb_test_gemm_bench",
    "rb_pipeline_nn_score", "rb_pipeline_nn_score_dev",
    "rb_search_create", "rb_search_destroy", "rb_search_decode", "rb_search_decode_dev", "rb_search_traceback", "rb_search_traceback_all", "rb_pipeline_search",
    "rb_postproc_create", "rb_postproc_destroy", "rb_postproc_dim_out", "rb_postproc_process", "rb_postproc_process_dev",
    "rb_comm_create", "rb_comm_destroy", "rb_comm_world", "rb_comm_rank", "rb_comm_window_alloc", "rb_comm_window_attach",
    "rb_comm_window_ptr", "rb_comm_gather_scores_dev", "rb_comm_push_rows_dev", "rb_comm_barrier_dev", "rb_comm_nccl_unique_id", "rb_comm_nccl_init",
    "rb_comm_nccl_version",
]
COMM_P2P, COMM_NCCL = 0, 1
COMM_HANDLE_BYTES, COMM_ID_BYTES = 64, 128


class RasrB200Error(RuntimeError):
    def __init__(self, status, text):
        super().__init__("%s: %s" % (STATUS.get(status, status), text))
        self.status = status


class FrontendCfg(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("window_length_s", C.c_double), ("window_shift_s", C.c_double),
                ("fft_max_input_s", C.c_double), ("filter_width", C.c_double), ("preemphasis_alpha", C.c_float),
                ("n_cepstra", C.c_int), ("derivatives", C.c_int), ("device", C.c_int), ("window_type", C.c_int)]


WINDOW_TYPES = {"hamming": 0, "rectangular": 1, "hanning": 2, "periodic-hanning": 3, "bartlett": 4, "blackman": 5, "kaiser": 6}


class FrontendGeometry(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("win_length", "win_shift", "fft_length", "n_bins", "n_filters", "n_weights", "feat_dim")]


class DcCfg(C.Structure):
    _fields_ = [("min_dc_length_s", C.c_double), ("max_dc_increment", C.c_float),
                ("min_non_dc_segment_length_s", C.c_double), ("maximal_output_size", C.c_int)]


class PostprocCfg(C.Structure):
    _fields_ = [("norm_type", C.c_int), ("norm_length", C.c_long), ("norm_right", C.c_long),
                ("splice_length", C.c_int), ("splice_right", C.c_int), ("matrix_rows", C.c_int),
                ("matrix_cols", C.c_int), ("matrix", C.POINTER(C.c_float)), ("contraction", C.c_int),
                ("device", C.c_int)]


class LexiconC(C.Structure):
    _fields_ = [("n_words", C.c_uint32), ("word_offsets", C.POINTER(C.c_uint32)),
                ("state_emission", C.POINTER(C.c_uint32)), ("state_tdp_model", C.POINTER(C.c_uint32)),
                ("n_models", C.c_uint32), ("tdp", C.POINTER(C.c_float)), ("entry_model", C.c_uint32),
                ("unigram", C.POINTER(C.c_float)), ("word_regular", C.POINTER(C.c_uint8)), ("single_word", C.c_int32)]


class MixtureSetC(C.Structure):
    _fields_ = [("dim", C.c_uint32), ("n_mixtures", C.c_uint32), ("n_densities", C.c_uint32),
                ("n_means", C.c_uint32), ("n_covariances", C.c_uint32), ("mix_offsets", C.POINTER(C.c_uint32)),
                ("mix_density", C.POINTER(C.c_uint32)), ("mix_log_weight", C.POINTER(C.c_double)),
                ("dens_mean", C.POINTER(C.c_uint32)), ("dens_cov", C.POINTER(C.c_uint32)),
                ("means", C.POINTER(C.c_float)), ("variances", C.POINTER(C.c_float))]


_lib = None


def build():
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc"), "all"])


def lib():
    """The loaded C-ABI library.  Raises if it has not been built: the product has no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RasrB200Error(-2, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64p, fp, dp, u32p, ip = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_double), \
        C.POINTER(C.c_uint32), C.POINTER(C.c_int)
    L.rb_last_error.restype = C.c_char_p
    L.rb_version.restype = C.c_char_p
    L.rb_launch_count.restype = C.c_uint64
    L.rb_frontend_default_cfg.argtypes = [C.POINTER(FrontendCfg)]
    L.rb_frontend_default_cfg.restype = None
    L.rb_frontend_create.argtypes = [C.POINTER(FrontendCfg), C.POINTER(vp)]
    L.rb_frontend_destroy.argtypes = [vp]
    L.rb_frontend_destroy.restype = None
    L.rb_frontend_get_geometry.argtypes = [vp, C.POINTER(FrontendGeometry)]
    L.rb_frontend_get_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rb_frontend_nframes_for.argtypes = [vp, C.c_long]
    L.rb_frontend_nframes_for.restype = C.c_long
    L.rb_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.rb_host_free.argtypes = [vp]
    L.rb_host_free.restype = None
    L.rb_host_register.argtypes = [vp, C.c_size_t]
    L.rb_host_unregister.argtypes = [vp]
    L.rb_host_is_pinned.argtypes = [vp]
    L.rb_frontend_timestamps.argtypes = [vp, C.c_long, C.c_double, vp, vp]
    L.rb_frontend_timestamps.restype = C.c_long
    L.rb_frontend_reset.argtypes = [vp]
    L.rb_frontend_push.argtypes = [vp, vp, C.c_long, C.c_double]
    L.rb_frontend_finish.argtypes = [vp]
    L.rb_frontend_nframes.argtypes = [vp]
    L.rb_frontend_nframes.restype = C.c_long
    L.rb_frontend_read.argtypes = [vp, vp, vp, vp]
    L.rb_frontend_count_frames.argtypes = [vp, vp, C.c_int, vp]
    L.rb_frontend_count_frames.restype = C.c_long
    L.rb_frontend_process.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp]
    L.rb_frontend_process_s16.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp]
    L.rb_frontend_process_dev.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    L.rb_dc_default_cfg.argtypes = [C.POINTER(DcCfg)]
    L.rb_dc_default_cfg.restype = None
    L.rb_frontend_dc_max_frames.argtypes = [vp, C.POINTER(DcCfg), vp, C.c_int]
    L.rb_frontend_dc_max_frames.restype = C.c_long
    L.rb_frontend_process_dc.argtypes = [vp, C.POINTER(DcCfg), vp, vp, C.c_int, vp, C.c_long, vp, vp, vp]
    L.rb_frontend_dc_runs.argtypes = [vp, vp, vp, vp, vp, C.c_long, vp]
    L.rb_frontend_set_dc_detection.argtypes = [vp, C.POINTER(DcCfg)]
    L.rb_frontend_dc_runs.restype = C.c_long
    L.rb_frontend_set_debug.argtypes = [vp, C.c_int]
    L.rb_frontend_read_stages.argtypes = [vp, vp, vp, vp]
    L.rb_gmm_create.argtypes = [C.POINTER(MixtureSetC), C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                C.POINTER(vp)]
    L.rb_gmm_destroy.argtypes = [vp]
    L.rb_gmm_destroy.restype = None
    L.rb_gmm_configure_preselection.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float]
    L.rb_gmm_get_clustering.argtypes = [vp, vp, vp, vp]
    L.rb_test_glibc_rand.argtypes = [C.c_uint, C.c_int, vp]
    L.rb_test_glibc_rand.restype = None
    L.rb_test_introsort.argtypes = [vp, C.c_int, vp]
    L.rb_test_introsort.restype = None
    L.rb_gmm_n_mixtures.argtypes = [vp]
    L.rb_gmm_dim.argtypes = [vp]
    L.rb_gmm_score_fanout_dev.argtypes = [vp, vp, C.c_long, C.c_int, vp, vp]
    L.rb_pipeline_score_fanout_dev.argtypes = [vp, vp, vp, vp, C.c_int, vp, C.c_int, vp, vp]
    L.rb_gmm_set_timing.argtypes = [vp, C.c_int]
    L.rb_gmm_get_timing.argtypes = [vp, vp]
    L.rb_gmm_score.argtypes = [vp, vp, C.c_long, vp, vp]
    L.rb_gmm_score_dev.argtypes = [vp, vp, C.c_long, vp, vp, vp]
    L.rb_nn_create.argtypes = [C.c_int, vp, vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, C.POINTER(vp)]
    L.rb_nn_destroy.argtypes = [vp]
    L.rb_nn_destroy.restype = None
    L.rb_nn_n_outputs.argtypes = [vp]
    L.rb_nn_n_inputs.argtypes = [vp]
    L.rb_nn_set_class_mapping.argtypes = [vp, C.c_int, vp]
    L.rb_nn_n_emissions.argtypes = [vp]
    L.rb_nn_score.argtypes = [vp, vp, C.c_long, vp]
    L.rb_nn_score_dev.argtypes = [vp, vp, C.c_long, vp, vp]
    L.rb_nn_forward.argtypes = [vp, vp, C.c_long, vp]
    L.rb_nn_forward_dev.argtypes = [vp, vp, C.c_long, vp, vp]
    L.rb_pipeline_score.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
    L.rb_pipeline_score_s16.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp]
    L.rb_pipeline_score_dev.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp]
    L.rb_pipeline_nn_score.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp]
    L.rb_pipeline_nn_score_dev.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
    L.rb_search_create.argtypes = [C.POINTER(LexiconC), C.c_int, C.POINTER(vp)]
    L.rb_search_destroy.argtypes = [vp]
    L.rb_search_destroy.restype = None
    L.rb_search_decode.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.rb_search_decode_dev.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp]
    L.rb_search_traceback.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.rb_search_traceback.restype = C.c_long
    L.rb_search_traceback_all.argtypes = [vp, vp, vp, vp, vp, vp, C.c_long]
    L.rb_search_traceback_all.restype = C.c_long
    L.rb_pipeline_search.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp, C.c_int]
    L.rb_postproc_create.argtypes = [C.POINTER(PostprocCfg), C.c_int, C.POINTER(vp)]
    L.rb_postproc_destroy.argtypes = [vp]
    L.rb_postproc_destroy.restype = None
    L.rb_postproc_dim_out.argtypes = [vp]
    L.rb_postproc_process.argtypes = [vp, vp, vp, C.c_int, vp]
    L.rb_postproc_process_dev.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    L.rb_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.rb_comm_destroy.argtypes = [vp]
    L.rb_comm_destroy.restype = None
    L.rb_comm_world.argtypes = [vp]
    L.rb_comm_rank.argtypes = [vp]
    L.rb_comm_window_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp), vp]
    L.rb_comm_window_attach.argtypes = [vp, vp]
    L.rb_comm_window_ptr.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.rb_comm_gather_scores_dev.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    L.rb_comm_push_rows_dev.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int, C.c_int, vp]
    L.rb_comm_barrier_dev.argtypes = [vp, vp]
    L.rb_comm_nccl_unique_id.argtypes = [vp]
    L.rb_comm_nccl_init.argtypes = [vp, vp]
    L.rb_test_gemm_bf16.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.rb_test_gemm_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int]
    _lib = L
    return L


def check(rc):
    if rc != RB_OK:
        raise RasrB200Error(rc, lib().rb_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Address of a numpy array / torch tensor / int (device pointer) / None, as c_void_p."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError("cannot take the address of %r" % type(a))


class _PinnedBlock:
    """Owns one rb_host_alloc allocation; freed when the last numpy view of it goes away."""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        check(lib().rb_host_alloc(max(1, int(nbytes)), C.byref(self.ptr)))
        self.nbytes = int(nbytes)

    def __del__(self):
        if getattr(self, "ptr", None) and C is not None and _lib is not None:
            _lib.rb_host_free(self.ptr)
            self.ptr = None


def host_empty(shape, dtype=np.float32):
    """numpy array in page-locked host memory from rb_host_alloc (what the adapters keep their feature / score buffers
    in): copies to and from it run asynchronously at the PCIe rate."""
    dtype = np.dtype(dtype)
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(v) for v in shape)
    n = int(np.prod(shape)) if shape else 1
    block = _PinnedBlock(n * dtype.itemsize)
    buf = (C.c_char * max(1, n * dtype.itemsize)).from_address(block.ptr.value)
    a = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    buf._block = block  # the ctypes buffer is the numpy array's base: the block lives as long as any view of it
    return a


def host_is_pinned(a):
    return bool(lib().rb_host_is_pinned(ptr(a)))


def device_count():
    return int(lib().rb_device_count())


def launch_count():
    return int(lib().rb_launch_count())


def version():
    return lib().rb_version().decode()
