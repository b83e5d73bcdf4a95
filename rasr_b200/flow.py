"""Host-side mirror of the Flow::Node contract for the MFCC front-end (src/Flow/Node.hh:38-185).

`MfccNode` behaves like the adapter class adapters/B200MfccNode.cc: XML attributes arrive through
`set_parameter(name, value)` as text (src/Flow/AbstractNode.hh:149-159), `configure()` reads the
incoming "sample-rate" attribute and publishes the outgoing attributes, samples arrive as packets
with a start time, and on the EOS sentinel the whole segment is computed on the device and emitted
one `Vector<f32>` packet per frame (the batch-node pattern of src/Nn/NeuralNetworkForwardNode.cc:187-256).
"""
import ctypes as C

import numpy as np

from . import capi


class Packet:
    """Flow::Vector<f32> = Timestamp{start,end} + std::vector<f32> (src/Flow/Vector.hh:32-57)."""
    __slots__ = ("start", "end", "data")

    def __init__(self, data, start, end):
        self.data, self.start, self.end = data, start, end


EOS = object()  # Flow::Data::eos() sentinel


class FrontEnd:
    """Thin RAII wrapper of rb_frontend_* (one handle = one node instance = one CUDA stream)."""

    def __init__(self, sample_rate=16000.0, window_length=0.025, window_shift=0.01, fft_max_input=0.025,
                 filter_width=268.258, alpha=1.0, n_cepstra=13, derivatives=True, device=0, window_type="hamming"):
        L = capi.lib()
        if window_type not in capi.WINDOW_TYPES:
            raise capi.RasrB200Error(-4, "unknown window type '%s'" % window_type)
        self.cfg = capi.FrontendCfg(sample_rate, window_length, window_shift, fft_max_input, filter_width, alpha,
                                    n_cepstra, int(derivatives), device, capi.WINDOW_TYPES[window_type])
        self._h = C.c_void_p()
        capi.check(L.rb_frontend_create(C.byref(self.cfg), C.byref(self._h)))
        g = capi.FrontendGeometry()
        capi.check(L.rb_frontend_get_geometry(self._h, C.byref(g)))
        self.geometry = g
        self.feat_dim = g.feat_dim

    def close(self):
        if getattr(self, "_h", None) and capi is not None:  # capi is None while the interpreter shuts down
            capi.lib().rb_frontend_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def tables(self):
        g = self.geometry
        window = np.zeros(g.win_length, np.float32)
        start = np.zeros(g.n_filters, np.int32)
        end = np.zeros(g.n_filters, np.int32)
        w = np.zeros((g.n_filters, g.n_bins), np.float32)
        dct = np.zeros((self.cfg.n_cepstra, g.n_filters), np.float32)
        capi.check(capi.lib().rb_frontend_get_tables(self._h, capi.ptr(window), capi.ptr(start), capi.ptr(end),
                                                     capi.ptr(w), capi.ptr(dct)))
        return dict(window=window, fb_start=start, fb_end=end, fb_weights=w, dct=dct)

    def nframes_for(self, n_samples):
        return int(capi.lib().rb_frontend_nframes_for(self._h, n_samples))

    def count_frames(self, offsets):
        offsets = np.ascontiguousarray(offsets, np.int64)
        fo = np.zeros(offsets.size, np.int64)
        n = capi.lib().rb_frontend_count_frames(self._h, capi.ptr(offsets), offsets.size - 1, capi.ptr(fo))
        if n < 0:
            raise capi.RasrB200Error(-1, "bad offsets")
        return fo

    def process(self, samples, offsets=None, timestamps=True, stages=False, out=None):
        """Batch of independent segments (host buffers: numpy, or pinned torch CPU tensors through their
        data_ptr; `out` receives the features).  Returns dict(feats, frame_offsets, t_start, t_end)."""
        if isinstance(samples, np.ndarray) or not hasattr(samples, "data_ptr"):
            samples = np.ascontiguousarray(samples, np.float32)
        if offsets is None:
            offsets = np.array([0, samples.size], np.int64)
        offsets = np.ascontiguousarray(offsets, np.int64)
        fo = self.count_frames(offsets)
        T = int(fo[-1])
        feats = out if out is not None else np.zeros((T, self.feat_dim), np.float32)
        ts = np.zeros(T, np.float64) if timestamps else None
        te = np.zeros(T, np.float64) if timestamps else None
        L = capi.lib()
        if stages:
            capi.check(L.rb_frontend_set_debug(self._h, 1))
        try:
            capi.check(L.rb_frontend_process(self._h, capi.ptr(samples), capi.ptr(offsets), offsets.size - 1,
                                             capi.ptr(feats), capi.ptr(ts), capi.ptr(te)))
            out = dict(feats=feats, frame_offsets=fo, t_start=ts, t_end=te)
            if stages:
                g = self.geometry
                amp = np.zeros((T, g.n_bins), np.float32)
                fb = np.zeros((T, g.n_filters), np.float32)
                cep = np.zeros((T, self.cfg.n_cepstra), np.float32)
                capi.check(L.rb_frontend_read_stages(self._h, capi.ptr(amp), capi.ptr(fb), capi.ptr(cep)))
                out.update(amplitude=amp, fbank=fb, cepstra=cep)
        finally:
            if stages:
                L.rb_frontend_set_debug(self._h, 0)
        return out

    def process_dc(self, samples, offsets=None, dc=None, timestamps=True):
        """signal-dc-detection (samples.flow:34-37) in front of the chain.  `dc`: capi.DcCfg or a dict of its fields
        (default: the values samples.flow sets).  Returns dict(feats, frame_offsets, t_start, t_end, runs) with
        runs = dict(utt, begin, end, start, sequential_path)."""
        samples = np.ascontiguousarray(samples, np.float32)
        if offsets is None:
            offsets = np.array([0, samples.size], np.int64)
        offsets = np.ascontiguousarray(offsets, np.int64)
        L = capi.lib()
        cfg = capi.DcCfg()
        L.rb_dc_default_cfg(C.byref(cfg))
        if isinstance(dc, capi.DcCfg):
            cfg = dc
        elif dc:
            for k, v in dc.items():
                setattr(cfg, k, v)
        n_utt = offsets.size - 1
        cap = int(L.rb_frontend_dc_max_frames(self._h, C.byref(cfg), capi.ptr(offsets), n_utt))
        if cap < 0:
            capi.check(cap)
        feats = np.zeros((cap, self.feat_dim), np.float32)
        ts = np.zeros(cap, np.float64) if timestamps else None
        te = np.zeros(cap, np.float64) if timestamps else None
        fo = np.zeros(n_utt + 1, np.int64)
        capi.check(L.rb_frontend_process_dc(self._h, C.byref(cfg), capi.ptr(samples), capi.ptr(offsets), n_utt,
                                            capi.ptr(feats), cap, capi.ptr(fo), capi.ptr(ts), capi.ptr(te)))
        T = int(fo[-1])
        seq = C.c_int(0)
        n = int(L.rb_frontend_dc_runs(self._h, None, None, None, None, 0, C.byref(seq)))
        ru, rb, re = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
        rs = np.zeros(n, np.float64)
        L.rb_frontend_dc_runs(self._h, capi.ptr(ru), capi.ptr(rb), capi.ptr(re), capi.ptr(rs), n, None)
        return dict(feats=feats[:T], frame_offsets=fo, t_start=ts[:T] if timestamps else None,
                    t_end=te[:T] if timestamps else None,
                    runs=dict(utt=ru, begin=rb, end=re, start=rs, sequential_path=bool(seq.value)))

    def process_s16(self, pcm, offsets=None, n_channels=1, track=0, timestamps=True, out=None):
        """16-bit PCM in (numpy int16 [frames] or [frames, n_channels] interleaved, or a pinned torch tensor);
        demultiplexed and converted on the device (samples.flow:13-18)."""
        if isinstance(pcm, np.ndarray) or not hasattr(pcm, "data_ptr"):
            pcm = np.ascontiguousarray(pcm, np.int16)
        n_frames = int(pcm.shape[0]) if len(pcm.shape) == 2 else int(pcm.shape[0]) // n_channels
        if offsets is None:
            offsets = np.array([0, n_frames], np.int64)
        offsets = np.ascontiguousarray(offsets, np.int64)
        fo = self.count_frames(offsets)
        T = int(fo[-1])
        feats = out if out is not None else np.zeros((T, self.feat_dim), np.float32)
        ts = np.zeros(T, np.float64) if timestamps else None
        te = np.zeros(T, np.float64) if timestamps else None
        capi.check(capi.lib().rb_frontend_process_s16(self._h, capi.ptr(pcm), int(n_channels), int(track),
                                                      capi.ptr(offsets), offsets.size - 1, capi.ptr(feats),
                                                      capi.ptr(ts), capi.ptr(te)))
        return dict(feats=feats, frame_offsets=fo, t_start=ts, t_end=te)

    def process_dev(self, d_samples, offsets, d_feats, stream=None):
        """Device buffers (torch tensors or raw pointers); only enqueues + syncs the tile tables."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        capi.check(capi.lib().rb_frontend_process_dev(self._h, capi.ptr(d_samples), capi.ptr(offsets),
                                                      offsets.size - 1, capi.ptr(d_feats), capi.ptr(stream)))

    def set_dc_detection(self, dc=None, enabled=True):
        """streaming protocol: run signal-dc-detection over the pushed samples at finish()"""
        L = capi.lib()
        if not enabled:
            capi.check(L.rb_frontend_set_dc_detection(self._h, None))
            return
        cfg = capi.DcCfg()
        L.rb_dc_default_cfg(C.byref(cfg))
        for k, v in (dc or {}).items():
            setattr(cfg, k, v)
        capi.check(L.rb_frontend_set_dc_detection(self._h, C.byref(cfg)))

    # streaming protocol
    def reset(self):
        capi.check(capi.lib().rb_frontend_reset(self._h))

    def push(self, samples, start_time):
        samples = np.ascontiguousarray(samples, np.float32)
        capi.check(capi.lib().rb_frontend_push(self._h, capi.ptr(samples), samples.size, float(start_time)))

    def finish(self):
        L = capi.lib()
        capi.check(L.rb_frontend_finish(self._h))
        T = int(L.rb_frontend_nframes(self._h))
        feats = np.zeros((T, self.feat_dim), np.float32)
        ts = np.zeros(T, np.float64)
        te = np.zeros(T, np.float64)
        capi.check(L.rb_frontend_read(self._h, capi.ptr(feats), capi.ptr(ts), capi.ptr(te)))
        return feats, ts, te


class MfccNode:
    """Flow::Node-shaped adapter: filter name, setParameter, configure, work (pull until EOS, then emit)."""

    PARAMS = {"alpha": ("alpha", float), "length": ("window_length", float), "shift": ("window_shift", float),
              "maximum-input-size": ("fft_max_input", float), "filter-width": ("filter_width", float),
              "nr-outputs": ("n_cepstra", int), "derivatives": ("derivatives", lambda v: v in ("true", "1", "yes")),
              "device": ("device", int), "window-type": ("window_type", str)}
    # signal-dc-detection in front of the chain, parameter names of src/Signal/DcDetection.cc:231-241
    DC_PARAMS = {"min-dc-length": ("min_dc_length_s", float), "max-dc-increment": ("max_dc_increment", float),
                 "min-non-dc-segment-length": ("min_non_dc_segment_length_s", float),
                 "maximal-output-size": ("maximal_output_size", int)}

    @staticmethod
    def filter_name():
        return "b200-mfcc"

    def __init__(self):
        self._kw = {}
        self._dc, self._dc_on = {}, False
        self._fe = None
        self._out = []
        self._segment_open = False
        self.output_attributes = {}

    def set_parameter(self, name, value):
        """Returns False for unknown names, like Flow::AbstractNode::setParameter."""
        if name == "dc-detection":
            self._dc_on = value in ("true", "1", "yes")
            self._fe = None
            return True
        if name in self.DC_PARAMS:
            key, conv = self.DC_PARAMS[name]
            self._dc[key] = conv(value)
            self._fe = None
            return True
        if name not in self.PARAMS:
            return False
        key, conv = self.PARAMS[name]
        self._kw[key] = conv(value)
        self._fe = None
        return True

    def configure(self, input_attributes):
        """input_attributes: dict of text attributes; needs "sample-rate" (src/Signal/Window.cc:158-177)."""
        if input_attributes.get("datatype", "vector-f32") != "vector-f32":
            return False
        if "sample-rate" not in input_attributes:
            return False
        sr = float(input_attributes["sample-rate"])
        self._fe = FrontEnd(sample_rate=sr, **self._kw)
        self._fe.set_dc_detection(self._dc, self._dc_on)
        self.output_attributes = self.output_attributes_for(self._kw.get("window_shift", 0.01))
        return True

    @staticmethod
    def output_attributes_for(window_shift_s):
        """What the replaced chain leaves in the attributes: the window node adds "frame-shift" = its shift parameter
        (src/Signal/Window.cc:166), the cosine transform sets "sample-rate" to 1 (src/Signal/CosineTransform.cc:208);
        values travel as text with 6 significant digits (src/Flow/Attributes.hh:104-113)."""
        return {"datatype": "vector-f32", "sample-rate": "1", "frame-shift": "%g" % window_shift_s}

    def put(self, packet):
        """Input side of work(): a Packet of samples, or EOS."""
        if self._fe is None:
            raise capi.RasrB200Error(-5, "node used before configure()")
        if packet is EOS:
            feats, ts, te = self._fe.finish()
            self._out = [Packet(feats[t], ts[t], te[t]) for t in range(feats.shape[0])]
            self._out.append(EOS)
            self._fe.reset()
            self._segment_open = False
            return
        if not self._segment_open:
            self._fe.reset()
            self._segment_open = True
        self._fe.push(packet.data, packet.start)

    def work(self):
        """Output side: one packet per call; EOS terminates the segment."""
        if not self._out:
            return None
        return self._out.pop(0)
