"""Host-side mirror of the score consumer: Search::LinearSearch (src/Search/LinearSearch.cc:233-468) over the dense
score matrix (rb_search_*).  The reference protocol restart() / feed(scorer) per frame / getCurrentBestSentence()
becomes one decode() over whole segments: the scores of a segment are already on the device."""
import ctypes as C

import numpy as np

from . import capi


class LinearSearch:
    """lexicon: dict(word_offsets, state_emission, state_tdp_model, tdp [n_models x 4: loop, forward, skip, exit],
    entry_model, unigram[, word_regular, single_word]) -- see rasr_b200.synth.lexicon.  single_word = the
    recognizer's "single-word-recognition" (the reference's default is true; here it must be asked for), word_regular
    marks the lemmata with a non-empty evaluation sequence (0: silence / noise)."""

    def __init__(self, lexicon, device=0):
        self._a = dict(word_offsets=np.ascontiguousarray(lexicon["word_offsets"], np.uint32),
                       state_emission=np.ascontiguousarray(lexicon["state_emission"], np.uint32),
                       state_tdp_model=np.ascontiguousarray(lexicon["state_tdp_model"], np.uint32),
                       tdp=np.ascontiguousarray(lexicon["tdp"], np.float32).reshape(-1, 4),
                       unigram=np.ascontiguousarray(lexicon["unigram"], np.float32))
        a = self._a
        if lexicon.get("word_regular") is not None:
            a["word_regular"] = np.ascontiguousarray(lexicon["word_regular"], np.uint8)
            if a["word_regular"].size != a["unigram"].size:
                raise ValueError("word_regular needs one flag per word")
        P = lambda x, t: x.ctypes.data_as(C.POINTER(t))
        c = capi.LexiconC(a["word_offsets"].size - 1, P(a["word_offsets"], C.c_uint32),
                          P(a["state_emission"], C.c_uint32), P(a["state_tdp_model"], C.c_uint32), a["tdp"].shape[0],
                          P(a["tdp"], C.c_float), int(lexicon["entry_model"]), P(a["unigram"], C.c_float),
                          P(a["word_regular"], C.c_uint8) if "word_regular" in a else None,
                          int(bool(lexicon.get("single_word", False))))
        self._h = C.c_void_p()
        capi.check(capi.lib().rb_search_create(C.byref(c), int(device), C.byref(self._h)))
        self._fo = None

    def close(self):
        if getattr(self, "_h", None) and capi is not None:  # capi is None while the interpreter shuts down
            capi.lib().rb_search_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def _result(self):
        n_utt = self._fo.size - 1
        cap = max(1, int(self._fo[-1] - self._fo[0]))
        wo = np.zeros(n_utt + 1, np.int64)
        words, times = np.empty(cap, np.uint32), np.empty(cap, np.int32)  # filled up to the returned count
        am, lm = np.empty(cap, np.float32), np.empty(cap, np.float32)
        n = capi.lib().rb_search_traceback_all(self._h, capi.ptr(wo), capi.ptr(words), capi.ptr(times), capi.ptr(am),
                                               capi.ptr(lm), cap)
        if n < 0:
            capi.check(int(n))
        return [dict(words=words[wo[u]:wo[u + 1]], times=times[wo[u]:wo[u + 1]], am=am[wo[u]:wo[u + 1]],
                     lm=lm[wo[u]:wo[u + 1]]) for u in range(n_utt)]

    def traceback(self, utt):
        """one segment of the last decode (rb_search_traceback)"""
        cap = max(1, int(self._fo[utt + 1] - self._fo[utt]))
        words, times = np.zeros(cap, np.uint32), np.zeros(cap, np.int32)
        am, lm = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        n = capi.lib().rb_search_traceback(self._h, utt, capi.ptr(words), capi.ptr(times), capi.ptr(am), capi.ptr(lm))
        if n < 0:
            capi.check(int(n))
        return dict(words=words[:n], times=times[:n], am=am[:n], lm=lm[:n])

    def decode(self, scores, frame_offsets=None):
        """Host score matrix [frames x n_emissions]; returns one traceback dict per segment."""
        if isinstance(scores, np.ndarray) or not hasattr(scores, "data_ptr"):
            scores = np.ascontiguousarray(scores, np.float32)
        self._fo = np.ascontiguousarray(frame_offsets if frame_offsets is not None else [0, scores.shape[0]], np.int64)
        capi.check(capi.lib().rb_search_decode(self._h, capi.ptr(scores), int(scores.shape[1]), capi.ptr(self._fo),
                                               self._fo.size - 1))
        return self._result()

    def decode_dev(self, d_scores, n_emissions, frame_offsets, stream=None, want_result=True):
        """Score matrix on the device.  stream=None is this handle's own stream: the producer of `d_scores` must have
        finished (or run on the stream passed here)."""
        self._fo = np.ascontiguousarray(frame_offsets, np.int64)
        capi.check(capi.lib().rb_search_decode_dev(self._h, capi.ptr(d_scores), int(n_emissions), capi.ptr(self._fo),
                                                   self._fo.size - 1, capi.ptr(stream)))
        return self._result() if want_result else None
