"""LPC features with the reference's Python-facing API (src/feature/LPC.py:12-75), computed by csrc/lpc.cu.

    extract(fs, signal=None, diff=False, **kwargs)                                   LPC.py:66-75
    get_lpc_extractor(fs, win_length_ms=32, win_shift_ms=16, n_lpc=15, pre_emphasis_coef=0.95)   :59-63
    LPCExtractor(...).extract(signal)                                                :46-57

The reference delegates the arithmetic to scikits.talkbox (`levinson_lpc.lpc(frame, n_lpc)[0][1:]`, :40-42): the
autocorrelation method solved by Levinson-Durbin.  talkbox is third-party and unpinned; parity is pinned against the
in-repo prototype of the same algorithm (src/lpcc/lpc_jp.py:13-66) -- see DESIGN.md section 5.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .utils import cached_func, diff_feature


class LPCExtractor(object):
    def __init__(self, fs, win_length_ms, win_shift_ms, n_lpc, pre_emphasis_coef):
        self.PRE_EMPH = pre_emphasis_coef
        self.n_lpc = n_lpc
        L = _lib.lib()
        flen = int(float(win_length_ms) / 1000 * fs)
        nfft = 64
        while nfft < flen:
            nfft *= 2
        # the plan only supplies the frame geometry and the float64 window tables to the LPC kernel
        self._h = _lib.check_ptr(L.srb200_mfcc_create(int(fs), float(win_length_ms), float(win_shift_ms), int(max(nfft, 256)), 20, 12,
                                                      float(pre_emphasis_coef)), "LPCExtractor")
        self.FRAME_LEN = L.srb200_mfcc_frame_len(self._h)
        self.FRAME_SHIFT = L.srb200_mfcc_frame_shift(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().srb200_mfcc_free(self._h)
                self._h = None
        except Exception:
            pass

    def extract(self, signal):
        signal = np.asarray(signal)
        if signal.ndim > 1:
            signal = np.mean(signal, axis=1)
        n = len(signal)
        frames = (n - self.FRAME_LEN) // self.FRAME_SHIFT + 1          # LPC.py:47
        out = np.empty((max(frames, 0), self.n_lpc), dtype=np.float64)
        got = C.c_int(0)
        L = _lib.lib()
        if signal.dtype == np.int16:
            sig = np.ascontiguousarray(signal)
            rc = L.srb200_lpc_extract_i16(self._h, sig.ctypes.data_as(C.POINTER(C.c_int16)), n, self.n_lpc, _lib.dptr(out), C.byref(got))
        elif signal.dtype == np.float32:
            sig = np.ascontiguousarray(signal)
            rc = L.srb200_lpc_extract_f32(self._h, _lib.fptr(sig), n, self.n_lpc, _lib.dptr(out), C.byref(got))
        else:
            sig = np.ascontiguousarray(signal, dtype=np.float64)
            rc = L.srb200_lpc_extract_f64(self._h, _lib.dptr(sig), n, self.n_lpc, _lib.dptr(out), C.byref(got))
        _lib.check(rc, "LPC.extract")
        return out


@cached_func
def get_lpc_extractor(fs, win_length_ms=32, win_shift_ms=16, n_lpc=15, pre_emphasis_coef=0.95):
    return LPCExtractor(fs, win_length_ms, win_shift_ms, n_lpc, pre_emphasis_coef)


def extract(fs, signal=None, diff=False, **kwargs):
    """accept two argument, or one as a tuple (LPC.py:66-75)"""
    if signal is None:
        assert type(fs) == tuple
        fs, signal = fs[0], fs[1]
    ret = get_lpc_extractor(fs, **kwargs).extract(signal)
    if diff:
        return diff_feature(ret)
    return ret
