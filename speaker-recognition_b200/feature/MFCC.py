"""MFCC front-end with the reference's Python-facing API, computed by the fused CUDA kernels of
csrc/mfcc.cu through the C ABI (include/srb200.h, srb200_mfcc_*).

Mirrors src/feature/MFCC.py:
    extract(fs, signal=None, diff=False, **kwargs)          :123-132
    get_mfcc_extractor(fs, win_length_ms=32, win_shift_ms=16, FFT_SIZE=2048, n_filters=50, n_ceps=13,
                       pre_emphasis_coef=0.95)              :115-121 (memoised per argument tuple)
    MFCCExtractor(...).extract(signal)                      :20-40, 49-79
Returned arrays are float64 like the reference's; the arithmetic is fp32 on the GPU (float64 DCT accumulate
and CMVN statistics), parity bound 1e-4 * max(1, |ref|) per coefficient (tests/test_mfcc_gpu.py).
There is no CPU path: without libsrb200.so and a CUDA device every call raises.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .utils import cached_func, diff_feature  # noqa: F401  (re-exported like the reference module)

POWER_SPECTRUM_FLOOR = 1e-100  # MFCC.py:7; honoured for digital silence, see csrc/mfcc.cu


class MFCCExtractor(object):
    """Plan object: owns the device tables (window, mel CSR, DCT, twiddles) for one parameter set."""

    def __init__(self, fs, win_length_ms, win_shift_ms, FFT_SIZE, n_bands, n_coefs, PRE_EMPH, verbose=False):
        self.PRE_EMPH = PRE_EMPH
        self.fs = fs
        self.n_bands = n_bands
        self.coefs = n_coefs
        self.FFT_SIZE = FFT_SIZE
        self.verbose = verbose
        L = _lib.lib()
        self._h = _lib.check_ptr(L.srb200_mfcc_create(int(fs), float(win_length_ms), float(win_shift_ms), int(FFT_SIZE),
                                                      int(n_bands), int(n_coefs), float(PRE_EMPH)), "MFCCExtractor")
        self.FRAME_LEN = L.srb200_mfcc_frame_len(self._h)
        self.FRAME_SHIFT = L.srb200_mfcc_frame_shift(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().srb200_mfcc_free(self._h)
                self._h = None
        except Exception:
            pass

    def dprint(self, msg):
        if self.verbose:
            print(msg)

    def num_frames(self, n_samples):
        return (n_samples - self.FRAME_LEN) // self.FRAME_SHIFT + 1

    def extract(self, signal, nd=0, cmvn=True):
        """MFCC of one utterance: float64 [frames - nd, n_ceps * (nd + 1)].  nd=0 is the reference's
        MFCCExtractor.extract; nd=1/2 fuses diff_feature(.., nd) into the device epilogue."""
        signal = np.asarray(signal)
        if signal.ndim > 1:
            self.dprint("INFO: Input signal has more than 1 channel; the channels will be averaged.")
            signal = np.mean(signal, axis=1)                       # MFCC.py:53-55
        assert len(signal) > 5 * self.FRAME_LEN, "Signal too short!"   # MFCC.py:56
        L = _lib.lib()
        n = len(signal)
        rows = self.num_frames(n) - nd
        cols = self.coefs * (nd + 1)
        out = np.empty((rows, cols), dtype=np.float64)
        got = C.c_int(0)
        if signal.dtype == np.int16:
            sig = np.ascontiguousarray(signal)
            rc = L.srb200_mfcc_extract_i16(self._h, sig.ctypes.data_as(C.POINTER(C.c_int16)), n, nd, int(cmvn),
                                           _lib.dptr(out), C.byref(got))
        elif signal.dtype == np.float32:
            sig = np.ascontiguousarray(signal)
            rc = L.srb200_mfcc_extract_f32(self._h, _lib.fptr(sig), n, nd, int(cmvn), _lib.dptr(out), C.byref(got))
        else:
            sig = np.ascontiguousarray(signal, dtype=np.float64)   # cast['float'], MFCC.py:128
            rc = L.srb200_mfcc_extract_f64(self._h, _lib.dptr(sig), n, nd, int(cmvn), _lib.dptr(out), C.byref(got))
        _lib.check(rc, "MFCC.extract")
        assert got.value == rows
        return out


@cached_func
def get_mfcc_extractor(fs, win_length_ms=32, win_shift_ms=16, FFT_SIZE=2048, n_filters=50, n_ceps=13,
                       pre_emphasis_coef=0.95):
    return MFCCExtractor(fs, win_length_ms, win_shift_ms, FFT_SIZE, n_filters, n_ceps, pre_emphasis_coef)


def extract(fs, signal=None, diff=False, **kwargs):
    """accept two argument, or one as a tuple (MFCC.py:123-132).  diff=True appends first differences
    (diff_feature nd=1) exactly as the reference does; pass nd=2 for the 39-dim MFCC+delta+delta-delta."""
    if signal is None:
        assert type(fs) == tuple
        fs, signal = fs[0], fs[1]
    nd = kwargs.pop("nd", 1 if diff else 0)
    return get_mfcc_extractor(fs, **kwargs).extract(signal, nd=nd)
