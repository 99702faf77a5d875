"""Energy-based silence removal with the reference's signature (src/filters/silence.py:11-50), on the device.
Integer arithmetic end to end -> bit-exact with the reference for int16 and uint8 input."""
import ctypes as C

import numpy as np

from .. import _lib


def remove_silence(fs, signal, frame_duration=0.02, frame_shift=0.01, perc=0.15):
    signal = np.ascontiguousarray(signal)
    L = _lib.lib()
    n_out = C.c_longlong(0)
    out = np.empty_like(signal)
    if signal.dtype == np.int16:
        rc = L.srb200_remove_silence_i16(int(fs), signal.ctypes.data_as(C.POINTER(C.c_int16)), len(signal), float(frame_duration),
                                         float(frame_shift), float(perc), out.ctypes.data_as(C.POINTER(C.c_int16)), C.byref(n_out))
    elif signal.dtype == np.uint8:
        rc = L.srb200_remove_silence_u8(int(fs), signal.ctypes.data_as(C.POINTER(C.c_uint8)), len(signal), float(frame_duration),
                                        float(frame_shift), float(perc), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(n_out))
    else:
        raise TypeError("remove_silence supports int16 and uint8 PCM (the wav sample types the reference handles), got %s" % signal.dtype)
    _lib.check(rc, "remove_silence")
    return out[:n_out.value].copy()
