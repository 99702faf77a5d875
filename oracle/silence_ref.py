"""oracle/silence_ref.py -- TEST INFRASTRUCTURE ONLY.

Restatement of remove_silence, src/filters/silence.py:11-50, under Python 3 semantics that reproduce what the
reference computed under Python 2 (integer `/` at :19 and :49 -> `//`; the debug print at :35 dropped).
Integer arithmetic: the CUDA path must match bit for bit."""
import numpy as np


def remove_silence(fs, signal, frame_duration=0.02, frame_shift=0.01, perc=0.15):
    orig_dtype = type(signal[0])
    typeinfo = np.iinfo(orig_dtype)
    is_unsigned = typeinfo.min >= 0
    signal = signal.astype(np.int64)
    if is_unsigned:
        signal = signal - (typeinfo.max + 1) // 2                   # :19
    siglen = len(signal)
    retsig = np.zeros(siglen, dtype=np.int64)
    frame_length = int(frame_duration * fs)
    frame_shift_length = int(frame_shift * fs)
    new_siglen = 0
    i = 0
    average_energy = np.sum(signal ** 2) / float(siglen)            # :34
    while i < siglen:                                               # :37-46
        subsig = signal[i:i + frame_length]
        ave_energy = np.sum(subsig ** 2) / float(len(subsig))
        if ave_energy < average_energy * perc:
            i += frame_length
        else:
            sigaddlen = min(frame_shift_length, len(subsig))
            retsig[new_siglen:new_siglen + sigaddlen] = subsig[:sigaddlen]
            new_siglen += sigaddlen
            i += frame_shift_length
    retsig = retsig[:new_siglen]
    if is_unsigned:
        retsig = retsig + typeinfo.max // 2                         # :49
    return retsig.astype(orig_dtype)
