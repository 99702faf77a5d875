"""oracle/lpc_ref.py -- TEST INFRASTRUCTURE ONLY.

Float64 restatement of the reference's LPC feature path:

  frames, window, pre-emphasis   <- LPCExtractor.extract      src/feature/LPC.py:46-57 (window from MFCC.hamming)
  lpc(frame, order)[0][1:]       <- LPCExtractor.lpcc         src/feature/LPC.py:40-42 -> scikits.talkbox.linpred.levinson_lpc.lpc

talkbox is third-party, absent from /root/reference and unpinned (Dockerfile:100 installs the latest).  Its published
algorithm (levinson_lpc.py: acorr_lpc = biased autocorrelation via FFT of size 2^nextpow2(2n-1), then levinson_1d) is
restated here; the autocorrelation scale cancels in the coefficients.  Parity status: pinned against the reference
repo's own prototype of the same algorithm, src/lpcc/lpc_jp.py:13-66 (autocorr + LevinsonDurbin), executed by
tests/golden/make_golden.py (fixture tests/golden/lpc_ref.npz); UNPINNED against talkbox itself.
"""
import numpy as np

from . import mfcc_ref


def acorr_biased(x, maxlag):
    """talkbox acorr_lpc: real(ifft(|fft(x, nfft)|^2))[:maxlag+1] / len(x)"""
    n = len(x)
    nfft = 1 << int(np.ceil(np.log2(2 * n - 1)))
    a = np.real(np.fft.ifft(np.abs(np.fft.fft(x, nfft)) ** 2))
    return a[:maxlag + 1] / n


def levinson(r, order):
    """talkbox levinson_1d: returns a[0..order] with a[0] = 1"""
    a = np.zeros(order + 1)
    a[0] = 1.0
    e = r[0]
    for i in range(1, order + 1):
        acc = r[i]
        for j in range(1, i):
            acc += a[j] * r[i - j]
        k = -acc / e
        t = a.copy()
        for j in range(1, i):
            a[j] = t[j] + k * t[i - j]
        a[i] = k
        e *= 1 - k * k
    return a


def extract(fs, signal, win_length_ms=32, win_shift_ms=16, n_lpc=15, pre_emphasis_coef=0.95):
    signal = np.asarray(signal, dtype=np.float64)
    flen = int(float(win_length_ms) / 1000 * fs)
    fshift = int(float(win_shift_ms) / 1000 * fs)
    window = mfcc_ref.hamming(flen)
    frames = (len(signal) - flen) // fshift + 1
    out = []
    for f in range(frames):
        frame = signal[f * fshift: f * fshift + flen] * window
        frame[1:] -= frame[:-1] * pre_emphasis_coef
        r = acorr_biased(frame, n_lpc)
        out.append(levinson(r, n_lpc)[1:] if r[0] != 0 else np.zeros(n_lpc))
    out = np.array(out)
    out[np.isnan(out)] = 0
    return out
