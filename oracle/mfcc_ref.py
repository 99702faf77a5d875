"""oracle/mfcc_ref.py -- TEST INFRASTRUCTURE ONLY.

Float64 NumPy restatement of the reference's own MFCC front-end.  The arithmetic order of the
per-frame path is kept (one numpy.fft.fft of the zero-padded frame per frame, dense mel matrix
product, natural log, DCT product, then whole-utterance CMVN) so this is a float64 statement of
what the reference computes; the table builders are written in vectorised form.

  tables.window        <- hamming                     src/feature/MFCC.py:14-16
  tables (lengths, D)  <- MFCCExtractor.__init__      :20-40   FRAME_LEN = int(float(ms)/1000*fs); D = dctmtx[1:n_ceps+1]
  raw_cepstra          <- MFCCExtractor.extract       :49-71   window, THEN pre-emphasis, |FFT|^2, floor, D.log(M.X)
  cmvn                 <- MFCCExtractor.extract       :72-79   population std (ddof=0), skipped for a single frame
  mel_matrix           <- _mel_filterbank             :81-105  (melfb.m style; `min` at :94 is the builtin)
  dct_matrix           <- dctmtx                      :107-113
  defaults             <- get_mfcc_extractor          :115-121 32 ms / 16 ms / 2048 / 50 / 13 / 0.95
  extract              <- extract                     :123-132
  diff_feature         <- diff_feature                src/feature/utils.py:24-31

Python-2 / old-NumPy behaviour reproduced: integer `/` at :57 and :66 is floor
division; cast['float'] is a float64 cast.  Parity: PINNED against the reference file itself,
executed in this container by tests/golden/make_golden.py (fixtures tests/golden/mfcc_*.npz) and
live in tests/test_oracle_pinning.py when /root/reference is present.
"""
from collections import namedtuple

import numpy as np

POWER_SPECTRUM_FLOOR = 1e-100  # MFCC.py:7

Tables = namedtuple("Tables", "fs frame_len frame_shift nfft window mel dct pre_emph")


def hamming(n):
    """MFCC.py:14-16: half-sample phase, period n (not n-1)."""
    i = np.arange(n, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2 * np.pi / n * (i + 0.5))


def dct_matrix(n):
    """MFCC.py:107-113: orthonormal DCT-II, row y, column x."""
    row = np.arange(n, dtype=np.float64)[:, None]
    col = np.arange(n, dtype=np.float64)[None, :]
    D = np.sqrt(2.0 / n) * np.cos(np.pi * (2 * col + 1) * row / (2 * n))
    D[0] /= np.sqrt(2)
    return D


def mel_matrix(fs, nfft, n_bands):
    """MFCC.py:81-105.  Returns the dense [n_bands, nfft//2+1] matrix."""
    f0 = 700.0 / fs
    half = int(np.floor(nfft / 2))
    lr = np.log(1 + 0.5 / f0) / (n_bands + 1)
    edge = nfft * f0 * (np.exp(np.array([0, 1, n_bands, n_bands + 1]) * lr) - 1)   # :91
    b1 = int(np.floor(edge[0])) + 1
    b2 = int(np.ceil(edge[1]))
    b3 = int(np.floor(edge[2]))
    b4 = min(half, int(np.ceil(edge[3]))) - 1                                       # :94 builtin min
    pos = np.log(1 + np.arange(b1, b4 + 1) / f0 / nfft) / lr                        # :95 fractional band index per bin
    lo = np.floor(pos)
    frac = pos - lo
    M = np.zeros((n_bands, 1 + half))
    c = np.arange(b2 - 1, b4)                                                       # :99-101 falling slopes
    np.add.at(M, (lo[c].astype(int) - 1, c + 1), 2 * (1 - frac[c]))
    c = np.arange(b3)                                                               # :102-104 rising slopes
    np.add.at(M, (lo[c].astype(int), c + 1), 2 * frac[c])
    return M


_cache = {}


def tables(fs, win_length_ms=32, win_shift_ms=16, FFT_SIZE=2048, n_filters=50, n_ceps=13,
           pre_emphasis_coef=0.95):
    """MFCC.py:20-40 + :115-121 (memoised per argument tuple like cached_func, utils.py:11-21)."""
    key = (fs, win_length_ms, win_shift_ms, FFT_SIZE, n_filters, n_ceps, pre_emphasis_coef)
    if key not in _cache:
        flen = int(float(win_length_ms) / 1000 * fs)
        fshift = int(float(win_shift_ms) / 1000 * fs)
        _cache[key] = Tables(fs, flen, fshift, FFT_SIZE, hamming(flen), mel_matrix(fs, FFT_SIZE, n_filters),
                             dct_matrix(n_filters)[1:n_ceps + 1], pre_emphasis_coef)
    return _cache[key]


def num_frames(n_samples, tb):
    return (n_samples - tb.frame_len) // tb.frame_shift + 1                         # :57 (Python 2 int division)


def raw_cepstra(signal, tb):
    """MFCC.py:52-71: cepstra before normalisation, float64 [frames, n_ceps]."""
    if signal.ndim > 1:                                                             # :53-55
        signal = np.mean(signal, axis=1)
    if not len(signal) > 5 * tb.frame_len:                                          # :56
        raise AssertionError("Signal too short!")
    nbins = tb.nfft // 2 + 1
    rows = []
    for f in range(num_frames(len(signal), tb)):
        seg = signal[f * tb.frame_shift: f * tb.frame_shift + tb.frame_len] * tb.window      # :61-62
        seg[1:] -= seg[:-1] * tb.pre_emph              # :64 (RHS evaluated first: non-recursive)
        power = np.abs(np.fft.fft(seg, tb.nfft)[:nbins]) ** 2                                # :66
        power[power < POWER_SPECTRUM_FLOOR] = POWER_SPECTRUM_FLOOR                           # :67
        rows.append(np.dot(tb.dct, np.log(np.dot(tb.mel, power))))                           # :69
    return np.vstack(rows)


def cmvn(feature):
    """MFCC.py:74-77."""
    if feature.shape[0] > 1:
        feature = (feature - np.mean(feature, axis=0)) / np.std(feature, axis=0)
    return feature


def diff_feature(feat, nd=1):
    """utils.py:24-31: plain first differences; rows shrink by nd."""
    d1 = feat[1:] - feat[:-1]
    if nd == 1:
        return np.concatenate((feat[1:], d1), axis=1)
    if nd == 2:
        d2 = d1[1:] - d1[:-1]
        return np.concatenate((feat[2:], d1[1:], d2), axis=1)
    return None                                         # the reference falls off the end for other nd


def extract(fs, signal=None, diff=False, **kwargs):
    """MFCC.py:123-132."""
    if signal is None:
        assert type(fs) == tuple
        fs, signal = fs
    signal = np.asarray(signal, dtype=np.float64)
    out = cmvn(raw_cepstra(signal, tables(fs, **kwargs)))
    return diff_feature(out) if diff else out


def extract39(fs, signal, **kwargs):
    """The 39-dim MFCC+delta+delta-delta features of BASELINE configs 2-4 = extract() then
    diff_feature(nd=2) (extract(diff=True) only reaches nd=1)."""
    return diff_feature(extract(fs, signal, **kwargs), nd=2)
