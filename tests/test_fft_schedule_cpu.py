"""The MFCC kernel's FFT schedule (csrc/fft_core.h: Stockham radix-4 passes + radix-2 tail + real-FFT unpack) is
host/device code; replay exactly what the CUDA threads do, serially, and compare with numpy.fft (CPU only)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(ROOT, "tests", "_build", "fft_emul.so")
    src = os.path.join(ROOT, "tests", "csrc", "fft_emul.cc")
    hdr = os.path.join(ROOT, "speaker-recognition_b200", "csrc", "fft_core.h")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-fPIC", "-shared", "-I", os.path.dirname(hdr), "-o", out, src])
    return C.CDLL(out)


@pytest.mark.parametrize("N", [64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("workers", [1, 256])
def test_power_spectrum_matches_numpy(emul, N, workers):
    rng = np.random.default_rng(N)
    batch = 4
    y = np.zeros((batch, N), np.float32)
    y[:, : N // 4] = (rng.standard_normal((batch, N // 4)) * 3000).astype(np.float32)   # zero-padded like MFCC.py:66
    P = np.zeros((batch, N // 2 + 1), np.float32)
    fp = C.POINTER(C.c_float)
    emul.emul_power_spectrum(y.ctypes.data_as(fp), N, batch, workers, P.ctypes.data_as(fp))
    ref = np.abs(np.fft.fft(y.astype(np.float64), axis=1)[:, : N // 2 + 1]) ** 2
    assert np.max(np.abs(P - ref) / ref.max(axis=1, keepdims=True)) < 2e-6


def test_warp_per_frame_power_spectrum_matches_numpy(emul):
    """the warp-per-frame schedule (fft_core.h wf_step_a / wf_step_b_* / wf_power_lane: swizzled transpose tiles, in-place second
    step, bins k and 1024 - k from one pair) replayed lane by lane; every bin 0..1024 must be written exactly once"""
    rng = np.random.default_rng(6)
    batch = 5
    y = np.zeros((batch, 2048), np.float32)
    y[:, :512] = (rng.standard_normal((batch, 512)) * 3000).astype(np.float32)
    y[1, 256:512] = 0           # shorter frame (8 kHz geometry)
    y[2, :] = 0                 # digital silence
    y[3, :512] = 12000.0        # DC: one dominant bin
    P = np.full((batch, 1025), -1.0, np.float32)
    fp = C.POINTER(C.c_float)
    emul.emul_warp_power_spectrum(y.ctypes.data_as(fp), batch, P.ctypes.data_as(fp))
    ref = np.abs(np.fft.fft(y.astype(np.float64), axis=1)[:, :1025]) ** 2
    scale = np.maximum(ref.max(axis=1, keepdims=True), 1e-30)
    assert np.all(P >= 0)
    assert np.max(np.abs(P - ref) / scale) < 2e-6
    assert np.all(P[2] == 0)
