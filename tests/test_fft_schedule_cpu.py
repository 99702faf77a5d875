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


def _wavefronts(addrs, width):
    """shared-memory wavefronts of one warp-wide access: 32 banks of 4 bytes; a 64-bit access is served one half-warp at a time;
    within a phase a bank takes one wavefront per DISTINCT word mapped to it"""
    lanes_per_phase = {4: 32, 8: 16, 16: 8}[width]
    tot = 0
    for p in range(0, 32, lanes_per_phase):
        bank = {}
        for a in addrs[p:p + lanes_per_phase]:
            for w in range(width // 4):
                bank.setdefault((a // 4 + w) % 32, set()).add(a // 4 + w)
        tot += max(len(v) for v in bank.values())
    return tot


def test_warp_per_frame_tiles_are_bank_conflict_free(emul):
    """the transpose tiles of the warp-per-frame FFT (fft_core.h wf_t_index / wf_z_index: padded rows, tile stride 4 mod 16): every
    column write of step A, column read and in-place write of step B and the power stage's Z[k] reads cost exactly the two
    wavefronts a 64-bit access needs; the mirrored reads Z[1024 - k] cost at most four"""
    ti, zi, TS = emul.emul_wf_t_index, emul.emul_wf_z_index, emul.emul_wf_tile()
    assert TS % 16 == 4 and TS >= 16 * 17
    lanes = [(l >> 4, l & 15) for l in range(32)]
    for rnd in range(2):
        for c in range(16):
            assert _wavefronts([8 * ((h + 2 * rnd) * TS + ti(u, c)) for h, u in lanes], 8) == 2          # step A store, column q2 = c
            assert _wavefronts([8 * ((h + 2 * rnd) * TS + ti(c, u)) for h, u in lanes], 8) == 2          # step B load, row n1 = c
            assert _wavefronts([8 * zi(h + 2 * rnd, 16 * c + u) for h, u in lanes], 8) == 2               # step B store, q1 = c
            # in place: a lane stores exactly where it loaded
            for h, u in lanes:
                r = h + 2 * rnd
                assert sorted(zi(r, 16 * q1 + u) for q1 in range(16)) == sorted(r * TS + ti(n1, u) for n1 in range(16))
    for i in range(16):
        ks = [l + 32 * i for l in range(32)]
        assert _wavefronts([8 * zi(k & 3, k >> 2) for k in ks], 8) == 2
        kms = [(1024 - k) & 1023 for k in ks]
        assert _wavefronts([8 * zi(k & 3, k >> 2) for k in kms], 8) <= 4
