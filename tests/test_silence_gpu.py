"""Bit-exact parity of the device silence removal with the reference algorithm (src/filters/silence.py:11-50)."""
import numpy as np
import pytest

from oracle import silence_ref
from speaker_recognition_b200 import synth
from speaker_recognition_b200.filters import remove_silence

pytestmark = pytest.mark.gpu


def gated(n, seed, fs=16000):
    """audio with silent and quiet stretches"""
    s = synth.audio(n, seed=seed, fs=fs).astype(np.int64)
    rng = np.random.default_rng(seed)
    for _ in range(6):
        w = max(8, min(4000, n // 8))
        a = int(rng.integers(0, n - w))
        s[a:a + int(rng.integers(w // 8, w))] //= int(rng.choice([1000, 40, 7]))
    return s.astype(np.int16)


@pytest.mark.parametrize("fs,n,seed", [(16000, 80000, 1), (8000, 30001, 2), (22050, 50000, 3), (16000, 333, 4), (44100, 200000, 5)])
def test_remove_silence_bit_exact_int16(fs, n, seed):
    sig = gated(n, seed, fs)
    ref = silence_ref.remove_silence(fs, sig)
    got = remove_silence(fs, sig)
    assert got.dtype == np.int16 and got.shape == ref.shape
    assert np.array_equal(got, ref)
    assert 0 < len(ref) < len(sig) or n < 1000


def test_remove_silence_params_uint8_and_extremes():
    sig = gated(60000, 9)
    for kw in (dict(perc=0.5), dict(frame_duration=0.03, frame_shift=0.007), dict(perc=0.0), dict(perc=1e9)):
        assert np.array_equal(remove_silence(16000, sig, **kw), silence_ref.remove_silence(16000, sig, **kw))
    u8 = ((sig.astype(np.int32) >> 8) + 128).astype(np.uint8)
    assert np.array_equal(remove_silence(16000, u8), silence_ref.remove_silence(16000, u8))
    loud = np.full(5000, 32767, dtype=np.int16)             # squares near 2^30, sums need int64
    loud[1000:3000] = -32768
    assert np.array_equal(remove_silence(16000, loud), silence_ref.remove_silence(16000, loud))
    with pytest.raises(TypeError):
        remove_silence(16000, sig.astype(np.float32))


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e", "f", "u8"])
def test_remove_silence_matches_the_reference_outputs(golden_silence, tag):
    """against the outputs of the reference's own src/filters/silence.py (tests/golden/silence_ref.npz), bit for bit"""
    import ast
    c = golden_silence.case(tag)
    got = remove_silence(int(c["fs"]), c["signal"], **ast.literal_eval(str(c["kwargs"])))
    assert got.dtype == c["out"].dtype and np.array_equal(got, c["out"])
