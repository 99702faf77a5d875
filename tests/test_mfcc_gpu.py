"""GPU parity tests for the fused MFCC path through the C ABI, against the fixtures generated from the reference's
own MFCC.py (tests/golden/mfcc_ref.npz) and the float64 oracle (oracle/mfcc_ref.py).
Tolerance: |ours - ref| <= 1e-4 * max(1, |ref|) per coefficient (fp32 arithmetic vs the float64 reference)."""
import ast

import numpy as np
import pytest

from conftest import rel_err
from oracle import mfcc_ref
from speaker_recognition_b200 import _lib, synth
from speaker_recognition_b200.feature import MFCC, mix_feature
from speaker_recognition_b200.feature.utils import diff_feature

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("name", ["c1_16k", "diff_16k", "d39_16k", "c_8k", "kw_16k"])
def test_extract_matches_reference_fixture(golden_mfcc, name):
    c = golden_mfcc.case(name)
    fs, sig, kw = int(c["fs"]), c["signal"], ast.literal_eval(str(c["kwargs"]))
    if name == "diff_16k":
        got = MFCC.extract((fs, sig), diff=True)
    elif name == "d39_16k":
        got = MFCC.extract(fs, sig, nd=2)
    else:
        got = MFCC.extract(fs, sig, **kw)
    assert got.shape == c["out"].shape and got.dtype == np.float64
    assert np.max(rel_err(got, c["out"])) < TOL


def test_raw_cepstra_and_input_dtypes():
    sig = synth.audio(40000, seed=11)
    tb = mfcc_ref.tables(16000)
    ref_raw = mfcc_ref.raw_cepstra(sig.astype(np.float64), tb)
    ex = MFCC.get_mfcc_extractor(16000)
    assert ex.FRAME_LEN == 512 and ex.FRAME_SHIFT == 256
    raw = ex.extract(sig, cmvn=False)
    assert np.max(rel_err(raw, ref_raw)) < TOL
    ref = mfcc_ref.extract(16000, sig)
    for arr in (sig, sig.astype(np.float32), sig.astype(np.float64), sig.astype(np.int32), list(sig[:20000])):
        got = MFCC.extract(16000, arr)
        r = ref if len(arr) == len(sig) else mfcc_ref.extract(16000, np.asarray(arr))
        assert np.max(rel_err(got, r)) < TOL
    # host-side diff_feature on device output == fused epilogue
    a = diff_feature(MFCC.extract(16000, sig), nd=2)
    b = MFCC.extract(16000, sig, nd=2)
    assert np.max(np.abs(a - b)) < 2e-6


def test_reference_edge_cases():
    with pytest.raises(AssertionError):                      # MFCC.py:56
        MFCC.extract(16000, np.zeros(5 * 512, dtype=np.int16))
    st = np.stack([synth.audio(9000, 5), synth.audio(9000, 6)], axis=1)     # multi-channel input is averaged (:53-55)
    assert np.max(rel_err(MFCC.extract(16000, st), mfcc_ref.extract(16000, st))) < TOL
    # smallest legal signal and a length that is not a multiple of the shift
    for n in (5 * 512 + 1, 7777):
        s = synth.audio(n, seed=n)
        got, ref = MFCC.extract(16000, s), mfcc_ref.extract(16000, s)
        assert got.shape == ref.shape and np.max(rel_err(got, ref)) < 2e-4
    # digital silence inside the signal: the 1e-100 power floor (MFCC.py:7,67) decides those frames
    s = synth.audio(20000, seed=3)
    s[6000:9000] = 0
    raw_ref = mfcc_ref.raw_cepstra(s.astype(np.float64), mfcc_ref.tables(16000))
    raw = MFCC.get_mfcc_extractor(16000).extract(s, cmvn=False)
    assert np.max(rel_err(raw, raw_ref)) < TOL
    assert np.max(rel_err(MFCC.extract(16000, s), mfcc_ref.extract(16000, s))) < TOL
    assert mix_feature((16000, synth.audio(8000, 1))).shape[1] == 13 + 15      # MFCC || LPC


def test_full_size_properties_config2():
    """10 002 raw frames (BASELINE config 2 utterance): properties that need no oracle at this size."""
    n = synth.samples_for_frames(10002)
    sig = synth.audio(n, seed=21)
    f39 = MFCC.extract(16000, sig, nd=2)
    assert f39.shape == (10000, 39)
    f13 = MFCC.extract(16000, sig)
    assert f13.shape == (10002, 13)
    assert np.max(np.abs(f13.mean(axis=0))) < 1e-5 and np.max(np.abs(f13.std(axis=0) - 1)) < 1e-5     # CMVN
    assert np.max(np.abs(f39[:, :13] - f13[2:])) < 1e-6
    assert np.max(np.abs(f39[:, 13:26] - (f13[2:] - f13[1:-1]))) < 1e-5
    # scale invariance: cepstra 1..13 (c0 dropped) and CMVN features do not depend on the signal gain
    half = MFCC.extract(16000, (sig // 2).astype(np.int16) * 2)     # same signal up to LSB
    assert np.max(np.abs(MFCC.extract(16000, sig.astype(np.float32) * 0.25) - f13)) < 2e-4
    # oracle on a 400-frame slice of the same utterance: raw cepstra are local to a frame
    ex = MFCC.get_mfcc_extractor(16000)
    raw = ex.extract(sig, cmvn=False)
    sl = sig[256 * 5000: 256 * 5000 + synth.samples_for_frames(400)]
    ref = mfcc_ref.raw_cepstra(sl.astype(np.float64), mfcc_ref.tables(16000))
    assert np.max(rel_err(raw[5000:5400], ref)) < TOL
    assert half.shape == f13.shape


def test_batched_device_entry_point():
    """srb200_mfcc_extract_dev: several utterances of different lengths in one call, device in / device out"""
    import ctypes as C
    import torch
    L = _lib.lib()
    ex = MFCC.get_mfcc_extractor(16000)
    lens = [30000, 8000, 51234, 2561]
    sigs = [synth.audio(n, seed=100 + i) for i, n in enumerate(lens)]
    pcm = torch.from_numpy(np.concatenate(sigs)).cuda()
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    offc = (C.c_longlong * len(off))(*off.tolist())
    rows = L.srb200_mfcc_rows_total(ex._h, offc, len(lens), 2)
    feat = torch.full((rows, 40), 7.0, dtype=torch.float32, device="cuda")
    foff = (C.c_int * (len(lens) + 1))()
    _lib.check(L.srb200_mfcc_extract_dev(ex._h, pcm.data_ptr(), offc, len(lens), 2, feat.data_ptr(), 40, foff,
                                         torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    out = feat.cpu().numpy()
    assert np.all(out[:, 39] == 0)                               # pad column zero-filled
    for i, s in enumerate(sigs):
        ref = mfcc_ref.extract39(16000, s)
        got = out[foff[i]:foff[i + 1], :39]
        assert got.shape == ref.shape and np.max(rel_err(got, ref)) < TOL
