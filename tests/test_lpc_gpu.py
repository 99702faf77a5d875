"""GPU parity tests for the LPC half of mix_feature (csrc/lpc.cu) against the fixture generated from the reference repo's
own LPC prototype (tests/golden/lpc_ref.npz) and the float64 oracle (oracle/lpc_ref.py)."""
import ast
import ctypes as C
import os

import numpy as np
import pytest

from conftest import rel_err
from oracle import lpc_ref, mfcc_ref
from speaker_recognition_b200 import _lib, synth
from speaker_recognition_b200.feature import LPC, MFCC, mix_feature

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("case", ["lpc_16k", "lpc_8k_o10"])
def test_lpc_matches_reference_fixture(case):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lpc_ref.npz"))
    kw = ast.literal_eval(str(z[case + "/kwargs"]))
    got = LPC.extract(int(z[case + "/fs"]), z[case + "/signal"], **kw)
    assert got.shape == z[case + "/out"].shape and got.dtype == np.float64
    assert np.max(rel_err(got, z[case + "/out"])) < TOL


def test_lpc_dtypes_diff_and_mix_feature():
    sig = synth.audio(30000, seed=41)
    ref = lpc_ref.extract(16000, sig)
    for arr in (sig, sig.astype(np.float32), sig.astype(np.float64)):
        assert np.max(rel_err(LPC.extract(16000, arr), ref)) < TOL
    d = LPC.extract((16000, sig), diff=True)
    assert d.shape == (ref.shape[0] - 1, 30)
    mix = mix_feature((16000, sig))                      # MFCC || LPC, feature/__init__.py:25-30
    assert mix.shape == (ref.shape[0], 28)
    assert np.max(rel_err(mix[:, :13], mfcc_ref.extract(16000, sig))) < TOL
    assert np.max(rel_err(mix[:, 13:], ref)) < TOL
    # a digitally silent frame: talkbox raises (r[0] == 0); here the row is zeros and the rest is unaffected
    s2 = sig.copy()
    s2[5000:7000] = 0
    got = LPC.extract(16000, s2)
    assert np.all(np.isfinite(got))
    ref2 = lpc_ref.extract(16000, s2)
    assert np.max(rel_err(got, ref2)) < TOL


def test_lpc_device_entry_point_builds_mix_feature_in_place():
    import torch
    L = _lib.lib()
    ex = MFCC.get_mfcc_extractor(16000)
    lens = [20000, 9000, 33000]
    sigs = [synth.audio(n, seed=500 + i) for i, n in enumerate(lens)]
    pcm = torch.from_numpy(np.concatenate(sigs)).cuda()
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    offc = (C.c_longlong * len(off))(*off.tolist())
    rows = L.srb200_mfcc_rows_total(ex._h, offc, len(lens), 0)
    feat = torch.zeros((rows, 28), dtype=torch.float32, device="cuda")
    foff = (C.c_int * (len(lens) + 1))()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.srb200_mfcc_extract_dev(ex._h, pcm.data_ptr(), offc, len(lens), 0, feat.data_ptr(), 28, foff, st))
    _lib.check(L.srb200_lpc_extract_dev(ex._h, pcm.data_ptr(), offc, len(lens), 15, feat.data_ptr(), 28, 13, foff, st))
    torch.cuda.synchronize()
    out = feat.cpu().numpy()
    for i, s in enumerate(sigs):
        ref = np.concatenate((mfcc_ref.extract(16000, s), lpc_ref.extract(16000, s)), axis=1)
        assert np.max(rel_err(out[foff[i]:foff[i + 1]], ref)) < TOL
