"""CPU emulation of the tensor-core scorer from the operands the HOST PACKER produces (srb200_tc_pack_export): the fp16 head/tail
weight images are read back exactly as the kernel addresses them (block, CTA rank = column >> 6, canonical core-matrix offset),
the exactly corrected entries are applied from the per-block lists (prefix-ordered columns, scratch slots, null padding), the
valid-column masks drop speaker padding -- and the resulting log-likelihoods must equal the float64 oracle.  This pins the
packer's layout and bookkeeping without a GPU; the arithmetic of the kernel itself is covered by tests/test_gmm_tc_gpu.py."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err, shipped_ubm_pools
from oracle import gmm_ref
from speaker_recognition_b200 import _lib, synth
from speaker_recognition_b200.gmm.pygmm import GMM

NH, NB = 64, 128


def export(models):
    gs = [GMM(len(w)).set_params(w, m, s) for (w, m, s) in models]
    arr = (C.c_void_p * len(gs))(*[g.gmm for g in gs])
    info = (C.c_int * 8)()
    L = _lib.lib()
    assert L.srb200_tc_pack_export(arr, len(gs), info, None, None, None) == 1
    Kp, blocks, cols, bh, hs, hhead = info[0], info[1], info[2], info[3], info[4], info[5]
    D = models[0][1].shape[1]
    img = np.zeros(blocks * bh, dtype=np.uint16)
    hard = np.zeros(blocks * hs, dtype=np.int32)
    cs = np.zeros(2 * D, dtype=np.float32)
    assert L.srb200_tc_pack_export(arr, len(gs), info, img.ctypes.data_as(C.POINTER(C.c_uint16)), hard.ctypes.data_as(C.POINTER(C.c_int)),
                                   cs.ctypes.data_as(C.POINTER(C.c_float))) == 1
    return dict(Kp=Kp, blocks=blocks, cols=cols, bh=bh, hs=hs, hhead=hhead, D=D, img=img.view(np.float16), hard=hard.reshape(blocks, hs // 4, 4), cs=cs)


def emulate(P, X):
    """per-frame log2-likelihood matrix [T, blocks * 128] and the valid mask, as the kernel computes them"""
    D, Kp = P["D"], P["Kp"]
    x = ((X.astype(np.float32) - P["cs"][:D]) * P["cs"][D:]).astype(np.float32)
    phi = np.zeros((len(X), Kp), dtype=np.float64)
    phi[:, :D], phi[:, D:2 * D], phi[:, 2 * D] = (x * x).astype(np.float64), x, 1.0
    k = np.arange(Kp)
    L2 = np.zeros((len(X), P["blocks"] * NB))
    valid = np.zeros(P["blocks"] * NB, dtype=bool)
    for b in range(P["blocks"]):
        blk = P["img"][b * P["bh"]:(b + 1) * P["bh"]].astype(np.float64)
        W = np.zeros((NB, Kp))
        for col in range(NB):
            r = col & 63
            off = (col >> 6) * 2 * NH * Kp + ((k >> 3) * (NH // 8) + (r >> 3)) * 64 + (r & 7) * 8 + (k & 7)
            W[col] = blk[off] + blk[off + NH * Kp]                    # head + tail
        l = phi @ W.T
        hl = P["hard"][b]
        starts, ends, nflag, vmask = hl[:4, 0], hl[:4, 1], hl[:4, 2], hl[:4, 3]
        for g in range(4):
            assert nflag[g] % 4 == 0 and (ends[g] - starts[g]) % 4 == 0
            scratch = np.full((len(X), 16), np.nan)                    # stale slots must never be read
            csum = np.zeros(len(X))
            for e in range(starts[g], ends[g]):
                xo, a_bits, nb_bits, w = hl[P["hhead"] + e]
                a = np.array([a_bits], dtype=np.int32).view(np.float32)[0]
                nb = np.array([nb_bits], dtype=np.int32).view(np.float32)[0]
                t = x[:, xo // 4].astype(np.float64) * float(a) + float(nb)
                csum = csum + t * t
                if w:
                    assert (w - 1) % 4 == 0 and 0 <= (w - 1) // 4 < 16
                    scratch[:, (w - 1) // 4] = csum
                    csum = np.zeros(len(X))
            for p in range(nflag[g]):
                assert not np.isnan(scratch[:, p]).any(), "flagged slot never flushed"
                l[:, g * 32 + p] -= scratch[:, p]
            for p in range(32):
                valid[b * NB + g * 32 + p] = bool((int(vmask[g]) >> p) & 1)
        L2[:, b * NB:(b + 1) * NB] = l
    return L2, valid


def speaker_ll(P, L2, valid, s):
    cols = P["cols"]
    l = np.where(valid[s * cols:(s + 1) * cols][None, :], L2[:, s * cols:(s + 1) * cols], -np.inf)
    m = l.max(1, keepdims=True)
    return (m[:, 0] + np.log2(np.exp2(l - m).sum(1))) * np.log(2.0)


@pytest.mark.parametrize("S,K,D", [(3, 128, 39), (5, 32, 13), (3, 48, 28), (2, 200, 34), (6, 20, 39)])
def test_packed_operands_reproduce_the_oracle_synthetic(S, K, D):
    W, M, Sg = synth.models(S, K, D, seed0=4100 + K, floor_frac=0.03)
    P = export([(W[i], M[i], Sg[i]) for i in range(S)])
    assert P["cols"] == (32 if K <= 32 else 64 if K <= 64 else -(-K // 128) * 128) and P["Kp"] == -(-(2 * D + 1) // 16) * 16
    X = synth.draw(W[S // 2], M[S // 2], Sg[S // 2], 200, seed=K)
    L2, valid = emulate(P, X)
    assert valid.sum() == S * K
    for i in range(S):
        exact = gmm_ref.score(W[i], M[i], Sg[i], X, gmm_ref.MODE_LOGDOMAIN)
        assert np.max(rel_err(speaker_ll(P, L2, valid, i), exact)) < 2e-5


def test_packed_operands_reproduce_the_oracle_real_parameters(golden_gmm, golden_real):
    pools = shipped_ubm_pools(golden_gmm, golden_real)
    for w, m, s in pools:
        P = export([(w, m, s)])
        X = synth.draw(w, m, s, 150, seed=len(w))
        L2, valid = emulate(P, X)
        assert np.max(rel_err(speaker_ll(P, L2, valid, 0), gmm_ref.score(w, m, s, X, gmm_ref.MODE_LOGDOMAIN))) < 2e-5
    W, M, Sg = synth.realparam_models(12, 32, pools, seed0=5000)
    P = export([(W[i], M[i], Sg[i]) for i in range(12)])
    X = synth.draw(W[7], M[7], Sg[7], 150, seed=9)
    L2, valid = emulate(P, X)
    for i in range(12):
        assert np.max(rel_err(speaker_ll(P, L2, valid, i), gmm_ref.score(W[i], M[i], Sg[i], X, gmm_ref.MODE_LOGDOMAIN))) < 2e-5
