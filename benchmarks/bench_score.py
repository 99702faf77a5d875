"""Scorer-only timing (device-resident features drawn from the models): one launch of srb200_gmmset_score_dev per step, CUDA
events on the launching stream.  Usage: python benchmarks/bench_score.py [--path 1] [--S 100 --K 128 --D 39 --utts 64 --frames 10000]
[--real]  (--real: parameters resampled from the shipped UBMs, D = 34; needs tests/golden/*.npz)"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path", type=int, default=-1)
    ap.add_argument("--S", type=int, default=100)
    ap.add_argument("--K", type=int, default=128)
    ap.add_argument("--D", type=int, default=39)
    ap.add_argument("--utts", type=int, default=64)
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--real", action="store_true")
    a = ap.parse_args()
    import torch
    from speaker_recognition_b200 import _lib, synth
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    L = _lib.lib()
    if a.real:
        zg = np.load(os.path.join(ROOT, "tests", "golden", "gmm_ref.npz"))
        zr = np.load(os.path.join(ROOT, "tests", "golden", "realparams_ref.npz"))
        pools = [(zg[t + "/w"], zg[t + "/mean"], zg[t + "/sigma"]) for t in ("ubm32", "ubm64", "ubm256")]
        pools.append((zr["ubm32p/w"], zr["ubm32p/mean"], zr["ubm32p/sigma"]))
        W, M, Sg = synth.realparam_models(a.S, a.K, pools, seed0=6000)
        a.D = M.shape[2]
    else:
        W, M, Sg = synth.models(a.S, a.K, a.D, seed0=1000)
    gs = GMMSet(gmm_order=a.K)
    for i in range(a.S):
        gs.gmms.append(GMM(a.K).set_params(W[i], M[i], Sg[i]))
        gs.y.append(i)
    ld = (a.D + 3) & ~3
    n = a.utts * a.frames
    X = torch.zeros((n, ld), dtype=torch.float32, device="cuda")
    truth = []
    for u in range(a.utts):
        s = (u * 7) % a.S
        truth.append(s)
        X[u * a.frames:(u + 1) * a.frames, :a.D] = torch.from_numpy(synth.draw(W[s], M[s], Sg[s], a.frames, seed=100 + u).astype(np.float32)).cuda()
    off = (C.c_int * (a.utts + 1))(*[u * a.frames for u in range(a.utts + 1)])
    sums = torch.zeros((a.utts, a.S), dtype=torch.float64, device="cuda")
    best = torch.zeros(a.utts, dtype=torch.int32, device="cuda")
    _lib.check(L.srb200_set_score_path(a.path))
    st = torch.cuda.current_stream().cuda_stream
    h = gs._ensure_set()

    def step():
        _lib.check(L.srb200_gmmset_score_dev(h, X.data_ptr(), ld, a.D, off, a.utts, sums.data_ptr(), None, 0, best.data_ptr(), st))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for e0, e1 in ev:
        e0.record()
        step()
        e1.record()
    torch.cuda.synchronize()
    ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    flops = 3.0 * a.D * n * a.S * a.K
    print(json.dumps({"bench": "score_only", "path": a.path, "route": int(L.srb200_last_score_route()), "real_params": bool(a.real), "S": a.S, "K": a.K, "D": a.D,
                      "utts": a.utts, "frames": a.frames, "ms_median": float(np.median(ms)), "ms_min": float(np.min(ms)),
                      "algorithmic_tflops": flops / (np.median(ms) * 1e-3) / 1e12,
                      "frames_per_s": n / (np.median(ms) * 1e-3), "argmax_ok": bool(np.array_equal(best.cpu().numpy(), np.array(truth)))}))


if __name__ == "__main__":
    main()
