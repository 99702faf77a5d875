#!/usr/bin/env python
"""BASELINE config 5: 64 concurrent 16 kHz streams, 25 ms hop (400 samples), trailing 1.5 s window, 100 speakers x 128
mixtures, 39-dim features; reports p50/p99 per-tick latency from "new samples on the host" to "labels on the host"."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--ticks", type=int, default=400)
    ap.add_argument("--speakers", type=int, default=100)
    ap.add_argument("--mixtures", type=int, default=128)
    args = ap.parse_args()
    import torch
    from speaker_recognition_b200 import synth
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    from speaker_recognition_b200.streaming import StreamingPredictor
    S, K, D, B, hop = args.speakers, args.mixtures, 39, args.streams, 400
    W, M, Sg = synth.models(S, K, D, seed0=1000)
    gs = GMMSet(gmm_order=K)
    for i in range(S):
        gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
        gs.y.append(i)
    sp = StreamingPredictor(gs, n_streams=B, fs=16000, window_s=1.5, nd=2)
    fill = 24000 // hop
    total = fill + args.ticks + 20
    audio = np.stack([synth.audio(hop * total, seed=b) for b in range(B)])
    lat = []
    for t in range(total):
        chunk = audio[:, t * hop:(t + 1) * hop]
        t0 = time.perf_counter()
        out = sp.push(chunk)
        dt = time.perf_counter() - t0
        if out is not None and t >= fill + 20:
            lat.append(dt * 1e3)
    lat = np.array(lat)
    print(json.dumps({"metric": "stream_tick_latency_ms", "p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                      "mean": float(lat.mean()), "ticks": len(lat), "unit": "ms", "higher_is_better": False,
                      "config": {"workload": "C5", "streams": B, "hop_samples": hop, "window_s": 1.5, "frames_per_window": 90,
                                 "speakers": S, "mixtures": K, "dim": D},
                      "realtime_budget_ms": 25.0, "data": "synthetic"}), flush=True)


if __name__ == "__main__":
    main()
