#!/usr/bin/env python
"""BASELINE config 4: 1000 speakers x 256-mixture GMMs, 39-dim, speaker-sharded over the ranks: every rank extracts the
same features and scores them against its S/world speakers; one all-gather of the [B, S_local] sums; global argmax.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P benchmarks/bench_c4.py"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--speakers", type=int, default=1000)
    ap.add_argument("--mixtures", type=int, default=256)
    ap.add_argument("--utts", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    import torch
    import torch.distributed as dist
    from speaker_recognition_b200 import _lib, synth
    from speaker_recognition_b200 import dist as sdist
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    from speaker_recognition_b200.pipeline import BatchScorer
    L = _lib.lib()
    _lib.check(L.srb200_set_device(local_rank))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    S, K, D, B = args.speakers, args.mixtures, 39, args.utts
    lo, hi = sdist.shard_range(S, rank, world)
    gs = GMMSet(gmm_order=K)
    for i in range(lo, hi):
        w, m, s = synth.model(K, D, seed=1000 + i)
        gs.gmms.append(GMM(K).set_params(w, m, s))
        gs.y.append(i)
    n = (10002 - 1) * 256 + 512
    sigs = [synth.audio(n, seed=u) for u in range(min(B, 4))]
    host = torch.empty(B * n, dtype=torch.int16).pin_memory()
    for u in range(B):
        host.numpy()[u * n:(u + 1) * n] = sigs[u % len(sigs)]
    bs = BatchScorer(gs, fs=16000, nd=2)
    bs._prepare([n] * B)
    bs.upload(host)

    def step():
        bs.step_device()
        return sdist.speaker_sharded_argmax(bs.d_sums)

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        best, full = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        frames = B * 10000 * args.steps
        print(json.dumps({"metric": "mfcc_gmm_scored_frames_per_sec", "value": frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
                          "ms_per_step": ms / args.steps, "scaling": "strong",
                          "config": {"workload": "C4", "speakers": S, "mixtures": K, "dim": D, "utts_per_step": B, "frames_per_utt": 10000,
                                     "parallelism": "speaker-sharded x%d (%d speakers per GPU)" % (world, hi - lo)},
                          "frame_model_pairs_per_sec": frames * S / (ms * 1e-3), "argmax_sample": [int(x) for x in best[:4].cpu()],
                          "data": "synthetic"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
