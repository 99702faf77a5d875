#!/usr/bin/env python
"""BASELINE config 3: UBM EM, K=512 mixtures, D=39, N=1e6 synthetic frames, frame-sharded over the ranks with ONE
all-reduce of the float64 statistics per iteration (strong scaling: N is fixed).

  python benchmarks/bench_em.py [--frames 1000000] [--iters 10]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P benchmarks/bench_em.py

Prints one JSON line: frames/s per EM iteration (= N * iters / time, device-timed, max over ranks), the time of the
all-reduce alone, roofline numbers for the E-step kernels, and (rank 0, N=1) the reference's own C++ trainer on a
bounded sample of the same data as the CPU baseline."""
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
    ap.add_argument("--frames", type=int, default=1000000)
    ap.add_argument("--mixtures", type=int, default=512)
    ap.add_argument("--dim", type=int, default=39)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))

    import torch
    import torch.distributed as dist
    from speaker_recognition_b200 import _lib, synth
    from speaker_recognition_b200 import dist as sdist
    from speaker_recognition_b200.gmm.pygmm import GMM
    L = _lib.lib()
    _lib.check(L.srb200_set_device(local_rank))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N, K, D = args.frames, args.mixtures, args.dim
    # data: frames from a seeded K-component mixture, generated on the device (same on every rank, then sharded)
    g = torch.Generator(device="cuda").manual_seed(1234)
    wt, mt, st_ = synth.model(K, D, seed=77, floor_frac=0.0)
    comp = torch.multinomial(torch.from_numpy(wt).cuda(), N, replacement=True, generator=g)
    X = torch.from_numpy(mt).cuda()[comp] + torch.from_numpy(st_).cuda()[comp] * torch.randn((N, D), device="cuda", dtype=torch.float64, generator=g)
    lo, hi = sdist.shard_range(N, rank, world)
    ld = (D + 3) & ~3
    Xd = torch.zeros((hi - lo, ld), dtype=torch.float32, device="cuda")
    Xd[:, :D] = X[lo:hi].float()
    # injected init with init_gaussians semantics (gmm.cc:306-361): random rows, global unbiased std, uniform weights
    idx = torch.randint(0, N, (K,), generator=torch.Generator().manual_seed(5))
    m0 = X[idx.cuda()].cpu().numpy()
    s0 = np.tile(X.std(dim=0, unbiased=True).cpu().numpy(), (K, 1))
    w0 = np.ones(K) / K
    Xcpu = X[:args.cpu_sample].cpu().numpy() if rank == 0 else None
    del X, comp

    def run(iters):
        gm = GMM(K).set_params(w0, m0, s0)
        be = sdist.DeviceEM(gm, Xd, hi - lo, ld, D)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done, ll = sdist.em_train(be, N, iters, threshold=0.0)
        e1.record()
        torch.cuda.synchronize()
        return gm, done, ll, e0.elapsed_time(e1)

    run(2)                                   # warm-up (allocations, NCCL channels)
    gm, done, ll, ms = run(args.iters)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # all-reduce alone
    stats = torch.zeros(L.srb200_em_stats_len(K, D), dtype=torch.float64, device="cuda")
    ar_us = 0.0
    if world > 1:
        for _ in range(5):
            dist.all_reduce(stats)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            dist.all_reduce(stats)
        e1.record()
        torch.cuda.synchronize()
        ar_us = e0.elapsed_time(e1) / 50 * 1e3
    if rank == 0:
        per_iter_ms = ms / (done + 0.5)       # `done` M-steps + the extra monitoring E-steps (about half an iteration each 2)
        val = N * done / (ms * 1e-3)
        fma_issued = 6.25 * 2 * D * N * K       # (2 + 2 + 2.25) FMA-pipe ops per (n, k, d) counted as 2 flops
        line = {"metric": "ubm_em_frames_per_sec_per_iteration", "value": val, "unit": "frames/s", "n_gpus": world, "iters": done,
                "ms_total": ms, "ms_per_iteration": ms / max(done, 1), "scaling": "strong",
                "config": {"workload": "C3", "mixtures": K, "dim": D, "frames": N, "parallelism": "frame-sharded x%d" % world},
                "allreduce_us": ar_us, "allreduce_bytes": int(stats.numel() * 8), "final_ll": ll,
                "roofline_fma": {"achieved_tflops": fma_issued * done / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                 "note": "6.25 fp32 pipe ops per (frame, mixture, dim) per iteration incl. the monitoring passes"},
                "dtype": "f32 kernels, f64 statistics / M-step", "data": "synthetic"}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref_so
            if ref_so.available():
                ns = len(Xcpu)
                gref = ref_so.RefGMM(w0, m0, s0)
                threads = os.cpu_count() or 1
                t0 = time.perf_counter()
                for _ in range(2):
                    gref.em_iteration(Xcpu, 1e-3, threads)
                dt = (time.perf_counter() - t0) / 2
                line["cpu_baseline"] = {"value": ns / dt, "unit": "frames/s", "cores": threads, "kind": "reference",
                                        "sample": "GMMTrainerBaseline::iteration (compiled reference) on %d of the frames, concurrency=%d" % (ns, threads)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
