#!/usr/bin/env python
"""bench.py -- the headline benchmark of the MFCC + GMM hot path (BASELINE.json metric: scored frames/sec;
GMM kernel vs roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--utts B] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload = "C2"): BASELINE.json configs[1] -- 100 speakers x 128-mixture diagonal GMMs, 39-dim
MFCC+delta+delta-delta features, utterances of 10 000 scored frames (10 002 raw frames = 2 560 768 int16 samples
@16 kHz); one STEP = one pass of the whole hot path (PCM -> MFCC -> CMVN/deltas -> score vs all speakers ->
per-utterance argmax) over a batch of B such utterances per GPU.  A "scored frame" is one feature frame evaluated
against all 100 enrolled models.  Data: synthetic seeded audio, synthetic seeded models (speaker_recognition_b200.synth).

  value      whole-job frames/s with the PCM batch already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e        the same metric through the public batched API (pipeline.BatchScorer.predict_pcm): pinned host PCM ->
             H2D -> kernels -> D2H of the labels, every step
  roofline   the dominant kernel (gmm_score_kernel) timed alone inside this run with CUDA events; reported against
             HBM (algorithmic bytes, the contract's field) AND against the fp32 FMA pipe, which is the binding bound
  cpu_baseline / --impl reference   the reference's CPU path on this host: its MFCC restated in NumPy (its own is
             Python 2) + its own C++ GMM compiled from /root/reference (oracle/_ref), all host threads
Multi-GPU: utterances are independent -> each rank scores its own B utterances (weak scaling), no data-path collective.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S_SPK, K_MIX, N_CEPS, ND = 100, 128, 13, 2
DIM = N_CEPS * (ND + 1)
FRAMES_OUT = 10000
RAW_FRAMES = FRAMES_OUT + ND
UTT_SAMPLES = (RAW_FRAMES - 1) * 256 + 512          # 2 560 768
L2_BYTES = 126 * 1024 * 1024


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def wait_started(self, timeout=5.0):
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=2)

    def summary(self, windows):
        """only samples taken inside the timed windows [(t0, t1), ...] (host clock) count"""
        sm, smax, reasons = [], [], set()
        for ts, r in self.rows:
            if not any(a <= ts <= b for a, b in windows):
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference's CPU path on this host (checker code: oracle/ is used here only as the CPU baseline)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(models, sig, threads):
    """one C2 utterance through the reference's CPU path: MFCC (float64 NumPy restatement of src/feature/MFCC.py, per-frame
    loop kept) + diff_feature(nd=2) + its own compiled C++ scorer (score_all per speaker, concurrency = all threads)"""
    from oracle import mfcc_ref
    t0 = time.perf_counter()
    feat = mfcc_ref.extract39(16000, sig)
    t1 = time.perf_counter()
    scores = [m.score_all_abi(feat, threads) if hasattr(m, "score_all_abi") else m(feat, threads) for m in models]
    t2 = time.perf_counter()
    return len(feat), t1 - t0, t2 - t1, int(np.argmax(scores))


def cpu_models(W, M, Sg):
    """reference GMM handles (compiled reference when oracle/_ref exists, else the C restatement)"""
    from oracle import gmm_ref, ref_so
    if ref_so.available():
        return [ref_so.RefGMM(W[i], M[i], Sg[i]) for i in range(len(W))], "reference"
    return [(lambda X, t, i=i: gmm_ref.score_all(W[i], M[i], Sg[i], X, gmm_ref.MODE_FASTEXP, t)) for i in range(len(W))], "port"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, bounded sample per step"""
    if rank != 0:
        return
    from speaker_recognition_b200 import synth
    threads = os.cpu_count() or 1
    W, M, Sg = synth.models(S_SPK, K_MIX, DIM, seed0=1000)
    models, kind = cpu_models(W, M, Sg)
    sig = synth.audio(UTT_SAMPLES, seed=0)
    for _ in range(args.warmup):
        cpu_reference_step(models, sig, threads)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        n, _, _, _ = cpu_reference_step(models, sig, threads)
        frames += n
    dt = time.perf_counter() - t0
    val = frames / dt
    sample = "1 utterance (%d frames) x %d speakers per step; MFCC = float64 NumPy restatement (1 core), scoring = %s C++ score_all x %d speakers, concurrency=%d" % (
        FRAMES_OUT, S_SPK, "compiled reference" if kind == "reference" else "C port of the reference", S_SPK, threads)
    line = {"impl": "reference", "metric": "mfcc_gmm_scored_frames_per_sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2", "speakers": S_SPK, "mixtures": K_MIX, "dim": DIM, "frames_per_utt": FRAMES_OUT,
                       "utts_per_step": 1, "l2_policy": "cpu"},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--utts", type=int, default=64, help="utterances per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from speaker_recognition_b200 import _lib, synth
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    from speaker_recognition_b200.pipeline import BatchScorer

    L = _lib.lib()
    _lib.check(L.srb200_set_device(local_rank), "set_device")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic speakers and audio (per rank: its own utterances -> weak scaling) ----
    W, M, Sg = synth.models(S_SPK, K_MIX, DIM, seed0=1000)
    gs = GMMSet(gmm_order=K_MIX)
    for i in range(S_SPK):
        gs.gmms.append(GMM(K_MIX).set_params(W[i], M[i], Sg[i]))
        gs.y.append(i)
    B = args.utts
    n_distinct = min(B, 8)
    base = [synth.audio(UTT_SAMPLES, seed=rank * 1000 + u) for u in range(n_distinct)]
    host = torch.empty(B * UTT_SAMPLES, dtype=torch.int16).pin_memory()
    hv = host.numpy()
    for u in range(B):
        hv[u * UTT_SAMPLES:(u + 1) * UTT_SAMPLES] = base[u % n_distinct]
    lengths = [UTT_SAMPLES] * B
    bs = BatchScorer(gs, fs=16000, nd=ND)
    bs._prepare(lengths)
    frames_per_step = bs.frames_per_step()
    assert frames_per_step == B * FRAMES_OUT

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ----
    bs.upload(host)
    for _ in range(args.warmup):
        bs.step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_started()
    barrier()
    w0 = time.perf_counter()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    f_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    for i in range(args.steps):
        f_ev[i][0].record()
        bs.features()
        f_ev[i][1].record()
        k_ev[i][0].record()
        bs.score()
        k_ev[i][1].record()
    ev[1].record()
    barrier()
    w1 = time.perf_counter()
    dev_ms = ev[0].elapsed_time(ev[1])
    score_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    feat_ms = float(np.mean([a.elapsed_time(b) for a, b in f_ev]))
    labels_dev = bs.d_best.cpu().numpy().copy()

    # ---- end-to-end through the public batched API: pinned host PCM -> labels on host, every step ----
    for _ in range(2):
        bs.predict_pcm(host, lengths)
    barrier()
    w2 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lab = bs.predict_pcm(host, lengths)
    e1.record()
    barrier()
    w3 = time.perf_counter()
    e2e_ms = e0.elapsed_time(e1)
    time.sleep(0.12)
    sampler.stop()
    clocks = sampler.summary([(w0, w1 + 0.06), (w2, w3 + 0.06)])
    assert np.array_equal(lab.numpy(), labels_dev)

    t = torch.tensor([dev_ms, e2e_ms, score_ms, feat_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, score_ms, feat_ms = [float(x) for x in t.cpu()]
    total_frames = frames_per_step * args.steps * world
    value = total_frames / (dev_ms * 1e-3)
    e2e_value = total_frames / (e2e_ms * 1e-3)

    if rank == 0:
        hbm_peak, peak_src, peaks = load_peaks()
        # algorithmic bytes/flops of ONE scoring launch: B_s = 4 T D + 4 S K (2D+1) + 8 S per utterance batch
        T = frames_per_step
        bytes_alg = 4.0 * T * DIM + 4.0 * S_SPK * K_MIX * (2 * DIM + 1) + 8.0 * S_SPK * B
        flops_alg = 3.0 * DIM * T * S_SPK * K_MIX                   # survey convention: sub, mul, mul-add per dim
        fma_issued = 2.0 * DIM * T * S_SPK * K_MIX                   # what the kernel issues: 2 FFMA per (frame, mixture, dim)
        sec = score_ms * 1e-3
        route = int(L.srb200_last_score_route())          # 1: tcgen05 3xTF32 kernel (default), 0: fp32 FMA kernel
        kname = "gmm_score_tc_kernel" if route == 1 else "gmm_score_kernel"
        traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tpath):
            tj = json.load(open(tpath)).get(kname, {})
            if tj.get("utts") == B and tj.get("speakers") == S_SPK and tj.get("mixtures") == K_MIX:
                traffic = tj.get("dram_bytes")
        hbm = {"bound": "hbm", "kernel": kname, "achieved": bytes_alg / sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "frac": bytes_alg / sec / 1e9 / hbm_peak, "traffic": traffic, "peak_source": peak_src,
               "ms_per_launch": score_ms, "algorithmic_bytes": bytes_alg,
               "note": "scoring is compute-bound (~2700 flop/B at C2), not HBM-bound: see the binding roofline"}
        fma = None
        if route == 0:
            fma = {"bound": "fp32_fma", "kernel": "gmm_score_kernel", "unit": "TFLOP/s",
                   "achieved": 2.0 * fma_issued / sec / 1e12, "achieved_survey_convention": flops_alg / sec / 1e12}
            try:
                from speaker_recognition_b200 import build as b
                LB = C.CDLL(b.BENCH_OUT)
                LB.srb200_bench_fma_tflops.restype = C.c_double
                LB.srb200_bench_fma_tflops.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
                pk = {name: LB.srb200_bench_fma_tflops(v, 4000, 3, None) for v, name in ((0, "ffma_indep"), (1, "ffma_pair"), (2, "ffma2_packed"))}
                fma["peak"] = max(pk.values())
                fma["peak_variants"] = pk
                fma["peak_source"] = "measured in this run (csrc/bench/fma_peak.cu)"
                fma["frac"] = fma["achieved"] / fma["peak"]
            except Exception as e:  # measurement helper missing: nominal figure, flagged
                fma["peak"] = 148 * 128 * 2 * 1.965e9 / 1e12
                fma["peak_source"] = "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (helper failed: %s)" % e
                fma["frac"] = fma["achieved"] / fma["peak"]
        if route == 1:
            # tensor route: the Mahalanobis sum as a GEMM [T x (2D+1)] . [(2D+1) x S K].  ALGORITHMIC flops = one pass over the
            # unpadded K; the kernel executes 3 tf32 passes (3xTF32 split for fp32-grade accuracy) over
            # K padded to a multiple of 8, reported as `executed`.  Peak: dense tf32 = half the measured cuBLAS bf16 figure
            # (tensor-core rate halves from 16-bit to tf32 operands: 2.25 vs 1.1 PFLOP/s nominal, B200_PROFILING.md) --
            # the sustained one, because the kernel is timed inside the step loop.
            gemm_alg = 2.0 * (2 * DIM + 1) * T * S_SPK * K_MIX
            kp = ((2 * DIM + 1 + 7) // 8) * 8
            gemm_exec = 3.0 * 2.0 * kp * T * S_SPK * K_MIX
            if peaks.get("bf16_tflops_sustained"):
                tpeak, tsrc = peaks["bf16_tflops_sustained"] / 2.0, "measured bf16_tflops_sustained / 2 (MEASURED_PEAKS.json; tf32 = half the bf16 rate)"
            else:
                tpeak, tsrc = 1100.0, "fallback: nominal dense tf32 1.1 PFLOP/s (B200_PROFILING.md)"
            roofline = {"bound": "tensor", "kernel": kname, "achieved": gemm_alg / sec / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                        "frac": gemm_alg / sec / 1e12 / tpeak, "traffic": traffic, "peak_source": tsrc, "ms_per_launch": score_ms,
                        "algorithmic_flops": gemm_alg, "executed": gemm_exec / sec / 1e12, "executed_frac": gemm_exec / sec / 1e12 / tpeak,
                        "note": "3xTF32: the kernel issues 3 tensor passes per algorithmic pass, so frac <= 1/3 by construction; "
                                "executed_frac is the tensor-pipe utilisation against the same peak"}
            binding = "tensor_tf32"
        else:
            roofline = hbm
            binding = "fp32_fma"
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            models, kind = cpu_models(W, M, Sg)
            nfr, t_m, t_s, cpu_lab = cpu_reference_step(models, base[0], threads)
            cpu_baseline = {"value": nfr / (t_m + t_s), "unit": "frames/s", "cores": threads, "kind": kind,
                            "sample": "1 of the %d utterances of a step (%d frames x %d speakers): MFCC restatement %.2f s (1 core) + %s scorer %.2f s (%d threads)"
                                      % (B, nfr, S_SPK, t_m, kind, t_s, threads),
                            "mfcc_s": t_m, "score_s": t_s, "top1_matches_gpu": bool(cpu_lab == int(labels_dev[0]))}
        line = {"metric": "mfcc_gmm_scored_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "tf32x3+f32" if route == 1 else "f32", "data": "synthetic",
                "config": {"workload": "C2", "speakers": S_SPK, "mixtures": K_MIX, "dim": DIM, "frames_per_utt": FRAMES_OUT,
                           "utts_per_step_per_gpu": B, "parallelism": "utterance-sharded x%d" % world,
                           "l2_policy": "inputs larger than L2 (%.0f MB PCM per step per GPU vs %d MB L2)" % (B * UTT_SAMPLES * 2 / 1e6, L2_BYTES >> 20)},
                "frame_model_pairs_per_sec": value * S_SPK,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(B * UTT_SAMPLES * 2), "d2h_bytes_per_step": int(B * 4),
                        "ms_per_step": e2e_ms / args.steps, "api": "pipeline.BatchScorer.predict_pcm"},
                "gpu_launches": BatchScorer.KERNELS_PER_STEP * args.steps,
                "kernel_ms": {"mfcc_features": feat_ms, "gmm_score": score_ms},
                "score_route": "tcgen05_3xtf32" if route == 1 else "fp32_fma",
                "roofline": roofline, "roofline_hbm": hbm if route == 1 else None, "roofline_fma": fma if route == 0 else None,
                "binding_roofline": binding,
                "cpu_baseline": cpu_baseline, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
