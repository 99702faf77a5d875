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
  roofline   the dominant kernel (the tcgen05 scorer) timed alone inside this run with CUDA events: ALGORITHMIC flops
             3 D T S K / that time, against the measured fp16/bf16 tensor peak (MEASURED_PEAKS.json) -- the unit
             the kernel runs on; roofline_mfcc the same for the MFCC kernels against the measured fp32 FMA peak; roofline_hbm
             the contract's HBM line (this path is compute-bound: ~2700 flop/B)
  cpu_baseline / --impl reference   the reference's CPU path on this host: its MFCC restated in NumPy (its own is Python 2) +
             its own C++ GMM compiled from /root/reference (oracle/_ref); cpu_baseline.variants lists the compositions
             BASELINE.md s3 names (concurrency sweep, marshalling, sklearn, MFCC over a process pool), value = the fastest
  c2_realparams / c3_em / c4   sub-records for the other BASELINE configs, measured after the headline timing: the scorer on
             models with the parameter statistics of the UBMs the reference ships; frame-sharded UBM EM with its one
             all-reduce per iteration (strong scaling over --gpus); 1000 speakers x 256 mixtures speaker-sharded
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
def workload_config():
    """the SAME dict in both arms (the driver compares them)"""
    return {"workload": "C2", "speakers": S_SPK, "mixtures": K_MIX, "dim": DIM, "frames_per_utt": FRAMES_OUT, "fs": 16000,
            "l2_policy": "GPU arm: inputs larger than L2 (%.0f MB PCM per step per GPU vs %d MB L2); CPU arm: not applicable"
                         % (64 * UTT_SAMPLES * 2 / 1e6, L2_BYTES >> 20)}


def cpu_reference_step(models, sig, threads):
    """one C2 utterance through the reference's CPU path: MFCC (float64 NumPy restatement of src/feature/MFCC.py, per-frame
    loop kept) + diff_feature(nd=2) + its own compiled C++ scorer (score_all per speaker, concurrency = `threads`)"""
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


def _mfcc_pool_job(seed):
    from oracle import mfcc_ref
    from speaker_recognition_b200 import synth
    sig = synth.audio(_POOL_SAMPLES, seed=seed)
    t0 = time.perf_counter()
    f = mfcc_ref.extract39(16000, sig)
    return len(f), time.perf_counter() - t0


_POOL_SAMPLES = (2002 - 1) * 256 + 512        # 2 000 scored frames per pool job (a fifth of a C2 utterance)


def cpu_variants(W, M, Sg, sig, feat):
    """The CPU compositions BASELINE.md s3 names, each timed on a bounded sample of the C2 workload (one utterance, or a stated
    fraction scaled up).  Returns (variants dict, best scoring seconds per utterance, best MFCC seconds per utterance)."""
    from oracle import mfcc_ref, ref_so
    threads = os.cpu_count() or 1
    out = {}
    T = len(feat)
    models, kind = cpu_models(W, M, Sg)
    best_score = None
    if kind == "reference":
        # (a) compiled reference, contiguous input (ref_score_batch: the C++ side builds its vector<vector<double>>, pygmm.cc:16-23)
        for conc in sorted({1, 8, 32, threads}):
            n_m = 10 if conc == 1 else S_SPK            # the 1-thread point is timed on 10 models and scaled
            t0 = time.perf_counter()
            for m in models[:n_m]:
                m.score(feat, conc)
            dt = (time.perf_counter() - t0) * (S_SPK / n_m)
            out["ref_cpp_conc%d" % conc] = {"score_s_per_utt": dt, "scaled_from_models": n_m}
            best_score = dt if best_score is None else min(best_score, dt)
        # (b) the same through the exported score_all with the Python-side double** marshalling (what a ctypes caller pays)
        t0 = time.perf_counter()
        for m in models:
            m.score_all_abi(feat, threads)
        out["ref_cpp_score_all_abi_conc%d" % threads] = {"score_s_per_utt": time.perf_counter() - t0, "scaled_from_models": S_SPK}
        # (c) as shipped: pygmm.py:89-95 builds one ctypes array PER ROW per model call; timed for 1 model, scaled to S
        import ctypes as Ct
        t0 = time.perf_counter()
        rows = [(Ct.c_double * DIM)(*r) for r in feat.tolist()]
        arr = (Ct.POINTER(Ct.c_double) * T)(*[Ct.cast(r, Ct.POINTER(Ct.c_double)) for r in rows])
        t_marsh = time.perf_counter() - t0
        del rows, arr
        out["as_shipped_pygmm_marshalling"] = {"score_s_per_utt": best_score + t_marsh * S_SPK, "marshalling_s_per_model_call": t_marsh,
                                               "scaled_from_models": 1}
    # (d) sklearn GaussianMixture(diag).score_samples, looped over the speakers (timed on 10, scaled)
    try:
        from sklearn.mixture import GaussianMixture
        gms = []
        for i in range(10):
            g = GaussianMixture(n_components=K_MIX, covariance_type="diag")
            g.weights_, g.means_, g.covariances_ = W[i], M[i], Sg[i] ** 2
            g.precisions_cholesky_ = 1.0 / Sg[i]
            gms.append(g)
        t0 = time.perf_counter()
        for g in gms:
            g.score_samples(feat)
        dt = (time.perf_counter() - t0) * (S_SPK / 10)
        out["sklearn_diag_score_samples"] = {"score_s_per_utt": dt, "scaled_from_models": 10}
        best_score = dt if best_score is None else min(best_score, dt)
    except Exception as e:                                  # sklearn missing on the box: say so, do not fail the bench
        out["sklearn_diag_score_samples"] = {"unavailable": str(e)[:100]}
    # (e) MFCC: the restated per-frame loop on 1 core (timed by the caller), and over a process pool as the reference's scripts
    # do across files (src/test/extract-mfcc-data.py:41-47): `threads` jobs of 2 000 frames each
    mfcc_pool = None
    try:
        import multiprocessing as mp
        n_jobs = min(threads, 64)
        with mp.get_context("fork").Pool(n_jobs) as pool:
            pool.map(_mfcc_pool_job, range(n_jobs))          # warm the workers (imports, tables)
            t0 = time.perf_counter()
            res = pool.map(_mfcc_pool_job, range(n_jobs))
            wall = time.perf_counter() - t0
        fr = sum(r[0] for r in res)
        mfcc_pool = wall * FRAMES_OUT / fr                  # seconds of wall clock per C2 utterance's worth of frames
        out["mfcc_pool"] = {"mfcc_s_per_utt_equiv": mfcc_pool, "processes": n_jobs, "frames": fr, "wall_s": wall}
    except Exception as e:
        out["mfcc_pool"] = {"unavailable": str(e)[:100]}
    return out, best_score, mfcc_pool, kind


def best_concurrency(models, feat, threads):
    best = (None, threads)
    for conc in sorted({8, 32, threads}):
        t0 = time.perf_counter()
        for m in models[:10]:
            m.score_all_abi(feat, conc) if hasattr(m, "score_all_abi") else m(feat, conc)
        dt = time.perf_counter() - t0
        if best[0] is None or dt < best[0]:
            best = (dt, conc)
    return best[1]


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, bounded sample per step"""
    if rank != 0:
        return
    from oracle import mfcc_ref
    from speaker_recognition_b200 import synth
    threads = os.cpu_count() or 1
    W, M, Sg = synth.models(S_SPK, K_MIX, DIM, seed0=1000)
    models, kind = cpu_models(W, M, Sg)
    sig = synth.audio(UTT_SAMPLES, seed=0)
    conc = best_concurrency(models, mfcc_ref.extract39(16000, sig[:200000]), threads)       # the scorer's fastest thread count here
    for _ in range(args.warmup):
        cpu_reference_step(models, sig, conc)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        n, _, _, _ = cpu_reference_step(models, sig, conc)
        frames += n
    dt = time.perf_counter() - t0
    val = frames / dt
    sample = "1 utterance (%d frames) x %d speakers per step; MFCC = float64 NumPy restatement (1 core), scoring = %s C++ score_all x %d speakers, concurrency=%d (fastest of 8/32/%d)" % (
        FRAMES_OUT, S_SPK, "compiled reference" if kind == "reference" else "C port of the reference", S_SPK, conc, threads)
    line = {"impl": "reference", "metric": "mfcc_gmm_scored_frames_per_sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(), "batch": {"utts_per_step": 1},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# sub-records for the other BASELINE configs (all ranks call these; rank 0 keeps the result)
# ------------------------------------------------------------------------------------------------
def _maxed(ms, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sub_c2_realparams(L, steps=5):
    """the scorer alone on 100 speakers x 128 mixtures whose (mean, sigma) rows are resampled from the UBMs the reference ships
    (D = 34 = MFCC 19 + LPC 15, un-normalised scales, sigma floor present): frames drawn from the models, device-resident"""
    import torch
    from speaker_recognition_b200 import _lib, synth
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    gdir = os.path.join(ROOT, "tests", "golden")
    zg, zr = np.load(os.path.join(gdir, "gmm_ref.npz")), np.load(os.path.join(gdir, "realparams_ref.npz"))
    pools = [(zg[t + "/w"], zg[t + "/mean"], zg[t + "/sigma"]) for t in ("ubm32", "ubm64", "ubm256")]
    pools.append((zr["ubm32p/w"], zr["ubm32p/mean"], zr["ubm32p/sigma"]))
    S, K, B, T = S_SPK, K_MIX, 64, FRAMES_OUT
    W, M, Sg = synth.realparam_models(S, K, pools, seed0=6000)
    D = M.shape[2]
    gs = GMMSet(gmm_order=K)
    for i in range(S):
        gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
        gs.y.append(i)
    ld = (D + 3) & ~3
    X = torch.zeros((B * T, ld), dtype=torch.float32, device="cuda")
    truth = []
    for u in range(B):
        sp = (u * 7) % S
        truth.append(sp)
        X[u * T:(u + 1) * T, :D] = torch.from_numpy(synth.draw(W[sp], M[sp], Sg[sp], T, seed=100 + u).astype(np.float32)).cuda()
    off = (C.c_int * (B + 1))(*[u * T for u in range(B + 1)])
    sums = torch.zeros((B, S), dtype=torch.float64, device="cuda")
    best = torch.zeros(B, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    h = gs._ensure_set()

    def step():
        _lib.check(L.srb200_gmmset_score_dev(h, X.data_ptr(), ld, D, off, B, sums.data_ptr(), None, 0, best.data_ptr(), st))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in ev:
        e0.record()
        step()
        e1.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    return {"workload": "C2-realparams", "speakers": S, "mixtures": K, "dim": D, "utts": B, "frames_per_utt": T, "score_ms": ms,
            "route": int(L.srb200_last_score_route()), "algorithmic_tflops": 3.0 * D * B * T * S * K / (ms * 1e-3) / 1e12,
            "scored_frames_per_sec_scorer_only": B * T / (ms * 1e-3), "argmax_ok": bool(np.array_equal(best.cpu().numpy(), np.array(truth))),
            "note": "scorer kernel only; parameters resampled from the four shipped UBMs (tests/golden), frames drawn from the models"}


def sub_c3_em(L, rank, world, iters=10, N=1000000, K=512, D=39):
    """BASELINE config 3: UBM EM on 1 M frames, frame-sharded, ONE all-reduce of the float64 statistics per iteration"""
    import torch
    import torch.distributed as dist
    from speaker_recognition_b200 import dist as sdist, synth
    from speaker_recognition_b200.gmm.pygmm import GMM
    g = torch.Generator(device="cuda").manual_seed(1234)
    wt, mt, st_ = synth.model(K, D, seed=77, floor_frac=0.0)
    comp = torch.multinomial(torch.from_numpy(wt).cuda(), N, replacement=True, generator=g)
    X = torch.from_numpy(mt).cuda()[comp] + torch.from_numpy(st_).cuda()[comp] * torch.randn((N, D), device="cuda", dtype=torch.float64, generator=g)
    lo, hi = sdist.shard_range(N, rank, world)
    ld = (D + 3) & ~3
    Xd = torch.zeros((hi - lo, ld), dtype=torch.float32, device="cuda")
    Xd[:, :D] = X[lo:hi].float()
    idx = torch.randint(0, N, (K,), generator=torch.Generator().manual_seed(5))
    m0 = X[idx.cuda()].cpu().numpy()
    s0 = np.tile(X.std(dim=0, unbiased=True).cpu().numpy(), (K, 1))
    w0 = np.ones(K) / K
    del X, comp

    passes = [0]

    def run(n_it, n_local=None):
        n_local = hi - lo if n_local is None else n_local
        gm = GMM(K).set_params(w0, m0, s0)
        be = sdist.DeviceEM(gm, Xd, n_local, ld, D)
        passes[0] = 0
        inner = be.estep

        def counted():
            passes[0] += 1
            return inner()
        be.estep = counted
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done, ll = sdist.em_train(be, N if n_local == hi - lo else n_local * world, n_it, threshold=0.0)
        e1.record()
        torch.cuda.synchronize()
        return done, ll, e0.elapsed_time(e1)
    run(2)
    done, ll, ms = run(iters)
    ms = _maxed(ms, world)
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
        ar_us = _maxed(e0.elapsed_time(e1) / 50 * 1e3, world)
    n_estep = passes[0]                             # the monitor of iteration i is the E-step of iteration i + 1: one extra pass at the end at most
    route_main = int(L.srb200_last_em_route())
    per_it = ms / max(done, 1)
    # the part of an iteration that does not shrink with the shard (operand packing, the route flag read, M-step, launches, the
    # all-reduce): the same loop on the smallest shard the tensor route takes (n K = 2^24)
    n_small = min(hi - lo, (1 << 24) // K)
    run(2, n_small)
    d_s, _, ms_s = run(iters, n_small)
    fixed_ms = _maxed(ms_s, world) / max(d_s, 1)
    ar_share = ar_us * 1e-3 / per_it
    if world == 1:
        limit = "single GPU: no exchange; %.2f of %.2f ms per iteration is the fixed part (measured on a %d-frame shard)" % (fixed_ms, per_it, n_small)
    elif ar_share >= 0.2:
        limit = "exchange-bound: the all-reduce is %.0f %% of an iteration" % (100.0 * ar_share)
    elif fixed_ms >= 0.5 * per_it:
        limit = ("latency-bound: %.2f of %.2f ms per iteration is the fixed part that does not shrink with the shard (measured on a %d-frame shard: operand "
                 "packing, route flag read, M-step, launches, all-reduce %.0f us = %.0f %%)" % (fixed_ms, per_it, n_small, ar_us, 100.0 * ar_share))
    else:
        limit = "compute-bound: the all-reduce is %.1f %% of an iteration, the fixed part %.2f of %.2f ms" % (100.0 * ar_share, fixed_ms, per_it)
    return {"workload": "C3", "mixtures": K, "dim": D, "frames": N, "nranks": world, "frames_per_rank": hi - lo, "iterations": done,
            "ms_total": ms, "ms_per_iteration": per_it, "frames_per_sec_per_iteration": N / (per_it * 1e-3), "allreduce_us": ar_us,
            "allreduce_bytes": int(stats.numel() * 8), "scaling": "strong", "final_ll": ll,
            "algorithmic_tflops": 7.0 * D * N * K * done / (ms * 1e-3) / 1e12,
            "limit": limit, "fixed_ms_per_iteration": fixed_ms,
            "e_step_passes": n_estep, "em_route_last": route_main,
            "em_route_note": "1 = tcgen05 E-step (gmm_em_tc.cu), 0 = fp32 FMA kernels (gmm_em.cu)"}


def sub_c4(L, rank, world, S=1000, K=256, B=8, steps=5):
    """BASELINE config 4: 1000 speakers x 256 mixtures, speaker-sharded: every rank scores the same features against its S/world
    speakers, one all-gather of the [B, S_local] sums, global argmax"""
    import torch
    import torch.distributed as dist
    from speaker_recognition_b200 import dist as sdist, synth
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.gmmset import GMMSet
    from speaker_recognition_b200.pipeline import BatchScorer
    lo, hi = sdist.shard_range(S, rank, world)
    gs = GMMSet(gmm_order=K)
    for i in range(lo, hi):
        w, m, s = synth.model(K, DIM, seed=1000 + i)
        gs.gmms.append(GMM(K).set_params(w, m, s))
        gs.y.append(i)
    sigs = [synth.audio(UTT_SAMPLES, seed=u) for u in range(2)]
    host = torch.empty(B * UTT_SAMPLES, dtype=torch.int16).pin_memory()
    for u in range(B):
        host.numpy()[u * UTT_SAMPLES:(u + 1) * UTT_SAMPLES] = sigs[u % len(sigs)]
    bs = BatchScorer(gs, fs=16000, nd=ND)
    bs._prepare([UTT_SAMPLES] * B)
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
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        if i == steps - 1:
            bs.features()
            k0.record()
            bs.score()
            k1.record()
            best, full = sdist.speaker_sharded_argmax(bs.d_sums)
        else:
            best, full = step()
    e1.record()
    torch.cuda.synchronize()
    ms = _maxed(e0.elapsed_time(e1), world) / steps
    frames = B * FRAMES_OUT
    return {"workload": "C4", "speakers": S, "mixtures": K, "dim": DIM, "utts_per_step": B, "frames_per_utt": FRAMES_OUT, "nranks": world,
            "speakers_per_gpu": hi - lo, "ms_per_step": ms, "score_ms": _maxed(k0.elapsed_time(k1), world), "route": int(L.srb200_last_score_route()),
            "scored_frames_per_sec": frames / (ms * 1e-3), "frame_model_pairs_per_sec": frames * S / (ms * 1e-3), "scaling": "strong",
            "argmax_sample": [int(x) for x in best[:4].cpu()]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--utts", type=int, default=64, help="utterances per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true", help="skip the C2-realparams / C3 / C4 sub-records")
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

    # ---- end-to-end through the public batched API: pinned host PCM -> labels on host, every step.  (1) one step at a time
    # (predict_pcm = submit + result); (2) the same API used as a serving loop uses it: step i+1 is submitted before step i's labels
    # are collected, so its first H2D copies overlap step i's last kernels (two buffer slots) -- every step still copies its own
    # input from pinned host memory and reads its own labels back inside the timed region.  (2) is the e2e headline.
    for _ in range(2):
        bs.predict_pcm(host, lengths)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lab = bs.predict_pcm(host, lengths)
    e1.record()
    barrier()
    e2e_sync_ms = e0.elapsed_time(e1)
    assert np.array_equal(lab.numpy(), labels_dev)
    pending = bs.submit(host, lengths)                 # warm-up of the two-in-flight mode (its sub-batch shapes are new to the caches)
    for _ in range(max(args.warmup, 3)):
        nxt = bs.submit(host, lengths)
        bs.result(pending)
        pending = nxt
    bs.result(pending)
    barrier()
    w2 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pending = bs.submit(host, lengths)
    for _ in range(args.steps - 1):
        nxt = bs.submit(host, lengths)
        lab = bs.result(pending)
        pending = nxt
    lab = bs.result(pending)
    e1.record()
    barrier()
    w3 = time.perf_counter()
    e2e_ms = e0.elapsed_time(e1)
    time.sleep(0.12)
    sampler.stop()
    clocks = sampler.summary([(w0, w1 + 0.06), (w2, w3 + 0.06)])
    assert np.array_equal(lab.numpy(), labels_dev)

    t = torch.tensor([dev_ms, e2e_ms, score_ms, feat_ms, e2e_sync_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, score_ms, feat_ms, e2e_sync_ms = [float(x) for x in t.cpu()]
    total_frames = frames_per_step * args.steps * world
    value = total_frames / (dev_ms * 1e-3)
    e2e_value = total_frames / (e2e_ms * 1e-3)
    route = int(L.srb200_last_score_route())          # 1: tcgen05 3xFP16 kernel on CTA pairs (default), 0: fp32 FMA kernel

    # ---- sub-records for the other BASELINE configs (every rank takes part; timing is over, so they cannot disturb the headline)
    sub = {}
    if not args.no_subrecords:
        del bs
        torch.cuda.empty_cache()
        for name, fn in (("c3_em", lambda: sub_c3_em(L, rank, world)), ("c4", lambda: sub_c4(L, rank, world)),
                         ("c2_realparams", (lambda: sub_c2_realparams(L)) if rank == 0 else None)):
            if fn is None:
                continue
            try:
                sub[name] = fn()
            except Exception as e:          # a sub-record must never take the headline line down with it
                sub[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
            torch.cuda.empty_cache()

    if rank == 0:
        hbm_peak, peak_src, peaks = load_peaks()
        # algorithmic bytes/flops of ONE scoring launch: B_s = 4 T D + 4 S K (2D+1) + 8 S per utterance batch,
        # F_s = 3 D T S K (sub, mul, mul-add per dim; exps and logs not counted)
        T = frames_per_step
        bytes_alg = 4.0 * T * DIM + 4.0 * S_SPK * K_MIX * (2 * DIM + 1) + 8.0 * S_SPK * B
        flops_alg = 3.0 * DIM * T * S_SPK * K_MIX
        sec = score_ms * 1e-3
        kname = "gmm_score_tc2_kernel" if route == 1 else "gmm_score_kernel"
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
        # fp32 FMA pipe peak, measured in this run (csrc/bench/fma_peak.cu): the MFCC kernels' bound, and the scorer's on route 0
        fma_peak, fma_src, fma_variants = 148 * 128 * 2 * 1.965e9 / 1e12, "nominal 148 SM x 128 lanes x 2 x 1.965 GHz", None
        try:
            from speaker_recognition_b200 import build as b
            LB = C.CDLL(b.BENCH_OUT)
            LB.srb200_bench_fma_tflops.restype = C.c_double
            LB.srb200_bench_fma_tflops.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
            fma_variants = {name: LB.srb200_bench_fma_tflops(v, 4000, 3, None) for v, name in ((0, "ffma_indep"), (1, "ffma_pair"), (2, "ffma2_packed"))}
            fma_peak, fma_src = max(fma_variants.values()), "measured in this run (csrc/bench/fma_peak.cu)"
        except Exception as e:
            fma_src += " (helper failed: %s)" % e
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        if route == 1:
            # tensor route: the Mahalanobis sum as a GEMM [T x (2D+1)] . [(2D+1) x S K] on tcgen05 kind::f16, every fp32 operand split into
            # an fp16 head and tail and multiplied in 3 passes (fp32-grade accuracy): the kernel EXECUTES 3 x 2 Kp T S K flops for the
            # 3 D T S K algorithmic ones.  Peak = the measured cuBLAS bf16 figure (the same 16-bit tensor rate): the sustained one is the
            # denominator (the kernel is timed inside the step loop), burst and the clock-derived pipe peak are given beside it.
            kp = ((2 * DIM + 1 + 15) // 16) * 16
            flops_exec = 3.0 * 2.0 * kp * T * S_SPK * K_MIX
            sust, burst = peaks.get("bf16_tflops_sustained"), peaks.get("bf16_tflops")
            tpeak = sust or burst or 1590.0
            tsrc = ("measured bf16_tflops_sustained (MEASURED_PEAKS.json): fp16/bf16 dense tensor rate, the unit this kernel runs on" if sust
                    else "fallback: 1.59 PFLOP/s (B200_PROFILING.md)")
            ach = flops_alg / sec / 1e12
            roofline = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                        "traffic": traffic, "peak_source": tsrc, "ms_per_launch": score_ms, "algorithmic_flops": flops_alg,
                        "flops_convention": "3 D T S K (SURVEY.md s8d)",
                        "frac_vs_peak_burst": (ach / burst) if burst else None,
                        "peak_clock_derived": 148 * 8192 * sm_mhz * 1e6 / 1e12, "frac_vs_peak_clock_derived": ach / (148 * 8192 * sm_mhz * 1e6 / 1e12),
                        "frac_vs_tf32_peak_as_in_round1": ach / (tpeak / 2.0),
                        "executed_flops": flops_exec, "executed": flops_exec / sec / 1e12, "executed_frac": flops_exec / sec / 1e12 / tpeak,
                        "ceiling_note": "a 3-pass split caps `frac` at (3 D)/(3 * 2 Kp) = %.3f of the tensor peak; executed_frac is the tensor-pipe "
                                        "utilisation against the same peak" % (3.0 * DIM / (6.0 * kp))}
            binding = "tensor_fp16"
        else:
            ach = flops_alg / sec / 1e12
            roofline = {"bound": "fp32_fma", "kernel": kname, "achieved": ach, "peak": fma_peak, "unit": "TFLOP/s", "frac": ach / fma_peak,
                        "traffic": traffic, "peak_source": fma_src, "ms_per_launch": score_ms, "algorithmic_flops": flops_alg,
                        "flops_convention": "3 D T S K (SURVEY.md s8d); the kernel issues 2 FFMA = 4 flops per (frame, mixture, dim)"}
            binding = "fp32_fma"
        # MFCC kernels (frames + stats + norm, timed together): flops per frame by SURVEY.md s8d
        nfft, nnz = 2048, 1989
        mfcc_flops_frame = 2.5 * nfft * np.log2(nfft) + 3 * (nfft // 2 + 1) + 2 * nnz + 2 * N_CEPS * 50 + 4 * 512
        raw_frames = B * RAW_FRAMES
        mf_ach = mfcc_flops_frame * raw_frames / (feat_ms * 1e-3) / 1e12
        roofline_mfcc = {"bound": "fp32_fma", "kernel": "mfcc_frames_warp_kernel (+ mfcc_stats_kernel, mfcc_norm_kernel)", "achieved": mf_ach, "peak": fma_peak,
                         "unit": "TFLOP/s", "frac": mf_ach / fma_peak, "peak_source": fma_src, "peak_variants": fma_variants, "ms_per_step": feat_ms,
                         "algorithmic_flops_per_frame": mfcc_flops_frame, "frames": raw_frames,
                         "hbm_gbs": (2.0 * 256 + 4.0 * N_CEPS) * raw_frames / (feat_ms * 1e-3) / 1e9,
                         "note": "FFT butterflies are add-heavy (no FMA pairing), so the FMA-pipe fraction understates the issue-slot use; "
                                 "neither HBM (%.1f %% of peak) nor the FMA pipe binds: the kernel is bound by the shared-memory pipe (88 %% l1tex throughput, profiles/r2g_mfcc_warp.txt)"
                                 % (100.0 * (2.0 * 256 + 4.0 * N_CEPS) * raw_frames / (feat_ms * 1e-3) / 1e9 / hbm_peak)}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            from oracle import mfcc_ref
            t0 = time.perf_counter()
            feat_cpu = mfcc_ref.extract39(16000, base[0])
            t_mfcc1 = time.perf_counter() - t0
            variants, best_score, mfcc_pool, kind = cpu_variants(W, M, Sg, base[0], feat_cpu)
            variants["mfcc_1core"] = {"mfcc_s_per_utt": t_mfcc1}
            t_m = min(t_mfcc1, mfcc_pool) if mfcc_pool else t_mfcc1
            # top-1 of the CPU path on this utterance (compiled reference, or sklearn-free port)
            models, _ = cpu_models(W, M, Sg)
            cpu_lab = int(np.argmax([m.score(feat_cpu, threads).sum() if hasattr(m, "score") else m(feat_cpu, threads) for m in models]))
            cpu_baseline = {"value": FRAMES_OUT / (t_m + best_score), "unit": "frames/s", "cores": threads, "kind": kind,
                            "sample": "1 of the %d utterances of a step (%d frames x %d speakers); value = the FASTEST composition: MFCC %.3f s per utterance "
                                      "(%s) + scoring %.3f s (best of the variants)" % (B, FRAMES_OUT, S_SPK, t_m, "process pool" if mfcc_pool and mfcc_pool < t_mfcc1 else "1 core", best_score),
                            "mfcc_s": t_m, "score_s": best_score, "variants": variants, "top1_matches_gpu": bool(cpu_lab == int(labels_dev[0]))}
        line = {"metric": "mfcc_gmm_scored_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16x3+f32" if route == 1 else "f32", "data": "synthetic",
                "config": workload_config(),
                "batch": {"utts_per_step_per_gpu": B, "parallelism": "utterance-sharded x%d" % world},
                "frame_model_pairs_per_sec": value * S_SPK,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(B * UTT_SAMPLES * 2), "d2h_bytes_per_step": int(B * 4),
                        "ms_per_step": e2e_ms / args.steps, "api": "pipeline.BatchScorer.submit / result, two steps in flight",
                        "one_step_at_a_time": {"value": total_frames / (e2e_sync_ms * 1e-3), "ms_per_step": e2e_sync_ms / args.steps,
                                               "api": "pipeline.BatchScorer.predict_pcm"}},
                "gpu_launches": BatchScorer.KERNELS_PER_STEP * args.steps,
                "kernel_ms": {"mfcc_features": feat_ms, "gmm_score": score_ms},
                "score_route": "tcgen05_3xfp16_cta_pairs" if route == 1 else "fp32_fma",
                "roofline": roofline, "roofline_mfcc": roofline_mfcc, "roofline_hbm": hbm,
                "binding_roofline": binding,
                "cpu_baseline": cpu_baseline, "clocks": clocks}
        line.update(sub)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
