"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing in speaker_recognition_b200.dist: frame-sharded EM with
one all-reduce of the statistics vector per iteration, utterance sharding, speaker-sharded argmax.  The compute
callables are NumPy stand-ins here (the kernels need a GPU; their parity is tested in test_gmm_gpu.py) -- what is
under test is the control flow, the statistics layout that crosses the fabric, and rank consistency."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from oracle import gmm_ref
from speaker_recognition_b200 import dist as sdist
from speaker_recognition_b200 import synth


class NumpyEM(object):
    """float64 stand-in for DeviceEM with the same statistics layout: [K] N_k | [K*D] sum g x | [K*D] sum g x^2 | ll | frames"""

    def __init__(self, w, m, s, X, ubm_mean=None):
        self.w, self.m, self.s, self.X, self.ubm_mean = w.copy(), m.copy(), s.copy(), X, ubm_mean

    def estep(self):
        K, D = self.m.shape
        z = (self.X[:, None, :] - self.m[None]) / self.s[None]
        l = np.log(self.w)[None] - np.sum(np.log(np.sqrt(2 * np.pi) * self.s), axis=1)[None] - 0.5 * np.einsum("tkd,tkd->tk", z, z)
        mx = l.max(axis=1, keepdims=True)
        lse = mx[:, 0] + np.log(np.exp(l - mx).sum(axis=1))
        g = np.exp(l - lse[:, None])
        st = np.concatenate([g.sum(0), (g.T @ self.X).ravel(), (g.T @ (self.X ** 2)).ravel(), [lse.sum(), float(len(self.X))]])
        return torch.from_numpy(st)

    def mstep(self, stats, n_total, min_covar, mode):
        K, D = self.m.shape
        st = stats.numpy()
        Nk = st[:K].copy()
        Nk[Nk == 0] = 1e-6
        sx, sxx = st[K:K + K * D].reshape(K, D), st[K + K * D:K + 2 * K * D].reshape(K, D)
        if mode == 0:
            w = Nk / n_total
            self.w = w / w.sum()
            mu = sx / Nk[:, None]
            var = (sxx - 2 * mu * sx + mu * mu * st[:K][:, None]) / Nk[:, None]
            self.m, self.s = mu, np.maximum(np.sqrt(np.maximum(var, 0)), np.sqrt(min_covar))
        else:
            a = (Nk / (Nk + 16.0))[:, None]
            self.m = sx / Nk[:, None] * a + self.ubm_mean * (1 - a)

    def finish(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- frame-sharded EM ----
        K, D, n = 6, 5, 4001
        wt, mt, st_ = synth.model(K, D, seed=3, floor_frac=0.0)
        X = synth.draw(wt, mt, st_, n, seed=4)
        w0 = np.ones(K) / K
        m0 = X[np.random.default_rng(1).choice(n, K, replace=False)]
        s0 = np.tile(X.std(axis=0, ddof=1), (K, 1))
        lo, hi = sdist.shard_range(n, rank, world)
        be = NumpyEM(w0, m0, s0, X[lo:hi])
        iters, ll = sdist.em_train(be, n_total=n, nr_iteration=6, threshold=0.0)
        np.savez(os.path.join(out_dir, "em_rank%d.npz" % rank), w=be.w, m=be.m, s=be.s, iters=iters, ll=ll, lo=lo, hi=hi)
        # early stop: with a huge threshold the reference stops at the SECOND monitored iteration (the first has no
        # predecessor: last_ll = -DBL_MAX, gmm.cc:612,643-644), i.e. after 4 updates -- identically on all ranks
        be2 = NumpyEM(w0, m0, s0, X[lo:hi])
        it2, _ = sdist.em_train(be2, n_total=n, nr_iteration=50, threshold=1e9)
        # ---- MAP ----
        be3 = NumpyEM(wt, mt, st_, X[lo:hi] + 0.2, ubm_mean=mt)
        sdist.em_train(be3, n_total=n, nr_iteration=1, mode=1)
        # ---- utterance sharding + speaker-sharded argmax ----
        n_utt = 7
        ulo, uhi = sdist.shard_range(n_utt, rank, world)
        labels = sdist.gather_labels(["u%d" % u for u in range(ulo, uhi)], n_utt)
        S = 9
        slo, shi = sdist.shard_range(S, rank, world)
        full = torch.from_numpy(np.random.default_rng(9).standard_normal((4, S)))
        best, gathered = sdist.speaker_sharded_argmax(full[:, slo:shi].clone())
        # ---- speaker-sharded MAP enrolment: every rank adapts its share of the names, models are exchanged at the end ----
        names = ["n%d" % i for i in range(5)]
        feats = {nm: X[300 * i:300 * i + 200 + 37 * i] + 0.1 * i for i, nm in enumerate(names)}
        calls = []

        def fit_many(local_names, local_feats):
            calls.append(list(local_names))
            return [gmm_ref.map_iteration(wt, mt, st_, mt, f, exact_exp=True)[1].tobytes() for f in local_feats]
        enrolled = sdist.enroll_sharded(names, feats, fit_many)
        elo, ehi = sdist.shard_range(len(names), rank, world)
        assert calls == [names[elo:ehi]]                                   # one batched call per rank, over its own names only
        np.savez(os.path.join(out_dir, "misc_rank%d.npz" % rank), it2=it2, map_m=be3.m, labels=np.array(labels), best=best.numpy(),
                 gathered=gathered.numpy(), full=full.numpy(),
                 **{"enr_" + nm: np.frombuffer(enrolled[nm], dtype=np.float64) for nm in names})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_world2_gloo_em_and_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(str(tmp_path / ("em_rank%d.npz" % k))) for k in range(world)]
    assert r[0]["lo"] == 0 and r[0]["hi"] == r[1]["lo"] and r[1]["hi"] == 4001
    # every rank ends with bit-identical parameters (same reduced statistics, same deterministic M-step)
    for key in ("w", "m", "s"):
        assert np.array_equal(r[0][key], r[1][key])
    assert int(r[0]["iters"]) == 6 == int(r[1]["iters"])
    # and they equal the single-process float64 oracle (reference EM arithmetic) after 6 iterations
    K, D, n = 6, 5, 4001
    wt, mt, st_ = synth.model(K, D, seed=3, floor_frac=0.0)
    X = synth.draw(wt, mt, st_, n, seed=4)
    w = np.ones(K) / K
    m = X[np.random.default_rng(1).choice(n, K, replace=False)]
    s = np.tile(X.std(axis=0, ddof=1), (K, 1))
    for _ in range(6):
        w, m, s = gmm_ref.em_iteration(w, m, s, X, exact_exp=True)
    assert np.max(np.abs(r[0]["w"] - w)) < 1e-9 and np.max(np.abs(r[0]["m"] - m)) < 1e-8 and np.max(np.abs(r[0]["s"] - s)) < 1e-8
    # monitored ll = likelihood of the model after iteration index 5 (odd) = after 6 updates
    assert abs(float(r[0]["ll"]) - gmm_ref.score_all(w, m, s, X, gmm_ref.MODE_LOGDOMAIN)) < 1e-6 * abs(float(r[0]["ll"]))
    q = [np.load(str(tmp_path / ("misc_rank%d.npz" % k))) for k in range(world)]
    assert int(q[0]["it2"]) == 4 == int(q[1]["it2"])                       # reference stops after iteration index 3
    wm, mm, sm = gmm_ref.map_iteration(wt, mt, st_, mt, X + 0.2, exact_exp=True)
    assert np.max(np.abs(q[0]["map_m"] - mm)) < 1e-9 and np.array_equal(q[0]["map_m"], q[1]["map_m"])
    assert list(q[0]["labels"]) == ["u%d" % u for u in range(7)] == list(q[1]["labels"])
    assert np.array_equal(q[0]["gathered"], q[0]["full"]) and np.array_equal(q[1]["gathered"], q[0]["full"])
    assert np.array_equal(q[0]["best"], np.argmax(q[0]["full"], axis=1)) and np.array_equal(q[0]["best"], q[1]["best"])
    for i in range(5):                                                     # every rank holds every speaker's model, equal to the oracle's
        f = X[300 * i:300 * i + 200 + 37 * i] + 0.1 * i
        want = gmm_ref.map_iteration(wt, mt, st_, mt, f, exact_exp=True)[1].ravel()
        assert np.array_equal(q[0]["enr_n%d" % i], want) and np.array_equal(q[1]["enr_n%d" % i], want)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [sdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_em_train_control_flow_matches_reference_schedule():
    """single process: iterations applied and the monitoring schedule follow gmm.cc:613-652"""
    K, D, n = 4, 3, 800
    wt, mt, st_ = synth.model(K, D, seed=5, floor_frac=0.0)
    X = synth.draw(wt, mt, st_, n, seed=6)
    w0, m0, s0 = np.ones(K) / K, X[:K].copy(), np.tile(X.std(axis=0, ddof=1), (K, 1))
    for nr_iter in (1, 2, 3, 4):
        be = NumpyEM(w0, m0, s0, X)
        iters, ll = sdist.em_train(be, n, nr_iter, threshold=0.0)
        assert iters == nr_iter
        w, m, s = w0, m0, s0
        for _ in range(nr_iter):
            w, m, s = gmm_ref.em_iteration(w, m, s, X, exact_exp=True)
        assert np.max(np.abs(be.m - m)) < 1e-8
        assert (ll is None) == (nr_iter < 2)            # the first monitored model is the one after iteration index 1
