"""GPU parity tests of the tensor-core E-step (csrc/gmm_em_tc.cu): sufficient statistics and EM trajectories against the float64
oracle (oracle/gmm_oracle.c em_iteration, which follows gmm.cc:388-653) and against the fp32 FMA kernels of gmm_em.cu.

Tolerance: statistics 1e-4 relative to their scale, trained parameters 5e-4 after several iterations (the bound the FMA
route's trajectory test uses)."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import gmm_ref
from speaker_recognition_b200 import _lib, synth
from speaker_recognition_b200.gmm.pygmm import GMM

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def restore_em_path():
    yield
    _lib.check(_lib.lib().srb200_set_em_path(-1))


def estep(path, w, m, s, X, iters=1, min_covar=1e-3):
    import torch
    L = _lib.lib()
    _lib.check(L.srb200_set_em_path(path))
    g = GMM(len(w)).set_params(w, m, s)
    K, D = m.shape
    ld = (D + 3) & ~3
    Xd = torch.zeros((len(X), ld), dtype=torch.float32, device="cuda")
    Xd[:, :D] = torch.from_numpy(np.asarray(X, dtype=np.float32)).cuda()
    stats = torch.zeros(L.srb200_em_stats_len(K, D), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    hist, routes = [], []
    for _ in range(iters):
        _lib.check(L.srb200_em_estep_dev(g.gmm, Xd.data_ptr(), len(X), ld, D, stats.data_ptr(), st))
        torch.cuda.synchronize()
        routes.append(L.srb200_last_em_route())
        hist.append(stats.cpu().numpy().copy())
        _lib.check(L.srb200_em_mstep_dev(g.gmm, stats.data_ptr(), len(X), min_covar, 0, None, st))
    torch.cuda.synchronize()
    return g, hist, routes


def start(K, D, n, seed):
    wt, mt, st_ = synth.model(K, D, seed=seed, floor_frac=0.0)
    X = synth.draw(wt, mt, st_, n, seed=seed + 1)
    w0 = np.ones(K) / K
    m0 = X[np.random.default_rng(seed + 2).choice(n, K, replace=False)].astype(np.float64)
    s0 = np.tile(X.std(axis=0, ddof=1), (K, 1))
    return X, w0, m0, s0


def oracle_stats(w, m, s, X):
    """N_k, sum gamma x, sum gamma x^2, total log-likelihood in float64 numpy (the E-step of gmm.cc:440-520)"""
    X = np.asarray(X, dtype=np.float32).astype(np.float64)
    lp = (np.log(w)[None, :] - 0.5 * (((X[:, None, :] - m[None]) / s[None]) ** 2).sum(-1) - np.log(np.sqrt(2 * np.pi) * s).sum(-1)[None, :])
    mx = lp.max(1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(lp - mx).sum(1))
    g = np.exp(lp - lse[:, None])
    return g.sum(0), g.T @ X, g.T @ (X * X), lse.sum()


@pytest.mark.parametrize("K,D,n", [(128, 39, 6000), (200, 13, 5000), (512, 39, 3000), (130, 28, 2500), (65, 50, 1500)])
def test_statistics_match_float64_and_the_fma_route(K, D, n):
    X, w0, m0, s0 = start(K, D, n, seed=50 + K)
    _, h1, r1 = estep(1, w0, m0, s0, X)
    _, h0, r0 = estep(0, w0, m0, s0, X)
    assert r1 == [1] and r0 == [0]
    Nk, Sx, Sxx, ll = oracle_stats(w0, m0, s0, X)
    for h in (h1[0], h0[0]):
        assert np.max(np.abs(h[:K] - Nk)) < 1e-4 * max(1.0, Nk.max())
        assert np.max(np.abs(h[K:K + K * D].reshape(K, D) - Sx)) < 1e-4 * max(1.0, np.abs(Sx).max())
        assert np.max(np.abs(h[K + K * D:K + 2 * K * D].reshape(K, D) - Sxx)) < 1e-4 * max(1.0, np.abs(Sxx).max())
        assert abs(h[-2] - ll) < 1e-5 * abs(ll) and h[-1] == n
        assert abs(h[:K].sum() - n) < 1e-4 * n
    # the two routes agree far inside the oracle bound on the per-mixture occupancies
    assert np.max(np.abs(h1[0][:K] - h0[0][:K])) < 2e-5 * max(1.0, Nk.max())


@pytest.mark.parametrize("K,D,n", [(128, 39, 4000), (256, 13, 6000)])
def test_trajectory_matches_oracle(K, D, n):
    X, w0, m0, s0 = start(K, D, n, seed=90 + K)
    iters = 4
    g, _, routes = estep(1, w0, m0, s0, X, iters)
    assert routes[0] == 1                 # later iterations may fall back (a mixture tightening far from the centre): same results either way
    w, m, s = w0, m0, s0
    for _ in range(iters):
        w, m, s = gmm_ref.em_iteration(w, m, s, X)
    gw, gm, gsg = g.get_params()
    assert np.max(np.abs(gw - w) / np.maximum(w, 1e-6)) < 5e-4
    assert np.max(rel_err(gm, m)) < 5e-4 and np.max(np.abs(gsg - s) / s) < 5e-4


def test_offset_and_scaled_features_are_centred_away():
    """features with a large common offset and a small scale (LPC-like): the centring / power-of-two scaling keeps the expanded
    form accurate, and the back-transform of the moments is exact in float64"""
    K, D, n = 128, 20, 5000
    X, w0, m0, s0 = start(K, D, n, seed=7)
    off = np.linspace(-40.0, 60.0, D)
    X2 = (X * 0.03 + off).astype(np.float32)
    m2, s2 = m0 * 0.03 + off, s0 * 0.03
    _, h1, r1 = estep(1, w0, m2, s2, X2)
    assert r1 == [1]
    Nk, Sx, Sxx, ll = oracle_stats(w0, m2, s2, X2)
    h = h1[0]
    assert np.max(np.abs(h[:K] - Nk)) < 1e-4 * max(1.0, Nk.max())
    assert np.max(np.abs(h[K:K + K * D].reshape(K, D) - Sx)) < 1e-4 * np.abs(Sx).max()
    assert np.max(np.abs(h[K + K * D:K + 2 * K * D].reshape(K, D) - Sxx)) < 1e-4 * np.abs(Sxx).max()
    assert abs(h[-2] - ll) < 1e-5 * abs(ll)


def test_far_out_mixture_falls_back_to_the_fma_kernels():
    """a collapsed mixture far from the centre (sum nb''^2 above the budget): the packer raises its flag and the iteration runs on
    the FMA kernels with the same results"""
    K, D, n = 128, 39, 3000
    X, w0, m0, s0 = start(K, D, n, seed=3)
    m0 = m0.copy()
    s0 = s0.copy()
    m0[5] += 30.0 * s0[5]
    s0[5] *= 0.05
    _, h1, r1 = estep(1, w0, m0, s0, X)
    _, h0, r0 = estep(0, w0, m0, s0, X)
    assert r1 == [0] and r0 == [0]
    assert np.array_equal(h1[0], h0[0])


def test_underflowed_frames_are_skipped_like_the_reference():
    """frames whose likelihood underflows (gmm.cc:482-490) contribute nothing and are not counted"""
    K, D, n = 128, 13, 4000
    X, w0, m0, s0 = start(K, D, n, seed=11)
    X = X.copy()
    X[::97] += 400.0                      # log p far below the denormal limit
    _, h1, r1 = estep(1, w0, m0, s0, X)
    _, h0, r0 = estep(0, w0, m0, s0, X)
    assert r1 == [1] and r0 == [0]
    assert h1[0][-1] == h0[0][-1] < n
    assert abs(h1[0][:K].sum() - h1[0][-1]) < 1e-4 * n
    assert np.max(np.abs(h1[0][:K] - h0[0][:K])) < 2e-5 * max(1.0, h0[0][:K].max())
    assert abs(h1[0][-2] - h0[0][-2]) < 1e-6 * abs(h0[0][-2])


def test_default_route_takes_the_tensor_cores_for_large_jobs_only():
    L = _lib.lib()
    X, w0, m0, s0 = start(128, 39, 2000, seed=5)
    _, _, r = estep(-1, w0, m0, s0, X)
    assert r == [0]                       # 2000 x 128 < 2^24
    Xb = np.tile(X, (70, 1))              # 140000 x 128 >= 2^24
    _, _, r = estep(-1, w0, m0, s0, Xb)
    assert r == [1]
    assert L.srb200_last_em_route() == 1
