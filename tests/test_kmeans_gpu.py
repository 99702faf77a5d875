"""GPU parity tests of the K-MeansII initialisation (the reference's init_with_kmeans = 1, gmm.cc:331-344), through the C ABI
(srb200_kmeans_lloyd_f64 / srb200_kmeans_init_f64 / train with Parameter.init_with_kmeans).

 - Lloyd iterations from injected centroids against the COMPILED REFERENCE's own (tests/golden/kmeans_ref.npz,
   KMeansSolver::Lloyds_iteration[_weighted], kmeans.cc:150-340): deterministic, 1e-9;
 - the whole initialisation against the oracle (oracle/kmeans_ref.py) fed the same counter-based draws: 1e-9;
 - statistically against the reference's final costs over srand() seeds (same algorithm, different generator);
 - edge cases: K = 1, fewer points than 2K candidates, duplicated points (empty clusters)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import kmeans_ref
from speaker_recognition_b200 import _lib, synth
from speaker_recognition_b200.gmm.pygmm import GMM

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "kmeans_ref.npz"))


def lloyd_gpu(X, cent, weight=None):
    L = _lib.lib()
    X = _lib.as_f64(X)
    c = _lib.as_f64(cent).copy()
    w = None if weight is None else _lib.as_f64(weight)
    cost, it = C.c_double(), C.c_int()
    _lib.check(L.srb200_kmeans_lloyd_f64(_lib.dptr(X), X.shape[0], X.shape[1], c.shape[0], None if w is None else _lib.dptr(w), _lib.dptr(c),
                                         C.byref(cost), C.byref(it)), "kmeans_lloyd")
    return c, cost.value, it.value


def init_gpu(X, K, seed):
    L = _lib.lib()
    X = _lib.as_f64(X)
    c = np.zeros((K, X.shape[1]))
    cost = C.c_double()
    _lib.check(L.srb200_kmeans_init_f64(_lib.dptr(X), X.shape[0], X.shape[1], K, C.c_uint64(seed), _lib.dptr(c), C.byref(cost)), "kmeans_init")
    return c, cost.value


@pytest.mark.parametrize("case", ["a", "b"])
def test_lloyd_matches_compiled_reference(case):
    X, init, w = GOLD[case + "/X"], GOLD[case + "/init"], GOLD[case + "/weight"]
    c, cost, _ = lloyd_gpu(X, init)
    assert np.max(np.abs(c - GOLD[case + "/lloyd_cent"])) < 1e-9 and abs(cost - float(GOLD[case + "/lloyd_cost"])) < 1e-9 * cost
    c, cost, _ = lloyd_gpu(X, init, w)
    assert np.max(np.abs(c - GOLD[case + "/wlloyd_cent"])) < 1e-9 and abs(cost - float(GOLD[case + "/wlloyd_cost"])) < 1e-9 * cost


@pytest.mark.parametrize("case,seed", [("a", 1), ("a", 77), ("b", 5)])
def test_kmeansII_matches_oracle_with_the_same_draws(case, seed):
    X, K = GOLD[case + "/X"], int(GOLD[case + "/K"])
    c, cost = init_gpu(X, K, seed)
    co, costo, _ = kmeans_ref.kmeansII(X, K, seed)
    assert abs(cost - costo) < 1e-9 * costo
    assert np.max(np.abs(c - co)) < 1e-8
    c2, cost2 = init_gpu(X, K, seed)                    # reproducible from the seed, bit for bit
    assert np.array_equal(c, c2) and cost == cost2


@pytest.mark.parametrize("case", ["a", "b"])
def test_kmeansII_quality_against_reference_costs(case):
    X, K = GOLD[case + "/X"], int(GOLD[case + "/K"])
    ours = np.array([init_gpu(X, K, s)[1] for s in range(1, 13)])
    ref = GOLD[case + "/kmeansII_costs"]                 # the compiled reference over 30 srand() seeds
    good_ref = np.mean(ref < 1.02 * ref.min())
    assert np.mean(ours < 1.02 * ref.min()) >= 0.6 * good_ref and ours.min() < 1.01 * ref.min()


def test_edge_cases():
    rng = np.random.default_rng(3)
    X = rng.normal(size=(50, 4)).astype(np.float32).astype(np.float64)
    c, cost = init_gpu(X, 1, 9)                           # K = 1: the centroid is the mean
    assert np.max(np.abs(c[0] - X.mean(0))) < 1e-12 and abs(cost - ((X - X.mean(0)) ** 2).sum()) < 1e-9 * cost
    c, cost = init_gpu(X[:5], 4, 2)                       # n < 2K: the fill loop repeats rows (kmeansII.cc:157-160)
    assert np.all(np.isfinite(c)) and cost >= 0
    co, costo, _ = kmeans_ref.kmeansII(X[:5], 4, 2)
    assert np.max(np.abs(c - co)) < 1e-9
    Xd = np.repeat(X[:3], 20, axis=0)                     # 3 distinct points, K = 6: clusters run empty, no NaN (deviation from the reference)
    c, cost = init_gpu(Xd, 6, 4)
    assert np.all(np.isfinite(c)) and cost < 1e-18
    with pytest.raises(_lib.SrbError):
        init_gpu(X, 0, 1)


def test_fit_with_kmeans_init_trains_and_is_reproducible():
    K, D = 16, 13
    W, M, Sg = synth.models(1, K, D, seed0=41)
    X = synth.draw(W[0], M[0], Sg[0], 6000, seed=9)
    g1 = GMM(K, nr_iteration=15, init_with_kmeans=1, seed=123).fit(X)
    g2 = GMM(K, nr_iteration=15, init_with_kmeans=1, seed=123).fit(X)
    g0 = GMM(K, nr_iteration=15, init_with_kmeans=0, seed=123).fit(X)
    assert np.array_equal(g1.get_params()[1], g2.get_params()[1])
    ll1, ll0 = g1.score_all(X), g0.score_all(X)
    assert np.isfinite(ll1) and ll1 > ll0 - 0.02 * abs(ll0)          # k-means start is at least as good as random rows (to 2 %)
    # the means after ONE EM iteration from the k-means start equal the oracle's: init centroids -> same EM step
    cent, _ = init_gpu(X, K, 55)
    co, _, _ = kmeans_ref.kmeansII(X.astype(np.float32).astype(np.float64), K, 55)
    assert np.max(np.abs(cent - co)) < 1e-8
