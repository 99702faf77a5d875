"""oracle/kmeans_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/README.md): float64 numpy restatement of the reference's
K-MeansII initialisation, the `init_with_kmeans = 1` branch of GMMTrainerBaseline::init_gaussians (src/gmm/src/gmm.cc:331-344).

  lloyd(X, cent, weight=None)      KMeansSolver::Lloyds_iteration / _weighted      src/gmm/src/kmeans.cc:150-246, 249-340
                                   (calc_belonging / calc_belonging_weighted, kmeans.cc:72-147)
  kmeans_pp_weighted(...)          KMeansppSolver::cluster_weighted                src/gmm/src/kmeans++.cc:157-213
  kmeansII(X, K, seed)             KMeansIISolver::cluster                         src/gmm/src/kmeansII.cc:80-172

Pinning (tests/test_oracle_pinning.py, fixture tests/golden/kmeans_ref.npz made by tests/golden/make_golden.py from the
compiled reference through oracle/ref_harness.cc): `lloyd` is compared with the reference's own Lloyd iterations from
the same start (deterministic: centroids to 1e-9); the randomised stages cannot be compared draw for draw -- the
reference consumes libc rand() and std::default_random_engine, this restatement (like the product, csrc/kmeans.cu)
draws u(seed, stream, index) from a counter-based splitmix64 hash -- so `kmeansII` is pinned statistically: its final
cost over several seeds against the reference's cost over several srand() seeds on the same data.

One deliberate deviation, shared with the product: a cluster that loses all its points keeps its previous centroid (the
reference computes 0/0 = NaN, kmeans.cc:231-233 / 326-328)."""
import numpy as np

_M64 = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15


def uniform(seed, stream, idx):
    """u(seed, stream, idx) in [0, 1): the draw the product uses (csrc/kmeans.cu km_uniform); idx may be an array."""
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M64) + np.uint64(_GOLD) * ((np.uint64(stream) << np.uint64(40)) + idx + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _dist_to(X, c):
    """squared distances of all rows to one centroid, dimensions summed in index order (distancesqr, kmeans.cc:62-69)"""
    s = np.zeros(len(X))
    for d in range(X.shape[1]):
        t = X[:, d] - c[d]
        s = s + t * t
    return s


def assign(X, cent, weight=None):
    """calc_belonging(_weighted) (kmeans.cc:72-147): nearest centroid, strict <, first minimum wins"""
    best = np.full(len(X), np.finfo(np.float64).max)
    belong = np.zeros(len(X), dtype=np.int64)
    for j in range(len(cent)):
        dj = _dist_to(X, cent[j])
        if weight is not None:
            dj = dj * weight
        m = dj < best
        best[m] = dj[m]
        belong[m] = j
    return best, belong


def lloyd(X, cent, weight=None, max_iter=200):
    """returns (best centroids, best cost, iterations run)"""
    X = np.asarray(X, dtype=np.float64)
    cent = np.array(cent, dtype=np.float64)
    K = len(cent)
    best_cost, best_cent, last = np.finfo(np.float64).max, cent.copy(), np.finfo(np.float64).max
    it = 0
    for it in range(max_iter):
        d, belong = assign(X, cent, weight)
        cost = float(d.sum())
        if cost < best_cost:
            best_cost, best_cent = cost, cent.copy()
        if abs(last - cost) < 1e-6:                              # kmeans.cc:194 / 299
            break
        if weight is None and cost > best_cost * 1.5:            # kmeans.cc:199-203
            break
        w = np.ones(len(X)) if weight is None else weight
        size = np.bincount(belong, weights=w, minlength=K)
        new = np.zeros_like(cent)
        np.add.at(new, belong, X * w[:, None])
        nz = size > 0
        cent = cent.copy()
        cent[nz] = new[nz] / size[nz, None]                      # empty cluster: keeps its centroid (deviation, see header)
        last = cost
    return best_cent, best_cost, it


def kmeans_pp_weighted(C, weight, K, seed):
    """kmeans++.cc:157-213 over the candidate rows C with the product's draws (stream 200)"""
    m = len(C)
    cent = np.zeros((K, C.shape[1]))
    cent[0] = C[min(m - 1, int(uniform(seed, 200, 0) * m))]
    dist = np.full(m, np.finfo(np.float64).max)
    for k in range(1, K):
        dist = np.minimum(dist, _dist_to(C, cent[k - 1]) * weight)
        S = 0.0
        for v in dist:                                           # index order, like the product's host loop
            S += v
        r = float(uniform(seed, 200, k)) * S
        chosen, last_pos = -1, -1
        for i in range(m):
            if dist[i] > 0.0:
                last_pos = i
            r -= dist[i]
            if r <= 0.0:
                chosen = i
                break
        if chosen < 0:
            chosen = last_pos if last_pos >= 0 else 0
        cent[k] = C[chosen]
    return lloyd(C, cent, weight)


def kmeansII(X, K, seed, l=2.0, size_factor=2.0):
    """kmeansII.cc:80-172.  Returns (centroids [K][D], cost, candidate row indices)."""
    X = np.asarray(X, dtype=np.float64)
    n = len(X)
    cand = [min(n - 1, int(uniform(seed, 0, 0) * n))]
    dist = np.full(n, np.finfo(np.float64).max)
    belong = np.zeros(n, dtype=np.int64)
    last, rnd = 0, 0
    while True:
        for j in range(last, len(cand)):                         # update_distances, kmeansII.cc:52-75
            dj = _dist_to(X, X[cand[j]])
            m = dj < dist
            dist[m] = dj[m]
            belong[m] = j
        if len(cand) > size_factor * K:
            break
        S = float(dist.sum())
        last = len(cand)
        u = uniform(seed, 1 + rnd, np.arange(n))
        cand.extend(np.nonzero(u * S < dist * (l * K))[0].tolist())     # kmeansII.cc:135-141
        if len(cand) == last:
            break
        rnd += 1
    k = 0
    while len(cand) <= size_factor * K:                          # kmeansII.cc:157-160
        cand.append(min(n - 1, int(uniform(seed, 100, k) * n)))
        k += 1
    weight = np.bincount(belong, minlength=len(cand)).astype(np.float64)
    C = X[np.array(cand)]
    cent, _, _ = kmeans_pp_weighted(C, weight, K, seed)
    cent, cost, _ = lloyd(X, cent)
    return cent, cost, np.array(cand)
