"""oracle/ref_so.py -- TEST INFRASTRUCTURE ONLY.

Python 3 ctypes loader for the reference's own C++ GMM compiled in-container by oracle/Makefile
(oracle/_ref/pygmm.so = src/gmm/src/*.cc, unmodified; oracle/_ref/ref_harness.so = the same objects
plus oracle/ref_harness.cc).  Unlike the reference's pygmm.py (src/gmm/python/pygmm.py:29-31, which
sets restype for only two functions and so truncates GMM* to a C int on LP64) every signature is
declared.  score_instance is never called: it passes no exp buffer and trips assert(buffer != NULL)
(pygmm.cc:116 -> gmm.cc:185).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")


class Parameter(C.Structure):
    """struct Parameter, src/gmm/src/pygmm.hh:12-26."""
    _fields_ = [("nr_instance", C.c_int), ("nr_dim", C.c_int), ("nr_mixture", C.c_int),
                ("min_covar", C.c_double), ("threshold", C.c_double), ("nr_iteration", C.c_int),
                ("init_with_kmeans", C.c_int), ("concurrency", C.c_int), ("verbosity", C.c_int)]


def available():
    return os.path.isfile(os.path.join(_REF_DIR, "ref_harness.so"))


_lib = None


def lib():
    """The harness library (contains every reference symbol plus ref_*)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_REF_DIR, "ref_harness.so")
    if not os.path.isfile(path):
        raise RuntimeError("compiled reference missing: run `make -C oracle` where /root/reference exists")
    L = C.CDLL(path)
    vp, dp, ci, cd = C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_double
    dpp = C.POINTER(dp)
    L.new_gmm.restype, L.new_gmm.argtypes = vp, [ci, ci]
    L.load.restype, L.load.argtypes = vp, [C.c_char_p]
    L.dump.restype, L.dump.argtypes = None, [vp, C.c_char_p]
    L.train_model.restype, L.train_model.argtypes = None, [vp, dpp, C.POINTER(Parameter)]
    L.train_model_from_ubm.restype, L.train_model_from_ubm.argtypes = None, [vp, vp, dpp, C.POINTER(Parameter)]
    L.score_all.restype, L.score_all.argtypes = cd, [vp, dpp, ci, ci, ci]
    L.score_batch.restype, L.score_batch.argtypes = None, [vp, dpp, dp, ci, ci, ci]
    L.get_dim.restype, L.get_dim.argtypes = ci, [vp]
    L.get_nr_mixtures.restype, L.get_nr_mixtures.argtypes = ci, [vp]
    L.ref_new.restype, L.ref_new.argtypes = vp, [ci, ci, dp, dp, dp]
    L.ref_get.restype, L.ref_get.argtypes = None, [vp, dp, dp, dp]
    L.ref_free.restype, L.ref_free.argtypes = None, [vp]
    L.ref_em_iteration.restype, L.ref_em_iteration.argtypes = None, [vp, dp, ci, ci, cd, ci]
    L.ref_map_iteration.restype, L.ref_map_iteration.argtypes = None, [vp, vp, dp, ci, ci, ci]
    L.ref_score_batch.restype, L.ref_score_batch.argtypes = None, [vp, dp, dp, ci, ci, ci]
    L.ref_score_batch_exactexp.restype, L.ref_score_batch_exactexp.argtypes = None, [vp, dp, dp, ci, ci]
    L.ref_lloyd.restype, L.ref_lloyd.argtypes = cd, [dp, ci, ci, ci, dp, dp, ci]
    L.ref_kmeansII.restype, L.ref_kmeansII.argtypes = cd, [dp, ci, ci, ci, C.c_uint, dp, ci]
    L.ref_kmeanspp_weighted.restype, L.ref_kmeanspp_weighted.argtypes = cd, [dp, ci, ci, dp, ci, C.c_uint, dp, ci]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def rows_pp(X):
    """double** view of a contiguous float64 matrix (what pygmm.cc:16-23 consumes)."""
    X = _c64(X)
    n, d = X.shape
    arr = (C.POINTER(C.c_double) * n)()
    base = X.ctypes.data
    for i in range(n):
        arr[i] = C.cast(base + i * d * 8, C.POINTER(C.c_double))
    return arr, X


class RefGMM:
    """A reference `GMM*` with injected parameters."""

    def __init__(self, w=None, mean=None, sigma=None, handle=None):
        L = lib()
        if handle is not None:
            self.h = handle
        else:
            w, mean, sigma = _c64(w), _c64(mean), _c64(sigma)
            K, D = mean.shape
            self.h = L.ref_new(K, D, _dp(w), _dp(mean), _dp(sigma))

    @classmethod
    def load(cls, path):
        if not os.path.isfile(path):
            raise FileNotFoundError(path)  # the reference silently yields an empty model (gmm.cc:218-221)
        return cls(handle=lib().load(path.encode()))

    def dump(self, path):
        lib().dump(self.h, path.encode())

    @property
    def K(self):
        return lib().get_nr_mixtures(self.h)

    @property
    def D(self):
        return lib().get_dim(self.h)

    def params(self):
        K, D = self.K, self.D
        w, m, s = np.empty(K), np.empty((K, D)), np.empty((K, D))
        lib().ref_get(self.h, _dp(w), _dp(m), _dp(s))
        return w, m, s

    def score(self, X, concurrency=1):
        X = _c64(X)
        out = np.empty(len(X))
        lib().ref_score_batch(self.h, _dp(X), _dp(out), X.shape[0], X.shape[1], concurrency)
        return out

    def score_exactexp(self, X):
        X = _c64(X)
        out = np.empty(len(X))
        lib().ref_score_batch_exactexp(self.h, _dp(X), _dp(out), X.shape[0], X.shape[1])
        return out

    def score_all_abi(self, X, concurrency=1):
        """through the reference's own exported score_all (pygmm.cc:98-102), double** marshalling included"""
        pp, keep = rows_pp(X)
        return lib().score_all(self.h, pp, keep.shape[0], keep.shape[1], concurrency)

    def em_iteration(self, X, min_covar=1e-3, concurrency=1):
        X = _c64(X)
        lib().ref_em_iteration(self.h, _dp(X), X.shape[0], X.shape[1], min_covar, concurrency)

    def map_iteration(self, ubm, X, concurrency=1):
        X = _c64(X)
        lib().ref_map_iteration(self.h, ubm.h, _dp(X), X.shape[0], X.shape[1], concurrency)

    def train(self, X, nr_mixture, nr_iteration=200, min_covar=1e-3, threshold=0.01,
              init_with_kmeans=0, concurrency=1, verbosity=0, ubm=None):
        """the reference's full train_model / train_model_from_ubm (pygmm.cc:63-96); RNG-seeded init,
        writes gmm-training-intermediate-dump.model into the CWD (gmm.cc:624-630)."""
        pp, keep = rows_pp(X)
        p = Parameter(keep.shape[0], keep.shape[1], nr_mixture, min_covar, threshold, nr_iteration,
                      init_with_kmeans, concurrency, verbosity)
        if ubm is None:
            lib().train_model(self.h, pp, C.byref(p))
        else:
            lib().train_model_from_ubm(self.h, ubm.h, pp, C.byref(p))

    def free(self):
        if self.h:
            lib().ref_free(self.h)
            self.h = None


def _quiet_call(fn, *args):
    """the reference's k-means prints progress with printf: silence fd 1 around the call"""
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        return fn(*args)
    finally:
        C.CDLL(None).fflush(None)
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


def ref_lloyd(X, cent, weight=None, concurrency=4):
    """KMeansSolver::Lloyds_iteration[_weighted] (kmeans.cc:150-340) from the given centroids: (centroids, best cost)"""
    X = np.ascontiguousarray(X, dtype=np.float64)
    c = np.ascontiguousarray(cent, dtype=np.float64).copy()
    dp = C.POINTER(C.c_double)
    w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
    cost = _quiet_call(lib().ref_lloyd, X.ctypes.data_as(dp), X.shape[0], X.shape[1], c.shape[0],
                       None if w is None else w.ctypes.data_as(dp), c.ctypes.data_as(dp), concurrency)
    return c, float(cost)


def ref_kmeansII(X, K, seed, concurrency=4):
    """KMeansIISolver::cluster (kmeansII.cc:80-172) after srand(seed): (centroids, cost)"""
    X = np.ascontiguousarray(X, dtype=np.float64)
    c = np.zeros((K, X.shape[1]))
    dp = C.POINTER(C.c_double)
    cost = _quiet_call(lib().ref_kmeansII, X.ctypes.data_as(dp), X.shape[0], X.shape[1], K, int(seed), c.ctypes.data_as(dp), concurrency)
    return c, float(cost)


def ref_kmeanspp_weighted(Cand, weight, K, seed, concurrency=4):
    """KMeansppSolver::cluster_weighted (kmeans++.cc:157-213) after srand(seed): (centroids, weighted cost)"""
    Cand = np.ascontiguousarray(Cand, dtype=np.float64)
    w = np.ascontiguousarray(weight, dtype=np.float64)
    c = np.zeros((K, Cand.shape[1]))
    dp = C.POINTER(C.c_double)
    cost = _quiet_call(lib().ref_kmeanspp_weighted, Cand.ctypes.data_as(dp), Cand.shape[0], Cand.shape[1], w.ctypes.data_as(dp), K, int(seed),
                       c.ctypes.data_as(dp), concurrency)
    return c, float(cost)
