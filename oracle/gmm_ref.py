"""oracle/gmm_ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes front for oracle/gmm_oracle.c (plain-C float64 restatement of src/gmm/src/gmm.cc scoring + EM)
plus a restatement of the reference's text model format and a NumPy log-sum-exp scorer.
Parity: pinned against the compiled reference in tests/test_oracle_pinning.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "gmm_oracle.so")
_lib = None

MODE_FASTEXP, MODE_LIBMEXP, MODE_LOGDOMAIN = 0, 1, 2
SAFE_LOG_FLOOR = float(np.log(1e-15))      # gmm.cc:34-38  -> -34.538776...
DENORM_LOG_MIN = float(np.log(5e-324))     # smallest positive double: log = -744.44


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            build()
        L = C.CDLL(_SO)
        dp, ci, cd = C.POINTER(C.c_double), C.c_int, C.c_double
        L.orc_fastexp_remez5.restype, L.orc_fastexp_remez5.argtypes = cd, [cd]
        L.orc_score_batch.restype = None
        L.orc_score_batch.argtypes = [ci, ci, dp, dp, dp, dp, ci, dp, ci, ci]
        L.orc_score_all.restype = cd
        L.orc_score_all.argtypes = [ci, ci, dp, dp, dp, dp, ci, ci, ci]
        L.orc_em_iteration.restype = None
        L.orc_em_iteration.argtypes = [ci, ci, dp, dp, dp, dp, ci, cd, ci]
        L.orc_map_iteration.restype = None
        L.orc_map_iteration.argtypes = [ci, ci, dp, dp, dp, dp, dp, ci, ci]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def score(w, mean, sigma, X, mode=MODE_FASTEXP, concurrency=1):
    """per-frame log-likelihoods, float64[n]."""
    w, mean, sigma, X = _c64(w), _c64(mean), _c64(sigma), _c64(X)
    K, D = mean.shape
    out = np.empty(len(X))
    lib().orc_score_batch(K, D, _dp(w), _dp(mean), _dp(sigma), _dp(X), len(X), _dp(out), concurrency, mode)
    return out


def score_all(w, mean, sigma, X, mode=MODE_FASTEXP, concurrency=1):
    w, mean, sigma, X = _c64(w), _c64(mean), _c64(sigma), _c64(X)
    K, D = mean.shape
    return lib().orc_score_all(K, D, _dp(w), _dp(mean), _dp(sigma), _dp(X), len(X), concurrency, mode)


def score_numpy(w, mean, sigma, X):
    """exact float64 log-domain scorer in pure NumPy (independent of the C file)."""
    w, mean, sigma, X = _c64(w), _c64(mean), _c64(sigma), _c64(X)
    const = np.log(w) - np.sum(np.log(np.sqrt(2 * np.pi) * sigma), axis=1)          # [K]
    out = np.empty(len(X))
    for lo in range(0, len(X), 2048):
        x = X[lo:lo + 2048]
        z = (x[:, None, :] - mean[None]) / sigma[None]
        l = const[None] - 0.5 * np.einsum("tkd,tkd->tk", z, z)
        m = l.max(axis=1)
        out[lo:lo + 2048] = m + np.log(np.exp(l - m[:, None]).sum(axis=1))
    return out


def compat_clamp(ll):
    """What the reference's linear-domain scorer returns for a frame whose total probability underflows
    float64 completely: safe_log(0) = log(1e-15) (gmm.cc:34-38).  Applied to exact log-domain values."""
    ll = np.array(ll, dtype=np.float64, copy=True)
    ll[ll < DENORM_LOG_MIN] = SAFE_LOG_FLOOR
    return ll


def em_iteration(w, mean, sigma, X, min_covar=1e-3, exact_exp=False):
    """one GMMTrainerBaseline::iteration; returns new (w, mean, sigma)."""
    w, mean, sigma, X = _c64(w).copy(), _c64(mean).copy(), _c64(sigma).copy(), _c64(X)
    K, D = mean.shape
    lib().orc_em_iteration(K, D, _dp(w), _dp(mean), _dp(sigma), _dp(X), len(X), min_covar, int(exact_exp))
    return w, mean, sigma


def map_iteration(w, mean, sigma, ubm_mean, X, exact_exp=False):
    """one GMMUBMTrainerBaseline iteration (means-only MAP, relevance 16)."""
    w, mean, sigma, X = _c64(w).copy(), _c64(mean).copy(), _c64(sigma).copy(), _c64(X)
    ubm_mean = _c64(ubm_mean)
    K, D = mean.shape
    lib().orc_map_iteration(K, D, _dp(w), _dp(mean), _dp(sigma), _dp(ubm_mean), _dp(X), len(X), int(exact_exp))
    return w, mean, sigma


# ---- text model format: GMM::dump/load gmm.cc:655-682, Gaussian::dump/load gmm.cc:101-150 ----

def load_model(path):
    """-> (w[K], mean[K,D], sigma[K,D]).  Format: K \\n w.. \\n then per Gaussian 'dim covtype' \\n mean.. \\n sigma.."""
    with open(path) as f:
        tok = f.read().split()
    it = iter(tok)
    K = int(next(it))
    w = np.array([float(next(it)) for _ in range(K)])
    means, sigmas = [], []
    for _ in range(K):
        D = int(next(it))
        cov = int(next(it))
        if cov != 1:
            raise ValueError("only diagonal covariance (1) supported, gmm.cc:211-215")
        means.append([float(next(it)) for _ in range(D)])
        sigmas.append([float(next(it)) for _ in range(D)])
    return w, np.array(means), np.array(sigmas)


def dump_model(path, w, mean, sigma, precision=6):
    """precision=6 reproduces the reference's default ostream formatting (gmm.cc:655-662, %g-like)."""
    fmt = "%%.%dg" % precision
    K, D = mean.shape
    with open(path, "w") as f:
        f.write("%d\n" % K)
        f.write("".join(fmt % v + " " for v in w) + "\n")
        for k in range(K):
            f.write("%d 1\n" % D)
            f.write("".join(fmt % v + " " for v in mean[k]) + "\n")
            f.write("".join(fmt % v + " " for v in sigma[k]) + "\n")
