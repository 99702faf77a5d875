"""sklearn-shaped GMM over the C ABI -- a Python 3 mirror of src/gmm/python/pygmm.py:37-139.

Same constructor defaults (pygmm.py:39-46), same methods: fit(X, ubm=None), score(X) -> ndarray[N],
score_all(X) -> float, dump/load/dumps/loads, get_dim, get_nr_mixtures.  Differences, all deliberate
(SURVEY.md s7 hard part 6): X goes down as ONE contiguous float64 matrix instead of one ctypes array per
row (pygmm.py:89-95); handles are freed; load() keeps the Python-side parameters; dumps() is lossless.
"""
import ctypes as C
import os
import tempfile
from multiprocessing import cpu_count

import numpy as np

from .. import _lib
from .._lib import Parameter as GMMParameter

COVTYPE_SPHEREICAL, COVTYPE_DIAGONAL, COVTYPE_FULL = 0, 1, 2   # (sic) pygmm.py:33-35

_PARAM_NAMES = ["nr_mixture", "min_covar", "threshold", "nr_iteration", "init_with_kmeans", "concurrency", "verbosity"]


class GMM(object):
    def __init__(self, nr_mixture=10, covariance_type=COVTYPE_DIAGONAL, min_covar=1e-3, threshold=0.01,
                 nr_iteration=200, init_with_kmeans=0, concurrency=cpu_count(), verbosity=0, seed=None):
        self.nr_mixture = nr_mixture
        self.covariance_type = covariance_type
        self.min_covar = min_covar
        self.threshold = threshold
        self.nr_iteration = nr_iteration
        self.init_with_kmeans = init_with_kmeans
        self.concurrency = concurrency          # accepted for compatibility; the GPU path ignores it
        self.verbosity = verbosity
        self.seed = seed                        # None: per-process call counter (the reference seeds from rand())
        self.last_iterations = None
        self.last_log_likelihood = None
        self._version = 0                       # bumped whenever the parameters change (GMMSet keys its packed device copy on it)
        self.gmm = _lib.check_ptr(_lib.lib().new_gmm(int(nr_mixture), int(covariance_type)), "GMM")

    # ---- lifetime ----
    def __del__(self):
        try:
            if getattr(self, "gmm", None):
                _lib.lib().free_gmm(self.gmm)
                self.gmm = None
        except Exception:
            pass

    def __getstate__(self):
        st = {k: getattr(self, k) for k in _PARAM_NAMES + ["covariance_type", "seed"]}
        st["model_text"] = self.dumps() if self.get_dim() > 0 else None
        return st

    def __setstate__(self, st):
        txt = st.pop("model_text")
        self.__dict__.update(st)
        self.last_iterations = self.last_log_likelihood = None
        self._version = 0
        if txt is None:
            self.gmm = _lib.check_ptr(_lib.lib().new_gmm(int(self.nr_mixture), int(self.covariance_type)), "GMM")
        else:
            self.gmm = GMM._load_handle_from_text(txt)

    # ---- I/O (text format of gmm.cc:655-682, 101-150) ----
    @staticmethod
    def load(model_file):
        h = _lib.check_ptr(_lib.lib().load(os.fsencode(model_file)), "GMM.load")
        g = GMM.__new__(GMM)
        GMM._init_defaults(g)
        g.gmm = h
        g.nr_mixture = _lib.lib().get_nr_mixtures(h)
        return g

    @staticmethod
    def _init_defaults(g):
        g.covariance_type = COVTYPE_DIAGONAL
        g.min_covar, g.threshold, g.nr_iteration = 1e-3, 0.01, 200
        g.init_with_kmeans, g.concurrency, g.verbosity, g.seed = 0, cpu_count(), 0, None
        g.nr_mixture = 10
        g.last_iterations = g.last_log_likelihood = None
        g._version = 0

    def dump(self, model_file, precision=6):
        """precision=6 is the reference's ostream default (lossy); 17 is lossless."""
        if precision == 6:
            _lib.lib().dump(self.gmm, os.fsencode(model_file))
            err = _lib.last_error()
            if err:
                raise _lib.SrbError("GMM.dump: " + err)
        else:
            _lib.check(_lib.lib().srb200_gmm_dump_precise(self.gmm, os.fsencode(model_file), int(precision)), "GMM.dump")

    def dumps(self):
        fd, tmp = tempfile.mkstemp(suffix=".gmm")
        os.close(fd)
        try:
            self.dump(tmp, precision=17)
            with open(tmp, "r") as f:
                return f.read()
        finally:
            os.unlink(tmp)

    @staticmethod
    def _load_handle_from_text(s):
        fd, tmp = tempfile.mkstemp(suffix=".gmm")
        try:
            with os.fdopen(fd, "wb") as f:
                f.write(s.encode() if isinstance(s, str) else s)
            return _lib.check_ptr(_lib.lib().load(os.fsencode(tmp)), "GMM.loads")
        finally:
            os.unlink(tmp)

    @staticmethod
    def loads(s):
        g = GMM.__new__(GMM)
        GMM._init_defaults(g)
        g.gmm = GMM._load_handle_from_text(s)
        g.nr_mixture = _lib.lib().get_nr_mixtures(g.gmm)
        return g

    # ---- parameters (not in the reference binding; its C++ members are public) ----
    def set_params(self, weights, means, sigmas):
        w, m, s = _lib.as_f64(weights), _lib.as_f64(means), _lib.as_f64(sigmas)
        K, D = m.shape
        assert w.shape == (K,) and s.shape == (K, D)
        _lib.check(_lib.lib().srb200_gmm_set_params(self.gmm, K, D, _lib.dptr(w), _lib.dptr(m), _lib.dptr(s)), "GMM.set_params")
        self.nr_mixture = K
        self._version += 1
        return self

    def get_params(self):
        K, D = self.get_nr_mixtures(), self.get_dim()
        w, m, s = np.empty(K), np.empty((K, D)), np.empty((K, D))
        _lib.check(_lib.lib().srb200_gmm_get_params(self.gmm, _lib.dptr(w), _lib.dptr(m), _lib.dptr(s)), "GMM.get_params")
        return w, m, s

    # ---- training / scoring ----
    def _gen_param(self, X):
        p = GMMParameter()
        for name in _PARAM_NAMES:
            setattr(p, name, type(getattr(p, name))(getattr(self, name)))
        p.nr_instance = len(X)
        p.nr_dim = len(X[0]) if len(X) else 0
        return p

    @staticmethod
    def _matrix(X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("X must be a 2-D array-like of frames")
        return X

    def fit(self, X, ubm=None):
        """:param ubm is None or a GMM instance (MAP adaptation of the means, gmmubm.cc)"""
        X = self._matrix(X)
        param = self._gen_param(X)
        it, ll = C.c_int(0), C.c_double(0.0)
        L = _lib.lib()
        if ubm is None:
            seed = self.seed if self.seed is not None else GMM._next_seed()
            rc = L.srb200_train_f64(self.gmm, _lib.dptr(X), C.byref(param), C.c_uint64(seed), C.byref(it), C.byref(ll))
        else:
            rc = L.srb200_train_from_ubm_f64(self.gmm, ubm.gmm, _lib.dptr(X), C.byref(param), C.byref(it), C.byref(ll))
        _lib.check(rc, "GMM.fit")
        self.nr_mixture = L.get_nr_mixtures(self.gmm)
        self.last_iterations, self.last_log_likelihood = it.value, ll.value
        self._version += 1
        return self

    @staticmethod
    def fit_many_from_ubm(gmms, Xs, ubm):
        """MAP-adapt `ubm` to every X in Xs at once (gmms[i] receives the model of Xs[i]): one host->device copy of all frames, the
        speakers' EM kernels back to back, one synchronisation per log-likelihood check for all speakers (C ABI:
        srb200_train_many_from_ubm_f64).  Same result as [g.fit(X, ubm) for g, X in zip(gmms, Xs)]; the training parameters
        (nr_iteration, threshold, min_covar, ...) are taken from gmms[0]."""
        assert len(gmms) == len(Xs) and len(gmms) >= 1
        mats = [GMM._matrix(X) for X in Xs]
        D = mats[0].shape[1]
        off = np.zeros(len(mats) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(m) for m in mats])
        X = np.ascontiguousarray(np.concatenate(mats, axis=0))
        param = gmms[0]._gen_param(X)
        arr = (C.c_void_p * len(gmms))(*[g.gmm for g in gmms])
        its = (C.c_int * len(gmms))()
        lls = (C.c_double * len(gmms))()
        L = _lib.lib()
        _lib.check(L.srb200_train_many_from_ubm_f64(arr, len(gmms), ubm.gmm, _lib.dptr(X), off.ctypes.data_as(C.POINTER(C.c_int)), D,
                                                    C.byref(param), its, lls), "GMM.fit_many_from_ubm")
        for i, g in enumerate(gmms):
            g.nr_mixture = L.get_nr_mixtures(g.gmm)
            g.last_iterations, g.last_log_likelihood = its[i], lls[i]
            g._version += 1
        return gmms

    _seed_counter = [0]

    @staticmethod
    def _next_seed():
        GMM._seed_counter[0] += 1
        return 0x5EED + GMM._seed_counter[0]

    def score(self, X):
        """per-frame log-likelihoods, ndarray[N] (pygmm.py:120-126)"""
        X = self._matrix(X)
        prob = np.empty(len(X), dtype=np.float64)
        _lib.check(_lib.lib().srb200_score_batch_f64(self.gmm, _lib.dptr(X), X.shape[0], X.shape[1], _lib.dptr(prob)), "GMM.score")
        return prob

    def score_all(self, X):
        """sum of the per-frame log-likelihoods (pygmm.py:128-132)"""
        X = self._matrix(X)
        out = C.c_double(0.0)
        _lib.check(_lib.lib().srb200_score_all_f64(self.gmm, _lib.dptr(X), X.shape[0], X.shape[1], C.byref(out)), "GMM.score_all")
        return out.value

    def get_dim(self):
        return _lib.lib().get_dim(self.gmm)

    def get_nr_mixtures(self):
        return _lib.lib().get_nr_mixtures(self.gmm)
