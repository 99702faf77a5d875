"""Speaker-set classifiers -- a Python 3 mirror of src/testbench/gmmset.py:16-105 (GMMSet, GMMSetPyGMM; also the
shape of src/gui/skgmm.py:11-39) on top of the C ABI.

The reference scores an utterance with a Python loop over speakers, one `.so` call (and one full re-marshalling
of X) per speaker (gmmset.py:59-64, 95-99).  Here all enrolled models live packed on the device in one
`srb200_gmmset` and every predict is ONE kernel launch over (frames x speakers x mixtures).
"""
import ctypes as C
import operator
from collections import defaultdict

import numpy as np

from . import _lib
from .gmm.pygmm import GMM


class GMMSet(object):
    def __init__(self, gmm_order=32, ubm=None, reject_threshold=10, **kwargs):
        self.kwargs = kwargs
        self.gmms = []
        self.ubm = ubm
        self.reject_threshold = reject_threshold
        if ubm is not None:
            self.gmm_order = ubm.get_nr_mixtures()
        else:
            self.gmm_order = gmm_order
        self.y = []
        self._set = None          # device-packed models (srb200_gmmset*), rebuilt lazily
        self._set_ids = None

    # ---- device set management ----
    def _drop_set(self):
        if getattr(self, "_set", None):
            _lib.lib().srb200_gmmset_free(self._set)
        self._set = None
        self._set_ids = None

    def __del__(self):
        try:
            self._drop_set()
        except Exception:
            pass

    def _ensure_set(self):
        """the device-packed set: every enrolled model and, when a UBM is attached, the UBM as one extra model at the end (so that
        predict_one_with_rejection is ONE scoring launch).  The C side copies the parameters at creation, so the cache key carries
        every model's identity AND its parameter version (a refit / set_params on an enrolled GMM rebuilds the set)."""
        models = list(self.gmms) + ([self.ubm] if self.ubm is not None else [])
        ids = tuple((id(g), getattr(g, "_version", 0)) for g in models)
        if self._set is not None and ids == self._set_ids:
            return self._set
        self._drop_set()
        arr = (C.c_void_p * len(models))(*[g.gmm for g in models])
        self._set = _lib.check_ptr(_lib.lib().srb200_gmmset_create(arr, len(models)), "GMMSet")
        if self.ubm is not None and self.gmms:
            _lib.check(_lib.lib().srb200_gmmset_set_ranked(self._set, len(self.gmms)), "GMMSet")      # the UBM is scored, never ranked
        self._set_ids = ids
        return self._set

    def _score_all_models(self, x):
        """[sum_t log p_s(x_t) for the enrolled speakers ..., (the UBM's, if one is attached)] in one launch"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = len(self.gmms) + (1 if self.ubm is not None else 0)
        sums = np.empty(n, dtype=np.float64)
        _lib.check(_lib.lib().srb200_gmmset_score_f64(self._ensure_set(), _lib.dptr(x), x.shape[0], x.shape[1], _lib.dptr(sums)),
                   "GMMSet.predict_one_scores")
        return sums

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_set"] = None
        st["_set_ids"] = None
        return st

    # ---- enrolment ----
    def fit_new(self, x, label):
        self.y.append(label)
        gmm = GMM(self.gmm_order, **self.kwargs)
        gmm.fit(x, self.ubm)
        self.gmms.append(gmm)

    def fit_many(self, xs, labels):
        """enrol several speakers at once.  With a UBM this is ONE batched MAP adaptation (GMM.fit_many_from_ubm); without one it is
        the reference's loop of independent EM trainings (gmmset.py:30-34)."""
        if self.ubm is None or not len(labels):
            for x, label in zip(xs, labels):
                self.fit_new(x, label)
            return
        gmms = [GMM(self.gmm_order, **self.kwargs) for _ in labels]
        GMM.fit_many_from_ubm(gmms, xs, self.ubm)
        self.y.extend(labels)
        self.gmms.extend(gmms)

    def cluster_by_label(self, X, y):
        Xtmp = defaultdict(list)
        for ind, x in enumerate(X):
            label = y[ind]
            Xtmp[label].extend(x)
        yp, Xp = zip(*Xtmp.items())
        return Xp, yp

    def auto_tune_parameter(self, X, y):
        if self.ubm is None:
            return
        # TODO in the reference too (gmmset.py:44-47)

    def fit(self, X, y):
        X, y = self.cluster_by_label(X, y)
        self.fit_many(X, y)
        self.auto_tune_parameter(X, y)

    # ---- scoring ----
    def gmm_score(self, gmm, x):
        return np.sum(gmm.score(x))

    def predict_one_scores(self, x):
        """sum_t log p_s(x_t) for every enrolled speaker s, one launch (gmmset.py:59-60)"""
        if not self.gmms:
            return []
        return list(self._score_all_models(x)[:len(self.gmms)])

    def predict_one(self, x):
        scores = self.predict_one_scores(x)
        return self.y[max(enumerate(scores), key=operator.itemgetter(1))[0]]

    def predict(self, X):
        return list(map(self.predict_one, X))

    def predict_one_with_rejection(self, x):
        """gmmset.py:69-81: reject when the best speaker's mean log-likelihood is less than reject_threshold above the UBM's.
        The UBM rides in the same launch as model S + 1 (the reference makes a second pass over x for it)."""
        assert self.ubm is not None, "UBM must be given prior to conduct reject prediction."
        allsc = self._score_all_models(x)
        x_len = len(x)  # normalize score
        scores = [v / x_len for v in allsc[:len(self.gmms)]]
        max_tup = max(enumerate(scores), key=operator.itemgetter(1))
        ubm_score = allsc[len(self.gmms)] / x_len
        if max_tup[1] - ubm_score < self.reject_threshold:
            return None
        return self.y[max_tup[0]]

    def predict_with_reject(self, X):
        return list(map(self.predict_one_with_rejection, X))

    def load_gmm(self, label, fname):
        self.y.append(label)
        gmm = GMM.load(fname)
        for key, val in self.kwargs.items():
            setattr(gmm, key, val)
        self.gmms.append(gmm)

    def before_pickle(self):
        pass

    def after_pickle(self):
        pass


class GMMSetPyGMM(GMMSet):
    def predict_one(self, x):
        scores = [s / len(x) for s in self.predict_one_scores(x)]     # score_all(x) / len(x), gmmset.py:96
        return self.y[max(enumerate(scores), key=operator.itemgetter(1))[0]]

    def before_pickle(self):
        self._drop_set()
        self.gmms = [x.dumps() for x in self.gmms]

    def after_pickle(self):
        self.gmms = [GMM.loads(x) for x in self.gmms]
