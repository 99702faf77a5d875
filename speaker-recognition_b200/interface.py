"""Model facade -- a Python 3 mirror of src/gui/interface.py:28-109 (ModelInterface.enroll / train / predict /
dump / load) wired to the CUDA feature and GMM paths.  VAD (init_noise/filter, interface.py:37-53) depends on the
un-vendored `pyssp` and is out of scope; those two methods raise.
"""
import os
import pickle
import sys
import time
import traceback as tb
from collections import defaultdict

import numpy as np

from .feature import mix_feature
from .gmm.pygmm import GMM
from .gmmset import GMMSetPyGMM as GMMSet

CHECK_ACTIVE_INTERVAL = 1       # seconds


class ModelInterface(object):

    UBM_MODEL_FILE = None

    def __init__(self):
        self.features = defaultdict(list)
        self.gmmset = GMMSet()

    def init_noise(self, fs, signal):
        raise NotImplementedError("VAD (pyssp LTSD, src/filters/ltsd.py) is outside the B200 hot path")

    def filter(self, fs, signal):
        raise NotImplementedError("VAD (pyssp LTSD, src/filters/ltsd.py) is outside the B200 hot path")

    def enroll(self, name, fs, signal):
        """add the signal to this person's training dataset (interface.py:55-61)"""
        feat = mix_feature((fs, signal))
        self.features[name].extend(feat)

    def _get_gmm_set(self):
        if self.UBM_MODEL_FILE and os.path.isfile(self.UBM_MODEL_FILE):
            return GMMSet(ubm=GMM.load(self.UBM_MODEL_FILE))
        return GMMSet()

    def train(self, quiet=False):
        self.gmmset = self._get_gmm_set()
        start = time.time()
        if not quiet:
            print("Start training...")
        # the reference fits one name after the other (interface.py:81-82); with a UBM all names are MAP-adapted in ONE batched
        # launch sequence (GMMSet.fit_many), without one it is the same loop of independent EM trainings
        names = list(self.features.keys())
        self.gmmset.fit_many([np.asarray(self.features[n]) for n in names], names)
        if not quiet:
            print(time.time() - start, " seconds")

    def predict(self, fs, signal):
        """return a label (name); None when feature extraction fails (interface.py:85-94)"""
        try:
            feat = mix_feature((fs, signal))
        except Exception:
            print(tb.format_exc(), file=sys.stderr)
            return None
        return self.gmmset.predict_one(feat)

    def dump(self, fname):
        """dump all models to file"""
        self.gmmset.before_pickle()
        with open(fname, "wb") as f:
            pickle.dump(self, f, -1)
        self.gmmset.after_pickle()

    @staticmethod
    def load(fname):
        """load from a dumped model file"""
        with open(fname, "rb") as f:
            R = pickle.load(f)
            R.gmmset.after_pickle()
            return R
