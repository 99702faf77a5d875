"""Feature package front, mirroring src/feature/__init__.py:20-30.

The reference prefers `bob.ap` when importable and otherwise its own MFCC.py; BASELINE.json pins the own
MFCC.py, which is what the CUDA path implements.  `mix_feature` is MFCC || LPC (feature/__init__.py:25-30):
13 CMVN-normalised cepstra followed by 15 raw LPC coefficients per frame.
"""
import numpy as np

from . import LPC, MFCC  # noqa: F401
from .utils import diff_feature  # noqa: F401

MIX_FEATURE_HAS_LPC = True


def get_extractor(extract_func, **kwargs):
    def f(tup):
        return extract_func(*tup, **kwargs)
    return f


def mix_feature(tup):
    """(fs, signal) -> feature rows used by ModelInterface.enroll/predict (interface.py:59,90)"""
    mfcc = MFCC.extract(tup)
    lpc = LPC.extract(tup)
    return np.concatenate((mfcc, lpc), axis=1)
