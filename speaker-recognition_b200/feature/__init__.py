"""Feature package front, mirroring src/feature/__init__.py:20-30.

The reference prefers `bob.ap` when importable and otherwise its own MFCC.py; BASELINE.json pins the own
MFCC.py, which is what the CUDA path implements.  `mix_feature` in the reference is MFCC || LPC
(feature/__init__.py:25-30); the LPC half depends on the un-vendored scikits.talkbox and is the first
"next" row of SURVEY.md s8(f) -- until it is built, mix_feature returns the MFCC block only and says so.
"""
from . import MFCC  # noqa: F401
from .utils import diff_feature  # noqa: F401

MIX_FEATURE_HAS_LPC = False


def get_extractor(extract_func, **kwargs):
    def f(tup):
        return extract_func(*tup, **kwargs)
    return f


def mix_feature(tup):
    """(fs, signal) -> feature rows used by ModelInterface.enroll/predict (interface.py:59,90)."""
    return MFCC.extract(tup)
