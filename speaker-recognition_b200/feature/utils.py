"""Feature helpers with the reference's names and semantics (src/feature/utils.py)."""
import numpy

kwd_mark = object()


def cached_func(function):
    """memoise per (args, sorted kwargs) -- src/feature/utils.py:11-21"""
    cache = {}

    def wrapper(*args, **kwargs):
        key = args + (kwd_mark,) + tuple(sorted(kwargs.items()))
        if key not in cache:
            cache[key] = function(*args, **kwargs)
        return cache[key]
    wrapper.cache = cache
    return wrapper


def diff_feature(feat, nd=1):
    """delta (nd=1) / delta + delta-delta (nd=2) by plain first differences; rows shrink by nd
    (src/feature/utils.py:24-31).  Host-side convenience for arrays already on the host; the fused device
    path is MFCC.extract(..., nd=...) / the srb200_mfcc_* entry points."""
    feat = numpy.asarray(feat)
    diff = feat[1:] - feat[:-1]
    feat = feat[1:]
    if nd == 1:
        return numpy.concatenate((feat, diff), axis=1)
    elif nd == 2:
        d2 = diff[1:] - diff[:-1]
        return numpy.concatenate((feat[1:], diff[1:], d2), axis=1)
