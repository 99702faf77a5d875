"""speaker-recognition_b200: the MFCC + diagonal-GMM/UBM hot path of ppwwyyxx/speaker-recognition,
built B200-native (hand-written sm_100a CUDA behind a C ABI, include/srb200.h) and exposed through
the reference's own Python-facing API:

    feature.MFCC.extract(fs, signal=None, diff=False, **kw)     <- src/feature/MFCC.py:123-132
    feature.utils.diff_feature(feat, nd)                         <- src/feature/utils.py:24-31
    gmm.pygmm.GMM(...).fit/score/score_all/dump/load             <- src/gmm/python/pygmm.py:37-139
    gmmset.GMMSet / GMMSetPyGMM                                  <- src/testbench/gmmset.py:16-105
    interface.ModelInterface.enroll/train/predict/dump/load      <- src/gui/interface.py:28-109

There is no CPU fallback: importing `_lib` raises if csrc/libsrb200.so is missing.
"""
__version__ = "0.1.0"
