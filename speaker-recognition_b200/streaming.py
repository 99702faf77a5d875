"""Real-time "conversation mode" predictor for many concurrent streams (BASELINE config 5).

Semantics follow the reference GUI (src/gui/gui.py:46-47,169-190): on every tick the trailing WINDOW seconds of each
stream are re-extracted (MFCC + CMVN over the window only) and re-scored against all enrolled speakers -- there is no
incremental state in the reference and none here, so results equal ModelInterface.predict on the same window.  What
changes is the mechanics: all streams share ONE launch sequence per tick (H2D of the new samples, one window-slide
launch (srb200_stream_push_i16), srb200_mfcc_extract_dev over the B windows as B utterances, srb200_gmmset_score_dev, D2H of B labels).
A window must exceed 5*FRAME_LEN samples or the extractor refuses it (MFCC.py:56), as in the reference.
"""
import ctypes as C

import numpy as np

from . import _lib
from .feature import MFCC as _MFCC


class StreamingPredictor(object):
    def __init__(self, gmmset, n_streams=64, fs=16000, window_s=1.5, nd=2, **mfcc_kwargs):
        import torch
        self.torch = torch
        _lib.require_gpu()
        self.L = _lib.lib()
        self.gmmset = gmmset
        self.S = self.L.srb200_gmmset_size(gmmset._ensure_set())
        self.ex = _MFCC.get_mfcc_extractor(fs, **mfcc_kwargs)
        self.nd = nd
        self.dim = self.L.srb200_mfcc_out_dim(self.ex._h, nd)
        if self.dim != self.L.srb200_gmmset_dim(gmmset._ensure_set()):
            raise ValueError("feature dim %d != model dim %d" % (self.dim, self.L.srb200_gmmset_dim(gmmset._ensure_set())))
        self.ld = (self.dim + 3) & ~3
        self.B = int(n_streams)
        self.W = int(round(window_s * fs))
        assert self.W > 5 * self.ex.FRAME_LEN, "window too short (MFCC.py:56)"
        dev = torch.device("cuda", self.L.srb200_get_device())
        self.ring = torch.zeros((self.B, self.W), dtype=torch.int16, device=dev)
        self.tmp = torch.empty_like(self.ring)
        off = (np.arange(self.B + 1, dtype=np.int64) * self.W).tolist()
        self.smp_off = (C.c_longlong * (self.B + 1))(*off)
        rows = int(self.L.srb200_mfcc_rows_total(self.ex._h, self.smp_off, self.B, nd))
        self.feat = torch.empty((rows, self.ld), dtype=torch.float32, device=dev)
        self.feat_off = (C.c_int * (self.B + 1))()
        self.sums = torch.empty((self.B, self.S), dtype=torch.float64, device=dev)
        self.best = torch.empty(self.B, dtype=torch.int32, device=dev)
        self.h_best = torch.empty(self.B, dtype=torch.int32).pin_memory()
        self.h_new = None
        self.h2d_done = torch.cuda.Event()      # the previous tick's H2D copy has read the pinned staging buffer
        self.filled = 0

    def push(self, chunk):
        """chunk: int16 array [n_streams, hop] of new samples (host).  Returns int32 labels [n_streams] (indices into
        gmmset.y) once a full window has been seen, else None."""
        torch = self.torch
        chunk = np.ascontiguousarray(chunk, dtype=np.int16)
        hop = chunk.shape[1]
        assert chunk.shape[0] == self.B and 0 < hop < self.W, "need [n_streams, hop] samples with 0 < hop < window"
        if self.h_new is None or self.h_new.shape[1] != hop:
            torch.cuda.current_stream().synchronize()
            self.h_new = torch.empty((self.B, hop), dtype=torch.int16).pin_memory()
            self.d_new = torch.empty((self.B, hop), dtype=torch.int16, device=self.ring.device)
        else:
            self.h2d_done.synchronize()         # warm-up ticks return without a sync: do not overwrite a staging buffer still being read
        self.h_new.numpy()[:] = chunk
        self.d_new.copy_(self.h_new, non_blocking=True)
        self.h2d_done.record()
        st = torch.cuda.current_stream().cuda_stream
        # slide every window and append the new samples: one launch, [old[hop:], new]
        _lib.check(self.L.srb200_stream_push_i16(self.ring.data_ptr(), self.d_new.data_ptr(), self.tmp.data_ptr(), self.B, self.W, hop, st),
                   "stream_push")
        self.ring, self.tmp = self.tmp, self.ring
        self.filled = min(self.W, self.filled + hop)
        if self.filled < self.W:
            return None
        set_h = self.gmmset._ensure_set()       # re-validated every tick: enrolling a speaker rebuilds the packed set
        if self.L.srb200_gmmset_size(set_h) != self.S:
            self.S = self.L.srb200_gmmset_size(set_h)
            self.sums = torch.empty((self.B, self.S), dtype=torch.float64, device=self.ring.device)
        _lib.check(self.L.srb200_mfcc_extract_dev(self.ex._h, self.ring.data_ptr(), self.smp_off, self.B, self.nd,
                                                  self.feat.data_ptr(), self.ld, self.feat_off, st), "mfcc_extract_dev")
        _lib.check(self.L.srb200_gmmset_score_dev(set_h, self.feat.data_ptr(), self.ld, self.dim, self.feat_off, self.B,
                                                  self.sums.data_ptr(), None, 0, self.best.data_ptr(), st), "gmmset_score_dev")
        self.h_best.copy_(self.best, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h_best.numpy().copy()

    def labels(self, idx):
        return [self.gmmset.y[int(i)] for i in idx]
