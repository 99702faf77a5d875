"""Batched device pipeline: PCM -> MFCC(+delta, CMVN) -> score against a speaker set -> labels, for many
utterances per call.  This is the throughput form of ModelInterface.predict (src/gui/interface.py:85-94):
same arithmetic, but B utterances per launch sequence and every intermediate stays in HBM.

PyTorch is used only as the owner of device buffers / streams (and torch.distributed for multi-GPU plumbing);
all compute goes through the C ABI (include/srb200.h: srb200_mfcc_extract_dev, srb200_gmmset_score_dev).
Multi-GPU: utterances are independent, so ranks simply take disjoint utterance shards -- no collective on the
data path (SURVEY.md s8e "Predict, many utterances").
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .feature import MFCC as _MFCC


class BatchScorer(object):
    """gmmset: a GMMSet/GMMSetPyGMM with enrolled models.  nd: 0/1/2 delta order (feature dim = n_ceps*(nd+1))."""

    KERNELS_PER_STEP = 5   # mfcc_frames, mfcc_stats, mfcc_norm, gmm_score, score_reduce

    def __init__(self, gmmset, fs=16000, nd=2, device=None, **mfcc_kwargs):
        import torch
        self.torch = torch
        _lib.require_gpu()
        self.L = _lib.lib()
        self.gmmset = gmmset
        self.S = self.L.srb200_gmmset_size(self._set())
        self.ex = _MFCC.get_mfcc_extractor(fs, **mfcc_kwargs)
        self.nd = nd
        self.dim = self.L.srb200_mfcc_out_dim(self.ex._h, nd)
        if self.dim != self.L.srb200_gmmset_dim(self._set()):
            raise ValueError("feature dim %d != model dim %d" % (self.dim, self.L.srb200_gmmset_dim(self._set())))
        self.ld = (self.dim + 3) & ~3
        self.device = torch.device("cuda", self.L.srb200_get_device()) if device is None else device
        self._shape = None

    def _set(self):
        """the GMMSet's packed device set, re-validated on every use: enrolling / loading / refitting a speaker after this scorer
        was built frees and rebuilds that set, and a cached raw pointer would dangle"""
        h = self.gmmset._ensure_set()
        S = self.L.srb200_gmmset_size(h)
        if getattr(self, "S", S) != S:
            self.S = S
            self._shape = None              # the sums buffer is sized by S
        return h

    def _prepare(self, lengths):
        torch = self.torch
        key = tuple(int(x) for x in lengths)
        if self._shape == key:
            return
        B = len(key)
        off = np.zeros(B + 1, dtype=np.int64)
        off[1:] = np.cumsum(key)
        self.smp_off = (C.c_longlong * (B + 1))(*off.tolist())
        self.n_samples = int(off[-1])
        rows = self.L.srb200_mfcc_rows_total(self.ex._h, self.smp_off, B, self.nd)
        if rows <= 0:
            raise ValueError("an utterance is too short for MFCC extraction")
        self.rows = int(rows)
        self.feat_off = (C.c_int * (B + 1))()
        # device / pinned buffers only ever GROW (with headroom): a stream of ragged batches stops allocating once it has seen its
        # largest batch; the tensors used below are views of the capacity buffers
        cap = getattr(self, "_cap", None)
        if cap is None or self.n_samples > cap["samples"] or self.rows > cap["rows"] or B > cap["B"] or self.S != cap["S"]:
            grow = lambda need, have: max(need + need // 4 + 64, have)
            self._cap = dict(samples=grow(self.n_samples, cap["samples"] if cap else 0), rows=grow(self.rows, cap["rows"] if cap else 0),
                             B=grow(B, cap["B"] if cap else 0), S=self.S)
            torch.cuda.synchronize(self.device)    # nothing in flight may still use the buffers that are about to go
            self._slots = {}
        self.B = B
        self._use_slot(0)
        self._shape = key
        # sub-batches for the end-to-end path: H2D of chunk c+1 overlaps the kernels of chunk c.
        # One step at a time ("sync"): a small first sub-batch starts the kernels early (its copy is the only one nothing overlaps),
        # later ones grow so the persistent kernels' waves stay full, and the LAST one is small again: its kernels are the only ones no
        # copy overlaps.  Two steps in flight ("pipe"): the next step's copies fill those gaps anyway, so few, large sub-batches win
        # (per-launch ramps and tails cost 0.1 ms per sub-batch).  Measured at C2 x 64 (profiles/r2_e2e_sweep.txt; the copy alone
        # takes 6.0 ms, the kernels 6.16): sync 12.2 / 9.5 / 8.2 / 7.5 ms per step with 1 / 2 / 4 / 8 sub-batches, pipe 6.47 / 6.43 /
        # 6.57 / 6.96.  SRB200_E2E_CHUNKS="f1,f2,..." (cumulative fractions) overrides both.
        env = os.environ.get("SRB200_E2E_CHUNKS")
        fr_sync = [1 / 32, 1 / 8, 0.25, 0.4, 0.55, 0.7, 0.85, 1.0] if B >= 32 else ([0.25, 0.5, 0.75, 1.0] if B >= 8 else [1.0])
        fr_pipe = [0.5, 1.0] if B >= 8 else [1.0]
        if env:
            fr_sync = fr_pipe = [float(x) for x in env.split(",")]
        self.chunk_sets = {"sync": self._make_chunks(fr_sync, off, key), "pipe": self._make_chunks(fr_pipe, off, key)}
        self.chunks = self.chunk_sets["sync"]
        if getattr(self, "copy_stream", None) is None:
            self.copy_stream = torch.cuda.Stream(device=self.device)
        self._direct_use = True

    def _make_chunks(self, fr, off, key):
        B = len(key)
        bounds = sorted(set([0] + [min(B, max(1, round(f * B))) for f in fr] + [B]))
        if B >= 32 and len(bounds) > 3 and os.environ.get("SRB200_E2E_ALIGN", "1") != "0":
            bounds = self._wave_align(bounds, key)
        chunks = []
        row0 = 0
        for c in range(len(bounds) - 1):
            u0, u1 = bounds[c], bounds[c + 1]
            rel = (off[u0:u1 + 1] - off[u0]).tolist()
            smp = (C.c_longlong * (u1 - u0 + 1))(*rel)
            rows_c = int(self.L.srb200_mfcc_rows_total(self.ex._h, smp, u1 - u0, self.nd))
            chunks.append(dict(u0=u0, u1=u1, s0=int(off[u0]), s1=int(off[u1]), smp=smp, row0=row0, rows=rows_c,
                               feat_off=(C.c_int * (u1 - u0 + 1))()))
            row0 += rows_c
        return chunks

    def _slot(self, k):
        """buffers of pipeline slot k at the current capacity (slot 0: the direct API; slots 0/1 alternate in submit())"""
        torch = self.torch
        sl = self._slots.get(k)
        if sl is None:
            cap = self._cap
            sl = dict(pcm=torch.empty(cap["samples"], dtype=torch.int16, device=self.device),
                      feat=torch.empty((cap["rows"], self.ld), dtype=torch.float32, device=self.device),
                      sums=torch.empty((cap["B"], self.S), dtype=torch.float64, device=self.device),
                      best=torch.empty(cap["B"], dtype=torch.int32, device=self.device),
                      hbest=torch.empty(cap["B"], dtype=torch.int32).pin_memory(),
                      done=torch.cuda.Event(), copied=[], used=False, B=0)
            self._slots[k] = sl
        return sl

    def _use_slot(self, k):
        """point the views the stages work on (d_pcm, d_feat, d_sums, d_best, h_best) at slot k"""
        sl = self._slot(k)
        self.d_pcm = sl["pcm"][:self.n_samples]
        self.d_feat = sl["feat"][:self.rows]
        self.d_sums = sl["sums"][:self.B]
        self.d_best = sl["best"][:self.B]
        self.h_best = sl["hbest"][:self.B]
        return sl

    def _wave_align(self, bounds, lengths):
        """move the sub-batch boundaries by up to two utterances so that every sub-batch fills whole waves of the persistent scorer:
        a sub-batch is its own launch, its work units are PAIRS of 128-frame tiles dealt to the chip's CTA pairs (gmm_score_tc2.cu), and
        a last, partly filled wave idles most of the chip (ten 10 000-frame utterances: 5.3 waves -> 6)"""
        one = (C.c_longlong * 2)()
        tiles = []
        for n in lengths:
            one[0], one[1] = 0, n
            tiles.append((int(self.L.srb200_mfcc_rows_total(self.ex._h, one, 1, self.nd)) + 127) // 128)
        ct = np.concatenate([[0], np.cumsum(tiles)])
        clusters = max(1, self.torch.cuda.get_device_properties(self.device).multi_processor_count // 2)

        def idle(u0, u1):
            w = ((ct[u1] - ct[u0] + 1) // 2) / clusters
            return 0.0 if w <= 1.0 else float(np.ceil(w) - w)
        out = [0]
        B = len(lengths)
        for b in bounds[1:-1]:
            cands = [x for x in range(b - 2, b + 3) if out[-1] < x < B]
            if cands:
                out.append(min(cands, key=lambda x: (idle(out[-1], x), abs(x - b))))
        out.append(B)
        return sorted(set(out))

    # ---- stages (each enqueues on the current torch stream) ----
    def upload(self, pcm_host):
        """pcm_host: pinned int16 torch tensor holding the utterances back to back"""
        self._direct_use = True
        self.d_pcm.copy_(pcm_host, non_blocking=True)

    def features(self):
        self._direct_use = True
        st = self.torch.cuda.current_stream().cuda_stream
        _lib.check(self.L.srb200_mfcc_extract_dev(self.ex._h, self.d_pcm.data_ptr(), self.smp_off, self.B, self.nd,
                                                  self.d_feat.data_ptr(), self.ld, self.feat_off, st), "mfcc_extract_dev")

    def score(self):
        self._direct_use = True
        st = self.torch.cuda.current_stream().cuda_stream
        _lib.check(self.L.srb200_gmmset_score_dev(self._set(), self.d_feat.data_ptr(), self.ld, self.dim, self.feat_off, self.B,
                                                  self.d_sums.data_ptr(), None, 0, self.d_best.data_ptr(), st), "gmmset_score_dev")

    def step_device(self):
        """one pass of the hot path with the PCM already resident in HBM"""
        self.features()
        self.score()

    def submit(self, pcm_host, lengths):
        """enqueue one end-to-end step -- pinned host PCM in, speaker indices out -- and return a ticket for result().  The batch
        is cut into sub-batches of utterances: the H2D copy of sub-batch c+1 (copy stream) overlaps the kernels of sub-batch c.
        Two buffer slots alternate, so a caller that submits step i+1 before collecting step i also overlaps the first copies of
        step i+1 with the last kernels of step i (the copy engine and the SMs are then both busy all the time)."""
        torch = self.torch
        set_h = self._set()
        self._prepare(lengths)
        k = getattr(self, "_next", 0)
        self._next = k ^ 1
        other = self._slots.get(k ^ 1)
        in_flight = other is not None and other["used"] and not other["done"].query()      # the caller keeps two steps in flight
        self.chunks = self.chunk_sets["pipe" if in_flight else "sync"]
        sl = self._use_slot(k)
        cur = torch.cuda.current_stream()
        if self._direct_use:
            self.copy_stream.wait_stream(cur)          # direct users of the buffers (upload / features / score) are done
            self._direct_use = False
        elif sl["used"]:
            self.copy_stream.wait_event(sl["done"])    # the step that used this slot before (two submits ago) has finished
        while len(sl["copied"]) < len(self.chunks):
            sl["copied"].append(torch.cuda.Event())
        with torch.cuda.stream(self.copy_stream):
            for c, ev in zip(self.chunks, sl["copied"]):
                self.d_pcm[c["s0"]:c["s1"]].copy_(pcm_host[c["s0"]:c["s1"]], non_blocking=True)
                ev.record(self.copy_stream)
        st = cur.cuda_stream
        for c, ev in zip(self.chunks, sl["copied"]):
            cur.wait_event(ev)
            nu = c["u1"] - c["u0"]
            _lib.check(self.L.srb200_mfcc_extract_dev(self.ex._h, self.d_pcm.data_ptr() + 2 * c["s0"], c["smp"], nu, self.nd,
                                                      self.d_feat.data_ptr() + 4 * self.ld * c["row0"], self.ld, c["feat_off"], st),
                       "mfcc_extract_dev")
            _lib.check(self.L.srb200_gmmset_score_dev(set_h, self.d_feat.data_ptr() + 4 * self.ld * c["row0"], self.ld, self.dim,
                                                      c["feat_off"], nu, self.d_sums.data_ptr() + 8 * self.S * c["u0"], None, 0,
                                                      self.d_best.data_ptr() + 4 * c["u0"], st), "gmmset_score_dev")
        self.h_best.copy_(self.d_best, non_blocking=True)
        sl["done"].record(cur)
        sl["used"] = True
        sl["B"] = self.B
        return sl

    def result(self, ticket):
        """wait for a submitted step; returns its speaker indices (pinned host int32 tensor, valid until the slot's next submit)"""
        sl = ticket                                     # the slot itself: stays valid if a later, larger batch re-allocates the slots
        sl["done"].synchronize()
        return sl["hbest"][:sl["B"]]

    def predict_pcm(self, pcm_host, lengths):
        """end to end and synchronous: pinned host PCM in, speaker indices (host int32 tensor, pinned) out"""
        return self.result(self.submit(pcm_host, lengths))

    def predict(self, signals):
        """convenience: list of int16 numpy signals -> list of labels"""
        torch = self.torch
        lengths = [len(s) for s in signals]
        self._prepare(lengths)
        host = torch.from_numpy(np.concatenate([np.asarray(s, dtype=np.int16) for s in signals])).pin_memory()
        idx = self.predict_pcm(host, lengths)
        return [self.gmmset.y[int(i)] for i in idx]

    def frames_per_step(self):
        return self.rows
