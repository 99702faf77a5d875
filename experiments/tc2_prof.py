"""Cycle counters of the tensor scorer's warp roles (CTA pair 0), from a library built with -DSRB_T2_PROF=1
(experiments/ab_build.py prof -DSRB_T2_PROF=1; SRB200_LIB=experiments/ab/libsrb200_prof.so python experiments/tc2_prof.py)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from speaker_recognition_b200 import _lib, synth  # noqa: E402
from speaker_recognition_b200.gmm.pygmm import GMM  # noqa: E402
from speaker_recognition_b200.gmmset import GMMSet  # noqa: E402

S, K, D, utts, frames = 100, 128, 39, 64, 10000
L = _lib.lib()
W, M, Sg = synth.models(S, K, D, seed0=1000)
gs = GMMSet(gmm_order=K)
for i in range(S):
    gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
    gs.y.append(i)
ld = (D + 3) & ~3
n = utts * frames
X = torch.zeros((n, ld), dtype=torch.float32, device="cuda")
for u in range(utts):
    s = (u * 7) % S
    X[u * frames:(u + 1) * frames, :D] = torch.from_numpy(synth.draw(W[s], M[s], Sg[s], frames, seed=100 + u).astype(np.float32)).cuda()
off = (C.c_int * (utts + 1))(*[u * frames for u in range(utts + 1)])
sums = torch.zeros((utts, S), dtype=torch.float64, device="cuda")
best = torch.zeros(utts, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
h = gs._ensure_set()
buf = (C.c_ulonglong * (2 * 20 * 8))()
fn = C.CDLL(_lib.LIB_PATH).srb200_debug_tc2_prof
for it in range(3):
    _lib.check(L.srb200_gmmset_score_dev(h, X.data_ptr(), ld, D, off, utts, sums.data_ptr(), None, 0, best.data_ptr(), st))
    assert fn(buf) == 0
a = np.array(buf[:], dtype=np.float64).reshape(2, 20, 8)
FINE = os.environ.get("T2_PROF_FINE") == "1"
names = {"epi": ["wait+ldtm", "fence+arrive tm_empty", "add-back", "arrive empty_h+rings+mask", "lse", "handoff: fin_empty wait", "TOTAL", "handoff: rest"] if FINE else ["wait tm_full", "ldtm", "addback/arrive/mask", "lse", "handoff", "phi+tail", "TOTAL", "blocks"],
         "fin": ["wait fin_full", "process", "loop", "-", "-", "-", "TOTAL", "-"],
         "prod": ["wait empty_h", "wait empty_b", "issue", "-", "-", "-", "TOTAL", "-"],
         "mma": ["wait a_ready", "wait tm_empty", "wait full_b", "issue", "unit tail", "wait peer_full", "TOTAL", "-"]}
for r in range(2):
    for w in range(20):
        role = "epi" if w < 16 else ("fin" if w < 18 else ("prod" if w == 18 else "mma"))
        v = a[r, w]
        nb = max(v[7], 1.0) if (role == "epi" and not FINE) else 3400.0
        print("rank %d warp %2d %-4s" % (r, w, role), " ".join("%s=%.0f" % (nm, x / nb) for nm, x in zip(names[role], v) if nm != "-"))
