import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.')
from speaker_recognition_b200 import _lib, synth
L = _lib.lib()
K, D, n = 512, 39, 1000000
W, M, Sg = synth.models(1, K, D, seed0=7)
X = synth.draw(W[0], M[0], Sg[0], n, seed=1)
X = _lib.as_f64(X)
c = np.zeros((K, D)); cost = C.c_double()
for rep in range(2):
    t = time.time()
    _lib.check(L.srb200_kmeans_init_f64(_lib.dptr(X), n, D, K, C.c_uint64(5 + rep), _lib.dptr(c), C.byref(cost)))
    print("kmeansII init n=%d K=%d D=%d: %.2f s, cost %.6g" % (n, K, D, time.time() - t, cost.value), flush=True)
