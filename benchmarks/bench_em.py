"""E-step timing at BASELINE config 3 (K = 512, D = 39, N = 1e6): tensor-core route vs FMA kernels, per iteration of a short training run.
usage: python benchmarks/bench_em.py [iters]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from speaker_recognition_b200 import _lib, synth  # noqa: E402
from speaker_recognition_b200.gmm.pygmm import GMM  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    N, K, D = 1000000, 512, 39
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(1234)
    wt, mt, st_ = synth.model(K, D, seed=77, floor_frac=0.0)
    comp = torch.multinomial(torch.from_numpy(wt).cuda(), N, replacement=True, generator=g)
    X = torch.from_numpy(mt).cuda()[comp] + torch.from_numpy(st_).cuda()[comp] * torch.randn((N, D), device="cuda", dtype=torch.float64, generator=g)
    ld = (D + 3) & ~3
    Xd = torch.zeros((N, ld), dtype=torch.float32, device="cuda")
    Xd[:, :D] = X.float()
    idx = torch.randint(0, N, (K,), generator=torch.Generator().manual_seed(5))
    m0 = X[idx.cuda()].cpu().numpy()
    s0 = np.tile(X.std(dim=0, unbiased=True).cpu().numpy(), (K, 1))
    w0 = np.ones(K) / K
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for path in (1, 0):
        _lib.check(L.srb200_set_em_path(path))
        gm = GMM(K).set_params(w0, m0, s0)
        stats = torch.zeros(L.srb200_em_stats_len(K, D), dtype=torch.float64, device="cuda")
        rows = []
        for it in range(iters):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            _lib.check(L.srb200_em_estep_dev(gm.gmm, Xd.data_ptr(), N, ld, D, stats.data_ptr(), st))
            e1.record()
            _lib.check(L.srb200_em_mstep_dev(gm.gmm, stats.data_ptr(), N, 1e-3, 0, None, st))
            e2.record()
            torch.cuda.synchronize()
            rows.append({"it": it, "route": int(L.srb200_last_em_route()), "estep_ms": e0.elapsed_time(e1), "mstep_ms": e1.elapsed_time(e2),
                         "ll": float(stats[-2].item())})
        out["path%d" % path] = rows
    _lib.check(L.srb200_set_em_path(-1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
