"""Follow-up probe: 3-pass vs 4-pass (adds lo*lo) TF32 and plain fp32 GEMM on the datasets of tests/test_gmm_tc_gpu.py,
mismatched speakers included (frames drawn from one speaker, scored on all)."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speaker_recognition_b200 import synth
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"
A0 = 0.84932180028801904272

def split_tf32(v32):
    hi = (v32.view(torch.int32) & ~0x1FFF).view(torch.float32)
    return hi, (v32 - hi)

def run(S, K, D, T, seed0, floor_frac=0.01, src=None):
    W_, M_, S_ = synth.models(S, K, D, seed0=seed0, floor_frac=floor_frac)
    src = S // 2 if src is None else src
    X = synth.draw(W_[src], M_[src], S_[src], T, seed=K + T)
    x = torch.from_numpy(X).to(dev)
    w = torch.from_numpy(W_).to(dev).reshape(S * K); mu = torch.from_numpy(M_).to(dev).reshape(S * K, D); sg = torch.from_numpy(S_).to(dev).reshape(S * K, D)
    a = A0 / sg; nb = -mu * a
    c2 = torch.log2(w) - torch.log2(2.5066282746310002 * sg).sum(1)
    l_exact = c2[None] - ((x[:, None, :] * a[None] + nb[None]) ** 2).sum(-1)
    def lse(l):
        l = l.reshape(T, S, K); m = l.max(-1, keepdim=True).values
        return (m[..., 0] + torch.log2(torch.exp2(l - m).sum(-1))) * np.log(2.0)
    LL64 = lse(l_exact)
    C = c2 - (nb ** 2).sum(1)
    Wm = torch.cat([-(a ** 2), -2 * a * nb, C[:, None]], 1).float()
    phi = torch.cat([x ** 2, x, torch.ones(T, 1, dtype=x.dtype, device=dev)], 1).float()
    Wh, Wl = split_tf32(Wm); Ph, Pl = split_tf32(phi)
    res = {}
    l3 = Ph @ Wh.T + (Ph @ Wl.T + Pl @ Wh.T)
    l4 = Ph @ Wh.T + (Ph @ Wl.T + Pl @ Wh.T) + Pl @ Wl.T
    # accumulate all passes into one fp32 accumulator in one GEMM (what the kernel does): concatenate along K
    l3c = torch.cat([Ph, Ph, Pl], 1) @ torch.cat([Wh, Wl, Wh], 1).T
    l4c = torch.cat([Ph, Ph, Pl, Pl], 1) @ torch.cat([Wh, Wl, Wh, Wl], 1).T
    torch.backends.cuda.matmul.allow_tf32 = False
    l32 = phi @ Wm.T
    torch.backends.cuda.matmul.allow_tf32 = True
    for name, l in (("3pass", l3), ("4pass", l4), ("3pass_oneacc", l3c), ("4pass_oneacc", l4c), ("fp32", l32)):
        LL = lse(l.double())
        res[name] = float(((LL - LL64).abs() / LL64.abs().clamp(min=1)).max())
    return res

cases = [(3, 128, 39, 1000, 3000 + 3 * 7 + 128), (5, 32, 39, 300, 3000 + 5 * 7 + 32), (3, 256, 39, 500, 3000 + 3 * 7 + 256), (7, 100, 34, 640, 3000 + 49 + 100),
         (20, 128, 39, 4096, 1000), (10, 128, 39, 2000, 7777)]
for c in cases:
    print(c, run(*c))
print("floor 5%:", run(10, 128, 39, 2000, 4242, floor_frac=0.05))
print("floor 20%:", run(10, 128, 39, 2000, 4243, floor_frac=0.2))
