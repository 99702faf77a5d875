"""Numerics probe for a tensor-core scorer: how accurate is the expanded-square GEMM  l = C + W.[x^2, x]  when run as
3xTF32 (hi*hi + hi*lo + lo*hi) or 6-pass 3xBF16 on the B200 tensor cores (cuBLAS through torch), against float64?
The metric is the per-frame log-likelihood error |LL - LL64| / max(1, |LL64|) (north_star tolerance 1e-4)."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speaker_recognition_b200 import synth

torch.backends.cuda.matmul.allow_tf32 = True
S, K, D, T = 20, 128, 39, 4096
W_, M_, S_ = synth.models(S, K, D, seed0=1000)
X = synth.draw(W_[3], M_[3], S_[3], T, seed=5)
dev = "cuda"
x = torch.from_numpy(X).to(dev)                      # float64
w = torch.from_numpy(W_).to(dev).reshape(S * K)
mu = torch.from_numpy(M_).to(dev).reshape(S * K, D)
sg = torch.from_numpy(S_).to(dev).reshape(S * K, D)
A0 = 0.84932180028801904272
a = A0 / sg
nb = -mu * a
c2 = torch.log2(w) - torch.log2(2.5066282746310002 * sg).sum(1)
# exact float64
l_exact = c2[None, :] - ((x[:, None, :] * a[None] + nb[None]) ** 2).sum(-1)            # [T, SK]
def lse(l):
    l = l.reshape(T, S, K)
    m = l.max(-1, keepdim=True).values
    return ((m[..., 0] + torch.log2(torch.exp2(l - m).sum(-1))) * np.log(2.0))
LL64 = lse(l_exact)

def split_tf32(v32):
    hi = (v32.view(torch.int32) & ~0x1FFF).view(torch.float32)
    return hi, (v32 - hi)

def split_bf16(v32, n=3):
    parts, r = [], v32.clone()
    for _ in range(n):
        p = r.to(torch.bfloat16)
        parts.append(p)
        r = r - p.float()
    return parts

def run(hard_thresh=None, mode="tf32", order="blocked"):
    am, nbm = a.clone(), nb.clone()
    hard = torch.zeros_like(a, dtype=torch.bool)
    if hard_thresh is not None:
        hard = sg < hard_thresh
        am = torch.where(hard, torch.zeros_like(a), a)
        nbm = torch.where(hard, torch.zeros_like(nb), nb)
    C = (c2 - (nbm ** 2).sum(1))
    Wm = torch.cat([-(am ** 2), -2 * am * nbm, C[:, None]], dim=1)       # [SK, 2D+1]
    phi = torch.cat([x ** 2, x, torch.ones(T, 1, dtype=x.dtype, device=dev)], dim=1)    # [T, 2D+1]
    if order == "interleaved":
        idx = torch.stack([torch.arange(D), torch.arange(D) + D], 1).reshape(-1).tolist() + [2 * D]
        Wm, phi = Wm[:, idx], phi[:, idx]
    W32, P32 = Wm.float(), phi.float()
    if mode == "tf32":
        Wh, Wl = split_tf32(W32); Ph, Pl = split_tf32(P32)
        l = Ph @ Wh.T + (Ph @ Wl.T + Pl @ Wh.T)
    elif mode == "fp32":
        torch.backends.cuda.matmul.allow_tf32 = False
        l = P32 @ W32.T
        torch.backends.cuda.matmul.allow_tf32 = True
    else:
        Wp, Pp = split_bf16(W32), split_bf16(P32)
        l = torch.zeros(T, S * K, device=dev)
        for i, j in [(2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0)]:
            l += (Pp[i] @ Wp[j].T).float()
    if hard_thresh is not None:   # exact fp32 correction for the excluded dims
        t = (x.float()[:, None, :] * a.float()[None] + nb.float()[None]) ** 2
        l = l - (t * hard[None].float()).sum(-1)
    l = l.double()
    LL = lse(l)
    err = ((LL - LL64).abs() / LL64.abs().clamp(min=1)).max().item()
    rel = (l_exact.reshape(T, S, K) - l_exact.reshape(T, S, K).max(-1, keepdim=True).values) > -40
    el = ((l - l_exact).abs().reshape(T, S, K))[rel].max().item()
    return {"LL_relerr_max": err, "l2_abs_err_relevant_max": el, "argmax_ok": bool((LL.argmax(1) == LL64.argmax(1)).all())}

out = {}
for name, kw in [("fp32_gemm", dict(mode="fp32")), ("3xtf32", dict(mode="tf32")), ("3xtf32_interleaved", dict(mode="tf32", order="interleaved")),
                 ("3xtf32_hard0.1", dict(mode="tf32", hard_thresh=0.1)), ("3xtf32_hard0.1_interleaved", dict(mode="tf32", hard_thresh=0.1, order="interleaved")),
                 ("6xbf16", dict(mode="bf16")), ("6xbf16_hard0.1", dict(mode="bf16", hard_thresh=0.1))]:
    out[name] = run(**kw)
    print(name, out[name])
print(json.dumps(out))
