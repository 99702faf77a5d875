"""CPU emulation of the fp16x3 tensor route (round 2): features centred per dimension (c_d) and scaled by a power of two (alpha_d),
PHI = [x''^2, x'', 1] and W split into fp16 hi + fp16 lo by round-to-nearest, three passes (hi*hi + lo*hi + hi*lo) accumulated in
fp32; entries whose |nb''| makes the expanded form cancel beyond a budget are taken out and added back exactly.
Reports per-frame relative error vs float64 and the relative error of utterance sums, as a function of the budget."""
import sys, os, glob
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speaker_recognition_b200 import synth
from oracle import gmm_ref
A0 = 0.84932180028801904272

def split16(v):
    v = v.astype(np.float32)
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)

def split_tf32_trunc(v):
    v = v.astype(np.float32)
    hi = (v.view(np.int32) & ~0x1FFF).view(np.float32)
    lo = ((v - hi).view(np.int32) & ~0x1FFF).view(np.float32)
    return hi, lo

def lse(l, S, K):
    l = l.reshape(l.shape[0], S, K); m = l.max(-1, keepdims=True)
    return (m[..., 0] + np.log2(np.exp2(l - m).sum(-1))) * np.log(2.0)

def run(W_, M_, S_, X, budget, split=split16, centre=True, verbose=False):
    S, K, D = M_.shape
    w = W_.reshape(S * K); mu = M_.reshape(S * K, D); sg = S_.reshape(S * K, D)
    # centre + scale from the model set itself
    gm = (w[:, None] * mu).sum(0) / S
    gstd = np.sqrt((w[:, None] * (sg ** 2 + (mu - gm) ** 2)).sum(0) / S)
    c = gm if centre else np.zeros(D)
    alpha = np.exp2(-np.round(np.log2(gstd))) if centre else np.ones(D)
    x = (X - c) * alpha                      # x''
    a = A0 / (sg * alpha); nb = -(mu - c) * alpha * a
    c2 = np.log2(w) - np.log2(2.5066282746310002 * sg).sum(1)
    t_exact = (x[:, None, :] * a[None] + nb[None])
    l_exact = c2[None] - (t_exact ** 2).sum(-1)
    LL64 = lse(l_exact, S, K)
    # hard entries: per row, take out the largest nb^2 until the rest is within the budget
    nb2 = nb ** 2
    hard = np.zeros_like(nb2, dtype=bool)
    order = np.argsort(-nb2, axis=1)
    for r in range(S * K):
        tot = nb2[r].sum()
        for d in order[r]:
            if tot <= budget: break
            hard[r, d] = True; tot -= nb2[r, d]
    ah = np.where(hard, 0.0, a); nbh = np.where(hard, 0.0, nb)
    C = c2 - (nbh ** 2).sum(1)
    Wm = np.concatenate([-(ah ** 2), -2 * ah * nbh, C[:, None]], 1)
    x32 = x.astype(np.float32)
    phi = np.concatenate([x32 * x32, x32, np.ones((len(X), 1), np.float32)], 1)
    Wh, Wl = split(Wm); Ph, Pl = split(phi)
    l = (np.concatenate([Ph, Pl, Ph], 1) @ np.concatenate([Wh, Wh, Wl], 1).T).astype(np.float64)     # one fp32 accumulator
    # exact add-back in fp32
    a32 = a.astype(np.float32); nb32 = nb.astype(np.float32)
    rr, dd = np.nonzero(hard)
    for r, d in zip(rr, dd):
        t = x32[:, d] * a32[r, d] + nb32[r, d]
        l[:, r] -= (t * t).astype(np.float64)
    LL = lse(l, S, K)
    rel = np.abs(LL - LL64) / np.maximum(1, np.abs(LL64))
    sums = np.abs(LL.sum(0) - LL64.sum(0)) / np.abs(LL64.sum(0))
    return dict(hard_frac=float(hard.mean()), hard_per_block=float(hard.reshape(-1, min(128, S * K), D).sum((1, 2)).max()) if (S * K) % min(128, S * K) == 0 else -1,
                frame_rel_max=float(rel.max()), sum_rel_max=float(sums.max()), bias=float(((LL - LL64).mean())), argmax_ok=bool((LL.sum(0).argmax() == LL64.sum(0).argmax())))

if __name__ == "__main__":
    T = 1500
    print("== synthetic C2-like (10 spk x 128 mix, D=39, 1% floor), frames from speaker 5")
    W_, M_, S_ = synth.models(10, 128, 39, seed0=1000)
    X = synth.draw(W_[5], M_[5], S_[5], T, seed=7)
    for b in (50, 200, 800, 2000, 8000, 1e9):
        print(" f16 RN budget %8.0f" % b, run(W_, M_, S_, X, b))
    print(" tf32 trunc budget 1e9 ", run(W_, M_, S_, X, 1e9, split=split_tf32_trunc))
    print(" tf32 trunc budget 72  ", run(W_, M_, S_, X, 72, split=split_tf32_trunc))
    print("== shipped UBMs (D=34), frames drawn from the model")
    for p in sorted(glob.glob('/root/reference/src/gui/model/*.model')):
        w, m, s = gmm_ref.load_model(p)
        X = synth.draw(w, m, s, T, seed=3)
        for b in (200, 800, 2000, 1e9):
            print(" %-45s budget %6.0f" % (os.path.basename(p)[:45], b), run(w[None], m[None], s[None], X, b))
        print(" %-45s uncentred b=800 " % "", run(w[None], m[None], s[None], X, 800, centre=False))
