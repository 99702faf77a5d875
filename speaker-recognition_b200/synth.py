"""Seeded synthetic inputs for tests and benchmarks (no dataset ships with the reference, and there is
no network): 16 kHz int16 "voiced" audio and diagonal-GMM speaker models, as specified in
SURVEY.md s8(d) / BASELINE.md s3.  Pure NumPy, deterministic (numpy PCG64 streams).
"""
import numpy as np

SIGMA_FLOOR = 0.0316228  # sqrt(min_covar = 1e-3) as printed in the reference's shipped models (gmm.hh:58)


def audio(n_samples, seed=0, fs=16000, dtype=np.int16):
    """3-5 harmonics of f0 in [90, 250] Hz with slow AM + white noise at -20 dB, peak ~ 0.3*32767.
    Never digitally silent, so the 1e-100 power floor (MFCC.py:7) is never active."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / fs
    f0 = rng.uniform(90.0, 250.0)
    nh = int(rng.integers(3, 6))
    sig = np.zeros(n_samples)
    for h in range(1, nh + 1):
        amp = rng.uniform(0.3, 1.0) / h
        ph = rng.uniform(0, 2 * np.pi)
        sig += amp * np.sin(2 * np.pi * f0 * h * t + ph)
    am = 1.0 + 0.5 * np.sin(2 * np.pi * rng.uniform(1.0, 4.0) * t + rng.uniform(0, 2 * np.pi))
    sig *= am
    sig /= np.max(np.abs(sig))
    noise = rng.standard_normal(n_samples)
    sig = sig + noise * (np.sqrt(np.mean(sig ** 2)) * 0.1)  # -20 dB
    sig *= 0.3 * 32767.0 / np.max(np.abs(sig))
    if dtype == np.int16:
        return np.round(sig).astype(np.int16)
    return sig.astype(dtype)


def samples_for_frames(n_frames, frame_len=512, frame_shift=256):
    return (n_frames - 1) * frame_shift + frame_len


def model(K, D, seed, floor_frac=0.01):
    """means ~ N(0,1), sigma ~ U[0.3,1.2] with floor_frac of entries at the sigma floor, weights ~ Dirichlet(1)."""
    rng = np.random.default_rng(seed)
    mean = rng.standard_normal((K, D))
    sigma = rng.uniform(0.3, 1.2, size=(K, D))
    sigma[rng.random((K, D)) < floor_frac] = SIGMA_FLOOR
    w = rng.dirichlet(np.ones(K))
    return w, mean, sigma


def models(S, K, D, seed0=1000, floor_frac=0.01):
    ws, ms, ss = [], [], []
    for s in range(S):
        w, m, sg = model(K, D, seed0 + s, floor_frac)
        ws.append(w), ms.append(m), ss.append(sg)
    return np.stack(ws), np.stack(ms), np.stack(ss)


def draw(w, mean, sigma, n, seed):
    """n frames drawn from the diagonal GMM (w, mean, sigma)."""
    rng = np.random.default_rng(seed)
    k = rng.choice(len(w), size=n, p=w / w.sum())
    return mean[k] + sigma[k] * rng.standard_normal((n, mean.shape[1]))


def realparam_models(S, K, pools, seed0=5000, shift=0.25):
    """Speakers whose parameter statistics are those of REAL trained models: every mixture of every speaker is a whole
    (mean, sigma) row resampled from `pools` -- a list of (w, mean, sigma) of trained models sharing D, e.g. the UBMs the
    reference ships (src/gui/model/*.model, D = 34 = MFCC 19 + LPC 15: un-normalised scales, sigma floor 0.0316 present,
    collapsed mixtures included) -- with the mean moved by shift * sigma * N(0,1) (a MAP-adaptation-sized displacement) and
    Dirichlet weights.  Returns (W [S,K], M [S,K,D], Sg [S,K,D])."""
    pm = np.concatenate([np.asarray(p[1], dtype=np.float64) for p in pools])
    ps = np.concatenate([np.asarray(p[2], dtype=np.float64) for p in pools])
    D = pm.shape[1]
    W, M, Sg = np.empty((S, K)), np.empty((S, K, D)), np.empty((S, K, D))
    for s in range(S):
        rng = np.random.default_rng(seed0 + s)
        rows = rng.integers(0, len(pm), size=K)
        Sg[s] = ps[rows]
        M[s] = pm[rows] + shift * ps[rows] * rng.standard_normal((K, D))
        W[s] = rng.dirichlet(np.ones(K))
    return W, M, Sg
