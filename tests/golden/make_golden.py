#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF, run in this container.

  * MFCC: the reference's own src/feature/MFCC.py + utils.py are read from /root/reference and
    executed in memory.  They are Python 2; the only edits applied (to the in-memory text, never
    written anywhere) are the mechanical ones SURVEY.md s7 item 3 lists:
        `/` -> `//` at the two integer divisions, xrange -> range, cast['float'] -> asarray(float),
        `from utils import` resolved by pre-seeding sys.modules, builtin min kept.
  * GMM: the reference's own C++ compiled by oracle/Makefile into oracle/_ref (unmodified sources).

Inputs are seeded (speaker_recognition_b200.synth) and stored next to the outputs so the fixtures
are self-contained: /root/reference does not exist on the GPU box.
Run:  python tests/golden/make_golden.py      (needs /root/reference and `make -C oracle`)
"""
import builtins
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from speaker_recognition_b200 import synth  # noqa: E402
from oracle import ref_so  # noqa: E402


def load_reference_mfcc():
    """exec the reference's feature/utils.py and feature/MFCC.py (py3-patched in memory)."""
    with open(os.path.join(REF, "src/feature/utils.py")) as f:
        utils_src = f.read()
    utils_mod = types.ModuleType("utils")
    exec(compile(utils_src, "ref:feature/utils.py", "exec"), utils_mod.__dict__)
    sys.modules["utils"] = utils_mod
    with open(os.path.join(REF, "src/feature/MFCC.py")) as f:
        src = f.read()
    patched = (src
               .replace("(len(signal) - self.FRAME_LEN) / self.FRAME_SHIFT + 1",
                        "(len(signal) - self.FRAME_LEN) // self.FRAME_SHIFT + 1")
               .replace("[:self.FFT_SIZE / 2 + 1]", "[:self.FFT_SIZE // 2 + 1]")
               .replace("xrange", "range")
               .replace("signal = cast['float'](signal)", "signal = asarray(signal, dtype=float)"))
    assert patched.count("//") == 3 and "xrange" not in patched   # 2 divisions + the URL in a docstring
    mod = types.ModuleType("MFCC_ref")
    exec(compile(patched, "ref:feature/MFCC.py", "exec"), mod.__dict__)
    # `from numpy import *` shadows the builtin min under NumPy 2; the reference ran under NumPy 1
    # where it did not.  Restore the builtin the reference used at MFCC.py:94.
    mod.__dict__["min"] = builtins.min
    del sys.modules["utils"]
    return mod, utils_mod


def gen_mfcc():
    MFCC, utils = load_reference_mfcc()
    cases = {}
    # C1: 5 s @ 16 kHz defaults
    sig = synth.audio(80000, seed=0, fs=16000)
    cases["c1_16k"] = dict(fs=16000, signal=sig, kwargs={}, out=MFCC.extract(16000, sig))
    # diff=True on a shorter clip
    sig = synth.audio(24000, seed=1, fs=16000)
    cases["diff_16k"] = dict(fs=16000, signal=sig, kwargs={}, out=MFCC.extract((16000, sig), diff=True))
    # 39-dim: extract then diff_feature(nd=2)
    cases["d39_16k"] = dict(fs=16000, signal=sig, kwargs={}, out=utils.diff_feature(MFCC.extract(16000, sig), nd=2))
    # 8 kHz (the GUI's recording rate, gui.py:49) -> FRAME_LEN 256, SHIFT 128
    sig = synth.audio(16000, seed=2, fs=8000)
    cases["c_8k"] = dict(fs=8000, signal=sig, kwargs={}, out=MFCC.extract(8000, sig))
    # non-default kwargs (the sweep in src/test/test-feature.py): 19 ceps, 55 filters, 20/10 ms
    sig = synth.audio(20000, seed=3, fs=16000)
    kw = dict(n_ceps=19, n_filters=55, win_length_ms=20, win_shift_ms=10, FFT_SIZE=1024, pre_emphasis_coef=0.97)
    cases["kw_16k"] = dict(fs=16000, signal=sig, kwargs=kw, out=MFCC.extract(16000, sig, **kw))
    # raw (un-normalised) cepstra + tables for the default extractor, to localise failures
    ex = MFCC.get_mfcc_extractor(16000)
    tab = dict(window=ex.window, mel=ex.M, dct=ex.D)
    flat = {}
    for name, c in cases.items():
        flat[name + "/fs"] = np.int64(c["fs"])
        flat[name + "/signal"] = c["signal"]
        flat[name + "/out"] = c["out"]
        flat[name + "/kwargs"] = np.array(repr(c["kwargs"]))
    for k, v in tab.items():
        flat["tables16k/" + k] = v
    np.savez_compressed(os.path.join(HERE, "mfcc_ref.npz"), **flat)
    print("mfcc_ref.npz:", {k: v["out"].shape for k, v in cases.items()})


def gen_gmm():
    from oracle import gmm_ref
    flat = {}
    rng = np.random.default_rng(7)
    # (1) shipped real-parameter models (D=34): params + frames drawn from them + reference scores
    for tag, fname, n in [("ubm32", "ubm.mixture-32.utt-300.model", 500),
                          ("ubm64", "ubm.mixture-64.utt-300.model", 300),
                          ("ubm256", "ubm.mixture-256.nperson-300.immature.model", 300)]:
        path = os.path.join(REF, "src/gui/model", fname)
        g = ref_so.RefGMM.load(path)
        w, m, s = g.params()
        X = synth.draw(w, m, s, n, seed=11)
        X[: n // 10] *= 1.7                    # some poorly-explained frames
        flat[tag + "/w"], flat[tag + "/mean"], flat[tag + "/sigma"], flat[tag + "/X"] = w, m, s, X
        flat[tag + "/score_fastexp"] = g.score(X, 4)
        flat[tag + "/score_libmexp"] = g.score_exactexp(X)
        flat[tag + "/score_all"] = np.float64(g.score_all_abi(X, 4))
        g.free()
    # (2) synthetic config-2-shaped speakers (D=39, K=128): 4 speakers, frames from speaker 3
    W, M, S = synth.models(4, 128, 39, seed0=1000)
    X = synth.draw(W[3], M[3], S[3], 500, seed=12)
    flat["c2/w"], flat["c2/mean"], flat["c2/sigma"], flat["c2/X"] = W, M, S, X
    sc = []
    for i in range(4):
        g = ref_so.RefGMM(W[i], M[i], S[i])
        sc.append(g.score(X, 4))
        g.free()
    flat["c2/score_fastexp"] = np.stack(sc)
    # (3) underflow behaviour: frames scaled so that some totals underflow (safe_log clamp, gmm.cc:34-38)
    w, m, s = flat["ubm32/w"], flat["ubm32/mean"], flat["ubm32/sigma"]
    Xu = synth.draw(w, m, s, 400, seed=13) * np.linspace(1.0, 8.0, 400)[:, None]
    g = ref_so.RefGMM(w, m, s)
    flat["underflow/X"] = Xu
    flat["underflow/score_fastexp"] = g.score(Xu, 2)
    g.free()
    # (4) one EM iteration and one MAP iteration from an injected start (K=8, D=13 and K=32, D=34)
    w0, m0, s0 = synth.model(8, 13, seed=21, floor_frac=0.0)
    Xt = synth.draw(*synth.model(8, 13, seed=22, floor_frac=0.0), 1500, seed=23)
    g = ref_so.RefGMM(w0, m0, s0)
    g.em_iteration(Xt, 1e-3, 2)
    flat["em/w0"], flat["em/mean0"], flat["em/sigma0"], flat["em/X"] = w0, m0, s0, Xt
    flat["em/w1"], flat["em/mean1"], flat["em/sigma1"] = g.params()
    g.em_iteration(Xt, 1e-3, 2)
    flat["em/w2"], flat["em/mean2"], flat["em/sigma2"] = g.params()
    g.free()
    ubm = ref_so.RefGMM(w, m, s)
    spk = ref_so.RefGMM(w, m, s)
    Xa = synth.draw(w, m + 0.3 * rng.standard_normal(m.shape), s, 600, seed=24)
    spk.map_iteration(ubm, Xa, 2)
    flat["map/X"] = Xa
    flat["map/w1"], flat["map/mean1"], flat["map/sigma1"] = spk.params()
    ubm.free(), spk.free()
    np.savez_compressed(os.path.join(HERE, "gmm_ref.npz"), **flat)
    print("gmm_ref.npz keys:", len(flat))


def gen_realparams():
    """(a) the fourth shipped UBM; (b) a 100-speaker set whose parameter statistics are resampled from the four shipped UBMs
    (synth.realparam_models), scored by the compiled reference: utterance sums for every speaker, per-frame values for three."""
    flat = {}
    fname = "ubm.mixture-32.person-20.immature.model"
    g = ref_so.RefGMM.load(os.path.join(REF, "src/gui/model", fname))
    w, m, s = g.params()
    X = synth.draw(w, m, s, 400, seed=31)
    X[:40] *= 1.7
    flat["ubm32p/w"], flat["ubm32p/mean"], flat["ubm32p/sigma"], flat["ubm32p/X"] = w, m, s, X
    flat["ubm32p/score_fastexp"] = g.score(X, 4)
    flat["ubm32p/score_all"] = np.float64(g.score_all_abi(X, 4))
    g.free()
    z = np.load(os.path.join(HERE, "gmm_ref.npz"))
    pools = [(z[t + "/w"], z[t + "/mean"], z[t + "/sigma"]) for t in ("ubm32", "ubm64", "ubm256")] + [(w, m, s)]
    S, K, src = 100, 32, 41
    W, M, Sg = synth.realparam_models(S, K, pools, seed0=5000)
    X = synth.draw(W[src], M[src], Sg[src], 600, seed=32)
    sums, frames = np.empty(S), {}
    for i in range(S):
        g = ref_so.RefGMM(W[i], M[i], Sg[i])
        sums[i] = g.score_all_abi(X, 4)
        if i in (0, src, 99):
            frames[i] = g.score(X, 4)
        g.free()
    flat["realset/S"], flat["realset/K"], flat["realset/src"], flat["realset/seed0"] = np.int64(S), np.int64(K), np.int64(src), np.int64(5000)
    flat["realset/X"], flat["realset/sums_fastexp"] = X, sums
    for i, f in frames.items():
        flat["realset/frames_%d" % i] = f
    assert int(np.argmax(sums)) == src
    # (c) batched MAP enrolment: six speakers with different amounts of data, each adapted from the shipped 32-mixture UBM by TWO
    # iterations of the reference's own GMMUBMTrainerBaseline (ref_harness: ref_map_iteration)
    wu, mu, su = z["ubm32/w"], z["ubm32/mean"], z["ubm32/sigma"]
    rng = np.random.default_rng(77)
    lens = [150, 333, 900, 64, 512, 257]
    Xs = [synth.draw(wu, mu + 0.4 * su * rng.standard_normal(mu.shape), su, n, seed=200 + i) for i, n in enumerate(lens)]
    ubm = ref_so.RefGMM(wu, mu, su)
    means = []
    for X in Xs:
        spk = ref_so.RefGMM(wu, mu, su)
        spk.map_iteration(ubm, X, 2)
        spk.map_iteration(ubm, X, 2)
        w2, m2, s2 = spk.params()
        assert np.array_equal(w2, wu) and np.array_equal(s2, su)       # MAP moves the means only (gmmubm.cc:40-81)
        means.append(m2)
        spk.free()
    ubm.free()
    flat["mapset/X"] = np.concatenate(Xs)
    flat["mapset/frame_off"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat["mapset/mean2"] = np.stack(means)
    np.savez_compressed(os.path.join(HERE, "realparams_ref.npz"), **flat)
    print("realparams_ref.npz keys:", len(flat))


def load_reference_silence():
    """exec the reference's filters/silence.py (py3-patched in memory: the two integer `/` -> `//`, the debug print dropped)."""
    with open(os.path.join(REF, "src/filters/silence.py")) as f:
        src = f.read()
    patched = (src.replace("signal = signal - (typeinfo.max + 1) / 2", "signal = signal - (typeinfo.max + 1) // 2")
               .replace("retsig = retsig + typeinfo.max / 2", "retsig = retsig + typeinfo.max // 2")
               .replace("    print average_energy\n", ""))
    assert patched.count("//") == 2 and "print average_energy" not in patched
    mod = types.ModuleType("silence_ref_exec")
    mod.__dict__["__name__"] = "silence_ref_exec"      # not "__main__": the file's main() must not run
    exec(compile(patched, "ref:filters/silence.py", "exec"), mod.__dict__)
    return mod


def gated_audio(n, seed, fs=16000):
    """audio with silent and quiet stretches (same generator as tests/test_silence_gpu.py)"""
    s = synth.audio(n, seed=seed, fs=fs).astype(np.int64)
    rng = np.random.default_rng(seed)
    for _ in range(6):
        w = max(8, min(4000, n // 8))
        a = int(rng.integers(0, n - w))
        s[a:a + int(rng.integers(w // 8, w))] //= int(rng.choice([1000, 40, 7]))
    return s.astype(np.int16)


def gen_silence():
    """outputs of the reference's own remove_silence (src/filters/silence.py:11-50) on seeded gated audio, int16 and uint8"""
    mod = load_reference_silence()
    flat = {}
    cases = [("a", 16000, 80000, 1, {}), ("b", 8000, 30001, 2, {}), ("c", 22050, 50000, 3, {}), ("d", 16000, 333, 4, {}),
             ("e", 16000, 60000, 9, dict(perc=0.5)), ("f", 16000, 60000, 9, dict(frame_duration=0.03, frame_shift=0.007))]
    for tag, fs, n, seed, kw in cases:
        sig = gated_audio(n, seed, fs)
        flat[tag + "/fs"], flat[tag + "/signal"], flat[tag + "/kwargs"] = np.int64(fs), sig, np.array(repr(kw))
        flat[tag + "/out"] = mod.remove_silence(fs, sig, **kw)
    sig = gated_audio(60000, 9)
    u8 = ((sig.astype(np.int32) >> 8) + 128).astype(np.uint8)
    flat["u8/fs"], flat["u8/signal"], flat["u8/kwargs"] = np.int64(16000), u8, np.array(repr({}))
    flat["u8/out"] = mod.remove_silence(16000, u8)
    np.savez_compressed(os.path.join(HERE, "silence_ref.npz"), **flat)
    print("silence_ref.npz:", {k: v.shape for k, v in flat.items() if k.endswith("/out")})


def gen_lpc():
    """LPC: the reference's feature path (src/feature/LPC.py:46-57) with its talkbox call replaced by the reference repo's
    own prototype of the same algorithm, src/lpcc/lpc_jp.py:13-66 (autocorr + LevinsonDurbin), exec'd from the tree.
    (talkbox itself is third-party and absent; lpc_jp.py imports it and pylab at module level, so only the two function
    definitions are compiled out of the file's AST.)"""
    import ast
    MFCC, _ = load_reference_mfcc()
    with open(os.path.join(REF, "src/lpcc/lpc_jp.py")) as f:
        src = f.read()
    src = src[:src.index('if __name__ == "__main__":')]          # the script part below is Python 2 (print statements)
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("autocorr", "LevinsonDurbin")]
    assert len(keep) == 2
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "ref:lpcc/lpc_jp.py", "exec"), ns)
    flat = {}
    for name, fs, n, seed, kw in [("lpc_16k", 16000, 12000, 31, dict(win_length_ms=32, win_shift_ms=16, n_lpc=15, pre_emphasis_coef=0.95)),
                                  ("lpc_8k_o10", 8000, 5000, 32, dict(win_length_ms=32, win_shift_ms=16, n_lpc=10, pre_emphasis_coef=0.95))]:
        sig = synth.audio(n, seed=seed, fs=fs)
        flen = int(float(kw["win_length_ms"]) / 1000 * fs)
        fshift = int(float(kw["win_shift_ms"]) / 1000 * fs)
        window = MFCC.hamming(flen)
        x = np.asarray(sig, dtype=float)
        frames = (len(x) - flen) // fshift + 1                      # LPC.py:47
        out = []
        for f in range(frames):
            frame = x[f * fshift: f * fshift + flen] * window       # LPC.py:50-51
            frame[1:] -= frame[:-1] * kw["pre_emphasis_coef"]       # LPC.py:53
            r = ns["autocorr"](frame, kw["n_lpc"] + 1)
            a, _e = ns["LevinsonDurbin"](r, kw["n_lpc"])
            out.append(a[1:])                                       # LPC.py:41-42  lpc[1:]
        flat[name + "/fs"] = np.int64(fs)
        flat[name + "/signal"] = sig
        flat[name + "/out"] = np.array(out)
        flat[name + "/kwargs"] = np.array(repr(kw))
    np.savez_compressed(os.path.join(HERE, "lpc_ref.npz"), **flat)
    print("lpc_ref.npz:", {k: v.shape for k, v in flat.items() if k.endswith("/out")})


def gen_kmeans():
    """K-MeansII initialisation (gmm.cc:331-344): the compiled reference's Lloyd iterations from injected centroids
    (KMeansSolver::Lloyds_iteration[_weighted], kmeans.cc:150-340 -- deterministic) and the final cost of its whole
    KMeansIISolver::cluster (kmeansII.cc:80-172) for several srand() seeds (randomised: a statistical pin)."""
    rng = np.random.default_rng(2024)
    flat = {}
    for name, n_per, D, n_clusters, K in [("a", 400, 6, 10, 8), ("b", 150, 13, 24, 32)]:
        centers = rng.normal(0, 4, (n_clusters, D))
        X = np.concatenate([c + rng.normal(0, 1, (n_per, D)) for c in centers])
        rng.shuffle(X)
        X = X.astype(np.float32).astype(np.float64)                 # the product stages fp32 frames: keep the data exactly representable
        init = X[:K].copy()
        w = rng.integers(0, 5, len(X)).astype(np.float64)
        c_u, cost_u = ref_so.ref_lloyd(X, init)
        c_w, cost_w = ref_so.ref_lloyd(X, init, w)
        costs = [ref_so.ref_kmeansII(X, K, seed)[1] for seed in range(100, 130)]
        flat.update({name + "/X": X, name + "/K": np.int64(K), name + "/init": init, name + "/weight": w,
                     name + "/lloyd_cent": c_u, name + "/lloyd_cost": np.float64(cost_u),
                     name + "/wlloyd_cent": c_w, name + "/wlloyd_cost": np.float64(cost_w),
                     name + "/kmeansII_costs": np.array(costs)})
    np.savez_compressed(os.path.join(HERE, "kmeans_ref.npz"), **flat)
    print("kmeans_ref.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in flat.items() if "cost" in k})


if __name__ == "__main__":
    which = sys.argv[1:] or ["mfcc", "gmm", "lpc", "kmeans", "realparams", "silence"]
    if "mfcc" in which:
        gen_mfcc()
    if "gmm" in which:
        gen_gmm()
    if "lpc" in which:
        gen_lpc()
    if "kmeans" in which:
        gen_kmeans()
    if "realparams" in which:
        gen_realparams()
    if "silence" in which:
        gen_silence()
