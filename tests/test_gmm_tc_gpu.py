"""GPU parity tests of the tensor-core scoring route (csrc/gmm_score_tc.cuh: tcgen05, TMEM accumulators)
against the float64 oracle and against the default fp32-FMA route.  Bound: the north_star's 1e-4 * max(1, |ref|) per frame,
identical argmax; this route's measured worst case on sigma-floor models is ~4e-5 (the FMA route: ~2e-7)."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import gmm_ref
from speaker_recognition_b200 import _lib, synth
from speaker_recognition_b200.gmm.pygmm import GMM
from speaker_recognition_b200.gmmset import GMMSet

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture()
def tc_route():
    L = _lib.lib()
    _lib.check(L.srb200_set_score_path(1))
    yield L
    _lib.check(L.srb200_set_score_path(-1))


def build_set(S, K, D, seed0=1000, floor_frac=0.01):
    W, M, Sg = synth.models(S, K, D, seed0=seed0, floor_frac=floor_frac)
    gs = GMMSet(gmm_order=K)
    for i in range(S):
        gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
        gs.y.append(i)
    return gs, W, M, Sg


@pytest.mark.parametrize("S,K,D,T", [(3, 128, 39, 1000), (5, 32, 39, 300), (4, 10, 13, 129), (2, 64, 26, 128), (3, 256, 39, 500),
                                     (1, 512, 39, 257), (7, 100, 34, 640), (2, 40, 50, 200)])
def test_tc_route_matches_oracle_per_frame_and_sums(tc_route, S, K, D, T):
    gs, W, M, Sg = build_set(S, K, D, seed0=3000 + S * 7 + K)
    X = synth.draw(W[S // 2], M[S // 2], Sg[S // 2], T, seed=K + T)
    sums = np.array(gs.predict_one_scores(X))
    for i in range(S):
        exact = gmm_ref.score(W[i], M[i], Sg[i], X, gmm_ref.MODE_LOGDOMAIN)
        got = gs.gmms[i].score(X)                        # per-frame values through the single-model entry point
        assert np.max(rel_err(got, exact)) < TOL
        assert abs(sums[i] - exact.sum()) < 2e-5 * abs(exact.sum())
    assert int(np.argmax(sums)) == S // 2


def test_tc_route_real_parameter_models_and_clamp(tc_route, golden_gmm, golden_real):
    """the four UBMs the reference ships (D = 34 = MFCC + LPC: un-normalised scales, 12-17 % of the sigmas below 0.1) run ON the
    tensor route -- asserted, not assumed -- and match the float64 oracle per frame and the compiled reference's sums"""
    L = tc_route
    for fx, tag in ((golden_gmm, "ubm32"), (golden_gmm, "ubm64"), (golden_gmm, "ubm256"), (golden_real, "ubm32p")):
        c = fx.case(tag)
        g = GMM(len(c["w"])).set_params(c["w"], c["mean"], c["sigma"])
        got = g.score(c["X"])
        assert L.srb200_last_score_route() == 1, tag
        exact = gmm_ref.score(c["w"], c["mean"], c["sigma"], c["X"], gmm_ref.MODE_LOGDOMAIN)
        assert np.max(rel_err(got, exact)) < TOL
        assert np.max(rel_err(got, c["score_fastexp"])) < 5e-4         # the reference's own per-frame values (fast-exp error included)
        assert abs(g.score_all(c["X"]) - float(c["score_all"])) < 1e-5 * abs(float(c["score_all"]))
        assert L.srb200_last_score_route() == 1, tag
    c, u = golden_gmm.case("ubm32"), golden_gmm.case("underflow")
    g = GMM(32).set_params(c["w"], c["mean"], c["sigma"])
    got = g.score(u["X"])
    assert L.srb200_last_score_route() == 1
    exact = gmm_ref.score(c["w"], c["mean"], c["sigma"], u["X"], gmm_ref.MODE_LOGDOMAIN)
    deep = exact < gmm_ref.DENORM_LOG_MIN - 5.0
    assert deep.sum() > 3 and np.allclose(got[deep], gmm_ref.SAFE_LOG_FLOOR, atol=1e-5)      # safe_log clamp (gmm.cc:34-38)
    dom = exact > -600
    assert np.max(rel_err(got[dom], exact[dom])) < TOL


def test_tc_route_hundred_speakers_with_real_parameter_statistics(tc_route, golden_gmm, golden_real):
    """100 speakers x 32 mixtures, D = 34, every mixture a (mean, sigma) row resampled from the shipped UBMs
    (synth.realparam_models): one launch on the tensor route; sums and argmax against the compiled reference
    (tests/golden/realparams_ref.npz), per-frame values against float64"""
    from conftest import shipped_ubm_pools
    L = tc_route
    c = golden_real.case("realset")
    S, K, src = int(c["S"]), int(c["K"]), int(c["src"])
    W, M, Sg = synth.realparam_models(S, K, shipped_ubm_pools(golden_gmm, golden_real), seed0=int(c["seed0"]))
    gs = GMMSet(gmm_order=K)
    for i in range(S):
        gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
        gs.y.append(i)
    sums = np.array(gs.predict_one_scores(c["X"]))
    assert L.srb200_last_score_route() == 1
    assert np.max(np.abs(sums - c["sums_fastexp"]) / np.abs(c["sums_fastexp"])) < 1e-5
    assert int(np.argmax(sums)) == int(np.argmax(c["sums_fastexp"])) == src
    for i in (0, src, 99):
        got = gs.gmms[i].score(c["X"])
        assert L.srb200_last_score_route() == 1
        exact = gmm_ref.score(W[i], M[i], Sg[i], c["X"], gmm_ref.MODE_LOGDOMAIN)
        assert np.max(rel_err(got, exact)) < TOL
        assert np.max(rel_err(got, c["frames_%d" % i])) < 5e-4


def test_tc_route_sees_retrained_parameters(tc_route):
    """a handle that was scored (tensor-core images built) and is then trained -- by fit(), or by the device EM building blocks --
    must score with its NEW parameters afterwards (round-1 advisor finding: the images of the old parameters survived)"""
    import torch
    from speaker_recognition_b200 import dist
    L = tc_route
    K, D = 16, 13
    w0, m0, s0 = synth.model(K, D, seed=71, floor_frac=0.0)
    X = synth.draw(*synth.model(K, D, seed=72, floor_frac=0.0), 3000, seed=73)
    # (a) set_params -> score -> fit -> score
    g = GMM(K, nr_iteration=6, threshold=0.0, seed=3).set_params(w0, m0, s0)
    before = g.score_all(X)
    assert L.srb200_last_score_route() == 1
    g.fit(X)
    assert g.last_iterations == 6                                   # the monitor did not stop early on a frozen likelihood
    w1, m1, s1 = g.get_params()
    after = g.score_all(X)
    assert L.srb200_last_score_route() == 1
    exact = gmm_ref.score_all(w1, m1, s1, X, gmm_ref.MODE_LOGDOMAIN)
    assert abs(after - exact) < 2e-5 * abs(exact) and after > before + 10.0
    # (b) device EM building blocks: set_params -> E-step (packs) -> M-step -> score
    g2 = GMM(K).set_params(w0, m0, s0)
    g2.score_all(X[:200])
    Xd = torch.zeros((len(X), 16), dtype=torch.float32, device="cuda")
    Xd[:, :D] = torch.from_numpy(X.astype(np.float32)).cuda()
    em = dist.DeviceEM(g2, Xd, len(X), 16, D)
    for _ in range(2):
        em.mstep(em.estep(), len(X), 1e-3, 0)
    # the use the advisor described: score right after the device M-step (the host arrays are still the old ones)
    mid = g2.score_all(X)
    em.finish()
    w2, m2, s2 = g2.get_params()
    got = g2.score_all(X)
    assert L.srb200_last_score_route() == 1
    exact2 = gmm_ref.score_all(w2, m2, s2, X, gmm_ref.MODE_LOGDOMAIN)
    assert abs(got - exact2) < 2e-5 * abs(exact2) and abs(mid - exact2) < 2e-5 * abs(exact2)
    assert abs(exact2 - before) > 1e-3 * abs(before)


def test_tc_route_equals_fma_route_on_batched_utterances(tc_route):
    """device entry point with several ragged utterances: both routes give the same sums (to the TC route's accuracy) and argmax"""
    import ctypes as C
    import torch
    L = tc_route
    S, K, D = 9, 128, 39
    gs, W, M, Sg = build_set(S, K, D)
    lens = [700, 129, 1, 2048, 333]
    X = np.concatenate([synth.draw(W[i % S], M[i % S], Sg[i % S], n, seed=50 + i) for i, n in enumerate(lens)])
    Xd = torch.zeros((len(X), 40), dtype=torch.float32, device="cuda")
    Xd[:, :D] = torch.from_numpy(X.astype(np.float32)).cuda()
    off = (C.c_int * (len(lens) + 1))(*np.concatenate([[0], np.cumsum(lens)]).tolist())
    res = []
    for path in (1, 0):
        _lib.check(L.srb200_set_score_path(path))
        sums = torch.zeros((len(lens), S), dtype=torch.float64, device="cuda")
        best = torch.zeros(len(lens), dtype=torch.int32, device="cuda")
        _lib.check(L.srb200_gmmset_score_dev(gs._ensure_set(), Xd.data_ptr(), 40, D, off, len(lens), sums.data_ptr(), None, 0,
                                             best.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        res.append((sums.cpu().numpy(), best.cpu().numpy()))
    _lib.check(L.srb200_set_score_path(1))
    assert np.max(np.abs(res[0][0] - res[1][0]) / np.abs(res[1][0])) < 2e-5
    assert np.array_equal(res[0][1], res[1][1])
    assert list(res[0][1]) == [i % S for i in range(len(lens))]


def test_tc_route_full_size_config2(tc_route):
    S, K, D, T = 100, 128, 39, 10000
    gs, W, M, Sg = build_set(S, K, D)
    X = synth.draw(W[37], M[37], Sg[37], T, seed=5)
    scores = np.array(gs.predict_one_scores(X))
    assert int(np.argmax(scores)) == 37
    a, b = np.array(gs.predict_one_scores(X[:3333])), np.array(gs.predict_one_scores(X[3333:]))
    assert np.max(np.abs((a + b) - scores) / np.abs(scores)) < 1e-7      # additivity (per-tile partials are fp32 warp sums here)
    for i in (0, 37, 99):
        ex = gmm_ref.score_all(W[i], M[i], Sg[i], X[:1000], gmm_ref.MODE_LOGDOMAIN)
        assert abs(gs.gmms[i].score_all(X[:1000]) - ex) < 2e-5 * abs(ex)


def test_tc_route_limits_and_fallbacks(tc_route):
    """the shapes the tensor route cannot take fall back to the FMA kernel (and say so); the largest one it can take works"""
    L = tc_route
    rng = np.random.default_rng(5)
    for D, K, expect in [(40, 40, 1), (55, 40, 1), (63, 40, 1), (64, 40, 0), (100, 16, 0)]:   # 2 D + 1 <= 127: PHI head + tail fit the 128 TMEM columns beside three slots
        W, M, Sg = synth.models(2, K, D, seed0=900 + D)
        gs = GMMSet(gmm_order=K)
        for i in range(2):
            gs.gmms.append(GMM(K).set_params(W[i], M[i], Sg[i]))
            gs.y.append(i)
        X = synth.draw(W[1], M[1], Sg[1], 300, seed=D)
        sums = np.array(gs.predict_one_scores(X))
        assert L.srb200_last_score_route() == expect
        for i in range(2):
            ex = gmm_ref.score(W[i], M[i], Sg[i], X, gmm_ref.MODE_LOGDOMAIN).sum()
            assert abs(sums[i] - ex) < 2e-5 * abs(ex)
        assert int(np.argmax(sums)) == 1
    # every sigma tiny: far more exact add-back entries than a block's list holds -> the whole set stays on the FMA route
    K, D = 64, 20
    w = np.full(K, 1.0 / K)
    m = rng.normal(size=(K, D))
    s = np.full((K, D), 0.05)
    g = GMM(K).set_params(w, m, s)
    X = m[rng.integers(0, K, 200)] + 0.05 * rng.normal(size=(200, D))
    got = g.score(X)
    assert L.srb200_last_score_route() == 0
    assert np.max(rel_err(got, gmm_ref.score(w, m, s, X, gmm_ref.MODE_LOGDOMAIN))) < TOL
    # within the limits (16 exactly corrected rows per 32-column group): 3 hard dims in a third of the mixtures stays on route 1
    s2 = np.full((K, D), 0.7)
    s2[::3, :3] = 0.04
    g2 = GMM(K).set_params(w, m, s2)
    X2 = m[rng.integers(0, K, 300)] + s2[0] * rng.normal(size=(300, D))
    got2 = g2.score(X2)
    assert L.srb200_last_score_route() == 1
    assert np.max(rel_err(got2, gmm_ref.score(w, m, s2, X2, gmm_ref.MODE_LOGDOMAIN))) < TOL


@pytest.mark.parametrize("T", [1, 2, 127, 128, 129, 255, 256, 257])
def test_tc_route_tile_boundaries(tc_route, T):
    S, K, D = 3, 48, 13
    gs, W, M, Sg = build_set(S, K, D, seed0=77)
    X = synth.draw(W[0], M[0], Sg[0], T, seed=T)
    sums = np.array(gs.predict_one_scores(X))
    for i in range(S):
        ex = gmm_ref.score(W[i], M[i], Sg[i], X, gmm_ref.MODE_LOGDOMAIN)
        assert np.max(rel_err(gs.gmms[i].score(X), ex)) < TOL
        assert abs(sums[i] - ex.sum()) < 2e-5 * abs(ex.sum())
