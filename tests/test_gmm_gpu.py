"""GPU parity tests for the diagonal-GMM path, through the C ABI (include/srb200.h), against the oracle
(oracle/gmm_oracle.c, pinned to the compiled reference) and the committed reference fixtures (tests/golden).

Tolerance (north_star): |ours - ref64| <= 1e-4 * max(1, |ref64|) per frame vs the exact float64 value;
utterance sums and argmax vs the reference's own (fast-exp) results; EM parameters 1e-4 relative."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err
from oracle import gmm_ref
from speaker_recognition_b200 import _lib, synth
from speaker_recognition_b200.gmm.pygmm import GMM
from speaker_recognition_b200.gmmset import GMMSet, GMMSetPyGMM

pytestmark = pytest.mark.gpu
TOL = 1e-4

# Every test of this module runs on both scoring routes (include/srb200.h: srb200_set_score_path): the fp32 FMA kernel and
# the tcgen05 3xTF32 kernel.  The north-star bound is TOL for both; utterance sums and exact-arithmetic properties get a
# tighter, route-specific bound: rt(fma_bound, tc_bound).
_ROUTE = {"path": 0}


@pytest.fixture(autouse=True, params=[0, 1], ids=["fma", "tc"])
def score_route(request):
    L = _lib.lib()
    _lib.check(L.srb200_set_score_path(request.param))
    _ROUTE["path"] = request.param
    yield request.param
    _lib.check(L.srb200_set_score_path(-1))


def rt(fma, tc):
    return tc if _ROUTE["path"] == 1 else fma



def make_gmm(w, m, s):
    return GMM(len(w)).set_params(w, m, s)


@pytest.mark.parametrize("tag", ["ubm32", "ubm64", "ubm256"])
def test_score_matches_reference_fixture_real_models(golden_gmm, tag):
    c = golden_gmm.case(tag)
    w, m, s, X = c["w"], c["mean"], c["sigma"], c["X"]
    g = make_gmm(w, m, s)
    got = g.score(X)
    exact = gmm_ref.score(w, m, s, X, gmm_ref.MODE_LOGDOMAIN)
    assert np.max(rel_err(got, exact)) < TOL                     # vs float64 truth
    assert np.max(rel_err(got, c["score_fastexp"])) < 5e-4       # vs the reference's own fast-exp values (its error, s7 2b)
    tot = g.score_all(X)
    assert abs(tot - float(c["score_all"])) < rt(1e-6, 1e-5) * abs(float(c["score_all"]))   # utterance sum vs compiled reference
    assert abs(tot - got.sum()) < 1e-5 * abs(tot)


def test_reference_abi_double_pp_entry_points(golden_gmm):
    """score_all / score_batch / score_instance exactly as pygmm.py calls them (double** rows, pygmm.cc:98-120)."""
    c = golden_gmm.case("ubm32")
    w, m, s, X = c["w"], c["mean"], c["sigma"], np.ascontiguousarray(c["X"])
    g = make_gmm(w, m, s)
    L = _lib.lib()
    n, d = X.shape
    rows = (C.POINTER(C.c_double) * n)(*[C.cast(X.ctypes.data + i * d * 8, C.POINTER(C.c_double)) for i in range(n)])
    out = np.zeros(n)
    L.score_batch(g.gmm, rows, _lib.dptr(out), n, d, 8)
    assert _lib.last_error() == ""
    assert np.max(rel_err(out, c["score_fastexp"])) < 5e-4
    tot = L.score_all(g.gmm, rows, n, d, 8)
    assert abs(tot - float(c["score_all"])) < rt(1e-6, 1e-5) * abs(float(c["score_all"]))
    one = L.score_instance(g.gmm, _lib.dptr(X[7].copy()), d)     # asserts in the reference; works here
    assert abs(one - out[7]) < 1e-5 * max(1, abs(out[7]))
    assert np.isnan(L.score_all(g.gmm, rows, n, d - 1, 1)) and "nr_dim" in _lib.last_error()


def test_multispeaker_set_argmax_and_sums(golden_gmm):
    c = golden_gmm.case("c2")
    W, M, S, X = c["w"], c["mean"], c["sigma"], c["X"]
    gs = GMMSet(gmm_order=128)
    for i in range(len(W)):
        gs.gmms.append(make_gmm(W[i], M[i], S[i]))
        gs.y.append("spk%d" % i)
    scores = np.array(gs.predict_one_scores(X))
    ref = c["score_fastexp"].sum(axis=1)
    assert np.max(np.abs(scores - ref) / np.abs(ref)) < rt(1e-6, 1e-5)
    assert gs.predict_one(X) == "spk3"
    # per-model path agrees with the batched path
    assert abs(gs.gmm_score(gs.gmms[1], X) - scores[1]) < rt(1e-6, 1e-5) * abs(scores[1])
    g2 = GMMSetPyGMM(gmm_order=128)
    g2.gmms, g2.y = gs.gmms, gs.y
    assert g2.predict_one(X) == "spk3"


def test_underflow_clamp_matches_reference(golden_gmm):
    """safe_log (gmm.cc:34-38): frames whose probability underflows float64 score log(1e-15)."""
    c, u = golden_gmm.case("ubm32"), golden_gmm.case("underflow")
    w, m, s, X, ref = c["w"], c["mean"], c["sigma"], u["X"], u["score_fastexp"]
    g = make_gmm(w, m, s)
    L = _lib.lib()
    assert L.srb200_get_compat_clamp() == 1
    got = g.score(X)
    exact = gmm_ref.score(w, m, s, X, gmm_ref.MODE_LOGDOMAIN)
    dom = exact > -700                       # parity domain: outside the float64 denormal band
    assert np.max(rel_err(got[dom], ref[dom])) < 5e-4
    deep = exact < gmm_ref.DENORM_LOG_MIN - 1.0
    assert deep.sum() > 3
    assert np.allclose(got[deep], gmm_ref.SAFE_LOG_FLOOR, atol=1e-5) and np.allclose(ref[deep], gmm_ref.SAFE_LOG_FLOOR)
    try:
        _lib.check(L.srb200_set_compat_clamp(0))
        raw = g.score(X)
        assert np.max(rel_err(raw, exact)) < TOL      # exact log-domain values everywhere, no clamp
    finally:
        _lib.check(L.srb200_set_compat_clamp(1))


@pytest.mark.parametrize("K,D,n", [(1, 1, 1), (3, 5, 7), (10, 13, 311), (32, 26, 255), (33, 39, 257), (100, 40, 1000), (7, 128, 300)])
def test_shapes_ragged_and_padding(K, D, n):
    w, m, s = synth.model(K, D, seed=K * 100 + D, floor_frac=0.05)
    X = synth.draw(w, m, s, n, seed=n)
    g = make_gmm(w, m, s)
    got = g.score(X)
    exact = gmm_ref.score(w, m, s, X, gmm_ref.MODE_LOGDOMAIN)
    assert got.shape == (n,)
    assert np.max(rel_err(got, exact)) < TOL
    assert abs(g.score_all(X) - exact.sum()) < 1e-5 * max(1.0, abs(exact.sum()))


def test_empty_and_bad_inputs():
    w, m, s = synth.model(4, 3, seed=1)
    g = make_gmm(w, m, s)
    assert g.score(np.zeros((0, 3))).shape == (0,)
    assert g.score_all(np.zeros((0, 3))) == 0.0
    with pytest.raises(_lib.SrbError):
        g.score(np.zeros((5, 4)))
    with pytest.raises(_lib.SrbError):
        GMM(4).score(np.zeros((5, 3)))            # no parameters yet
    w0 = w.copy()
    w0[1] = 0.0                                    # a zero-weight mixture must simply vanish
    g0 = make_gmm(w0, m, s)
    X = synth.draw(w, m, s, 64, seed=2)
    assert np.max(rel_err(g0.score(X), gmm_ref.score_numpy(np.where(w0 > 0, w0, 1e-300), m, s, X))) < TOL


def test_full_size_properties_config2():
    """BASELINE config 2 shape (100 speakers x 128 mixtures, D=39, 10 000 frames): size-independent properties."""
    S, K, D, T = 100, 128, 39, 10000
    W, M, Sg = synth.models(S, K, D, seed0=1000)
    true_spk = 37
    X = synth.draw(W[true_spk], M[true_spk], Sg[true_spk], T, seed=5)
    gs = GMMSet(gmm_order=K)
    for i in range(S):
        gs.gmms.append(make_gmm(W[i], M[i], Sg[i]))
        gs.y.append(i)
    scores = np.array(gs.predict_one_scores(X))
    assert int(np.argmax(scores)) == true_spk and gs.predict_one(X) == true_spk
    # additivity over a frame split (sum of parts == whole)
    a, b = np.array(gs.predict_one_scores(X[:3333])), np.array(gs.predict_one_scores(X[3333:]))
    assert np.max(np.abs((a + b) - scores) / np.abs(scores)) < rt(1e-9, 1e-7)
    # permutation invariance of the utterance sum
    perm = np.random.default_rng(0).permutation(T)
    assert np.max(np.abs(np.array(gs.predict_one_scores(X[perm])) - scores) / np.abs(scores)) < rt(1e-9, 1e-7)
    # spot-check 5 speakers against the float64 oracle on a 1 000-frame slice
    for i in (0, 17, 37, 63, 99):
        ex = gmm_ref.score_all(W[i], M[i], Sg[i], X[:1000], gmm_ref.MODE_LOGDOMAIN)
        got = gs.gmms[i].score_all(X[:1000])
        assert abs(got - ex) < rt(1e-6, 1e-5) * abs(ex)


# ------------------------------------------------------------------------------------------------
# EM
# ------------------------------------------------------------------------------------------------
def em_iterations_dev(w, m, s, X, iters, min_covar=1e-3, mode=0, ubm=None):
    """drive the device building blocks (estep_dev -> [all-reduce] -> mstep_dev) from an injected start"""
    import torch
    L = _lib.lib()
    g = make_gmm(w, m, s)
    K, D = m.shape
    ld = (D + 3) & ~3
    Xd = torch.zeros((len(X), ld), dtype=torch.float32, device="cuda")
    Xd[:, :D] = torch.from_numpy(np.asarray(X, dtype=np.float32)).cuda()
    stats = torch.zeros(L.srb200_em_stats_len(K, D), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    hist = []
    for _ in range(iters):
        _lib.check(L.srb200_em_estep_dev(g.gmm, Xd.data_ptr(), len(X), ld, D, stats.data_ptr(), st))
        hist.append(stats.cpu().numpy().copy())
        _lib.check(L.srb200_em_mstep_dev(g.gmm, stats.data_ptr(), len(X), min_covar, mode, ubm.gmm if ubm else None, st))
    torch.cuda.synchronize()
    return g, hist


def test_em_iteration_matches_reference_fixture(golden_gmm):
    e = golden_gmm.case("em")
    g, hist = em_iterations_dev(e["w0"], e["mean0"], e["sigma0"], e["X"], 2)
    w, m, s = g.get_params()
    assert np.max(np.abs(w - e["w2"]) / e["w2"]) < TOL
    assert np.max(rel_err(m, e["mean2"])) < TOL and np.max(np.abs(s - e["sigma2"]) / e["sigma2"]) < TOL
    g1, _ = em_iterations_dev(e["w0"], e["mean0"], e["sigma0"], e["X"], 1)
    w, m, s = g1.get_params()
    assert np.max(np.abs(w - e["w1"]) / e["w1"]) < TOL
    assert np.max(rel_err(m, e["mean1"])) < TOL and np.max(np.abs(s - e["sigma1"]) / e["sigma1"]) < TOL
    # statistics tail: total log-likelihood of the start model and frames used
    K, D = e["mean0"].shape
    ll0 = gmm_ref.score_all(e["w0"], e["mean0"], e["sigma0"], e["X"])
    assert abs(hist[0][-2] - ll0) < 1e-6 * abs(ll0) and hist[0][-1] == len(e["X"])
    assert abs(hist[0][:K].sum() - len(e["X"])) < 1e-4 * len(e["X"])     # sum_k N_k = n


def test_map_adaptation_matches_reference_fixture(golden_gmm):
    c, mp = golden_gmm.case("ubm32"), golden_gmm.case("map")
    ubm = make_gmm(c["w"], c["mean"], c["sigma"])
    spk = GMM(32, nr_iteration=1)
    spk.fit(mp["X"], ubm=ubm)                       # train_model_from_ubm, one iteration from the UBM copy
    w, m, s = spk.get_params()
    assert np.array_equal(w, c["w"]) and np.array_equal(s, c["sigma"])      # frozen (gmmubm.cc:40-42,76-78)
    assert np.max(rel_err(m, mp["mean1"])) < TOL
    assert spk.last_iterations == 1


@pytest.mark.parametrize("K,D,n", [(16, 13, 5000), (40, 39, 3000), (5, 50, 2000)])
def test_em_trajectory_matches_oracle(K, D, n):
    """several iterations from an injected start vs the float64 oracle (gmm_oracle.c em_iteration)"""
    wt, mt, st_ = synth.model(K, D, seed=31, floor_frac=0.0)
    X = synth.draw(wt, mt, st_, n, seed=32)
    w0 = np.ones(K) / K
    m0 = X[np.random.default_rng(3).choice(n, K, replace=False)]
    s0 = np.tile(X.std(axis=0, ddof=1), (K, 1))
    iters = 4
    g, _ = em_iterations_dev(w0, m0, s0, X, iters)
    w, m, s = w0, m0, s0
    for _ in range(iters):
        w, m, s = gmm_ref.em_iteration(w, m, s, X)
    gw, gm, gsg = g.get_params()
    assert np.max(np.abs(gw - w) / np.maximum(w, 1e-6)) < 5e-4
    assert np.max(rel_err(gm, m)) < 5e-4 and np.max(np.abs(gsg - s) / s) < 5e-4


def test_fit_converges_and_improves_likelihood():
    """full GMM.fit (reference-style random-row init, gmm.cc:306-361): statistical parity with the reference trainer"""
    K, D, n = 8, 13, 6000
    wt, mt, st_ = synth.model(K, D, seed=41, floor_frac=0.0)
    X = synth.draw(wt, mt, st_, n, seed=42)
    g = GMM(K, nr_iteration=30, threshold=1e-6, seed=7)
    g.fit(X)
    assert g.get_dim() == D and 2 <= g.last_iterations <= 30
    ll = g.score_all(X)
    ll_true = gmm_ref.score_all(wt, mt, st_, X, gmm_ref.MODE_LOGDOMAIN)
    assert ll > ll_true - 0.05 * abs(ll_true)           # close to the generating model's likelihood
    w, m, s = g.get_params()
    assert abs(w.sum() - 1) < 1e-9 and s.min() >= np.sqrt(1e-3) - 1e-12
    # determinism under a fixed seed
    g2 = GMM(K, nr_iteration=30, threshold=1e-6, seed=7).fit(X)
    assert np.array_equal(g2.get_params()[1], m)
    # reference ABI: train_model with double** rows
    L = _lib.lib()
    Xc = np.ascontiguousarray(X)
    rows = (C.POINTER(C.c_double) * n)(*[C.cast(Xc.ctypes.data + i * D * 8, C.POINTER(C.c_double)) for i in range(n)])
    h = L.new_gmm(K, 1)
    p = _lib.Parameter(n, D, K, 1e-3, 1e-6, 10, 0, 4, 0)
    L.train_model(h, rows, C.byref(p))
    assert _lib.last_error() == "" and L.get_dim(h) == D
    L.free_gmm(h)


def test_pickle_roundtrip_of_speaker_set(tmp_path):
    import pickle
    W, M, S = synth.models(3, 8, 13, seed0=50)
    gs = GMMSetPyGMM(gmm_order=8)
    for i in range(3):
        gs.gmms.append(make_gmm(W[i], M[i], S[i]))
        gs.y.append("p%d" % i)
    X = synth.draw(W[2], M[2], S[2], 200, seed=9)
    before = gs.predict_one_scores(X)
    gs.before_pickle()
    blob = pickle.dumps(gs)
    gs.after_pickle()
    g2 = pickle.loads(blob)
    g2.after_pickle()
    assert g2.predict_one(X) == "p2"
    assert np.allclose(g2.predict_one_scores(X), before, rtol=1e-12)


def test_dist_em_train_with_device_backend_single_rank():
    """speaker_recognition_b200.dist.em_train + DeviceEM (the frame-sharded trainer, world size 1) vs the float64 oracle"""
    import torch
    from speaker_recognition_b200 import dist as sdist
    K, D, n = 12, 20, 4000
    wt, mt, st_ = synth.model(K, D, seed=61, floor_frac=0.0)
    X = synth.draw(wt, mt, st_, n, seed=62)
    w0 = np.ones(K) / K
    m0 = X[np.random.default_rng(4).choice(n, K, replace=False)]
    s0 = np.tile(X.std(axis=0, ddof=1), (K, 1))
    g = make_gmm(w0, m0, s0)
    ld = (D + 3) & ~3
    Xd = torch.zeros((n, ld), dtype=torch.float32, device="cuda")
    Xd[:, :D] = torch.from_numpy(X.astype(np.float32)).cuda()
    iters, ll = sdist.em_train(sdist.DeviceEM(g, Xd, n, ld, D), n, 4, threshold=0.0)
    assert iters == 4
    w, m, s = w0, m0, s0
    for _ in range(4):
        w, m, s = gmm_ref.em_iteration(w, m, s, X)
    gw, gm, gsg = g.get_params()
    assert np.max(rel_err(gm, m)) < 5e-4 and np.max(np.abs(gsg - s) / s) < 5e-4 and np.max(np.abs(gw - w) / w) < 5e-4
    assert abs(ll - gmm_ref.score_all(w, m, s, X)) < 1e-4 * abs(ll)


def test_batched_map_enrolment_matches_the_reference_per_speaker(golden_gmm, golden_real):
    """srb200_train_many_from_ubm_f64 (what ModelInterface.train runs with a UBM): six speakers with different amounts of data in
    ONE call equal, speaker by speaker, two iterations of the reference's GMMUBMTrainerBaseline (fixture from ref_map_iteration),
    and equal the one-speaker-at-a-time path of this library"""
    c, m = golden_gmm.case("ubm32"), golden_real.case("mapset")
    ubm = GMM(32).set_params(c["w"], c["mean"], c["sigma"])
    off = m["frame_off"]
    Xs = [m["X"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    gmms = [GMM(32, nr_iteration=2, threshold=0.0) for _ in Xs]
    GMM.fit_many_from_ubm(gmms, Xs, ubm)
    for i, g in enumerate(gmms):
        w, mean, sg = g.get_params()
        assert g.last_iterations == 2
        assert np.max(np.abs(mean - m["mean2"][i]) / np.maximum(1.0, np.abs(m["mean2"][i]))) < 1e-4
        assert np.array_equal(w, c["w"]) and np.array_equal(sg, c["sigma"])        # weights and sigmas stay the UBM's (gmmubm.cc:40-81)
        one = GMM(32, nr_iteration=2, threshold=0.0).fit(Xs[i], ubm)
        assert np.max(np.abs(one.get_params()[1] - mean)) < 1e-9                   # same kernels, same order: batching changes nothing
        ex = gmm_ref.score_all(w, mean, sg, Xs[i], gmm_ref.MODE_LOGDOMAIN)
        assert abs(g.last_log_likelihood - ex) < 1e-5 * abs(ex)
    # the convergence rule is applied per speaker (gmm.cc:643-648): a huge threshold stops everyone at the second check
    gm2 = [GMM(32, nr_iteration=20, threshold=1e9) for _ in Xs]
    GMM.fit_many_from_ubm(gm2, Xs, ubm)
    assert [g.last_iterations for g in gm2] == [4] * len(Xs)


def test_model_interface_train_with_ubm_enrols_everyone_in_one_batch(tmp_path, golden_gmm):
    """ModelInterface.train with a UBM file (src/gui/interface.py:63-83): 100 speakers enrolled by ONE batched MAP adaptation; the
    result equals the per-speaker path and the enrolled set predicts its own speakers"""
    from speaker_recognition_b200.interface import ModelInterface
    c = golden_gmm.case("ubm32")
    ubm = GMM(32).set_params(c["w"], c["mean"], c["sigma"])
    ubm_file = str(tmp_path / "ubm.model")
    ubm.dump(ubm_file, precision=17)
    rng = np.random.default_rng(5)
    mi = ModelInterface()
    mi.UBM_MODEL_FILE = ubm_file
    data = {}
    for i in range(100):
        shift = 0.6 * c["sigma"] * rng.standard_normal(c["mean"].shape)
        data["p%03d" % i] = synth.draw(c["w"], c["mean"] + shift, c["sigma"], 120 + 7 * (i % 13), seed=900 + i)
        mi.features["p%03d" % i].extend(data["p%03d" % i])
    L = _lib.lib()
    mi.train(quiet=True)
    assert len(mi.gmmset.gmms) == 100 and mi.gmmset.ubm is not None
    for k in (0, 37, 99):
        nm = mi.gmmset.y[k]
        one = GMM(32).fit(data[nm], mi.gmmset.ubm)
        assert np.max(np.abs(one.get_params()[1] - mi.gmmset.gmms[k].get_params()[1])) < 1e-9
        assert mi.gmmset.predict_one(data[nm]) == nm
