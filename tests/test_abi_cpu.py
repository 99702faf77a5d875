"""CPU-only checks of the drop-in boundary: libsrb200.so loads, exports every symbol include/srb200.h declares,
mirrors the reference's struct layout, and FAILS LOUDLY (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from speaker_recognition_b200 import _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "srb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_symbols_exported_and_bound():
    L = _lib.lib()
    names = header_functions()
    assert len(names) >= 45
    for ref in ["new_gmm", "load", "dump", "train_model", "train_model_from_ubm", "score_all", "score_batch",
                "score_instance", "get_dim", "get_nr_mixtures"]:            # src/gmm/src/pygmm.hh:28-41
        assert ref in names
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for n in names:
        assert n in exported, "header declares %s but the library does not export it" % n
        assert n in _lib.PROTOTYPES, "no ctypes prototype for %s" % n
        assert getattr(L, n) is not None
    assert set(_lib.PROTOTYPES) <= set(names)


def test_library_is_sm100a_cuda_and_self_contained():
    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH], text=True)
    assert "sm_100a" in out
    sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True, stderr=subprocess.STDOUT)
    assert "gmm_score_kernel" in sass and "mfcc_frames_kernel" in sass and "em_accumulate_kernel" in sass
    assert "UBLKCP" in sass and "FFMA" in sass       # TMA bulk copy (cp.async.bulk) + the FMA hot loop
    ldd = subprocess.check_output(["ldd", _lib.LIB_PATH], text=True)
    assert "libtorch" not in ldd and "libcudart" not in ldd   # static cudart, no torch types behind the C ABI


def test_parameter_struct_layout_matches_reference():
    # struct Parameter, src/gmm/src/pygmm.hh:12-26: 3 ints, pad, 2 doubles, 4 ints
    P = _lib.Parameter
    assert C.sizeof(P) == 48
    assert P.nr_instance.offset == 0 and P.nr_dim.offset == 4 and P.nr_mixture.offset == 8
    assert P.min_covar.offset == 16 and P.threshold.offset == 24
    assert P.nr_iteration.offset == 32 and P.init_with_kmeans.offset == 36
    assert P.concurrency.offset == 40 and P.verbosity.offset == 44


def test_model_text_io_is_host_side_and_reference_compatible(tmp_path, golden_gmm):
    from oracle import gmm_ref
    L = _lib.lib()
    c = golden_gmm.case("ubm32")
    ref_path = str(tmp_path / "ref.model")
    gmm_ref.dump_model(ref_path, c["w"], c["mean"], c["sigma"], precision=6)      # what the reference writes
    g = _lib.check_ptr(L.load(ref_path.encode()))
    assert L.get_dim(g) == 34 and L.get_nr_mixtures(g) == 32
    out6 = str(tmp_path / "ours6.model")
    L.dump(g, out6.encode())
    assert open(out6).read() == open(ref_path).read()                             # byte-identical 6-digit format
    K, D = 32, 34
    w, m, s = np.empty(K), np.empty((K, D)), np.empty((K, D))
    _lib.check(L.srb200_gmm_get_params(g, _lib.dptr(w), _lib.dptr(m), _lib.dptr(s)))
    rw, rm, rs = gmm_ref.load_model(ref_path)
    assert np.array_equal(w, rw) and np.array_equal(m, rm) and np.array_equal(s, rs)
    # lossless variant
    _lib.check(L.srb200_gmm_set_params(g, K, D, _lib.dptr(c["w"]), _lib.dptr(np.ascontiguousarray(c["mean"])),
                                       _lib.dptr(np.ascontiguousarray(c["sigma"]))))
    out17 = str(tmp_path / "ours17.model")
    _lib.check(L.srb200_gmm_dump_precise(g, out17.encode(), 17))
    rw, rm, rs = gmm_ref.load_model(out17)
    assert np.array_equal(rw, c["w"]) and np.array_equal(rm, c["mean"]) and np.array_equal(rs, c["sigma"])
    L.free_gmm(g)


def test_error_paths_do_not_abort():
    L = _lib.lib()
    assert not L.new_gmm(4, 2)                                   # non-diagonal: the reference throws (gmm.cc:211-215)
    assert "diagonal" in _lib.last_error()
    assert not L.load(b"/nonexistent/file.model")                # the reference silently returns an empty model
    assert "cannot open" in _lib.last_error()
    g = L.new_gmm(4, 1)
    w = np.ones(4) / 4
    bad = np.zeros((4, 3))
    assert L.srb200_gmm_set_params(g, 4, 3, _lib.dptr(w), _lib.dptr(bad), _lib.dptr(bad)) != 0   # sigma must be > 0
    L.free_gmm(g)


@pytest.mark.skipif(_lib.lib().srb200_device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when it cannot run on the GPU."""
    from speaker_recognition_b200.gmm.pygmm import GMM
    from speaker_recognition_b200.feature import MFCC
    g = GMM(2)
    g.set_params([0.5, 0.5], [[0.0, 0.0], [1.0, 1.0]], [[1.0, 1.0], [1.0, 1.0]])
    with pytest.raises(_lib.SrbError, match="no CPU path|no usable CUDA"):
        g.score(np.zeros((3, 2)))
    with pytest.raises(_lib.SrbError):
        g.fit(np.zeros((10, 2)))
    with pytest.raises(_lib.SrbError):
        MFCC.extract(16000, np.zeros(8000, dtype=np.int16))
    X = np.zeros((1, 2))
    pp = (C.POINTER(C.c_double) * 1)(X.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.isnan(_lib.lib().score_all(g.gmm, pp, 1, 2, 1))    # reference ABI returns double: NaN + last_error


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "speaker-recognition_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt and "gmm_oracle" not in txt, f


def _tc_stats(models):
    """(eligible, blocks, hard entries, max per block) from the host-side packer of the tensor-core route"""
    from speaker_recognition_b200.gmm.pygmm import GMM
    gs = [GMM(len(w)).set_params(w, m, s) for (w, m, s) in models]
    arr = (C.c_void_p * len(gs))(*[g.gmm for g in gs])
    b, h, mx = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = _lib.lib().srb200_tc_pack_stats(arr, len(gs), C.byref(b), C.byref(h), C.byref(mx))
    return rc, b.value, h.value, mx.value


def test_tensor_route_packer_takes_real_parameter_models(golden_gmm, golden_real):
    """host-side eligibility of the tensor-core scorer (no GPU needed): every UBM the reference ships, and a 100-speaker set
    resampled from them, fit the exact add-back list with room to spare (round 1 keyed the list on sigma < 0.1 and overflowed:
    12-17 % of the entries of these models)"""
    from conftest import shipped_ubm_pools
    from speaker_recognition_b200 import synth
    pools = shipped_ubm_pools(golden_gmm, golden_real)
    for w, m, s in pools:
        rc, blocks, hard, mx = _tc_stats([(w, m, s)])
        assert rc == 1 and mx <= 64 and hard <= 0.01 * m.size, (rc, blocks, hard, mx)
    W, M, Sg = synth.realparam_models(100, 32, pools, seed0=5000)
    rc, blocks, hard, mx = _tc_stats([(W[i], M[i], Sg[i]) for i in range(100)])
    assert rc == 1 and blocks == 25 and mx <= 100 and hard <= 0.01 * M.size, (rc, blocks, hard, mx)
    W, M, Sg = synth.models(8, 128, 39, seed0=1000)            # the synthetic C2 models: 1 % of the entries at the sigma floor
    rc, blocks, hard, mx = _tc_stats([(W[i], M[i], Sg[i]) for i in range(8)])
    assert rc == 1 and hard <= 0.005 * M.size, (rc, blocks, hard, mx)
    rng = np.random.default_rng(5)                              # every sigma tiny against the spread of the means: not eligible
    m = rng.normal(size=(64, 20))
    rc, _, _, _ = _tc_stats([(np.full(64, 1 / 64), m, np.full((64, 20), 0.05))])
    assert rc == 0
