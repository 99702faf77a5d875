"""ctypes binding of csrc/libsrb200.so (include/srb200.h).  This is the ONLY compute backend of the package:
if the library is missing or no CUDA device is usable, calls raise -- there is no CPU fallback.

Unlike the reference's binding (src/gmm/python/pygmm.py:16,29-31: restype declared for two functions
only, so GMM* is truncated to a C int on LP64) every prototype is declared."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRB200_LIB") or os.path.join(_HERE, "csrc", "libsrb200.so")   # SRB200_LIB: A/B builds of the same library (experiments)


class Parameter(C.Structure):
    """struct Parameter (include/srb200.h; src/gmm/src/pygmm.hh:12-26; GMMParameter pygmm.py:18-27)."""
    _fields_ = [("nr_instance", C.c_int), ("nr_dim", C.c_int), ("nr_mixture", C.c_int),
                ("min_covar", C.c_double), ("threshold", C.c_double), ("nr_iteration", C.c_int),
                ("init_with_kmeans", C.c_int), ("concurrency", C.c_int), ("verbosity", C.c_int)]


class SrbError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes): every symbol include/srb200.h declares
_vp, _ci, _cd, _cll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
_dp, _fp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
_dpp, _llp, _i16p = C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_longlong), C.POINTER(C.c_int16)
_PP = C.POINTER(Parameter)
PROTOTYPES = {
    # the reference's ten symbols (pygmm.hh:28-41)
    "new_gmm": (_vp, [_ci, _ci]),
    "load": (_vp, [C.c_char_p]),
    "dump": (None, [_vp, C.c_char_p]),
    "train_model": (None, [_vp, _dpp, _PP]),
    "train_model_from_ubm": (None, [_vp, _vp, _dpp, _PP]),
    "score_all": (_cd, [_vp, _dpp, _ci, _ci, _ci]),
    "score_batch": (None, [_vp, _dpp, _dp, _ci, _ci, _ci]),
    "score_instance": (_cd, [_vp, _dp, _ci]),
    "get_dim": (_ci, [_vp]),
    "get_nr_mixtures": (_ci, [_vp]),
    # additions
    "free_gmm": (None, [_vp]),
    "srb200_last_error": (C.c_char_p, []),
    "srb200_version": (C.c_char_p, []),
    "srb200_device_count": (_ci, []),
    "srb200_set_device": (_ci, [_ci]),
    "srb200_get_device": (_ci, []),
    "srb200_set_compat_clamp": (_ci, [_ci]),
    "srb200_get_compat_clamp": (_ci, []),
    "srb200_set_score_path": (_ci, [_ci]),
    "srb200_get_score_path": (_ci, []),
    "srb200_last_score_route": (_ci, []),
    "srb200_last_em_route": (_ci, []),
    "srb200_set_em_path": (_ci, [_ci]),
    "srb200_score_launch_count": (_cll, []),
    "srb200_tc_pack_stats": (_ci, [C.POINTER(_vp), _ci, _ip, _ip, _ip]),
    "srb200_tc_pack_export": (_ci, [C.POINTER(_vp), _ci, _ip, C.POINTER(C.c_uint16), _ip, _fp]),
    "srb200_gmm_set_params": (_ci, [_vp, _ci, _ci, _dp, _dp, _dp]),
    "srb200_gmm_get_params": (_ci, [_vp, _dp, _dp, _dp]),
    "srb200_gmm_dump_precise": (_ci, [_vp, C.c_char_p, _ci]),
    "srb200_score_batch_f64": (_ci, [_vp, _dp, _ci, _ci, _dp]),
    "srb200_score_all_f64": (_ci, [_vp, _dp, _ci, _ci, _dp]),
    "srb200_score_batch_f32": (_ci, [_vp, _fp, _ci, _ci, _fp]),
    "srb200_train_f64": (_ci, [_vp, _dp, _PP, C.c_uint64, _ip, _dp]),
    "srb200_train_from_ubm_f64": (_ci, [_vp, _vp, _dp, _PP, _ip, _dp]),
    "srb200_train_many_from_ubm_f64": (_ci, [C.POINTER(_vp), _ci, _vp, _dp, _ip, _ci, _PP, _ip, _dp]),
    "srb200_kmeans_init_f64": (_ci, [_dp, _ci, _ci, _ci, C.c_uint64, _dp, _dp]),
    "srb200_kmeans_lloyd_f64": (_ci, [_dp, _ci, _ci, _ci, _dp, _dp, _dp, _ip]),
    "srb200_gmmset_create": (_vp, [C.POINTER(_vp), _ci]),
    "srb200_gmmset_free": (None, [_vp]),
    "srb200_gmmset_size": (_ci, [_vp]),
    "srb200_gmmset_dim": (_ci, [_vp]),
    "srb200_gmmset_set_ranked": (_ci, [_vp, _ci]),
    "srb200_gmmset_score_f64": (_ci, [_vp, _dp, _ci, _ci, _dp]),
    "srb200_gmmset_score_f32": (_ci, [_vp, _fp, _ci, _ci, _dp]),
    "srb200_gmmset_score_dev": (_ci, [_vp, _vp, _ci, _ci, _ip, _ci, _vp, _vp, _ci, _vp, _vp]),
    "srb200_em_stats_len": (_ci, [_ci, _ci]),
    "srb200_em_estep_dev": (_ci, [_vp, _vp, _ci, _ci, _ci, _vp, _vp]),
    "srb200_em_mstep_dev": (_ci, [_vp, _vp, _cll, _cd, _ci, _vp, _vp]),
    "srb200_gmm_sync_from_device": (_ci, [_vp]),
    "srb200_mfcc_create": (_vp, [_ci, _cd, _cd, _ci, _ci, _ci, _cd]),
    "srb200_mfcc_free": (None, [_vp]),
    "srb200_mfcc_frame_len": (_ci, [_vp]),
    "srb200_mfcc_frame_shift": (_ci, [_vp]),
    "srb200_mfcc_num_frames": (_ci, [_vp, _cll]),
    "srb200_mfcc_out_dim": (_ci, [_vp, _ci]),
    "srb200_mfcc_extract_i16": (_ci, [_vp, _i16p, _cll, _ci, _ci, _dp, _ip]),
    "srb200_mfcc_extract_f32": (_ci, [_vp, _fp, _cll, _ci, _ci, _dp, _ip]),
    "srb200_mfcc_extract_f64": (_ci, [_vp, _dp, _cll, _ci, _ci, _dp, _ip]),
    "srb200_mfcc_extract_dev": (_ci, [_vp, _vp, _llp, _ci, _ci, _vp, _ci, _ip, _vp]),
    "srb200_mfcc_rows_total": (_cll, [_vp, _llp, _ci, _ci]),
    "srb200_lpc_extract_i16": (_ci, [_vp, _i16p, _cll, _ci, _dp, _ip]),
    "srb200_lpc_extract_f32": (_ci, [_vp, _fp, _cll, _ci, _dp, _ip]),
    "srb200_lpc_extract_f64": (_ci, [_vp, _dp, _cll, _ci, _dp, _ip]),
    "srb200_lpc_extract_dev": (_ci, [_vp, _vp, _llp, _ci, _ci, _vp, _ci, _ci, _ip, _vp]),
    "srb200_stream_push_i16": (_ci, [_vp, _vp, _vp, _ci, _ci, _ci, _vp]),
    "srb200_remove_silence_i16": (_ci, [_ci, _i16p, _cll, _cd, _cd, _cd, _i16p, _llp]),
    "srb200_remove_silence_u8": (_ci, [_ci, C.POINTER(C.c_uint8), _cll, _cd, _cd, _cd, C.POINTER(C.c_uint8), _llp]),
}


def lib():
    """Load libsrb200.so (once).  Raises if it has not been built: the product has no other backend."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SrbError("libsrb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `python -m speaker_recognition_b200.build`; there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)          # AttributeError = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().srb200_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise SrbError("%s%s" % (what + ": " if what else "", last_error() or "unknown failure (rc=%r)" % rc))


def check_ptr(p, what=""):
    if not p:
        raise SrbError("%s%s" % (what + ": " if what else "", last_error() or "null handle"))
    return p


def require_gpu():
    if lib().srb200_device_count() <= 0:
        raise SrbError("no usable CUDA device: speaker_recognition_b200 runs only on the GPU (no CPU fallback)")


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def dptr(a):
    return a.ctypes.data_as(_dp)


def fptr(a):
    return a.ctypes.data_as(_fp)
