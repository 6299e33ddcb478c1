"""ctypes binding of libscl_b200.so (include/scl_b200.h).  No fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscl_b200.so")

_lock = threading.Lock()
_lib = None

c_f32p = C.c_void_p
c_ptr = C.c_void_p


class MsParams(C.Structure):
    _fields_ = [("d_alpha", C.c_float), ("d_beta", C.c_float), ("alpha", C.c_float), ("beta", C.c_float),
                ("lamb", C.c_float), ("eps", C.c_float), ("ms_mining", C.c_int32), ("wfunction", C.c_int32),
                ("sumfunction", C.c_int32)]


class TupleParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dist_term", C.c_int32), ("m1", C.c_float), ("m2", C.c_float),
                ("lam", C.c_float), ("d_max_squared", C.c_float), ("f_max_squared", C.c_float)]


WF = {"exp": 0, "lin": 1, "tanh": 2}
SUMF = {"ms": 0, "plain": 1}
TUPLE_KIND = {"triplet_loss": 0, "lazy_triplet_loss": 1, "quadruplet_loss": 2, "lazy_quadruplet_loss": 3,
              "evil_triplet_loss": 4, "evil_quadruplet_loss": 5,
              "distance_quadruplet_loss": 6, "distance_lazy_quadruplet_loss": 7}
DIST_TERM = {"none": 0, "distance_loss": 1, "huber_distance_loss": 2}

# name -> (restype, argtypes); mirrors include/scl_b200.h one to one
_SIZE_P = C.POINTER(C.c_size_t)
PROTOTYPES = {
    "scl_version": (C.c_int, []),
    "scl_strerror": (C.c_char_p, [C.c_int]),
    "scl_last_error": (C.c_char_p, []),
    "scl_device_ok": (C.c_int, []),
    "scl_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "scl_get_tuning": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "scl_wms_tuple_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, _SIZE_P]),
    "scl_wms_tuple_fwd_bwd": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.POINTER(MsParams), c_ptr, c_ptr,
                                        c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_ms_flat_workspace_bytes": (C.c_int, [C.c_int, C.c_int, _SIZE_P]),
    "scl_wms_flat_fwd_bwd": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, C.POINTER(MsParams), c_ptr, c_ptr, c_ptr, c_ptr,
                                       C.c_size_t, c_ptr]),
    "scl_ms_flat_fwd_bwd": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, C.POINTER(MsParams), c_ptr, c_ptr, c_ptr, c_ptr,
                                      C.c_size_t, c_ptr]),
    "scl_tuple_loss_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _SIZE_P]),
    "scl_tuple_loss_fwd_bwd": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, C.POINTER(TupleParams), c_ptr,
                                         c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_logratio_fwd_bwd": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, C.c_int, c_ptr, c_ptr,
                                       c_ptr, C.c_size_t, c_ptr]),
    "scl_pairwise_sqdist": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
    "scl_pairwise_distance_loss_fwd_bwd": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                                     c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_netvlad_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _SIZE_P]),
    "scl_netvlad_fwd": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, C.c_size_t,
                                  c_ptr]),
    "scl_netvlad_bwd": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr,
                                  c_ptr, C.c_size_t, c_ptr]),
    "scl_pca_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, _SIZE_P]),
    "scl_pca_fwd": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_pca_bwd": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_pca_shadow_bytes": (C.c_int, [C.c_int, C.c_int, _SIZE_P]),
    "scl_pca_prepare": (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, C.c_size_t, c_ptr]),
    "scl_pca_prepared_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, _SIZE_P]),
    "scl_pca_fwd_prepared": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_pca_bwd_prepared": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_pca_center_workspace_bytes": (C.c_int, [C.c_int, C.c_int, _SIZE_P]),
    "scl_pca_center": (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_set_gemm_precision": (C.c_int, [C.c_int]),
    "scl_get_gemm_precision": (C.c_int, []),
    "scl_gemm_tf32": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                c_ptr, C.c_int, c_ptr]),
    "scl_knn_shadow_bytes": (C.c_int, [C.c_int64, C.c_int, _SIZE_P]),
    "scl_knn_build": (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, C.c_size_t, c_ptr]),
    "scl_knn_query_workspace_bytes": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, _SIZE_P]),
    "scl_knn_query": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int, C.c_int, C.c_int64, C.c_int, c_ptr,
                                c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_knn_query_begin": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int, C.c_int, c_ptr, c_ptr, C.c_size_t,
                                      c_ptr]),
    "scl_knn_query_groups": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "scl_knn_query_launch": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int, C.c_int, c_ptr, C.c_size_t, c_ptr]),
    "scl_knn_query_begin_group": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr,
                                            C.c_size_t, c_ptr]),
    "scl_knn_query_end_group": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int, C.c_int, C.c_int64, C.c_int, c_ptr,
                                          c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_knn_bound_reduce": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
    "scl_knn_query_end": (C.c_int, [c_ptr, c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int, C.c_int, C.c_int64, c_ptr, c_ptr,
                                    c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "scl_knn_timing": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "scl_knn_set_debug_scores": (C.c_int, [c_ptr, C.c_size_t]),
    "scl_topk_merge": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int64, c_ptr, c_ptr, c_ptr]),
    "scl_geo_topn": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int64, c_ptr, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "scl_recall_curves": (C.c_int, [c_ptr, C.c_int, C.c_int, c_ptr, C.c_int, c_ptr, c_ptr]),
}


class SclError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise SclError(
                        f"{LIB_PATH} is missing. Build it with `python -m soft_contrastive_learning_b200.build` "
                        "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in PROTOTYPES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


TUNING_UNSET = -2 ** 31


def set_tuning(name: str, value=None) -> None:
    """Process-wide kernel-selection knob (include/scl_b200.h: scl_set_tuning); ``None`` restores the built-in choice."""
    check(lib().scl_set_tuning(name.encode(), TUNING_UNSET if value is None else int(value)), f"scl_set_tuning({name})")


def get_tuning(name: str):
    v = C.c_int()
    check(lib().scl_get_tuning(name.encode(), C.byref(v)), f"scl_get_tuning({name})")
    return None if v.value == TUNING_UNSET else v.value


class tuning:
    """``with tuning(SCL_WMS_STREAM=1, SCL_WMS_STREAM_CFG=2): ...`` -- set knobs, restore the previous values on exit."""

    def __init__(self, **knobs):
        self.knobs = knobs

    def __enter__(self):
        self.saved = {k: get_tuning(k) for k in self.knobs}
        for k, v in self.knobs.items():
            set_tuning(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            set_tuning(k, v)
        return False


def check(status: int, what: str = "") -> None:
    if status != 0:
        L = lib()
        msg = L.scl_strerror(status).decode()
        detail = L.scl_last_error().decode() if status == -5 else ""
        raise SclError(f"{what}: {msg}" + (f" [{detail}]" if detail else ""))
