"""CPU: the C-ABI library loads, exports every symbol include/scl_b200.h declares, refuses to compute without a
B200 (no silent fallback), and the host-side logic mirrors the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from soft_contrastive_learning_b200.build import build_library
    build_library()
    from soft_contrastive_learning_b200 import _lib
    return _lib.lib()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "scl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scl_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(L):
    from soft_contrastive_learning_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in scl_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes and header disagree"


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "scl_b200.h")).read()
    assert "torch" not in src and "at::" not in src and "std::" not in src


def test_status_strings_and_version(L):
    assert L.scl_version() >= 100
    assert L.scl_strerror(0) == b"ok"
    for code in range(-7, 0):
        assert len(L.scl_strerror(code)) > 3
    assert b"unknown" in L.scl_strerror(-99)


def test_pure_size_queries_work_without_a_gpu(L):
    n = C.c_size_t()
    assert L.scl_wms_tuple_workspace_bytes(32, 25, 4096, C.byref(n)) == 0 and n.value >= 32 * 4
    assert L.scl_wms_tuple_workspace_bytes(32, 33, 4096, C.byref(n)) == -2        # S > 32
    assert L.scl_wms_tuple_workspace_bytes(32, 25, 4095, C.byref(n)) == -2        # D % 4
    assert L.scl_ms_flat_workspace_bytes(1024, 4096, C.byref(n)) == 0 and n.value >= 3 * 1024 * 1024 * 4
    assert L.scl_knn_shadow_bytes(1000, 4096, C.byref(n)) == 0 and n.value >= 1000 * 4096 * 2
    assert L.scl_knn_shadow_bytes(1000, 4097, C.byref(n)) == -1
    assert L.scl_netvlad_workspace_bytes(2, 12, 512, 64, C.byref(n)) == 0
    assert L.scl_netvlad_workspace_bytes(2, 12, 512, 32, C.byref(n)) == -7        # only K=64 exists in the reference
    assert L.scl_knn_query_workspace_bytes(100000, 256, 64, 25, C.byref(n)) == 0 and n.value > 0
    # query groups of the sharded two-phase protocol: a pure function of D and Q (every rank must see the same groups)
    ng, gq = C.c_int(), C.c_int()
    assert L.scl_knn_query_groups(4096, 10000, C.byref(ng), C.byref(gq)) == 0 and (ng.value, gq.value) == (2, 5120)
    assert L.scl_knn_query_groups(4096, 300, C.byref(ng), C.byref(gq)) == 0 and (ng.value, gq.value) == (1, 300)
    assert L.scl_knn_query_groups(4096, 5121, C.byref(ng), C.byref(gq)) == 0 and ng.value == 2 and gq.value * 2 >= 5121
    assert L.scl_knn_query_groups(4095, 300, C.byref(ng), C.byref(gq)) == -1
    assert L.scl_knn_query_groups(4096, 300, None, C.byref(gq)) == -1


def test_compute_entry_points_fail_loudly_without_b200(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    assert L.scl_device_ok() != 0
    from soft_contrastive_learning_b200 import _lib, losses
    with pytest.raises(_lib.SclError):
        losses.wms_loss(np.zeros((4, 4), np.float32), np.zeros((4, 8), np.float32), 0.8, 15.0)
    # argument validation happens before any device work
    p = _lib.MsParams()
    buf = (C.c_float * 64)()
    assert L.scl_wms_tuple_fwd_bwd(None, None, 1, 4, 8, C.byref(p), None, None, None, None, None, 0, None) == -1
    rc = L.scl_wms_tuple_fwd_bwd(C.cast(buf, C.c_void_p), C.cast(buf, C.c_void_p), 1, 4, 8, C.byref(p),
                                 C.cast(buf, C.c_void_p), None, None, None, C.cast(buf, C.c_void_p), 256, None)
    assert rc in (-5, -6)          # CUDA error / wrong arch: never a computed result
    # the sharded two-phase entry points validate their arguments the same way
    raw = (C.c_char * 1024)()
    fb = C.c_void_p((C.addressof(raw) + 255) & ~255)               # 256-byte aligned like a workspace
    assert L.scl_knn_query_begin(None, fb, 4096, 64, fb, 8, 5, fb, fb, 1 << 20, None) == -1
    assert L.scl_knn_query_begin(fb, fb, 4096, 63, fb, 8, 5, fb, fb, 1 << 20, None) == -2          # D % 4
    assert L.scl_knn_query_begin(fb, fb, 512, 64, fb, 8, 5, fb, fb, 1 << 20, None) == -7           # shard too small: plain query
    assert L.scl_knn_query_begin(fb, fb, 4096, 64, fb, 8, 40, fb, fb, 1 << 20, None) == -7         # k > 32
    assert L.scl_knn_bound_reduce(None, 2, 8, 5, fb, None) == -1
    assert L.scl_knn_query_end(fb, fb, 4096, 64, fb, 8, 5, 0, None, fb, fb, None, fb, 1 << 20, None) == -1
    assert L.scl_knn_query_begin(fb, fb, 4096, 64, fb, 8, 5, fb, fb, 1 << 20, None) in (-3, -5, -6)   # alignment / no device


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "soft_contrastive_learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
    bench = open(os.path.join(ROOT, "bench.py")).read() if os.path.exists(os.path.join(ROOT, "bench.py")) else ""
    assert "torch.compile" not in bench


def test_host_logic_mirrors_reference():
    from oracle import losses as ol
    from oracle import retrieval as orr
    from soft_contrastive_learning_b200 import losses, retrieval
    assert np.array_equal(losses.ms_labels(3, 4, 5), ol.ms_labels(3, 4, 5))        # train.py:822-826
    rng = np.random.default_rng(0)
    xy = np.cumsum(rng.uniform(0, 2, size=(300, 2)), axis=0)
    for l in (0.0, 0.3, 1.0, 5.0):                                                 # top-n.py:34 L sweep
        assert retrieval.subsample_refs(xy, l) == orr.subsample_refs(xy, l)
    # contiguous shards cover [0,R) exactly once
    for R, G in ((1000, 8), (1001, 8), (7, 8), (1_000_000, 4)):
        spans = [retrieval.shard_bounds(R, G, r) for r in range(G)]
        assert spans[0][0] == 0 and spans[-1][1] == R
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(KeyError):
        losses.get_loss("residual_det")                                            # out of scope: loud, not silent
    for name in losses.LOSS_NAMES:
        assert callable(losses.get_loss(name))
    with pytest.raises(AttributeError):
        losses.distance_triplet_loss(None, None, None, 0.1, 0.5, None, 225.0, 2.0, "no_such_loss")


def test_synth_workloads_follow_the_sampler_contract():
    from soft_contrastive_learning_b200 import synth
    emb, dist, xy = synth.wms_batch(T=3, P=12, N=12, D=64, seed=42)
    assert emb.shape == (3, 25, 64) and dist.shape == (3, 25, 25) and emb.dtype == np.float32
    assert np.allclose(np.diagonal(dist, axis1=1, axis2=2), 0) and np.allclose(dist, dist.transpose(0, 2, 1))
    assert (dist[:, 0, 1:13] <= 15.0 + 1e-4).all()                                 # positives within MAX_POS_RADIUS
    assert (dist[:, 0, 13:] >= 15.0 - 1e-4).all()                                  # negatives beyond MIN_NEG_RADIUS
    iu = np.triu_indices(12, 1)
    assert (dist[:, 13:, 13:][:, iu[0], iu[1]] >= 15.0 - 1e-4).all()               # mutually exclusive negatives
