"""Exact brute-force top-N retrieval, geo bookkeeping and recall@N on B200, by the reference's call shapes.

Host-side mirror of /root/reference/evaluation/top-n.py:69-119 (and train/train.py:1181-1185, 363-386;
evaluation/roc.py:200-216).  ``KDTree`` below is a drop-in for the one call the reference makes on
``sklearn.neighbors.KDTree``:  ``KDTree(ref_f).query(query_f, k=N, return_distance=True, sort_results=True)``.
The neighbour search runs in libscl_b200.so (csrc/knn.cu, csrc/knn_tc.cu): an fp16 tcgen05 candidate pass, an exact
float64 rescore and a per-query exactness certificate with an exact-scan fallback -- results are the exact float64
kNN ordered by (distance, index).  With ``torch.distributed`` initialised, ``ShardedKDTree`` splits the database rows
over the ranks and merges per-shard lists after one NCCL all-gather.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np
import torch

from ._lib import check, lib
from .losses import _dev, _f32, _p, _stream, _ws


def _f64(x, dev):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    return x.to(dev, dtype=torch.float64).contiguous()


class KDTree:
    """Exact Euclidean kNN index over the rows of ``X`` (the database shard stays resident in HBM).

    Only the part of sklearn's KDTree API that the reference uses is provided: construction from an
    [R,D] float array and ``query``.  ``index_offset`` is added to returned indices (shards)."""

    def __init__(self, X, index_offset: int = 0):
        self.db = _f32(X)
        if self.db.dim() != 2:
            raise ValueError("X must be [n_samples, n_features]")
        self.R, self.D = self.db.shape
        if self.D % 4:
            raise ValueError("feature dimension must be a multiple of 4")
        self.index_offset = int(index_offset)
        L = lib()
        nbytes = C.c_size_t()
        check(L.scl_knn_shadow_bytes(self.R, self.D, C.byref(nbytes)), "scl_knn_shadow_bytes")
        self.shadow = _ws(nbytes.value, self.db.device)
        check(L.scl_knn_build(_p(self.db), self.R, self.D, _p(self.shadow), self.shadow.numel(), _stream()),
              "scl_knn_build")
        self.last_stats = None

    MAX_K = 1024

    def query_device(self, queries, k=1, force_path=0, out=None):
        """Device tensors in, device tensors out: (dist [Q,k] float64, idx [Q,k] int64).
        ``out`` = (dist, idx) pre-allocated contiguous device tensors (e.g. the two halves of one packed buffer)."""
        k = int(k)
        if k < 1:
            raise ValueError("k must be >= 1")
        if k > self.MAX_K:
            raise ValueError(f"k={k} exceeds the supported maximum of {self.MAX_K} neighbours per query "
                             "(scl_knn_query); split the request or raise kSelThreads/kTieCap in csrc/knn.cu")
        q = _f32(queries)
        if q.dim() == 1:
            q = q[None]
        Q = q.shape[0]
        if q.shape[1] != self.D:
            raise ValueError("query dimension mismatch")
        L = lib()
        nbytes = C.c_size_t()
        check(L.scl_knn_query_workspace_bytes(self.R, self.D, Q, k, C.byref(nbytes)), "scl_knn_query_workspace_bytes")
        ws = _ws(nbytes.value, q.device)
        if out is None:
            dist = torch.empty((Q, k), dtype=torch.float64, device=q.device)
            idx = torch.empty((Q, k), dtype=torch.int64, device=q.device)
        else:
            dist, idx = out
            assert dist.shape == (Q, k) and idx.shape == (Q, k) and dist.is_contiguous() and idx.is_contiguous()
        stats = torch.zeros(8, dtype=torch.int32, device=q.device)
        check(L.scl_knn_query(_p(self.db), _p(self.shadow), self.R, self.D, _p(q), Q, k, self.index_offset,
                              int(force_path), _p(dist), _p(idx), _p(stats), _p(ws), ws.numel(), _stream()),
              "scl_knn_query")
        self.last_stats = stats
        return dist, idx

    # -- sharded two-phase query (include/scl_b200.h: scl_knn_query_launch / _begin_group / _end_group) -----------------
    UNSUPPORTED = -7

    @staticmethod
    def query_groups(D, Q):
        """(n_groups, queries per group) of the first phase's tensor launch; the same on every rank."""
        n, gq = C.c_int(), C.c_int()
        check(lib().scl_knn_query_groups(int(D), int(Q), C.byref(n), C.byref(gq)), "scl_knn_query_groups")
        return n.value, gq.value

    def query_launch(self, q, k):
        """Query preparation + the tensor launch over all of ``q`` [Q,D] f32 on the current stream.  Returns the state to
        hand to ``query_begin_group`` / ``query_end_group``; None when this shard does not take the tensor pass."""
        Q = q.shape[0]
        L = lib()
        nbytes = C.c_size_t()
        check(L.scl_knn_query_workspace_bytes(self.R, self.D, Q, k, C.byref(nbytes)), "scl_knn_query_workspace_bytes")
        ws = _ws(nbytes.value, q.device)
        rc = L.scl_knn_query_launch(_p(self.db), _p(self.shadow), self.R, self.D, _p(q), Q, k, _p(ws), ws.numel(), _stream())
        if rc == self.UNSUPPORTED:
            return None
        check(rc, "scl_knn_query_launch")
        self._group_stats = []
        return ws

    def query_begin_group(self, state, q, k, group, ub):
        """On the CURRENT stream: wait for the tensor kernel's completion signal of ``group`` (-1: all queries), then
        fill ``ub`` [nq,k] f32 with the upper bounds on the distances of this shard's k best candidates."""
        check(lib().scl_knn_query_begin_group(_p(self.db), _p(self.shadow), self.R, self.D, _p(q), q.shape[0], k, int(group),
                                              _p(ub), _p(state), state.numel(), _stream()), "scl_knn_query_begin_group")

    def query_end_group(self, state, q, k, group, bound, out):
        """Second phase for the queries of ``group``: ``bound`` [nq] f32, ``out`` = (dist [nq,k] f64, idx [nq,k] i64)."""
        dist, idx = out
        stats = torch.zeros(8, dtype=torch.int32, device=q.device)
        check(lib().scl_knn_query_end_group(_p(self.db), _p(self.shadow), self.R, self.D, _p(q), q.shape[0], k,
                                            self.index_offset, int(group), _p(bound), _p(dist), _p(idx), _p(stats), _p(state),
                                            state.numel(), _stream()), "scl_knn_query_end_group")
        self._group_stats.append(stats)
        tot = torch.stack(self._group_stats).sum(0)
        tot[3], tot[6] = 2, len(self._group_stats)          # path; "chunks" = groups processed
        self.last_stats = tot
        return dist, idx

    def query_begin(self, q, k, bound):
        """Unpipelined first phase: launch + candidates of all queries; ``bound`` [Q,k] f32.  None (bounds = +inf) when this
        shard does not take the tensor pass."""
        state = self.query_launch(q, k)
        if state is None:
            bound.fill_(float("inf"))
            return None
        self.query_begin_group(state, q, k, -1, bound)
        return state

    def query_end(self, state, q, k, bound, out):
        return self.query_end_group(state, q, k, -1, bound, out)

    def query(self, X, k=1, return_distance=True, sort_results=True, force_path=0):
        """``KDTree.query`` as called at evaluation/top-n.py:106.  Results are always sorted ascending."""
        dist, idx = self.query_device(X, k, force_path)
        if isinstance(X, np.ndarray) or (isinstance(X, torch.Tensor) and not X.is_cuda):
            idx_h = idx.cpu().numpy()
            return (dist.cpu().numpy(), idx_h) if return_distance else idx_h
        return (dist, idx) if return_distance else idx

    def stats(self):
        """Counters of the last query (path 1 = exact scan, 2 = tensor pass).  ``n_fallback`` = queries the first tensor
        pass could not certify; of those ``n_stage2`` were resolved by the second tensor stage and ``n_scan`` went to
        the exact float64 scan; ``n_bound`` = queries of a two-phase (sharded) call settled by the reduced bound."""
        s = self.last_stats.cpu().tolist()
        return {"n_queries": s[0], "n_certified": s[1], "n_fallback": s[2], "path": s[3], "n_stage2": s[4],
                "n_scan": s[5], "chunks": s[6], "n_bound": s[7]}


def bound_reduce(ub_all):
    """[G,Q,k] gathered upper bounds (ascending per rank and query) -> [Q] k-th smallest of each query's union."""
    G, Q, k = ub_all.shape
    bound = torch.empty(Q, dtype=torch.float32, device=ub_all.device)
    check(lib().scl_knn_bound_reduce(_p(ub_all), G, Q, k, _p(bound), _stream()), "scl_knn_bound_reduce")
    return bound


def topk_merge(d_all, i_all):
    """Merge G sorted per-shard lists: d_all [G,Q,k] float64, i_all [G,Q,k] int64 -> ([Q,k], [Q,k])."""
    G, Q, k = d_all.shape
    d = torch.empty((Q, k), dtype=torch.float64, device=d_all.device)
    i = torch.empty((Q, k), dtype=torch.int64, device=d_all.device)
    check(lib().scl_topk_merge(_p(d_all.contiguous()), _p(i_all.contiguous()), G, Q, k, 0, _p(d), _p(i), _stream()),
          "scl_topk_merge")
    return d, i


def topk_merge_packed(packed, G, Q, k, out=None):
    """Merge G packed per-rank messages ``packed`` [G, 2, Q, k] (8-byte words: [g,0] = float64 distances, [g,1] = int64
    indices) as one all-gather delivers them, without a repack.  ``out`` = (d, i) contiguous [Q,k] destinations."""
    assert packed.is_contiguous() and packed.element_size() == 8 and packed.numel() == G * 2 * Q * k
    if out is None:
        d = torch.empty((Q, k), dtype=torch.float64, device=packed.device)
        i = torch.empty((Q, k), dtype=torch.int64, device=packed.device)
    else:
        d, i = out
        assert d.shape == (Q, k) and i.shape == (Q, k) and d.is_contiguous() and i.is_contiguous()
    base = packed.data_ptr()
    check(lib().scl_topk_merge(C.c_void_p(base), C.c_void_p(base + Q * k * 8), G, Q, k, 2 * Q * k, _p(d), _p(i), _stream()),
          "scl_topk_merge")
    return d, i


class ShardedKDTree:
    """Database rows split contiguously over the ranks of ``group``; queries replicated.

    Each rank answers against its shard (global indices = local + offset); one all-gather of the [Q,k] lists
    (NCCL over NVLink) followed by the merge kernel gives every rank the exact global top-k.  With ``two_phase`` (default)
    the ranks first exchange their candidates' score bounds so that each rescores only what can reach the GLOBAL top-k."""

    TWO_PHASE_MAX_K = 32          # the tensor pass serves k <= 32 (scl_knn_query)

    def __init__(self, X_local, index_offset, group=None, two_phase=True, pipelined=False):
        import torch.distributed as dist
        self.group = group
        self.two_phase = bool(two_phase)
        # two-phase: run the second phase per query group on a second stream under the tensor kernel.  Off by default:
        # kernels that co-run with the persistent tensor kernel crawl (and slow it by 1-3 %), measured gain +0.7 % at
        # N = 2 and none at N = 8
        self.pipelined = bool(pipelined)
        self._side = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local = KDTree(X_local, index_offset=index_offset)

    def query_device(self, queries, k=1, force_path=0):
        import torch.distributed as dist
        if self.world == 1:
            return self.local.query_device(queries, k, force_path)
        q = queries if isinstance(queries, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(queries))
        q = q.to(self.local.db.device, dtype=torch.float32).contiguous()      # the device the shard lives on
        if q.dim() == 1:
            q = q[None]
        Q = q.shape[0]
        if not (self.two_phase and force_path == 0 and k <= self.TWO_PHASE_MAX_K):
            # one packed message per rank: [2, Q, k] 8-byte words (float64 distances | int64 indices), ONE all-gather
            mine = torch.empty((2, Q, k), dtype=torch.int64, device=q.device)
            self.local.query_device(q, k, force_path, out=(mine[0].view(torch.float64), mine[1]))
            packed = torch.empty((self.world, 2, Q, k), dtype=torch.int64, device=q.device)
            dist.all_gather_into_tensor(packed.view(self.world * 2 * Q, k), mine.view(2 * Q, k), group=self.group)
            return topk_merge_packed(packed, self.world, Q, k)
        return self._query_two_phase(q, k)

    def _query_two_phase(self, q, k):
        """Two phases, pipelined over the query groups of ONE tensor launch (include/scl_b200.h: scl_knn_query_launch).

        Phase 1 on the current stream: the tensor kernel works through the queries group by group and signals each group
        in device memory.  On a second stream, per group: wait for the signal, select the candidates, all-gather the ranks'
        [nq,k] score bounds, take the k-th smallest of the union as the bound on the global k-th distance, rescore only what
        can still make the GLOBAL top-k, all-gather the packed lists, merge -- all of it under the tensor kernel's work on
        the next group.  The branch condition and the groups depend only on values every rank shares (k, Q, D), so the
        collectives match; a rank whose shard is too small for the tensor pass contributes +inf and answers with the
        plain query."""
        import torch.distributed as dist
        G, rank = self.world, dist.get_rank(self.group)
        Q, dev = q.shape[0], q.device
        n_groups, gq = KDTree.query_groups(self.local.D, Q) if self.pipelined else (1, Q)
        on_gpu = dev.type == "cuda" and n_groups > 1          # (one group, or the CPU protocol tests: no second stream)
        if on_gpu:
            main = torch.cuda.current_stream(dev)
            if self._side is None:
                self._side = torch.cuda.Stream(dev)
            side = self._side
        # every buffer is allocated on the main stream (its allocator pool) and main waits for the side stream at the end
        d = torch.empty((Q, k), dtype=torch.float64, device=dev)
        i = torch.empty((Q, k), dtype=torch.int64, device=dev)
        bufs = []
        for g in range(n_groups):
            nq = min(gq, Q - g * gq)
            bufs.append((torch.empty((G, nq, k), dtype=torch.float32, device=dev),          # gathered score bounds
                         torch.empty((2, nq, k), dtype=torch.int64, device=dev),            # this rank's packed lists
                         torch.empty((G, 2, nq, k), dtype=torch.int64, device=dev)))        # gathered packed lists
        state = self.local.query_launch(q, k)
        if on_gpu and state is None:
            side.wait_stream(main)                            # (tensor path: begin_group waits for the launch's own event)
        with (torch.cuda.stream(side) if on_gpu else contextlib.nullcontext()):
            for g in range(n_groups):
                q0, nq = g * gq, min(gq, Q - g * gq)
                ub_all, mine, packed = bufs[g]
                ub = ub_all[rank]
                gid = g if n_groups > 1 else -1
                if state is None:
                    ub.fill_(float("inf"))
                else:
                    self.local.query_begin_group(state, q, k, gid, ub)
                dist.all_gather_into_tensor(ub_all.view(G * nq, k), ub, group=self.group)
                bound = bound_reduce(ub_all)
                out = (mine[0].view(torch.float64), mine[1])
                if state is None:
                    self.local.query_device(q[q0:q0 + nq], k, 0, out=out)
                else:
                    self.local.query_end_group(state, q, k, gid, bound, out)
                dist.all_gather_into_tensor(packed.view(G * 2 * nq, k), mine.view(2 * nq, k), group=self.group)
                topk_merge_packed(packed, G, nq, k, out=(d[q0:q0 + nq], i[q0:q0 + nq]))
        if on_gpu:
            main.wait_stream(side)
        return d, i

    def query_from_host(self, q_host, k=1):
        """Queries that live in (pinned) host memory, identical on every rank: each rank copies only its 1/G slice over
        PCIe and the slices are all-gathered over NVLink, instead of G full host->device copies competing for the host
        links.  Returns device tensors like ``query_device``."""
        import torch.distributed as dist
        if self.world == 1:
            return self.local.query_device(q_host.to(self.local.db.device, non_blocking=True), k)
        rank = dist.get_rank(self.group)
        Q, D = q_host.shape
        per = (Q + self.world - 1) // self.world
        dev = self.local.db.device
        mine = torch.zeros((per, D), dtype=torch.float32, device=dev)
        lo, hi = min(Q, rank * per), min(Q, (rank + 1) * per)
        if hi > lo:
            mine[:hi - lo].copy_(q_host[lo:hi], non_blocking=True)
        full = torch.empty((self.world * per, D), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        return self.query_device(full[:Q], k)

    def query(self, X, k=1, return_distance=True, sort_results=True):
        d, i = self.query_device(X, k)
        if isinstance(X, np.ndarray):
            return (d.cpu().numpy(), i.cpu().numpy()) if return_distance else i.cpu().numpy()
        return (d, i) if return_distance else i


def shard_bounds(R, world, rank):
    """Contiguous row split R/G per rank."""
    per = (R + world - 1) // world
    lo = min(R, rank * per)
    return lo, min(R, lo + per)


# ----------------------------------------------------------------------------------------------
# R2 / R3
# ----------------------------------------------------------------------------------------------
def geo_topn(query_xy, ref_xy, top_i):
    """top-n.py:69,110-113 without the [Q,R] distance matrix: (top_g_dists [Q,k], gt_i [Q], gt_g_dist [Q])."""
    dev = _dev()
    qxy, rxy = _f64(query_xy, dev), _f64(ref_xy, dev)
    ti = top_i if isinstance(top_i, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(top_i, dtype=np.int64))
    ti = ti.to(dev, dtype=torch.int64).contiguous()
    Q, k = ti.shape
    R = rxy.shape[0]
    tg = torch.empty((Q, k), dtype=torch.float64, device=dev)
    gi = torch.empty(Q, dtype=torch.int64, device=dev)
    gd = torch.empty(Q, dtype=torch.float64, device=dev)
    check(lib().scl_geo_topn(_p(qxy), _p(rxy), Q, R, _p(ti), k, _p(tg), _p(gi), _p(gd), _stream()), "scl_geo_topn")
    return tg, gi, gd


def subsample_refs(ref_xy, l):
    """top-n.py:91-94 greedy subsampling (sequential by construction; host side like the reference)."""
    ref_xy = np.asarray(ref_xy)
    ref_idx = [0]
    last = ref_xy[0]
    l2 = l ** 2
    for i in range(len(ref_xy)):
        d = ref_xy[i] - last
        if d[0] * d[0] + d[1] * d[1] >= l2:
            ref_idx.append(i)
            last = ref_xy[i]
    return ref_idx


def top_n(ref_f, query_f, ref_xy, query_xy, N=25, l=0.0):
    """Body of get_top_n for one (dimension, l) cell (top-n.py:91-119) on already PCA-projected features.
    Returns the pickle payload [top_i, top_g_dists, top_f_dists, gt_i, gt_g_dist, ref_idx] (NumPy / lists)."""
    ref_idx = subsample_refs(ref_xy, l)                                        # :91-94
    if len(ref_idx) < N:                                                       # :96-97
        return None
    ref_idx_np = np.asarray(ref_idx, dtype=np.int64)
    ref_f = np.asarray(ref_f)
    sub_f = ref_f[ref_idx_np]                                                  # :99
    tree = KDTree(sub_f)                                                       # :103
    top_f_dists, top_i = tree.query(np.asarray(query_f), k=N, return_distance=True, sort_results=True)   # :106
    sub_xy = np.asarray(ref_xy, dtype=np.float64)[ref_idx_np]
    tg, gi, gd = geo_topn(np.asarray(query_xy, dtype=np.float64), sub_xy, top_i)                          # :110-113
    top_i_orig = ref_idx_np[top_i]                                             # :116
    gt_i = ref_idx_np[gi.cpu().numpy()]                                        # :117
    return [top_i_orig.tolist(), tg.cpu().numpy().tolist(), top_f_dists, gt_i.tolist(), gd.cpu().numpy(), ref_idx]


def recall_curves(top_g_dists, thresholds):
    """curves[n, x] = % of queries with min_{j<=n} top_g_dists[q,j] < thresholds[x]
    (train.py:368-375; row 0 is roc.py:213-216's top-1 curve)."""
    dev = _dev()
    tg = _f64(np.asarray(top_g_dists, dtype=np.float64) if not isinstance(top_g_dists, torch.Tensor) else top_g_dists, dev)
    th = _f64(np.asarray(thresholds, dtype=np.float64), dev)
    Q, k = tg.shape
    out = torch.empty((k, th.numel()), dtype=torch.float64, device=dev)
    check(lib().scl_recall_curves(_p(tg), Q, k, _p(th), th.numel(), _p(out), _stream()), "scl_recall_curves")
    return out.cpu().numpy()


def recall_at_n(top_g_dists, rad=25.0, num=25):
    """train.py:373-375: X = linspace(0, rad, 25); returns (X, Y[n, x])."""
    X = np.linspace(0, rad, num=num)
    return X, recall_curves(top_g_dists, X)
