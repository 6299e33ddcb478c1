"""CPU, world_size 2 over gloo: the sharded-retrieval host protocol (contiguous row shards, global index offsets,
the all-gather of the per-rank score bounds between the two phases of the local query, ONE all-gather of the packed
per-shard [2,Q,k] messages, merge by (distance, index)) returns the exact global top-k.

The two device-side pieces (local kNN, merge kernel) are replaced by oracle stand-ins injected from here; what is
exercised is the product's distributed plumbing in soft_contrastive_learning_b200.retrieval."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleLocal:
    def __init__(self, X, index_offset=0):
        self.X, self.off = np.asarray(X), index_offset
        self.D = self.X.shape[1]
        self.db = torch.zeros(1)                 # the device the shard lives on (CPU in this test)

    # two-phase protocol (scl_knn_query_launch / _begin_group / _end_group) restated exactly: the bounds are the shard's k
    # smallest exact squared distances, the second phase returns the shard's rows at or below the reduced bound, padded
    # with (inf, -1); queries in two groups (ragged: 9 = 5 + 4)
    @staticmethod
    def query_groups(D, Q):
        gq = (Q + 1) // 2
        return (Q + gq - 1) // gq, gq

    def query_launch(self, q, k):
        from oracle import retrieval as orr
        if self.X.shape[0] < max(k, self.min_rows):           # "this shard does not take the tensor pass"
            return None
        self.kept = 0
        return orr.knn_bruteforce_exact(self.X, np.asarray(q), k)

    def _rows(self, state, group):
        if group < 0:                                          # all queries (unpipelined protocol)
            return slice(0, state[0].shape[0])
        gq = self.query_groups(self.D, state[0].shape[0])[1]
        return slice(group * gq, min(state[0].shape[0], (group + 1) * gq))

    def query_begin_group(self, state, q, k, group, ub):
        d = state[0][self._rows(state, group)]
        ub.copy_(torch.from_numpy((d ** 2).astype(np.float32) * (1 + 1e-6)))

    def query_end_group(self, state, q, k, group, bound, out):
        d, i = (a[self._rows(state, group)] for a in state)
        keep = d ** 2 <= bound.numpy()[:, None].astype(np.float64)
        self.kept += int(keep.sum())
        out[0].copy_(torch.from_numpy(np.where(keep, d, np.inf)))
        out[1].copy_(torch.from_numpy(np.where(keep, i + self.off, -1)))
        return out

    min_rows = 0

    def query_device(self, q, k=1, force_path=0, out=None):
        from oracle import retrieval as orr
        d, i = orr.knn_bruteforce_exact(self.X, np.asarray(q), k)
        if d.shape[1] < k:      # shard smaller than k: pad like the device path (inf, -1)
            pad = k - d.shape[1]
            d = np.concatenate([d, np.full((d.shape[0], pad), np.inf)], 1)
            i = np.concatenate([i, np.full((i.shape[0], pad), -1 - self.off, dtype=np.int64)], 1)
        d, i = torch.from_numpy(d), torch.from_numpy(i + self.off)
        if out is not None:     # the product hands over the two halves of its packed all-gather message
            out[0].copy_(d)
            out[1].copy_(i)
            return out
        return d, i


def _numpy_merge(d_all, i_all):
    G, Q, k = d_all.shape
    d = d_all.permute(1, 0, 2).reshape(Q, G * k).numpy()
    i = i_all.permute(1, 0, 2).reshape(Q, G * k).numpy()
    od = np.empty((Q, k))
    oi = np.empty((Q, k), dtype=np.int64)
    for q in range(Q):
        key_i = np.where(i[q] < 0, np.iinfo(np.int64).max, i[q])
        order = np.lexsort((key_i, d[q]))[:k]
        od[q], oi[q] = d[q][order], i[q][order]
    return torch.from_numpy(od), torch.from_numpy(oi)


def _worker(rank, world, port, R, D, Q, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from soft_contrastive_learning_b200 import retrieval, synth
        db, qry, *_ = synth.retrieval_problem(R=R, Q=Q, D=D, seed=5)
        lo, hi = retrieval.shard_bounds(R, world, rank)
        retrieval.KDTree = _OracleLocal            # test doubles for the two CUDA pieces
        retrieval.topk_merge = _numpy_merge
        retrieval.bound_reduce = lambda ub_all: torch.from_numpy(                  # [G,Q,k] -> k-th smallest of the union
            np.sort(ub_all.permute(1, 0, 2).reshape(ub_all.shape[1], -1).numpy(), axis=1)[:, ub_all.shape[2] - 1].copy())
        # the packed message [G, 2, Q, k] of 8-byte words, as ONE all-gather delivers it
        def merge_packed(packed, G, Q, k, out=None):
            d, i = _numpy_merge(packed[:, 0].contiguous().view(torch.float64), packed[:, 1].contiguous())
            if out is None:
                return d, i
            out[0].copy_(d)
            out[1].copy_(i)
            return out
        retrieval.topk_merge_packed = merge_packed
        tree = retrieval.ShardedKDTree(db[lo:hi], index_offset=lo, pipelined=True)     # per query group (two groups here)
        d, i = tree.query_device(qry, k)
        whole = retrieval.ShardedKDTree(db[lo:hi], index_offset=lo)                    # default: all queries at once
        dw, iw = whole.query_device(qry, k)
        assert torch.equal(iw, i) and torch.equal(dw, d)
        if hi - lo >= k:                           # the two-phase protocol ran and trimmed the per-rank lists
            kept = torch.tensor([tree.local.kept])
            dist.all_reduce(kept)
            assert Q * k <= int(kept) < 1.2 * Q * k
        # same answer from the single-phase protocol, and with one rank's shard refusing the first phase
        plain = retrieval.ShardedKDTree(db[lo:hi], index_offset=lo, two_phase=False)
        d1, i1 = plain.query_device(qry, k)
        assert torch.equal(i1, i) and torch.equal(d1, d)
        if rank == 1:
            tree.local.min_rows = 10 ** 9
        d1, i1 = tree.query_device(qry, k)
        assert torch.equal(i1, i) and torch.equal(d1, d)
        tree.local.min_rows = 0
        # queries in host memory: every rank copies its 1/G slice (ragged: Q = 9 over 2 ranks) and the slices are all-gathered
        d2, i2 = tree.query_from_host(torch.from_numpy(qry), k)
        assert torch.equal(i2, i) and torch.equal(d2, d)
        if rank == 0:
            np.savez(out, d=d.numpy(), i=i.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("R,k", [(257, 5), (40, 25)])
def test_sharded_protocol_world2(tmp_path, R, k):
    D, Q = 16, 9
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), R, D, Q, k, out), nprocs=2, join=True)
    from oracle import retrieval as orr
    from soft_contrastive_learning_b200 import synth
    db, qry, *_ = synth.retrieval_problem(R=R, Q=Q, D=D, seed=5)
    ref_d, ref_i = orr.knn_bruteforce_exact(db, qry, k)
    got = np.load(out)
    assert np.array_equal(got["i"], ref_i)
    assert np.allclose(got["d"], ref_d, rtol=1e-12)


# ---------------------------------------------------------------------------------------------
# tuple-mode losses shard by tuples; only the scalar mean is exchanged
# ---------------------------------------------------------------------------------------------
def _tuple_worker(rank, world, port, counts, out):
    import torch.distributed as dist
    from oracle import losses as ol
    from soft_contrastive_learning_b200 import sharded, synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        emb, dmat, _ = synth.wms_batch(T=sum(counts), P=4, N=5, D=32, seed=7)
        lo = sum(counts[:rank])
        e = torch.tensor(emb[lo:lo + counts[rank]], dtype=torch.float64, requires_grad=True)
        d = torch.tensor(dmat[lo:lo + counts[rank]], dtype=torch.float64)
        local = ol.wms_loss_tuples(d, e, 0.8, 15.0)                 # the oracle stands in for the CUDA kernel on CPU
        (g,) = torch.autograd.grad(local, e)
        loss, grad = sharded.combine_tuple_shards(local.detach(), g, counts[rank])
        torch.save({"loss": loss, "grad": grad}, out + f".{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(3, 3), (5, 2)])
def test_tuple_shards_reproduce_the_single_call(tmp_path, counts):
    """world_size 2 over gloo: the combined mean and the re-weighted local gradients equal one call over all tuples,
    also when the ranks hold different numbers of tuples."""
    import torch.multiprocessing as mp
    from oracle import losses as ol
    from soft_contrastive_learning_b200 import synth
    out = str(tmp_path / "tuple_shard")
    mp.spawn(_tuple_worker, args=(2, _free_port(), counts, out), nprocs=2, join=True)
    emb, dmat, _ = synth.wms_batch(T=sum(counts), P=4, N=5, D=32, seed=7)
    e = torch.tensor(emb, dtype=torch.float64, requires_grad=True)
    whole = ol.wms_loss_tuples(torch.tensor(dmat, dtype=torch.float64), e, 0.8, 15.0)
    (g,) = torch.autograd.grad(whole, e)
    parts = [torch.load(out + f".{r}") for r in range(2)]
    for p in parts:
        assert abs(float(p["loss"]) - float(whole)) <= 1e-12 * abs(float(whole))
    got = torch.cat([p["grad"] for p in parts])
    assert torch.allclose(got, g, rtol=1e-12, atol=1e-15)


# ---------------------------------------------------------------------------------------------
# NetVLAD head batch-parallel; the only exchange is one all-reduce of [dW | dC]
# ---------------------------------------------------------------------------------------------
def _oracle_netvlad_local(x_local, assign_w, centers, dout_fn):
    """float64 oracle stand-in for the CUDA head on CPU (same contract as sharded._netvlad_local_fwd_bwd)."""
    from oracle import netvlad as onv
    xt = x_local.detach().clone().requires_grad_(True)
    wt = assign_w.detach().clone().requires_grad_(True)
    ct = centers.detach().clone().requires_grad_(True)
    vlad = onv.netvlad_head(xt, wt, ct)
    vlad.backward(dout_fn(vlad.detach()))
    return vlad.detach(), xt.grad, wt.grad, ct.grad


def _netvlad_worker(rank, world, port, counts, out):
    import torch.distributed as dist
    from soft_contrastive_learning_b200 import sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        B = sum(counts)
        x = torch.randn((B, 12, 32), generator=g, dtype=torch.float64)
        w = 0.3 * torch.randn((32, 64), generator=g, dtype=torch.float64)
        c = 0.3 * torch.randn((32, 64), generator=g, dtype=torch.float64)
        dout = torch.randn((B, 32 * 64), generator=g, dtype=torch.float64) / B      # gradient of a global-batch mean
        lo = sum(counts[:rank])
        sl = slice(lo, lo + counts[rank])
        vlad, dx, dw, dc = sharded.netvlad_step_sharded(x[sl], w, c, lambda v: dout[sl], local_fwd_bwd=_oracle_netvlad_local)
        torch.save({"vlad": vlad, "dx": dx, "dw": dw, "dc": dc}, out + f".{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(4, 4), (5, 2)])
def test_netvlad_batch_parallel_reproduces_the_single_call(tmp_path, counts):
    """world_size 2 over gloo: per-rank images, replicated weights, ONE all-reduce of the packed [dW | dC]: both ranks end
    with the gradients of one call over the whole batch; outputs and dx stay with the rank that owns the images."""
    out = str(tmp_path / "nv_shard")
    mp.spawn(_netvlad_worker, args=(2, _free_port(), counts, out), nprocs=2, join=True)
    g = torch.Generator().manual_seed(3)
    B = sum(counts)
    x = torch.randn((B, 12, 32), generator=g, dtype=torch.float64)
    w = 0.3 * torch.randn((32, 64), generator=g, dtype=torch.float64)
    c = 0.3 * torch.randn((32, 64), generator=g, dtype=torch.float64)
    dout = torch.randn((B, 32 * 64), generator=g, dtype=torch.float64) / B
    vlad, dx, dw, dc = _oracle_netvlad_local(x, w, c, lambda v: dout)
    parts = [torch.load(out + f".{r}") for r in range(2)]
    for p in parts:
        assert torch.allclose(p["dw"], dw, rtol=1e-12, atol=1e-15) and torch.allclose(p["dc"], dc, rtol=1e-12, atol=1e-15)
    assert torch.allclose(torch.cat([p["vlad"] for p in parts]), vlad, rtol=1e-12, atol=1e-15)
    assert torch.allclose(torch.cat([p["dx"] for p in parts]), dx, rtol=1e-12, atol=1e-15)
