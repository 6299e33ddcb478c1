"""Tuple-mode losses over the GPUs of one box (SURVEY.md section 8e, row "Tuple-mode losses").

Tuples are independent units (the reference trains ``tuples_per_batch`` tuples per step, train/train.py:654, and its
loss is the mean over tuples), so the batch is split by tuples, one process per GPU: every rank runs the fused
forward+backward kernel on its own tuples and there is NO data-path collective -- gradients are per descriptor and stay
on the rank that owns the descriptor.  The only exchange is the scalar: the global mean needs the tuple counts and the
local loss sums (one all-reduce of two numbers).  With unequal tuple counts the local results are re-weighted by
``T_local / T_global`` so the outcome is identical to one call over the whole batch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def combine_tuple_shards(local_loss, local_grad, t_local, group=None):
    """(mean loss over ALL tuples of all ranks, gradient of that mean w.r.t. this rank's descriptors).

    ``local_loss`` is the mean over this rank's ``t_local`` tuples and ``local_grad`` the gradient of that local mean
    (what every ``*_value_and_grad`` / ``_wms_tuple_raw`` call returns).  Works on any backend (NCCL on the GPUs,
    gloo in the CPU tests)."""
    loss = local_loss if isinstance(local_loss, torch.Tensor) else torch.tensor(float(local_loss))
    pair = torch.stack((loss.detach().reshape(()).to(torch.float64) * float(t_local),
                        torch.tensor(float(t_local), dtype=torch.float64, device=loss.device)))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(pair, op=dist.ReduceOp.SUM, group=group)
    # everything stays on the device: no host synchronisation inside a training step
    grad = None if local_grad is None else local_grad * (float(t_local) / pair[1]).to(local_grad.dtype)
    return (pair[0] / pair[1]).to(loss.dtype), grad


def wms_loss_sharded(distances, embeddings, d_alpha, d_beta, alpha=2.0, beta=50.0, lamb=1.0, eps=0.1, ms_mining=True,
                     wfunction="exp", sumfunction="ms", group=None):
    """``wms_loss`` (model/losses.py:5-60) in tuple mode over this rank's tuples ``distances [T_local,S,S]``,
    ``embeddings [T_local,S,D]`` (device tensors): returns the global mean loss and d(global mean)/d(local embeddings)."""
    from . import losses
    params = losses._ms_params(d_alpha, d_beta, alpha, beta, lamb, eps, ms_mining, wfunction, sumfunction)
    emb = losses._f32(embeddings)
    loss, grad, _, _ = losses._wms_tuple_raw(emb, losses._f32(distances), params, need_grad=True)
    return combine_tuple_shards(loss.reshape(()), grad, emb.shape[0], group)


# ----------------------------------------------------------------------------------------------
# NetVLAD head + PCA projection, batch-parallel
# ----------------------------------------------------------------------------------------------
def _netvlad_local_fwd_bwd(x_local, assign_w, centers, dout_fn):
    """This rank's images through the CUDA head (csrc/netvlad*.cu): returns (vlad, dx, dW_local, dC_local)."""
    from . import netvlad
    xt = x_local.detach().requires_grad_(True)
    wt = assign_w.detach().requires_grad_(True)
    ct = centers.detach().requires_grad_(True)
    vlad = netvlad.netVLAD(xt, wt, ct)
    dout = dout_fn(vlad.detach())
    vlad.backward(dout)
    return vlad.detach(), xt.grad, wt.grad, ct.grad


def allreduce_netvlad_grads(dw_local, dc_local, group=None):
    """ONE all-reduce (sum) of the packed [dW | dC] buffer (2 x C x K floats = 256 KB for the reference's 512 x 64):
    the only parameters of the head are 'assignment/kernel' and 'cluster_centers' (model/nets.py:12), restored and trained
    as ordinary variables (train/train.py:874-892).  Returns (dW, dC) views of the reduced buffer."""
    packed = torch.stack((dw_local.reshape(-1), dc_local.reshape(-1)))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[0].reshape(dw_local.shape), packed[1].reshape(dc_local.shape)


def netvlad_step_sharded(x_local, assign_w, centers, dout_fn, group=None, local_fwd_bwd=None):
    """One data-parallel forward+backward of the NetVLAD head (model/nets.py:66-67 and what ``optimizer.minimize``
    differentiates, train/train.py:874-878): the images of the batch are split over the ranks (``x_local`` [B_local,
    HW, C] on this rank), the weights are replicated; every rank runs the head on its images and the only exchange is
    the all-reduce of ``[dW | dC]``.  ``dout_fn(vlad_local)`` returns d(global loss)/d(vlad_local) -- e.g. the PCA
    backward of the gradient a tuple-sharded loss produced (``V`` replicated, 537 MB at 32768 -> 4096); it must already
    carry the global-batch normalisation (``combine_tuple_shards`` does that for the tuple losses).
    Returns (vlad_local, dx_local, dW, dC) with dW, dC identical on every rank."""
    fn = local_fwd_bwd or _netvlad_local_fwd_bwd
    vlad, dx, dw, dc = fn(x_local, assign_w, centers, dout_fn)
    dw, dc = allreduce_netvlad_grads(dw, dc, group)
    return vlad, dx, dw, dc
