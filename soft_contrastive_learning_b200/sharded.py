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
