"""NumPy stand-in for the ~40 TF-1.x ops that /root/reference/model/losses.py uses.

TEST TOOLING.  TensorFlow 1.10 cannot be installed here (Python 3.12, no wheel, no network), so the
reference's loss file cannot run on TF.  This shim lets the reference's OWN SOURCE execute on NumPy
float64 arrays: ``load_reference_losses()`` imports ``/root/reference/model/losses.py`` unmodified
with ``tensorflow`` -> this module and ``pointnetvlad_cls`` -> a restatement of the four external
PointNetVLAD losses.  Only the forward value is produced (no autodiff); gradients are pinned by
central finite differences of these forwards (tests/golden/make_golden.py).

Each op below is the plain NumPy equivalent of the TF op of the same name (same broadcasting, same
axis/keepdims semantics).  Nothing here is imported by the product.
"""
from __future__ import annotations

import contextlib
import importlib.util
import sys
import types

import numpy as np


class _Dim(int):
    pass


class _Shape(tuple):
    def as_list(self):
        return [int(d) for d in self]


class T(np.ndarray):
    """ndarray with the slice of the tf.Tensor API that losses.py touches."""

    def get_shape(self):
        return _Shape(_Dim(d) for d in self.shape)


def _w(x):
    return np.asarray(x).view(T)


def _make_tf():
    tf = types.ModuleType("tensorflow")
    # dtypes are honoured: what the reference declares float32 (the GPS masks, losses.py:22) runs in float32 --
    # tf.exp overflowing to inf there is part of the loss's behaviour -- while descriptors fed as float64 keep the
    # similarity / log-sum-exp arithmetic in float64 (NumPy promotes mixed operands)
    tf.float32 = np.float32
    tf.float64 = np.float64
    tf.bool = np.bool_

    def _axis(axis):
        return tuple(axis) if isinstance(axis, (list, tuple)) else axis

    tf.constant = lambda v, dtype=None: _w(np.array(v, dtype=np.float64))
    tf.cast = lambda x, dtype=None: _w(np.asarray(x).astype(dtype if dtype is not None else np.float64))
    tf.eye = lambda n, dtype=None: _w(np.eye(int(n), dtype=dtype if dtype is not None else np.float32))
    tf.zeros = lambda shape, dtype=None: _w(np.zeros([int(s) for s in shape]))
    tf.ones_like = lambda x: _w(np.ones_like(x))
    tf.zeros_like = lambda x: _w(np.zeros_like(x))
    tf.fill = lambda shape, v: _w(np.full([int(s) for s in shape], v, dtype=np.float64))
    tf.where = lambda c, a, b: _w(np.where(c, a, b))
    tf.divide = lambda a, b: _w(np.divide(a, b))
    tf.div = tf.divide
    tf.truediv = tf.divide
    tf.multiply = lambda a, b: _w(np.multiply(a, b))
    tf.add = lambda a, b: _w(np.add(a, b))
    tf.subtract = lambda a, b: _w(np.subtract(a, b))
    tf.exp = lambda x: _w(np.exp(x))
    tf.log = lambda x: _w(np.log(x))
    # correctly rounded in the operand's dtype (see oracle/losses.py wms_masks)
    tf.tanh = lambda x: _w(np.tanh(np.asarray(x, dtype=np.float64)).astype(np.asarray(x).dtype))
    tf.sqrt = lambda x: _w(np.sqrt(x))
    tf.maximum = lambda a, b: _w(np.maximum(a, b))
    tf.minimum = lambda a, b: _w(np.minimum(a, b))
    tf.equal = lambda a, b: _w(np.equal(a, b))
    tf.logical_not = lambda a: _w(np.logical_not(a))
    tf.squared_difference = lambda a, b: _w((np.asarray(a) - np.asarray(b)) ** 2)
    tf.reshape = lambda x, shape: _w(np.reshape(x, [int(s) for s in shape]))
    tf.tile = lambda x, m: _w(np.tile(x, [int(s) for s in m]))
    tf.concat = lambda xs, axis: _w(np.concatenate(xs, axis=axis))

    def transpose(x, perm=None):
        return _w(np.transpose(x, perm))
    tf.transpose = transpose

    def matmul(a, b, transpose_a=False, transpose_b=False, adjoint_b=False):
        a = np.asarray(a)
        b = np.asarray(b)
        if transpose_a:
            a = np.swapaxes(a, -1, -2)
        if transpose_b or adjoint_b:
            b = np.swapaxes(b, -1, -2)
        return _w(a @ b)
    tf.matmul = matmul
    tf.einsum = lambda eq, *ops: _w(np.einsum(eq, *ops))

    def _red(fn):
        def f(x, axis=None, keepdims=False, keep_dims=False):
            return _w(fn(np.asarray(x), axis=_axis(axis), keepdims=bool(keepdims or keep_dims)))
        return f
    tf.reduce_sum = _red(np.sum)
    tf.reduce_mean = _red(np.mean)
    tf.reduce_max = _red(np.max)
    tf.reduce_min = _red(np.min)

    nn = types.ModuleType("tensorflow.nn")

    def l2_normalize(x, axis=None, epsilon=1e-12, dim=None):
        ax = axis if axis is not None else dim
        x = np.asarray(x)
        ss = np.sum(x * x, axis=ax, keepdims=True)
        return _w(x / np.sqrt(np.maximum(ss, epsilon)))
    nn.l2_normalize = l2_normalize
    nn.relu = lambda x: _w(np.maximum(x, 0))
    tf.nn = nn

    losses = types.ModuleType("tensorflow.losses")

    class Reduction:
        NONE = "none"
        SUM_BY_NONZERO_WEIGHTS = "weighted_sum_by_nonzero_weights"
    losses.Reduction = Reduction

    def huber_loss(labels, predictions, weights=1.0, delta=1.0, reduction=Reduction.SUM_BY_NONZERO_WEIGHTS):
        err = np.asarray(predictions) - np.asarray(labels)
        abs_err = np.abs(err)
        quad = np.minimum(abs_err, delta)
        lin = abs_err - quad
        l = 0.5 * quad ** 2 + delta * lin
        if reduction == Reduction.NONE:
            return _w(l)
        return _w(np.mean(l))
    losses.huber_loss = huber_loss
    tf.losses = losses

    linalg = types.ModuleType("tensorflow.linalg")
    tf.linalg = linalg

    @contextlib.contextmanager
    def name_scope(name):
        yield name
    tf.name_scope = name_scope
    return tf


def _make_pointnetvlad(tf):
    """Restatement of mikacuy/pointnetvlad loss/pointnetvlad_loss.py (README.md:11, un-pinned), written
    with the shim's tf ops in the same op order as the in-repo twin evil_triplet_loss (losses.py:63-73)."""
    m = types.ModuleType("pointnetvlad_cls")

    def best_pos_distance(query, pos_vecs):
        num_pos = pos_vecs.get_shape()[1]
        query_copies = tf.tile(query, [1, int(num_pos), 1])
        return tf.reduce_min(tf.reduce_sum(tf.squared_difference(pos_vecs, query_copies), 2), 1)

    def _second(ref, a, neg_vecs, margin, inner):
        num_neg = neg_vecs.get_shape()[1]
        batch = a.get_shape()[0]
        copies = tf.tile(a, [1, int(num_neg), 1])
        ref = tf.tile(tf.reshape(ref, (-1, 1)), [1, int(num_neg)])
        mm = tf.fill([int(batch), int(num_neg)], margin)
        return tf.reduce_mean(inner(tf.maximum(
            tf.add(mm, tf.subtract(ref, tf.reduce_sum(tf.squared_difference(neg_vecs, copies), 2))),
            tf.zeros([int(batch), int(num_neg)])), 1))

    def triplet_loss(q_vec, pos_vecs, neg_vecs, margin):
        return _second(best_pos_distance(q_vec, pos_vecs), q_vec, neg_vecs, margin, tf.reduce_sum)

    def lazy_triplet_loss(q_vec, pos_vecs, neg_vecs, margin):
        return _second(best_pos_distance(q_vec, pos_vecs), q_vec, neg_vecs, margin, tf.reduce_max)

    def quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2):
        trip = triplet_loss(q_vec, pos_vecs, neg_vecs, m1)
        return trip + _second(best_pos_distance(q_vec, pos_vecs), other_neg, neg_vecs, m2, tf.reduce_sum)

    def lazy_quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2):
        trip = lazy_triplet_loss(q_vec, pos_vecs, neg_vecs, m1)
        return trip + _second(best_pos_distance(q_vec, pos_vecs), other_neg, neg_vecs, m2, tf.reduce_max)

    m.best_pos_distance = best_pos_distance
    m.triplet_loss = triplet_loss
    m.lazy_triplet_loss = lazy_triplet_loss
    m.quadruplet_loss = quadruplet_loss
    m.lazy_quadruplet_loss = lazy_quadruplet_loss
    return m


def load_reference_losses(path="/root/reference/model/losses.py"):
    """Import the reference's losses.py, unmodified, over the shim.  Returns (module, pointnetvlad_shim)."""
    tf = _make_tf()
    pn = _make_pointnetvlad(tf)
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "pointnetvlad_cls")}
    sys.modules["tensorflow"] = tf
    sys.modules["pointnetvlad_cls"] = pn
    try:
        spec = importlib.util.spec_from_file_location("_reference_losses", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod, pn


def A(x, dtype=np.float64):
    """Wrap an array so reference code can call .get_shape() on it (descriptors float64, GPS distances float32)."""
    return _w(np.asarray(x, dtype=dtype))
