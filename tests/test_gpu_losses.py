"""GPU parity: CUDA losses (through the C ABI) vs the float64 oracle and the committed golden vectors.

Tolerance (BASELINE.json north_star): loss and gradients within 1e-5 relative of the reference math in fp32.
Gradients are compared relative to the gradient's max magnitude (element-wise relative error is meaningless for
entries that cancel to ~0); thresholds: loss 1e-5, gradient 1e-5, masks identical.  The error each test MEASURES is
recorded (conftest `measured`) and written to gpurun_out/parity_measured.json at the end of the session."""
import numpy as np
import pytest
import torch

from oracle import losses as ol
from soft_contrastive_learning_b200 import synth

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-5
GRAD_TOL = 1e-5


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def grad_err(g, ref):
    return np.abs(np.asarray(g, dtype=np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30)


WMS_VARIANTS = {
    "exp_ms_mine": dict(wfunction="exp", sumfunction="ms", ms_mining=True),
    "exp_ms_nomine": dict(wfunction="exp", sumfunction="ms", ms_mining=False),
    "lin_ms_mine": dict(wfunction="lin", sumfunction="ms", ms_mining=True),
    "tanh_ms_mine": dict(wfunction="tanh", sumfunction="ms", ms_mining=True),
    "exp_plain_mine": dict(wfunction="exp", sumfunction="plain", ms_mining=True),
}


def _kept_bits(kept, S):
    k = kept.cpu().numpy().astype(np.uint32)        # [T,S,2]
    bits = ((k[..., None] >> np.arange(S, dtype=np.uint32)) & 1).astype(bool)      # [T,S,2,S]
    return bits[:, :, 0, :], bits[:, :, 1, :]


@pytest.mark.parametrize("tag", list(WMS_VARIANTS))
def test_wms_golden_tuple_kernel(cuda_lib, golden, tag):
    from soft_contrastive_learning_b200 import losses
    g = golden("wms_flat_S25_D64")
    loss, grad, kept = losses.wms_loss_value_and_grad(g["dist"], g["emb"], 0.8, 15.0, return_kept=True,
                                                       **WMS_VARIANTS[tag])
    assert rel(loss, float(g["loss_" + tag])) < LOSS_TOL
    assert grad_err(grad, g["grad_" + tag]) < GRAD_TOL
    kp, kn = _kept_bits(kept, 25)
    assert np.array_equal(kp[0], g["keptpos_" + tag]) and np.array_equal(kn[0], g["keptneg_" + tag])


def test_wms_golden_tuples_T4(cuda_lib, golden):
    from soft_contrastive_learning_b200 import losses
    g = golden("wms_tuples_T4_S25_D256")
    loss, grad = losses.wms_loss_value_and_grad(g["dist"], g["emb"], 0.8, 15.0)
    assert rel(loss, float(g["loss"])) < LOSS_TOL
    assert grad_err(grad, g["grad"]) < GRAD_TOL


def _oracle_wms(emb, dist, **kw):
    return ol.value_and_grad(lambda e: ol.wms_loss_tuples(torch.as_tensor(dist.astype(np.float64)), e, 0.8, 15.0, **kw),
                             [emb.astype(np.float64)])


@pytest.mark.parametrize("T,P,N,D", [(32, 12, 12, 4096),      # BASELINE config 1
                                     (3, 15, 16, 1024),       # S = 32 (config 3 tuple shape)
                                     (2, 12, 12, 32768),      # the published model's 32768-d descriptors: chunked path
                                     (5, 3, 4, 64), (2, 14, 14, 512)])
def test_wms_tuple_vs_oracle(cuda_lib, T, P, N, D):
    from soft_contrastive_learning_b200 import losses
    emb, dist, _ = synth.wms_batch(T=T, P=P, N=N, D=D, seed=42)
    loss, grad, kept = losses.wms_loss_value_and_grad(dist, emb, 0.8, 15.0, return_kept=True)
    ref, (rg,) = _oracle_wms(emb, dist)
    assert rel(loss, ref) < LOSS_TOL, (loss, ref)
    assert grad_err(grad, rg) < GRAD_TOL
    # identical mining decisions, pair by pair
    S = 1 + P + N
    kp, kn = _kept_bits(kept, S)
    for t in range(T):
        _, mp, mn = ol.wms_loss(dist[t].astype(np.float64), emb[t].astype(np.float64), 0.8, 15.0, return_masks=True)
        assert np.array_equal(kp[t], mp.numpy()) and np.array_equal(kn[t], mn.numpy())


@pytest.mark.parametrize("cluster", ["1", "2", "4", "8"])
def test_wms_cluster_sizes_agree(cuda_lib, cluster, tune):
    from soft_contrastive_learning_b200 import losses
    tune("SCL_WMS_CLUSTER", int(cluster))
    emb, dist, _ = synth.wms_batch(T=6, P=12, N=12, D=2048, seed=3)
    loss, grad = losses.wms_loss_value_and_grad(dist, emb, 0.8, 15.0)
    ref, (rg,) = _oracle_wms(emb, dist)
    assert rel(loss, ref) < LOSS_TOL and grad_err(grad, rg) < GRAD_TOL


# the tuple-mode kernels: streaming with the FFMA2 backward and with the tensor-core backward (large batches),
# cluster-resident (small batches), cluster-chunked (slice too large for shared memory); each is forced through the same shapes, incl. ragged D (not a multiple of the 256-column
# ring stage), odd S (distance block not 16-byte sized), S = 32 (widest register tile) and a forward-only call
WMS_PATHS = {"stream": {"SCL_WMS_STREAM": 1, "SCL_WMS_STREAM_CFG": 2}, "resident": {"SCL_WMS_STREAM": 0},
             "stream_mma": {"SCL_WMS_STREAM": 1, "SCL_WMS_STREAM_CFG": 6},
             "chunked": {"SCL_WMS_STREAM": 0, "SCL_WMS_CHUNKED": 1}}


@pytest.mark.parametrize("path", list(WMS_PATHS))
@pytest.mark.parametrize("T,P,N,D", [(7, 12, 12, 4096), (3, 15, 16, 1024), (5, 3, 4, 64), (4, 12, 12, 1000),
                                     (3, 13, 14, 772), (2, 12, 12, 256), (1, 12, 12, 8192)])
def test_wms_kernel_paths_vs_oracle(cuda_lib, tune, path, T, P, N, D):
    from soft_contrastive_learning_b200 import losses
    for k, v in WMS_PATHS[path].items():
        tune(k, v)
    emb, dist, _ = synth.wms_batch(T=T, P=P, N=N, D=D, seed=11)
    loss, grad, kept = losses.wms_loss_value_and_grad(dist, emb, 0.8, 15.0, return_kept=True)
    ref, (rg,) = _oracle_wms(emb, dist)
    assert rel(loss, ref) < LOSS_TOL, (loss, ref)
    assert grad_err(grad, rg) < GRAD_TOL
    S = 1 + P + N
    kp, kn = _kept_bits(kept, S)
    for t in range(T):
        _, mp, mn = ol.wms_loss(dist[t].astype(np.float64), emb[t].astype(np.float64), 0.8, 15.0, return_masks=True)
        assert np.array_equal(kp[t], mp.numpy()) and np.array_equal(kn[t], mn.numpy())


@pytest.mark.parametrize("path", list(WMS_PATHS))
def test_wms_kernel_paths_variants_and_forward_only(cuda_lib, tune, golden, path):
    from soft_contrastive_learning_b200 import losses
    for k, v in WMS_PATHS[path].items():
        tune(k, v)
    g = golden("wms_flat_S25_D64")
    for tag, kw in WMS_VARIANTS.items():
        loss, grad = losses.wms_loss_value_and_grad(g["dist"], g["emb"], 0.8, 15.0, **kw)
        assert rel(loss, float(g["loss_" + tag])) < LOSS_TOL
        assert grad_err(grad, g["grad_" + tag]) < GRAD_TOL
    e = torch.tensor(g["emb"], device="cuda")
    d = torch.tensor(g["dist"], device="cuda")
    with torch.no_grad():
        l = losses.wms_loss(d, e, 0.8, 15.0)               # forward only: demb == NULL in the C call
    assert rel(float(l), float(g["loss_exp_ms_mine"])) < LOSS_TOL


@pytest.mark.parametrize("cfg", ["2", "6"])
def test_wms_stream_large_batch_matches_small_batch_kernels(cuda_lib, tune, cfg):
    """T = 600 tuples (several per persistent CTA): the streaming kernels (FFMA2 backward, tensor-core backward)
    against the cluster kernel, per tuple."""
    tune("SCL_WMS_STREAM_CFG", int(cfg))
    from soft_contrastive_learning_b200 import losses
    emb, dist, _ = synth.wms_batch(T=40, P=12, N=12, D=1024, seed=21)
    emb = np.tile(emb, (15, 1, 1)) + 1e-3 * np.random.default_rng(0).standard_normal((600, 25, 1024)).astype(np.float32)
    dist = np.tile(dist, (15, 1, 1))
    p = losses._ms_params(0.8, 15.0)
    e, d = torch.tensor(emb, device="cuda"), torch.tensor(dist, device="cuda")
    tune("SCL_WMS_STREAM", 1)
    l1, g1, k1, t1 = losses._wms_tuple_raw(e, d, p, need_grad=True, want_kept=True, want_per_tuple=True)
    tune("SCL_WMS_STREAM", 0)
    l0, g0, k0, t0 = losses._wms_tuple_raw(e, d, p, need_grad=True, want_kept=True, want_per_tuple=True)
    assert torch.equal(k1, k0)
    assert torch.allclose(t1, t0, rtol=2e-6, atol=0)
    assert rel(float(l1), float(l0)) < 2e-6
    assert float((g1 - g0).abs().max() / g0.abs().max()) < GRAD_TOL


def test_wms_autograd_wrapper_and_2d_call(cuda_lib):
    from soft_contrastive_learning_b200 import losses
    emb, dist, _ = synth.wms_batch(T=1, P=12, N=12, D=256, seed=9)
    e = torch.tensor(emb[0], device="cuda", requires_grad=True)
    d3 = torch.tensor(dist, device="cuda")               # [1,S,S], as train.py:684-686 feeds it
    l3 = losses.wms_loss(d3, e, d_alpha=0.8, d_beta=15.0)
    (3.0 * l3).backward()
    ref, (rg,) = ol.value_and_grad(lambda x: ol.wms_loss(torch.as_tensor(dist[0].astype(np.float64)), x, 0.8, 15.0),
                                   [emb[0].astype(np.float64)])
    assert rel(float(l3), ref) < LOSS_TOL
    assert grad_err(e.grad.cpu().numpy() / 3.0, rg) < GRAD_TOL
    l2 = losses.wms_loss(d3[0], e.detach(), 0.8, 15.0)   # the 2-D form of losses.py:5
    assert float(l2) == float(l3)


def test_wms_is_deterministic(cuda_lib):
    from soft_contrastive_learning_b200 import losses
    emb, dist, _ = synth.wms_batch(T=16, P=12, N=12, D=1024, seed=5)
    a = losses.wms_loss_value_and_grad(dist, emb, 0.8, 15.0)
    b = losses.wms_loss_value_and_grad(dist, emb, 0.8, 15.0)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


def test_wms_zero_row_edge_case(cuda_lib):
    """An all-zero descriptor hits the 1e-12 clamp of tf.nn.l2_normalize; loss and gradient stay finite and match."""
    from soft_contrastive_learning_b200 import losses
    emb, dist, _ = synth.wms_batch(T=2, P=4, N=4, D=64, seed=1)
    emb[0, 3] = 0.0
    loss, grad = losses.wms_loss_value_and_grad(dist, emb, 0.8, 15.0)
    ref, (rg,) = _oracle_wms(emb, dist)
    assert np.isfinite(loss) and np.isfinite(grad).all()
    assert rel(loss, ref) < LOSS_TOL


# ---------------- flat mode ----------------
@pytest.mark.parametrize("tag,mining", [("mine", True), ("nomine", False)])
def test_ms_golden_flat(cuda_lib, golden, tag, mining):
    from soft_contrastive_learning_b200 import losses
    g = golden("ms_T3_P4_N5_D48")
    loss, grad = losses.ms_loss_value_and_grad(g["labels"], g["emb"], ms_mining=mining)
    assert rel(loss, float(g["loss_" + tag])) < LOSS_TOL
    assert grad_err(grad, g["grad_" + tag]) < GRAD_TOL


@pytest.mark.parametrize("T,D", [(8, 1024), (32, 4096)], ids=["B256_D1024", "config3_B1024_D4096"])
def test_flat_wms_and_ms_batch(cuda_lib, measured, T, D):
    """Flat-mode W1 / W2 on a whole batch of T tuples x 32 descriptors; the second case is BASELINE config 3 at its full
    size (B = 1024, D = 4096: one 1024 x 1024 Gram over 4096 columns, float64 oracle included).  The batch is
    guard-banded (tests/_guardband.py: no pair within 1e-4 of a mining threshold), so the kept-masks must be IDENTICAL."""
    from soft_contrastive_learning_b200 import losses
    import _guardband as gb
    P, N = 15, 16
    rng = np.random.default_rng(42)
    xy = synth.tuple_xy(rng, T, P, N).reshape(-1, 2)
    emb = synth.tuple_descriptors(rng, T, P, N, D).reshape(T * 32, D)
    dist = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1)).astype(np.float32)
    labels = losses.ms_labels(T, P, N)
    emb, rounds, margin = gb.guard_band(emb, [gb.wms_masks64(dist, 0.8, 15.0), gb.ms_masks64(labels)], band=1e-4)
    assert margin >= 1e-4
    loss, grad, kept = losses.wms_loss_value_and_grad(dist, emb, 0.8, 15.0, return_kept=True)
    ref, (rg,) = ol.value_and_grad(lambda e: ol.wms_loss(torch.as_tensor(dist.astype(np.float64)), e, 0.8, 15.0),
                                   [emb.astype(np.float64)])
    measured(f"wms_flat_B{T * 32}_D{D}", loss=rel(loss, ref), grad=grad_err(grad, rg), guard_rounds=rounds)
    assert rel(loss, ref) < LOSS_TOL and grad_err(grad, rg) < GRAD_TOL
    _, mp, mn = ol.wms_loss(dist.astype(np.float64), emb.astype(np.float64), 0.8, 15.0, return_masks=True)
    k = kept.cpu().numpy().astype(bool)
    assert np.array_equal(k[0], mp.numpy()) and np.array_equal(k[1], mn.numpy())
    assert 0 < mp.sum() < mp.numel() and 0 < mn.sum()                          # the mining branch is exercised
    for mining in (True, False):
        loss, grad, kept = losses.ms_loss_value_and_grad(labels, emb, ms_mining=mining, return_kept=True)
        ref, (rg,) = ol.value_and_grad(lambda e: ol.ms_loss(labels, e, ms_mining=mining), [emb.astype(np.float64)])
        measured(f"ms_flat_B{T * 32}_D{D}_mining{int(mining)}", loss=rel(loss, ref), grad=grad_err(grad, rg))
        assert rel(loss, ref) < LOSS_TOL and grad_err(grad, rg) < GRAD_TOL
        _, mp, mn = ol.ms_loss(labels, emb.astype(np.float64), ms_mining=mining, return_masks=True)
        k = kept.cpu().numpy().astype(bool)
        assert np.array_equal(k[0], mp.numpy()) and np.array_equal(k[1], mn.numpy())


# ---------------- triplet family ----------------
TUPLE_CASES = ["triplet", "lazy_triplet", "quadruplet", "lazy_quadruplet", "evil_triplet", "evil_quadruplet",
               "huber_distance_triplet", "huber_distance_lazy_triplet", "distance_triplet",
               "distance_quadruplet", "huber_distance_quadruplet", "distance_lazy_quadruplet",
               "huber_distance_lazy_quadruplet"]


def _run_named(losses, tag, emb_t, sq, P, N, m1, m2, lam, dmax, fmax):
    q, p, n, o = (emb_t[:, 0:1], emb_t[:, 1:1 + P], emb_t[:, 1 + P:1 + P + N], emb_t[:, 1 + P + N:])
    if tag == "triplet":
        return losses.triplet_loss(q, p, n, m1)
    if tag == "lazy_triplet":
        return losses.lazy_triplet_loss(q, p, n, m1)
    if tag == "quadruplet":
        return losses.quadruplet_loss(q, p, n, o, m1, m2)
    if tag == "lazy_quadruplet":
        return losses.lazy_quadruplet_loss(q, p, n, o, m1, m2)
    if tag == "evil_triplet":
        return losses.evil_triplet_loss(q, p, n, m1)
    if tag == "evil_quadruplet":
        return losses.evil_quadruplet_loss(q, p, n, o, m1, m2)
    trip = "lazy_triplet_loss" if "lazy" in tag else "triplet_loss"
    dl = "huber_distance_loss" if "huber" in tag else "distance_loss"
    if "quadruplet" in tag:
        return losses.distance_quadruplet_loss(q, p, n, o, m1, m2, lam, sq, dmax, fmax, trip, dl)
    return losses.distance_triplet_loss(q, p, n, m1, lam, sq, dmax, fmax, trip, dl)


@pytest.mark.parametrize("tag", TUPLE_CASES)
def test_tuple_losses_golden(cuda_lib, golden, tag):
    from soft_contrastive_learning_b200 import losses
    g = golden("tuple_losses_T3_P4_N6_D40")
    P, N = int(g["P"]), int(g["N"])
    quad = "quadruplet" in tag
    emb = g["emb"] if quad else g["emb"][:, :1 + P + N]
    e = torch.tensor(emb, device="cuda", requires_grad=True)
    loss = _run_named(losses, tag, e, g["sq_d_dists"], P, N, float(g["m1"]), float(g["m2"]), float(g["lam"]),
                      float(g["d_max_squared"]), float(g["f_max_squared"]))
    loss.backward()
    assert rel(float(loss), float(g["loss_" + tag])) < LOSS_TOL
    ref_grad = g["grad_" + tag] if quad else g["grad_" + tag][:, :1 + P + N]
    assert grad_err(e.grad.cpu().numpy(), ref_grad) < GRAD_TOL


@pytest.mark.parametrize("name", ["triplet_loss", "lazy_quadruplet_loss", "quadruplet_loss"])
def test_tuple_losses_config3_shape(cuda_lib, name):
    """T=32, S=32, D=4096 (config 3).  Descriptors scaled so hinges are partly active."""
    from soft_contrastive_learning_b200 import losses
    T, P, D = 32, 15, 4096
    quad = "quadruplet" in name
    N = 15 if quad else 16
    rng = np.random.default_rng(7)
    emb = (0.011 * synth.tuple_descriptors(rng, T, P, N, D, other=quad, pos_noise=1.35)).astype(np.float32)
    loss, grad = losses.tuple_loss_value_and_grad(name, emb.reshape(T * 32, D), T, P, N, m1=0.1, m2=0.2)
    fn = getattr(ol, name)

    def f(e):
        parts = ol.split_tuple(e, P, N, other=quad)
        return fn(*parts, 0.1, 0.2) if quad else fn(*parts, 0.1)
    ref, (rg,) = ol.value_and_grad(f, [emb.astype(np.float64)])
    assert ref > 0
    assert rel(loss, ref) < LOSS_TOL and grad_err(grad.reshape(emb.shape), rg) < GRAD_TOL


@pytest.mark.parametrize("dl", ["distance_loss", "huber_distance_loss"])
@pytest.mark.parametrize("trip", ["triplet_loss", "lazy_triplet_loss"])
def test_distance_quadruplet_active_second_hinge(cuda_lib, trip, dl):
    """distance_quadruplet_loss (losses.py:267-307) with the `other` negative close enough to the
    negatives that the distance-term hinge is active in most tuples (the golden case only exercises it for 'distance')."""
    from soft_contrastive_learning_b200 import losses
    T, P, N, D = 16, 12, 11, 512
    rng = np.random.default_rng(31)
    xy = synth.tuple_xy(rng, T, P, N, other=True)
    emb = (0.03 * synth.tuple_descriptors(rng, T, P, N, D, other=True, pos_noise=1.2)).astype(np.float32)
    emb[:, -1] = emb[:, 1 + P] + 0.02 * rng.standard_normal((T, D)).astype(np.float32)     # other ~ first negative
    sqd = synth.anchor_sq_dists(xy, P).astype(np.float32)
    e = torch.tensor(emb, device="cuda", requires_grad=True)
    q, p, n, o = e[:, 0:1], e[:, 1:1 + P], e[:, 1 + P:1 + P + N], e[:, 1 + P + N:]
    loss = losses.distance_quadruplet_loss(q, p, n, o, 0.1, 0.2, 0.5, sqd, 225.0, 2.0, trip, dl)
    loss.backward()

    def f(x):
        parts = ol.split_tuple(x, P, N, other=True)
        return ol.distance_quadruplet_loss(*parts, 0.1, 0.2, 0.5, torch.as_tensor(sqd.astype(np.float64)), 225.0, 2.0,
                                           trip, dl)
    ref, (rg,) = ol.value_and_grad(f, [emb.astype(np.float64)])
    base, _ = ol.value_and_grad(lambda x: ol.distance_triplet_loss(*ol.split_tuple(x, P, N, other=True)[:3], 0.1, 0.5,
                                                                   torch.as_tensor(sqd.astype(np.float64)), 225.0, 2.0,
                                                                   trip, dl), [emb.astype(np.float64)])
    assert ref > base + 1e-3                                   # the second hinge contributes
    assert rel(float(loss), ref) < LOSS_TOL and grad_err(e.grad.cpu().numpy(), rg) < GRAD_TOL
    # the --loss name dispatch of train.py:729-763 reaches the same kernel
    name = ("huber_" if "huber" in dl else "") + "distance_" + ("lazy_" if "lazy" in trip else "") + "quadruplet"
    cfg = dict(TUPLES_PER_BATCH=T, POSITIVES_PER_TUPLE=P, NEGATIVES_PER_TUPLE=N, MARGIN_1=0.1, MARGIN_2=0.2, LAM=0.5)
    with torch.no_grad():
        l2 = losses.get_loss(name)(e.detach().reshape(T * (P + N + 2), D), sqd, cfg)
    assert float(l2) == float(loss)


@pytest.mark.parametrize("tag,dl", [("squared", "distance_loss"), ("huber", "huber_distance_loss")])
def test_pairwise_distance_loss_golden_and_shapes(cuda_lib, golden, tag, dl):
    """pairwise_distance_loss (losses.py:627-646): golden from the reference source, then a wider shape
    (P = 12, D = 4096, T = 32) against the oracle."""
    from soft_contrastive_learning_b200 import losses
    g = golden("pairwise_distance_loss_T3_P5_D40")
    e = torch.tensor(g["emb"], device="cuda", requires_grad=True)
    loss = losses.pairwise_distance_loss(e[:, :1], e[:, 1:], g["pairwise_sq_d"], float(g["d_max_squared"]),
                                         float(g["f_max_squared"]), dl)
    (2.0 * loss).backward()
    assert rel(float(loss), float(g["loss_" + tag])) < LOSS_TOL
    assert grad_err(e.grad.cpu().numpy() / 2.0, g["grad_" + tag]) < GRAD_TOL
    T, P, D = 32, 12, 4096
    rng = np.random.default_rng(17)
    xy = synth.tuple_xy(rng, T, P, 2)[:, :P + 1]
    emb = (0.02 * synth.tuple_descriptors(rng, T, P, 2, D, pos_noise=1.2)[:, :P + 1]).astype(np.float32)
    sqd = ((xy[:, :, None, :] - xy[:, None, :, :]) ** 2).sum(-1).astype(np.float32)
    lv, gv = losses.pairwise_distance_loss_value_and_grad(emb, sqd, 225.0, 2.0, dl)
    ref, (rg,) = ol.value_and_grad(lambda x: ol.pairwise_distance_loss(x[:, :1], x[:, 1:], torch.as_tensor(sqd.astype(np.float64)),
                                                                       225.0, 2.0, dl), [emb.astype(np.float64)])
    assert rel(lv, ref) < LOSS_TOL and grad_err(gv, rg) < GRAD_TOL


def test_logratio_golden_and_tuples(cuda_lib, golden):
    from soft_contrastive_learning_b200 import losses
    g = golden("logratio_P5_N5_D32")
    e = torch.tensor(g["emb"], device="cuda", requires_grad=True)
    loss = losses.logratio_loss(e[:, :1], e[:, 1:6], e[:, 6:], g["sq_pos"], g["sq_neg"])
    loss.backward()
    assert rel(float(loss), float(g["loss"])) < LOSS_TOL
    assert grad_err(e.grad.cpu().numpy(), g["grad"]) < GRAD_TOL
    # tuple mode and the non-strict (all-pairs) variant against the oracle
    rng = np.random.default_rng(4)
    T, P, N, D = 6, 12, 12, 512
    xy = synth.tuple_xy(rng, T, P, N)
    emb = synth.tuple_descriptors(rng, T, P, N, D)
    sp, sn = synth.logratio_sq_dists(xy, P, N)
    loss, grad = losses.logratio_loss_value_and_grad(emb.reshape(T * 25, D), T, P, N, sp.astype(np.float32),
                                                     sn.astype(np.float32))
    ref, (rg,) = ol.value_and_grad(
        lambda x: ol.logratio_loss_tuples(*ol.split_tuple(x, P, N), torch.as_tensor(sp.astype(np.float32).astype(np.float64)),
                                          torch.as_tensor(sn.astype(np.float32).astype(np.float64))),
        [emb.astype(np.float64)])
    assert rel(loss, ref) < LOSS_TOL and grad_err(grad.reshape(emb.shape), rg) < GRAD_TOL


def test_pairwise_sqdist(cuda_lib, golden):
    from soft_contrastive_learning_b200 import losses
    g = golden("pairwise_sqdist")
    out = losses._pairwise_squared_distances(g["selfcheck_in"].astype(np.float32))
    assert np.array_equal(out, g["selfcheck_out"].astype(np.float32))          # the reference's own self-check tensor
    out = losses._pairwise_squared_distances(g["x"])
    assert np.allclose(out, g["d"], rtol=1e-5, atol=1e-4)


def test_get_loss_by_name(cuda_lib):
    from soft_contrastive_learning_b200 import losses
    emb, dist, xy = synth.wms_batch(T=2, P=12, N=12, D=128, seed=8)
    out = torch.tensor(emb.reshape(50, 128), device="cuda", requires_grad=True)
    cfg = dict(TUPLES_PER_BATCH=2)
    l = losses.get_loss("wms")(out, torch.tensor(dist, device="cuda"), cfg)
    l.backward()
    ref, _ = _oracle_wms(emb, dist)
    assert rel(float(l), ref) < LOSS_TOL and out.grad.abs().sum() > 0
    sq = synth.anchor_sq_dists(xy, 12).astype(np.float32)
    l2 = losses.get_loss("huber_distance_triplet")(out.detach(), sq, cfg)
    q, p, n = ol.split_tuple(torch.as_tensor(emb.astype(np.float64)), 12, 12)
    ref2 = float(ol.distance_triplet_loss(q, p, n, 0.1, 0.5, torch.as_tensor(sq.astype(np.float64)), 225.0, 2.0))
    assert rel(float(l2), ref2) < LOSS_TOL
