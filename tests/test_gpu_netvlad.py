"""GPU parity: NetVLAD head and PCA projection (forward and backward) vs the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import netvlad as onv
from soft_contrastive_learning_b200 import synth

pytestmark = pytest.mark.gpu


def relmax(a, b):
    return np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30)


NV_TOL = 1e-5      # north_star: 1e-5 relative (fp32); gradients relative to their max-norm


@pytest.mark.parametrize("B,H,W", [(2, 3, 4), (3, 11, 15), (2, 30, 40), (5, 9, 13)])
def test_netvlad_forward_backward(cuda_lib, measured, B, H, W):
    from soft_contrastive_learning_b200 import netvlad
    x, aw, cc, *_ = synth.netvlad_problem(B=B, H=H, W=W, seed=42)
    rng = np.random.default_rng(0)
    dout = rng.standard_normal((B, 512 * 64)).astype(np.float32)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    wt = torch.tensor(aw, device="cuda", requires_grad=True)
    ct = torch.tensor(cc, device="cuda", requires_grad=True)
    out = netvlad.netVLAD(xt, wt.reshape(1, 1, 512, 64), ct.reshape(1, 1, 1, 512, 64))
    (out * torch.tensor(dout, device="cuda")).sum().backward()
    xo = torch.tensor(x.astype(np.float64), requires_grad=True)
    wo = torch.tensor(aw.astype(np.float64), requires_grad=True)
    co = torch.tensor(cc.astype(np.float64), requires_grad=True)
    ro = onv.netvlad_head(xo, wo, co)
    (ro * torch.tensor(dout.astype(np.float64))).sum().backward()
    errs = dict(out=relmax(out.detach().cpu().numpy(), ro.detach().numpy()), dx=relmax(xt.grad.cpu().numpy(), xo.grad.numpy()),
                dW=relmax(wt.grad.cpu().numpy(), wo.grad.numpy()), dC=relmax(ct.grad.cpu().numpy(), co.grad.numpy()))
    measured(f"netvlad_B{B}_{H}x{W}", **errs)
    assert max(errs.values()) < NV_TOL, errs
    assert np.allclose((out.detach().cpu().numpy() ** 2).sum(1), 1.0, atol=1e-5)


def test_netvlad_and_pca_config2_full_size(cuda_lib, measured):
    """BASELINE config 2 at its size: B = 256 maps of 30x40x512, K = 64, PCA 32768 -> 4096, forward and backward on the GPU.
    The float64 oracle runs on a 16-image subset (NetVLAD is per image; dW / dC are sums over images, checked by
    linearity: the B = 256 result equals the sum of sixteen 16-image calls, one of which is held to the oracle) and on a
    subset of the PCA's output / input columns (all 256 rows)."""
    from soft_contrastive_learning_b200 import netvlad
    free, _ = torch.cuda.mem_get_info()
    if free < 12 * 2 ** 30:
        pytest.skip("needs ~8 GB of HBM")
    B, H, W, Cc, K, Dout = 256, 30, 40, 512, 64, 4096
    Din = Cc * K
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.randn((B, H, W, Cc), generator=g, device="cuda")
    aw = 0.05 * torch.randn((Cc, K), generator=g, device="cuda")
    cc = 0.05 * torch.randn((Cc, K), generator=g, device="cuda")
    V = torch.randn((Dout, Din), generator=g, device="cuda") / Din ** 0.5
    m = 0.01 * torch.randn(Din, generator=g, device="cuda")
    var = 0.5 + 1.5 * torch.rand(Dout, generator=g, device="cuda")
    dy = torch.randn((B, Dout), generator=g, device="cuda")

    def run(xs, dys):
        xt, wt, ct = xs.clone().requires_grad_(True), aw.clone().requires_grad_(True), cc.clone().requires_grad_(True)
        vlad = netvlad.netVLAD(xt, wt, ct)
        vlad.retain_grad()
        y = netvlad.pca_project(vlad, V, m, var)
        (y * dys).sum().backward()
        return vlad.detach(), y.detach(), vlad.grad, xt.grad, wt.grad, ct.grad

    vlad, y, dvlad, dx, dw, dc = run(x, dy)
    assert torch.isfinite(y).all() and torch.isfinite(dx).all()
    # (1) linearity in the batch: sum of 16-image calls == the 256-image call (dW, dC); per-image results identical
    dw_sum, dc_sum = torch.zeros_like(dw), torch.zeros_like(dc)
    for b0 in range(0, B, 16):
        v16, y16, dv16, dx16, dw16, dc16 = run(x[b0:b0 + 16], dy[b0:b0 + 16])
        dw_sum += dw16
        dc_sum += dc16
        if b0 in (0, 112, 240):
            assert relmax(v16.cpu().numpy(), vlad[b0:b0 + 16].double().cpu().numpy()) < 2e-6
            assert relmax(dx16.cpu().numpy(), dx[b0:b0 + 16].double().cpu().numpy()) < 2e-6
    lin = dict(dW=relmax(dw_sum.cpu().numpy(), dw.double().cpu().numpy()), dC=relmax(dc_sum.cpu().numpy(), dc.double().cpu().numpy()))
    measured("netvlad_config2_linearity", **lin)
    assert max(lin.values()) < 5e-6, lin
    # (2) the float64 oracle on images 96..111 with the gradient that actually arrived from the PCA backward
    sl = slice(96, 112)
    xo = x[sl].double().cpu().requires_grad_(True)
    wo, co = aw.double().cpu().requires_grad_(True), cc.double().cpu().requires_grad_(True)
    ro = onv.netvlad_head(xo, wo, co)
    (ro * dvlad[sl].double().cpu()).sum().backward()
    _, _, _, dx_s, dw_s, dc_s = run(x[sl], dy[sl])
    errs = dict(out=relmax(vlad[sl].cpu().numpy(), ro.detach().numpy()), dx=relmax(dx[sl].cpu().numpy(), xo.grad.numpy()),
                dW16=relmax(dw_s.cpu().numpy(), wo.grad.numpy()), dC16=relmax(dc_s.cpu().numpy(), co.grad.numpy()))
    measured("netvlad_config2_vs_oracle_16_images", **errs)
    assert max(errs.values()) < NV_TOL, errs
    # (3) PCA 32768 -> 4096 against float64 on 256 output columns (forward) and 512 input columns (backward), all rows
    gi = torch.Generator().manual_seed(1)
    oc = torch.randperm(Dout, generator=gi)[:256].cuda()
    ic = torch.randperm(Din, generator=gi)[:512].cuda()
    yo = ((vlad.double() - m.double()) @ V[oc].double().t()) / var[oc].double().sqrt()
    dxo = (dy.double() / var.double().sqrt()) @ V[:, ic].double()
    perr = dict(y=relmax(y[:, oc].cpu().numpy(), yo.cpu().numpy()), dvlad=relmax(dvlad[:, ic].cpu().numpy(), dxo.cpu().numpy()))
    measured("pca_config2_32768_to_4096", **perr)
    assert max(perr.values()) < NV_TOL, perr


def test_pca_forward_backward_and_sklearn(cuda_lib, measured):
    from sklearn.decomposition import PCA
    from soft_contrastive_learning_b200 import netvlad
    x, aw, cc, V, m, var = synth.netvlad_problem(B=4, H=2, W=2, Dout=256, seed=1)
    rng = np.random.default_rng(1)
    feats = rng.standard_normal((4, 32768)).astype(np.float32) * 0.01
    ft = torch.tensor(feats, device="cuda", requires_grad=True)
    y = netvlad.pca_project(ft, torch.tensor(V, device="cuda"), torch.tensor(m, device="cuda"),
                            torch.tensor(var, device="cuda"))
    dy = rng.standard_normal(y.shape).astype(np.float32)
    (y * torch.tensor(dy, device="cuda")).sum().backward()
    fo = torch.tensor(feats.astype(np.float64), requires_grad=True)
    yo = onv.pca_project(fo, V.astype(np.float64), m.astype(np.float64), var.astype(np.float64))
    (yo * torch.tensor(dy.astype(np.float64))).sum().backward()
    measured("pca_B4_Dout256", y=relmax(y.detach().cpu().numpy(), yo.detach().numpy()), dx=relmax(ft.grad.cpu().numpy(), fo.grad.numpy()))
    assert relmax(y.detach().cpu().numpy(), yo.detach().numpy()) < 1e-5
    assert relmax(ft.grad.cpu().numpy(), fo.grad.numpy()) < 1e-5
    # evaluation twin: sklearn PCA(whiten=True).transform (top-n.py:74-77)
    X = (rng.standard_normal((300, 64)) @ rng.standard_normal((64, 64))).astype(np.float32)
    pca = PCA(whiten=True, n_components=16).fit(X)
    v2, m2, var2 = netvlad.pca_from_sklearn(pca)
    got = netvlad.pca_project(X[:32], v2, m2, var2)
    assert np.allclose(got, pca.transform(X[:32]), rtol=2e-4, atol=2e-4)


# ---------------- tcgen05 GEMM behind PCA (tc_gemm.cu) ----------------
@pytest.mark.parametrize("B,Din,Dout", [(96, 2048, 256),      # full tiles
                                        (70, 1000, 132),      # ragged: K % 32 != 0, M and N tails
                                        (256, 4096, 64),      # narrow output tile (BN = 64)
                                        (3, 36, 8), (130, 520, 260)])
@pytest.mark.parametrize("precision", [0, 1])
def test_pca_tensor_core_gemm_shapes(cuda_lib, B, Din, Dout, precision):
    """fp32-grade 3xTF32 (precision 0) must meet the fp32 tolerance of the reference graph; the single-pass TF32 mode has
    a stated tolerance of 2e-3 of the result's max magnitude.  Backward reads V MN-major (no transposed copy)."""
    from soft_contrastive_learning_b200 import netvlad
    rng = np.random.default_rng(B * 7 + Dout)
    x = rng.standard_normal((B, Din)).astype(np.float32)
    V = (rng.standard_normal((Dout, Din)) / np.sqrt(Din)).astype(np.float32)
    m = (0.1 * rng.standard_normal(Din)).astype(np.float32)
    var = rng.uniform(0.5, 2.0, Dout).astype(np.float32)
    dy = rng.standard_normal((B, Dout)).astype(np.float32)
    netvlad.set_gemm_precision(precision)
    try:
        xt = torch.tensor(x, device="cuda", requires_grad=True)
        y = netvlad.pca_project(xt, torch.tensor(V, device="cuda"), torch.tensor(m, device="cuda"), torch.tensor(var, device="cuda"))
        (y * torch.tensor(dy, device="cuda")).sum().backward()
    finally:
        netvlad.set_gemm_precision(0)
    yo = ((x.astype(np.float64) - m) @ V.astype(np.float64).T) / np.sqrt(var.astype(np.float64))
    dxo = (dy.astype(np.float64) / np.sqrt(var.astype(np.float64))) @ V.astype(np.float64)
    tol = 1e-5 if precision == 0 else 2e-3
    assert relmax(y.detach().cpu().numpy(), yo) < tol
    assert relmax(xt.grad.cpu().numpy(), dxo) < tol


def test_pca_tensor_core_matches_simt_fallback(cuda_lib, tune):
    from soft_contrastive_learning_b200 import netvlad
    rng = np.random.default_rng(5)
    x = rng.standard_normal((64, 1024)).astype(np.float32)
    V = (rng.standard_normal((128, 1024)) / 32).astype(np.float32)
    m = (0.1 * rng.standard_normal(1024)).astype(np.float32)
    var = rng.uniform(0.5, 2.0, 128).astype(np.float32)
    a = netvlad.pca_project(x, V, m, var)
    tune("SCL_GEMM_SIMT", 1)
    b = netvlad.pca_project(x, V, m, var)
    assert relmax(a, b) < 1e-5


@pytest.mark.parametrize("n,D,d", [(300, 1024, 16), (512, 130, 16), (200, 4096, 12)])
def test_pca_fit_matches_oracle_and_sklearn(cuda_lib, n, D, d):
    """PCA(whiten=True, n_components=d).fit on the GPU (Gram route n <= D, covariance route n > D,
    ragged D) against the float64 oracle and scikit-learn's exact solver; the fitted triple then drives pca_project.
    d stays inside the components that stand clear of the noise floor of the synthetic spectrum: below it the
    eigen-gaps vanish and any two solvers rotate the basis differently."""
    from sklearn.decomposition import PCA
    from soft_contrastive_learning_b200 import netvlad
    X = synth.pca_features(n, D, rank=40, seed=17, dtype=np.float32)
    v, m, var = netvlad.pca_fit(X, d)
    vo, mo, varo = onv.pca_fit(X, d)
    assert v.shape == (d, D) and m.shape == (D,) and var.shape == (d,)
    assert np.allclose(m, mo, rtol=1e-6, atol=1e-6)
    assert np.allclose(var, varo, rtol=2e-5)
    assert np.abs(v - vo).max() < 2e-4                      # unit-norm rows; eigenvector error ~ eps * |G| / gap
    pca = PCA(whiten=True, n_components=d, svd_solver="full").fit(X.astype(np.float64))
    got = netvlad.pca_project(X[:64], v, m, var)
    assert np.allclose(got, pca.transform(X[:64].astype(np.float64)), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(2500, 1300, 96),       # 20 x 11 = 220 tiles > #SMs: the persistent tile loop wraps
                                   (130, 68, 1000),        # ragged M, N, K (K not a multiple of the 32-float stage)
                                   (64, 36, 40)])          # narrow output (BN = 64), single partial stage
@pytest.mark.parametrize("precision", [0, 1])
def test_gemm_engine_all_layouts(cuda_lib, M, N, K, a_mn, b_mn, precision):
    """scl_gemm_tf32 (csrc/tc_gemm.cu) directly: K-major / MN-major operands, ragged shapes, more tiles than SMs."""
    import ctypes as C
    from soft_contrastive_learning_b200._lib import check, lib
    g = torch.Generator(device="cuda").manual_seed(M + 7 * N + 13 * K)
    pad = lambda n: (n + 3) // 4 * 4
    A = torch.randn((K, pad(M)) if a_mn else (M, pad(K)), generator=g, device="cuda")
    Bm = torch.randn((K, pad(N)) if b_mn else (N, pad(K)), generator=g, device="cuda")
    Ad = (A[:, :M].t() if a_mn else A[:, :K]).double()
    Bd = (Bm[:, :N].t() if b_mn else Bm[:, :K]).double()
    ref = Ad @ Bd.t()
    out = torch.full((M, pad(N)), float("nan"), device="cuda")
    scale = torch.rand(pad(N), generator=g, device="cuda") + 0.5
    check(lib().scl_gemm_tf32(C.c_void_p(A.data_ptr()), C.c_void_p(Bm.data_ptr()), C.c_void_p(out.data_ptr()), M, N, K,
                              A.stride(0), Bm.stride(0), out.stride(0), a_mn, b_mn, C.c_void_p(scale.data_ptr()), precision,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "scl_gemm_tf32")
    got = out[:, :N].double()
    ref = ref * scale[:N].double()
    tol = 2e-6 if precision == 0 else 2e-3                 # fp32-grade 3xTF32 vs one TF32 pass, relative to the result norm
    assert float((got - ref).abs().max() / ref.abs().max()) < tol
    if pad(N) != N:
        assert torch.isnan(out[:, N:]).all()               # nothing written past the logical width


@pytest.mark.parametrize("B,Din,Dout", [(5, 1000, 72), (130, 4096, 264), (3, 32768, 128)])
def test_pca_prepared_matches_float64_and_tracks_the_matrix(cuda_lib, measured, B, Din, Dout):
    """The prepared projection (matrix split once into fp16 hi / lo halves, csrc/tc_gemm_h3.cu) against float64, ragged tile
    shapes included, forward (matrix K-major) and backward (same shadow read MN-major); the shadow follows in-place updates
    of the tensor it was built from (version counter) and other tensors never see it."""
    from soft_contrastive_learning_b200 import netvlad
    g = torch.Generator(device="cuda").manual_seed(Din + Dout)
    x = torch.randn((B, Din), generator=g, device="cuda") * torch.logspace(-3, 2, B, device="cuda")[:, None]
    V = torch.randn((Dout, Din), generator=g, device="cuda") / Din ** 0.5
    m = 0.1 * torch.randn(Din, generator=g, device="cuda")
    var = 0.5 + 1.5 * torch.rand(Dout, generator=g, device="cuda")
    dy = torch.randn((B, Dout), generator=g, device="cuda")

    def run(Vt):
        xt = x.clone().requires_grad_(True)
        y = netvlad.pca_project(xt, Vt, m, var)
        (y * dy).sum().backward()
        return y.detach(), xt.grad

    def ref(Vt):
        y = ((x.double() - m.double()) @ Vt.double().t()) / var.double().sqrt()
        dx = (dy.double() / var.double().sqrt()) @ Vt.double()
        return y, dx

    assert netvlad._prepared_for(V) is not None            # this shape takes the prepared path
    y, dx = run(V)
    yr, dxr = ref(V)
    rowrel = lambda a, b: float(((a.double() - b).abs().amax(1) / b.abs().amax(1).clamp_min(1e-300)).max())
    errs = dict(y=rowrel(y, yr), dx=rowrel(dx, dxr))       # per row: the rows span five decades
    measured(f"pca_prepared_B{B}_{Din}_to_{Dout}", **errs)
    assert max(errs.values()) < NV_TOL, errs
    p0 = netvlad._prepared_for(V)
    assert netvlad._prepared_for(V) is p0                  # cached while unchanged
    V.mul_(-2.0)                                           # in-place update: the shadow must be rebuilt
    y2, dx2 = run(V)
    assert netvlad._prepared_for(V) is not p0
    yr2, dxr2 = ref(V)
    assert rowrel(y2, yr2) < NV_TOL and rowrel(dx2, dxr2) < NV_TOL
    W = V.clone()                                          # another tensor, same shape: its own shadow
    assert netvlad._prepared_for(W) is not netvlad._prepared_for(V)


@pytest.mark.parametrize("B,H,W", [(40, 30, 40), (200, 16, 24), (37, 13, 17)])
def test_netvlad_fused_kernels_against_the_generic_path(cuda_lib, measured, B, H, W):
    """The one-pass kernels of csrc/netvlad_fused.cu (forward; first half of the backward) against the generic
    rownorm / GEMM / softmax / GEMM path on batches where a persistent CTA walks several tiles and crosses image
    boundaries (the per-image accumulators are drained mid-CTA there; the small oracle shapes never do that).  Both paths
    are fp32-grade, so they agree to ~1e-6; repeated runs are bit-identical (no atomics)."""
    from soft_contrastive_learning_b200 import _lib, netvlad
    g = torch.Generator(device="cuda").manual_seed(B)
    x = torch.randn((B, H, W, 512), generator=g, device="cuda") * (0.1 + 3.0 * torch.rand((B, 1, 1, 1), generator=g, device="cuda"))
    aw = 0.05 * torch.randn((512, 64), generator=g, device="cuda")
    cc = 0.05 * torch.randn((512, 64), generator=g, device="cuda")
    dout = torch.randn((B, 512 * 64), generator=g, device="cuda") * torch.logspace(-2, 1, B, device="cuda")[:, None]

    def run(fused):
        with _lib.tuning(SCL_NV_FUSED=int(fused)):
            xt, wt, ct = x.clone().requires_grad_(True), aw.clone().requires_grad_(True), cc.clone().requires_grad_(True)
            out = netvlad.netVLAD(xt, wt, ct)
            (out * dout).sum().backward()
        return out.detach(), xt.grad, wt.grad, ct.grad

    f, r = run(True), run(False)
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
    # dx per image: the gradients of the images span three decades
    dx_err = float(((f[1].double() - r[1].double()).abs().amax((1, 2, 3)) / r[1].double().abs().amax((1, 2, 3))).max())
    errs = dict(out=rel(f[0], r[0]), dx=dx_err, dW=rel(f[2], r[2]), dC=rel(f[3], r[3]))
    measured(f"netvlad_fused_vs_generic_B{B}_{H}x{W}", **errs)
    assert max(errs.values()) < 5e-6, errs
    f2 = run(True)
    assert all(torch.equal(a, b) for a, b in zip(f, f2)), "fused kernels are not deterministic"
