import numpy as np, torch, sys
sys.path.insert(0, '.')
from soft_contrastive_learning_b200 import netvlad
def relmax(a, b): return np.abs(a - b).max() / np.abs(b).max()
rng = np.random.default_rng(0)
for (B, Din, Dout) in [(96, 2048, 256), (70, 1000, 132), (256, 4096, 64), (4, 32768, 256)]:
    x = rng.standard_normal((B, Din)).astype(np.float32)
    V = (rng.standard_normal((Dout, Din)) / np.sqrt(Din)).astype(np.float32)
    m = (0.1 * rng.standard_normal(Din)).astype(np.float32)
    var = rng.uniform(0.5, 2.0, Dout).astype(np.float32)
    dy = rng.standard_normal((B, Dout)).astype(np.float32)
    yo = ((x.astype(np.float64) - m) @ V.astype(np.float64).T) / np.sqrt(var.astype(np.float64))
    dxo = (dy.astype(np.float64) / np.sqrt(var.astype(np.float64))) @ V.astype(np.float64)
    for prec in (0, 1):
        netvlad.set_gemm_precision(prec)
        xt = torch.tensor(x, device="cuda", requires_grad=True)
        y = netvlad.pca_project(xt, torch.tensor(V, device="cuda"), torch.tensor(m, device="cuda"), torch.tensor(var, device="cuda"))
        (y * torch.tensor(dy, device="cuda")).sum().backward()
        print(B, Din, Dout, 'prec', prec, 'fwd err', relmax(y.detach().cpu().numpy(), yo), 'bwd err', relmax(xt.grad.cpu().numpy(), dxo), flush=True)
netvlad.set_gemm_precision(0)
