import numpy as np, torch, sys
sys.path.insert(0, '.')
from soft_contrastive_learning_b200._lib import lib, check
from soft_contrastive_learning_b200.losses import _p, _stream
L = lib()
torch.manual_seed(0)
def run(M, N, K, a_mn, b_mn, prec):
    A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda')
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    C = torch.full((M, N), float('nan'), device='cuda')
    check(L.scl_gemm_tf32(_p(As), _p(Bs), _p(C), M, N, K, As.shape[1], Bs.shape[1], N, int(a_mn), int(b_mn), None, prec, _stream()), 'gemm')
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    nz = (C == 0).float().mean().item(); nn = torch.isnan(C).float().mean().item()
    print(f'M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn} prec{prec}: err {err:.3e} zeros {nz:.3f} nan {nn:.3f}', flush=True)
for (M, N, K) in [(128, 128, 32), (128, 128, 256), (256, 256, 1024), (128, 64, 64)]:
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            run(M, N, K, a_mn, b_mn, 1)
run(256, 256, 2048, 0, 0, 0); run(256, 256, 2048, 0, 1, 0); run(256, 256, 2048, 1, 1, 0)
run(256, 256, 32768, 0, 0, 0); run(128, 128, 4096, 0, 1, 0); run(128, 64, 1200, 1, 1, 0)
def bench(M, N, K, a_mn, b_mn, prec, tag):
    A = torch.randn(M, K, device='cuda'); B = torch.randn(N, K, device='cuda')
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    C = torch.empty((M, N), device='cuda')
    f = lambda: check(L.scl_gemm_tf32(_p(As), _p(Bs), _p(C), M, N, K, As.shape[1], Bs.shape[1], N, int(a_mn), int(b_mn), None, prec, _stream()), 'gemm')
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'{tag}: M{M} N{N} K{K} prec{prec}: {ms:.3f} ms, {2.0*M*N*K/ms/1e9:.1f} TFLOP/s (algorithmic)', flush=True)
for prec in (0, 1):
    bench(256, 4096, 32768, 0, 0, prec, 'pca fwd')
    bench(256, 32768, 4096, 0, 1, prec, 'pca bwd')
    bench(1024, 1024, 4096, 0, 0, prec, 'flat gram')
    bench(1024, 4096, 1024, 0, 1, prec, 'flat dE')
