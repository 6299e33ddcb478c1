"""wms streaming kernel, T = 4096: forward only vs forward + backward (what the two passes cost)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import losses, synth
T, S, D = 4096, 25, 4096
emb_s, dist_s, _ = synth.wms_batch(T=64, P=12, N=12, D=D, seed=42)
emb = torch.tensor(emb_s, device="cuda").repeat(T // 64, 1, 1).contiguous(); dist = torch.tensor(dist_s, device="cuda").repeat(T // 64, 1, 1).contiguous()
emb += 1e-3 * torch.randn_like(emb)
params = losses._ms_params(0.8, 15.0)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"fwd+bwd {timeit(lambda: losses._wms_tuple_raw(emb, dist, params, need_grad=True)):.4f} ms")
print(f"fwd     {timeit(lambda: losses._wms_tuple_raw(emb, dist, params, need_grad=False)):.4f} ms")
