"""wms config 1 (T = 32): launch time forward-only vs forward+backward, and vs the cluster size knob."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import losses, synth, _lib
T, S, D = 32, 25, 4096
emb_s, dist_s, _ = synth.wms_batch(T=T, P=12, N=12, D=D, seed=42)
emb = torch.tensor(emb_s, device="cuda"); dist = torch.tensor(dist_s, device="cuda")
params = losses._ms_params(0.8, 15.0)
def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(f"fwd+bwd {timeit(lambda: losses._wms_tuple_raw(emb, dist, params, need_grad=True)):.2f} us   fwd {timeit(lambda: losses._wms_tuple_raw(emb, dist, params, need_grad=False)):.2f} us")
for name in ("SCL_WMS_CLUSTER", "SCL_WMS_RESIDENT_CLUSTER", "SCL_WMS_C"):
    try:
        for c in (2, 4, 8, 16):
            with _lib.tuning(**{name: c}):
                print(name, c, f"{timeit(lambda: losses._wms_tuple_raw(emb, dist, params, need_grad=True)):.2f} us")
    except Exception as e:
        print(name, "n/a", str(e)[:60])
