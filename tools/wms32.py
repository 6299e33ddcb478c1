import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import losses, synth, _lib
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
emb, dist, _ = synth.wms_batch(T=T, P=12, N=12, D=4096, seed=1)
e = torch.tensor(emb, device="cuda"); d = torch.tensor(dist, device="cuda")
p = losses._ms_params(0.8, 15.0)
for cl in (None, 4, 8):
    with _lib.tuning(**({} if cl is None else {"SCL_WMS_CLUSTER": cl})):
        for _ in range(5): losses._wms_tuple_raw(e, d, p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): losses._wms_tuple_raw(e, d, p)
        e1.record(); torch.cuda.synchronize()
        print(f"T={T} cluster={cl}: {e0.elapsed_time(e1) / 50 * 1000:.1f} us per call", flush=True)
