import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import netvlad
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn((B, 30, 40, 512), generator=g, device="cuda"); aw = 0.05 * torch.randn((512, 64), generator=g, device="cuda"); cc = 0.05 * torch.randn((512, 64), generator=g, device="cuda")
for _ in range(3): netvlad.netVLAD(x, aw, cc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): netvlad.netVLAD(x, aw, cc)
e1.record(); torch.cuda.synchronize()
print(f"dbg={os.environ.get('SCL_NV_DBG','0')} fwd {e0.elapsed_time(e1)/10:.4f} ms", flush=True)
