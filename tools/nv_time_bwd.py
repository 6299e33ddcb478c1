import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import netvlad
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn((B, 30, 40, 512), generator=g, device="cuda").requires_grad_(True); aw = (0.05 * torch.randn((512, 64), generator=g, device="cuda")).requires_grad_(True); cc = (0.05 * torch.randn((512, 64), generator=g, device="cuda")).requires_grad_(True)
dout = torch.randn((B, 512 * 64), generator=g, device="cuda")
for _ in range(3):
    out = netvlad.netVLAD(x, aw, cc); out.backward(dout)
torch.cuda.synchronize()
