"""Diagnostics for the wms tuple kernel (prints, no asserts)."""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import losses as ol
from soft_contrastive_learning_b200 import losses, synth

def bits(kept, S):
    k = kept.cpu().numpy().astype(np.uint32)
    b = ((k[..., None] >> np.arange(S, dtype=np.uint32)) & 1).astype(bool)
    return b[:, :, 0, :], b[:, :, 1, :]

for (T, P, N, D) in [(1, 12, 12, 64), (1, 12, 12, 512), (2, 12, 12, 4096), (1, 3, 4, 64)]:
    emb, dist, _ = synth.wms_batch(T=T, P=P, N=N, D=D, seed=42)
    S = 1 + P + N
    for mining in (False, True):
        params = losses._ms_params(0.8, 15.0, ms_mining=mining)
        loss, grad, kept, per = losses._wms_tuple_raw(torch.tensor(emb, device="cuda"), torch.tensor(dist, device="cuda"),
                                                     params, True, True, True)
        torch.cuda.synchronize()
        ref, (rg,) = ol.value_and_grad(lambda e: ol.wms_loss_tuples(torch.as_tensor(dist.astype(np.float64)), e, 0.8, 15.0, ms_mining=mining),
                                       [emb.astype(np.float64)])
        kp, kn = bits(kept, S)
        _, mp, mn = ol.wms_loss(dist[0].astype(np.float64), emb[0].astype(np.float64), 0.8, 15.0, return_masks=True, ms_mining=mining)
        g = grad.cpu().numpy()
        print(f"T={T} S={S} D={D} mining={mining}: loss {loss.item():.7f} ref {ref:.7f} per0 {per[0].item():.7f} "
              f"gerr {np.abs(g-rg).max()/np.abs(rg).max():.2e} kp diff {(kp[0]!=mp.numpy()).sum()} kn diff {(kn[0]!=mn.numpy()).sum()} "
              f"kp rows kernel {kp[0].sum(1)[:8]} oracle {mp.numpy().sum(1)[:8]}")
