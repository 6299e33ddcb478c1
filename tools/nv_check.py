#!/usr/bin/env python
"""Fused vs generic NetVLAD forward/backward: agreement and device time at BASELINE config 2 (B=256, 30x40x512).
    gpurun -- 'python tools/nv_check.py [B]'"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import _lib, netvlad, synth  # noqa: E402


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def run(x, aw, cc, dout, fused):
    with _lib.tuning(SCL_NV_FUSED=int(fused)):
        xt, wt, ct = x.clone().requires_grad_(True), aw.clone().requires_grad_(True), cc.clone().requires_grad_(True)
        out = netvlad.netVLAD(xt, wt, ct)
        (out * dout).sum().backward()
        torch.cuda.synchronize()
    return out.detach(), xt.grad, wt.grad, ct.grad


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    small = [(2, 3, 4), (3, 11, 15), (2, 30, 40), (5, 9, 13), (40, 30, 40)]
    g = torch.Generator(device="cuda").manual_seed(1)
    for B, H, W in small:
        x = torch.randn((B, H, W, 512), generator=g, device="cuda")
        aw = 0.05 * torch.randn((512, 64), generator=g, device="cuda")
        cc = 0.05 * torch.randn((512, 64), generator=g, device="cuda")
        dout = torch.randn((B, 512 * 64), generator=g, device="cuda")
        f = run(x, aw, cc, dout, True)
        r = run(x, aw, cc, dout, False)
        print(f"B={B} {H}x{W}: fused vs generic  out {relmax(f[0], r[0]):.2e} dx {relmax(f[1], r[1]):.2e} "
              f"dW {relmax(f[2], r[2]):.2e} dC {relmax(f[3], r[3]):.2e}", flush=True)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    x = torch.randn((B, 30, 40, 512), generator=g, device="cuda")
    aw = 0.05 * torch.randn((512, 64), generator=g, device="cuda")
    cc = 0.05 * torch.randn((512, 64), generator=g, device="cuda")
    dout = torch.randn((B, 512 * 64), generator=g, device="cuda")
    for fused in (1, 0):
        with _lib.tuning(SCL_NV_FUSED=fused):
            xt, wt, ct = x.clone().requires_grad_(True), aw.clone().requires_grad_(True), cc.clone().requires_grad_(True)
            tf = timed(lambda: netvlad.netVLAD(xt, wt, ct))
            out = netvlad.netVLAD(xt, wt, ct)
            tb = timed(lambda: out.backward(dout, retain_graph=True))
            print(f"B={B} fused={fused}: fwd {tf:.4f} ms  bwd {tb:.4f} ms  (x {x.numel() * 4 / 1e6:.0f} MB: "
                  f"{x.numel() * 4 / tf / 1e6:.0f} GB/s fwd)", flush=True)


if __name__ == "__main__":
    main()
