#!/usr/bin/env python
"""Development aid: forward and backward time of one bidirectional GRU layer (recurrence + projections) for the
generator-sized (H=300) and discriminator-sized (H=64) configurations, with the cluster kernels (default where
supported) and the L2-exchange kernels (s2ag_debug_flags 1024 | 8192).  usage: python tools/gru_layer_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from speech2affective_gestures_b200 import _C, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _C.lib()
T = 34


def run(B, In, H, flags, reps=10):
    g = torch.Generator(device="cpu").manual_seed(0)
    ps = []
    for d in range(2):
        ps += [torch.randn(3 * H, In, generator=g) * 0.05, torch.randn(3 * H, H, generator=g) * 0.05,
               torch.randn(3 * H, generator=g) * 0.05, torch.randn(3 * H, generator=g) * 0.05]
    ps = [t.to(dev).requires_grad_(True) for t in ps]
    x = torch.randn(B, T, In, generator=g).to(dev).requires_grad_(True)
    gy = torch.randn(B, T, 2 * H, generator=g).to(dev)
    lib.s2ag_debug_flags(flags)
    tf = tb = 0.0
    for it in range(reps + 3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        y = ops.bigru(x, ps, 1, H, 0.0, True)
        e[1].record()
        y.backward(gy)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 3:
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    lib.s2ag_debug_flags(0)
    return tf / reps * 1e3, tb / reps * 1e3, y.detach(), [q.grad.clone() for q in ps] + [x.grad.clone()]


for name, B, In, H in (("G-sized", 256, 600, 300), ("D-sized pair", 512, 128, 64), ("D-sized", 256, 128, 64)):
    f0, b0, y0, g0 = run(B, In, H, 1024 | 8192)
    f1, b1, y1, g1 = run(B, In, H, 0)
    err = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-9)) for a, b in zip([y1] + g1, [y0] + g0))
    print("%-13s B=%d H=%d: L2-exchange fwd %.0f us bwd %.0f us | default (cluster where supported) fwd %.0f us bwd %.0f us"
          " | max rel diff %.1e" % (name, B, H, f0, b0, f1, b1, err))
