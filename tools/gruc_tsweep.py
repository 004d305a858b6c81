#!/usr/bin/env python
"""Development aid: duration of the recurrent kernels of one bidirectional GRU layer as a function of the sequence length
(fixed per-launch cost = intercept).  Run under `ncu --metrics gpu__time_duration.sum -k regex:gru_` or read the event
times printed here (projection GEMM included).  usage: python tools/gruc_tsweep.py [clips] [H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from speech2affective_gestures_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
In = 8 if H == 64 else 88
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
ps = []
for d in range(2):
    ps += [torch.randn(3 * H, In, generator=g) * 0.05, torch.randn(3 * H, H, generator=g) * 0.05,
           torch.randn(3 * H, generator=g) * 0.05, torch.randn(3 * H, generator=g) * 0.05]
ps = [t.to(dev) for t in ps]
for T in (2, 34, 68):
    x = torch.randn(B, T, In, generator=g).to(dev)
    with torch.no_grad():
        for _ in range(3):
            ops.bigru(x, ps, 1, H, 0.0, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.bigru(x, ps, 1, H, 0.0, False)
        e1.record()
        torch.cuda.synchronize()
        print("T=%3d: layer forward (projection + recurrence) %.1f us" % (T, e0.elapsed_time(e1) * 50))
