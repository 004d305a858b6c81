#!/usr/bin/env python
"""Development aid: clock64 timeline of one CTA of the cluster BPTT kernel (csrc/umma_gru_cluster.cu).
usage: python tools/grucb_timeline.py [clips] [H]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from speech2affective_gestures_b200 import _C, ops  # noqa: E402

B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 34
H = int(sys.argv[2]) if len(sys.argv) > 2 else 300
In = 2 * H
dev = torch.device("cuda:0")
lib = _C.lib()
g = torch.Generator(device="cpu").manual_seed(0)
ps = []
for d in range(2):
    ps += [torch.randn(3 * H, In, generator=g) * 0.05, torch.randn(3 * H, H, generator=g) * 0.05,
           torch.randn(3 * H, generator=g) * 0.05, torch.randn(3 * H, generator=g) * 0.05]
ps = [t.to(dev).requires_grad_(True) for t in ps]
x = torch.randn(B, T, In, generator=g).to(dev).requires_grad_(True)
gy = torch.randn(B, T, 2 * H, generator=g).to(dev)
for _ in range(3):
    ops.bigru(x, ps, 1, H, 0.0, True).backward(gy)
torch.cuda.synchronize()
y = ops.bigru(x, ps, 1, H, 0.0, True)
lib.s2ag_debug_flags(8)
y.backward(gy)
torch.cuda.synchronize()
lib.s2ag_debug_flags(0)
buf = (ctypes.c_longlong * (64 * 16))()
lib.s2ag_debug_read_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.s2ag_debug_read_timeline(buf, 64 * 16) == 0
tl = [[buf[s * 16 + i] for i in range(16)] for s in range(T)]
names = {1: "operand loads issued", 2: "rx_full seen", 3: "summed + credit sent", 4: "gate math, B operand staged",
         5: "dgi / dgh stored", 6: "mma_done seen", 7: "credits ok (bar)", 8: "DSMEM stores issued", 9: "arrivals sent",
         10: "[mma] b_full seen", 11: "[mma] committed"}
print("BPTT cluster kernel B=%d H=%d: marks relative to the worker's step start (cycles), steps 3..T-4 averaged" % (B, H))
for i in sorted(names):
    v = [tl[s][i] - tl[s][0] for s in range(3, T - 3)]
    print("  %-30s avg %8.0f  min %8d  max %8d" % (names[i], sum(v) / len(v), min(v), max(v)))
per = [tl[s + 1][0] - tl[s][0] for s in range(3, T - 3)]
print("  step period                    avg %8.0f cycles" % (sum(per) / len(per)))
