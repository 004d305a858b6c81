#!/usr/bin/env python
"""Development aid: per-phase clock64 timeline of one CTA of the persistent BPTT kernel."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from speech2affective_gestures_b200 import _C, ops  # noqa: E402

B, T, In, H = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 34, 600, int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda:0")
lib = _C.lib()
g = torch.Generator(device="cpu").manual_seed(0)
ps = []
for d in range(2):
    ps += [torch.randn(3 * H, In, generator=g) * 0.05, torch.randn(3 * H, H, generator=g) * 0.05,
           torch.randn(3 * H, generator=g) * 0.05, torch.randn(3 * H, generator=g) * 0.05]
ps = [t.to(dev).requires_grad_(True) for t in ps]
x = torch.randn(B, T, In, generator=g).to(dev).requires_grad_(True)
for it in range(3):
    y = ops.bigru(x, ps, 1, H, 0.0, False)
    if it == 2:
        torch.cuda.synchronize()
        lib.s2ag_debug_flags(8)
    y.sum().backward()
torch.cuda.synchronize()
lib.s2ag_debug_flags(0)
buf = (ctypes.c_longlong * (64 * 16))()
lib.s2ag_debug_read_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.s2ag_debug_read_timeline(buf, 64 * 16) == 0
tl = [[buf[s * 16 + i] for i in range(16)] for s in range(T)]
names = {1: "gate math + A stored", 2: "sync done", 3: "MMA done", 4: "partials stored", 5: "sync done", 6: "arrived",
         7: "dgi/dgh stored", 8: "[t0] others arrived", 9: "sync done", 10: "partials summed"}
print("BPTT B=%d H=%d: marks relative to step start (cycles), steps 2..T-3 averaged" % (B, H))
for i in sorted(names):
    v = [tl[s][i] - tl[s][0] for s in range(2, T - 2)]
    print("  %-22s avg %8.0f  min %8d  max %8d" % (names[i], sum(v) / len(v), min(v), max(v)))
per = [tl[s + 1][0] - tl[s][0] for s in range(2, T - 3)]
print("  step period            avg %8.0f cycles" % (sum(per) / len(per)))
