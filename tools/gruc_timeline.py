#!/usr/bin/env python
"""Development aid: clock64 timeline of one CTA (cluster rank 0, tile 0, direction 0) of the cluster GRU forward
kernel (csrc/umma_gru_cluster.cu).  usage: python tools/gruc_timeline.py [clips] [H]"""
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
ps = [t.to(dev) for t in ps]
x = torch.randn(B, T, In, generator=g).to(dev)
lib.s2ag_debug_flags(2048)
with torch.no_grad():
    for _ in range(3):
        ops.bigru(x, ps, 1, H, 0.0, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.bigru(x, ps, 1, H, 0.0, False)
    e1.record()
    torch.cuda.synchronize()
    print("layer forward (projection + recurrence): %.1f us" % (e0.elapsed_time(e1) * 100))
    lib.s2ag_debug_flags(2 | 2048 | (4096 if os.environ.get("GRUC_WAITALL") else 0))
    ops.bigru(x, ps, 1, H, 0.0, False)
    torch.cuda.synchronize()
    lib.s2ag_debug_flags(0)
buf = (ctypes.c_longlong * (64 * 16))()
lib.s2ag_debug_read_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.s2ag_debug_read_timeline(buf, 64 * 16) == 0
tl = [[buf[s * 16 + i] for i in range(16)] for s in range(T)]
names = {1: "gi loads issued", 2: "mma_done seen", 3: "tmem loaded", 4: "gate math done", 5: "staged + fence",
         6: "bar.sync passed", 7: "credits obtained", 8: "copies issued", 9: "loop end (stores issued)",
         10: "[mma] tfree", 11: "[mma] first slice ready", 12: "[mma] last slice ready", 13: "[mma] committed",
         14: "[mma] done observed", 15: "[mma] credits sent"}
print("B=%d H=%d: marks relative to the worker's step start (cycles), steps 3..T-3 averaged" % (B, H))
for i in sorted(names):
    v = [tl[s][i] - tl[s][0] for s in range(3, T - 2)]
    print("  %-26s avg %8.0f  min %8d  max %8d" % (names[i], sum(v) / len(v), min(v), max(v)))
per = [tl[s + 1][0] - tl[s][0] for s in range(3, T - 2)]
print("  step period                avg %8.0f cycles" % (sum(per) / len(per)))
print("step 10: slice i observed ready (cycles after the worker's start of step 9's copy issue [mark 8 of step 9]):")
print("  ", [buf[40 * 16 + i] - tl[9][8] for i in range(8)])
print("   next step start at", tl[10][0] - tl[9][8])
print("per-step period (cycles), all steps:", [tl[s + 1][0] - tl[s][0] for s in range(T - 1)])
print("worker marks of step 0 relative to its start:", [tl[0][i] - tl[0][0] for i in range(1, 10)])
print("worker marks of the last step:", [tl[T - 1][i] - tl[T - 1][0] for i in range(1, 10)])
