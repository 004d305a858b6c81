#!/usr/bin/env python
"""Development aid: per-phase clock64 timeline of one CTA of the persistent GRU forward kernel (G config)."""
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
for _ in range(1):
    for d in range(2):
        ps += [torch.randn(3 * H, In, generator=g) * 0.05, torch.randn(3 * H, H, generator=g) * 0.05,
               torch.randn(3 * H, generator=g) * 0.05, torch.randn(3 * H, generator=g) * 0.05]
ps = [t.to(dev) for t in ps]
x = torch.randn(B, T, In, generator=g).to(dev)
with torch.no_grad():
    for _ in range(3):
        ops.bigru(x, ps, 1, H, 0.0, False)
    torch.cuda.synchronize()
    lib.s2ag_debug_flags(2)
    ops.bigru(x, ps, 1, H, 0.0, False)
    torch.cuda.synchronize()
    lib.s2ag_debug_flags(0)
NS = 24
buf = (ctypes.c_longlong * (64 * 16 + 64 * 3 * NS + NS * 8))()
lib.s2ag_debug_read_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.s2ag_debug_read_timeline(buf, 64 * 16 + 64 * 3 * NS + NS * 8) == 0
tl = [[buf[s * 16 + i] for i in range(16)] for s in range(T)]
names = {0: "step start", 14: "gi loads issued", 15: "loop end", 12: "flag seen (pre-fence)", 13: "after proxy fence", 1: "after acq_rel fence", 2: "copies issued", 3: "mma_bar done", 4: "tmem read", 5: "stores issued",
         6: "bar.sync done", 7: "flag released", 8: "[mma] tfree", 9: "[mma] ready[0]", 10: "[mma] ready[S-1]", 11: "[mma] committed"}
print("B=%d H=%d: per-step marks relative to step start (cycles), steps 2..T-2 averaged" % (B, H))
acc = {}
for s in range(2, T - 1):
    for i in names:
        if i == 0:
            continue
        acc.setdefault(i, []).append(tl[s][i] - tl[s][0])
for i in sorted(acc):
    v = acc[i]
    print("  %-22s avg %8.0f  min %8d  max %8d" % (names[i], sum(v) / len(v), min(v), max(v)))
per = [tl[s + 1][0] - tl[s][0] for s in range(2, T - 2)]
print("  step period            avg %8.0f cycles" % (sum(per) / len(per)))

S = (H + 15) // 16
st = 10
base = tl[st][0]
off = 64 * 16
print("step %d per-slice (cycles after step start): flag seen / copy issued / landed(in-order)" % st)
for sl in range(S):
    v = [buf[off + ((st * 3 + k) * NS) + sl] - base for k in range(3)]
    print("  slice %2d  %7d %7d %7d" % (sl, v[0], v[1], v[2]))

o3 = 64 * 16 + 64 * 3 * NS
t0 = min(buf[o3 + sl * 8 + 0] for sl in range(S))
print("every slice CTA of group 0 (globaltimer ns after the earliest step-9 release): rel9 | step10: polled+issued, stores done, mma_bar, pre-release, released")
for sl in range(S):
    v = [buf[o3 + sl * 8 + k] - t0 for k in range(6)]
    print("  slice %2d  %6d | %6d %6d %6d %6d %6d" % (sl, *v))
