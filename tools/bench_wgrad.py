#!/usr/bin/env python
"""Development aid: conv weight-gradient shapes of the G/D nets at 256 clips, shifted-window kernel vs implicit GEMM."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from speech2affective_gestures_b200 import _C  # noqa: E402

dev = torch.device("cuda:0")
lib = _C.lib()
st = torch.cuda.current_stream().cuda_stream
# (N, H, W, Cin, Cout, KH, KW)
SHAPES = [(256, 34, 9, 3, 16, 1, 1), (256, 34, 9, 80, 16, 9, 1), (256, 34, 3, 48, 80, 9, 1), (256, 34, 3, 48, 16, 9, 1),
          (256, 34, 9, 3, 80, 9, 1), (256, 68, 1, 64, 192, 1, 1), (256, 34, 1, 64, 192, 1, 1), (256, 34, 1, 80, 16, 3, 1),
          (256, 34, 3, 16, 16, 3, 1), (256, 34, 1, 16, 8, 3, 1), (256, 37, 1, 71, 64, 5, 1), (256, 37, 1, 64, 64, 5, 1)]


def run(shape, force_gemm):
    N, H, W, Cin, Cout, KH, KW = shape
    ph, pw = (KH - 1) // 2, (KW - 1) // 2
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(N, H, W, Cin, generator=g).to(dev)
    dy = torch.randn(N, H, W, Cout, generator=g).to(dev)
    dw = torch.zeros(Cout, Cin, KH, KW, device=dev)
    lib.s2ag_debug_flags(force_gemm)

    def call():
        _C.call("s2ag_conv_bwd_weight", dy.data_ptr(), Cout, x.data_ptr(), Cin, N, H, W, Cin, dw.data_ptr(), None, Cout,
                KH, KW, 1, 1, ph, pw, 1, 1, st)
    call()
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (Cout, Cin, KH, KW), dy.permute(0, 3, 1, 2), padding=(ph, pw))
    err = ((dw - ref).norm() / ref.norm()).item()
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    lib.s2ag_debug_flags(0)
    return e0.elapsed_time(e1) * 100.0, err


print("%-34s %12s %12s   rel.err (shift / gemm)" % ("N,H,W,Cin,Cout,KH,KW", "shift us", "gemm us"))
if len(sys.argv) > 1:
    SHAPES = [SHAPES[int(sys.argv[1])]]
for s in SHAPES:
    a, ea = run(s, 128)
    b, eb = run(s, 4)
    c, ec = run(s, 64 + 128)
    print("%-34s %12.1f %12.1f   %.2e / %.2e   (no atomics: %.1f us)" % (str(s), a, b, ea, eb, c))
