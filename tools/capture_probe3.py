"""Dev tool: direct C-ABI calls under CUDA-graph capture (no autograd), run on the GPU box."""
import os, sys, ctypes
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from speech2affective_gestures_b200 import ops, _C
dev = torch.device("cuda:0")
R = lambda *s: torch.randn(*s, device=dev) * 0.3
p = ops._p

def probe(name, fn):
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("OK  ", name, flush=True)
    except Exception as e:
        print("FAIL", name, "::", str(e).split("\n")[0][:140], flush=True)
        try: torch.cuda.synchronize()
        except Exception: pass

N, L, Ci, Co, k = 8, 34, 40, 24, 3
x = R(N, L, Ci); w = R(Co, Ci, k); dy = R(N, L, Co); dw = torch.zeros(Co, Ci, k, device=dev); db = torch.zeros(Co, device=dev)
dx = torch.zeros(N, L, Ci, device=dev); y = torch.zeros(N, L, Co, device=dev); bias = R(Co)
st = lambda: ops._stream(x)
probe("conv_fwd", lambda: _C.call("s2ag_conv_fwd", p(x), Ci, N, L, 1, Ci, p(w), p(bias), p(y), Co, Co, k, 1, 1, 1, 1, 0, 1, 1, 0, 0.0, st()))
probe("conv_bwd_weight(no bias)", lambda: _C.call("s2ag_conv_bwd_weight", p(dy), Co, p(x), Ci, N, L, 1, Ci, p(dw), None, Co, k, 1, 1, 1, 1, 0, 1, 1, st()))
probe("conv_bwd_weight(+bias)", lambda: _C.call("s2ag_conv_bwd_weight", p(dy), Co, p(x), Ci, N, L, 1, Ci, p(dw), p(db), Co, k, 1, 1, 1, 1, 0, 1, 1, st()))
probe("conv_bwd_data", lambda: _C.call("s2ag_conv_bwd_data", p(dy), Co, N, L, 1, Ci, p(w), p(dx), Ci, Co, k, 1, 1, 0, 1, 1, 0, st()))
M, Nn, K = 272, 24, 40
a = R(M, K); g = R(M, Nn); dwl = torch.zeros(Nn, K, device=dev); dbl = torch.zeros(Nn, device=dev)
probe("linear_bwd_weight", lambda: _C.call("s2ag_linear_bwd_weight", p(g), Nn, p(a), K, p(dwl), p(dbl), M, Nn, K, st()))
C = 24; B, T = 8, 34
xt = R(B, T, C); w1 = R(C, 2, C); w2 = R(C, 2, C); b1 = R(C); b2 = R(C)
y1 = torch.zeros(B, T, C, device=dev); y2 = torch.zeros_like(y1); out = torch.zeros_like(y1); dout = R(B, T, C)
dxt = torch.zeros_like(y1); dw1 = torch.zeros(C, 2, C, device=dev); dw2 = torch.zeros_like(dw1); db1 = torch.zeros(C, device=dev); db2 = torch.zeros(C, device=dev)
ws = torch.zeros(2 * B * T * C, device=dev)
probe("tcn_fwd", lambda: _C.call("s2ag_tcn_block_fwd", p(xt), p(w1), p(b1), p(w2), p(b2), p(y1), p(y2), p(out), B, T, C, 2, 0.0, ctypes.c_uint64(1), None, st()))
probe("tcn_bwd", lambda: _C.call("s2ag_tcn_block_bwd", p(dout), p(xt), p(y1), p(y2), p(out), p(w1), p(w2), p(dxt), p(dw1), p(db1), p(dw2), p(db2), p(ws), B, T, C, 2, 0.0, st()))
