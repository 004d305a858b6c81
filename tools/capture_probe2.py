"""Dev tool: per-op CUDA-graph capturability of forward+backward (run on the GPU box)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
from speech2affective_gestures_b200 import ops
dev = torch.device("cuda:0")
P = lambda *s: (torch.randn(*s, device=dev) * 0.3).requires_grad_(True)

def probe(name, fn, mode="global"):
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("OK  ", mode, name, flush=True)
    except Exception as e:
        print("FAIL", mode, name, "::", str(e).split("\n")[0][:140], flush=True)
        try: torch.cuda.synchronize()
        except Exception: pass

x = P(8, 34, 40); w = P(24, 40); b = P(24)
bn = nn.BatchNorm1d(24).to(dev)
conv = nn.Conv1d(40, 24, 3, padding=1).to(dev)
gru = nn.GRU(24, 16, num_layers=2, bidirectional=True, batch_first=True).to(dev)
gp = []
for l in range(2):
    for sfx in ("", "_reverse"):
        gp += [getattr(gru, "weight_ih_l%d%s" % (l, sfx)), getattr(gru, "weight_hh_l%d%s" % (l, sfx)),
               getattr(gru, "bias_ih_l%d%s" % (l, sfx)), getattr(gru, "bias_hh_l%d%s" % (l, sfx))]
x24 = P(8, 34, 24)
tcnp = [P(24, 24, 2), P(24, 1, 1), P(24), P(24, 24, 2), P(24, 1, 1), P(24)]
table = P(50, 24); idx = torch.randint(0, 50, (8, 34), device=dev)
A = torch.rand(5, 9, 9, device=dev); xg = P(8, 34, 9, 80)
hw = [P(1, 16), P(1), P(1, 34), P(1)]

tests = {
 "linear fwd": lambda: ops.linear(x, w, b, 2, 0.3),
 "linear f+b": lambda: ops.linear(x, w, b, 2, 0.3).sum().backward(),
 "torch f+b (control)": lambda: torch.nn.functional.linear(x, w, b).sum().backward(),
 "bn f+b": lambda: ops.bn_act(x24, bn, 2, 0.3).sum().backward(),
 "conv_bn_act f+b": lambda: ops.conv_bn_act(x, conv.weight, conv.bias, (1, 1, 1, 0, 1, 1), bn, 2, 0.3).sum().backward(),
 "graph f+b": lambda: ops.graph_contract(xg, A).sum().backward(),
 "tcn f+b": lambda: ops.tcn_block(x24, *tcnp, 2, 0.3, True).sum().backward(),
 "embedding f+b": lambda: ops.embedding(idx, table, 0.1).sum().backward(),
 "bigru f+b": lambda: ops.bigru(x24, gp, 2, 16, 0.3, True).sum().backward(),
 "dhead f+b": lambda: ops.dhead(ops.bigru(x24, gp, 2, 16, 0.0, False), *hw).sum().backward(),
}
for mode in ("global", "thread_local"):
    for k, f in tests.items():
        probe(k, f, mode)
