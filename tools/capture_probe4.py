"""Dev tool: localise the first op that sees an invalidated capture (S2AG_DEBUG_CAPTURE=1)."""
import os, sys
os.environ["S2AG_DEBUG_CAPTURE"] = "1"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
from speech2affective_gestures_b200 import ops, _C
dev = torch.device("cuda:0")
P = lambda *s: (torch.randn(*s, device=dev) * 0.3).requires_grad_(True)
x = P(8, 34, 40)
conv = nn.Conv1d(40, 24, 3, padding=1).to(dev)
which = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "global"
orig_call = _C.call
def traced(name, *a):
    st0 = _C.lib().s2ag_stream_capture_status(a[-1])
    orig_call(name, *a)
    st1 = _C.lib().s2ag_stream_capture_status(a[-1])
    if st0 or st1:
        import threading
        print("   %-26s status %d -> %d  stream=%s thread=%s" % (name, st0, st1, a[-1], threading.current_thread().name), flush=True)
_C.call = traced
ops._C.call = traced
if which == "conv_nobn":
    fn = lambda: ops.conv_bn_act(x, conv.weight, conv.bias, (1, 1, 1, 0, 1, 1), None, 2, 0.3).sum().backward()
elif which == "conv_nobias_grad":
    conv.bias.requires_grad_(False)
    fn = lambda: ops.conv_bn_act(x, conv.weight, conv.bias, (1, 1, 1, 0, 1, 1), None, 2, 0.3).sum().backward()
elif which == "conv_noxgrad":
    x.requires_grad_(False)
    fn = lambda: ops.conv_bn_act(x, conv.weight, conv.bias, (1, 1, 1, 0, 1, 1), None, 2, 0.3).sum().backward()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    fn(); fn()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
print("--- capture", which, mode, flush=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, capture_error_mode=mode):
        fn()
    g.replay(); torch.cuda.synchronize(); print("OK")
except Exception as e:
    print("FAIL", str(e).split("\n")[0][:150])
