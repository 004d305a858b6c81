"""ncu target: one fused TCN residual block forward (256 clips x 34 frames x 300 channels, d = 2), no-grad and grad."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech2affective_gestures_b200.net.tcn import TemporalBlock
tb = TemporalBlock(300, 300, 2, 1, 2, 2, dropout=0.0).cuda()
x = torch.randn(256, 34, 300, device='cuda')
with torch.no_grad():
    for _ in range(3):
        tb.forward_cl(x)
torch.cuda.synchronize()
