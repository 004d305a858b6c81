import torch, sys
sys.path.insert(0, '/root/repo')
from speech2affective_gestures_b200.net.multimodal_context_net_v2 import WavEncoder
we = WavEncoder().cuda(); we.train()
a = torch.rand(256, 36267, device='cuda') - 0.5
with torch.no_grad():
    for _ in range(3): we(a)
torch.cuda.synchronize()
