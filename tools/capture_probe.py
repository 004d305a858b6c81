"""Dev tool: find which part of the GAN step is not CUDA-graph capturable (run on the GPU box)."""
import os, sys, traceback
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_step import make_processor
from speech2affective_gestures_b200 import ops
from speech2affective_gestures_b200.synthetic import synthetic_batch

dev = torch.device("cuda:0")
pr, c = make_processor("tiny", 40, 12, dev)
for n in (pr.s2ag_generator, pr.s2ag_discriminator, pr.trimodal_generator):
    n.train()
    for m in n.modules():
        if isinstance(m, torch.nn.GRU):
            m.dropout = 0.3
B = 8
text, audio, mfcc, target, vid = synthetic_batch(B, dev, 40, 12, 36267, 1)
pre = pr.make_pre_seq(target)
G, D, T = pr.s2ag_generator, pr.s2ag_discriminator, pr.trimodal_generator


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
        print("CAPTURE OK  ", name, flush=True)
    except Exception as e:
        print("CAPTURE FAIL", name, "::", str(e).split("\n")[0][:160], flush=True)
        torch.cuda.synchronize()


def g_fwd_nograd():
    with torch.no_grad():
        G(pre, text, mfcc, vid)

def t_fwd():
    with torch.no_grad():
        T(pre, text, audio, vid)

def d_fwd_bwd():
    D.zero_grad(); o = D(target); o.sum().backward()

def g_fwd_bwd():
    G.zero_grad(); o = G(pre, text, mfcc, vid); (o[0].sum() + o[2].sum() + o[3].sum()).backward()

def adam():
    ops.adam_step(G.flat_params, G.flat_grads, pr.gen_m, pr.gen_v, 1e-4, 0.5, 0.999, 1e-8, pr.gen_step)

def losses():
    with torch.no_grad():
        o = G(pre, text, mfcc, vid)
    ops.gen_loss(o[0], target, o[0], o[1], o[1], o[2], o[3], None, (1, 1, 1, 0), pr.metrics[1:6])
    ops.l1_mean(o[0], target, pr.metrics[6:7])

probe("make_pre_seq", lambda: pr.make_pre_seq(target))
probe("seed_advance", lambda: ops.advance_seed_nonce(dev))
probe("rand argsort", lambda: vid[torch.rand(B, device=dev).argsort()])
probe("randn_like", lambda: torch.randn_like(target))
probe("G fwd no_grad", g_fwd_nograd)
probe("T fwd", t_fwd)
probe("D fwd+bwd", d_fwd_bwd)
probe("G fwd+bwd", g_fwd_bwd)
probe("adam", adam)
probe("losses", losses)
probe("full step", lambda: pr.gan_step_async(text, audio, mfcc, target, vid, True))
