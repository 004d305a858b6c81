"""Shared helpers for the parity tests: builds the nets under test with the oracle's deterministic
weights, and the configurations ('full' = the reference's YAML config, 'tiny' = the same graph with
small widths so that the CPU kernel-logic emulator finishes in seconds)."""
import os
import sys
from types import SimpleNamespace as NS

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import s2ag_oracle as O  # noqa: E402  (test infrastructure)

from speech2affective_gestures_b200.net import multimodal_context_net_v2 as M  # noqa: E402
from speech2affective_gestures_b200.net import embedding_net as men  # noqa: E402
from speech2affective_gestures_b200.synthetic import Vocab  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "s2ag_reference_golden.npz")


def cfg_dict(kind):
    c = dict(O.CFG)
    if kind == "tiny":
        c.update(hidden_size=24, hidden_size_s2eg=24, wordembed_dim=24, n_layers=2)
    return c


def derand(net):
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.GRU):
            m.dropout = 0.0
    return net


def build_nets(kind, n_words, n_spk, dev, seeds=(100, 101, 102, 103)):
    """-> (G, T, D, C) on `dev`, de-randomised, weights = oracle.fill_state_dict(seed)."""
    c = cfg_dict(kind)
    cfg = NS(**c)
    spk = Vocab("vid", n_spk)
    G = M.PoseGenerator(cfg, 27, n_words, c["wordembed_dim"], None, 71, 37, 34, z_obj=spk)
    T = M.PoseGeneratorTriModal(cfg, 27, n_words, c["wordembed_dim"], None, z_obj=spk)
    D = M.AffDiscriminator(27)
    C = M.ConvDiscriminatorTriModal(27)
    nets = []
    for net, seed in zip((G, T, D, C), seeds):
        derand(net)
        sd = net.state_dict()
        O.fill_state_dict(sd, seed)  # in place: state_dict tensors alias the parameters
        nets.append(net.to(dev))
    return nets


def sd_cpu(net):
    """reference-layout state_dict on the CPU for the oracle (aliases preserved)"""
    out, seen = {}, {}
    for k, v in net.state_dict().items():
        key = (v.data_ptr(), tuple(v.shape))
        if key not in seen:
            seen[key] = v.detach().cpu().clone()
        out[k] = seen[key]
    return out


def inject_eps(seq):
    """make the package's re-parametrisation draw from `seq` (cycled), like gen_golden.py does for the reference"""
    state = {"i": 0}

    def src(like):
        e = seq[state["i"] % len(seq)]
        state["i"] += 1
        return e.to(like.device)
    men.eps_source = src
    return state


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(b.abs().max().item(), 1e-9))


def check_grads(named_grads, ref_grads, tol=2e-3, what=""):
    """Gradient parity that is robust to ReLU-mask flips.  Two fp32-grade implementations differ by ~1e-6 in
    the pre-activations, so a handful of the ~1e5-1e6 ReLU/LeakyReLU inputs of a pass land on the other side of 0
    and each flip moves ONE gradient element by O(1) of its size (a conv-bias gradient channel by a few %).
    Criteria (a real kernel bug -- wrong tap, missing term, bf16-only operand -- violates all three):
      * per tensor: at most max(4, 3%) of the elements off by more than tol * (tensor max + 1e-3 global max) -- or, when
        more, all of them confined to at most two output channels of the tensor (one flipped unit behind a
        BatchNorm -> ReLU moves the whole weight-gradient row group of its channel: 135 of the 2160 elements of
        st_gcn1.gcn.conv.weight, whose rows are k*16 + c, share one c);
      * per tensor: relative L2 error <= 10 * tol;
      * all tensors together: relative L2 error <= tol."""
    gmax = max([v.abs().max().item() for v in ref_grads.values()] + [1e-30])
    num = den = 0.0
    bad = []
    for name, g in named_grads.items():
        if name not in ref_grads:
            continue
        r = ref_grads[name].detach().cpu().double()
        d = (g.detach().cpu().double() - r).abs()
        lim = tol * (r.abs().max().item() + 1e-3 * gmax)
        n_off = int((d > lim).sum())
        l2 = d.pow(2).sum().item()
        ref2 = r.pow(2).sum().item()
        num += l2
        den += ref2
        if n_off > max(4, 0.03 * d.numel()):
            rows = (d > lim).nonzero()[:, 0]
            chans = set((rows % 16).tolist()) if name.endswith("gcn.conv.weight") or name.endswith("gcn.conv.bias") \
                else set(rows.tolist())
            if len(chans) > 2:
                bad.append((name, "outliers", n_off, d.numel(), d.max().item(), r.abs().max().item()))
        if l2 ** 0.5 > 10 * tol * (ref2 ** 0.5 + 1e-3 * gmax * d.numel() ** 0.5):
            bad.append((name, "relL2", (l2 / max(ref2, 1e-60)) ** 0.5))
    assert not bad, (what, bad[:8])
    assert (num / max(den, 1e-60)) ** 0.5 <= tol, (what, "global relL2", (num / max(den, 1e-60)) ** 0.5)
