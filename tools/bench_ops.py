#!/usr/bin/env python
"""Per-op / per-module CUDA-event timings of the hot path on one GPU (development aid; prints a table).
    python tools/bench_ops.py [--batch 256] [--engine 0|1] [--precision 0|1]"""
import argparse
import os
import sys
from types import SimpleNamespace as NS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from speech2affective_gestures_b200 import _C, ops  # noqa: E402
from speech2affective_gestures_b200.net import multimodal_context_net_v2 as M  # noqa: E402
from speech2affective_gestures_b200.synthetic import Vocab, synthetic_batch  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--engine", type=int, default=0)
    ap.add_argument("--precision", type=int, default=0)
    a = ap.parse_args()
    lib = _C.lib()
    lib.s2ag_set_engine(a.engine)
    lib.s2ag_set_precision(a.precision)
    dev = torch.device("cuda:0")
    B, T = a.batch, 34
    from speech2affective_gestures_b200.config import namespace as config_namespace
    cfg = config_namespace()
    spk = Vocab("vid", 1370)
    G = M.PoseGenerator(cfg, 27, 20000, 300, None, 71, 37, 34, z_obj=spk).to(dev)
    Tn = M.PoseGeneratorTriModal(cfg, 27, 20000, 300, None, z_obj=spk).to(dev)
    D = M.AffDiscriminator(27).to(dev)
    text, audio, mfcc, target, vid = synthetic_batch(B, dev, seed=5)
    pre = target.new_zeros(B, T, 28)
    pre[:, :4, :-1] = target[:, :4]
    pre[:, :4, -1] = 1
    rows = []

    def add(name, fn, **kw):
        n0 = lib.s2ag_launch_count()
        fn()
        n1 = lib.s2ag_launch_count()
        rows.append((name, timeit(fn, **kw), n1 - n0))

    for net in (G, Tn, D):
        net.train()

    def g_fwd_nograd():
        with torch.no_grad():
            G(pre, text, mfcc, vid)

    def g_fwd_bwd():
        G.zero_grad()
        out, z, mu, lv = G(pre, text, mfcc, vid)
        (out.sum() + mu.sum() + lv.sum()).backward()

    def t_fwd():
        with torch.no_grad():
            Tn(pre, text, audio, vid)

    def d_fwd_nograd():
        with torch.no_grad():
            D(target)

    def d_fwd_bwd():
        D.zero_grad()
        D(target).sum().backward()

    add("G fwd (no grad)", g_fwd_nograd)
    add("G fwd+bwd", g_fwd_bwd)
    add("T fwd (no grad)", t_fwd)
    add("D fwd (no grad)", d_fwd_nograd)
    add("D fwd+bwd", d_fwd_bwd)

    # sub-modules of G
    def sub(name, fn):
        add("  " + name, fn)

    with torch.no_grad():
        sub("MFCCEncoder fwd", lambda: G.audio_encoder(mfcc))
        sub("TextEncoderTCN fwd", lambda: G.text_encoder(text))
        sub("AffEncoder fwd", lambda: G.aff_encoder(pre[:, :, :-1]))
        sub("WavEncoder fwd", lambda: Tn.audio_encoder(audio))
        x88 = torch.randn(B, T, 88, device=dev)
        gp = M._gru_param_list(G.gru)
        if gp is not None:
            sub("G bi-GRU fwd (4 layers)", lambda: ops.bigru(x88, gp, 4, 300, 0.0, False, True))
    if gp is not None:
        xg = torch.randn(B, T, 88, device=dev, requires_grad=True)

        def gru_fb():
            G.zero_grad()
            ops.bigru(xg, gp, 4, 300, 0.0, False, True).sum().backward()
        sub("G bi-GRU fwd+bwd", gru_fb)

    print("batch %d engine %d precision %d" % (B, a.engine, a.precision))
    print("%-34s %10s %9s" % ("op", "ms", "launches"))
    for n, ms, k in rows:
        print("%-34s %10.3f %9d" % (n, ms, k))


if __name__ == "__main__":
    main()
