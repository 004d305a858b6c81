"""GPU A/B check of the two dense-contraction engines: every module forward/backward of the 'full' configuration
run once on the tcgen05 engine (bf16x3 operand split, the default) and once with s2ag_set_engine(1) (exact-fp32
SIMT kernel everywhere) must agree to fp32-grade tolerance, output by output and gradient by gradient.  Also
checks that the bf16x1 precision mode stays within bf16 error of the fp32-grade result."""
import ctypes

import numpy as np
import pytest
import torch

from common import O, build_nets, inject_eps, rel, check_grads
from speech2affective_gestures_b200 import _C

pytestmark = pytest.mark.gpu
N_WORDS, N_SPK = 64, 25


def _run(net_name, B, engine, precision=0):
    dev = torch.device("cuda:0")
    lib = _C.lib()
    assert lib.s2ag_set_engine(engine) == 0 and lib.s2ag_set_precision(precision) == 0
    try:
        G, T, D, C = build_nets("full", N_WORDS, N_SPK, dev)
        batch, eps_list, _ = O.synthetic_batch(B, N_WORDS, N_SPK, 36267, 11)
        text, audio, mfcc, target, vid = (x.to(dev) for x in batch)
        pre = target.new_zeros(B, 34, 28)
        pre[:, :4, :-1] = target[:, :4]
        pre[:, :4, -1] = 1
        inject_eps([eps_list[0]])
        g = torch.from_numpy(np.random.RandomState(1).normal(size=(B, 34, 27)).astype(np.float32)).to(dev)
        if net_name == "G":
            net = G
            net.train()
            out = net(pre, text, mfcc, vid)[0]
            (out * g).sum().backward()
        elif net_name == "D":
            net = D
            net.train()
            x = target.clone().requires_grad_(True)
            out = net(x)
            (out * g[:, :1, 0]).sum().backward()
        else:
            net = T
            net.train()
            with torch.no_grad():
                out = net(pre, text, audio, vid)[0]
        grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
        torch.cuda.synchronize()
        return out.detach().clone(), grads
    finally:
        lib.s2ag_set_engine(0)
        lib.s2ag_set_precision(0)


@pytest.mark.parametrize("net_name,B", [("G", 3), ("G", 16), ("D", 16), ("T", 8)])
def test_tcgen05_engine_matches_simt_engine(net_name, B):
    out_s, g_s = _run(net_name, B, engine=1)
    out_u, g_u = _run(net_name, B, engine=0)
    assert rel(out_u, out_s) < 2e-5, rel(out_u, out_s)
    if g_s:
        # bf16x3 carries ~2^-17 per operand (4e-6 per contraction); through the 4x34-step BPTT and the cancelling
        # sums of the TCN weight gradients this grows to ~1e-3 on the text-encoder gradients: same 2e-3 bar as the
        # oracle comparisons (robust to ReLU-mask flips, see tests/common.py)
        check_grads(g_u, g_s, tol=2e-3, what=net_name)


def test_bf16x1_mode_is_bf16_close():
    out_3, _ = _run("G", 8, engine=0, precision=0)
    out_1, _ = _run("G", 8, engine=0, precision=1)
    r = rel(out_1, out_3)
    assert 1e-6 < r < 3e-2, r  # really a different (single-pass bf16) computation, and still close


def test_persistent_gru_batch_chunking_matches_per_step_kernels():
    """B = 400 clips at H = 300 needs 4 batch tiles x 19 slices x 2 directions = 152 CTAs > 148: the persistent
    recurrence / BPTT run as two launches (3 + 1 tiles).  Forward output and all gradients must match the per-step
    SIMT kernels (s2ag_set_engine(1))."""
    from speech2affective_gestures_b200 import ops
    dev = torch.device("cuda:0")
    lib = _C.lib()
    B, T, In, H = 400, 34, 24, 300
    g = torch.Generator().manual_seed(5)
    base = [torch.randn(3 * H, In, generator=g) * 0.1, torch.randn(3 * H, H, generator=g) * 0.05,
            torch.randn(3 * H, generator=g) * 0.1, torch.randn(3 * H, generator=g) * 0.1] * 2
    x0 = torch.randn(B, T, In, generator=g)
    gy = torch.randn(B, T, 2 * H, generator=g).to(dev)
    res = []
    for engine in (1, 0):
        assert lib.s2ag_set_engine(engine) == 0
        try:
            ps = [t.clone().to(dev).requires_grad_(True) for t in base]
            x = x0.clone().to(dev).requires_grad_(True)
            y = ops.bigru(x, ps, 1, H, 0.0, False)
            y.backward(gy)
            torch.cuda.synchronize()
            res.append((y.detach().clone(), x.grad.clone(), [p.grad.clone() for p in ps]))
        finally:
            lib.s2ag_set_engine(0)
    (y1, dx1, g1), (y0, dx0, g0) = res
    assert rel(y0, y1) < 2e-5 and rel(dx0, dx1) < 2e-4
    for a, b in zip(g0, g1):
        assert rel(a, b) < 5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,act,accumulate", [(2176, 900, 600, 0, False), (1100, 300, 88, 2, False), (4352, 600, 1800, 0, True)])
def test_two_tma_contraction_matches_packed_kernel(M, N, K, act, accumulate):
    """gemm_umma_tt.cuh (both operands pre-packed and fetched by TMA, 256-row tiles, three MMA issuer threads; opt-in
    through s2ag_debug_flags bit 16384) against the default packed-B kernel and fp64: linear forward (store epilogue,
    ragged last row tile / column tile) and the accumulating data-gradient epilogue"""
    import torch
    from speech2affective_gestures_b200 import _C, ops
    dev = torch.device("cuda:0")
    if _C.is_emulated():
        _C._lib, _C._emulated = None, False
    torch.manual_seed(5)
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    st = ops._stream(x)
    res = {}
    for flag in (0, 16384):
        _C.lib().s2ag_debug_flags(flag)
        try:
            if accumulate:   # dx[M,K'] += dy[M,N'] @ w[N',K']  with (N', K') = (K, N) of the forward naming
                dx = torch.ones(M, N, device=dev)
                _C.call("s2ag_linear_bwd_data", ops._p(x), K, ops._p(w.t().contiguous()), ops._p(dx), N, M, K, N, 1, st)
                res[flag] = dx
            else:
                y = torch.empty(M, N, device=dev)
                _C.call("s2ag_linear_fwd", ops._p(x), K, ops._p(w), ops._p(b), ops._p(y), N, M, N, K, act, 0.3, st)
                res[flag] = y
        finally:
            _C.lib().s2ag_debug_flags(0)
    if accumulate:
        want = 1.0 + x.double() @ w.double().t()
    else:
        want = x.double() @ w.double().t() + b.double()
        if act == 2:
            want = torch.where(want > 0, want, 0.3 * want)
    scale = want.abs().max().item()
    assert (res[16384].double() - want).abs().max().item() <= 2e-5 * scale
    assert (res[16384] - res[0]).abs().max().item() <= 2e-5 * scale
