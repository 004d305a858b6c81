"""GPU checks of the two kernels added beside the generic tcgen05 engine, through the C ABI:
  * gemm_umma_pk_kernel (weight operand packed once per call, fetched by TMA) against the on-the-fly kernel
    (s2ag_debug_flags bit 8 disables the packed route) and an fp64 reference, including ragged N / K and split rows;
  * conv_wgrad_shift_kernel (MN-major shifted-window weight gradient) against the exact-fp32 SIMT engine and an fp64
    reference on every role assignment (X or dY on the M side), tap split and padding case."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from speech2affective_gestures_b200 import _C, ops

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(b.abs().max().item(), 1e-12))


@pytest.mark.parametrize("M,N,K", [(8704, 900, 600), (1024, 300, 150), (640, 88, 900), (515, 27, 37), (2048, 256, 64)])
def test_packed_weight_operand_contraction(M, N, K):
    dev = torch.device("cuda:0")
    lib = _C.lib()
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.1).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    gy = torch.randn(M, N, generator=g).to(dev)
    res = {}
    for name, flags in (("packed", 0), ("on_the_fly", 256)):
        lib.s2ag_debug_flags(flags)
        try:
            xa, wa, ba = (t.clone().requires_grad_(True) for t in (x, w, b))
            y = ops.linear(xa, wa, ba)  # no activation: an fp64 reference would flip ReLU masks of near-zero outputs
            y.backward(gy)
            torch.cuda.synchronize()
            res[name] = (y.detach(), xa.grad, wa.grad, ba.grad)
        finally:
            lib.s2ag_debug_flags(0)
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    yr = F.linear(xd, wd, bd)
    yr.backward(gy.double())
    ref = (yr, xd.grad, wd.grad, bd.grad)
    for i, what in enumerate(("y", "dx", "dw", "db")):
        assert _rel(res["packed"][i], ref[i]) < 2e-5, (what, _rel(res["packed"][i], ref[i]))
        # same bf16x3 arithmetic, same k order: the two routes agree far inside the fp32-grade tolerance
        assert _rel(res["packed"][i], res["on_the_fly"][i]) < 2e-6, (what, _rel(res["packed"][i], res["on_the_fly"][i]))


def test_packed_route_needs_registered_scratch():
    """without a scratch buffer for the stream the same entry point runs the on-the-fly kernel (same result)"""
    dev = torch.device("cuda:0")
    lib = _C.lib()
    x = torch.randn(1024, 96, device=dev); w = torch.randn(64, 96, device=dev); b = torch.zeros(64, device=dev)
    y0 = ops.linear(x, w, b)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        h = ctypes.c_void_p(st.cuda_stream)
        y1 = torch.empty(1024, 64, device=dev)
        st.wait_stream(torch.cuda.default_stream())
        assert lib.s2ag_register_scratch(h, None, 0) == 0  # explicitly none for this stream
        _C.call("s2ag_linear_fwd", x.data_ptr(), 96, w.data_ptr(), b.data_ptr(), y1.data_ptr(), 64, 1024, 64, 96, 0, 0.0, h)
    st.synchronize()
    assert _rel(y1, y0) < 2e-6
    assert lib.s2ag_register_scratch(ctypes.c_void_p(1), ctypes.c_void_p(8), 64) < 0  # misaligned buffer is refused


WGRAD = [  # N, H, W, Cin, Cout, KH, KW, ph, pw
    (3, 34, 9, 3, 16, 1, 1, 0, 0),      # narrower than a tensor-core tile: the GEMM route would be SIMT
    (3, 34, 9, 80, 16, 9, 1, 4, 0),     # X on the M side
    (3, 34, 3, 16, 80, 9, 3, 4, 1),     # dY on the M side, 27 taps
    (2, 34, 9, 16, 16, 3, 3, 1, 1),
    (3, 34, 1, 80, 16, 3, 1, 1, 0),     # Conv1d
    (2, 37, 1, 71, 64, 5, 1, 2, 0),     # ragged channel count
    (2, 34, 1, 64, 192, 1, 1, 0, 0),    # 1x1: k-steps split over accumulator sets
    (2, 20, 5, 8, 24, 3, 2, 0, 0),      # no padding, even kernel width
]


@pytest.mark.parametrize("N,H,W,Cin,Cout,KH,KW,ph,pw", WGRAD)
def test_conv_wgrad_shift_kernel(N, H, W, Cin, Cout, KH, KW, ph, pw):
    dev = torch.device("cuda:0")
    lib = _C.lib()
    st = ops._stream(torch.empty(1, device=dev))
    g = torch.Generator().manual_seed(N * H + Cin * Cout + KH)
    Ho, Wo = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1
    x = torch.randn(N, H, W, Cin, generator=g).to(dev)
    dy = torch.randn(N, Ho, Wo, Cout, generator=g).to(dev)
    dw0 = torch.randn(Cout, Cin, KH, KW, generator=g).to(dev)   # the kernel accumulates into dw

    def run(engine, flags):
        dw = dw0.clone(); db = torch.zeros(Cout, device=dev)
        assert lib.s2ag_set_engine(engine) == 0
        lib.s2ag_debug_flags(flags)
        try:
            _C.call("s2ag_conv_bwd_weight", dy.data_ptr(), Cout, x.data_ptr(), Cin, N, H, W, Cin, dw.data_ptr(), db.data_ptr(),
                    Cout, KH, KW, 1, 1, ph, pw, 1, 1, st)
            torch.cuda.synchronize()
        finally:
            lib.s2ag_set_engine(0); lib.s2ag_debug_flags(0)
        return dw - dw0, db

    dw_shift, db_shift = run(0, 128)   # bit 7: take the shifted-window kernel wherever it is applicable
    dw_simt, db_simt = run(1, 0)
    ref = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (Cout, Cin, KH, KW), dy.double().permute(0, 3, 1, 2),
                                      padding=(ph, pw))
    assert _rel(dw_simt, ref) < 1e-5
    assert _rel(dw_shift, ref) < 2e-5, _rel(dw_shift, ref)
    assert _rel(db_shift, dy.double().sum((0, 1, 2))) < 1e-5


@pytest.mark.parametrize("In,H,B", [(8, 64, 200), (24, 40, 5), (88, 300, 70), (16, 24, 33)])
def test_cluster_gru_kernels_match_l2_exchange_kernels(In, H, B):
    """umma_gru_cluster.cu (thread-block clusters exchanging h / reduce-scattering the BPTT partials through
    distributed shared memory; forward default for H <= 80, s2ag_debug_flags bit 2048 forces it for every size)
    against the L2-exchange kernels of umma_gru.cu (bits 1024 | 8192): forward output and every gradient, ragged clip
    tiles and partial last slices included."""
    dev = torch.device("cuda:0")
    lib = _C.lib()
    T, L = 34, 2
    g = torch.Generator().manual_seed(In + H)
    # (weights ~ 1/sqrt(H) like nn.GRU's init: larger weights make the recurrence chaotic and any two fp32-grade
    # implementations drift apart -- both kernels are then 6e-4 from an fp64 nn.GRU at H = 300)
    sc = 0.2 if H <= 64 else 0.05
    ps0 = [torch.randn(s, generator=g) * sc for l in range(L) for s in
           [(3 * H, In if l == 0 else 2 * H), (3 * H, H), (3 * H,), (3 * H,)] * 2]
    x0 = torch.randn(B, T, In, generator=g)
    gy = torch.randn(B, T, 2 * H, generator=g).to(dev)
    res = {}
    for name, flags in (("l2", 1024 | 8192), ("cluster", 2048)):
        lib.s2ag_debug_flags(flags)
        try:
            ps = [t.clone().to(dev).requires_grad_(True) for t in ps0]
            x = x0.clone().to(dev)
            y = ops.bigru(x, ps, L, H, 0.0, False)
            y.backward(gy)
            torch.cuda.synchronize()
            res[name] = [y.detach()] + [q.grad for q in ps]
        finally:
            lib.s2ag_debug_flags(0)
    for a, b in zip(res["cluster"], res["l2"]):
        assert _rel(a, b) < 2e-5, _rel(a, b)
