"""Per-kernel parity: every C-ABI entry point against the torch.nn op it replaces (the reference's
operator call sites), forward and backward, fp32, tolerance 2e-4 relative-to-max
(well inside north_star's 1e-3)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from speech2affective_gestures_b200 import ops

TOL = 2e-4


def close(a, b, tol=TOL, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = (a - b).abs().max().item()
    ref = max(b.abs().max().item(), 1e-6)
    assert err <= tol * ref + 1e-6, "%s: max err %.3e vs ref max %.3e" % (what, err, ref)


def P(t, dev, grad=True):
    return t.clone().to(dev).requires_grad_(grad)


@pytest.mark.parametrize("M,N,K,act", [(70, 45, 37, 2), (3, 150, 300, 0), (130, 27, 150, 1)])
def test_linear(dev, M, N, K, act):
    torch.manual_seed(0)
    x, w, b = torch.randn(M, K), torch.randn(N, K) * 0.1, torch.randn(N)
    xr, wr, br = P(x, "cpu"), P(w, "cpu"), P(b, "cpu")
    ref = F.linear(xr, wr, br)
    ref = [ref, F.relu(ref), F.leaky_relu(ref, 0.3)][act]
    g = torch.randn_like(ref)
    ref.backward(g)
    xd, wd, bd = P(x, dev), P(w, dev), P(b, dev)
    y = ops.linear(xd, wd, bd, act, 0.3)
    y.backward(g.to(dev))
    close(y, ref, what="y"); close(xd.grad, xr.grad, what="dx"); close(wd.grad, wr.grad, what="dw")
    close(bd.grad, br.grad, what="db")


def test_linear_into_slice(dev):
    torch.manual_seed(1)
    B, T = 3, 5
    x, w, b = torch.randn(B, T, 20), torch.randn(8, 20), torch.randn(8)
    buf = torch.zeros(B, T, 24, device=dev)
    xd, wd, bd = P(x, dev), P(w, dev), P(b, dev)
    y = ops.linear(xd, wd, bd, out=ops.col_slice(buf, 4, 12))
    ref = F.linear(x, w, b)
    close(buf[:, :, 4:12], ref)
    assert buf[:, :, :4].abs().max() == 0 and buf[:, :, 12:].abs().max() == 0
    g = torch.randn(B, T, 24)
    y.backward(g.to(dev)[:, :, 4:12])
    close(xd.grad, g[:, :, 4:12] @ w, what="dx")


def test_linear_t(dev):
    torch.manual_seed(21)
    B, L, C, N = 5, 37, 34, 32
    x, w, b = torch.randn(B, L, C), torch.randn(N, L), torch.randn(N)
    xr, wr, br = P(x, "cpu"), P(w, "cpu"), P(b, "cpu")
    ref = F.leaky_relu(F.linear(xr.transpose(1, 2), wr, br), 0.3)
    g = torch.randn_like(ref)
    ref.backward(g)
    xd, wd, bd = P(x, dev), P(w, dev), P(b, dev)
    buf = torch.zeros(B, C, 40, device=dev)
    y = ops.linear_t(xd, wd, bd, ops.ACT_LEAKY, 0.3, out=ops.col_slice(buf, 8, 40))
    y.backward(g.to(dev))
    close(y, ref); close(xd.grad, xr.grad, what="dx"); close(wd.grad, wr.grad, what="dw"); close(bd.grad, br.grad)


CONV1D = [  # L, Cin, Cout, k, s, p, d
    (37, 71, 64, 5, 1, 2, 1), (37, 48, 34, 3, 1, 1, 1), (200, 1, 16, 15, 5, 30, 1), (120, 16, 32, 15, 6, 0, 1),
    (34, 27, 16, 3, 1, 0, 1),
]


@pytest.mark.parametrize("L,Cin,Cout,k,s,p,d", CONV1D)
@pytest.mark.parametrize("with_bn", [False, True])
def test_conv1d_bn_act(dev, L, Cin, Cout, k, s, p, d, with_bn):
    torch.manual_seed(2)
    N = 3
    conv = nn.Conv1d(Cin, Cout, k, stride=s, padding=p, dilation=d)
    bn = nn.BatchNorm1d(Cout)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
    x = torch.randn(N, L, Cin)
    xr = P(x, "cpu")
    c = conv(xr.transpose(1, 2))
    ref = F.leaky_relu(bn(c) if with_bn else c, 0.3).transpose(1, 2)
    g = torch.randn_like(ref)
    ref.backward(g)
    import copy
    conv2, bn2 = copy.deepcopy(conv).to(dev), nn.BatchNorm1d(Cout).to(dev)
    with torch.no_grad():
        bn2.weight.copy_(bn.weight); bn2.bias.copy_(bn.bias)
    for q in list(conv2.parameters()) + list(bn2.parameters()):
        q.grad = None
    xd = P(x, dev, grad=(s == 1))
    y = ops.conv_bn_act(xd, conv2.weight, conv2.bias, (s, 1, p, 0, d, 1), bn2 if with_bn else None, ops.ACT_LEAKY, 0.3)
    y.backward(g.to(dev))
    close(y, ref, what="y")
    close(conv2.weight.grad, conv.weight.grad, what="dw")
    if s == 1:
        close(xd.grad, xr.grad, what="dx")
    if with_bn:
        close(bn2.weight.grad, bn.weight.grad, what="dgamma"); close(bn2.bias.grad, bn.bias.grad, what="dbeta")
        close(bn2.running_mean, bn.running_mean, what="rm"); close(bn2.running_var, bn.running_var, what="rv")
    else:
        close(conv2.bias.grad, conv.bias.grad, what="db")


@pytest.mark.parametrize("Cin,Cout,KH,KW,V", [(3, 80, 9, 1, 9), (16, 16, 9, 5, 9), (48, 16, 1, 1, 3), (16, 16, 9, 3, 3)])
def test_conv2d(dev, Cin, Cout, KH, KW, V):
    torch.manual_seed(3)
    N, T = 2, 34
    conv = nn.Conv2d(Cin, Cout, (KH, KW), padding=((KH - 1) // 2, (KW - 1) // 2))
    x = torch.randn(N, T, V, Cin)
    xr = P(x, "cpu")
    ref = conv(xr.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    g = torch.randn_like(ref)
    ref.backward(g)
    import copy
    c2 = copy.deepcopy(conv).to(dev)
    for q in c2.parameters():
        q.grad = None
    xd = P(x, dev)
    y = ops.conv_bn_act(xd, c2.weight, c2.bias, (1, 1, (KH - 1) // 2, (KW - 1) // 2, 1, 1))
    y.backward(g.to(dev))
    close(y, ref); close(xd.grad, xr.grad, what="dx"); close(c2.weight.grad, conv.weight.grad, what="dw")
    close(c2.bias.grad, conv.bias.grad, what="db")


@pytest.mark.parametrize("training", [True, False])
def test_bn_maps_add(dev, training):
    torch.manual_seed(4)
    M, C = 333, 48
    bn = nn.BatchNorm1d(C)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2)
    bn.train(training)
    import copy
    bn2 = copy.deepcopy(bn).to(dev)
    x, add = torch.randn(M, C) * 2 + 1, torch.randn(M, C)
    perm = torch.randperm(C)
    cmap = torch.randperm(C)
    # reference: column c uses parameter perm[c] and lands at output column cmap[c]
    xr, ar = P(x, "cpu"), P(add, "cpu")
    xin = torch.empty(M, C).index_copy(1, perm, xr)  # xin[:, perm[c]] = x[:, c]
    yb = bn(xin.unsqueeze(-1)).squeeze(-1)[:, perm]  # back to column order c
    ref = torch.zeros(M, C).index_copy(1, cmap, yb) + ar
    ref = F.leaky_relu(ref, 0.01)
    g = torch.randn(M, C)
    ref.backward(g)
    xd, ad = P(x, dev), P(add, dev)
    y = ops.bn_act(xd, bn2, ops.ACT_LEAKY, 0.01, add=ad, cmap=cmap.int().to(dev), pmap=perm.int().to(dev))
    y.backward(g.to(dev))
    close(y, ref, what="y"); close(xd.grad, xr.grad, what="dx"); close(ad.grad, ar.grad, what="dadd")
    close(bn2.weight.grad, bn.weight.grad, what="dgamma"); close(bn2.bias.grad, bn.bias.grad, what="dbeta")
    close(bn2.running_mean, bn.running_mean, what="rm"); close(bn2.running_var, bn.running_var, what="rv")


def test_bn_groups_equal_successive_calls(dev):
    """ops.bn_groups(2) over two stacked batches == two successive calls of the same BatchNorm (processor_v2.py:808-809:
    D(target) then D(out)): outputs, input / parameter gradients, running statistics and batch counter"""
    torch.manual_seed(14)
    M, C = 200, 24
    bn = nn.BatchNorm1d(C)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2)
    import copy
    bn2 = copy.deepcopy(bn).to(dev)
    xa, xb = torch.randn(M, C) * 2 + 1, torch.randn(M, C) * 0.5 - 3
    ga, gb = torch.randn(M, C), torch.randn(M, C)
    ra, rb = P(xa, "cpu"), P(xb, "cpu")
    ya = F.leaky_relu(bn(ra.unsqueeze(-1)).squeeze(-1), 0.3)
    yb = F.leaky_relu(bn(rb.unsqueeze(-1)).squeeze(-1), 0.3)
    torch.autograd.backward([ya, yb], [ga, gb])
    xs = P(torch.cat([xa, xb]), dev)
    with ops.bn_groups(2):
        y = ops.bn_act(xs, bn2, ops.ACT_LEAKY, 0.3)
    y.backward(torch.cat([ga, gb]).to(dev))
    close(y, torch.cat([ya, yb]), what="y")
    close(xs.grad, torch.cat([ra.grad, rb.grad]), what="dx")
    close(bn2.weight.grad, bn.weight.grad, what="dgamma"); close(bn2.bias.grad, bn.bias.grad, what="dbeta")
    close(bn2.running_mean, bn.running_mean, what="rm"); close(bn2.running_var, bn.running_var, what="rv")
    assert bn2._s2ag_batches == 2


def test_bn_large_mean(dev):
    """statistics must survive |mean| >> std (shifted accumulation)"""
    torch.manual_seed(5)
    M, C = 4000, 8
    x = torch.randn(M, C) * 0.01 + 100.0
    bn = nn.BatchNorm1d(C).to(dev)
    y = ops.bn_act(x.to(dev), bn)
    ref = F.batch_norm(x.double(), None, None, training=True).float()
    close(y, ref, tol=5e-3)


@pytest.mark.parametrize("M,C,groups,act,with_add,train", [(1000, 16, 1, 2, True, True), (777, 48, 1, 1, False, True),
                                                            (1024, 144, 2, 0, False, True), (513, 4, 1, 2, True, True),
                                                            (300, 64, 1, 2, False, False), (4001, 8, 1, 0, False, True)])
def test_bn_float4_vs_scalar_and_torch(dev, M, C, groups, act, with_add, train):
    """the float4 BatchNorm kernels (4 channels per thread) against the scalar kernels (s2ag_debug_bn_flags bit 0) and
    against F.batch_norm in fp64: output, dx, d(add), dgamma, dbeta, running statistics; ragged row counts, statistics
    groups (D(real) / D(fake) in one pass), eval mode"""
    from speech2affective_gestures_b200 import _C
    torch.manual_seed(90 + C)
    x = torch.randn(M, C) * 1.5 + 0.7
    add = torch.randn(M, C) if with_add else None
    g = torch.randn(M, C)
    slope = 0.2
    res = {}
    for flags in (0, 1):
        _C.lib().s2ag_debug_bn_flags(flags)
        try:
            bn = nn.BatchNorm1d(C)
            torch.manual_seed(7)
            with torch.no_grad():
                bn.weight.copy_(torch.rand(C) + 0.5); bn.bias.copy_(torch.rand(C) - 0.5)
                bn.running_mean.copy_(torch.rand(C) * 0.4 - 0.2); bn.running_var.copy_(torch.rand(C) + 0.5)
            bn = bn.to(dev).train(train)
            xd = P(x, dev)
            ad = P(add, dev) if with_add else None
            ctxm = ops.bn_groups(groups) if groups > 1 else __import__("contextlib").nullcontext()
            with ctxm:
                y = ops.bn_act(xd, bn, act, slope, add=ad)
                y.backward(g.to(dev))
            res[flags] = [t.detach().cpu().clone() for t in (y, xd.grad, bn.weight.grad, bn.bias.grad, bn.running_mean,
                                                              bn.running_var)] + ([ad.grad.cpu().clone()] if with_add else [])
        finally:
            _C.lib().s2ag_debug_bn_flags(0)
    for a, b, what in zip(res[0], res[1], ("y", "dx", "dgamma", "dbeta", "rmean", "rvar", "dadd")):
        close(a, b, tol=2e-5, what="float4 vs scalar " + what)
    # torch reference (fp64), one statistics group at a time
    xr = x.double().requires_grad_(True)
    outs = []
    w64, b64 = res[0][2] * 0, None
    bnr = nn.BatchNorm1d(C).double()
    with torch.no_grad():
        torch.manual_seed(7)
        bnr.weight.copy_(torch.rand(C) + 0.5); bnr.bias.copy_(torch.rand(C) - 0.5)
    bnr.train(train)
    if not train:
        with torch.no_grad():
            bnr.running_mean.copy_(res[0][4].double()); bnr.running_var.copy_(res[0][5].double())
    Mg = M // groups
    for q in range(groups):
        v = bnr(xr[q * Mg:(q + 1) * Mg])
        if with_add:
            v = v + add[q * Mg:(q + 1) * Mg].double()
        outs.append(v if act == 0 else (F.relu(v) if act == 1 else F.leaky_relu(v, slope)))
    ref = torch.cat(outs)
    ref.backward(g.double())
    close(res[0][0], ref, tol=1e-4, what="y vs torch")
    close(res[0][1], xr.grad, tol=1e-3, what="dx vs torch")
    close(res[0][2], bnr.weight.grad, tol=1e-3, what="dgamma vs torch")


def test_graph_contract(dev):
    torch.manual_seed(6)
    N, T, V, K, C = 2, 34, 9, 5, 16
    x, A = torch.randn(N, T, V, K * C), torch.rand(K, V, V)
    xr = P(x, "cpu")
    xx = xr.view(N, T, V, K, C).permute(0, 3, 4, 1, 2)  # n k c t v
    ref = torch.einsum("nkctv,kvw->nctw", xx, A).permute(0, 2, 3, 1)
    g = torch.randn_like(ref)
    ref.backward(g)
    xd = P(x, dev)
    y = ops.graph_contract(xd, A.to(dev))
    y.backward(g.to(dev))
    close(y, ref); close(xd.grad, xr.grad, what="dx")


@pytest.mark.parametrize("window", [False, True])
@pytest.mark.parametrize("V,Cin,B", [(9, 3, 2), (3, 48, 2), (9, 3, 33)])
def test_gcn_composed_conv(dev, V, Cin, B, window, monkeypatch):
    """ConvTemporalGraphical as one composed temporal convolution (ops.GcnFn) vs Conv2d((9,1)) + einsum of the reference
    (net/utils/tgcn.py:52-69): output, input gradient, gradients of the ORIGINAL conv parameters (accumulated: a second
    backward doubles them)"""
    monkeypatch.setattr(ops, "GCN_WGRAD_WINDOW", [window])   # both weight-gradient routes (ops.GcnFn.backward)
    torch.manual_seed(60 + V)
    T, K, C = 34, 5, 16
    conv = nn.Conv2d(Cin, K * C, (9, 1), padding=(4, 0))
    A = torch.rand(K, V, V)
    A[-1] = 0   # the body-part graph's last partition is all-zero
    x = torch.randn(B, T, V, Cin)
    xr = P(x, "cpu")
    yr = conv(xr.permute(0, 3, 1, 2))                     # n (k c) t v
    n, kc, t, v = yr.shape
    ref = torch.einsum("nkctv,kvw->nctw", yr.view(n, K, C, t, v), A).permute(0, 2, 3, 1)   # n t w c
    g = torch.randn_like(ref)
    ref.backward(g)
    w, b = P(conv.weight.detach(), dev), P(conv.bias.detach(), dev)
    xd = P(x, dev)
    for rep in (1, 2):
        xd.grad = None
        y = ops.gcn_conv(xd, w, b, A.to(dev), 4)
        y.backward(g.to(dev))
        close(y, ref); close(xd.grad, xr.grad, what="dx")
        close(w.grad, rep * conv.weight.grad, what="dW"); close(b.grad, rep * conv.bias.grad, what="db")


@pytest.mark.parametrize("d", [1, 2, 4, 8])
def test_tcn_block(dev, d):
    from torch.nn.utils import weight_norm
    torch.manual_seed(7)
    B, T, C = 3, 34, 40
    c1 = weight_norm(nn.Conv1d(C, C, 2, padding=d, dilation=d))
    c2 = weight_norm(nn.Conv1d(C, C, 2, padding=d, dilation=d))
    with torch.no_grad():
        for c in (c1, c2):
            c.weight_v.normal_(0, 0.2); c.weight_g.uniform_(0.5, 2.0)
    x = torch.randn(B, T, C)
    xr = P(x, "cpu")
    xt = xr.transpose(1, 2)
    y1 = F.relu(c1(xt)[:, :, :-d])
    y2 = F.relu(c2(y1)[:, :, :-d])
    ref = F.relu(y2 + xt).transpose(1, 2)
    g = torch.randn_like(ref)
    ref.backward(g)
    ps = [P(t.detach(), dev) for t in (c1.weight_v, c1.weight_g, c1.bias, c2.weight_v, c2.weight_g, c2.bias)]
    xd = P(x, dev)
    y = ops.tcn_block(xd, *ps, d, 0.3, training=False)
    y.backward(g.to(dev))
    close(y, ref); close(xd.grad, xr.grad, what="dx")
    for q, r, nm in zip(ps, (c1.weight_v, c1.weight_g, c1.bias, c2.weight_v, c2.weight_g, c2.bias),
                        ("v1", "g1", "b1", "v2", "g2", "b2")):
        close(q.grad, r.grad, what=nm)


def test_tcn_block_dropout_consistency(dev):
    """train-mode dropout: backward must use the same mask as forward (finite-difference-free check:
    d(sum out)/dx through the kept units only, compared against torch autograd on the saved mask)."""
    torch.manual_seed(8)
    B, T, C, d = 2, 34, 16, 2
    ps = [torch.randn(C, C, 2) * 0.3, torch.rand(C, 1, 1) + 0.5, torch.randn(C) * 0.1,
          torch.randn(C, C, 2) * 0.3, torch.rand(C, 1, 1) + 0.5, torch.randn(C) * 0.1]
    ps = [P(t, dev) for t in ps]
    xd = P(torch.randn(B, T, C), dev)
    y = ops.tcn_block(xd, *ps, d, 0.3, training=True)
    frac_zero = (y == 0).float().mean().item()
    assert 0.05 < frac_zero < 0.95
    y.sum().backward()
    assert torch.isfinite(xd.grad).all() and xd.grad.abs().sum() > 0


def test_embedding(dev):
    torch.manual_seed(9)
    V, D = 50, 300
    table = torch.randn(V, D)
    idx = torch.randint(0, V, (4, 34))
    tr = P(table, "cpu")
    ref = F.embedding(idx, tr)
    g = torch.randn_like(ref)
    ref.backward(g)
    td = P(table, dev)
    y = ops.embedding(idx.to(dev), td, 0.0)
    y.backward(g.to(dev))
    close(y, ref); close(td.grad, tr.grad, what="dtable")
    # dropout keeps ~90% and rescales
    y2 = ops.embedding(idx.to(dev), td, 0.1)
    kept = (y2 != 0).float().mean().item()
    assert 0.85 < kept < 0.95
    m = y2 != 0
    close(y2[m], (ref.to(dev) / 0.9)[m].detach())


@pytest.mark.parametrize("B,p", [(7, 0.0), (5, 0.1), (40, 0.0)])
def test_embedding_backward_padded_text(dev, B, p):
    """TED-shaped token grids (24 of 34 positions hold the padding id 0, extend_word_seq): the backward pre-aggregates the
    chunk's most frequent index before its atomics; row counts that are not a multiple of the 64-row chunk; with dropout the
    table gradient is the scatter of the upstream gradient through the SAME mask the forward drew"""
    torch.manual_seed(90 + B)
    V, D = 80, 300
    table = torch.randn(V, D)
    idx = torch.zeros(B, 34, dtype=torch.int64)
    pos = torch.rand(B, 34).argsort(dim=1)[:, :10]
    idx.scatter_(1, pos, torch.randint(4, V, (B, 10)))
    td = P(table, dev)
    y = ops.embedding(idx.to(dev), td, p)
    g = torch.randn(B, 34, D)
    y.backward(g.to(dev))
    mask = (y.detach().cpu() != 0).float() / (1.0 - p) if p > 0 else torch.ones(B, 34, D)
    want = torch.zeros(V, D, dtype=torch.float64)
    want.index_add_(0, idx.reshape(-1), (g * mask).reshape(-1, D).double())
    close(td.grad, want, tol=1e-5, what="dtable")


@pytest.mark.parametrize("In,H,sum_halves", [(24, 40, True), (8, 64, False)])
def test_bigru(dev, In, H, sum_halves):
    torch.manual_seed(10)
    B, T, L = 5, 34, 4
    gru = nn.GRU(In, H, num_layers=L, batch_first=True, bidirectional=True, dropout=0.0)
    x = torch.randn(B, T, In)
    xr = P(x, "cpu")
    o, _ = gru(xr)
    ref = o[:, :, :H] + o[:, :, H:] if sum_halves else o
    g = torch.randn_like(ref)
    ref.backward(g)
    names = []
    for l in range(L):
        for sfx in ("", "_reverse"):
            names += ["weight_ih_l%d%s" % (l, sfx), "weight_hh_l%d%s" % (l, sfx), "bias_ih_l%d%s" % (l, sfx),
                      "bias_hh_l%d%s" % (l, sfx)]
    ps = [P(getattr(gru, n).detach(), dev) for n in names]
    buf = torch.zeros(B, T, In, device=dev)
    xd = P(x, dev)
    # feed x through an identity Linear into the buffer so that dx flows back through `pieces`
    eye = torch.eye(In, device=dev)
    pc = ops.linear(xd, eye, None, out=ops.col_slice(buf, 0, In))
    y = ops.bigru(buf, ps, L, H, 0.0, False, sum_halves, pieces=(pc,), slices=((0, In),))
    y.backward(g.to(dev))
    close(y, ref, what="y"); close(xd.grad, xr.grad, what="dx")
    for q, n in zip(ps, names):
        close(q.grad, getattr(gru, n).grad, what=n, tol=5e-4)


def test_bigru_nograd_and_dropout(dev):
    torch.manual_seed(11)
    B, T, In, H, L = 3, 10, 6, 16, 2
    ps = [P(torch.randn(s) * 0.3, dev) for _ in range(L) for s in
          [(3 * H, In if _ == 0 else 2 * H), (3 * H, H), (3 * H,), (3 * H,)] * 2]
    x = torch.randn(B, T, In, device=dev)
    with torch.no_grad():
        y0 = ops.bigru(x, ps, L, H, 0.0, False)
    y1 = ops.bigru(x, ps, L, H, 0.5, True)
    assert y0.shape == (B, T, 2 * H) and (y0 - y1).abs().max() > 1e-3
    y1.sum().backward()
    assert all(torch.isfinite(q.grad).all() for q in ps)


def test_reparam_and_dhead(dev):
    torch.manual_seed(12)
    B, T, Z, H = 4, 34, 16, 64
    mu, lv, eps = torch.randn(B, Z), torch.randn(B, Z) * 0.3, torch.randn(B, Z)
    mr, lr = P(mu, "cpu"), P(lv, "cpu")
    zr = mr + eps * torch.exp(0.5 * lr)
    tiled = zr.unsqueeze(1).repeat(1, T, 1)
    g = torch.randn(B, T, Z)
    tiled.backward(g)
    buf = torch.zeros(B, T, 20, device=dev)
    md, ld = P(mu, dev), P(lv, dev)
    z, sl = ops.reparam_tile(md, ld, eps.to(dev), buf, 4)
    sl.backward(g.to(dev))
    close(z, zr); close(buf[:, :, 4:], tiled); close(md.grad, mr.grad, what="dmu"); close(ld.grad, lr.grad, what="dlv")
    # discriminator head
    gg = torch.randn(B, T, 2 * H)
    lin1, lin2 = nn.Linear(H, 1), nn.Linear(T, 1)
    gr = P(gg, "cpu")
    o = torch.sigmoid(lin2(lin1((gr[:, :, :H] + gr[:, :, H:]).reshape(-1, H)).view(B, -1)))
    go = torch.randn_like(o)
    o.backward(go)
    ps = [P(t.detach(), dev) for t in (lin1.weight, lin1.bias, lin2.weight, lin2.bias)]
    gd = P(gg, dev)
    od = ops.dhead(gd, *ps)
    od.backward(go.to(dev))
    close(od, o); close(gd.grad, gr.grad, what="dg")
    for q, r in zip(ps, (lin1.weight, lin1.bias, lin2.weight, lin2.bias)):
        close(q.grad, r.grad, what="dhead param")


def test_losses(dev):
    torch.manual_seed(13)
    B, T, Pd, Z = 6, 34, 27, 16
    out, tgt, rnd = torch.randn(B, T, Pd) * 0.3, torch.randn(B, T, Pd) * 0.3, torch.randn(B, T, Pd) * 0.3
    z, zr, mu, lv = torch.randn(B, Z), torch.randn(B, Z), torch.randn(B, Z), torch.randn(B, Z) * 0.2
    zr[0] = z[0] + 1e-4  # force the clamp(-1000) branch on one clip
    dis = torch.rand(B, 1) * 0.8 + 0.1
    o, m, l, d = P(out, "cpu"), P(mu, "cpu"), P(lv, "cpu"), P(dis, "cpu")
    huber = F.smooth_l1_loss(o / 0.1, tgt / 0.1) * 0.1
    gen = -torch.mean(torch.log(d + 1e-8))
    pl = (F.smooth_l1_loss(o / 0.05, rnd / 0.05, reduction="none") * 0.05).sum(1).sum(1)
    zl = F.l1_loss(z, zr, reduction="none").mean(1)
    div = torch.clamp(-(pl / (zl + 1e-5)), min=-1000).mean()
    kld = -0.5 * torch.mean(1 + l - m.pow(2) - l.exp())
    W = (500.0, 0.1, 0.05, 5.0)
    total = W[0] * huber + W[1] * kld + W[2] * div + W[3] * gen
    total.backward()
    losses = torch.zeros(5, device=dev)
    t = lambda a: a.to(dev)
    g_out, g_dis, g_mu, g_lv = ops.gen_loss(t(out), t(tgt), t(rnd), t(z), t(zr), t(mu), t(lv), t(dis), W, losses)
    close(losses, torch.stack([huber, gen, kld, div, total]), what="losses")
    close(g_out, o.grad, what="g_out"); close(g_dis, d.grad, what="g_dis"); close(g_mu, m.grad); close(g_lv, l.grad)
    # D loss
    dr, df = torch.rand(B, 1) * 0.8 + 0.1, torch.rand(B, 1) * 0.8 + 0.1
    a, b = P(dr, "cpu"), P(df, "cpu")
    dl = torch.sum(-torch.mean(torch.log(a + 1e-8) + torch.log(1 - b + 1e-8)))
    dl.backward()
    slot = torch.zeros(1, device=dev)
    gr, gf = ops.dis_loss(t(dr), t(df), slot)
    close(slot, dl.reshape(1)); close(gr, a.grad); close(gf, b.grad)
    # L1 metric
    ops.l1_mean(t(out), t(tgt), slot)
    close(slot, F.l1_loss(out, tgt).reshape(1))


def test_adam(dev):
    torch.manual_seed(14)
    n = 5000
    p0 = torch.randn(n)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=5e-4, betas=(0.5, 0.999))
    p = p0.clone().to(dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    for i in range(3):
        g = torch.randn(n)
        pr.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.to(dev), m, v, 5e-4, 0.5, 0.999, 1e-8, step)
    assert int(step.item()) == 3
    close(p, pr, tol=1e-6)


def test_attention(dev):
    """named off-path kernel: net/ser_att_conv_rnn_v2.py:16-34 (Attention.forward)"""
    torch.manual_seed(15)
    N, T, Hd, A = 3, 150, 32, 32
    l1, l2 = nn.Linear(Hd, A), nn.Linear(A, 1)
    x = torch.randn(N, T, Hd)
    v = torch.sigmoid(l1(x))
    alphas = torch.softmax(l2(v), dim=-2)
    ref = torch.sum(x * alphas, dim=1)
    t = lambda a: a.detach().to(dev)
    out, al = ops.attention(t(x), t(l1.weight), t(l1.bias), t(l2.weight), t(l2.bias))
    close(out, ref); close(al, alphas)


def test_no_reference_cycles(dev):
    """An op's output must die by reference counting alone: a ctx <-> output cycle keeps the autograd
    graph (and its AccumulateGrad nodes, bound to the warm-up stream) alive until the cyclic GC runs,
    which breaks CUDA-graph capture of the backward pass and leaks activations."""
    import gc
    import weakref
    torch.manual_seed(30)
    B, T, C, H = 2, 34, 8, 6
    x = P(torch.randn(B, T, C), dev)
    w, b = P(torch.randn(5, C), dev), P(torch.randn(5), dev)
    bn = nn.BatchNorm1d(5).to(dev)
    conv = nn.Conv1d(C, 5, 3, padding=1).to(dev)
    tp = [P(torch.randn(C, C, 2), dev), P(torch.rand(C, 1, 1) + 0.5, dev), P(torch.randn(C), dev),
          P(torch.randn(C, C, 2), dev), P(torch.rand(C, 1, 1) + 0.5, dev), P(torch.randn(C), dev)]
    gp = [P(torch.randn(s) * 0.3, dev) for s in [(3 * H, C), (3 * H, H), (3 * H,), (3 * H,)] * 2]
    hw = [P(torch.randn(1, H), dev), P(torch.randn(1), dev), P(torch.randn(1, T), dev), P(torch.randn(1), dev)]
    buf = torch.zeros(B, T, 16, device=dev)
    makers = {
        "linear": lambda: ops.linear(x, w, b, ops.ACT_LEAKY, 0.3),
        "linear_out": lambda: ops.linear(x, w, b, ops.ACT_LEAKY, 0.3, out=ops.col_slice(buf, 0, 5)),
        "linear_t": lambda: ops.linear_t(x, P(torch.randn(4, T), dev), None, ops.ACT_LEAKY, 0.3),
        "conv_bn_act": lambda: ops.conv_bn_act(x, conv.weight, conv.bias, (1, 1, 1, 0, 1, 1), bn, ops.ACT_LEAKY, 0.3),
        "conv_act": lambda: ops.conv_bn_act(x, conv.weight, conv.bias, (1, 1, 1, 0, 1, 1), None, ops.ACT_LEAKY, 0.3),
        "bn_act": lambda: ops.bn_act(ops.linear(x, w, b), bn, ops.ACT_RELU),
        "tcn": lambda: ops.tcn_block(x, *tp, 2, 0.0, False),
        "bigru": lambda: ops.bigru(x, gp, 1, H, 0.0, False),
        "bigru_sum": lambda: ops.bigru(x, gp, 1, H, 0.0, False, sum_halves=True),
        "dhead": lambda: ops.dhead(ops.bigru(x, gp, 1, H, 0.0, False), *hw),
    }
    gc.collect()
    gc.disable()
    try:
        for name, mk in makers.items():
            y = mk()
            r = weakref.ref(y)
            del y
            assert r() is None, "reference cycle through the output of %s" % name
    finally:
        gc.enable()


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,C,d,p", [(7, 34, 300, 1, 0.0), (7, 34, 300, 8, 0.3), (2, 100, 300, 8, 0.3), (5, 34, 24, 4, 0.2),
                                        (150, 34, 300, 2, 0.3)])
@pytest.mark.parametrize("pair", [False, True])
def test_tcn_block_fused_vs_two_launch(B, T, C, d, p, pair):
    """single-kernel TCN block (csrc/umma_tcn.cu) against the two conv-as-GEMM launches (csrc/tcn.cu) from the same
    dropout seed (identical masks): output, saved activations through the backward pass, all parameter gradients.
    Covers the column split of the accumulator (C = 300 -> 160 + 144), partial last tiles, one clip per tile."""
    dev = torch.device("cuda:0")
    from speech2affective_gestures_b200 import _C
    if _C.is_emulated():
        _C._lib, _C._emulated = None, False
    torch.manual_seed(21)
    raw = [torch.randn(C, C, 2) * 0.08, torch.rand(C, 1, 1) + 0.5, torch.randn(C) * 0.1,
           torch.randn(C, C, 2) * 0.08, torch.rand(C, 1, 1) + 0.5, torch.randn(C) * 0.1]
    x = torch.randn(B, T, C)
    g = torch.randn(B, T, C)
    res = {}
    default = ops.TCN_FUSED[0]
    _C.lib().s2ag_debug_flags(8192 if pair else 0)   # CTA-pair (cta_group::2) variant of the fused kernel
    for fused in (True, False):
        ops.TCN_FUSED[0] = fused
        try:
            ps = [P(t, dev) for t in raw]
            xd = P(x, dev)
            ops.manual_seed(77)
            y = ops.tcn_block(xd, *ps, d, p, training=True)
            y.backward(g.to(dev))
            res[fused] = [y.detach(), xd.grad] + [q.grad for q in ps]
        finally:
            ops.TCN_FUSED[0] = default
    if p > 0:
        assert 0.02 < (res[True][0] == 0).float().mean().item() < 0.98
    names = ("out", "dx", "dv1", "dg1", "db1", "dv2", "dg2", "db2")
    close(res[True][0], res[False][0], what="out")
    if B * T * C > 200000:
        # two fp32-grade summation orders: a few of the ~1e6 ReLU inputs land on the other side of 0 and each flip
        # moves single gradient elements by O(1) -- the flip-robust criteria of the module tests (common.check_grads)
        from common import check_grads
        check_grads(dict(zip(names[1:], res[True][1:])), dict(zip(names[1:], res[False][1:])), tol=2e-3, what="fused TCN")
    else:
        for a, b, nm in zip(res[True][1:], res[False][1:], names[1:]):
            close(a, b, what=nm)
    # no-grad call (y1 / y2 never written): same output
    ps = [t.to(dev) for t in raw]
    ops.manual_seed(77)
    ops.TCN_FUSED[0] = True
    try:
        with torch.no_grad():
            y0 = ops.tcn_block(x.to(dev), *ps, d, p, training=True)
    finally:
        ops.TCN_FUSED[0] = default
    _C.lib().s2ag_debug_flags(0)
    close(y0, res[True][0], what="no-grad out")


@pytest.mark.gpu
def test_tcn_block_fused_bf16x1_mode():
    """single-pass bf16 operands (s2ag_set_precision(1)) through the single-kernel TCN block: hi-plane-only path"""
    dev = torch.device("cuda:0")
    from speech2affective_gestures_b200 import _C
    if _C.is_emulated():
        _C._lib, _C._emulated = None, False
    torch.manual_seed(3)
    B, T, C, d = 7, 34, 300, 4
    raw = [torch.randn(C, C, 2) * 0.08, torch.rand(C, 1, 1) + 0.5, torch.randn(C) * 0.1,
           torch.randn(C, C, 2) * 0.08, torch.rand(C, 1, 1) + 0.5, torch.randn(C) * 0.1]
    ps = [t.to(dev) for t in raw]
    x = torch.randn(B, T, C, device=dev)
    default = ops.TCN_FUSED[0]
    out = {}
    try:
        ops.TCN_FUSED[0] = True
        with torch.no_grad():
            out[0] = ops.tcn_block(x, *ps, d, 0.0, training=False)
            assert _C.lib().s2ag_set_precision(1) == 0
            out[1] = ops.tcn_block(x, *ps, d, 0.0, training=False)
    finally:
        _C.lib().s2ag_set_precision(0)
        ops.TCN_FUSED[0] = default
    err = (out[1] - out[0]).abs().max().item() / out[0].abs().max().item()
    assert 1e-6 < err < 3e-2
