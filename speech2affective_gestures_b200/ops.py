"""Host-side operator layer: torch.autograd.Function wrappers over the C ABI (include/s2ag.h).

PyTorch is used here for device memory (torch.empty), stream handles and the autograd tape; every
FLOP of the hot path is executed by libs2ag_b200.so.  There is no fallback: tensors must live on
a CUDA device (or, in tests/emu only, on the CPU with the kernel-logic emulator injected).

Layout convention (see include/s2ag.h): activations are channels-last; parameter gradients are
ACCUMULATED by the kernels straight into `param.grad` (which the network modules alias onto one
flat buffer per network), so the backward functions return None for parameters.
"""
import contextlib
import ctypes
import itertools
import os

import torch

from . import _C

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2

_seed_counter = itertools.count(1)
_base_seed = 0x5A2A6000
_seed_dev = {}


def manual_seed(seed):
    """Re-seed the dropout stream of this process (mirrors torch.manual_seed at processor_v2.py:37)."""
    global _seed_counter, _base_seed
    _base_seed = int(seed) & 0xFFFFFFFF
    _seed_counter = itertools.count(1)


def seed_state():
    return _base_seed


def next_seed():
    return (_base_seed << 20) + next(_seed_counter) * 0x9E3779B1


def seed_nonce(device):
    """Device-resident uint64 added to every dropout seed; advanced once per training step so
    CUDA-graph replays draw fresh masks."""
    key = str(device)
    if key not in _seed_dev:
        _seed_dev[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return _seed_dev[key]


def advance_seed_nonce(device, inc=0x100000001B3):
    t = seed_nonce(device)
    _C.call("s2ag_seed_advance", _p(t), ctypes.c_uint64(inc), _stream(t))


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


# Scratch for the packed weight-operand images of the tcgen05 contractions (s2ag_register_scratch): one buffer per
# stream that launches library kernels, owned here.  Never allocated while a CUDA graph is being captured (a stream
# first seen during capture simply runs the contractions that stage both operands on the fly).
SCRATCH_BYTES = 192 << 20   # packed weight image + packed activation image of the largest contraction (gemm_umma_tt.cuh)
_SCRATCH = {}


def _handle(stream, device):
    h = stream.cuda_stream
    key = (device.index, h)
    if key not in _SCRATCH and not torch.cuda.is_current_stream_capturing():
        buf = torch.empty(SCRATCH_BYTES, dtype=torch.uint8, device=device)
        _C.call("s2ag_register_scratch", ctypes.c_void_p(h), ctypes.c_void_p(buf.data_ptr()), SCRATCH_BYTES)
        _SCRATCH[key] = buf
    return ctypes.c_void_p(h)


def has_scratch(stream, device):
    return (device.index, stream.cuda_stream) in _SCRATCH


def _stream(t):
    if t.is_cuda:
        return _handle(torch.cuda.current_stream(t.device), t.device)
    return None


def _check(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda and not _C.is_emulated():
            raise _C.S2agError("s2ag ops need CUDA tensors: this path has no CPU fallback")
        if t.dtype not in (torch.float32, torch.int64, torch.int32, torch.float64):
            raise _C.S2agError("unsupported dtype %s" % t.dtype)


def _rows(t, ncols):
    """View t as [M, ncols] rows with a uniform row stride (last dim contiguous). -> (tensor, ld)"""
    if t.shape[-1] != ncols:
        raise _C.S2agError("expected last dim %d, got %s" % (ncols, tuple(t.shape)))
    if t.dim() == 1:
        t = t.unsqueeze(0)
    ok = t.stride(-1) == 1 or ncols == 1
    ld = t.stride(-2) if t.shape[-2] > 1 else max(ncols, t.stride(-2))
    # outer dims must be expressible as multiples of the row stride
    exp = ld
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] > 1 and t.stride(d) != exp:
            ok = False
        exp *= t.shape[d]
    if ld < ncols:
        ok = False
    if not ok:
        t = t.contiguous()
        ld = ncols
    return t, ld


def col_slice(buf, a, b):
    """Fresh (non-view) alias of buf[..., a:b] for a contiguous `buf`: producers write their
    features straight into their column range of the GRU input buffer, and because the alias is
    not an autograd view of `buf`, each producer's output is an ordinary graph node."""
    assert buf.is_contiguous()
    size = tuple(buf.shape[:-1]) + (b - a,)
    return buf.new_empty(0).set_(buf.untyped_storage(), buf.storage_offset() + a, size, buf.stride())


class Out:
    """Holder that smuggles a pre-allocated destination past autograd's input bookkeeping."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t


def _grad_of(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


def _empty(shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


# ------------------------------------------------------------------------------------------ Linear
class LinearFn(torch.autograd.Function):
    """y = act(x @ w.T + b) on the last dim; `out` (optional) is a pre-allocated, possibly
    column-sliced destination (this is how torch.cat at net/multimodal_context_net_v2.py:526 is
    avoided).  Reference: nn.Linear call sites listed in include/s2ag.h."""

    @staticmethod
    def forward(ctx, x, w, b, act, slope, out):
        _check(x, w, b)
        N, K = w.shape
        xr, ldx = _rows(x, K)
        M = x.numel() // K
        y = out.t if out is not None else _empty(x.shape[:-1] + (N,), x)
        yr, ldy = _rows(y, N)
        assert yr.data_ptr() == y.data_ptr(), "out must be row-strided"
        _C.call("s2ag_linear_fwd", _p(xr), ldx, _p(w), _p(b), _p(yr), ldy, M, N, K, act, float(slope), _stream(x))
        ctx.t = (xr, w, y.detach())  # detached alias: storing `y` itself would tie ctx <-> output in a cycle
        ctx.cfg = (act, float(slope), M, N, K, ldx, b)
        ctx.xshape = x.shape
        ctx.want_w = w.requires_grad  # decided at forward time (a caller may freeze the weights around one pass)
        return y

    @staticmethod
    def backward(ctx, dy):
        xr, w, y = ctx.t
        act, slope, M, N, K, ldx, b = ctx.cfg
        dyr, lddy = _rows(dy, N)
        st = _stream(dy)
        if act != ACT_NONE:
            yr, ldy = _rows(y, N)
            dpre = _empty((M, N), dy)
            _C.call("s2ag_act_bwd", _p(dyr), lddy, _p(yr), ldy, _p(dpre), N, M, N, act, slope, st)
            dyr, lddy = dpre, N
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _empty(ctx.xshape, dy)
            _C.call("s2ag_linear_bwd_data", _p(dyr), lddy, _p(w), _p(dx), K, M, N, K, 0, st)
        if ctx.want_w:
            db = _grad_of(b) if (b is not None and b.requires_grad) else None
            _C.call("s2ag_linear_bwd_weight", _p(dyr), lddy, _p(xr), ldx, _p(_grad_of(w)), _p(db), M, N, K, st)
        return dx, None, None, None, None, None


def linear(x, w, b=None, act=ACT_NONE, slope=0.0, out=None):
    return LinearFn.apply(x, w, b, act, slope, None if out is None else Out(out))


class LinearTFn(torch.autograd.Function):
    """MFCCEncoder.linear1 (net/multimodal_context_net_v2.py:49,57) on channels-last data:
    y[b,c,:] = act(w @ x[b,:,c] + bias);  x[B,L,C] -> y[B,C,N] (optionally into a column slice)."""

    @staticmethod
    def forward(ctx, x, w, b, act, slope, out):
        _check(x, w, b)
        x = x.contiguous()
        B, L, C = x.shape
        N = w.shape[0]
        y = out.t if out is not None else _empty((B, C, N), x)
        yr, ldy = _rows(y, N)
        assert yr.data_ptr() == y.data_ptr()
        _C.call("s2ag_linear_t_fwd", _p(x), _p(w), _p(b), _p(yr), ldy, B, L, C, N, act, float(slope), _stream(x))
        ctx.t = (x, w, b, y.detach())
        ctx.cfg = (B, L, C, N, act, float(slope))
        ctx.want_w = w.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b, y = ctx.t
        B, L, C, N, act, slope = ctx.cfg
        st = _stream(dy)
        dyr, lddy = _rows(dy, N)
        if act != ACT_NONE:
            yr, ldy = _rows(y, N)
            dpre = _empty((B * C, N), dy)
            _C.call("s2ag_act_bwd", _p(dyr), lddy, _p(yr), ldy, _p(dpre), N, B * C, N, act, slope, st)
            dyr, lddy = dpre, N
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _empty((B, L, C), dy)
            _C.call("s2ag_linear_t_bwd_data", _p(dyr), lddy, _p(w), _p(dx), B, L, C, N, st)
        if ctx.want_w:
            db = _grad_of(b) if (b is not None and b.requires_grad) else None
            _C.call("s2ag_linear_t_bwd_weight", _p(dyr), lddy, _p(x), _p(_grad_of(w)), _p(db), B, L, C, N, st)
        return dx, None, None, None, None, None


def linear_t(x, w, b=None, act=ACT_NONE, slope=0.0, out=None):
    return LinearTFn.apply(x, w, b, act, slope, None if out is None else Out(out))


# ------------------------------------------------------------------------------------------ BatchNorm helper
class _BnState:
    """The per-call record a BN forward leaves for its backward."""
    __slots__ = ("x", "ldx", "M", "C", "mean", "invstd", "training", "act", "slope", "y", "ldy", "cmap", "pmap",
                 "want_w", "groups")


_bn_repeat = [1]
_side_stream = [None]
_side_stream_conv = [None]
_side_stream_tcn = [None]


def set_side_stream(stream, conv_stream=None, tcn_stream=None):
    """Second CUDA stream for work that is off the critical path of the recurrent kernels (weight-gradient GEMMs of a
    GRU layer).  `conv_stream` (default: the same stream): where the weight gradients of the convolutions go -- a stream
    of their own keeps them from queueing behind the text encoder's backward and the GRU weight gradients.  The caller
    must make its main stream wait for both before consuming parameter gradients."""
    _side_stream[0] = stream
    _side_stream_conv[0] = conv_stream if conv_stream is not None else stream
    _side_stream_tcn[0] = tcn_stream if tcn_stream is not None else _side_stream_conv[0]   # TCN-block weight gradients



class bn_repeat:
    """Context: a train-mode BatchNorm forward inside it stands for `n` identical forward calls of the
    reference (same input, same weights => same batch statistics and output); the running statistics are
    updated as n successive momentum updates would, r <- (1-m)^n r + (1-(1-m)^n) s, and
    num_batches_tracked advances by n.  Used where the reference runs the SAME encoder on the SAME input
    several times per step (processor_v2.py:798,823,909: three generator passes)."""

    def __init__(self, n):
        self.n = int(n)

    def __enter__(self):
        self.prev = _bn_repeat[0]
        _bn_repeat[0] = self.n

    def __exit__(self, *a):
        _bn_repeat[0] = self.prev


_bn_groups = [1]


class bn_groups:
    """Context: every BatchNorm forward inside it sees `n` consecutive, equally sized, INDEPENDENT batches stacked
    along the leading dimension -- what the reference computes as n successive calls of the same module
    (processor_v2.py:808-809: D(target), D(out.detach())).  Statistics are per group; the running statistics get the
    groups' momentum updates in call order.  All other kernels are per-sample, so stacking is otherwise invisible."""

    def __init__(self, n):
        self.n = int(n)

    def __enter__(self):
        self.prev = _bn_groups[0]
        _bn_groups[0] = self.n

    def __exit__(self, *a):
        _bn_groups[0] = self.prev


def _bn_forward(x2, ldx, M, C, bn, training, act, slope, y2, ldy, add2=None, ldadd=0, cmap=None, pmap=None):
    """x2/y2: row views. bn: module-like with weight,bias,running_mean,running_var,momentum,eps."""
    groups = _bn_groups[0]
    mean = _empty((groups * C,), x2)
    invstd = _empty((groups * C,), x2)
    ws = torch.empty(2 * C * groups, dtype=torch.float64, device=x2.device)
    rep = _bn_repeat[0] if training else 1
    momentum = float(bn.momentum) if rep == 1 else 1.0 - (1.0 - float(bn.momentum)) ** rep
    _C.call("s2ag_bn_fwd", _p(x2), ldx, M, C, _p(bn.weight), _p(bn.bias), _p(pmap), _p(bn.running_mean),
            _p(bn.running_var), 1 if training else 0, momentum, float(bn.eps), _p(add2), ldadd, _p(y2), ldy,
            _p(cmap), act, float(slope), _p(mean), _p(invstd), _p(ws), groups, _stream(x2))
    if training:
        bn._s2ag_batches = getattr(bn, "_s2ag_batches", 0) + rep * groups
    s = _BnState()
    s.x, s.ldx, s.M, s.C, s.mean, s.invstd, s.training = x2, ldx, M, C, mean, invstd, training
    s.act, s.slope, s.y, s.ldy, s.cmap, s.pmap = act, float(slope), y2, ldy, cmap, pmap
    s.want_w = bn.weight.requires_grad
    s.groups = groups
    return s


def _bn_backward(s, bn, dy2, lddy, need_dx=True, dadd2=None, lddadd=0):
    ws = torch.empty(2 * s.C * s.groups, dtype=torch.float64, device=dy2.device)
    dx = _empty((s.M, s.C), dy2) if need_dx else None
    wg = s.want_w
    _C.call("s2ag_bn_bwd", _p(dy2), lddy, _p(s.y), s.ldy, _p(s.cmap), _p(s.x), s.ldx, s.M, s.C, _p(bn.weight),
            _p(s.pmap), _p(s.mean), _p(s.invstd), 1 if s.training else 0, s.act, s.slope, _p(dx), s.C,
            _p(_grad_of(bn.weight)) if wg else None, _p(_grad_of(bn.bias)) if wg else None, _p(dadd2), lddadd,
            _p(ws), s.groups, _stream(dy2))
    return dx


class BnActFn(torch.autograd.Function):
    """y = act(BN(x) [+ add]) over channels-last x[..., C]; optional column maps (see s2ag_bn_fwd).
    Reference: nn.BatchNorm1d/2d call sites in include/s2ag.h."""

    @staticmethod
    def forward(ctx, x, add, gamma, beta, bn, training, act, slope, cmap, pmap, out):
        _check(x, add)
        C = x.shape[-1]
        x2, ldx = _rows(x, C)
        M = x.numel() // C
        y = out.t if out is not None else _empty(x.shape, x)
        y2, ldy = _rows(y.detach(), C)
        assert y2.data_ptr() == y.data_ptr()
        add2, ldadd = (None, 0) if add is None else _rows(add, C)
        ctx.s = _bn_forward(x2, ldx, M, C, bn, training, act, slope, y2, ldy, add2, ldadd, cmap, pmap)
        ctx.bn = bn
        ctx.has_add = add is not None
        ctx.xshape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        s = ctx.s
        dy2, lddy = _rows(dy, s.C)
        dadd = _empty(ctx.xshape, dy) if (ctx.has_add and ctx.needs_input_grad[1]) else None
        dx = _bn_backward(s, ctx.bn, dy2, lddy, ctx.needs_input_grad[0], dadd, s.C)
        if dx is not None:
            dx = dx.view(ctx.xshape)
        return dx, dadd, None, None, None, None, None, None, None, None, None


def bn_act(x, bn, act=ACT_NONE, slope=0.0, add=None, cmap=None, pmap=None, out=None):
    return BnActFn.apply(x, add, bn.weight, bn.bias, bn, bn.training, act, slope, cmap, pmap,
                         None if out is None else Out(out))


# ------------------------------------------------------------------------------------------ Conv (+BN +act)
def _conv_out(L, k, s, p, d):
    return (L + 2 * p - d * (k - 1) - 1) // s + 1


class ConvBnActFn(torch.autograd.Function):
    """channels-last Conv1d/Conv2d -> [BatchNorm] -> activation.
    x: [N,H,W,Cin] (Conv1d: [N,L,Cin], handled as W=1); weight in the reference layout
    [Cout,Cin,KH,KW] / [Cout,Cin,K].  Reference: WavEncoder/MFCCEncoder/AffEncoder/STGraphConv."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, conv, bn, training, act, slope, cmap, pmap, out):
        _check(x, w, b)
        is1d = x.dim() == 3
        N, H = x.shape[0], x.shape[1]
        W = 1 if is1d else x.shape[2]
        Cin = x.shape[-1]
        Cout = w.shape[0]
        KH, KW = (w.shape[2], 1) if is1d else (w.shape[2], w.shape[3])
        sh, sw, ph, pw, dh, dw = conv
        Ho, Wo = _conv_out(H, KH, sh, ph, dh), _conv_out(W, KW, sw, pw, dw)
        x2, ldx = _rows(x, Cin)
        oshape = (N, Ho, Cout) if is1d else (N, Ho, Wo, Cout)
        st = _stream(x)
        M = N * Ho * Wo
        if bn is None:
            y = out.t if out is not None else _empty(oshape, x)
            y2, ldy = _rows(y.detach(), Cout)
            _C.call("s2ag_conv_fwd", _p(x2), ldx, N, H, W, Cin, _p(w), _p(b), _p(y2), ldy, Cout, KH, KW, sh, sw, ph,
                    pw, dh, dw, act, float(slope), st)
            ctx.s = None
            ctx.c = None
        else:
            c = _empty((M, Cout), x)
            _C.call("s2ag_conv_fwd", _p(x2), ldx, N, H, W, Cin, _p(w), _p(b), _p(c), Cout, Cout, KH, KW, sh, sw, ph, pw,
                    dh, dw, ACT_NONE, 0.0, st)
            y = out.t if out is not None else _empty(oshape, x)
            y2, ldy = _rows(y.detach(), Cout)
            ctx.s = _bn_forward(c, Cout, M, Cout, bn, training, act, slope, y2, ldy, None, 0, cmap, pmap)
        assert y2.data_ptr() == y.data_ptr()
        ctx.bn = bn
        ctx.geom = (N, H, W, Cin, Cout, KH, KW, sh, sw, ph, pw, dh, dw, M, act, float(slope), ldx, ldy)
        ctx.x2, ctx.w, ctx.b, ctx.y2 = x2, w, b, y2
        ctx.xshape = x.shape
        ctx.want_w = w.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, Cin, Cout, KH, KW, sh, sw, ph, pw, dh, dw, M, act, slope, ldx, ldy = ctx.geom
        st = _stream(dy)
        dy2, lddy = _rows(dy, Cout)
        if ctx.s is not None:
            dc = _bn_backward(ctx.s, ctx.bn, dy2, lddy, True)
            lddc = Cout
        elif act != ACT_NONE:
            dc = _empty((M, Cout), dy)
            _C.call("s2ag_act_bwd", _p(dy2), lddy, _p(ctx.y2), ldy, _p(dc), Cout, M, Cout, act, slope, st)
            lddc = Cout
        else:
            dc, lddc = dy2, lddy
        w, b = ctx.w, ctx.b
        if ctx.want_w:
            db = _grad_of(b) if (b is not None and b.requires_grad) else None
            side = _side_stream_conv[0]
            cur = torch.cuda.current_stream(dy.device) if dy.is_cuda else None
            wst = st
            if side is not None and cur is not None and cur != side:
                # (also when no data gradient follows: every accumulation into a convolution's parameter gradients is
                # serialised on this one stream, whichever stream the backward chain of the module instance runs on)
                # the weight gradient is a leaf of the backward graph: it runs on the side stream beside the
                # data-gradient chain (the caller joins the side stream before it consumes parameter gradients)
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
                wst = _handle(side, dy.device)
                for t_ in (dc, ctx.x2):
                    t_.record_stream(side)
            _C.call("s2ag_conv_bwd_weight", _p(dc), lddc, _p(ctx.x2), ldx, N, H, W, Cin, _p(_grad_of(w)), _p(db), Cout,
                    KH, KW, sh, sw, ph, pw, dh, dw, wst)
        dx = None
        if ctx.needs_input_grad[0]:
            if sh != 1 or sw != 1:
                raise _C.S2agError("conv data-gradient implemented for stride 1 only (strided convs are frozen)")
            dx = _empty(ctx.xshape, dy)
            _C.call("s2ag_conv_bwd_data", _p(dc), lddc, N, H, W, Cin, _p(w), _p(dx), Cin, Cout, KH, KW, ph, pw, dh, dw,
                    0, st)
        return (dx,) + (None,) * 12


def conv_bn_act(x, conv_w, conv_b, geom, bn=None, act=ACT_NONE, slope=0.0, cmap=None, pmap=None, out=None):
    """geom = (sh, sw, ph, pw, dh, dw)"""
    g, be, tr = (None, None, False) if bn is None else (bn.weight, bn.bias, bn.training)
    return ConvBnActFn.apply(x, conv_w, conv_b, g, be, tuple(geom), bn, tr, act, slope, cmap, pmap,
                             None if out is None else Out(out))


@torch.no_grad()
def wavencoder_fwd(audio, convs, bns, slope, out=None):
    """Fused WavEncoder forward (s2ag_wavencoder_fwd, csrc/umma_wav.cu): audio [B, L] -> [B, 34, 32] with the reference
    layer geometry (net/multimodal_context_net_v2.py:17-28).  convs: the four nn.Conv1d, bns: the three nn.BatchNorm1d
    parameter containers (train mode: batch statistics + running-statistic update; eval: running statistics).  No
    autograd: the encoder is frozen on the hot path (PoseGeneratorTriModal)."""
    _check(audio)
    import ctypes as _ct
    geom = [(c.in_channels, c.out_channels, c.kernel_size[0], c.stride[0], c.padding[0]) for c in convs]
    if geom != [(1, 16, 15, 5, 1600), (16, 32, 15, 6, 0), (32, 64, 15, 6, 0), (64, 32, 15, 6, 0)]:
        raise _C.S2agError("wavencoder_fwd: layer geometry differs from the reference WavEncoder: %s" % (geom,))
    training = bns[0].training
    if any(b.training != training for b in bns) or _bn_groups[0] != 1:
        raise _C.S2agError("wavencoder_fwd: mixed BatchNorm modes / statistic groups are not supported")
    B, L = audio.shape
    audio = audio.contiguous()
    st = _stream(audio)
    n_ws = _C.lib().s2ag_wavencoder_ws_floats(B, L)
    if n_ws <= 0:
        raise _C.S2agError("wavencoder_fwd: audio too short (%d samples)" % L)
    ws = torch.empty(n_ws, dtype=torch.float32, device=audio.device)
    L1 = _conv_out(L, 15, 5, 1600, 1)
    Lo = _conv_out(_conv_out(_conv_out(L1, 15, 6, 0, 1), 15, 6, 0, 1), 15, 6, 0, 1)
    y = out if out is not None else _empty((B, Lo, 32), audio)
    y2, ldy = _rows(y, 32)
    arr = lambda ts: (_ct.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])
    rep = _bn_repeat[0] if training else 1
    mom = float(bns[0].momentum) if rep == 1 else 1.0 - (1.0 - float(bns[0].momentum)) ** rep
    _C.call("s2ag_wavencoder_fwd", _p(audio), B, L, arr([c.weight for c in convs]), arr([c.bias for c in convs]),
            arr([b.weight for b in bns]), arr([b.bias for b in bns]), arr([b.running_mean for b in bns]),
            arr([b.running_var for b in bns]), 1 if training else 0, mom, float(bns[0].eps), float(slope), _p(y2), ldy,
            _p(ws), st)
    if training:
        for b in bns:
            b._s2ag_batches = getattr(b, "_s2ag_batches", 0) + rep
    return y


# ------------------------------------------------------------------------------------------ ST-GCN graph contraction
class GraphFn(torch.autograd.Function):
    """y[n,t,w,c] = sum_{k,v} x[n,t,v,k*C+c] A[k,v,w]   (net/utils/tgcn.py:66-69)"""

    @staticmethod
    def forward(ctx, x, A):
        _check(x, A)
        K, V, _ = A.shape
        KC = x.shape[-1]
        C = KC // K
        x = x.contiguous()
        M = x.numel() // (V * KC)
        y = _empty(x.shape[:-1] + (C,), x)
        _C.call("s2ag_graph_fwd", _p(x), _p(A), _p(y), M, V, K, C, _stream(x))
        ctx.A = A
        ctx.dims = (M, V, K, C, x.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        M, V, K, C, xshape = ctx.dims
        dy = dy.contiguous()
        dx = _empty(xshape, dy)
        _C.call("s2ag_graph_bwd", _p(dy), _p(ctx.A), _p(dx), M, V, K, C, _stream(dy))
        return dx, None


def graph_contract(x, A):
    return GraphFn.apply(x, A)


# A/B: 1 = weight gradient of the composed convolution as a plain contraction over a sliding-window view of the padded
# input (s2ag_window_wgrad); parity-tested, measured SLOWER inside the step (11.88 vs 11.72 ms: two padding passes + a
# 1296-row contraction whose B operand must be packed), so the transposed-im2col route stays the default
GCN_WGRAD_WINDOW = [os.environ.get("S2AG_GCN_WGRAD_WINDOW", "0") == "1"]


class GcnFn(torch.autograd.Function):
    """ConvTemporalGraphical (net/utils/tgcn.py:15-71) as ONE temporal convolution: the (Kt x 1) Conv2d Cin -> K*C and
    einsum('nkctv,kvw->nctw') are composed into a Conv1d over the channels-last rows [N, T, V*Cin] -> [N, T, V*C]
    (s2ag_gcn_compose_fwd builds the composed weight every call: the weights change every step), so the [N, T, V, K*C]
    intermediate is never written.  Backward: data gradient and weight gradient of that Conv1d, then the composed
    weight gradient is folded back onto the Conv2d's parameters (s2ag_gcn_compose_bwd, += into their .grad)."""

    @staticmethod
    def forward(ctx, x, w, b, A, pad):
        _check(x, w, b, A)
        N, T, V, Cin = x.shape
        K = A.shape[0]
        KC, _, Kt, _ = w.shape
        C = KC // K
        x = x.contiguous()
        st = _stream(x)
        weff = _empty((V * C, V * Cin, Kt), x)
        beff = _empty((V * C,), x)
        _C.call("s2ag_gcn_compose_fwd", _p(w), _p(b), _p(A), _p(weff), _p(beff), V, K, C, Cin, Kt, st)
        y = _empty((N, T, V, C), x)
        _C.call("s2ag_conv_fwd", _p(x), V * Cin, N, T, 1, V * Cin, _p(weff), _p(beff), _p(y), V * C, V * C, Kt, 1, 1, 1,
                int(pad), 0, 1, 1, ACT_NONE, 0.0, st)
        ctx.t = (x, w, b, A, weff)
        ctx.dims = (N, T, V, Cin, K, C, Kt, int(pad))
        ctx.want_w = w.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b, A, weff = ctx.t
        N, T, V, Cin, K, C, Kt, pad = ctx.dims
        dy = dy.contiguous()
        st = _stream(dy)
        if ctx.want_w:
            side = _side_stream_conv[0]
            cur = torch.cuda.current_stream(dy.device) if dy.is_cuda else None
            wst = st
            if side is not None and cur is not None and cur != side:
                ev = torch.cuda.Event()   # weight gradient beside the data-gradient chain (see ConvBnActFn.backward)
                ev.record(cur)
                side.wait_event(ev)
                wst = _handle(side, dy.device)
                for t_ in (dy, x):
                    t_.record_stream(side)
                ctxm = torch.cuda.stream(side)
            else:
                ctxm = contextlib.nullcontext()
            has_b = b is not None and b.requires_grad
            if GCN_WGRAD_WINDOW[0] and Kt == 2 * pad + 1:
                # plain contraction over the sliding-window view of the zero-padded input (s2ag_window_wgrad): no
                # transposed im2col gather
                Tp = T + 2 * pad
                with ctxm:
                    xp = _empty((N * Tp + Kt - 1, V * Cin), dy)
                    dyp = _empty((N * Tp, V * C), dy)
                    dwt = torch.zeros((Kt * V * Cin, V * C), dtype=torch.float32, device=dy.device)
                    dbeff = torch.zeros(V * C, dtype=torch.float32, device=dy.device)
                _C.call("s2ag_pad_time", _p(x), V * Cin, _p(xp), N, T, V * Cin, Tp, pad, Kt - 1, wst)
                _C.call("s2ag_pad_time", _p(dy), V * C, _p(dyp), N, T, V * C, Tp, 0, 0, wst)
                _C.call("s2ag_window_wgrad", _p(dyp), V * C, _p(xp), V * Cin, _p(dwt), N * Tp, V * C, Kt * V * Cin, wst)
                if has_b:
                    _C.call("s2ag_colsum", _p(dy), V * C, _p(dbeff), N * T, V * C, wst)
                _C.call("s2ag_gcn_compose_bwd", _p(dwt), _p(dbeff), _p(A), _p(_grad_of(w)),
                        _p(_grad_of(b)) if has_b else None, V, K, C, Cin, Kt, 1, wst)
            else:
                with ctxm:
                    dweff = torch.zeros_like(weff)
                    dbeff = torch.zeros(V * C, dtype=torch.float32, device=dy.device)
                _C.call("s2ag_conv_bwd_weight", _p(dy), V * C, _p(x), V * Cin, N, T, 1, V * Cin, _p(dweff), _p(dbeff),
                        V * C, Kt, 1, 1, 1, pad, 0, 1, 1, wst)
                _C.call("s2ag_gcn_compose_bwd", _p(dweff), _p(dbeff), _p(A), _p(_grad_of(w)),
                        _p(_grad_of(b)) if has_b else None, V, K, C, Cin, Kt, 0, wst)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _empty(x.shape, dy)
            _C.call("s2ag_conv_bwd_data", _p(dy), V * C, N, T, 1, V * Cin, _p(weff), _p(dx), V * Cin, V * C, Kt, 1, pad, 0,
                    1, 1, 0, st)
        return dx, None, None, None, None


def gcn_conv(x, w, b, A, pad):
    """x [N, T, V, Cin] -> [N, T, V, C]; w / b: the reference's gcn.conv parameters, A [K, V, V]"""
    return GcnFn.apply(x, w, b, A, pad)


# ------------------------------------------------------------------------------------------ Embedding
class EmbeddingFn(torch.autograd.Function):
    """out = dropout(table[idx])   (net/multimodal_context_net_v2.py:88, :513)"""

    @staticmethod
    def forward(ctx, idx, table, p, seed):
        _check(idx, table)
        idx = idx.contiguous()
        V, D = table.shape
        out = _empty(tuple(idx.shape) + (D,), table)
        nonce = seed_nonce(table.device) if p > 0 else None
        _C.call("s2ag_embedding_fwd", _p(idx), _p(table), _p(out), D, idx.numel(), D, V, float(p),
                ctypes.c_uint64(seed), _p(nonce), _stream(table))
        ctx.idx, ctx.table = idx, table
        ctx.cfg = (float(p), seed, V, D, nonce)
        ctx.want_w = table.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        p, seed, V, D, nonce = ctx.cfg
        if ctx.want_w:
            d2, ld = _rows(dout, D)
            _C.call("s2ag_embedding_bwd", _p(ctx.idx), _p(d2), ld, _p(_grad_of(ctx.table)), ctx.idx.numel(), D, V, p,
                    ctypes.c_uint64(seed), _p(nonce), _stream(dout))
        return None, None, None, None


def embedding(idx, table, p=0.0):
    return EmbeddingFn.apply(idx, table, p, next_seed() if p > 0 else 0)


# ------------------------------------------------------------------------------------------ Dropout
class DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        _check(x)
        x = x.contiguous()
        y = torch.empty_like(x)
        nonce = seed_nonce(x.device)
        _C.call("s2ag_dropout", _p(x), _p(y), x.numel(), float(p), ctypes.c_uint64(seed), _p(nonce), _stream(x))
        ctx.cfg = (float(p), seed, nonce)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, seed, nonce = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        _C.call("s2ag_dropout", _p(dy), _p(dx), dy.numel(), p, ctypes.c_uint64(seed), _p(nonce), _stream(dy))
        return dx, None, None


def dropout(x, p, training):
    if not training or p <= 0:
        return x
    return DropoutFn.apply(x, p, next_seed())


# ------------------------------------------------------------------------------------------ TCN residual block
# Single-kernel TCN block (csrc/umma_tcn.cu) vs two conv-as-GEMM launches (csrc/tcn.cu).  Measured at 256 clips (B200):
# 72 us (2 launches, 86 SMs) against 77 us (6 launches, 136 SMs) timed alone, but 13.80 vs 13.50 ms per GAN step where the
# block shares the device with the recurrent kernels -- the fused kernel is bound by streaming the 1.5 MB weight image
# through every SM, so the two-launch path stays the default; S2AG_TCN_FUSED=1 selects the fused kernel
# (parity-tested either way).
TCN_FUSED = [os.environ.get("S2AG_TCN_FUSED", "0") == "1"]


class TcnBlockFn(torch.autograd.Function):
    """weight_norm + TemporalBlock (net/tcn.py:16-46) over channels-last x[B,T,C]."""

    @staticmethod
    def forward(ctx, x, v1, g1, b1, v2, g2, b2, dilation, p, seed):
        _check(x, v1, g1, b1, v2, g2, b2)
        x = x.contiguous()
        B, T, C = x.shape
        k = v1.shape[2]
        if k != 2 or v1.shape[0] != C or v1.shape[1] != C:
            raise _C.S2agError("TCN block kernel supports kernel_size=2 and n_inputs == n_outputs")
        st = _stream(x)
        w1, w2 = _empty((C, k, C), x), _empty((C, k, C), x)
        n1, n2 = _empty((C,), x), _empty((C,), x)
        out = _empty(x.shape, x)
        nonce = seed_nonce(x.device) if p > 0 else None
        n_ws = 0 if (_C.is_emulated() or not TCN_FUSED[0]) else _C.lib().s2ag_tcn_fused_ws_floats(T, C, dilation)
        if n_ws > 0:
            # one kernel per block (csrc/umma_tcn.cu): y1 stays in shared memory; it and y2 reach HBM only when a
            # backward pass will read them
            need_bwd = any(ctx.needs_input_grad)
            y1 = _empty(x.shape, x) if need_bwd else None
            y2 = _empty(x.shape, x) if need_bwd else None
            ws = _empty((n_ws,), x)
            _C.call("s2ag_tcn_block_fused_fwd", _p(x), _p(v1), _p(g1), _p(b1), _p(v2), _p(g2), _p(b2), _p(w1), _p(w2),
                    _p(n1), _p(n2), _p(y1), _p(y2), _p(out), _p(ws), B, T, C, dilation, float(p), ctypes.c_uint64(seed),
                    _p(nonce), st)
        else:
            _C.call("s2ag_weight_norm_fwd", _p(v1), _p(g1), _p(w1), _p(n1), C, C, k, st)
            _C.call("s2ag_weight_norm_fwd", _p(v2), _p(g2), _p(w2), _p(n2), C, C, k, st)
            y1, y2 = _empty(x.shape, x), _empty(x.shape, x)
            _C.call("s2ag_tcn_block_fwd", _p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(y1), _p(y2), _p(out), B, T, C,
                    dilation, float(p), ctypes.c_uint64(seed), _p(nonce), st)
        ctx.t = (x, y1, y2, out.detach(), w1, w2, n1, n2, v1, g1, b1, v2, g2, b2)
        ctx.cfg = (B, T, C, k, dilation, float(p))
        ctx.want_w = v1.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y1, y2, out, w1, w2, n1, n2, v1, g1, b1, v2, g2, b2 = ctx.t
        B, T, C, k, dilation, p = ctx.cfg
        st = _stream(dout)
        dout = dout.contiguous()
        dx = _empty(x.shape, x)
        ws = _empty((2, B * T * C), x)
        train_w = ctx.want_w
        side = _side_stream_tcn[0] if (dout.is_cuda and train_w) else None
        cur = torch.cuda.current_stream(dout.device) if side is not None else None
        if side is not None and cur != side:
            # data path here; the weight gradients (split-K GEMMs, bias sums, weight_norm backward) on the conv
            # weight-gradient stream, out of the text encoder's data-gradient chain
            _C.call("s2ag_tcn_block_bwd_phased", _p(dout), _p(x), _p(y1), _p(y2), _p(out), _p(w1), _p(w2), _p(dx), None,
                    None, None, None, _p(ws), B, T, C, dilation, p, 1, st)
            ev = torch.cuda.Event()
            ev.record(cur)
            side.wait_event(ev)
            wst = _handle(side, dout.device)
            with torch.cuda.stream(side):
                dw = torch.zeros((2, C, k, C), dtype=torch.float32, device=x.device)
            for t_ in (ws, x, y1, w1, w2):
                t_.record_stream(side)
            _C.call("s2ag_tcn_block_bwd_phased", None, _p(x), _p(y1), None, None, _p(w1), _p(w2), None, _p(dw[0]),
                    _p(_grad_of(b1)), _p(dw[1]), _p(_grad_of(b2)), _p(ws), B, T, C, dilation, p, 2, wst)
        else:
            wst = st
            dw = torch.zeros((2, C, k, C), dtype=torch.float32, device=x.device)
            db1 = _grad_of(b1) if train_w else torch.zeros_like(b1)
            db2 = _grad_of(b2) if train_w else torch.zeros_like(b2)
            _C.call("s2ag_tcn_block_bwd", _p(dout), _p(x), _p(y1), _p(y2), _p(out), _p(w1), _p(w2), _p(dx), _p(dw[0]),
                    _p(db1), _p(dw[1]), _p(db2), _p(ws), B, T, C, dilation, p, st)
        if train_w:
            _C.call("s2ag_weight_norm_bwd", _p(dw[0]), _p(v1), _p(g1), _p(n1), _p(_grad_of(v1)), _p(_grad_of(g1)), C, C,
                    k, wst)
            _C.call("s2ag_weight_norm_bwd", _p(dw[1]), _p(v2), _p(g2), _p(n2), _p(_grad_of(v2)), _p(_grad_of(g2)), C, C,
                    k, wst)
        return (dx,) + (None,) * 9


def tcn_block(x, v1, g1, b1, v2, g2, b2, dilation, p, training):
    p = p if training else 0.0
    return TcnBlockFn.apply(x, v1, g1, b1, v2, g2, b2, dilation, p, next_seed() if p > 0 else 0)


# ------------------------------------------------------------------------------------------ bidirectional multi-layer GRU
class BiGruFn(torch.autograd.Function):
    """nn.GRU(num_layers=L, bidirectional=True, batch_first=True, dropout=p), h0 = 0.
    `x` is the full [B,T,In] input buffer; `pieces` are the differentiable tensors that were
    written into column slices of it (so the concat is never materialised and the backward hands
    each producer a view of dx).  Output [B,T,2H], or [B,T,H] = fwd+rev when sum_halves.
    params: flat list per layer of (w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r)."""

    @staticmethod
    def forward(ctx, x, slices, nlayers, H, p, training, sum_halves, *rest):
        params = rest[:8 * nlayers]
        pieces = rest[8 * nlayers:]
        _check(x, *params)
        B, T, In0 = x.shape
        x2, ldx = _rows(x, In0)
        st = _stream(x)
        need_bwd = any(ctx.needs_input_grad)
        gi_ws = _empty((_C.lib().s2ag_gru_fwd_ws_floats(B, T, H),), x)
        layers = []
        cur, ldcur, In = x2, ldx, In0
        for l in range(nlayers):
            wif, whf, bif, bhf, wir, whr, bir, bhr = params[8 * l:8 * l + 8]
            out = _empty((B, T, 2 * H), x)
            gates = _empty((B * T * 2 * 4 * H,), x) if need_bwd else None
            _C.call("s2ag_gru_layer_fwd", _p(cur), ldcur, _p(wif), _p(wir), _p(bif), _p(bir), _p(whf), _p(whr), _p(bhf),
                    _p(bhr), _p(gi_ws), _p(out), _p(gates), B, T, In, H, st)
            rec = {"x": cur, "ldx": ldcur, "In": In, "out": out.detach(), "gates": gates, "drop": None}
            nxt = out
            if training and p > 0 and l < nlayers - 1:
                seed = next_seed()
                nonce = seed_nonce(x.device)
                nxt = torch.empty_like(out)
                _C.call("s2ag_dropout", _p(out), _p(nxt), out.numel(), float(p), ctypes.c_uint64(seed), _p(nonce), st)
                rec["drop"] = (float(p), seed, nonce)
            layers.append(rec)
            cur, ldcur, In = nxt, 2 * H, 2 * H
            last = out  # (ctx keeps a detached alias; the returned tensor must be a distinct object)
        if sum_halves:
            y = _empty((B, T, H), x)
            _C.call("s2ag_add_halves", _p(last), _p(y), B * T, H, st)
        else:
            y = last
        if need_bwd:
            ctx.layers, ctx.params = layers, params
            ctx.cfg = (B, T, In0, H, nlayers, sum_halves, slices, len(pieces))
            ctx.want_w = params[0].requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        B, T, In0, H, nlayers, sum_halves, slices, npieces = ctx.cfg
        st = _stream(dy)
        M = B * T
        side = _side_stream[0] if dy.is_cuda else None
        main = torch.cuda.current_stream(dy.device) if side is not None else None
        ws_floats = _C.lib().s2ag_gru_bwd_ws_floats(B, T, H)
        ws = None if side is not None else _empty((ws_floats,), dy)
        if sum_halves:
            d, ldd = _rows(dy, H)
            dstride = 0
        else:
            d, ldd = _rows(dy, 2 * H)
            dstride = H
        dx = None
        for l in range(nlayers - 1, -1, -1):
            rec = ctx.layers[l]
            wif, whf, bif, bhf, wir, whr, bir, bhr = ctx.params[8 * l:8 * l + 8]
            need_dx = l > 0 or npieces > 0 or ctx.needs_input_grad[0]
            dxl = _empty((B, T, rec["In"]), dy) if need_dx else None
            want_w = ctx.want_w
            gw = [_p(_grad_of(q)) for q in (wif, wir, bif, bir, whf, whr, bhf, bhr)] if want_w else [None] * 8
            args = [_p(d), ldd, dstride, _p(rec["x"]), rec["ldx"], _p(rec["out"]), _p(rec["gates"]),
                    _p(wif), _p(wir), _p(whf), _p(whr), _p(dxl), rec["In"], *gw]
            if side is not None and want_w:
                # recurrence + dx on the main stream; the time-batched weight-gradient GEMMs of this layer on the
                # side stream, overlapping the next (lower) layer's latency-bound BPTT kernel.  Per-layer workspace.
                wsl = _empty((ws_floats,), dy)
                _C.call("s2ag_gru_layer_bwd", *args, _p(wsl), B, T, rec["In"], H, 3, st)
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                _C.call("s2ag_gru_layer_bwd", *args, _p(wsl), B, T, rec["In"], H, 4,
                        _handle(side, dy.device))
                for t_ in (wsl, rec["x"], rec["out"], rec["gates"], d):
                    if t_ is not None:
                        t_.record_stream(side)
            else:
                if ws is None:
                    ws = _empty((ws_floats,), dy)
                _C.call("s2ag_gru_layer_bwd", *args, _p(ws), B, T, rec["In"], H, 7, st)
            if l > 0:
                drop = ctx.layers[l - 1]["drop"]
                if drop is not None:
                    pp, seed, nonce = drop
                    _C.call("s2ag_dropout", _p(dxl), _p(dxl), dxl.numel(), pp, ctypes.c_uint64(seed), _p(nonce), st)
                d, ldd, dstride = dxl, 2 * H, H
            else:
                dx = dxl
        grads = [dx if ctx.needs_input_grad[0] else None, None, None, None, None, None, None]
        grads += [None] * (8 * nlayers)
        for (a, b_) in slices:
            grads.append(dx[:, :, a:b_])
        return tuple(grads)


def bigru(x, gru_params, nlayers, H, p, training, sum_halves=False, pieces=None, slices=None):
    """x: [B,T,In] input buffer.  pieces/slices: the differentiable producers of its column ranges
    (default: x itself as one piece)."""
    if pieces is None:
        pieces, slices = (x,), ((0, x.shape[-1]),)
    return BiGruFn.apply(x.detach(), tuple(slices), nlayers, H, p, training, sum_halves, *gru_params, *pieces)


# ------------------------------------------------------------------------------------------ speaker z
class ReparamTileFn(torch.autograd.Function):
    """z = mu + eps*exp(0.5*logvar) and its tiling over T into columns [off, off+Z) of `dst`
    (net/embedding_net.py:10-13; net/multimodal_context_net_v2.py:536-539).  Returns (z, dst_slice)."""

    @staticmethod
    def forward(ctx, mu, logvar, eps, dst, off):
        dst = dst.t
        _check(mu, logvar, eps, dst)
        B, Z = mu.shape
        if tuple(eps.shape) != (B, Z) or tuple(logvar.shape) != (B, Z):
            raise _C.S2agError("re-parametrisation noise must be [%d, %d], got %s" % (B, Z, tuple(eps.shape)))
        mu, logvar, eps = mu.contiguous(), logvar.contiguous(), eps.contiguous()
        z = _empty((B, Z), mu)
        T = dst.shape[1]
        d2, ld = _rows(dst, dst.shape[-1])
        _C.call("s2ag_reparam_tile_fwd", _p(mu), _p(logvar), _p(eps), _p(z), _p(d2), ld, off, B, T, Z, _stream(mu))
        ctx.t = (logvar, eps)
        ctx.cfg = (B, T, Z, off)
        return z, col_slice(dst, off, off + Z)

    @staticmethod
    def backward(ctx, dz, dsl):
        logvar, eps = ctx.t
        B, T, Z, off = ctx.cfg
        dmu, dlv = _empty((B, Z), logvar), _empty((B, Z), logvar)
        if dsl is not None:
            d2, ld = _rows(dsl, Z)
        else:
            d2, ld = None, 0
        dz = dz.contiguous() if dz is not None else None
        _C.call("s2ag_reparam_tile_bwd", _p(d2), ld, 0, _p(dz), _p(logvar), _p(eps), _p(dmu), _p(dlv), B, T, Z,
                _stream(logvar))
        return dmu, dlv, None, None, None


def reparam_tile(mu, logvar, eps, dst, off):
    return ReparamTileFn.apply(mu, logvar, eps, Out(dst), off)


# ------------------------------------------------------------------------------------------ discriminator head
class DHeadFn(torch.autograd.Function):
    """sigmoid(out2(out(fwd+rev)))  (net/multimodal_context_net_v2.py:579-585)"""

    @staticmethod
    def forward(ctx, g, w1, b1, w2, b2):
        _check(g, w1, b1, w2, b2)
        g = g.contiguous()
        B, T, H2 = g.shape
        H = H2 // 2
        lin1 = _empty((B, T), g)
        out = _empty((B, 1), g)
        _C.call("s2ag_dhead_fwd", _p(g), _p(w1), _p(b1), _p(w2), _p(b2), _p(lin1), _p(out), B, T, H, _stream(g))
        ctx.t = (g, lin1, out.detach(), w1, b1, w2, b2)
        ctx.want_w = w1.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        g, lin1, out, w1, b1, w2, b2 = ctx.t
        B, T, H2 = g.shape
        dout = dout.contiguous()
        dg = _empty(g.shape, g) if ctx.needs_input_grad[0] else None
        if ctx.want_w:
            gw1, gb1, gw2, gb2 = _grad_of(w1), _grad_of(b1), _grad_of(w2), _grad_of(b2)
        else:
            gw1, gb1, gw2, gb2 = torch.zeros_like(w1), torch.zeros_like(b1), torch.zeros_like(w2), torch.zeros_like(b2)
        _C.call("s2ag_dhead_bwd", _p(dout), _p(out), _p(g), _p(lin1), _p(w1), _p(w2), _p(dg), _p(gw1), _p(gb1), _p(gw2),
                _p(gb2), B, T, H2 // 2, _stream(g))
        return dg, None, None, None, None


def dhead(g, w1, b1, w2, b2):
    return DHeadFn.apply(g, w1, b1, w2, b2)


# ------------------------------------------------------------------------------------------ losses / optimiser / metric
def dis_loss(d_real, d_fake, loss_out, want_grads=True):
    """processor_v2.py:811.  loss_out: float[1] device slot.  -> (g_real, g_fake)"""
    B = d_real.shape[0]
    gr = torch.empty_like(d_real) if want_grads else None
    gf = torch.empty_like(d_fake) if want_grads else None
    _C.call("s2ag_dis_loss", _p(d_real.contiguous()), _p(d_fake.contiguous()), _p(loss_out), _p(gr), _p(gf), B,
            _stream(d_real))
    return gr, gf


def gen_loss(out, tgt, out_rand, z, z_rand, mu, logvar, dis_out, weights, losses_out, want_grads=True):
    """processor_v2.py:893-937.  weights = (w_huber, w_kld, w_div, w_gan); losses_out: float[5]
    (huber, gen, kld, div_reg, total).  -> (g_out, g_dis, g_mu, g_logvar)"""
    B = out.shape[0]
    TP = out.numel() // B
    Z = 0 if z is None else z.shape[1]
    out, tgt = out.contiguous(), tgt.contiguous()
    g_out = torch.empty_like(out) if want_grads else None
    g_dis = torch.empty_like(dis_out) if (want_grads and dis_out is not None) else None
    g_mu = torch.empty_like(mu) if (want_grads and mu is not None) else None
    g_lv = torch.empty_like(logvar) if (want_grads and logvar is not None) else None
    c = lambda t: None if t is None else t.contiguous()
    _C.call("s2ag_gen_loss", _p(out), _p(tgt), _p(c(out_rand)), _p(c(z)), _p(c(z_rand)), _p(c(mu)), _p(c(logvar)),
            _p(c(dis_out)), float(weights[0]), float(weights[1]), float(weights[2]), float(weights[3]),
            _p(losses_out), _p(g_out), _p(g_dis), _p(g_mu), _p(g_lv), B, TP, Z, _stream(out))
    return g_out, g_dis, g_mu, g_lv


def l1_mean(a, b, dst):
    a, b = a.contiguous(), b.contiguous()
    _C.call("s2ag_l1_mean", _p(a), _p(b), _p(dst), a.numel(), _stream(a))


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step_count, grad_scale=1.0):
    """Flat-buffer Adam (processor_v2.py:215-220).  step_count: int32[1] device tensor."""
    _C.call("s2ag_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
            float(grad_scale), _p(step_count), _stream(p))


class AttentionFn(torch.autograd.Function):
    """sigmoid-MLP score + softmax over time + weighted sum (net/ser_att_conv_rnn_v2.py:30-34), forward and backward."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        _check(x, w1, b1, w2, b2)
        x = x.contiguous()
        N, T, Hd = x.shape
        A = w1.shape[0]
        out = _empty((N, Hd), x)
        alpha = _empty((N, T, 1), x)
        w1c, w2c = w1.contiguous(), w2.contiguous()
        _C.call("s2ag_attention_fwd", _p(x), _p(w1c), _p(b1), _p(w2c), _p(b2), _p(out), _p(alpha), N, T, Hd, A,
                _stream(x))
        ctx.t = (x, w1, b1, w2, b2, w1c, w2c, alpha.detach())
        ctx.want_w = w1.requires_grad
        return out, alpha

    @staticmethod
    def backward(ctx, d_out, d_alpha):
        x, w1, b1, w2, b2, w1c, w2c, alpha = ctx.t
        N, T, Hd = x.shape
        A = w1.shape[0]
        dx = _empty(x.shape, x)
        if ctx.want_w:
            gw1, gb1, gw2, gb2 = _grad_of(w1), _grad_of(b1), _grad_of(w2), _grad_of(b2)
        else:
            gw1, gb1, gw2, gb2 = torch.zeros_like(w1), torch.zeros_like(b1), torch.zeros_like(w2), torch.zeros_like(b2)
        if d_out is None:
            d_out = torch.zeros(N, Hd, dtype=torch.float32, device=x.device)
        _C.call("s2ag_attention_bwd", _p(x), _p(w1c), _p(b1), _p(w2c), _p(alpha), _p(d_out.contiguous()),
                _p(None if d_alpha is None else d_alpha.contiguous()), _p(dx), _p(gw1), _p(gb1), _p(gw2), _p(gb2),
                N, T, Hd, A, _stream(x))
        return dx, None, None, None, None


def attention(x, w1, b1, w2, b2):
    """-> (out[N,Hd], alphas[N,T,1]); differentiable w.r.t. x and (accumulating into .grad) the four parameters."""
    return AttentionFn.apply(x, w1, b1, w2, b2)


# ------------------------------------------------------------------------------------------ input front-end
_MFCC_CONST = {}


def mfcc_features(audio, sr=16000, num_mfcc=14):
    """utils/common.py:340-349 `get_mfcc_features` for a batch of raw-audio chunks on the device:
    audio [B, L] fp32 -> [B, 3*num_mfcc - 5, 1 + L // 512] (37 x 71 for the reference's 36 266-sample chunks)."""
    from .utils import audio_features as af
    _check(audio)
    audio, lda = _rows(audio, audio.shape[-1])
    B, L = audio.shape
    key = (str(audio.device), sr, num_mfcc)
    if key not in _MFCC_CONST:
        bank, span = af.mel_bank(sr)
        _MFCC_CONST[key] = (torch.from_numpy(bank).to(audio.device), torch.from_numpy(span).to(audio.device),
                            torch.from_numpy(af.dct_rows(num_mfcc)).to(audio.device))
    bank, span, dct = _MFCC_CONST[key]
    F = 1 + L // af.HOP
    ws = _empty((B, F, af.N_MELS), audio)
    out = _empty((B, 3 * num_mfcc - 5, F), audio)
    _C.call("s2ag_mfcc_features", _p(audio), lda, B, L, af.HOP, _p(bank), _p(span), af.N_MELS, _p(dct), num_mfcc,
            float(af.TOP_DB), 1.0 / 1000.0, _p(ws), _p(out), _stream(audio))
    return out


def expand_inputs(audio_i16=None, audio_max=None, mfcc_f16=None, audio_out=None, mfcc_out=None):
    """processor_v2.py:606-610 on the device: int16 audio * audio_max / 32767 -> fp32, fp16 MFCC -> fp32.
    The compressed tensors are what crosses PCIe.  -> (audio fp32 or None, mfcc fp32 or None)"""
    ref = audio_i16 if audio_i16 is not None else mfcc_f16
    B = L = 0
    is64 = 0
    if audio_i16 is not None:
        assert audio_i16.dtype == torch.int16 and audio_i16.is_contiguous()
        B, L = audio_i16.shape
        assert audio_max.dtype in (torch.float32, torch.float64) and audio_max.numel() == B
        is64 = 1 if audio_max.dtype == torch.float64 else 0
        if audio_out is None:
            audio_out = torch.empty((B, L), dtype=torch.float32, device=ref.device)
    n = 0
    if mfcc_f16 is not None:
        assert mfcc_f16.dtype == torch.float16 and mfcc_f16.is_contiguous()
        n = mfcc_f16.numel()
        if mfcc_out is None:
            mfcc_out = torch.empty(mfcc_f16.shape, dtype=torch.float32, device=ref.device)
    _C.call("s2ag_expand_inputs", _p(audio_i16), _p(audio_max), is64, _p(audio_out if audio_i16 is not None else None),
            B, L, _p(mfcc_f16), _p(mfcc_out if mfcc_f16 is not None else None), n,
            _stream(ref) if ref.is_cuda else None)
    return (audio_out if audio_i16 is not None else None), (mfcc_out if mfcc_f16 is not None else None)


# ------------------------------------------------------------------------------------------ long-form pipeline / metrics
def longform_blend(out, result, chunk, n_pre, pre_next=None, n_chunks=None):
    """processor_v2.py:1282-1327 for a batch in lock-step (see s2ag_longform_blend)."""
    B, T, P = out.shape
    assert result.is_contiguous() and result.shape[0] == B and result.shape[2] == P
    _C.call("s2ag_longform_blend", _p(out.contiguous()), _p(result), result.shape[1] * P, _p(pre_next), _p(n_chunks),
            int(chunk), B, T, P, int(n_pre), _stream(out))


def fade_out(seq, lengths, start_frames, n_smooth):
    """processor_v2.py:1334-1391 on [B, Lmax, P] sequences; lengths/start_frames int32 [B].  -> new lengths int32 [B]"""
    assert seq.is_contiguous()
    B, Lmax, P = seq.shape
    new_len = torch.empty_like(lengths)
    _C.call("s2ag_fade_out", _p(seq), Lmax * P, _p(lengths), _p(start_frames), _p(new_len), B, P, int(n_smooth), Lmax,
            _stream(seq))
    return new_len


def dir_vec_to_pose(vec, mean=None):
    """utils/ted_db_utils.py:81-102 (+ mean direction vector): [..., 27] -> [..., 10, 3]"""
    vec = vec.contiguous()
    N = vec.numel() // 27
    pose = _empty(tuple(vec.shape[:-1]) + (10, 3), vec)
    _C.call("s2ag_dir_vec_to_pose", _p(vec), _p(mean), _p(pose), N, _stream(vec))
    return pose


def pose_metrics(out, target, mean, n_pre, dst=None):
    """processor_v2.py:738-774: -> float32[3] device tensor (L1, joint MAE, accel difference)"""
    out, target = out.contiguous(), target.contiguous()
    B, T, P = out.shape
    assert P == 27
    if dst is None:
        dst = _empty((3,), out)
    ws = torch.empty(3, dtype=torch.float64, device=out.device)
    _C.call("s2ag_pose_metrics", _p(out), _p(target), _p(mean), _p(ws), _p(dst), B, T, int(n_pre), _stream(out))
    return dst


# ------------------------------------------------------------------------------------------ Frechet gesture distance
def fgd_new_accumulator(device, D=32):
    """zeroed fp64 moment buffer of s2ag_fgd_accumulate (net/embedding_space_evaluator.py:28-40: the four host lists)"""
    return torch.zeros(int(_C.lib().s2ag_fgd_acc_doubles(int(D))), dtype=torch.float64, device=device)


def fgd_accumulate(acc, gen_feat, real_feat):
    """net/embedding_space_evaluator.py:45-57 without the D2H copies: fold [N, D] feature pairs into `acc`."""
    _check(gen_feat, real_feat)
    N, D = gen_feat.shape
    assert real_feat.shape == gen_feat.shape and gen_feat.stride(1) == 1 and real_feat.stride(1) == 1
    _C.call("s2ag_fgd_accumulate", _p(gen_feat), gen_feat.stride(0), _p(real_feat), real_feat.stride(0), N, D, _p(acc),
            _stream(gen_feat))


def fgd_scores(acc, D=32):
    """net/embedding_space_evaluator.py:73-101 -> float64[2] device tensor (frechet_dist, feat_dist)"""
    out = torch.empty(2, dtype=torch.float64, device=acc.device)
    _C.call("s2ag_fgd_scores", _p(acc), int(D), _p(out), _stream(acc))
    return out


def frechet_distance(mu1, sigma1, mu2, sigma2):
    """net/embedding_space_evaluator.py:104-152 for given moments (fp64 device tensors) -> float64[2], [0] = distance"""
    D = mu1.numel()
    ts = [t.to(torch.float64).contiguous() for t in (mu1, sigma1, mu2, sigma2)]
    out = torch.empty(2, dtype=torch.float64, device=ts[0].device)
    _C.call("s2ag_frechet_distance", _p(ts[0]), _p(ts[1]), _p(ts[2]), _p(ts[3]), int(D), _p(out), _stream(ts[0]))
    return out
