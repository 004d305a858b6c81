"""Causal dilated temporal convolution network, drop-in for the reference's net/tcn.py.

Same classes, constructor arguments and state_dict keys (including the old-style weight_norm
`weight_g`/`weight_v` pairs and the `net.0.*` / `net.4.*` aliases that come from the reference
building `self.net = nn.Sequential(self.conv1, ...)`, net/tcn.py:31), but activations are
channels-last [B, T, C] and a whole residual block runs as two fused conv-as-GEMM launches of
libs2ag_b200.so (weight-norm folded, chomp never materialised, bias+ReLU+dropout+residual+ReLU in the
epilogue).  `forward` accepts/returns the reference's [B, C, T] layout; `forward_cl` is the
channels-last entry the text encoder uses.
"""
import torch
import torch.nn as nn

from .. import ops


class Chomp1d(nn.Module):
    """Kept for structural (state_dict index) parity with net/tcn.py:7-13; the fused kernel never
    computes the chomped columns, so this module has no work to do."""

    def __init__(self, chomp_size):
        super().__init__()
        self.chomp_size = chomp_size

    def forward(self, x):
        return x[:, :, :-self.chomp_size].contiguous()


class _WeightNormConv1d(nn.Module):
    """Parameter container with the reference's weight_norm(nn.Conv1d) parameter names/shapes:
    bias [Co], weight_g [Co,1,1], weight_v [Co,Ci,k] (net/tcn.py:19-20)."""

    def __init__(self, n_in, n_out, kernel_size, stride, padding, dilation):
        super().__init__()
        conv = nn.Conv1d(n_in, n_out, kernel_size, stride=stride, padding=padding, dilation=dilation)
        self.bias = nn.Parameter(conv.bias.detach().clone())
        v = conv.weight.detach().clone()
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(-1, 1, 1))
        self.weight_v = nn.Parameter(v)
        self.kernel_size, self.stride, self.padding, self.dilation = kernel_size, stride, padding, dilation


class TemporalBlock(nn.Module):
    def __init__(self, n_inputs, n_outputs, kernel_size, stride, dilation, padding, dropout=0.2):
        super().__init__()
        if n_inputs != n_outputs or kernel_size != 2 or stride != 1 or padding != (kernel_size - 1) * dilation:
            raise NotImplementedError(
                "the fused TCN kernel covers the configuration the reference instantiates "
                "(n_inputs == n_outputs, kernel_size 2, stride 1, causal padding): "
                "net/multimodal_context_net_v2.py:75-76")
        self.conv1 = _WeightNormConv1d(n_inputs, n_outputs, kernel_size, stride, padding, dilation)
        self.chomp1 = Chomp1d(padding)
        self.relu1 = nn.ReLU()
        self.dropout1 = nn.Dropout(dropout)
        self.conv2 = _WeightNormConv1d(n_outputs, n_outputs, kernel_size, stride, padding, dilation)
        self.chomp2 = Chomp1d(padding)
        self.relu2 = nn.ReLU()
        self.dropout2 = nn.Dropout(dropout)
        self.net = nn.Sequential(self.conv1, self.chomp1, self.relu1, self.dropout1,
                                 self.conv2, self.chomp2, self.relu2, self.dropout2)
        self.downsample = None
        self.relu = nn.ReLU()
        self.dilation = dilation

    def forward_cl(self, x):
        c1, c2 = self.conv1, self.conv2
        return ops.tcn_block(x, c1.weight_v, c1.weight_g, c1.bias, c2.weight_v, c2.weight_g, c2.bias,
                             self.dilation, self.dropout1.p, self.training)

    def forward(self, x):
        return self.forward_cl(x.transpose(1, 2)).transpose(1, 2)


class TemporalConvNet(nn.Module):
    def __init__(self, num_inputs, num_channels, kernel_size=2, dropout=0.2):
        super().__init__()
        layers = []
        for i, out_channels in enumerate(num_channels):
            dilation_size = 2 ** i
            in_channels = num_inputs if i == 0 else num_channels[i - 1]
            layers.append(TemporalBlock(in_channels, out_channels, kernel_size, stride=1, dilation=dilation_size,
                                        padding=(kernel_size - 1) * dilation_size, dropout=dropout))
        self.network = nn.Sequential(*layers)

    def forward_cl(self, x):
        for blk in self.network:
            x = blk.forward_cl(x)
        return x

    def forward(self, x):
        return self.forward_cl(x.transpose(1, 2)).transpose(1, 2)
