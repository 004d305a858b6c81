"""Spatial-temporal graph convolution block, drop-in for the reference's net/utils/tgcn.py
(ConvTemporalGraphical :15-71, STGraphConv :133-218) on channels-last [N, T, V, C] activations.

Submodule/parameter names mirror the reference so state_dicts interchange: gcn.conv (Conv2d
(k_t x 1)), tcn = Sequential(BN2d, ReLU, Conv2d(k_t x k_s), BN2d, Dropout), residual =
Sequential(Conv2d 1x1, BN2d).  The nn.* objects are parameter containers only; compute goes
through ops.* (libs2ag_b200.so)."""
import torch.nn as nn

import os

from ... import ops

COMPOSED = os.environ.get("S2AG_GCN_COMPOSED", "1") != "0"   # A/B switch: 0 = two-kernel route (conv, then contraction)


class ConvTemporalGraphical(nn.Module):
    def __init__(self, in_channels, out_channels, A_channels, temporal_kernel_size, temporal_stride=1,
                 temporal_padding=0, temporal_dilation=1, bias=True):
        super().__init__()
        if temporal_stride != 1:
            raise NotImplementedError("temporal stride 1 only (the reference never uses another)")
        self.conv = nn.Conv2d(in_channels, out_channels * A_channels, kernel_size=(temporal_kernel_size, 1),
                              padding=(temporal_padding, 0), stride=(temporal_stride, 1),
                              dilation=(temporal_dilation, 1), bias=bias)
        self.geom = (1, 1, temporal_padding, 0, temporal_dilation, 1)

    def forward(self, x, A):
        """x [N,T,V,Cin] -> [N,T,V,Cout];  einsum('nkctv,kvw->nctw') of the reference (:66-69)."""
        if COMPOSED:
            # conv and adjacency contraction composed into one temporal convolution over [N, T, V*Cin] rows: the
            # [N, T, V, K*Cout] intermediate (25 MB at 256 clips) is never formed (ops.GcnFn)
            return ops.gcn_conv(x, self.conv.weight, self.conv.bias, A, self.geom[2]), A
        y = ops.conv_bn_act(x, self.conv.weight, self.conv.bias, self.geom)
        return ops.graph_contract(y, A), A


class STGraphConv(nn.Module):
    def __init__(self, in_channels, out_channels, A_channels, kernel_size, stride=(1, 1), padding=(0, 0),
                 dropout=0, activation='LeakyRelU', residual=True):
        super().__init__()
        assert len(kernel_size) == 2 and kernel_size[0] % 2 == 1
        if tuple(stride) != (1, 1) or dropout != 0 or not residual:
            raise NotImplementedError("configuration outside the reference's AffEncoder usage")
        self.gcn = ConvTemporalGraphical(in_channels, out_channels, A_channels, kernel_size[0],
                                         temporal_stride=stride[0], temporal_padding=padding[0])
        self.tcn = nn.Sequential(
            nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size, stride, padding),
            nn.BatchNorm2d(out_channels),
            nn.Dropout(dropout, inplace=True),
        )
        # the reference always takes the conv+BN residual branch here (tgcn.py:195 compares a tuple with 1)
        self.residual = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=stride),
                                      nn.BatchNorm2d(out_channels))
        self.tcn_geom = (1, 1, padding[0], padding[1], 1, 1)
        if activation.lower() == 'leakyrelu':
            self.act = (ops.ACT_LEAKY, 0.01)
        elif activation.lower() == 'relu':
            self.act = (ops.ACT_RELU, 0.0)
        else:
            raise ValueError(activation)

    def forward(self, x, A):
        res = ops.conv_bn_act(x, self.residual[0].weight, self.residual[0].bias, (1, 1, 0, 0, 1, 1),
                              bn=self.residual[1])
        g, A = self.gcn(x, A)
        h = ops.bn_act(g, self.tcn[0], ops.ACT_RELU)
        h = ops.conv_bn_act(h, self.tcn[2].weight, self.tcn[2].bias, self.tcn_geom)
        out = ops.bn_act(h, self.tcn[3], self.act[0], self.act[1], add=res)
        return out, A
