"""re_parametrize of the reference's net/embedding_net.py:10-13 (the only function of that file on
the hot path).  `eps_source` lets the parity harness inject the noise the reference drew."""
import torch

eps_source = None  # callable(std_like) -> eps, or None for torch.randn_like


def draw_eps(like):
    if eps_source is not None:
        return eps_source(like)
    return torch.randn_like(like)


def re_parametrize(mu, log_var):
    """z = mu + eps * exp(0.5 * log_var); differentiable fused kernel (no tiling)."""
    from .. import ops
    buf = None
    z, _ = ops.reparam_tile(mu, log_var, draw_eps(mu), mu.new_empty((mu.shape[0], 1, mu.shape[1])), 0)
    return z


# ------------------------------------------------------------------------------------------------------------------
# The pose auto-encoder behind the Frechet gesture distance.  Reference: net/embedding_net.py:16-36
# (conv_norm_relu), :39-82 (PoseEncoderConv), :164-216 (PoseDecoderConv), :262-314 (EmbeddingNet).  The evaluator only
# ever builds mode='pose' and runs it in eval mode (net/embedding_space_evaluator.py:23-27), so that is what exists here:
# same containers => same state_dict keys as the reference checkpoint `outputs/embedding_net.pth.tar`
# ('embedding_dict'), forward passes on the library's channels-last kernels.
import torch.nn as nn  # noqa: E402


def conv_norm_relu(in_channels, out_channels, down_sample=False, padding=0, batch_norm=True):
    k, s = (4, 2) if down_sample else (3, 1)
    layers = [nn.Conv1d(in_channels, out_channels, kernel_size=k, stride=s, padding=padding)]
    if batch_norm:
        layers.append(nn.BatchNorm1d(out_channels))   # the reference builds it either way, registers it only here
    layers.append(nn.LeakyReLU(0.2, True))
    return nn.Sequential(*layers)


def _geom(conv):
    return (conv.stride[0], 1, conv.padding[0], 0, conv.dilation[0], 1)


class _Cached:
    """A re-laid-out copy of a parameter, rebuilt when the parameter changes (load_state_dict / .to / in-place update)."""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, p, fn):
        key = (p.data_ptr(), p._version, p.device)
        if key != self.key:
            with torch.no_grad():
                self.key, self.val = key, fn(p).contiguous()
        return self.val


class PoseEncoderConv(nn.Module):
    """poses [B, 34, dim] -> (z, mu, log_var) [B, 32] each (reference :39-82).  The conv stack runs channels-last, so
    the input needs no transpose; the reference flattens [B, 32 ch, 12] channel-major, here the first Linear's columns
    are permuted once instead (w'[:, l*32 + c] = w[:, c*12 + l])."""

    def __init__(self, length, dim):
        super().__init__()
        self.net = nn.Sequential(
            conv_norm_relu(dim, 32, batch_norm=True),
            conv_norm_relu(32, 64, batch_norm=True),
            conv_norm_relu(64, 64, True, batch_norm=True),
            nn.Conv1d(64, 32, 3),
        )
        self.out_net = nn.Sequential(
            nn.Linear(384, 256),  # 34 frames -> 12 positions x 32 channels
            nn.BatchNorm1d(256),
            nn.LeakyReLU(True),   # negative_slope=True == 1.0: the identity, as in the reference
            nn.Linear(256, 128),
            nn.BatchNorm1d(128),
            nn.LeakyReLU(True),
            nn.Linear(128, 32),
        )
        self.fc_mu = nn.Linear(32, 32)
        self.fc_log_var = nn.Linear(32, 32)
        self._w0 = _Cached()

    def forward(self, poses, variational_encoding):
        from .. import ops
        x = poses
        for blk in (self.net[0], self.net[1], self.net[2]):
            x = ops.conv_bn_act(x, blk[0].weight, blk[0].bias, _geom(blk[0]), bn=blk[1], act=ops.ACT_LEAKY, slope=0.2)
        x = ops.conv_bn_act(x, self.net[3].weight, self.net[3].bias, _geom(self.net[3]))   # [B, L, 32]
        B, L, C = x.shape
        lin0 = self.out_net[0]
        w0 = self._w0.get(lin0.weight, lambda w: w.view(-1, C, L).transpose(1, 2).reshape(-1, L * C))
        x = ops.linear(x.view(B, L * C), w0, lin0.bias)
        x = ops.bn_act(x, self.out_net[1], ops.ACT_LEAKY, float(self.out_net[2].negative_slope))
        x = ops.linear(x, self.out_net[3].weight, self.out_net[3].bias)
        x = ops.bn_act(x, self.out_net[4], ops.ACT_LEAKY, float(self.out_net[5].negative_slope))
        x = ops.linear(x, self.out_net[6].weight, self.out_net[6].bias)
        mu = ops.linear(x, self.fc_mu.weight, self.fc_mu.bias)
        log_var = ops.linear(x, self.fc_log_var.weight, self.fc_log_var.bias)
        z = re_parametrize(mu, log_var) if variational_encoding else mu
        return z, mu, log_var


class PoseDecoderConv(nn.Module):
    """latent [B, 32] -> poses [B, length, dim] (reference :164-216).  ConvTranspose1d(k=3, stride 1) is the stride-1
    convolution with flipped taps, swapped channel axes and padding 2; the [B, 4, L] view of the pre-net output is taken
    channels-last by permuting the rows of its last Linear."""

    def __init__(self, length, dim, use_pre_poses=False):
        super().__init__()
        self.use_pre_poses = use_pre_poses
        feat_size = 32
        if use_pre_poses:
            self.pre_pose_net = nn.Sequential(nn.Linear(dim * 4, 32), nn.BatchNorm1d(32), nn.ReLU(), nn.Linear(32, 32))
            feat_size += 32
        hidden, width = {64: (128, 256), 34: (64, 136)}[length]
        self.pre_net = nn.Sequential(nn.Linear(feat_size, hidden), nn.BatchNorm1d(hidden), nn.LeakyReLU(True),
                                     nn.Linear(hidden, width))
        self.net = nn.Sequential(
            nn.ConvTranspose1d(4, 32, 3), nn.BatchNorm1d(32), nn.LeakyReLU(0.2, True),
            nn.ConvTranspose1d(32, 32, 3), nn.BatchNorm1d(32), nn.LeakyReLU(0.2, True),
            nn.Conv1d(32, 32, 3),
            nn.Conv1d(32, dim, 3),
        )
        self._wl, self._bl, self._wt0, self._wt1 = _Cached(), _Cached(), _Cached(), _Cached()

    def forward(self, feat, pre_poses=None):
        from .. import ops
        if self.use_pre_poses:
            p = self.pre_pose_net
            h = ops.linear(pre_poses.reshape(pre_poses.shape[0], -1), p[0].weight, p[0].bias)
            h = ops.bn_act(h, p[1], ops.ACT_RELU)
            feat = torch.cat((ops.linear(h, p[3].weight, p[3].bias), feat), dim=1)
        pn = self.pre_net
        x = ops.linear(feat, pn[0].weight, pn[0].bias)
        x = ops.bn_act(x, pn[1], ops.ACT_LEAKY, float(pn[2].negative_slope))
        L = pn[3].out_features // 4
        wl = self._wl.get(pn[3].weight, lambda w: w.view(4, L, -1).transpose(0, 1).reshape(4 * L, -1))
        bl = self._bl.get(pn[3].bias, lambda b: b.view(4, L).t().reshape(-1))
        x = ops.linear(x, wl, bl).view(feat.shape[0], L, 4)                      # channels-last [B, L, 4]
        for ci, cache in ((0, self._wt0), (3, self._wt1)):
            ct = self.net[ci]
            w = cache.get(ct.weight, lambda w: w.transpose(0, 1).flip(2))        # [Cin, Cout, K] -> conv [Cout, Cin, K]
            x = ops.conv_bn_act(x, w, ct.bias, (1, 1, ct.kernel_size[0] - 1, 0, 1, 1), bn=self.net[ci + 1],
                                act=ops.ACT_LEAKY, slope=0.2)
        x = ops.conv_bn_act(x, self.net[6].weight, self.net[6].bias, _geom(self.net[6]))
        return ops.conv_bn_act(x, self.net[7].weight, self.net[7].bias, _geom(self.net[7]))   # [B, length, dim]


class EmbeddingNet(nn.Module):
    """mode='pose' only (what EmbeddingSpaceEvaluator builds, net/embedding_space_evaluator.py:23-27); the speech-context
    branch (ContextEncoder / PoseDecoderGRU, reference :218-260) belongs to the embedding-net TRAINING pipeline, which is
    outside the hot path.  Same forward signature and 7-tuple as the reference (:279-314)."""

    def __init__(self, args, pose_dim, n_frames, n_words, word_embed_size, word_embeddings, mode):
        super().__init__()
        if mode != 'pose':
            raise NotImplementedError("only the pose auto-encoder of the FGD evaluator is on this path (mode='pose')")
        self.context_encoder = None
        self.pose_encoder = PoseEncoderConv(n_frames, pose_dim)
        self.decoder = PoseDecoderConv(n_frames, pose_dim)
        self.mode = mode

    def forward(self, in_text, in_audio, pre_poses, poses, input_mode=None, variational_encoding=False):
        input_mode = self.mode if input_mode is None else input_mode
        if input_mode != 'pose':
            raise NotImplementedError("input_mode %r needs the speech-context encoder" % (input_mode,))
        poses_feat, pose_mu, pose_log_var = self.pose_encoder(poses, variational_encoding)
        out_poses = self.decoder(poses_feat, pre_poses)
        return None, None, None, poses_feat, pose_mu, pose_log_var, out_poses

    def freeze_pose_nets(self):
        for param in list(self.pose_encoder.parameters()) + list(self.decoder.parameters()):
            param.requires_grad = False
