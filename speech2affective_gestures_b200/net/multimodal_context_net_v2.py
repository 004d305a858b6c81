"""Drop-in mirror of the reference's net/multimodal_context_net_v2.py for the GAN-step hot path.

Same class names, constructor arguments, forward() signatures/returns and state_dict keys/shapes
(SURVEY.md section 8b), so `load_state_dict` of a reference checkpoint works.  The nn.Conv*/nn.BatchNorm*/
nn.Linear/nn.GRU/nn.Embedding objects below are PARAMETER CONTAINERS ONLY: their forward() is never
called.  All compute goes through `ops.*`, i.e. the hand-written sm_100a kernels of
libs2ag_b200.so behind the C ABI in include/s2ag.h.  There is no PyTorch or CPU fallback.

B200-first design decisions (DESIGN.md has the detail):
  * activations are channels-last, so the reference's permute/transpose/contiguous/view shuffles
    (:33, :53, :89, :155-173, :420-422) disappear: the AffEncoder regrouping is folded into the
    BatchNorm apply kernel's column map, MFCCEncoder's Linear-over-last-axis is a transposed GEMM;
  * the GRU input is never concatenated (:526, :539): every encoder writes its features straight
    into its column range of one [B, T, in_size] buffer;
  * parameters of a network live in ONE flat fp32 buffer and gradients in another (views are
    exposed as the usual nn.Parameters): one memset zeroes the gradients, one kernel runs Adam,
    one NCCL all-reduce averages them across ranks.
"""
import os

import torch
import torch.nn as nn

from .. import _C, ops
from ..utils import ted_db_utils as ted_db
from . import embedding_net as en
from .tcn import TemporalConvNet
from .utils.graph import Graph
from .utils.tgcn import STGraphConv


# --------------------------------------------------------------------------------------------- infrastructure
def _sync_bn_counters(module, *_):
    """num_batches_tracked is kept as a host counter on the hot path (no kernel launch per BN call)
    and folded into the int64 buffer whenever a state_dict is taken."""
    for m in module.modules():
        n = getattr(m, "_s2ag_batches", 0)
        if n and getattr(m, "num_batches_tracked", None) is not None:
            m.num_batches_tracked += n
            m._s2ag_batches = 0


def _reset_bn_counters(module, incompatible_keys):
    """a loaded state_dict carries its own num_batches_tracked: pending host-side increments are void"""
    for m in module.modules():
        if hasattr(m, "_s2ag_batches"):
            m._s2ag_batches = 0


class FlatParamNet(nn.Module):
    """Base of the four top-level networks: owns the flat parameter / gradient buffers."""

    def __init__(self):
        super().__init__()
        self._flat = None
        self._flat_grad = None
        self.register_state_dict_pre_hook(_sync_bn_counters)
        self.register_load_state_dict_post_hook(_reset_bn_counters)

    def flatten_parameters_(self):
        """(Re)pack every parameter into one contiguous fp32 buffer and alias .data/.grad onto it."""
        params = []
        seen = set()
        for p in self.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        if not params:
            return self
        dev = params[0].device
        n = sum((p.numel() + 7) // 8 * 8 for p in params)  # every tensor 32-byte aligned (one LDG.E.256 per operand item)
        flat = torch.zeros(n, dtype=torch.float32, device=dev)
        grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            flat[off:off + k].copy_(p.data.reshape(-1))
            if p.grad is not None:
                grad[off:off + k].copy_(p.grad.reshape(-1))
            p.data = flat[off:off + k].view(p.shape)
            p.grad = grad[off:off + k].view(p.shape)
            off += (k + 7) // 8 * 8
        self._flat, self._flat_grad = flat, grad
        return self

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        if self._flat is not None:
            self.flatten_parameters_()
        return self

    @property
    def flat_params(self):
        if self._flat is None:
            self.flatten_parameters_()
        return self._flat

    @property
    def flat_grads(self):
        if self._flat_grad is None:
            self.flatten_parameters_()
        return self._flat_grad

    def zero_grad(self, set_to_none=False):
        if self._flat_grad is not None:
            self._flat_grad.zero_()
        else:
            super().zero_grad(set_to_none=False)


def _gru_param_list(gru):
    ps = []
    for l in range(gru.num_layers):
        for sfx in ("", "_reverse"):
            ps += [getattr(gru, "weight_ih_l%d%s" % (l, sfx)), getattr(gru, "weight_hh_l%d%s" % (l, sfx)),
                   getattr(gru, "bias_ih_l%d%s" % (l, sfx)), getattr(gru, "bias_hh_l%d%s" % (l, sfx))]
    return ps


def _conv1d_geom(conv):
    return (conv.stride[0], 1, conv.padding[0], 0, conv.dilation[0], 1)


# --------------------------------------------------------------------------------------------- encoders
class WavEncoder(nn.Module):
    """Raw-audio strided Conv1d stack (reference :14-33).  in [B, n_samples] -> out [B, 34, 32]."""

    def __init__(self):
        super().__init__()
        self.feat_extractor = nn.Sequential(
            nn.Conv1d(1, 16, 15, stride=5, padding=1600),
            nn.BatchNorm1d(16),
            nn.LeakyReLU(0.3, inplace=True),
            nn.Conv1d(16, 32, 15, stride=6),
            nn.BatchNorm1d(32),
            nn.LeakyReLU(0.3, inplace=True),
            nn.Conv1d(32, 64, 15, stride=6),
            nn.BatchNorm1d(64),
            nn.LeakyReLU(0.3, inplace=True),
            nn.Conv1d(64, 32, 15, stride=6),
        )

    def forward(self, wav_data, out=None):
        f = self.feat_extractor
        if not _C.is_emulated() and not (torch.is_grad_enabled() and any(q.requires_grad for q in f.parameters())) \
                and os.environ.get("S2AG_WAV_FUSED", "1") != "0":
            # frozen on the hot path (PoseGeneratorTriModal): conv1 recomputed per tile, BatchNorm + LeakyReLU applied while
            # the next convolution stages its operand -- no normalised activation is ever written to HBM
            return ops.wavencoder_fwd(wav_data, [f[0], f[3], f[6], f[9]], [f[1], f[4], f[7]], 0.3, out=out)
        x = wav_data.unsqueeze(-1)  # channels-last [B, L, 1]
        for ci in (0, 3, 6):
            x = ops.conv_bn_act(x, f[ci].weight, f[ci].bias, _conv1d_geom(f[ci]), bn=f[ci + 1], act=ops.ACT_LEAKY,
                                slope=0.3)
        return ops.conv_bn_act(x, f[9].weight, f[9].bias, _conv1d_geom(f[9]), out=out)  # already (batch, seq, dim)


class MFCCEncoder(nn.Module):
    """reference :36-58.  in_mfcc [B, num_mfcc(37), mfcc_length(71)] -> [B, time_steps(34), 32].
    The reference permutes to [B, 71, 37] and convolves along the 37 axis; channels-last that is
    the input as given."""

    def __init__(self, mfcc_length, num_mfcc, time_steps):
        super().__init__()
        self.conv1 = nn.Conv1d(mfcc_length, 64, 5, padding=2)
        self.batch_norm1 = nn.BatchNorm1d(64)
        self.conv2 = nn.Conv1d(64, 64, 5, padding=2)
        self.batch_norm2 = nn.BatchNorm1d(64)
        self.conv3 = nn.Conv1d(64, 48, 3, padding=1)
        self.batch_norm3 = nn.BatchNorm1d(48)
        self.conv4 = nn.Conv1d(48, time_steps, 3, padding=1)
        self.batch_norm4 = nn.BatchNorm1d(time_steps)
        self.linear1 = nn.Linear(num_mfcc, 32)
        self.activation = nn.LeakyReLU(0.3, inplace=True)

    def forward(self, mfcc_data, out=None):
        x = mfcc_data
        for conv, bn in ((self.conv1, self.batch_norm1), (self.conv2, self.batch_norm2),
                         (self.conv3, self.batch_norm3), (self.conv4, self.batch_norm4)):
            x = ops.conv_bn_act(x, conv.weight, conv.bias, _conv1d_geom(conv), bn=bn, act=ops.ACT_LEAKY, slope=0.3)
        # x: [B, 37, time_steps]; Linear acts over the 37 axis
        return ops.linear_t(x, self.linear1.weight, self.linear1.bias, ops.ACT_LEAKY, 0.3, out=out)


class TextEncoderTCN(nn.Module):
    """reference :61-91: Embedding -> Dropout -> 4 TemporalBlocks -> Linear(300 -> 32)."""

    def __init__(self, args, n_words, embed_size=300, pre_trained_embedding=None,
                 kernel_size=2, dropout=0.3, emb_dropout=0.1):
        super().__init__()
        if pre_trained_embedding is not None:
            assert pre_trained_embedding.shape[0] == n_words
            assert pre_trained_embedding.shape[1] == embed_size
            self.embedding = nn.Embedding.from_pretrained(torch.FloatTensor(pre_trained_embedding),
                                                          freeze=args.freeze_wordembed)
        else:
            self.embedding = nn.Embedding(n_words, embed_size)
        num_channels = [args.hidden_size] * args.n_layers
        self.tcn = TemporalConvNet(embed_size, num_channels, kernel_size, dropout=dropout)
        self.decoder = nn.Linear(num_channels[-1], 32)
        self.drop = nn.Dropout(emb_dropout)
        self.emb_dropout = emb_dropout
        self.init_weights()

    def init_weights(self):
        self.decoder.bias.data.fill_(0)
        self.decoder.weight.data.normal_(0, 0.01)

    def forward(self, in_data, out=None):
        emb = ops.embedding(in_data, self.embedding.weight, self.drop.p if self.training else 0.0)
        y = self.tcn.forward_cl(emb)
        y = ops.linear(y, self.decoder.weight, self.decoder.bias, out=out)
        return y, 0


class AffEncoder(nn.Module):
    """reference :94-175: two ST-GCN blocks + BatchNorm1d regroupings + two Conv1d.
    poses [N, T, 27] -> [N, T, 8]."""

    def __init__(self, coords=3):
        super().__init__()
        self.coords = coords
        self.num_dir_vec_pairs = len(ted_db.dir_vec_pairs)
        graph1 = Graph(self.num_dir_vec_pairs, ted_db.dir_edge_pairs, strategy='spatial', max_hop=2)
        self.register_buffer("A1", torch.tensor(graph1.A, dtype=torch.float32), persistent=False)
        self.num_body_parts = len(ted_db.body_parts_edge_idx)
        graph2 = Graph(self.num_body_parts, ted_db.body_parts_edge_pairs, strategy='spatial', max_hop=2)
        self.register_buffer("A2", torch.tensor(graph2.A, dtype=torch.float32), persistent=False)

        self.st_gcn1 = STGraphConv(coords, 16, self.A1.size(0), (9, 5), stride=(1, 1), padding=(4, 2))
        self.batch_norm1 = nn.BatchNorm1d(16 * self.num_dir_vec_pairs)
        self.st_gcn2 = STGraphConv(48, 16, self.A2.size(0), (9, 3), stride=(1, 1), padding=(4, 1))
        self.batch_norm2 = nn.BatchNorm1d(16 * self.num_body_parts)
        self.conv3 = nn.Conv1d(48, 16, 5, padding=2)
        self.batch_norm3 = nn.BatchNorm1d(16)
        self.conv4 = nn.Conv1d(16, 8, 3, padding=1)
        self.batch_norm4 = nn.BatchNorm1d(8)
        self.activation = nn.LeakyReLU(inplace=True)

        # Column bookkeeping that replaces the reference's view/permute/zero-pad shuffles (:155-173).
        # Channels-last columns of st_gcn1's output are (v, c) -> v*16 + c.
        F1, V1, V2 = 16, self.num_dir_vec_pairs, self.num_body_parts
        part_of = {}
        for idx, edges in enumerate(ted_db.body_parts_edge_idx):
            for j, v in enumerate(edges):
                part_of[int(v)] = (idx, j)
        assert len(part_of) == V1 and all(len(e) == ted_db.max_body_part_edges for e in ted_db.body_parts_edge_idx)
        W2 = ted_db.max_body_part_edges * F1  # 48 channels per body part
        pm1, cm1 = [], []
        for v in range(V1):
            for c in range(F1):
                pm1.append(c * V1 + v)                       # BatchNorm1d(144) channel index is c*9 + v (:155-156)
                idx, j = part_of[v]
                cm1.append(idx * W2 + c * ted_db.max_body_part_edges + j)  # feat2_in[n,t,c*3+j,idx] (:161-163)
        pm2, cm2 = [], []
        for v in range(V2):
            for c in range(F1):
                pm2.append(c * V2 + v)                       # BatchNorm1d(48) channel index c*3 + v (:166-167)
                cm2.append(c * V2 + v)                       # conv3 sees channels ordered c*3 + v (:168)
        for name, val in (("pm1", pm1), ("cm1", cm1), ("pm2", pm2), ("cm2", cm2)):
            self.register_buffer(name, torch.tensor(val, dtype=torch.int32), persistent=False)

    def forward(self, poses, out=None):
        n, t, jc = poses.shape
        x = poses.reshape(n, t, jc // self.coords, self.coords)            # [N,T,V=9,C=3]
        f1, _ = self.st_gcn1(x, self.A1)                                    # [N,T,9,16]
        f1 = ops.bn_act(f1.view(n, t, -1), self.batch_norm1, cmap=self.cm1, pmap=self.pm1)
        f2, _ = self.st_gcn2(f1.view(n, t, self.num_body_parts, -1), self.A2)   # [N,T,3,16]
        f2 = ops.bn_act(f2.view(n, t, -1), self.batch_norm2, cmap=self.cm2, pmap=self.pm2)   # [N,T,48]
        f3 = ops.conv_bn_act(f2, self.conv3.weight, self.conv3.bias, _conv1d_geom(self.conv3), bn=self.batch_norm3,
                             act=ops.ACT_LEAKY, slope=0.01)
        return ops.conv_bn_act(f3, self.conv4.weight, self.conv4.bias, _conv1d_geom(self.conv4), bn=self.batch_norm4,
                               act=ops.ACT_LEAKY, slope=0.01, out=out)


# --------------------------------------------------------------------------------------------- generators
class _SpeakerMixin:
    def _build_speaker(self, z_obj):
        self.speaker_embedding = None
        if z_obj:
            self.z_size = 16
            self.in_size += self.z_size
            if z_obj.__class__.__name__ == 'Vocab':
                self.speaker_embedding = nn.Sequential(nn.Embedding(z_obj.n_words, self.z_size),
                                                       nn.Linear(self.z_size, self.z_size))
                self.speaker_mu = nn.Linear(self.z_size, self.z_size)
                self.speaker_log_var = nn.Linear(self.z_size, self.z_size)

    def _speaker_z(self, in_text, vid_indices, buf, off, eps=None):
        """-> (z_context, z_mu, z_log_var, tiled-slice-or-None); writes z tiled over T into buf[:, :, off:].
        eps: re-parametrisation noise drawn by the caller (to fix the order of the draws), else drawn here."""
        if not self.z_obj:
            return None, None, None, None
        if self.speaker_embedding:
            assert vid_indices is not None
            e = ops.embedding(vid_indices, self.speaker_embedding[0].weight, 0.0)
            zc = ops.linear(e, self.speaker_embedding[1].weight, self.speaker_embedding[1].bias)
            z_mu = ops.linear(zc, self.speaker_mu.weight, self.speaker_mu.bias)
            z_log_var = ops.linear(zc, self.speaker_log_var.weight, self.speaker_log_var.bias)
            z, tiled = ops.reparam_tile(z_mu, z_log_var, eps if eps is not None else en.draw_eps(z_mu), buf, off)
            return z, z_mu, z_log_var, tiled
        # random-noise style vector (:519-520): mu = 0, log_var = 0, eps = randn
        zeros = torch.zeros(in_text.shape[0], self.z_size, device=buf.device)
        eps = torch.randn(in_text.shape[0], self.z_size, device=buf.device)
        z, tiled = ops.reparam_tile(zeros, zeros, eps, buf, off)
        return z.detach(), None, None, tiled.detach()


class PoseGeneratorTriModal(FlatParamNet, _SpeakerMixin):
    """Frozen tri-modal baseline (reference :247-343), run forward once per GAN step."""

    def __init__(self, args, pose_dim, n_words, word_embed_size, word_embeddings, z_obj=None):
        super().__init__()
        self.pre_length = args.n_pre_poses
        self.gen_length = args.n_poses - args.n_pre_poses
        self.z_obj = z_obj
        self.input_context = args.input_context
        self.pose_in = pose_dim + 1
        if self.input_context == 'both':
            self.in_size = 32 + 32 + pose_dim + 1
        elif self.input_context == 'none':
            self.in_size = pose_dim + 1
        else:
            self.in_size = 32 + pose_dim + 1
        self.audio_encoder = WavEncoder()
        self.text_encoder = TextEncoderTCN(args, n_words, word_embed_size, pre_trained_embedding=word_embeddings,
                                           dropout=args.dropout_prob)
        self._build_speaker(z_obj)
        self.hidden_size = args.hidden_size
        self.gru = nn.GRU(self.in_size, hidden_size=self.hidden_size, num_layers=args.n_layers, batch_first=True,
                          bidirectional=True, dropout=args.dropout_prob)
        self.out = nn.Sequential(
            nn.Linear(self.hidden_size, self.hidden_size // 2),
            nn.LeakyReLU(True),  # sic (:285): `True` binds to negative_slope = 1.0, i.e. the identity
            nn.Linear(self.hidden_size // 2, pose_dim)
        )
        self.do_flatten_parameters = False
        self.flatten_parameters_()

    def encode_inputs(self, in_text, in_audio):
        """WavEncoder + TextEncoderTCN features (the part of forward() that depends only on the inputs); lets a
        driver run them ahead of / beside other work and hand them to forward(..., pre=...)."""
        a = self.audio_encoder(in_audio) if self.input_context in ('both', 'audio') else None
        t = self.text_encoder(in_text)[0] if self.input_context in ('both', 'text') else None
        return a, t

    def forward(self, pre_seq, in_text, in_audio, vid_indices=None, pre=None, eps=None):
        B, T = pre_seq.shape[0], pre_seq.shape[1]
        buf = torch.empty(B, T, self.in_size, dtype=torch.float32, device=pre_seq.device)
        pieces, slices = [], []
        col = self.pose_in
        buf[:, :, :col].copy_(pre_seq)
        if pre_seq.requires_grad:
            raise NotImplementedError("gradient w.r.t. pre_seq is not on the reference path")
        if self.input_context in ('both', 'audio'):
            if pre is None:
                a = self.audio_encoder(in_audio, out=ops.col_slice(buf, col, col + 32))
            else:
                a = pre[0]
                buf[:, :, col:col + 32].copy_(a.detach())
            pieces.append(a); slices.append((col, col + 32)); col += 32
        if self.input_context in ('both', 'text'):
            if pre is None:
                t, _ = self.text_encoder(in_text, out=ops.col_slice(buf, col, col + 32))
            else:
                t = pre[1]
                buf[:, :, col:col + 32].copy_(t.detach())
            if self.input_context == 'both':
                assert a.shape[1] == t.shape[1]
            pieces.append(t); slices.append((col, col + 32)); col += 32
        z, z_mu, z_log_var, tiled = self._speaker_z(in_text, vid_indices, buf, col, eps=eps)
        if tiled is not None and tiled.requires_grad:
            pieces.append(tiled); slices.append((col, col + self.z_size))
        g = ops.bigru(buf, _gru_param_list(self.gru), self.gru.num_layers, self.hidden_size, self.gru.dropout,
                      self.training, sum_halves=True, pieces=pieces, slices=slices)
        h = ops.linear(g, self.out[0].weight, self.out[0].bias)          # LeakyReLU(1.0) == identity
        y = ops.linear(h, self.out[2].weight, self.out[2].bias)
        return y, z, z_mu, z_log_var


class PoseGenerator(FlatParamNet, _SpeakerMixin):
    """The trained generator (reference :438-546)."""

    def __init__(self, args, pose_dim, n_words, word_embed_size, word_embeddings,
                 mfcc_length, num_mfcc, time_steps, z_obj=None):
        super().__init__()
        self.pre_length = args.n_pre_poses
        self.gen_length = args.n_poses - args.n_pre_poses
        self.z_obj = z_obj
        self.input_context = args.input_context
        self.mfcc_feature_length = 32
        self.text_feature_length = 32
        self.pose_feature_length = 8
        if self.input_context == 'both':
            self.in_size = self.mfcc_feature_length + self.text_feature_length + self.pose_feature_length
        elif self.input_context == 'audio':
            self.in_size = self.mfcc_feature_length + self.pose_feature_length
        elif self.input_context == 'text':
            self.in_size = self.text_feature_length + self.pose_feature_length
        elif self.input_context == 'none':
            self.in_size = self.pose_feature_length
        else:
            raise AssertionError(self.input_context)
        self.audio_encoder = MFCCEncoder(mfcc_length, num_mfcc, time_steps)
        self.text_encoder = TextEncoderTCN(args, n_words, word_embed_size, pre_trained_embedding=word_embeddings,
                                           dropout=args.dropout_prob)
        self.aff_encoder = AffEncoder()
        self._build_speaker(z_obj)
        self.hidden_size = args.hidden_size_s2eg
        self.gru = nn.GRU(self.in_size, hidden_size=self.hidden_size, num_layers=args.n_layers, batch_first=True,
                          bidirectional=True, dropout=args.dropout_prob)
        self.out = nn.Sequential(
            nn.Linear(self.hidden_size, self.hidden_size // 2),
            nn.LeakyReLU(inplace=True),
            nn.Linear(self.hidden_size // 2, pose_dim)
        )
        self.do_flatten_parameters = False
        self.flatten_parameters_()

    def encode_shared(self, pre_seq, in_mfcc, repeats=1, mfcc_stream=None):
        """The dropout-free, BatchNorm-only encoders (AffEncoder on the seed poses, MFCCEncoder): within one GAN
        iteration the reference evaluates them on the SAME inputs with the SAME weights in each of its generator
        passes (processor_v2.py:798, 823, 909), so they are computed once and handed to every pass through
        `forward(..., shared=...)`; `repeats` = number of passes they stand for (BatchNorm running statistics are
        advanced as that many identical updates, see ops.bn_repeat)."""
        with ops.bn_repeat(repeats):
            a = None
            if self.input_context in ('both', 'audio'):
                if mfcc_stream is not None and in_mfcc.is_cuda:
                    # the two encoders are independent: the MFCC encoder runs on `mfcc_stream` beside the AffEncoder
                    # (autograd runs its backward there as well); the caller's stream joins before returning
                    cur = torch.cuda.current_stream(in_mfcc.device)
                    ev = torch.cuda.Event(); ev.record(cur)
                    mfcc_stream.wait_event(ev)
                    with torch.cuda.stream(mfcc_stream):
                        a = self.audio_encoder(in_mfcc)
                    done = torch.cuda.Event(); done.record(mfcc_stream)
                    in_mfcc.record_stream(mfcc_stream)
                    a.record_stream(cur)
                else:
                    a = self.audio_encoder(in_mfcc)
            p = self.aff_encoder(pre_seq[..., :-1])
            if a is not None and mfcc_stream is not None and in_mfcc.is_cuda:
                cur.wait_event(done)
        return p, a

    def encode_text(self, in_text):
        """TextEncoderTCN features [B, T, 32] for forward(..., text_feat=...): depends only on the tokens and the
        weights, so a driver may evaluate it on a side stream while the recurrent kernels of another pass run."""
        return self.text_encoder(in_text)[0]

    def forward(self, pre_seq, in_text, in_mfcc, vid_indices=None, shared=None, text_feat=None, eps=None):
        B, T = pre_seq.shape[0], pre_seq.shape[1]
        buf = torch.empty(B, T, self.in_size, dtype=torch.float32, device=pre_seq.device)
        pieces, slices = [], []
        col = 0
        if shared is None:
            p = self.aff_encoder(pre_seq[..., :-1], out=ops.col_slice(buf, col, col + 8))
        else:
            p = shared[0]
            buf[:, :, col:col + 8].copy_(p.detach())
        pieces.append(p); slices.append((col, col + 8)); col += 8
        if self.input_context in ('both', 'audio'):
            if shared is None:
                a = self.audio_encoder(in_mfcc, out=ops.col_slice(buf, col, col + 32))
            else:
                a = shared[1]
                buf[:, :, col:col + 32].copy_(a.detach())
            pieces.append(a); slices.append((col, col + 32)); col += 32
        if self.input_context in ('both', 'text'):
            if text_feat is None:
                t, _ = self.text_encoder(in_text, out=ops.col_slice(buf, col, col + 32))
            else:
                t = text_feat
                buf[:, :, col:col + 32].copy_(t.detach())
            if self.input_context == 'both':
                assert a.shape[1] == t.shape[1], \
                    'Audio and text features must have the same number of time steps. ' \
                    'Found time steps: audio features: {}, text features: {}.'.format(a.shape[1], t.shape[1])
            pieces.append(t); slices.append((col, col + 32)); col += 32
        z, z_mu, z_log_var, tiled = self._speaker_z(in_text, vid_indices, buf, col, eps=eps)
        if tiled is not None and tiled.requires_grad:
            pieces.append(tiled); slices.append((col, col + self.z_size))
        g = ops.bigru(buf, _gru_param_list(self.gru), self.gru.num_layers, self.hidden_size, self.gru.dropout,
                      self.training, sum_halves=True, pieces=pieces, slices=slices)
        h = ops.linear(g, self.out[0].weight, self.out[0].bias, ops.ACT_LEAKY, 0.01)
        y = ops.linear(h, self.out[2].weight, self.out[2].bias)
        return y, z, z_mu, z_log_var


# --------------------------------------------------------------------------------------------- discriminators
class AffDiscriminator(FlatParamNet):
    """reference :549-585: AffEncoder -> 4-layer bi-GRU(64) -> Linear(64->1) -> Linear(34->1) -> sigmoid."""

    def __init__(self, input_size, coords=3):
        super().__init__()
        self.input_size = input_size
        self.coords = coords
        self.hidden_size = 64
        self.aff_encoder = AffEncoder(coords=coords)
        self.gru = nn.GRU(8, hidden_size=self.hidden_size, num_layers=4, bidirectional=True, dropout=0.3,
                          batch_first=True)
        self.out = nn.Linear(self.hidden_size, 1)
        self.out2 = nn.Linear(34, 1)
        self.activation = nn.LeakyReLU(inplace=True)
        self.do_flatten_parameters = False
        self.flatten_parameters_()

    def forward(self, poses, in_text=None):
        feat = self.aff_encoder(poses)
        g = ops.bigru(feat, _gru_param_list(self.gru), 4, self.hidden_size, self.gru.dropout, self.training)
        return ops.dhead(g, self.out.weight, self.out.bias, self.out2.weight, self.out2.bias)

    def forward_pair(self, poses_a, poses_b, feat_a=None):
        """D(poses_a), D(poses_b) as the reference computes them back to back with the same weights
        (processor_v2.py:808-809).  Nothing but BatchNorm couples the samples of a batch, so both batches go through
        every kernel together: the AffEncoder with per-call BatchNorm statistics (two statistics groups, running
        statistics updated in call order) and ONE latency-bound recurrent launch per GRU layer."""
        if feat_a is not None:
            # `feat_a` = self.aff_encoder(poses_a) evaluated EARLIER by the caller (D(target) depends on nothing the
            # generator produces; the caller orders it before this call, so BatchNorm's running statistics still
            # advance in call order): only the second batch goes through the encoder here
            feat = torch.cat([feat_a, self.aff_encoder(poses_b)], dim=0)
        else:
            # one pass over both stacked batches; BatchNorm keeps the two calls' statistics apart (ops.bn_groups)
            with ops.bn_groups(2):
                feat = self.aff_encoder(torch.cat([poses_a, poses_b], dim=0))
        g = ops.bigru(feat, _gru_param_list(self.gru), 4, self.hidden_size, self.gru.dropout, self.training)
        o = ops.dhead(g, self.out.weight, self.out.bias, self.out2.weight, self.out2.bias)
        n = poses_a.shape[0]
        return o[:n], o[n:]


class ConvDiscriminatorTriModal(FlatParamNet):
    """reference :390-435 (== ConvDiscriminator of the _abl_aff ablation): three valid Conv1d
    (BN, identity "LeakyReLU(True)") -> bi-GRU(64) -> Linear(64->1) -> Linear(28->1) -> sigmoid."""

    def __init__(self, input_size):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = 64
        self.pre_conv = nn.Sequential(
            nn.Conv1d(input_size, 16, 3),
            nn.BatchNorm1d(16),
            nn.LeakyReLU(True),
            nn.Conv1d(16, 8, 3),
            nn.BatchNorm1d(8),
            nn.LeakyReLU(True),
            nn.Conv1d(8, 8, 3),
        )
        self.gru = nn.GRU(8, hidden_size=self.hidden_size, num_layers=4, bidirectional=True, dropout=0.3,
                          batch_first=True)
        self.out = nn.Linear(self.hidden_size, 1)
        self.out2 = nn.Linear(28, 1)
        self.do_flatten_parameters = False
        self.flatten_parameters_()

    def forward(self, poses, in_text=None):
        f = self.pre_conv
        x = ops.conv_bn_act(poses, f[0].weight, f[0].bias, _conv1d_geom(f[0]), bn=f[1])
        x = ops.conv_bn_act(x, f[3].weight, f[3].bias, _conv1d_geom(f[3]), bn=f[4])
        feat = ops.conv_bn_act(x, f[6].weight, f[6].bias, _conv1d_geom(f[6]))
        g = ops.bigru(feat, _gru_param_list(self.gru), 4, self.hidden_size, self.gru.dropout, self.training)
        return ops.dhead(g, self.out.weight, self.out.bias, self.out2.weight, self.out2.bias)
