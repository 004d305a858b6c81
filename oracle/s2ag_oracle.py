"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

A plain-PyTorch CPU restatement of the reference algorithm for the GAN-step hot path
(UttaranB127/speech2affective_gestures @ 013cc487).  It works on the reference's state_dict
layout ([N,C,L] / [N,C,T,V] activations, PyTorch weight layouts) with torch.nn.functional ops and an
explicit GRU recurrence, so it shares NO code and NO data layout with the CUDA path it checks.

Parity pin: the reference has no tests / golden vectors for this path, so the pin is
the reference code itself, executed in the build container by oracle/gen_golden.py
(unmodified reference modules and the unmodified Processor.forward_pass_s2ag, imported from
/root/reference with stubbed third-party imports).  gen_golden.py asserts this oracle equals the
reference on seeded inputs and commits the reference's outputs as fixtures under tests/golden/;
tests/test_oracle_golden.py re-checks the oracle against those fixtures wherever the tests run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.

Each function cites the reference file:line it follows (paths relative to the reference root).
`sd` is a dict name -> tensor in the reference's state_dict naming; running statistics in `sd` are
updated in place when training=True, like nn.BatchNorm.
"""
import numpy as np
import torch
import torch.nn.functional as F

# utils/ted_db_utils.py:16-19
DIR_EDGE_PAIRS = [(0, 1), (1, 2), (0, 3), (3, 4), (4, 5), (0, 6), (6, 7), (7, 8)]
BODY_PARTS_EDGE_IDX = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
MAX_BODY_PART_EDGES = 3
BODY_PARTS_EDGE_PAIRS = [(0, 1), (0, 2)]


# ------------------------------------------------------------------ net/utils/graph.py:26-129
def spatial_adjacency(num_nodes, links, max_hop=2):
    """Graph(num_nodes, links, strategy='spatial', max_hop).A restated with matrix powers."""
    A = np.zeros((num_nodes, num_nodes))
    for i in range(num_nodes):
        A[i, i] = 1
    for a, b in links:
        A[a, b] = 1
        A[b, a] = 1
    hop = np.full((num_nodes, num_nodes), np.inf)
    powers = [np.linalg.matrix_power(A, d) > 0 for d in range(max_hop + 1)]
    for d in range(max_hop, -1, -1):
        hop[powers[d]] = d
    adj = (hop <= max_hop).astype(float)
    deg = adj.sum(0)
    norm = adj / np.where(deg > 0, deg, 1)
    out = []
    for h in range(max_hop + 1):
        root = np.zeros_like(adj); close = np.zeros_like(adj); further = np.zeros_like(adj)
        for i in range(num_nodes):
            for j in range(num_nodes):
                if hop[j, i] == h:
                    if hop[j, 0] == hop[i, 0]:
                        root[j, i] = norm[j, i]
                    elif hop[j, 0] > hop[i, 0]:
                        close[j, i] = norm[j, i]
                    else:
                        further[j, i] = norm[j, i]
        if h == 0:
            out.append(root)
        else:
            out.append(root + close)
            out.append(further)
    return torch.tensor(np.stack(out), dtype=torch.float32)


A1 = spatial_adjacency(9, DIR_EDGE_PAIRS)
A2 = spatial_adjacency(3, BODY_PARTS_EDGE_PAIRS)


# ------------------------------------------------------------------ building blocks
def _bn(x, sd, p, training):
    """nn.BatchNorm1d/2d: batch stats + running update (momentum 0.1, eps 1e-5) in train mode."""
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'], sd[p + '.bias'],
                        training, 0.1, 1e-5)


def gru(x, sd, p, num_layers, H):
    """nn.GRU(batch_first, bidirectional), h0 = 0, dropout 0; explicit recurrence."""
    B, T, _ = x.shape
    inp = x
    for l in range(num_layers):
        outs = []
        for sfx, order in (('', range(T)), ('_reverse', range(T - 1, -1, -1))):
            w_ih, w_hh = sd['%s.weight_ih_l%d%s' % (p, l, sfx)], sd['%s.weight_hh_l%d%s' % (p, l, sfx)]
            b_ih, b_hh = sd['%s.bias_ih_l%d%s' % (p, l, sfx)], sd['%s.bias_hh_l%d%s' % (p, l, sfx)]
            h = x.new_zeros(B, H)
            seq = [None] * T
            for t in order:
                gi = inp[:, t] @ w_ih.t() + b_ih
                gh = h @ w_hh.t() + b_hh
                r = torch.sigmoid(gi[:, :H] + gh[:, :H])
                z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
                n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
                h = (1 - z) * n + z * h
                seq[t] = h
            outs.append(torch.stack(seq, 1))
        inp = torch.cat(outs, 2)
    return inp


def wav_encoder(sd, p, wav, training):
    """net/multimodal_context_net_v2.py:14-33"""
    f = p + '.feat_extractor.'
    x = wav.unsqueeze(1)
    x = F.leaky_relu(_bn(F.conv1d(x, sd[f + '0.weight'], sd[f + '0.bias'], stride=5, padding=1600), sd, f + '1', training), 0.3)
    x = F.leaky_relu(_bn(F.conv1d(x, sd[f + '3.weight'], sd[f + '3.bias'], stride=6), sd, f + '4', training), 0.3)
    x = F.leaky_relu(_bn(F.conv1d(x, sd[f + '6.weight'], sd[f + '6.bias'], stride=6), sd, f + '7', training), 0.3)
    x = F.conv1d(x, sd[f + '9.weight'], sd[f + '9.bias'], stride=6)
    return x.transpose(1, 2)


def mfcc_encoder(sd, p, mfcc, training):
    """net/multimodal_context_net_v2.py:36-58"""
    x = mfcc.permute(0, 2, 1)
    for i, pad in ((1, 2), (2, 2), (3, 1), (4, 1)):
        x = F.conv1d(x, sd['%s.conv%d.weight' % (p, i)], sd['%s.conv%d.bias' % (p, i)], padding=pad)
        x = F.leaky_relu(_bn(x, sd, '%s.batch_norm%d' % (p, i), training), 0.3)
    return F.leaky_relu(F.linear(x, sd[p + '.linear1.weight'], sd[p + '.linear1.bias']), 0.3)


def _wn(sd, p):
    """old-style weight_norm, dim=0: w = g * v / ||v|| (norm over (Cin, k))"""
    v, g = sd[p + '.weight_v'], sd[p + '.weight_g']
    return v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))


def text_encoder_tcn(sd, p, in_text, n_layers=4):
    """net/multimodal_context_net_v2.py:61-91 + net/tcn.py:16-64 (dropout 0)"""
    x = F.embedding(in_text, sd[p + '.embedding.weight']).transpose(1, 2)
    for i in range(n_layers):
        d = 2 ** i
        q = '%s.tcn.network.%d' % (p, i)
        y = F.relu(F.conv1d(x, _wn(sd, q + '.conv1'), sd[q + '.conv1.bias'], padding=d, dilation=d)[:, :, :-d])
        y = F.relu(F.conv1d(y, _wn(sd, q + '.conv2'), sd[q + '.conv2.bias'], padding=d, dilation=d)[:, :, :-d])
        x = F.relu(y + x)
    return F.linear(x.transpose(1, 2), sd[p + '.decoder.weight'], sd[p + '.decoder.bias'])


def _st_gcn(sd, p, x, A, ks, training):
    """net/utils/tgcn.py:64-71 (gcn) and :212-218 (block): x [N,C,T,V]"""
    res = _bn(F.conv2d(x, sd[p + '.residual.0.weight'], sd[p + '.residual.0.bias']), sd, p + '.residual.1', training)
    g = F.conv2d(x, sd[p + '.gcn.conv.weight'], sd[p + '.gcn.conv.bias'], padding=(4, 0))
    n, kc, t, v = g.shape
    K = A.shape[0]
    g = torch.einsum('nkctv,kvw->nctw', g.view(n, K, kc // K, t, v), A)
    h = F.relu(_bn(g, sd, p + '.tcn.0', training))
    h = F.conv2d(h, sd[p + '.tcn.2.weight'], sd[p + '.tcn.2.bias'], padding=(4, (ks - 1) // 2))
    h = _bn(h, sd, p + '.tcn.3', training)
    return F.leaky_relu(h + res, 0.01)


def aff_encoder(sd, p, poses, training):
    """net/multimodal_context_net_v2.py:153-175"""
    n, t, jc = poses.shape
    f1 = _st_gcn(sd, p + '.st_gcn1', poses.view(n, t, -1, 3).permute(0, 3, 1, 2), A1, 5, training)
    c1 = f1.shape[1]
    f1 = _bn(f1.permute(0, 1, 3, 2).contiguous().view(n, -1, t), sd, p + '.batch_norm1', training)
    f1 = f1.view(n, -1, 9, t).permute(0, 1, 3, 2)
    f2_in = poses.new_zeros((n, t, MAX_BODY_PART_EDGES * c1, 3))
    for idx, part in enumerate(BODY_PARTS_EDGE_IDX):
        f2_in[..., :c1 * len(part), idx] = f1[..., part].permute(0, 2, 1, 3).contiguous().view(n, t, -1)
    f2 = _st_gcn(sd, p + '.st_gcn2', f2_in.permute(0, 2, 1, 3), A2, 3, training)
    f2 = _bn(f2.permute(0, 1, 3, 2).contiguous().view(n, -1, t), sd, p + '.batch_norm2', training)
    f2 = f2.view(n, -1, 3, t).permute(0, 1, 3, 2)
    f3_in = f2.permute(0, 2, 1, 3).contiguous().view(n, t, -1).permute(0, 2, 1)
    f3 = F.leaky_relu(_bn(F.conv1d(f3_in, sd[p + '.conv3.weight'], sd[p + '.conv3.bias'], padding=2), sd,
                          p + '.batch_norm3', training), 0.01)
    f4 = F.leaky_relu(_bn(F.conv1d(f3, sd[p + '.conv4.weight'], sd[p + '.conv4.bias'], padding=1), sd,
                          p + '.batch_norm4', training), 0.01)
    return f4.permute(0, 2, 1)


def _speaker(sd, vid, eps):
    """net/multimodal_context_net_v2.py:513-516 + net/embedding_net.py:10-13 (eps injected)"""
    zc = F.linear(F.embedding(vid, sd['speaker_embedding.0.weight']), sd['speaker_embedding.1.weight'],
                  sd['speaker_embedding.1.bias'])
    mu = F.linear(zc, sd['speaker_mu.weight'], sd['speaker_mu.bias'])
    lv = F.linear(zc, sd['speaker_log_var.weight'], sd['speaker_log_var.bias'])
    return mu + eps * torch.exp(0.5 * lv), mu, lv


def pose_generator(sd, pre_seq, in_text, in_mfcc, vid, eps, training, H=300, n_layers=4):
    """PoseGenerator.forward, net/multimodal_context_net_v2.py:492-546"""
    audio = mfcc_encoder(sd, 'audio_encoder', in_mfcc, training)
    text = text_encoder_tcn(sd, 'text_encoder', in_text, n_layers)
    z, mu, lv = _speaker(sd, vid, eps)
    pre = aff_encoder(sd, 'aff_encoder', pre_seq[..., :-1], training)
    x = torch.cat((pre, audio, text), 2)
    x = torch.cat((x, z.unsqueeze(1).repeat(1, x.shape[1], 1)), 2)
    o = gru(x, sd, 'gru', n_layers, H)
    o = o[:, :, :H] + o[:, :, H:]
    o = F.linear(F.leaky_relu(F.linear(o.reshape(-1, H), sd['out.0.weight'], sd['out.0.bias']), 0.01),
                 sd['out.2.weight'], sd['out.2.bias'])
    return o.reshape(x.shape[0], x.shape[1], -1), z, mu, lv


def pose_generator_trimodal(sd, pre_seq, in_text, in_audio, vid, eps, training, H=300, n_layers=4):
    """PoseGeneratorTriModal.forward, net/multimodal_context_net_v2.py:293-343 (LeakyReLU(True) == identity, :285)"""
    audio = wav_encoder(sd, 'audio_encoder', in_audio, training)
    text = text_encoder_tcn(sd, 'text_encoder', in_text, n_layers)
    z, mu, lv = _speaker(sd, vid, eps)
    x = torch.cat((pre_seq, audio, text), 2)
    x = torch.cat((x, z.unsqueeze(1).repeat(1, x.shape[1], 1)), 2)
    o = gru(x, sd, 'gru', n_layers, H)
    o = o[:, :, :H] + o[:, :, H:]
    o = F.linear(F.linear(o.reshape(-1, H), sd['out.0.weight'], sd['out.0.bias']), sd['out.2.weight'], sd['out.2.bias'])
    return o.reshape(x.shape[0], x.shape[1], -1), z, mu, lv


def aff_discriminator(sd, poses, training, H=64):
    """AffDiscriminator.forward, net/multimodal_context_net_v2.py:570-585"""
    n = poses.shape[0]
    g = gru(aff_encoder(sd, 'aff_encoder', poses, training), sd, 'gru', 4, H)
    g = g[:, :, :H] + g[:, :, H:]
    l1 = F.linear(g.contiguous().view(-1, H), sd['out.weight'], sd['out.bias']).view(n, -1)
    return torch.sigmoid(F.linear(l1, sd['out2.weight'], sd['out2.bias']))


def conv_discriminator(sd, poses, training, H=64):
    """ConvDiscriminatorTriModal.forward, net/multimodal_context_net_v2.py:415-435 (activations are identity)"""
    x = poses.transpose(1, 2)
    x = _bn(F.conv1d(x, sd['pre_conv.0.weight'], sd['pre_conv.0.bias']), sd, 'pre_conv.1', training)
    x = _bn(F.conv1d(x, sd['pre_conv.3.weight'], sd['pre_conv.3.bias']), sd, 'pre_conv.4', training)
    x = F.conv1d(x, sd['pre_conv.6.weight'], sd['pre_conv.6.bias']).transpose(1, 2)
    g = gru(x, sd, 'gru', 4, H)
    g = g[:, :, :H] + g[:, :, H:]
    l1 = F.linear(g.contiguous().view(-1, H), sd['out.weight'], sd['out.bias']).view(poses.shape[0], -1)
    return torch.sigmoid(F.linear(l1, sd['out2.weight'], sd['out2.bias']))


def attention(x, w1, b1, w2, b2):
    """Attention.forward, net/ser_att_conv_rnn_v2.py:30-34"""
    v = torch.sigmoid(F.linear(x, w1, b1))
    alphas = torch.softmax(F.linear(v, w2, b2), dim=-2)
    return torch.sum(x * alphas, dim=1), alphas


# ------------------------------------------------------------------ the GAN step, processor_v2.py:776-957
def as_leaves(sd):
    """clone a state_dict; float parameters (not running stats) become autograd leaves"""
    out, seen = {}, {}
    for k, v in sd.items():
        key = (v.data_ptr(), tuple(v.shape))
        if key in seen and v.numel() > 0:   # aliases (tcn `net.0.*` == `conv1.*`, net/tcn.py:31) stay aliases
            out[k] = seen[key]
            continue
        v = v.detach().clone()
        if v.dtype.is_floating_point and 'running_' not in k:
            v.requires_grad_(True)
        out[k] = v
        seen[key] = v
    return out


def _adam(sd, state, lr, betas=(0.5, 0.999), eps=1e-8):
    """torch.optim.Adam single step (processor_v2.py:215-220), bias-corrected"""
    state['t'] = state.get('t', 0) + 1
    t = state['t']
    done = set()
    with torch.no_grad():
        for k, p in sd.items():
            if not p.requires_grad or p.grad is None or id(p) in done:
                continue
            done.add(id(p))
            m = state.setdefault('m.' + k, torch.zeros_like(p))
            v = state.setdefault('v.' + k, torch.zeros_like(p))
            m.mul_(betas[0]).add_(p.grad, alpha=1 - betas[0])
            v.mul_(betas[1]).addcmul_(p.grad, p.grad, value=1 - betas[1])
            denom = (v.sqrt() / np.sqrt(1 - betas[1] ** t)).add_(eps)
            p.addcdiv_(m, denom, value=-lr / (1 - betas[0] ** t))


def gan_step(g_sd, d_sd, t_sd, batch, eps_list, rand_idx, cfg, opt_state, train=True, tri_training=True, epoch=1):
    """One iteration of Processor.forward_pass_s2ag with dropout disabled and the random draws
    injected: eps_list = noise for the reparametrisation of [G pass 1 (D step), T pass, G pass 2,
    G pass 3 (shuffled speakers)], rand_idx = the permutation of :903.
    Returns dict(losses..., out_dir_vec, out_trimodal, ret)."""
    in_text, in_audio, in_mfcc, target, vid = batch
    n_pre = cfg['n_pre_poses']
    pre_seq = target.new_zeros(target.shape[0], target.shape[1], target.shape[2] + 1)
    pre_seq[:, :n_pre, :-1] = target[:, :n_pre]
    pre_seq[:, :n_pre, -1] = 1
    res = {}
    gan_on = epoch > cfg['loss_warmup'] and cfg['loss_gan_weight'] > 0
    if gan_on:                                                                    # :791-814
        for p in d_sd.values():
            p.grad = None
        out1 = pose_generator(g_sd, pre_seq, in_text, in_mfcc, vid, eps_list[0], train)[0]
        d_real = aff_discriminator(d_sd, target, train)
        d_fake = aff_discriminator(d_sd, out1.detach(), train)
        dis = torch.sum(-torch.mean(torch.log(d_real + 1e-8) + torch.log(1 - d_fake + 1e-8)))
        res['dis'] = dis.item()
        if train:
            dis.backward()
            res['d_grads'] = {k: v.grad.clone() for k, v in d_sd.items() if v.requires_grad and v.grad is not None}
            _adam(d_sd, opt_state.setdefault('d', {}), cfg['learning_rate'] * cfg['discriminator_lr_weight'])
    for p in g_sd.values():                                                       # :818
        p.grad = None
    with torch.no_grad():
        out_tri = pose_generator_trimodal(t_sd, pre_seq, in_text, in_audio, vid, eps_list[1], tri_training)[0]  # :821
    out, z, mu, lv = pose_generator(g_sd, pre_seq, in_text, in_mfcc, vid, eps_list[2], train)          # :823
    huber = F.smooth_l1_loss(out / 0.1, target / 0.1) * 0.1                       # :894
    d_out = aff_discriminator(d_sd, out, train)
    gen = -torch.mean(torch.log(d_out + 1e-8))                                    # :896
    out_r, z_r, _, _ = pose_generator(g_sd, pre_seq, in_text, in_mfcc, vid[rand_idx], eps_list[3], train)  # :903-909
    pose_l1 = (F.smooth_l1_loss(out / 0.05, out_r.detach() / 0.05, reduction='none') * 0.05).sum(1).sum(1)
    z_l1 = F.l1_loss(z.detach(), z_r.detach(), reduction='none').mean(1)
    div = torch.clamp(-(pose_l1 / (z_l1 + 1.0e-5)), min=-1000).mean()              # :912-922
    kld = -0.5 * torch.mean(1 + lv - mu.pow(2) - lv.exp())                         # :926
    loss = cfg['loss_regression_weight'] * huber + cfg['loss_kld_weight'] * kld + cfg['loss_reg_weight'] * div
    if gan_on:
        loss = loss + cfg['loss_gan_weight'] * gen                                 # :936-937
    if train:
        loss.backward()
        res['g_grads'] = {k: v.grad.clone() for k, v in g_sd.items() if v.requires_grad and v.grad is not None}
        _adam(g_sd, opt_state.setdefault('g', {}), cfg['learning_rate'])
    res.update(huber=huber.item(), gen=gen.item(), kld=kld.item(), div=div.item(), total=loss.item(),
               out_dir_vec=out.detach(), out_trimodal=out_tri,
               ret=F.l1_loss(out, target).item() - F.l1_loss(out_tri, target).item())   # :956
    return res


# ------------------------------------------------------------------ deterministic, portable weights
def fill_state_dict(sd, seed, scale=1.0):
    """Overwrite every floating tensor of `sd` in place with values from numpy's MT19937 (stable
    across platforms), keyed by name order, so fixtures need not store 13 M weights.  Ranges are
    chosen to keep every layer well-conditioned: BN gamma in [0.5,1.5], running_var in [0.5,1.5],
    weight_g in [0.5,1.5], everything else uniform with a fan-in scaled bound."""
    rng = np.random.RandomState(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        if not v.dtype.is_floating_point:
            continue
        shape = tuple(v.shape)
        if k.endswith('running_var') or k.endswith('weight_g') or \
                (k.endswith('.weight') and v.dim() == 1):
            a = rng.uniform(0.5, 1.5, size=shape)
        elif k.endswith('running_mean') or v.dim() == 1:
            a = rng.uniform(-0.2, 0.2, size=shape)
        else:
            fan_in = int(np.prod(shape[1:])) if v.dim() > 1 else shape[0]
            if 'embedding' in k or 'speaker_embedding.0' in k:
                a = rng.normal(0, 1.0, size=shape)
            else:
                b = scale * np.sqrt(3.0 / fan_in)
                a = rng.uniform(-b, b, size=shape)
        with torch.no_grad():
            v.copy_(torch.from_numpy(a.astype(np.float32)).view(shape))
    return sd


def synthetic_batch(B, n_words, n_speakers, audio_length, seed):
    """SURVEY 8d synthetic inputs from numpy's RNG (portable)."""
    rng = np.random.RandomState(seed)
    T, P = 34, 27
    target = np.clip(rng.normal(0, 0.3, size=(B, T, P)), -2, 2).astype(np.float32)
    text = np.zeros((B, T), dtype=np.int64)
    for i in range(B):
        pos = rng.choice(T, size=10, replace=False)
        text[i, pos] = rng.randint(4, n_words, size=10)
    audio = rng.uniform(-0.5, 0.5, size=(B, audio_length)).astype(np.float32)
    mfcc = rng.normal(0, 0.1, size=(B, 37, 71)).astype(np.float32)
    vid = rng.randint(0, n_speakers, size=B).astype(np.int64)
    eps = rng.normal(0, 1, size=(4, B, 16)).astype(np.float32)
    rand_idx = rng.permutation(B).astype(np.int64)
    t = torch.from_numpy
    return (t(text), t(audio), t(mfcc), t(target), t(vid)), [t(e) for e in eps], t(rand_idx)


CFG = dict(n_pre_poses=4, n_poses=34, input_context='both', hidden_size=300, hidden_size_s2eg=300, n_layers=4,
           dropout_prob=0.3, freeze_wordembed=False, wordembed_dim=300, z_type='speaker', learning_rate=5e-4,
           discriminator_lr_weight=0.2, loss_regression_weight=500, loss_gan_weight=5.0, loss_warmup=0,
           loss_kld_weight=0.1, loss_reg_weight=0.05, motion_resampling_framerate=15, num_mfcc=14,
           mean_dir_vec=[0.0154009, -0.9690125, -0.0884354, -0.0022264, -0.8655276, 0.4342174, -0.0035145, -0.8755367, -0.4121039,
                         -0.9236511, 0.3061306, -0.0012415, -0.5155854, 0.8129665, 0.0871897, 0.2348464, 0.1846561, 0.8091402,
                         0.9271948, 0.2960011, -0.013189, 0.5233978, 0.8092403, 0.0725451, -0.2037076, 0.1924306, 0.8196916])
# config/multimodal_context_v2.yml:15-46 + parse_args.py defaults
