"""TEST INFRASTRUCTURE (never imported by the product path).  CPU restatement of the reference's Frechet-gesture-distance
evaluator: the eval-mode pose auto-encoder as torch.nn.functional calls over a reference-schema state_dict, and get_scores
/ calculate_frechet_distance in numpy + scipy exactly as the reference computes them.

Pinned against the UNMODIFIED reference classes by oracle/gen_golden.py (`fgd`): tests/golden/s2ag_fgd_golden.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy import linalg


def synthetic_pairs(seed, n_batches, nb, T=34, P=27):
    """seeded (real, generated) direction-vector clips (numpy MT19937: identical on every platform)"""
    rng = np.random.RandomState(seed)
    real = [np.clip(rng.normal(0, 0.3, size=(nb, T, P)), -2, 2).astype(np.float32) for _ in range(n_batches)]
    gen = [(r + rng.normal(0.05, 0.2, size=r.shape)).astype(np.float32) for r in real]
    return real, gen


def _bn(sd, pre, x):
    return F.batch_norm(x, sd[pre + 'running_mean'], sd[pre + 'running_var'], sd[pre + 'weight'], sd[pre + 'bias'],
                        False, 0.1, 1e-5)


def pose_encoder(sd, poses, prefix='pose_encoder.'):
    """net/embedding_net.py:39-82 (PoseEncoderConv.forward, variational_encoding=False) -> mu [B, 32]"""
    g = lambda k: sd[prefix + k]
    x = poses.transpose(1, 2)
    for i, stride in ((0, 1), (1, 1), (2, 2)):   # conv_norm_relu (:16-36): k3 s1, k3 s1, k4 s2 (down_sample)
        x = F.conv1d(x, g('net.%d.0.weight' % i), g('net.%d.0.bias' % i), stride=stride)
        x = F.leaky_relu(_bn(sd, prefix + 'net.%d.1.' % i, x), 0.2)
    x = F.conv1d(x, g('net.3.weight'), g('net.3.bias')).flatten(1)
    x = F.linear(x, g('out_net.0.weight'), g('out_net.0.bias'))
    x = F.leaky_relu(_bn(sd, prefix + 'out_net.1.', x), 1.0)   # nn.LeakyReLU(True): slope 1.0 (:58)
    x = F.linear(x, g('out_net.3.weight'), g('out_net.3.bias'))
    x = F.leaky_relu(_bn(sd, prefix + 'out_net.4.', x), 1.0)
    x = F.linear(x, g('out_net.6.weight'), g('out_net.6.bias'))
    return F.linear(x, g('fc_mu.weight'), g('fc_mu.bias'))


def pose_decoder(sd, feat, prefix='decoder.'):
    """net/embedding_net.py:164-216 (PoseDecoderConv, length 34, use_pre_poses=False) -> [B, 34, dim]"""
    g = lambda k: sd[prefix + k]
    x = F.linear(feat, g('pre_net.0.weight'), g('pre_net.0.bias'))
    x = F.leaky_relu(_bn(sd, prefix + 'pre_net.1.', x), 1.0)
    x = F.linear(x, g('pre_net.3.weight'), g('pre_net.3.bias')).view(feat.shape[0], 4, -1)
    for i in (0, 3):
        x = F.conv_transpose1d(x, g('net.%d.weight' % i), g('net.%d.bias' % i))
        x = F.leaky_relu(_bn(sd, prefix + 'net.%d.' % (i + 1), x), 0.2)
    x = F.conv1d(x, g('net.6.weight'), g('net.6.bias'))
    x = F.conv1d(x, g('net.7.weight'), g('net.7.bias'))
    return x.transpose(1, 2)


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """net/embedding_space_evaluator.py:104-152"""
    diff = mu1 - mu2
    cov_mean = linalg.sqrtm(sigma1.dot(sigma2))   # the reference passes disp=False (scipy 1.x: returns (sqrtm, errest))
    if not np.isfinite(cov_mean).all():
        off = np.eye(sigma1.shape[0]) * eps
        cov_mean = linalg.sqrtm((sigma1 + off).dot(sigma2 + off))
    if np.iscomplexobj(cov_mean):
        if not np.allclose(np.diagonal(cov_mean).imag, 0, atol=1e-3):
            raise ValueError('Imaginary component')
        cov_mean = cov_mean.real
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(cov_mean)


def get_scores(generated_feats, real_feats):
    """net/embedding_space_evaluator.py:73-101 -> (frechet_dist, feat_dist)"""
    A_mu, A_sigma = np.mean(generated_feats, axis=0), np.cov(generated_feats, rowvar=False)
    B_mu, B_sigma = np.mean(real_feats, axis=0), np.cov(real_feats, rowvar=False)
    try:
        fd = frechet_distance(A_mu, A_sigma, B_mu, B_sigma)
    except ValueError:
        fd = 1e+10
    feat_dist = np.mean([np.sum(np.absolute(real_feats[i] - generated_feats[i])) for i in range(real_feats.shape[0])])
    return float(fd), float(feat_dist)
