"""Front-end / long-form / metric kernels against the numpy restatement in oracle/frontend_oracle.py.
`dev` runs every test on the CPU kernel-logic emulator here and on the GPU (through the C ABI) on the B200 box."""
import os
import sys

import numpy as np
import pytest
import torch

from common import ROOT
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import frontend_oracle as FO  # noqa: E402  (test infrastructure)

from speech2affective_gestures_b200 import ops  # noqa: E402
from speech2affective_gestures_b200.utils import audio_features as af  # noqa: E402


def _speechlike(rng, L, sr=16000):
    """harmonic stack with a decaying envelope + a little noise: ~60 dB of spectral dynamic range"""
    t = np.arange(L) / sr
    f0 = rng.uniform(90, 220)
    y = sum(np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28)) / h ** 1.5 for h in range(1, 30))
    y = y * (0.3 + 0.7 * np.abs(np.sin(2 * np.pi * 3.1 * t))) + rng.normal(0, 1e-3, L)
    return (0.4 * y / np.abs(y).max()).astype(np.float32)


def test_mel_bank_and_dct_tables_match_the_restatement():
    bank, span = af.mel_bank()
    assert np.array_equal(bank, FO.mel_filterbank())
    for m in range(bank.shape[0]):
        nz = np.nonzero(bank[m])[0]
        assert span[m, 0] == nz[0] and span[m, 1] == nz[-1] + 1
    assert np.abs(af.dct_rows(14) - FO.dct_matrix(14, 128)).max() < 1e-7


@pytest.mark.parametrize("L", [36266, 36267, 5000])
def test_mfcc_features(dev, L):
    rng = np.random.RandomState(3)
    B = 3 if dev.type == "cuda" else 2
    audio = np.stack([rng.uniform(-0.5, 0.5, L).astype(np.float32)] + [_speechlike(rng, L) for _ in range(B - 1)])
    want = np.stack([FO.get_mfcc_features(a, 16000, 14) for a in audio])
    got = ops.mfcc_features(torch.from_numpy(audio).to(dev), 16000, 14).cpu().numpy()
    assert got.shape == want.shape == (B, 37, 1 + L // 512)
    # values are O(0.1) (MFCC / 1000); 1e-3 relative to the tensor maximum (north_star tolerance), typically 1e-5
    assert np.abs(got - want).max() <= 1e-3 * np.abs(want).max(), np.abs(got - want).max()
    assert np.abs(got - want).max() <= 2e-5


def test_expand_inputs_is_bit_exact(dev):
    rng = np.random.RandomState(5)
    B, L = 4, 1001
    a16 = rng.randint(-32768, 32767, size=(B, L)).astype(np.int16)
    h16 = rng.normal(0, 0.1, size=(B, 37, 71)).astype(np.float16)
    h16[0, 0, :4] = [0.0, 6e-8, -6e-5, 65504.0]  # zero, subnormal, small, max
    for dt in (np.float32, np.float64):
        amax = rng.uniform(0.1, 1.0, size=B).astype(dt)
        want_a = torch.from_numpy(a16 * amax[:, None] / 32767).float().numpy()   # processor_v2.py:606-608
        ga, gm = ops.expand_inputs(torch.from_numpy(a16).to(dev), torch.from_numpy(amax).to(dev),
                                   torch.from_numpy(h16).to(dev))
        assert np.array_equal(ga.cpu().numpy(), want_a)
        assert np.array_equal(gm.cpu().numpy(), h16.astype(np.float32))


def test_dir_vec_to_pose_and_metrics(dev):
    rng = np.random.RandomState(9)
    B, T = 5, 34
    out = rng.normal(0, 0.3, size=(B, T, 27)).astype(np.float32)
    tgt = rng.normal(0, 0.3, size=(B, T, 27)).astype(np.float32)
    mean = rng.normal(0, 0.5, size=27).astype(np.float32)
    pose = ops.dir_vec_to_pose(torch.from_numpy(out).to(dev), torch.from_numpy(mean).to(dev)).cpu().numpy()
    assert pose.shape == (B, T, 10, 3)
    assert np.abs(pose - FO.convert_dir_vec_to_pose(out + mean)).max() < 1e-6
    m = ops.pose_metrics(torch.from_numpy(out).to(dev), torch.from_numpy(tgt).to(dev), torch.from_numpy(mean).to(dev),
                         4).cpu().numpy()
    want = FO.push_samples_metrics(out, tgt, mean, 34, 4)
    assert np.allclose(m, want, rtol=1e-5), (m, want)


def test_longform_blend_ragged_and_fade_out(dev):
    rng = np.random.RandomState(11)
    B, T, P, n_pre = 3, 34, 27, 4
    n_chunks = np.array([4, 2, 3], dtype=np.int32)
    C = int(n_chunks.max())
    stride = T - n_pre
    chunks = rng.normal(0, 0.3, size=(C, B, T, P)).astype(np.float32)
    cap = T + stride * (C - 1) + 2 * n_pre   # room for the fade-out padding
    result = torch.zeros(B, cap, P, device=dev)
    pre = torch.zeros(B, T, P + 1, device=dev)
    nc = torch.from_numpy(n_chunks).to(dev)
    for c in range(C):
        ops.longform_blend(torch.from_numpy(chunks[c]).to(dev), result, c, n_pre, pre_next=pre, n_chunks=nc)
        live = np.nonzero(n_chunks > c)[0]
        p = pre.cpu().numpy()
        assert np.array_equal(p[live, :n_pre, :P], chunks[c][live, -n_pre:]) and np.all(p[live, :n_pre, P] == 1)
        assert np.all(p[live, n_pre:] == 0)
    res = result.cpu().numpy().copy()
    lens = []
    for b in range(B):
        want = FO.blend_chunks([chunks[c][b] for c in range(n_chunks[b])], n_pre)
        lens.append(len(want))
        assert np.array_equal(res[b, :len(want)], want), b   # same fp32 operation order as the reference's numpy
        assert np.all(res[b, len(want):] == 0)
    # fade-out: clip 0 ends 1.0 s before its last chunk does, clip 1 exactly at the end (needs padding), clip 2 0.3 s
    pad_samples = np.array([16000, 0, 4800])
    start = np.array([lens[b] - int(pad_samples[b] / 16000 * 15) for b in range(B)], dtype=np.int32)
    new_len = ops.fade_out(result, torch.tensor(lens, dtype=torch.int32, device=dev), torch.from_numpy(start).to(dev),
                           n_pre).cpu().numpy()
    res2 = result.cpu().numpy()
    for b in range(B):
        want, s0, e0 = FO.fade_out(res[b, :lens[b]], pad_samples[b], 16000, 15, n_pre, P)
        assert s0 == start[b] and new_len[b] == len(want)
        assert np.abs(res2[b, :len(want)] - want).max() < 1e-6, b


def test_attention_backward(dev):
    torch.manual_seed(4)
    N, T, Hd, A = 3, 37, 32, 32
    x = torch.randn(N, T, Hd)
    w1, b1, w2, b2 = torch.randn(A, Hd) * 0.3, torch.randn(A) * 0.1, torch.randn(1, A) * 0.5, torch.randn(1) * 0.1
    go, ga = torch.randn(N, Hd), torch.randn(N, T, 1)
    ref = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    v = torch.sigmoid(torch.nn.functional.linear(ref[0], ref[1], ref[2]))   # net/ser_att_conv_rnn_v2.py:30-34
    al = torch.softmax(torch.nn.functional.linear(v, ref[3], ref[4]), dim=-2)
    o = torch.sum(ref[0] * al, dim=1)
    torch.autograd.backward([o, al], [go, ga])
    mine = [t.clone().to(dev).requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    o2, al2 = ops.attention(*mine)
    assert torch.allclose(o2.cpu(), o, atol=1e-5) and torch.allclose(al2.cpu(), al, atol=1e-6)
    torch.autograd.backward([o2, al2], [go.to(dev), ga.to(dev)])
    for a, r in zip(mine, ref):
        assert a.grad is not None
        assert (a.grad.cpu() - r.grad).abs().max() <= 2e-4 * max(1.0, r.grad.abs().max().item()), a.shape
