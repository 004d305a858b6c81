"""Fused WavEncoder forward (csrc/umma_wav.cu, s2ag_wavencoder_fwd) against torch.nn built exactly like the reference
(net/multimodal_context_net_v2.py:17-28), BatchNorm in train mode (batch statistics + running-statistic update, which is
how the frozen tri-modal baseline runs during training: processor_v2.py:961-962) and in eval mode.  Tolerance 1e-4
relative to the output range (module bar of DESIGN.md; north_star bar is 1e-3)."""
import pytest
import torch
import torch.nn as nn

from common import rel

pytestmark = pytest.mark.gpu


def _reference_stack():
    return nn.Sequential(
        nn.Conv1d(1, 16, 15, stride=5, padding=1600), nn.BatchNorm1d(16), nn.LeakyReLU(0.3, inplace=True),
        nn.Conv1d(16, 32, 15, stride=6), nn.BatchNorm1d(32), nn.LeakyReLU(0.3, inplace=True),
        nn.Conv1d(32, 64, 15, stride=6), nn.BatchNorm1d(64), nn.LeakyReLU(0.3, inplace=True),
        nn.Conv1d(64, 32, 15, stride=6))


def _pair(seed):
    from speech2affective_gestures_b200.net.multimodal_context_net_v2 import WavEncoder
    torch.manual_seed(seed)
    ref = _reference_stack().double()
    with torch.no_grad():
        for m in ref:
            if isinstance(m, nn.BatchNorm1d):   # non-trivial affine parameters and running statistics
                m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.3, 0.3)
                m.running_mean.uniform_(-0.2, 0.2); m.running_var.uniform_(0.5, 2.0)
    we = WavEncoder()
    we.feat_extractor.load_state_dict({k: v.float() for k, v in ref.state_dict().items()}, strict=True)
    return ref, we.cuda()


@pytest.mark.parametrize("B,L", [(3, 36267), (1, 36000), (5, 20011), (2, 61234)])
@pytest.mark.parametrize("train", [True, False])
def test_wavencoder_fused_vs_torch(B, L, train):
    ref, we = _pair(7)
    ref.train(train); we.train(train)
    torch.manual_seed(B * 1000 + L)
    audio = torch.rand(B, L, dtype=torch.float64) - 0.5
    with torch.no_grad():
        want = ref(audio.unsqueeze(1)).transpose(1, 2)      # reference :33 -> (batch, seq, dim)
        got = we(audio.float().cuda())
    assert got.shape == want.shape
    assert rel(got, want) < 1e-4
    rsd, msd = ref.state_dict(), we.feat_extractor.state_dict()
    for k in rsd:
        if "running" in k:
            assert rel(msd[k], rsd[k]) < 1e-5, k
    if train:
        assert int(msd["1.num_batches_tracked"]) == 1 or hasattr(we.feat_extractor[1], "_s2ag_batches")


def test_wavencoder_fused_matches_unfused_chain(monkeypatch):
    """the fused kernels against this repo's own layer-by-layer path (same weights, batch statistics), written straight
    into a column slice of a wider buffer (how PoseGeneratorTriModal consumes it)"""
    ref, we = _pair(11)
    we.train()
    audio = (torch.rand(4, 36267) - 0.5).cuda()
    buf = torch.zeros(4, 34, 88, device="cuda")
    from speech2affective_gestures_b200 import ops
    with torch.no_grad():
        a = we(audio, out=ops.col_slice(buf, 27, 59))
        monkeypatch.setenv("S2AG_WAV_FUSED", "0")
        b = we(audio)
    assert rel(buf[:, :, 27:59], b) < 1e-4
    assert float(buf[:, :, :27].abs().max()) == 0.0 and float(buf[:, :, 59:].abs().max()) == 0.0
    assert a.data_ptr() == buf[:, :, 27:59].data_ptr()


def test_wavencoder_large_batch_property():
    """benchmark size (256 clips): clips are independent given the batch statistics, so permuting the clips permutes
    the features (size-independent property; the full-size reference does not run in seconds)"""
    ref, we = _pair(3)
    we.train()
    audio = (torch.rand(256, 36267, device="cuda") - 0.5)
    perm = torch.randperm(256, device="cuda")
    with torch.no_grad():
        a = we(audio)
        b = we(audio[perm].contiguous())
    assert torch.isfinite(a).all()
    assert rel(b, a[perm]) < 2e-5


def test_wavencoder_fused_bf16x1_mode():
    """BASELINE config 3's single-pass bf16 operand mode (s2ag_set_precision(1)): the hi-plane-only code path of the
    fused kernels; a throughput mode, bounded here at 3e-2 like the contraction engine's bf16x1 test"""
    from speech2affective_gestures_b200 import _C
    ref, we = _pair(5)
    ref.train(); we.train()
    audio = torch.rand(3, 36267, dtype=torch.float64) - 0.5
    with torch.no_grad():
        want = ref(audio.unsqueeze(1)).transpose(1, 2)
        assert _C.lib().s2ag_set_precision(1) == 0
        try:
            got = we(audio.float().cuda())
        finally:
            _C.lib().s2ag_set_precision(0)
    assert 1e-5 < rel(got, want) < 3e-2
