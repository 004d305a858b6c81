"""The Processor entry points of SURVEY 8(a2)/(a17)/(b): constructor, yield_batch, per_train_epoch / per_val_epoch /
train (checkpoint save, discovery incl. negative losses, resume), load_model_at_epoch with checkpoints WRITTEN BY THE
REFERENCE CLASSES (strict=True), generate_gestures (device-side push_samples metrics), the prefetch path bench.py's
`e2e` is measured through, and CUDA-graph capture hygiene."""
import json
import os
import shutil
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from common import O, ROOT, cfg_dict, derand, inject_eps, rel
import frontend_oracle as FO  # noqa: E402
from speech2affective_gestures_b200.net import multimodal_context_net_v2 as M
from speech2affective_gestures_b200.net import embedding_net as men
from speech2affective_gestures_b200.processor_v2 import Processor, get_epoch_and_loss
from speech2affective_gestures_b200.synthetic import make_data_loader, Vocab

CKPT = os.path.join(ROOT, "tests", "golden", "ref_ckpt_tiny")
SCHEMA = os.path.join(ROOT, "tests", "golden", "ref_state_dict_schema.json")


def make_processor(dev, tmp, kind="tiny", n_words=40, n_spk_rows=25, n_train=8, batch_size=4, n_val=8, **kw):
    c = cfg_dict(kind)
    args = NS(no_cuda=(dev.type != "cuda"), work_dir_s2ag=os.path.join(tmp, "work"), save_log=True, print_log=False,
              train_s2ag=True, batch_size=batch_size, s2ag_num_epoch=2, val_interval=1, save_interval=1,
              s2ag_load_last_best=False, s2ag_start_epoch='best', video_save_path=os.path.join(tmp, "videos"))
    dl = make_data_loader(n_train, n_val, 8, n_words=n_words, n_speakers=n_spk_rows)
    return Processor(tmp, args, NS(**c), dl, 27, 3, 16000, **kw), c


def test_get_epoch_and_loss_accepts_negative_losses(tmp_path):
    d = str(tmp_path)
    for name in ("epoch_000003_loss_0.0312_model.pth.tar", "epoch_000022_loss_-0.0123_model.pth.tar",
                 "epoch_000030_loss_0.0001_model.pth.tar", "log.txt"):
        open(os.path.join(d, name), "w").close()
    assert get_epoch_and_loss(d, 'best') == ("epoch_000022_loss_-0.0123_model.pth.tar", 22, -0.0123)
    assert get_epoch_and_loss(d, 22) == ("epoch_000022_loss_-0.0123_model.pth.tar", 22, -0.0123)
    assert get_epoch_and_loss(d, 3)[1:] == (3, 0.0312)
    assert get_epoch_and_loss(d, 4) == (None, None, np.inf)
    assert get_epoch_and_loss(None, 'best') == (None, None, np.inf)


def test_state_dict_schema_equals_reference():
    """keys, order, shapes and dtypes of the four networks in the shipped configuration == the reference classes'
    (fixture written by oracle/gen_golden.py from the reference's own state_dict())"""
    ref = json.load(open(SCHEMA))
    cfg = NS(**O.CFG)
    spk = Vocab("vid", ref["n_speakers"])
    nets = dict(gen=M.PoseGenerator(cfg, 27, ref["n_words"], 300, None, 71, 37, 34, z_obj=spk),
                tri=M.PoseGeneratorTriModal(cfg, 27, ref["n_words"], 300, None, z_obj=spk),
                dis=M.AffDiscriminator(27), cdis=M.ConvDiscriminatorTriModal(27))
    for k, net in nets.items():
        mine = [[n, list(v.shape), str(v.dtype)] for n, v in net.state_dict().items()]
        assert mine == ref["schema"][k], k


def test_load_reference_checkpoint_strict_and_reproduce_outputs(dev, tmp_path):
    """load_model_at_epoch('best') / load_trimodal on files torch.save'd by the REFERENCE classes (negative loss in the
    file name), strict key matching, then eval-mode outputs equal the reference's"""
    tmp = str(tmp_path)
    shutil.copytree(os.path.join(CKPT, "work"), os.path.join(tmp, "work"))
    shutil.copytree(os.path.join(CKPT, "outputs"), os.path.join(tmp, "outputs"))
    open(os.path.join(tmp, "work", "log.txt"), "w").close()
    pr, c = make_processor(dev, tmp)
    assert pr.load_model_at_epoch('best') and pr.best_s2ag_loss_epoch == 22 and pr.best_s2ag_loss == -0.0123
    assert pr.load_model_at_epoch(22)
    assert not pr.load_model_at_epoch(5)
    assert pr.load_trimodal()
    loaded = torch.load(os.path.join(tmp, "work", "epoch_000022_loss_-0.0123_model.pth.tar"), map_location="cpu")
    res = pr.s2ag_generator.load_state_dict(loaded['gen_model_dict'], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # parameters still alias the flat buffer after loading
    p0 = next(pr.s2ag_generator.parameters())
    assert p0.data_ptr() == pr.s2ag_generator.flat_params.data_ptr()
    want = np.load(os.path.join(CKPT, "expected.npz"))
    batch, eps_list, _ = O.synthetic_batch(3, 40, 24, 36267, seed=99)
    text, audio, mfcc, target, vid = (t.to(dev) for t in batch)
    pre = pr.make_pre_seq(target)
    for n in (pr.s2ag_generator, pr.trimodal_generator, pr.s2ag_discriminator):
        n.eval()
    inject_eps([eps_list[0]])
    with torch.no_grad():
        assert rel(pr.s2ag_generator(pre, text, mfcc, vid)[0], torch.from_numpy(want["g_out"])) < 1e-3
        assert rel(pr.trimodal_generator(pre, text, audio, vid)[0], torch.from_numpy(want["t_out"])) < 1e-3
        assert rel(pr.s2ag_discriminator(target), torch.from_numpy(want["d_out"])) < 1e-3


def test_yield_batch_contract(dev, tmp_path):
    pr, c = make_processor(dev, str(tmp_path), n_train=10, batch_size=4)
    np.random.seed(0)
    batches = list(pr.yield_batch(train=True))
    assert len(batches) == 3   # ceil(10 / 4) pseudo passes (processor_v2.py:597)
    text, vec, audio, mfcc, vids = batches[0]
    assert text.dtype == torch.int64 and tuple(text.shape) == (4, 34)
    assert vec.dtype == torch.float32 and tuple(vec.shape) == (4, 34, 27)
    assert audio.dtype == torch.float32 and tuple(audio.shape) == (4, 36267) and audio.abs().max() <= 0.5 + 1e-6
    assert mfcc.dtype == torch.float32 and tuple(mfcc.shape) == (4, 37, 71)
    assert vids.dtype == torch.int64 and tuple(vids.shape) == (4,)
    # the device-side expansion of the int16 / fp16 cache rows is bit-identical to the reference's host arithmetic
    np.random.seed(0)
    keys = np.random.choice(10, size=4, replace=True)
    s = pr.train_samples
    assert np.array_equal(audio.cpu().numpy(),
                          torch.from_numpy(s['audio'][keys] * s['audio_max'][keys, None] / 32767).float().numpy())
    assert np.array_equal(mfcc.cpu().numpy(), s['mfcc_features'][keys].astype(np.float32))
    # speakers are drawn from those NOT in the batch (:625-630)
    assert not set(vids.tolist()) & set(s['vid_indices'][keys].tolist())


def test_train_checkpoint_resume_generate(dev, tmp_path):
    """train() for two epochs (GAN branch from epoch 1), checkpoints written in the reference schema, resume, and the
    batched evaluation entry point with device-side metrics"""
    tmp = str(tmp_path)
    torch.manual_seed(3)
    np.random.seed(3)
    small = dict(n_train=4, n_val=4) if dev.type != "cuda" else {}   # the CPU emulator runs ~1 iteration / 10 s
    pr, c = make_processor(dev, tmp, min_train_epochs=-1, **small)
    w0 = pr.s2ag_generator.flat_params.clone()
    pr.train()
    assert not torch.equal(w0, pr.s2ag_generator.flat_params)
    assert np.isfinite(pr.epoch_info['mean_s2ag_loss'])
    files = sorted(f for f in os.listdir(pr.args.work_dir_s2ag) if f.endswith(".pth.tar"))
    assert files and all(f.startswith("epoch_00000") for f in files)
    ck = torch.load(os.path.join(pr.args.work_dir_s2ag, files[-1]), map_location="cpu")
    assert set(ck.keys()) == {'gen_model_dict', 'dis_model_dict'}
    bn = [k for k in ck['gen_model_dict'] if k.endswith("aff_encoder.batch_norm1.num_batches_tracked")]
    assert int(ck['gen_model_dict'][bn[0]]) > 0   # host-side BN batch counters are folded into the saved buffers
    # resume in a fresh Processor: strict load, continues at the discovered epoch
    pr2, _ = make_processor(dev, tmp, min_train_epochs=-1, **small)
    name, e_best, _ = get_epoch_and_loss(pr2.args.work_dir_s2ag, 'best')
    pr2.args.s2ag_load_last_best, pr2.args.s2ag_start_epoch, pr2.args.s2ag_num_epoch = True, 'best', e_best + 1
    seen = []
    orig = pr2.per_train_epoch
    pr2.per_train_epoch = lambda: (seen.append(pr2.meta_info['epoch']), orig())[1]
    pr2.train()
    assert seen == [e_best]
    # batched evaluation (generate_gestures): metrics equal push_samples on the returned tensors
    np.random.seed(5)
    r = pr2.generate_gestures(samples_to_generate=6, randomized=False, load_saved_model=True, s2ag_epoch='best')
    assert r['clips'] == 6 and all(np.isfinite(r[k]) for k in ('loss', 'joint_mae', 'accel', 'loss_trimodal'))
    test = pr2.data_loader['test_data_s2ag'].samples
    want = FO.push_samples_metrics(pr2.last_out.cpu().numpy(), test['vec_seq'][:6], c["mean_dir_vec"], 34, 4)
    assert np.allclose([r['loss'], r['joint_mae'], r['accel']], want, rtol=1e-4)


def test_speaker_id_outside_table_is_reported(dev, tmp_path):
    pr, c = make_processor(dev, str(tmp_path))
    with pytest.raises(IndexError):
        pr._check_speaker_ids(torch.tensor([3, 99], device=dev))


@pytest.mark.gpu
def test_capture_preserves_model_and_uses_packed_route(tmp_path):
    """capture_step() must leave the live model untouched (weights, Adam state, BN statistics, counters) and every
    stream of the captured step must own a scratch buffer, so that weight-side contractions take the packed-operand
    (TMA-fed) kernel inside the graph as well; then prefetch_inputs / swap_in_prefetched / replay_step (the path e2e is
    measured through), fp32 and compressed, give the same losses as an eager step on the same batch."""
    dev = torch.device("cuda:0")
    from speech2affective_gestures_b200 import ops
    from speech2affective_gestures_b200.synthetic import synthetic_batch
    pr, c = make_processor(dev, str(tmp_path), kind="full", n_words=500, n_spk_rows=50, batch_size=128)
    for net in (pr.s2ag_generator, pr.trimodal_generator, pr.s2ag_discriminator):
        derand(net)   # dropout off: graph replay and eager step are then comparable number by number
        net.train()
    pr.meta_info['epoch'] = 1
    snap = [n.flat_params.clone() for n in (pr.s2ag_generator, pr.s2ag_discriminator, pr.trimodal_generator)]
    bn = pr.s2ag_generator.aff_encoder.batch_norm1
    rm, nbt = bn.running_mean.clone(), int(pr.s2ag_generator.state_dict()['aff_encoder.batch_norm1.num_batches_tracked'])
    B = 128
    pr.capture_step(B, train=True, warmup=2)
    for a, n in zip(snap, (pr.s2ag_generator, pr.s2ag_discriminator, pr.trimodal_generator)):
        assert torch.equal(a, n.flat_params)
    assert torch.equal(rm, bn.running_mean) and int(pr.gen_step) == 0 and int(pr.dis_step) == 0
    assert float(pr.gen_m.abs().max()) == 0.0
    assert int(pr.s2ag_generator.state_dict()['aff_encoder.batch_norm1.num_batches_tracked']) == nbt
    for st in (pr._capture_stream, pr._side_stream, pr._side_stream_b):
        assert ops.has_scratch(st, dev)
    host = synthetic_batch(B, None, 500, 50, 36267, seed=7, pin=True)
    eps = [torch.randn(B, 16, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(4)]
    ridx = torch.randperm(B, generator=torch.Generator().manual_seed(9)).to(dev)
    # eager step on a twin processor with the same weights
    pr_e, _ = make_processor(dev, str(tmp_path), kind="full", n_words=500, n_spk_rows=50, batch_size=128)
    for a, b in zip((pr.s2ag_generator, pr.s2ag_discriminator, pr.trimodal_generator),
                    (pr_e.s2ag_generator, pr_e.s2ag_discriminator, pr_e.trimodal_generator)):
        derand(b)
        b.train()
        b.load_state_dict(a.state_dict())
    pr_e.meta_info['epoch'] = 1
    pr_e.injected_rand_idx = ridx
    inject_eps(eps)
    m_e = pr_e.gan_step_async(*[t.to(dev) for t in host], True).clone()
    # (the captured graph drew its noise with torch.randn at capture time: re-capture with injected draws)
    pr.injected_rand_idx = ridx
    men.eps_source = None
    st = {"i": 0}

    def src(like):   # static tensors so that the graph re-reads them on replay
        e = eps[st["i"] % 4]
        st["i"] += 1
        return e
    men.eps_source = src
    pr.capture_step(B, train=True, warmup=1)
    pr.prefetch_inputs(*host)
    pr.swap_in_prefetched()
    m_g = pr.replay_step().clone()
    torch.cuda.synchronize()
    assert torch.allclose(m_g, m_e, rtol=2e-3, atol=1e-6), (m_g, m_e)
    assert int(pr.gen_step) == 1
    # compressed wire format: int16 audio + scale, fp16 MFCC
    amax = host[1].abs().amax(dim=1)
    a16 = torch.round(host[1] / amax[:, None] * 32767).to(torch.int16).pin_memory()
    m16 = host[2].to(torch.float16).pin_memory()
    pr.prefetch_inputs_compressed(host[0], a16, amax.pin_memory(), m16, host[3], host[4])
    pr.swap_in_prefetched()
    torch.cuda.synchronize()
    assert torch.equal(pr.static_in[1].cpu(), (a16.float() * amax[:, None] / 32767))
    assert torch.equal(pr.static_in[2].cpu(), m16.float())
    men.eps_source = None
