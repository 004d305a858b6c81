"""Long-form entry points (`Processor.render_clip`, `generate_gestures_by_dataset`, `synthesize_long_form`,
BASELINE config 5) against
  * the fixture recorded from the reference's UNMODIFIED `render_clip` (oracle/gen_golden.py longform; full-width
    networks: GPU), and
  * the oracle chunk loop (oracle/s2ag_oracle.py generators + oracle/frontend_oracle.py MFCC / blend / fade-out /
    joint conversion) on small-width networks (CPU kernel-logic emulator and GPU).
Tolerance 1e-3 relative on the generated joint positions (north_star)."""
import copy
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from common import O, ROOT, cfg_dict, derand, sd_cpu, inject_eps, rel
import frontend_oracle as FO  # noqa: E402  (oracle/ is on sys.path through common)
from speech2affective_gestures_b200 import longform
from speech2affective_gestures_b200.processor_v2 import Processor
from speech2affective_gestures_b200.synthetic import make_data_loader, Vocab

LF_GOLDEN = os.path.join(ROOT, "tests", "golden", "s2ag_longform_golden.npz")


def lang_model():
    """word -> index exactly as the fixture's language model: 'w4'..'w55' known, everything else UNK (3)"""
    lm = Vocab("words", 64)
    lm.word2index = {"w%d" % i: i for i in range(4, 56)}
    return lm


def make_processor(kind, dev):
    c = cfg_dict(kind)
    args = NS(no_cuda=(dev.type != "cuda"), work_dir_s2ag=None, save_log=False, print_log=False, train_s2ag=True,
              batch_size=4, s2ag_num_epoch=1, val_interval=1, save_interval=10, video_save_path=None)
    dl = make_data_loader(4, 4, 4, n_words=64, n_speakers=25, lang_model=lang_model())
    pr = Processor("/nonexistent", args, NS(**c), dl, 27, 3, 16000)
    for net, seed in ((pr.s2ag_generator, 100), (pr.trimodal_generator, 101)):
        derand(net)
        O.fill_state_dict(net.state_dict(), seed)
        net.eval()
    return pr, c


def clip_args(seed, dur):
    clip = FO.synthetic_clip(seed, duration=dur)
    clip[6]['end_time'] += 1.0   # as oracle/gen_golden.py longform()
    return clip


def test_prepare_clip_matches_reference_schedule():
    """host-side preparation (resampling, chunk schedule, word placement, audio slicing) vs what the reference's
    render_clip fed its generator, chunk by chunk"""
    gold = np.load(LF_GOLDEN)
    cfg = NS(**O.CFG)
    for tag, seed, dur in (("a", 1, 20.0), ("b", 2, 9.3)):
        clip = clip_args(seed, dur)
        pc = longform.prepare_clip(cfg, lang_model(), 27, clip[6]['vid'], clip[1], clip[3], 16000,
                                   copy.deepcopy(clip[0]), [clip[6]['start_time'], clip[6]['end_time']],
                                   speaker_vid_idx=3)
        assert np.allclose(pc.clip_poses_resampled, gold[tag + "_fade0_resampled"], atol=1e-6)
        assert np.array_equal(pc.text_chunks, gold[tag + "_text"])
        assert (pc.text_chunks == 3).any() or tag == "b"   # an out-of-vocabulary word was mapped to UNK
        assert pc.audio_chunks.shape == (len(gold[tag + "_text"]), 36266)
        # the injected MFCC restatement on our audio slices reproduces the reference's per-chunk generator input
        for c in (0, pc.n_chunks - 1):
            m = FO.get_mfcc_features(pc.audio_chunks[c], 16000, 14)
            assert np.abs(m - gold[tag + "_mfcc"][c]).max() < 1e-6
        assert pc.speaker_vid_idx == 3


@pytest.mark.gpu
@pytest.mark.parametrize("fade", [False, True])
def test_render_clip_vs_reference_fixture(fade):
    dev = torch.device("cuda:0")
    gold = np.load(LF_GOLDEN)
    pr, c = make_processor("full", dev)
    eps = torch.from_numpy(gold["eps"])
    for tag, seed, dur in (("a", 1, 20.0), ("b", 2, 9.3)):
        clip = clip_args(seed, dur)
        name = '{}_{:.2f}_{:.2f}'.format(clip[6]['vid'], clip[6]['start_time'], clip[6]['end_time'])
        inject_eps([eps])
        res = pr.render_clip({'audio_sr': 16000, 'clip_duration_range': [5, 12]}, clip[6]['vid'], 0, 1, clip[1], clip[3],
                             16000, copy.deepcopy(clip[0]), [clip[6]['start_time'], clip[6]['end_time']],
                             test_samples=[name], speaker_vid_idx=3, check_duration=False, fade_out=fade)
        k = "%s_fade%d" % (tag, int(fade))
        assert np.allclose(res[0], gold[k + "_resampled"], atol=1e-6)
        for mine, want in ((res[1], gold[k + "_poses_tri"]), (res[2], gold[k + "_poses"])):
            assert mine.shape == want.shape, (k, mine.shape, want.shape)
            assert np.abs(mine - want).max() <= 1e-3 * np.abs(want).max(), (k, np.abs(mine - want).max())
    # filters of the reference signature
    assert pr.render_clip({'clip_duration_range': [5, 12]}, 'x', 0, 1, clip[1], clip[3], 16000, [], [0.0, 30.0],
                          check_duration=True) == (None, None, None)
    assert pr.render_clip({}, 'x', 0, 1, clip[1], clip[3], 16000, [], [0.0, 9.0], test_samples=['other']) == ([], [], [], [])


def oracle_render(pr, c, pc, eps, fade):
    """the reference's chunk loop restated with the oracle pieces, for one prepared clip"""
    g_sd, t_sd = sd_cpu(pr.s2ag_generator), sd_cpu(pr.trimodal_generator)
    mean = np.squeeze(np.array(c["mean_dir_vec"]))
    vid = torch.tensor([pc.speaker_vid_idx])
    outs = {"g": [], "t": []}
    pre = {}
    for k in outs:
        p = torch.zeros(1, 34, 28)
        p[0, :4, :-1] = torch.from_numpy(pc.seed_seq[:4].astype(np.float32))
        p[0, :4, -1] = 1
        pre[k] = p
    for ch in range(pc.n_chunks):
        audio = torch.from_numpy(pc.audio_chunks[ch])[None]
        mfcc = torch.from_numpy(FO.get_mfcc_features(pc.audio_chunks[ch], 16000, 14).astype(np.float32))[None]
        text = torch.from_numpy(pc.text_chunks[ch])[None]
        with torch.no_grad():
            og = O.pose_generator(g_sd, pre["g"], text, mfcc, vid, eps, False, H=c["hidden_size_s2eg"],
                                  n_layers=c["n_layers"])[0]
            ot = O.pose_generator_trimodal(t_sd, pre["t"], text, audio, vid, eps, False, H=c["hidden_size"],
                                           n_layers=c["n_layers"])[0]
        for k, o in (("g", og), ("t", ot)):
            outs[k].append(o[0].numpy())
            pre[k] = torch.zeros(1, 34, 28)
            pre[k][0, :4, :-1] = o[0, -4:]
            pre[k][0, :4, -1] = 1
    res = {}
    for k in outs:
        v = FO.blend_chunks(outs[k], 4)
        if fade:
            v, _, _ = FO.fade_out(v, pc.end_padding, 16000, 15, 4, 27)
        res[k] = FO.convert_dir_vec_to_pose(np.asarray(v, dtype=np.float32) + mean)
    return res["t"], res["g"]


def test_generate_gestures_by_dataset_lockstep_vs_oracle(dev):
    """three clips of different lengths (ragged chunk counts) rendered in ONE lock-step batch through the reference-
    signature entry point, each compared with the oracle's one-clip-at-a-time chunk loop; also exercises the merge
    of consecutive records of one video (processor_v2.py:1497-1523)"""
    pr, c = make_processor("tiny", dev)
    eps = torch.from_numpy(np.random.RandomState(3).normal(0, 1, size=(1, 16)).astype(np.float32))
    durs = (6.5, 9.3) if dev.type != "cuda" else (6.5, 9.3, 11.0)
    clips = [clip_args(10 + i, d) for i, d in enumerate(durs)]
    # split the first clip into two consecutive records of the same video: must be merged back
    w, poses, _, audio, _, _, meta = clips[0]
    cut_t, cut_f = 3.0, 75   # seconds into the clip / pose frames (25 fps)
    rec_a = [[x for x in w if x[1] < meta['start_time'] + cut_t], poses[:cut_f], None, audio[:int(cut_t * 16000)], None,
             None, dict(meta, end_frame_no=meta['start_frame_no'] + cut_f, end_time=meta['start_time'] + cut_t)]
    rec_b = [[x for x in w if x[1] >= meta['start_time'] + cut_t], poses[cut_f:], None, audio[int(cut_t * 16000):], None,
             None, dict(meta, start_frame_no=meta['start_frame_no'] + cut_f, start_time=meta['start_time'] + cut_t)]
    records = [rec_a, rec_b] + clips[1:]
    from speech2affective_gestures_b200.net import embedding_net as men
    for fade in (False, True):
        # every clip of the lock-step batch draws the noise the one-clip-at-a-time oracle loop uses
        men.eps_source = lambda like: eps.expand(like.shape[0], 16).contiguous().to(like.device)
        got = pr.generate_gestures_by_dataset('ted_db', {'clips': copy.deepcopy(records), 'audio_sr': 16000},
                                              check_duration=False, randomized=False, fade_out=fade,
                                              load_saved_model=False)
        assert [g[0] for g in got] == [cl[6]['vid'] for cl in clips]
        for (name, resampled, poses_tri, poses_g), cl in zip(got, clips):
            pc = longform.prepare_clip(NS(**c), lang_model(), 27, cl[6]['vid'], cl[1], cl[3], 16000,
                                       copy.deepcopy(cl[0]), [cl[6]['start_time'], cl[6]['end_time']], speaker_vid_idx=0)
            want_t, want_g = oracle_render(pr, c, pc, eps, fade)
            for mine, want in ((poses_tri, want_t), (poses_g, want_g)):
                assert mine.shape == want.shape, (name, mine.shape, want.shape)
                assert np.abs(mine - want).max() <= 1e-3 * np.abs(want).max(), (name, fade, np.abs(mine - want).max())


def test_synthesize_long_form_config5_shape(dev):
    """BASELINE config 5 (tensor-level API over precomputed per-chunk MFCC): lock-step batch vs the oracle loop"""
    pr, c = make_processor("tiny", dev)
    rng = np.random.RandomState(8)
    B, C = 3, 3
    text = torch.from_numpy(rng.randint(0, 64, size=(B, C, 34))).long()
    mfcc = torch.from_numpy(rng.normal(0, 0.1, size=(B, C, 37, 71)).astype(np.float32))
    vids = torch.from_numpy(rng.randint(0, 25, size=B)).long()
    seed = torch.from_numpy(rng.normal(0, 0.3, size=(B, 4, 27)).astype(np.float32))
    eps = torch.from_numpy(rng.normal(0, 1, size=(B, 16)).astype(np.float32))
    inject_eps([eps])
    got = pr.synthesize_long_form(text.to(dev), mfcc.to(dev), None, vids.to(dev), seed_poses=seed.to(dev)).cpu().numpy()
    assert got.shape == (B, 34 + 30 * (C - 1), 27)
    g_sd = sd_cpu(pr.s2ag_generator)
    for b in range(B):
        pre = torch.zeros(1, 34, 28)
        pre[0, :4, :-1] = seed[b]
        pre[0, :4, -1] = 1
        outs = []
        for ch in range(C):
            with torch.no_grad():
                o = O.pose_generator(g_sd, pre, text[b:b + 1, ch], mfcc[b:b + 1, ch], vids[b:b + 1], eps[b:b + 1], False,
                                     H=c["hidden_size_s2eg"], n_layers=c["n_layers"])[0]
            outs.append(o[0].numpy())
            pre = torch.zeros(1, 34, 28)
            pre[0, :4, :-1] = o[0, -4:]
            pre[0, :4, -1] = 1
        want = FO.blend_chunks(outs, 4)
        assert np.abs(got[b] - want).max() <= 1e-3 * np.abs(want).max()
