"""Long-form (whole-clip) gesture synthesis: the B200-side implementation behind `Processor.render_clip` and
`Processor.generate_gestures_by_dataset` (reference processor_v2.py:1144-1439, :1441-1567).

The reference renders ONE clip at a time at batch 1: per 34-frame chunk it slices the audio, runs librosa's MFCC on the
host, places the word indices on frames, runs the frozen tri-modal baseline and the generator, copies the result to the
host and blends the 4 overlapping frames in numpy.  The only sequential dependency is the 4 seed frames between
consecutive chunks of the same clip, so here MANY clips advance in lock-step: chunk c of every clip is one batch, the
MFCC front-end (csrc/frontend.cu), the seed hand-off + blend, the optional fade-out and the direction-vector -> joint
conversion (csrc/longform.cu) all run on the device, and nothing is copied back before the clips are finished.

Host-side preparation (clip-level metadata: resampling of the ground-truth poses, chunk schedule, word placement) follows
the reference line by line -- it is O(words) Python per clip, not on the hot path.
"""
import math

import numpy as np
import torch

from . import ops
from .utils import ted_db_utils as ted_db


def get_words_in_time_range(word_list, start_time, end_time):
    """utils/data_preprocessor.py:188-202"""
    words = []
    for word in word_list:
        if word[1] >= end_time:
            break
        if word[2] <= start_time:
            continue
        words.append(word)
    return words


class PreparedClip:
    """Everything `render_clip` derives on the host before it touches the networks (processor_v2.py:1152-1271)."""
    __slots__ = ("vid_name", "clip_time", "clip_poses_resampled", "target_dir_vec", "seed_seq", "n_chunks",
                 "audio_chunks", "text_chunks", "end_padding", "speaker_vid_idx", "clip_audio", "clip_words", "clip_idx")


def prepare_clip(cfg, lang_model, pose_dim, vid_name, clip_poses, clip_audio, sample_rate, clip_words, clip_time,
                 unit_time=None, speaker_vid_idx=0, n_speakers=None, clip_idx=0):
    n_frames, n_pre, fps = cfg.n_poses, cfg.n_pre_poses, cfg.motion_resampling_framerate
    mean_dir_vec = np.squeeze(np.array(cfg.mean_dir_vec))
    pc = PreparedClip()
    pc.vid_name, pc.clip_time, pc.clip_idx = vid_name, list(clip_time), clip_idx
    pc.clip_audio = clip_audio
    pc.clip_poses_resampled = ted_db.resample_pose_seq(clip_poses, clip_time[1] - clip_time[0], fps)
    tdv = ted_db.convert_pose_seq_to_dir_vec(pc.clip_poses_resampled)
    tdv = tdv.reshape(tdv.shape[0], -1)
    tdv -= mean_dir_vec
    pc.target_dir_vec = tdv
    pc.seed_seq = tdv[0:n_pre]
    words = [[w[0], w[1] - clip_time[0], w[2] - clip_time[0]] for w in clip_words]  # start of the input text at zero
    pc.clip_words = words
    clip_length = len(clip_audio) / sample_rate
    if unit_time is None:
        unit_time = n_frames / fps
    stride_time = (n_frames - n_pre) / fps
    num_subdivisions = 1 if clip_length < unit_time else math.ceil((clip_length - unit_time) / stride_time) + 1
    audio_sample_length = int(unit_time * sample_rate)
    audio_chunks, text_chunks = [], []
    pc.end_padding = 0
    for sub in range(num_subdivisions):
        t0 = min(sub * stride_time, clip_length)
        t1 = min(t0 + unit_time, clip_length)
        if t0 >= t1:
            continue
        a0 = math.floor(t0 / clip_length * len(clip_audio))
        a = np.asarray(clip_audio[a0:a0 + audio_sample_length], dtype=np.float32)
        if len(a) < audio_sample_length:
            if sub == num_subdivisions - 1:
                pc.end_padding = audio_sample_length - len(a)
            a = np.pad(a, (0, audio_sample_length - len(a)), 'constant')
        audio_chunks.append(a)
        ext = np.zeros(n_frames, dtype=np.int64)  # zero is the index of the padding token
        frame_duration = (t1 - t0) / n_frames
        for word in get_words_in_time_range(words, t0, t1):
            idx = max(0, int(np.floor((word[1] - t0) / frame_duration)))
            ext[idx] = lang_model.get_word_index(word[0])
        text_chunks.append(ext)
    pc.n_chunks = len(audio_chunks)
    pc.audio_chunks = np.stack(audio_chunks)
    pc.text_chunks = np.stack(text_chunks)
    if cfg.z_type == 'speaker':
        if speaker_vid_idx is None:
            speaker_vid_idx = np.random.randint(0, n_speakers)
        pc.speaker_vid_idx = int(speaker_vid_idx)
    else:
        pc.speaker_vid_idx = None
    return pc


@torch.no_grad()
def render_lockstep(processor, clips, fade_out=False, audio_sr=16000, run_trimodal=True):
    """Render a list of PreparedClip in lock-step on the device.
    -> list of (out_dir_vec_trimodal [L,27], out_dir_vec [L,27], out_poses_trimodal [L,10,3], out_poses [L,10,3])
    numpy arrays per clip (dir vecs WITHOUT the mean, poses with it, like render_clip's locals)."""
    cfg = processor.s2ag_config_args
    dev = processor.device
    G, Tri = processor.s2ag_generator, processor.trimodal_generator
    T, P, n_pre = cfg.n_poses, processor.pose_dim, cfg.n_pre_poses
    stride = T - n_pre
    B = len(clips)
    C = max(c.n_chunks for c in clips)
    A = clips[0].audio_chunks.shape[1]
    n_chunks = torch.tensor([c.n_chunks for c in clips], dtype=torch.int32, device=dev)
    audio = np.zeros((B, C, A), dtype=np.float32)
    text = np.zeros((B, C, T), dtype=np.int64)
    seed = np.zeros((B, T, P + 1), dtype=np.float32)
    for b, c in enumerate(clips):
        audio[b, :c.n_chunks] = c.audio_chunks
        text[b, :c.n_chunks] = c.text_chunks
        k = len(c.seed_seq)
        seed[b, :k, :-1] = c.seed_seq
        seed[b, :n_pre, -1] = 1   # indicating bit for seed poses
    audio_d = torch.from_numpy(audio).to(dev)
    text_d = torch.from_numpy(text).to(dev)
    vids = None
    if cfg.z_type == 'speaker':
        vids = torch.tensor([c.speaker_vid_idx for c in clips], dtype=torch.int64, device=dev)
    cap = T + stride * (C - 1) + 2 * n_pre          # room for the fade-out padding
    streams = {}
    names = (("tri", Tri),) if run_trimodal else ()
    names += (("gen", G),)
    for name, _ in names:
        streams[name] = dict(result=torch.zeros(B, cap, P, device=dev), pre=torch.from_numpy(seed).to(dev).clone(),
                             nxt=torch.empty(B, T, P + 1, device=dev))
    num_mfcc = getattr(cfg, "num_mfcc", 14)
    for c in range(C):
        a_c = audio_d[:, c]
        mfcc_c = ops.mfcc_features(a_c, audio_sr, num_mfcc)          # get_mfcc_features per chunk (:1249-1252)
        for name, net in names:
            st = streams[name]
            if name == "tri":
                out, *_ = net(st["pre"], text_d[:, c], a_c, vids)
            else:
                out, *_ = net(st["pre"], text_d[:, c], mfcc_c, vids)
            ops.longform_blend(out, st["result"], c, n_pre, pre_next=st["nxt"], n_chunks=n_chunks)
            # clips that are already finished keep feeding their (ignored) stale seed; live ones take the new one
            st["pre"], st["nxt"] = st["nxt"], st["pre"]
    lengths = T + stride * (n_chunks - 1)
    fps = cfg.motion_resampling_framerate
    mean = torch.tensor(np.squeeze(np.array(cfg.mean_dir_vec)), dtype=torch.float32, device=dev)
    outs = {}
    for name, _ in names:
        res, ln = streams[name]["result"], lengths
        if fade_out:
            start = torch.tensor([int(lengths[b]) - int(clips[b].end_padding / audio_sr * fps) for b in range(B)],
                                 dtype=torch.int32, device=dev)
            ln = ops.fade_out(res, lengths.to(torch.int32).contiguous(), start, n_pre)
        poses = ops.dir_vec_to_pose(res, mean)
        outs[name] = (res.cpu().numpy(), poses.cpu().numpy(), ln.cpu().numpy())
    ret = []
    for b in range(B):
        g_vec, g_pose, g_len = (outs["gen"][i][b] for i in range(3))
        if run_trimodal:
            t_vec, t_pose, t_len = (outs["tri"][i][b] for i in range(3))
            ret.append((t_vec[:t_len], g_vec[:g_len], t_pose[:t_len], g_pose[:g_len]))
        else:
            ret.append((None, g_vec[:g_len], None, g_pose[:g_len]))
    return ret
