"""TEST INFRASTRUCTURE (never imported by the product path).  numpy restatement of the host-side pieces around the
generator that SURVEY 8(f) ranks "next": the MFCC front-end, the long-form chunk blend / fade-out, the
direction-vector -> joint conversion and the evaluation metrics.  Every function cites the reference lines it
follows (paths relative to the reference checkout).

MFCC: the arithmetic lives in a third-party dependency that is ABSENT from the reference checkout and from this
image: `librosa.feature.mfcc` (requirements.txt lists `librosa` unpinned; the README's Python 3.7 / numpy 1.20.3 era
is librosa 0.8-0.9).  Its published algorithm is restated below; **parity unpinned** for that function (no librosa
here to produce a fixture) -- the call site utils/common.py:340-349 (scale 1/1000, first/second row differences) IS
pinned: oracle/gen_golden.py runs the reference's unmodified `get_mfcc_features` with this restatement injected as
`mfcc`.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------ librosa.feature.mfcc
def hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=16000, n_fft=2048, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(htk=False, norm='slaney') -> float32 [n_mels, 1 + n_fft//2]."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def dct_matrix(n_out, n_in):
    """scipy.fftpack.dct(type=2, norm='ortho') along an axis of length n_in, first n_out rows, as a matrix."""
    n = np.arange(n_in)
    k = np.arange(n_out)[:, None]
    m = 2.0 * np.cos(np.pi * k * (2 * n + 1) / (2.0 * n_in))
    m[0] *= np.sqrt(1.0 / (4 * n_in))
    m[1:] *= np.sqrt(1.0 / (2 * n_in))
    return m


def mfcc_librosa(y, sr=16000, n_mfcc=14, n_fft=2048, hop=512, n_mels=128, pad_mode="reflect"):
    """librosa.feature.mfcc(y, sr, n_mfcc): STFT (hann, center, reflect pad) -> |.|^2 -> mel(128, slaney) ->
    power_to_db(ref=1, amin=1e-10, top_db=80) -> DCT-II ortho -> first n_mfcc rows.  -> [n_mfcc, 1 + len(y)//hop]"""
    y = np.asarray(y, dtype=np.float32)
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (len(yp) - n_fft) // hop
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft))  # scipy get_window('hann', fftbins=True)
    frames = np.stack([yp[i * hop:i * hop + n_fft] for i in range(n_frames)], axis=1)  # [n_fft, frames]
    spec = np.fft.rfft(frames * win[:, None].astype(np.float32), axis=0).astype(np.complex64)
    power = np.abs(spec) ** 2
    mel = mel_filterbank(sr, n_fft, n_mels) @ power
    log_spec = 10.0 * np.log10(np.maximum(1e-10, mel))
    log_spec = np.maximum(log_spec, log_spec.max() - 80.0)
    return (dct_matrix(n_mfcc, n_mels) @ log_spec).astype(np.float32)


def get_mfcc_features(audio, sr, num_mfcc, mfcc=mfcc_librosa):
    """utils/common.py:340-349."""
    m = mfcc(audio, sr=sr, n_mfcc=num_mfcc) / 1000.
    d1 = m[2:] - m[1:-1]
    d2 = d1[1:] - d1[:-1]
    return np.concatenate((m, d1, d2), axis=0)


# ------------------------------------------------------------------------------------------------ skeleton
# utils/ted_db_utils.py:14 (parent, child, bone length)
DIR_VEC_PAIRS = [(0, 1, 0.26), (1, 2, 0.18), (2, 3, 0.14), (1, 4, 0.22), (4, 5, 0.36),
                 (5, 6, 0.33), (1, 7, 0.22), (7, 8, 0.36), (8, 9, 0.33)]


def convert_dir_vec_to_pose(vec):
    """utils/ted_db_utils.py:81-102: joint[child] = joint[parent] + length * dir_vec; joint 0 at the origin."""
    vec = np.asarray(vec, dtype=np.float64)
    vec = vec.reshape(vec.shape[:-1] + (-1, 3)) if vec.shape[-1] != 3 else vec
    joint = np.zeros(vec.shape[:-2] + (10, 3))
    for j, (a, b, ln) in enumerate(DIR_VEC_PAIRS):
        joint[..., b, :] = joint[..., a, :] + ln * vec[..., j, :]
    return joint


# ------------------------------------------------------------------------------------------------ long-form synthesis
def blend_chunks(chunks, n_pre):
    """processor_v2.py:1303-1331: chunks = list of [T, P] outputs; overlap of n_pre frames blended linearly,
    out[j] = prev[j]*(n-j)/(n+1) + next[j]*(j+1)/(n+1); previous chunk loses its last n_pre frames."""
    out_list = []
    for seq in chunks:
        seq = np.array(seq, dtype=np.float32, copy=True)
        if out_list:
            last = out_list[-1][-n_pre:]
            out_list[-1] = out_list[-1][:-n_pre]
            n = len(last)
            for j in range(n):
                seq[j] = last[j] * (n - j) / (n + 1) + seq[j] * (j + 1) / (n + 1)
        out_list.append(seq)
    return np.vstack(out_list)


def fade_out(out_dir_vec, end_padding_samples, audio_sr, fps, n_pre, pose_dim):
    """processor_v2.py:1334-1391 for one stream: zero the tail, then replace [start_frame, end_frame) by the weighted
    (w = 5 at both ends) quadratic least-squares fit of itself.  -> (array possibly padded, start_frame, end_frame)"""
    out = np.array(out_dir_vec, dtype=np.float64, copy=True)
    n_smooth = n_pre
    start_frame = len(out) - int(end_padding_samples / audio_sr * fps)
    end_frame = start_frame + n_smooth * 2
    if len(out) < end_frame:
        out = np.pad(out, [(0, end_frame - len(out)), (0, 0)], mode='constant')
    out[end_frame - n_smooth:] = np.zeros(pose_dim)
    y = out[start_frame:end_frame]
    x = np.arange(y.shape[0])
    w = np.ones(len(y)); w[0] = 5; w[-1] = 5
    co = np.polyfit(x, y, 2, w=w)
    out[start_frame:end_frame] = np.stack([np.poly1d(co[:, k])(x) for k in range(y.shape[1])], axis=1)
    return out, start_frame, end_frame


# ------------------------------------------------------------------------------------------------ metrics
def push_samples_metrics(out_dir_vec, target, mean_dir_vec, n_poses, n_pre):
    """processor_v2.py:738-774 -> (L1 loss, joint MAE, acceleration difference) of one batch."""
    out_dir_vec = np.asarray(out_dir_vec, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    mean = np.asarray(mean_dir_vec, dtype=np.float64).squeeze()
    loss = np.mean(np.abs(out_dir_vec - target))
    out_j = convert_dir_vec_to_pose(out_dir_vec + mean)
    tgt_j = convert_dir_vec_to_pose(target + mean)
    if out_j.shape[1] == n_poses:
        diff = out_j[:, n_pre:] - tgt_j[:, n_pre:]
    else:
        diff = out_j - tgt_j[:, n_pre:]
    mae = np.mean(np.abs(diff))
    accel = np.mean(np.abs(np.diff(tgt_j, n=2, axis=1) - np.diff(out_j, n=2, axis=1)))
    return loss, mae, accel


# ------------------------------------------------------------------------------------------------ synthetic clip source
def synthetic_clip(seed, duration=20.0, sr=16000, pose_fps=25.0, n_vocab_words=60, start_time=3.0):
    """A TED-shaped clip record in the LMDB layout generate_gestures_by_dataset reads (processor_v2.py:1486-1495):
    [words, poses, None, audio, None, None, {'vid','start_frame_no','end_frame_no','start_time','end_time'}].
    Poses: a smooth random upper-body skeleton (10 joints x 3); audio: harmonic 'speech' bursts; words 'w<i>' with
    (start, end) times in absolute video time."""
    rng = np.random.RandomState(seed)
    n_pose = int(round(duration * pose_fps))
    t = np.arange(n_pose) / pose_fps
    base = convert_dir_vec_to_pose(np.array([0.0154009, -0.9690125, -0.0884354, -0.0022264, -0.8655276, 0.4342174,
                                             -0.0035145, -0.8755367, -0.4121039, -0.9236511, 0.3061306, -0.0012415,
                                             -0.5155854, 0.8129665, 0.0871897, 0.2348464, 0.1846561, 0.8091402,
                                             0.9271948, 0.2960011, -0.013189, 0.5233978, 0.8092403, 0.0725451,
                                             -0.2037076, 0.1924306, 0.8196916]))
    wob = sum(rng.normal(0, 0.03, size=(1, 10, 3)) * np.sin(2 * np.pi * rng.uniform(0.1, 1.5) * t + rng.uniform(0, 6.28)
                                                            )[:, None, None] for _ in range(4))
    poses = (base[None] + wob).astype(np.float32).reshape(n_pose, -1)
    n = int(round(duration * sr))
    ta = np.arange(n) / sr
    f0 = rng.uniform(100, 200)
    audio = sum(np.sin(2 * np.pi * f0 * h * ta + rng.uniform(0, 6.28)) / h ** 1.3 for h in range(1, 20))
    audio = audio * (0.2 + 0.8 * (np.sin(2 * np.pi * 1.7 * ta) > -0.3)) + rng.normal(0, 3e-3, n)
    audio = (0.5 * audio / np.abs(audio).max()).astype(np.float32)
    words, cur = [], start_time + 0.1
    while cur < start_time + duration - 0.6:
        d = rng.uniform(0.15, 0.5)
        words.append(['w%d' % rng.randint(4, n_vocab_words), float(cur), float(cur + d)])
        cur += d + rng.uniform(0.02, 0.6)
    meta = {'vid': 'vid%03d' % seed, 'start_frame_no': 10, 'end_frame_no': 10 + n_pose,
            'start_time': start_time, 'end_time': start_time + duration}
    return [words, poses, None, audio, None, None, meta]
