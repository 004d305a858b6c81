"""Synthetic TED-shaped data that honours the tensor contracts of the reference's loader/npz cache
(loader_v2.py:470-539 `TedDBParams`, processor_v2.py:222-271 `load_cache`), so that the
Processor can be driven without the TED dataset (there is no network / dataset in this build).

The npz-cache schema the reference Processor samples from (processor_v2.py:589-638):
    extended_word_seq int64 [n, 34]     vec_seq float32 [n, 34, 27]     audio int16 [n, 36267]
    audio_max float [n]                 mfcc_features float16 [n, 37, 71]   vid_indices int64 [n]
"""
import numpy as np
import torch


class Vocab:
    """Speaker / language model stand-in.  The reference recognises a speaker model by its class
    NAME being 'Vocab' (net/multimodal_context_net_v2.py:469) and reads `.n_words`/`.word2index`."""

    PAD_token, SOS_token, EOS_token, UNK_token = 0, 1, 2, 3   # utils/vocab.py:9-12

    def __init__(self, name, n_words, word_embedding_weights=None):
        self.name = name
        self.n_words = n_words
        self.word2index = {"w%d" % i: i for i in range(n_words)}
        self.word_embedding_weights = word_embedding_weights

    def get_word_index(self, word):
        """utils/vocab.py:64-68"""
        return self.word2index.get(word, self.UNK_token)


class SyntheticTedData:
    """One split ('train'/'val'/'test') of synthetic clips with TedDBParams' derived sizes."""

    def __init__(self, n_samples, n_words=20000, n_speakers=1370, n_poses=34, pose_dim=27, audio_sr=16000,
                 fps=15, seed=1234, audio_length=None, lang_model=None, speaker_model=None):
        self.n_samples = n_samples
        self.n_poses = n_poses
        self.expected_audio_length = audio_length or int(round(n_poses / fps * audio_sr))  # 36267 (loader_v2.py:480)
        self.expected_spectrogram_length = int(round(n_poses / fps * audio_sr / 512)) + 1
        self.num_mfcc_combined = 37                                                       # loader_v2.py:483
        self.lang_model = lang_model or Vocab("words", n_words)
        self.speaker_model = speaker_model or Vocab("vid", n_speakers)
        rng = np.random.RandomState(seed)
        mfcc_len = int(np.ceil(self.expected_audio_length / 512))                         # processor_v2.py:124
        words = np.zeros((n_samples, n_poses), dtype=np.int64)
        for i in range(n_samples):  # "10-token text": ten frame slots carry word ids >= 4 (processor_v2.py:408-432)
            pos = rng.choice(n_poses, size=min(10, n_poses), replace=False)
            words[i, pos] = rng.randint(4, self.lang_model.n_words, size=len(pos))
        audio_f = rng.uniform(-0.5, 0.5, size=(n_samples, self.expected_audio_length)).astype(np.float32)
        audio_max = np.abs(audio_f).max(axis=1)
        self.samples = {
            "extended_word_seq": words,
            "vec_seq": np.clip(rng.normal(0, 0.3, size=(n_samples, n_poses, pose_dim)), -2, 2).astype(np.float32),
            "audio": np.round(audio_f / audio_max[:, None] * 32767).astype(np.int16),
            "audio_max": audio_max.astype(np.float32),
            "mfcc_features": rng.normal(0, 0.1, size=(n_samples, self.num_mfcc_combined, mfcc_len)).astype(np.float16),
            "vid_indices": rng.randint(0, self.speaker_model.n_words, size=n_samples).astype(np.int64),
        }


def make_data_loader(n_train=512, n_val=128, n_test=128, **kw):
    """dict with the keys the reference Processor reads (processor_v2.py:112-131)."""
    tr = SyntheticTedData(n_train, seed=1234, **kw)
    share = dict(lang_model=tr.lang_model, speaker_model=tr.speaker_model)
    kw2 = {k: v for k, v in kw.items() if k not in ("lang_model", "speaker_model")}
    return {"train_data_s2ag": tr,
            "val_data_s2ag": SyntheticTedData(n_val, seed=1235, **kw2, **share),
            "test_data_s2ag": SyntheticTedData(n_test, seed=1236, **kw2, **share)}


def synthetic_batch(B, device, n_words=20000, n_speakers=1370, audio_length=36267, seed=1234, pin=False):
    """Seeded device batch in the shapes of SURVEY 8d: (in_text, in_audio, in_mfcc, target_poses, vid_indices)."""
    g = torch.Generator().manual_seed(seed)
    T, P = 34, 27
    target = (torch.randn(B, T, P, generator=g) * 0.3).clamp_(-2, 2)
    text = torch.zeros(B, T, dtype=torch.int64)
    pos = torch.rand(B, T, generator=g).argsort(dim=1)[:, :10]
    text.scatter_(1, pos, torch.randint(4, n_words, (B, 10), generator=g))
    audio = torch.rand(B, audio_length, generator=g) - 0.5
    mfcc = torch.randn(B, 37, 71, generator=g) * 0.1
    vid = torch.randint(0, n_speakers, (B,), generator=g)
    out = (text, audio, mfcc, target, vid)
    if pin:
        out = tuple(t.pin_memory() for t in out)
    if device is not None and str(device) != "cpu":
        out = tuple(t.to(device, non_blocking=True) for t in out)
    return out
