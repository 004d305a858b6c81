"""TEST / BENCH INFRASTRUCTURE.  Imports the UNMODIFIED reference (recipes A and B) from
`$S2AG_REFERENCE` / `/root/reference` when that exists, else from the staged copy `oracle/_ref/`
(oracle/build_ref.py), with the third-party modules the image lacks stubbed at import time, and builds a
`Processor` around `Processor.__new__` (the real constructor needs the TED caches and two external checkpoints).

Only tests/, oracle/gen_golden.py and bench.py's reference arms import this.
"""
import os
import sys
import types
from types import SimpleNamespace as NS
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = ['librosa', 'librosa.feature', 'librosa.display', 'lmdb', 'matplotlib', 'matplotlib.pyplot',
         'matplotlib.ticker', 'matplotlib.animation', 'mpl_toolkits', 'mpl_toolkits.mplot3d',
         'python_speech_features', 'h5py', 'umap', 'soundfile', 'fasttext', 'transforms3d', 'configargparse',
         'pyttsx3', 'nltk', 'nltk.corpus']


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return MagicMock()


def reference_root():
    if os.environ.get("S2AG_REFERENCE") == "staged":  # force the staged copy (what the GPU box uses)
        p = os.path.join(HERE, "_ref")
        return p if os.path.isfile(os.path.join(p, "processor_v2.py")) else None
    for p in (os.environ.get("S2AG_REFERENCE"), "/root/reference", os.path.join(HERE, "_ref")):
        if p and os.path.isfile(os.path.join(p, "processor_v2.py")):
            return p
    return None


_loaded = {}


def load(device="cpu"):
    """-> the reference's `processor_v2` module (exposes .Processor, .PoseGenerator, .PGT, .AffDiscriminator, .CDT)."""
    if "RP" in _loaded:
        return _loaded["RP"]
    root = reference_root()
    if root is None:
        raise FileNotFoundError("no reference checkout and oracle/_ref is not staged (run oracle/build_ref.py)")
    sys.dont_write_bytecode = True  # the reference mount is read-only
    import importlib.util
    for n in STUBS:
        if n in sys.modules:
            continue
        try:
            if importlib.util.find_spec(n) is not None:
                continue
        except (ImportError, ValueError):
            pass
        m = _Stub(n)
        m.__path__ = []
        sys.modules[n] = m
    import torch
    if str(device) == "cpu":
        # AffEncoder hard-codes .cuda() (net/multimodal_context_net_v2.py:106,115,163): identity for the CPU arm,
        # also on a box that has a GPU (the CPU arm must stay on the host cores)
        torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, root)
    import processor_v2 as RP
    _loaded["RP"], _loaded["root"] = RP, root
    return RP


def root_used():
    return _loaded.get("root")


def speaker_vocab(n_spk):
    from utils.vocab import Vocab  # the reference's class: its NAME must be 'Vocab' (net/...v2.py:469)
    spk = Vocab('vid', insert_default_tokens=False)
    for i in range(n_spk):
        spk.index_word('v%d' % i)
    return spk


def make_processor(cfg_dict, n_words, n_spk, device="cpu", derand=False):
    """Recipe B: the unmodified reference Processor with its three networks and two Adam optimisers."""
    import torch
    RP = load(device)
    cfg = NS(**cfg_dict)
    spk = speaker_vocab(n_spk)
    pr = RP.Processor.__new__(RP.Processor)
    pr.s2ag_config_args, pr.meta_info, pr.use_mfcc = cfg, dict(epoch=1, iter=0), True
    pr.device = torch.device(device)
    pr.pose_dim = 27
    pr.trimodal_generator = RP.PGT(cfg, pose_dim=27, n_words=n_words, word_embed_size=cfg.wordembed_dim,
                                   word_embeddings=None, z_obj=spk).to(device)
    pr.s2ag_generator = RP.PoseGenerator(cfg, pose_dim=27, n_words=n_words, word_embed_size=cfg.wordembed_dim,
                                         word_embeddings=None, mfcc_length=71, num_mfcc=37, time_steps=34,
                                         z_obj=spk).to(device)
    pr.s2ag_discriminator = RP.AffDiscriminator(27).to(device)
    if derand:
        for net in (pr.trimodal_generator, pr.s2ag_generator, pr.s2ag_discriminator):
            for m in net.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
                if isinstance(m, torch.nn.GRU):
                    m.dropout = 0.0
    pr.s2ag_gen_optimizer = torch.optim.Adam(pr.s2ag_generator.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
    pr.s2ag_dis_optimizer = torch.optim.Adam(pr.s2ag_discriminator.parameters(),
                                             lr=cfg.learning_rate * cfg.discriminator_lr_weight, betas=(0.5, 0.999))
    return pr
