#!/usr/bin/env python
"""Development aid: device time of every dense contraction of one eager GAN step (s2ag_debug_flags bit 5).

usage: python tools/gemm_trace.py [clips] 2> trace.txt   -- prints the per-shape table on stdout
"""
import collections
import os
import re
import sys
import tempfile
from argparse import Namespace as NS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from speech2affective_gestures_b200 import _C, ops  # noqa: E402
from speech2affective_gestures_b200.config import namespace as config_namespace  # noqa: E402
from speech2affective_gestures_b200.processor_v2 import Processor  # noqa: E402
from speech2affective_gestures_b200.synthetic import make_data_loader, synthetic_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N_WORDS, N_SPEAKERS, AUDIO_LEN = 20000, 1370, 36267
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
lib = _C.lib()
cfg = config_namespace()
pargs = NS(no_cuda=False, work_dir_s2ag=None, save_log=False, print_log=False, train_s2ag=True, batch_size=B,
           s2ag_num_epoch=1, val_interval=1, save_interval=10)
dl = make_data_loader(8, 8, 8, n_words=N_WORDS, n_speakers=N_SPEAKERS)
torch.manual_seed(1234)
ops.manual_seed(1234)
pr = Processor(ROOT, pargs, cfg, dl, 27, 3, 16000)
pr.meta_info["epoch"] = 1
for net in (pr.s2ag_generator, pr.s2ag_discriminator):
    net.train()
pr.trimodal_generator.train()
pr.use_side_stream = False
host = synthetic_batch(B, None, N_WORDS, N_SPEAKERS, AUDIO_LEN, seed=1234, pin=True)
static_in = tuple(t.to(dev) for t in host)
for _ in range(3):
    pr.gan_step_async(*static_in, True)
torch.cuda.synchronize()

# capture the C library's stderr lines
sys.stderr.flush()
tmp = tempfile.TemporaryFile(mode="w+b")
saved = os.dup(2)
os.dup2(tmp.fileno(), 2)
lib.s2ag_debug_flags(32)
pr.gan_step_async(*static_in, True)
torch.cuda.synchronize()
lib.s2ag_debug_flags(0)
os.dup2(saved, 2)
tmp.seek(0)
lines = tmp.read().decode().splitlines()

pat = re.compile(r"\[gemm\]\s+([\d.]+) us\s+M=(\d+) N=(\d+) K=(\d+) batch=(\d+) splitk=(\d+)\s+([\d.]+) TF/s\s+(.*)")
agg = collections.OrderedDict()
for ln in lines:
    m = pat.match(ln)
    if not m:
        continue
    us = float(m.group(1))
    sig = re.search(r"LdA = ([^;]+); LdB = ([^;]+); Epi = ([^;\]]+)", m.group(8))
    kind = ("%s | %s | %s" % sig.groups()) if sig else m.group(8)[:60]
    kind = kind.replace("s2ag::", "")
    key = (int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)), int(m.group(6)), kind)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
print("one eager GAN step, %d clips: %d contractions, %.1f us total" % (B, sum(v[0] for v in agg.values()), tot))
print("%6s %9s %6s %8s  %6s %6s %8s %5s %3s  %s" % ("calls", "total_us", "share", "avg_us", "M", "N", "K", "batch", "sk", "TF/s  loaders"))
for key, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, K, nb, sk, kind = key
    tf = 2.0 * M * N * K * nb * n / (us * 1e-6) * 1e-12
    print("%6d %9.1f %5.1f%% %8.1f  %6d %6d %8d %5d %3d  %5.1f  %s" % (n, us, 100 * us / tot, us / n, M, N, K, nb, sk, tf, kind))
