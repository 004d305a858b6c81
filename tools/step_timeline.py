#!/usr/bin/env python
"""Development aid: kernel timeline (CUPTI via torch.profiler) of one graph-replayed GAN step: per-stream busy time,
gaps on each stream, top kernels.  usage: python tools/step_timeline.py [clips]"""
import collections
import os
import sys
from argparse import Namespace as NS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from speech2affective_gestures_b200 import _C, ops  # noqa: E402
from speech2affective_gestures_b200.config import namespace as config_namespace  # noqa: E402
from speech2affective_gestures_b200.processor_v2 import Processor  # noqa: E402
from speech2affective_gestures_b200.synthetic import make_data_loader, synthetic_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N_WORDS, N_SPEAKERS, AUDIO_LEN = 20000, 1370, 36267
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
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
host = synthetic_batch(B, None, N_WORDS, N_SPEAKERS, AUDIO_LEN, seed=1234, pin=True)
pr.capture_step(B, train=True, warmup=2)
pr.load_static_inputs(*[host[i] for i in (0, 1, 2, 3, 4)])
for _ in range(3):
    pr.replay_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pr.replay_step()
    torch.cuda.synchronize()
if os.environ.get("S2AG_TRACE"):
    prof.export_chrome_trace(os.environ["S2AG_TRACE"])
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_index", 0)) for e in evs))
if not ks:
    print("no kernel events"); sys.exit(0)
t0, t1 = ks[0][0], max(k[1] for k in ks)
print("step span %.1f us, %d kernels" % (t1 - t0, len(ks)))
# concurrency histogram: time with n kernels in flight
pts = []
for s, e, _, _ in ks:
    pts.append((s, 1)); pts.append((e, -1))
pts.sort()
hist = collections.Counter()
cur, last = 0, t0
for t, d in pts:
    hist[cur] += t - last
    last = t
    cur += d
print("time with n kernels in flight: " + ", ".join("%d: %.0f us" % (n, v) for n, v in sorted(hist.items())))
agg = collections.defaultdict(lambda: [0, 0.0])
for s_, e_, name, _ in ks:
    k = name.split("(")[0]
    k = k[:k.index("<")] if "<" in k and not k.startswith("void at::") else k[:60]
    agg[k][0] += 1
    agg[k][1] += e_ - s_
tot_k = sum(v[1] for v in agg.values())
print("kernel time summed over streams %.1f us (%.2fx the span)" % (tot_k, tot_k / (t1 - t0)))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("  %5d  %9.1f us  %5.1f%%  %s" % (n, us, 100 * us / tot_k, k[-70:]))
# coarse phases: every 0.5 ms, which kernels dominate
bins = collections.defaultdict(lambda: collections.Counter())
W = 500.0
for s, e, name, _ in ks:
    b0, b1 = int((s - t0) // W), int((e - t0) // W)
    for b in range(b0, b1 + 1):
        lo, hi = max(s, t0 + b * W), min(e, t0 + (b + 1) * W)
        if hi > lo:
            bins[b][name.split("(")[0].split("<")[0][-38:]] += hi - lo
for b in sorted(bins):
    tot = sum(bins[b].values())
    top = ", ".join("%s %.0f" % (n, v) for n, v in bins[b].most_common(3))
    print("  %5.1f ms  busy %4.0f%%  %s" % (b * W / 1000, 100 * tot / W, top))
