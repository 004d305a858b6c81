#!/usr/bin/env python
"""BASELINE config 5: inference generate path, batch 4096, 300-frame long-form synthesis (10 chunks of 34 frames, stride 30,
seed hand-off + linear blend on the device) on one GPU: new frames/s.  CPU arm: the oracle generator on a bounded sample."""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace as NS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--chunks", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=64)
    a = ap.parse_args()
    from speech2affective_gestures_b200 import _C
    from speech2affective_gestures_b200.processor_v2 import Processor
    from speech2affective_gestures_b200.synthetic import make_data_loader
    dev = torch.device("cuda:0")
    from speech2affective_gestures_b200.config import namespace as config_namespace
    cfg = config_namespace()
    pargs = NS(no_cuda=False, work_dir_s2ag=None, save_log=False, print_log=False, train_s2ag=True, batch_size=a.batch,
               s2ag_num_epoch=1, val_interval=1, save_interval=10)
    pr = Processor(ROOT, pargs, cfg, make_data_loader(8, 8, 8, n_words=20000, n_speakers=1370), 27, 3, 16000)
    B, C = a.batch, a.chunks
    g = torch.Generator().manual_seed(3)
    text = torch.zeros(B, C, 34, dtype=torch.int64)
    text[:, :, ::4] = torch.randint(4, 20000, (B, C, 9), generator=g)
    mfcc = (torch.randn(B, C, 37, 71, generator=g) * 0.1)
    vid = torch.randint(0, 1370, (B,), generator=g)
    text, mfcc, vid = text.to(dev), mfcc.to(dev), vid.to(dev)
    n0 = _C.lib().s2ag_launch_count()
    out = pr.synthesize_long_form(text, mfcc, None, vid)
    torch.cuda.synchronize()
    launches = _C.lib().s2ag_launch_count() - n0
    assert out.shape == (B, 34 + 30 * (C - 1), 27) and torch.isfinite(out).all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        pr.synthesize_long_form(text, mfcc, None, vid)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    frames = B * out.shape[1]
    # CPU arm: oracle PoseGenerator eval forward on a bounded sample, all host threads
    cpu = None
    if a.cpu_sample > 0:
        import s2ag_oracle as O  # CPU arm only
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from common import sd_cpu
        torch.set_num_threads(os.cpu_count() or 1)
        sd = O.as_leaves(sd_cpu(pr.s2ag_generator))
        b = a.cpu_sample
        pre = torch.zeros(b, 34, 28)
        eps = torch.zeros(b, 16)
        t0 = time.perf_counter()
        for c in range(C):
            with torch.no_grad():
                O.pose_generator(sd, pre, text[:b, c].cpu(), mfcc[:b, c].cpu(), vid[:b].cpu(), eps, False)
        dt = time.perf_counter() - t0
        cpu = {"value": b * out.shape[1] / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "%d clips x %d chunks, oracle PoseGenerator eval forward" % (b, C)}
    print(json.dumps({"metric": "long-form synthesis frames/sec (300-frame clips, 10 lock-step chunks)", "value": frames / (ms * 1e-3),
                      "unit": "frames/s", "n_gpus": 1, "ms_per_batch": ms, "batch": B, "frames_per_clip": out.shape[1],
                      "gpu_launches": int(launches), "dtype": "f32", "data": "synthetic", "cpu_baseline": cpu}))


if __name__ == "__main__":
    main()
