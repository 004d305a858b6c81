#!/usr/bin/env python
"""Benchmark of the hot path: one full GAN training iteration (processor_v2.py:776-957 semantics:
D step + G step, forward/backward/Adam, dropout as shipped) over a synthetic TED-shaped batch.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels), one rank per GPU
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU cores

A "step" = one GAN iteration over `--batch-per-gpu` clips per rank (weak scaling: 256 clips/GPU, i.e.
BASELINE.json's batch-2048 configuration at 8 GPUs).  Prints ONE JSON line (rank 0).
  value  : clips/s, whole job, inputs resident in HBM, the step replayed as a CUDA graph, CUDA-event timed
  e2e    : clips/s through the public Processor API with HOST (pinned) inputs: per step H2D of the batch,
           the graph replay, D2H read of the 8-float loss/metric vector
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace as NS

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WORDS, N_SPEAKERS, AUDIO_LEN = 20000, 1370, 36267
E2E_PASSES = 3   # the end-to-end loop is timed this many times (host jitter); fastest reported, all listed
FLOP_PER_CLIP = 3.396e9  # reference step, FlopCounterMode


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "bf16_burst": float(d.get("bf16_tflops", 1590.0)),
                    "bf16_sustained": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


def cfg_namespace():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import s2ag_oracle as O
    return NS(**O.CFG), O


# ----------------------------------------------------------------------------- CPU arm (oracle port of the reference)
def cpu_reference_clips_per_s(sample_b, steps, warmup, threads):
    """The reference algorithm (oracle/s2ag_oracle.py restatement; the Python reference itself cannot
    travel to the GPU box) on the host cores: full GAN iteration, fp32, `threads` torch threads."""
    import torch
    cfg, O = cfg_namespace()
    torch.set_num_threads(threads)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import build_nets, sd_cpu
    G, T, D, _ = build_nets("full", N_WORDS, N_SPEAKERS, torch.device("cpu"))
    g_sd, d_sd, t_sd = (O.as_leaves(sd_cpu(n)) for n in (G, D, T))
    del G, T, D
    batch, eps_list, rand_idx = O.synthetic_batch(sample_b, N_WORDS, N_SPEAKERS, AUDIO_LEN, 1234)
    state = {}
    for _ in range(warmup):
        O.gan_step(g_sd, d_sd, t_sd, batch, eps_list, rand_idx, O.CFG, state, train=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.gan_step(g_sd, d_sd, t_sd, batch, eps_list, rand_idx, O.CFG, state, train=True)
    dt = (time.perf_counter() - t0) / steps
    return sample_b / dt, dt


def workload_text(B):
    """the same workload description on both arms"""
    return ("full GAN step (D step + G step: 3 G fwd, 1 frozen tri-modal fwd incl. WavEncoder, 3 D fwd, G+D bwd, 2 Adam), "
            "%d clips/GPU, 34 frames x 27-D, audio %d samples, 10-token text, n_words=%d, dropout as shipped"
            % (B, AUDIO_LEN, N_WORDS))


def reference_step_clips_per_s(sample_b, steps, warmup, threads, device="cpu"):
    """The UNMODIFIED reference `Processor.forward_pass_s2ag` (oracle/ref_loader.py: the checkout, else the copy staged
    in git-ignored oracle/_ref by oracle/build_ref.py), train=True, epoch 1, dropout as shipped, fp32, TF32 off.
    device='cpu': the reference's own CPU path on `threads` host threads; device='cuda': context arm, stock
    PyTorch/cuDNN/cuBLAS on the same B200 ('the bar for each kernel')."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    cfg, O = cfg_namespace()
    torch.set_num_threads(threads)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1234)
    pr = ref_loader.make_processor(O.CFG, N_WORDS, N_SPEAKERS, device=device)
    for n in (pr.s2ag_generator, pr.s2ag_discriminator):
        n.train()
    batch, _, _ = O.synthetic_batch(sample_b, N_WORDS, N_SPEAKERS, AUDIO_LEN, 1234)
    batch = tuple(t.to(device) for t in batch)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    for _ in range(warmup):
        pr.forward_pass_s2ag(batch[0], batch[1], batch[2], batch[3], batch[4], train=True)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        pr.forward_pass_s2ag(batch[0], batch[1], batch[2], batch[3], batch[4], train=True)
    sync()
    dt = (time.perf_counter() - t0) / steps
    return sample_b / dt, dt, ref_loader.root_used()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    gpu = args.impl == "reference-gpu"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    have_ref = ref_loader.reference_root() is not None
    if gpu and not have_ref:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref is not staged (oracle/build_ref.py)"}))
        return
    if have_ref:
        sample_b = args.ref_batch if not gpu else args.batch_per_gpu
        steps, warm = (max(1, min(args.steps, 3)), 1) if not gpu else (max(1, min(args.steps, 10)), 2)
        v, dt, root = reference_step_clips_per_s(sample_b, steps, warm, threads, "cuda" if gpu else "cpu")
        kind = "reference"
        what = ("UNMODIFIED reference Processor.forward_pass_s2ag (%s), %s, train=True, dropout as shipped, fp32"
                % ("oracle/_ref staged copy" if root.endswith("_ref") else "reference checkout",
                   "stock PyTorch/cuDNN on cuda:0, TF32 off" if gpu else "%d host threads" % threads))
    else:
        sample_b, steps, warm = args.cpu_sample_batch, max(1, min(args.steps, 3)), 1
        v, dt = cpu_reference_clips_per_s(sample_b, steps, warm, threads)
        kind = "port"
        what = "oracle port of processor_v2.forward_pass_s2ag (reference not staged), dropout off"
    line = {
        "impl": args.impl, "metric": "gesture-clips/sec (34-frame, 27-D pose), full GAN training step",
        "value": v, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.batch_per_gpu),
                   "sample": "bounded sample of that workload: %d clips per timed iteration" % sample_b},
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": threads if not gpu else 0, "kind": kind,
                         "sample": "%d timed GAN iterations of %d clips after %d warm-up; %s" % (steps, sample_b, warm, what)},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of the same shapes (profiles/,
# 256 clips); None where no capture of that kernel is committed
NCU_TRAFFIC = {"gru_recurrence_fwd": 65.92e6 + 53.71e6,   # profiles/r02_final_ncu_gru_persist_fwd.txt
               "gemm_gru_projection": 38848512,         # profiles/r01_ncu_gemm_umma_pk.txt
               # profiles/r02_final_ncu_tcn_fused.txt: reads only (x + the weight image); the 10.4 MB output is still
               # L2-resident when the kernel ends, so DRAM writes are ~0 in the capture
               "tcn_residual_block_fused_fwd": 11.98e6,
               # profiles/r02_ncu_wavencoder.txt: sum over the four launches (ncu flushes L2 between kernels, so the
               # raw conv2/conv3 intermediates -- L2-resident in a real step -- are counted as DRAM reads here)
               "wavencoder_fwd": 37.45e6 + (37.22e6 + 2.99e6) + (43.21e6 + 0.42e6) + 14.39e6}


def _time_call(torch, call, flush, reps=10, warm=3):
    for _ in range(warm):
        call()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()  # L2 flush between timed launches (160 MB > 126 MB L2)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(); call(); k1.record()
        torch.cuda.synchronize()
        tot += k0.elapsed_time(k1)
    return tot / reps


def roofline(B, dev, lib, clips_per_s_per_gpu=None):
    """Per-kernel rooflines, each kernel timed ALONE inside this process with CUDA events on its launch stream, L2
    flushed between launches.  `roofline` = the DOMINANT kernel by time share (profiles/: gru_persist_fwd, ~20 % of
    the step's kernel time); `roofline_kernels` = the kernels the north star names, with SURVEY 8(d)'s algorithmic
    FLOPs / bytes per clip x the clips one launch processes.  Tensor-bound kernels are reported against the measured
    bf16 burst peak although they run fp32-grade bf16x3 (3 MMAs per algorithmic MAC: ceiling 1/3 of that peak)."""
    import torch
    from speech2affective_gestures_b200 import _C, ops
    from speech2affective_gestures_b200.net.multimodal_context_net_v2 import WavEncoder
    pk = peaks()
    T, H = 34, 300
    flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)
    st = ops._stream(flush)
    psrc = pk["source"] + " (MEASURED_PEAKS.json)"
    out = []

    def entry(name, kernel, bound, ms, flops, nbytes, extra=None):
        e = {"name": name, "kernel": kernel, "bound": bound, "kernel_ms": ms, "algorithmic_flops": flops,
             "algorithmic_bytes": nbytes, "traffic": NCU_TRAFFIC.get(name), "peak_source": psrc,
             "l2": "160 MB buffer zeroed between timed launches"}
        if bound == "tensor":
            ach = flops / (ms * 1e-3) / 1e12
            e.update(achieved=ach, peak=pk["bf16_burst"], unit="TFLOP/s", frac=ach / pk["bf16_burst"])
        else:
            ach = nbytes / (ms * 1e-3) / 1e9
            e.update(achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"],
                     tensor_tflops=flops / (ms * 1e-3) / 1e12)
        if extra:
            e.update(extra)
        out.append(e)
        return e

    # (1) the recurrence of one generator-sized bidirectional GRU layer: 34 dependent steps, W_hh stationary
    M = B * T
    ws_floats = lib.s2ag_gru_fwd_ws_floats(B, T, H)
    gi = torch.randn(ws_floats, device=dev) * 0.5
    whh = torch.randn(2, 3 * H, H, device=dev) * 0.05
    bhh = torch.zeros(2, 3 * H, device=dev)
    o = torch.empty(B, T, 2 * H, device=dev)
    gates = torch.empty(M * 2 * 4 * H, device=dev)
    rec = lambda: _C.call("s2ag_gru_recurrence_fwd", ops._p(gi), ops._p(whh[0]), ops._p(whh[1]), ops._p(bhh[0]),
                          ops._p(bhh[1]), ops._p(o), ops._p(gates), B, T, H, st)
    ms = _time_call(torch, rec, flush)
    dom = entry("gru_recurrence_fwd", "gru_persist_fwd_kernel (tcgen05 + TMA bulk copies; one bidirectional layer, "
                "%d clips x %d steps x H=%d, gates saved for BPTT)" % (B, T, H), "tensor", ms,
                2.0 * M * (3 * H) * H * 2, M * 6 * H * 4 + M * 2 * H * 4 + M * 8 * H * 4,
                {"step_period_us": ms * 1e3 / T})
    # (2) the dense-contraction engine on its largest shape: GRU input projection [B*34, 600] x [600, 2*900]
    N, K = 1800, 600
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.05; y = torch.empty(M, N, device=dev)
    bias = torch.zeros(N, device=dev)
    ms = _time_call(torch, lambda: _C.call("s2ag_linear_fwd", ops._p(x), K, ops._p(w), ops._p(bias), ops._p(y), N, M, N,
                                           K, 0, 0.0, st), flush)
    entry("gemm_gru_projection", "pack_operand_kernel + gemm_umma_pk_kernel (tcgen05; %dx%dx%d)" % (M, N, K), "tensor",
          ms, 2.0 * M * N * K, (M * K + N * K + M * N) * 4)
    # (3) one causal-TCN residual block forward (weight-norm + 2 dilated convs + ReLU/residual), C=300, k=2, d=2, through
    #     the C ABI with preallocated buffers: (a) the default two-launch path, (b) the single-kernel block
    import ctypes
    C, dil = 300, 2
    xt = torch.randn(B, T, C, device=dev)
    v1, v2 = torch.randn(C, C, 2, device=dev) * 0.05, torch.randn(C, C, 2, device=dev) * 0.05
    g1, g2 = torch.ones(C, device=dev), torch.ones(C, device=dev)
    b1, b2 = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    w1, w2 = torch.empty(C, 2, C, device=dev), torch.empty(C, 2, C, device=dev)
    n1, n2 = torch.empty(C, device=dev), torch.empty(C, device=dev)
    y1, y2, yo = (torch.empty(B, T, C, device=dev) for _ in range(3))
    P_ = ops._p

    def tcn_two_launch():
        _C.call("s2ag_weight_norm_fwd", P_(v1), P_(g1), P_(w1), P_(n1), C, C, 2, st)
        _C.call("s2ag_weight_norm_fwd", P_(v2), P_(g2), P_(w2), P_(n2), C, C, 2, st)
        _C.call("s2ag_tcn_block_fwd", P_(xt), P_(w1), P_(b1), P_(w2), P_(b2), P_(y1), P_(y2), P_(yo), B, T, C, dil, 0.0,
                ctypes.c_uint64(0), None, st)
    ms = _time_call(torch, tcn_two_launch, flush)
    entry("tcn_residual_block_fwd", "weight_norm_fwd x2 + pack_operand x2 + gemm_umma_pk_kernel<LdConv,EpiTcn> x2 (d=2; the "
          "default path)", "tensor", ms, 24.48e6 * B, 81600.0 * B)
    n_ws = lib.s2ag_tcn_fused_ws_floats(T, C, dil)
    if n_ws > 0:
        wsf = torch.empty(n_ws, device=dev)
        ms = _time_call(torch, lambda: _C.call(
            "s2ag_tcn_block_fused_fwd", P_(xt), P_(v1), P_(g1), P_(b1), P_(v2), P_(g2), P_(b2), P_(w1), P_(w2), P_(n1),
            P_(n2), None, None, P_(yo), P_(wsf), B, T, C, dil, 0.0, ctypes.c_uint64(0), None, st), flush)
        entry("tcn_residual_block_fused_fwd", "tcn_pack_kernel + tcn_block_fused_kernel (umma_tcn.cu: one kernel per block, "
              "y1 kept in shared memory; opt-in, bound by the weight stream: DESIGN.md)", "tensor", ms, 24.48e6 * B,
              81600.0 * B)
    # (4) the frozen baseline's WavEncoder stack (4 strided Conv1d + 3 train-mode BN + LeakyReLU), raw audio in
    we = WavEncoder().to(dev)
    audio = torch.rand(B, AUDIO_LEN, device=dev) - 0.5
    with torch.no_grad():
        ms = _time_call(torch, lambda: we(audio), flush, reps=5)
    entry("wavencoder_fwd", "wav_prep_kernel + wav_conv_kernel<16,32,1> + <32,64,0> + <64,32,0> (umma_wav.cu: conv1 "
          "recomputed per tile, BatchNorm + LeakyReLU applied while the next strided convolution stages its tcgen05 "
          "operand, train-mode batch statistics)", "hbm", ms, 39.38e6 * B, 149420.0 * B)
    r = dict(dom)
    r["roofline_kernels"] = out
    if clips_per_s_per_gpu is not None:
        r["step_frac_of_tensor_roofline"] = clips_per_s_per_gpu * FLOP_PER_CLIP / 1e12 / pk["bf16_sustained"]
    return r

# ----------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--ref-batch", type=int, default=128,
                    help="clips per timed iteration of the reference CPU arm (BASELINE config 2 batch)")
    ap.add_argument("--batch-per-gpu", type=int, default=256)
    ap.add_argument("--cpu-sample-batch", type=int, default=16)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16x1"],
                    help="tensor-core operand precision: bf16x3 = fp32-grade (parity configuration, default), bf16x1 = BASELINE config 3")
    ap.add_argument("--roofline-only", action="store_true", help="run only the roofline kernel (for ncu captures)")
    ap.add_argument("--e2e-input", default="cache", choices=["cache", "fp32"],
                    help="host format of the end-to-end leg: 'cache' = the reference npz cache's own precision (int16 audio "
                         "+ per-clip scale, fp16 MFCC: processor_v2.py:231,606-610), expanded on the device; 'fp32' = "
                         "host-expanded fp32 tensors")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl != "ours":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from speech2affective_gestures_b200 import _C, ops
    from speech2affective_gestures_b200.processor_v2 import Processor
    from speech2affective_gestures_b200.synthetic import make_data_loader, synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _C.lib()
    assert lib.s2ag_is_device_build() == 1, "CUDA extension missing: there is no fallback"
    assert lib.s2ag_set_precision(0 if args.precision == "bf16x3" else 1) == 0
    if args.roofline_only:
        print(json.dumps({"roofline": roofline(args.batch_per_gpu, dev, lib)}))
        return

    from speech2affective_gestures_b200.config import namespace as config_namespace
    cfg = config_namespace()  # (the oracle is only imported by the CPU legs below)
    B = args.batch_per_gpu
    pargs = NS(no_cuda=False, work_dir_s2ag=None, save_log=False, print_log=False, train_s2ag=True, batch_size=B,
               s2ag_num_epoch=1, val_interval=1, save_interval=10)
    dl = make_data_loader(8, 8, 8, n_words=N_WORDS, n_speakers=N_SPEAKERS)
    torch.manual_seed(1234 + rank)
    ops.manual_seed(1234 + rank)
    pr = Processor(ROOT, pargs, cfg, dl, 27, 3, 16000)
    pr.meta_info["epoch"] = 1  # GAN branch active (loss_warmup = 0)
    for net in (pr.s2ag_generator, pr.s2ag_discriminator):
        net.train()
    pr.trimodal_generator.train()  # the reference never calls .eval() on it before train() (processor_v2.py:961-962)

    host = synthetic_batch(B, None, N_WORDS, N_SPEAKERS, AUDIO_LEN, seed=1234 + rank, pin=True)
    if args.e2e_input == "cache":
        # the loader's on-disk form (processor_v2.py:601-611): audio int16 with its per-clip scale, MFCC fp16
        amax = host[1].abs().amax(dim=1).clamp_min(1e-8)
        a16 = (host[1] * (32767.0 / amax[:, None])).round().clamp_(-32767, 32767).to(torch.int16).pin_memory()
        host_c = (host[0], a16, amax.float().pin_memory(), host[2].to(torch.float16).pin_memory(), host[3], host[4])
    else:
        host_c = host
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_c)
    host_metrics = torch.zeros(8, dtype=torch.float32).pin_memory()

    use_graph = not args.no_graph
    n0 = lib.s2ag_launch_count()
    if use_graph:
        try:
            pr.capture_step(B, train=True, warmup=2)
        except Exception as e:  # e.g. NCCL capture unsupported: run the step eagerly
            if rank == 0:
                print("graph capture failed (%s); running eagerly" % str(e)[:200], file=sys.stderr)
            use_graph = False
    if not use_graph:
        pr.static_in = tuple(t.to(dev) for t in host)
    n1 = lib.s2ag_launch_count()
    pr.load_static_inputs(*[host[i] for i in (0, 1, 2, 3, 4)]) if use_graph else None
    torch.cuda.synchronize()

    def one_step():
        if use_graph:
            pr.replay_step()
        else:
            pr.gan_step_async(*pr.static_in, True)

    c0 = lib.s2ag_launch_count()
    one_step()
    launches_per_step = (lib.s2ag_launch_count() - c0) if not use_graph else None
    if use_graph:  # kernels recorded into the graph by the capture pass (last of the warmup+capture passes)
        launches_per_step = (n1 - n0) // 3
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    # ---- device-resident timing
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # ---- end-to-end timing (pinned host inputs -> H2D -> step -> D2H metrics), public Processor API
    # Host-side jitter (shared box: other tenants on the host cores / PCIe) moves this number by up to 30 % from one pass
    # to the next while the device-timed value is stable to 1 %: the loop is run E2E_PASSES times, every pass is listed
    # in e2e.passes_ms_per_step and the fastest one is reported.
    e2e_passes = []
    for _pass in range(E2E_PASSES):
        barrier()
        t0 = time.perf_counter()
        if use_graph:
            # the loader's prefetch: the H2D copy of batch i+1 (copy stream, pinned memory) overlaps step i; every step's
            # inputs cross PCIe once inside the timed region, the first copy is exposed
            prefetch = pr.prefetch_inputs_compressed if args.e2e_input == "cache" else pr.prefetch_inputs
            prefetch(*host_c)
            for i in range(args.steps):
                pr.swap_in_prefetched()
                if i + 1 < args.steps:
                    prefetch(*host_c)
                one_step()
                host_metrics.copy_(pr.metrics, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        else:
            for _ in range(args.steps):
                pr.load_static_inputs(*host)
                one_step()
                host_metrics.copy_(pr.metrics, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        e2e_passes.append((time.perf_counter() - t0) * 1e3)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    if world > 1:   # every timing is the MAX over ranks (per pass for the end-to-end loop)
        tt = torch.tensor([ms] + e2e_passes, device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_passes = tt[0].item(), tt[1:].tolist()
    t_e2e = min(e2e_passes)
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (t_e2e / 1e3)

    # ---- roofline of the dominant kernel family, timed alone with CUDA events on its launch stream
    roof = roofline(B, dev, lib, value / world) if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_loader
        if ref_loader.reference_root() is not None:
            v, dt, root = reference_step_clips_per_s(args.ref_batch, 2, 1, threads, "cpu")
            cpu = {"value": v, "unit": "clips/s", "cores": threads, "kind": "reference",
                   "sample": "2 timed GAN iterations of %d clips after 1 warm-up (UNMODIFIED reference "
                             "Processor.forward_pass_s2ag from %s, dropout as shipped, %d torch threads)"
                             % (args.ref_batch, "oracle/_ref" if root.endswith("_ref") else "the reference checkout",
                                threads)}
        else:
            v, dt = cpu_reference_clips_per_s(args.cpu_sample_batch, 1, 1, threads)
            cpu = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                   "sample": "1 timed GAN iteration of %d clips after 1 warm-up (oracle port of "
                             "processor_v2.forward_pass_s2ag, dropout off, %d torch threads)"
                             % (args.cpu_sample_batch, threads)}

    if rank == 0:
        line = {
            "metric": "gesture-clips/sec (34-frame, 27-D pose), full GAN training step", "value": value,
            "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "bf16x3" else "bf16", "data": "synthetic",
            "config": {"workload": workload_text(B),
                       "global_batch": B * world, "parallelism": "dp%d" % world,
                       "precision": "%s (fp32 storage and accumulation; tensor-core operands %s)" % (
                           args.precision, "split bf16 hi+lo, 3 MMAs" if args.precision == "bf16x3" else "single bf16"),
                       "l2": "per-step working set (activations+saved gates+params, >1 GB) exceeds the 126 MB L2; "
                             "no explicit flush", "cuda_graph": use_graph},
            "e2e": {"value": e2e, "unit": "clips/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 32,
                    "ms_per_step": t_e2e / args.steps,
                    "passes_ms_per_step": [round(t / args.steps, 3) for t in e2e_passes],
                    "host_format": ("reference npz-cache precision (int16 audio + per-clip scale, fp16 MFCC), expanded on "
                                    "the device inside the timed region") if args.e2e_input == "cache" and use_graph
                    else "fp32 host tensors"},
            "gpu_launches": int(launches_per_step * (2 * args.steps)) if launches_per_step else 0,
            "gpu_launches_per_step": int(launches_per_step or 0),
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        # Tear-down: the captured graph holds NCCL work; destroying the communicator underneath it can block
        # (observed on 2 GPUs).  Drop the graph, drain the device, rendezvous once more and leave without the
        # communicator destructor.
        pr._graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
