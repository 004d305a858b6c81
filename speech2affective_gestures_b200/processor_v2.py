"""Drop-in for the hot-path entry points of the reference's processor_v2.py `Processor`:
constructor, `forward_pass_s2ag` (one GAN iteration, :776-957), `per_train_epoch` / `per_val_epoch` /
`train` (:959-1069), `yield_batch` (:589-638), `generate_gestures` (:1071-1142), `render_clip` (:1144-1439) and
`generate_gestures_by_dataset` (:1441-1567); the long-form entry points render clips in lock-step batches on the
device (longform.py).

The FGD evaluators (`EmbeddingSpaceEvaluator`, :162-165) are built when `outputs/embedding_net.pth.tar` exists under
base_path and accumulate on the device (net/embedding_space_evaluator.py, csrc/fgd.cu).

Out of scope: LMDB/npz cache writers, video rendering, BVH export.

Execution model (B200-first): one process per GPU; every kernel of a step is enqueued on one
stream with no host synchronisation (the reference's ~9 `.item()` syncs per step, :943-956, become
one 8-float device buffer read on demand); the whole step can be captured once into a CUDA graph
(`capture_step`) and replayed; gradients of a network are one flat buffer, all-reduced with a single
NCCL call per optimiser when torch.distributed is initialised (replaces nn.DataParallel, :167-172);
Adam runs as one kernel over the flat parameter buffer with device-side bias correction.
"""
import os
import re
import time

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .net import embedding_net as en
from .net.embedding_space_evaluator import EmbeddingSpaceEvaluator
from .net.multimodal_context_net_v2 import PoseGeneratorTriModal as PGT, ConvDiscriminatorTriModal as CDT, \
    PoseGenerator, AffDiscriminator

torch.manual_seed(1234)  # processor_v2.py:37
ops.manual_seed(1234)

# slots of Processor.metrics (device float32[8])
M_DIS, M_HUBER, M_GEN, M_KLD, M_DIV, M_TOTAL, M_L1, M_L1_TRI = range(8)


def _prio(role):
    """Stream priorities of the captured step (CUDA: lower number = scheduled first when SMs free up).  The main stream
    carries the serial D step / D(out) / loss chain of short kernels; the side streams carry wide, latency-tolerant
    work (weight-gradient GEMMs, the other generator passes) that must not starve it.  S2AG_STREAM_PRIO="m,a,b[,c]"
    overrides (A/B measurements: raising any side stream to the main stream's priority costs 0.05-0.6 ms/step)."""
    env = os.environ.get("S2AG_STREAM_PRIO")   # "main,side,sideb[,sidec]"
    table = dict(main=-1, side=0, sideb=0)   # measured: 12.94 -> 12.50 ms/step (tools/ab_schedule.sh)
    if env:
        table.update(zip(("main", "side", "sideb", "sidec"), (int(v) for v in env.split(","))))
    table.setdefault("sidec", table["side"])
    return table[role]


_STREAMS = {}


def _stream_for(device, role):
    """The step's streams are a per-device resource shared by every Processor of the process (each owns a 192 MB
    packed-operand scratch registered with the library, ops._handle): created once per (device, role)."""
    key = (torch.device(device).index, role)
    if key not in _STREAMS:
        _STREAMS[key] = torch.cuda.Stream(device=device, priority=_prio(role if role in ("main", "side", "sideb", "sidec") else "side"))
    return _STREAMS[key]


def get_epoch_and_loss(path_to_model_files, epoch='best'):
    """Checkpoint discovery by filename, same scheme as processor_v2.py:53-83:
    epoch_{:06d}_loss_{:.4f}_model.pth.tar; 'best' = lowest loss."""
    if not path_to_model_files or not os.path.isdir(path_to_model_files):
        return None, None, np.inf
    found = []
    for f in os.listdir(path_to_model_files):
        # the saved metric is L1(ours) - L1(tri-modal): negative exactly when the model beats the baseline
        m = re.match(r"epoch_(\d+)_loss_(-?(?:[0-9.]+(?:e[-+]?\d+)?|inf|nan))_model\.pth\.tar$", f)
        if m:
            found.append((int(m.group(1)), float(m.group(2)), f))
    if not found:
        return None, None, np.inf
    if epoch == 'best':
        e, l, f = min(found, key=lambda r: r[1])
    else:
        cand = [r for r in found if r[0] == int(epoch)]
        if not cand:
            return None, None, np.inf
        e, l, f = cand[0]
    return f, e, l


class _Log:
    def __init__(self, work_dir, save_log=True, print_log=True):
        self.work_dir, self.save_log, self.do_print = work_dir, save_log, print_log

    def print_log(self, s):
        if self.do_print:
            print(s)
        if self.save_log and self.work_dir:
            os.makedirs(self.work_dir, exist_ok=True)
            with open(os.path.join(self.work_dir, "log.txt"), "a") as f:
                f.write(s + "\n")


class Processor(object):
    """Processor for emotive gesture generation (B200-native hot path)."""

    def __init__(self, base_path, args, s2ag_config_args, data_loader, pose_dim, coords, audio_sr,
                 min_train_epochs=20, zfill=6):
        self.base_path = base_path
        self.args = args
        if getattr(args, "no_cuda", False) or not torch.cuda.is_available():
            from . import _C
            if not _C.is_emulated():
                raise _C.S2agError("this Processor runs on a CUDA device only (no CPU fallback)")
            self.device = torch.device("cpu")  # tests/emu: kernel-logic emulator injected explicitly
        else:
            self.device = torch.device('cuda:{}'.format(torch.cuda.current_device()))
        self.s2ag_config_args = s2ag_config_args
        self.data_loader = data_loader
        self.result, self.iter_info, self.epoch_info = dict(), dict(), dict()
        self.meta_info = dict(epoch=0, iter=0)
        self.io = _Log(getattr(args, "work_dir_s2ag", None), getattr(args, "save_log", False),
                       getattr(args, "print_log", True))
        self.pose_dim, self.coords, self.audio_sr = pose_dim, coords, audio_sr

        part = 'train_data_s2ag' if getattr(args, "train_s2ag", True) else 'test_data_s2ag'
        d = self.data_loader[part]
        self.time_steps = d.n_poses
        self.audio_length = d.expected_audio_length
        self.num_mfcc = d.num_mfcc_combined
        self.lang_model = d.lang_model
        self.mfcc_length = int(np.ceil(self.audio_length / 512))
        self.best_s2ag_loss, self.best_s2ag_loss_epoch, self.s2ag_loss_updated = np.inf, None, False
        self.min_train_epochs, self.zfill = min_train_epochs, zfill
        self.train_speaker_model = self.data_loader['train_data_s2ag'].speaker_model
        self.val_speaker_model = self.data_loader['val_data_s2ag'].speaker_model
        self.test_speaker_model = self.data_loader['test_data_s2ag'].speaker_model

        cfg = self.s2ag_config_args
        self.trimodal_generator = PGT(cfg, pose_dim=pose_dim, n_words=self.lang_model.n_words,
                                      word_embed_size=cfg.wordembed_dim,
                                      word_embeddings=self.lang_model.word_embedding_weights,
                                      z_obj=self.train_speaker_model)
        self.trimodal_discriminator = CDT(pose_dim)  # constructed, never called (processor_v2.py:141)
        self.use_mfcc = True
        self.s2ag_generator = PoseGenerator(cfg, pose_dim=pose_dim, n_words=self.lang_model.n_words,
                                            word_embed_size=cfg.wordembed_dim,
                                            word_embeddings=self.lang_model.word_embedding_weights,
                                            mfcc_length=self.mfcc_length, num_mfcc=self.num_mfcc,
                                            time_steps=self.time_steps, z_obj=self.train_speaker_model)
        self.s2ag_discriminator = AffDiscriminator(pose_dim)
        for net in (self.trimodal_generator, self.trimodal_discriminator, self.s2ag_generator,
                    self.s2ag_discriminator):
            net.to(self.device)

        # processor_v2.py:162-165; the reference fails without the checkpoint, here the evaluators are simply absent
        self.evaluator_trimodal = self.evaluator = None
        if base_path and os.path.isfile(os.path.join(base_path, 'outputs/embedding_net.pth.tar')):
            self.evaluator_trimodal = EmbeddingSpaceEvaluator(base_path, cfg, pose_dim, self.lang_model, self.device)
            self.evaluator = EmbeddingSpaceEvaluator(base_path, cfg, pose_dim, self.lang_model, self.device)

        self.train_samples = self.data_loader['train_data_s2ag'].samples
        self.val_samples = self.data_loader['val_data_s2ag'].samples
        self.num_train_samples = self.data_loader['train_data_s2ag'].n_samples
        self.num_val_samples = self.data_loader['val_data_s2ag'].n_samples
        self.num_test_samples = self.data_loader['test_data_s2ag'].n_samples

        self.lr_s2ag_gen = cfg.learning_rate
        self.lr_s2ag_dis = cfg.learning_rate * cfg.discriminator_lr_weight
        self._init_optimizers()
        self._init_distributed()
        self.metrics = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._graph = None
        self._side_stream = None
        self._side_stream_c = None
        self._side_stream_d = None
        self._side_stream_b = None
        self._copy_stream = None
        self.use_side_stream = True
        self.injected_rand_idx = None  # parity harness: fixed speaker permutation for processor_v2.py:903

    # ------------------------------------------------------------------ optimiser / distributed state
    def _init_optimizers(self):
        """Adam(betas=(0.5, 0.999)) state for G and D over the flat buffers (processor_v2.py:215-220)."""
        G, D = self.s2ag_generator, self.s2ag_discriminator
        self.gen_m, self.gen_v = torch.zeros_like(G.flat_params), torch.zeros_like(G.flat_params)
        self.dis_m, self.dis_v = torch.zeros_like(D.flat_params), torch.zeros_like(D.flat_params)
        self.gen_step = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.dis_step = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _init_distributed(self):
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if self.world > 1:  # one initial broadcast of weights and BN buffers
            for net in (self.trimodal_generator, self.s2ag_generator, self.s2ag_discriminator):
                dist.broadcast(net.flat_params, 0)
                for b in net.buffers():
                    if b.dtype == torch.float32:
                        dist.broadcast(b, 0)

    def _allreduce_grads(self, net):
        """SUM over ranks; the 1/world average is folded into the Adam kernel's grad_scale."""
        if self.world > 1:
            dist.all_reduce(net.flat_grads)

    # ------------------------------------------------------------------ one GAN iteration
    def make_pre_seq(self, target_poses):
        n_pre = self.s2ag_config_args.n_pre_poses
        pre_seq = target_poses.new_zeros((target_poses.shape[0], target_poses.shape[1], target_poses.shape[2] + 1))
        pre_seq[:, 0:n_pre, :-1] = target_poses[:, 0:n_pre]
        pre_seq[:, 0:n_pre, -1] = 1  # indicating bit for constraints
        return pre_seq

    def gan_step_async(self, in_text, in_audio, in_mfcc, target_poses, vid_indices, train):
        """processor_v2.py:776-957 without host synchronisation; results land in self.metrics and
        self.last_out (generated poses of the main G pass) / self.last_out_trimodal."""
        cfg = self.s2ag_config_args
        G, D, Tri = self.s2ag_generator, self.s2ag_discriminator, self.trimodal_generator
        m = self.metrics
        gan_on = self.meta_info['epoch'] > cfg.loss_warmup and cfg.loss_gan_weight > 0.0
        pre_seq = self.make_pre_seq(target_poses)
        if train:
            ops.advance_seed_nonce(self.device)
        use_side = self.device.type == "cuda" and self.use_side_stream
        ev_start = None
        if use_side:
            # fork point of the side streams: the inputs are ready and the dropout nonce is advanced; nothing the
            # main stream launches from here on (the shared encoders first) holds the side streams back
            ev_start = torch.cuda.Event(); ev_start.record(torch.cuda.current_stream())
            if self._side_stream_c is None and os.environ.get("S2AG_CONV_WGRAD_STREAM", "1") != "0":
                self._side_stream_c = _stream_for(self.device, "sidec")
        feat_real = ev_real = None
        if use_side and gan_on and train and os.environ.get("S2AG_D_REAL_EARLY", "0") == "1":
            # (opt-in, measured SLOWER: 12.38 vs 11.92 ms/step -- two 256-clip encoder chains cost more than one 512-clip
            # chain on SMs shared with the persistent kernels.)
            # D(target)'s AffEncoder (:808) depends on nothing the generator produces: it runs on a stream of its own
            # from the step's start, beside generator pass #1; autograd runs its backward there too, beside the backward
            # of D(fake)'s encoder (parameter-gradient accumulations are serialised on the conv weight-gradient stream
            # or atomic, see ops.ConvBnActFn / csrc/bn.cu)
            sd_ = self._side_stream_d = _stream_for(self.device, "sided")
            sd_.wait_event(ev_start)
            with torch.cuda.stream(sd_):
                feat_real = D.aff_encoder(target_poses)
            ev_real = torch.cuda.Event(); ev_real.record(sd_)
            target_poses.record_stream(sd_)
            feat_real.record_stream(torch.cuda.current_stream())
        use_div = cfg.z_type in ('speaker', 'random') and cfg.loss_reg_weight > 0.0
        # AffEncoder(pre_seq) and MFCCEncoder(in_mfcc) have no dropout and G's weights do not change between the
        # generator passes of one iteration (:798, :823, :909): evaluate them once for all passes.
        n_passes = (1 if gan_on else 0) + 1 + (1 if use_div else 0)
        with torch.set_grad_enabled(train):
            shared = G.encode_shared(pre_seq, in_mfcc, repeats=n_passes if G.training else 1,
                                     mfcc_stream=self._side_stream_c if (use_side and not ops.TCN_FUSED[0] and os.environ.get(
                                         "S2AG_MFCC_STREAM", "1") != "0") else None)
        shared_ng = tuple(None if t is None else t.detach() for t in shared)
        # Input-only encoders (the generator's TextEncoderTCN for each of its passes, the frozen baseline's WavEncoder
        # and TextEncoderTCN) run on a side stream: they overlap the latency-bound recurrent kernels of the main
        # stream, which occupy only about half of the SMs.  Events order each consumer after its producer; autograd
        # runs the backward of the side-stream ops on the side stream and synchronises by itself.
        main_s = torch.cuda.current_stream() if use_side else None
        txt1 = txt2 = txt3 = tri_pre = None
        eps2 = eps3 = early3 = early2 = run_tri_late = out_tri = None
        ev = {}
        if use_side:
            if self._side_stream is None:
                self._side_stream = _stream_for(self.device, "side")
            side = self._side_stream
            if self._side_stream_c is None and os.environ.get("S2AG_CONV_WGRAD_STREAM", "1") != "0":
                self._side_stream_c = _stream_for(self.device, "sidec")
            # GRU weight-gradient GEMMs run on `side`, beside the next layer's BPTT kernel; the convolutions' weight
            # gradients on a stream of their own (they would otherwise queue behind the text encoder's backward)
            if self._side_stream_b is None:
                self._side_stream_b = _stream_for(self.device, "sideb")
            # (the TCN blocks' weight gradients go to the second side stream: it is idle once the generator's BPTT is done)
            ops.set_side_stream(side, self._side_stream_c,
                                self._side_stream_b if os.environ.get("S2AG_TCN_WGRAD_STREAM", "b") == "b" else None)
            # Known issue: with the opt-in single-kernel TCN block (S2AG_TCN_FUSED=1) the early fork ends in a launch failure
            # when that kernel runs beside the shared encoders at the step's start (memcheck-clean, kernel tests green,
            # fine with the late fork: 12.04 ms/step) -- unresolved, so the opt-in keeps the late fork.
            if os.environ.get("S2AG_EARLY_FORK", "1") != "0" and not ops.TCN_FUSED[0]:
                side.wait_event(ev_start)
            else:
                side.wait_stream(main_s)
            if self._side_stream_c is not None:
                self._side_stream_c.wait_stream(main_s)
            with torch.cuda.stream(side):
                if gan_on:
                    with torch.no_grad():
                        txt1 = G.encode_text(in_text)
                    ev[1] = torch.cuda.Event(); ev[1].record(side)
                with torch.no_grad():
                    tri_pre = Tri.encode_inputs(in_text, in_audio)
                ev['t'] = torch.cuda.Event(); ev['t'].record(side)
                with torch.set_grad_enabled(train):
                    txt2 = G.encode_text(in_text)
                ev[2] = torch.cuda.Event(); ev[2].record(side)
                if use_div:
                    with torch.no_grad():
                        txt3 = G.encode_text(in_text)
                    ev[3] = torch.cuda.Event(); ev[3].record(side)
            for t_ in (txt1, txt2, txt3) + (tuple(tri_pre) if tri_pre is not None else ()):
                if t_ is not None:
                    t_.record_stream(main_s)

        # ---- train D (processor_v2.py:791-814)
        if gan_on:
            if train:
                D.zero_grad()
            if use_side:
                main_s.wait_event(ev[1])
            with torch.no_grad():  # the reference builds and discards this graph; only .detach() is used (:809)
                out_for_d, *_ = G(pre_seq, in_text, in_mfcc, vid_indices, shared=shared_ng, text_feat=txt1)
            if use_side:
                # The frozen baseline's recurrent body runs on a second side stream beside the D step.  Its persistent GRU
                # kernels (76 resident CTAs, one SM each) may only overlap the discriminator-sized persistent kernels of
                # the D step (<= 32 small CTAs): it starts after generator pass #1 (event below) and the main stream waits
                # for it before generator pass #2, so two generator-sized persistent kernels (76 + 76 > 148 SMs) can
                # never be in flight together.
                if self._side_stream_b is None:
                    self._side_stream_b = _stream_for(self.device, "sideb")
                sb = self._side_stream_b
                ev_p1 = torch.cuda.Event(); ev_p1.record(main_s)
                sb.wait_event(ev_p1)
                sb.wait_event(ev['t'])
                # (the baseline's re-parametrisation noise is drawn here, at its place in the reference's order)
                # In training every consumer of these draws lives on the second side stream, so the draws (and the sort
                # behind the speaker permutation, ~45 us of small kernels) are enqueued there and leave the main
                # stream's serial chain; the generator state advances in program order either way.
                draw_s = sb if train else main_s
                eps_t = None
                if getattr(Tri, 'speaker_embedding', None) is not None:
                    with torch.cuda.stream(draw_s):
                        eps_t = en.draw_eps(torch.empty(vid_indices.shape[0], Tri.z_size, device=self.device))
                    ev_e = torch.cuda.Event(); ev_e.record(draw_s)
                    sb.wait_event(ev_e)

                def run_tri():
                    with torch.cuda.stream(sb):
                        with torch.no_grad():
                            o, *_ = Tri(pre_seq, in_text, in_audio, vid_indices, pre=tri_pre, eps=eps_t)
                    o.record_stream(main_s)
                    for t_ in tuple(tri_pre) + (eps_t, in_text, in_audio, vid_indices):
                        if t_ is not None:
                            t_.record_stream(sb)
                    return o

                # training (measured order, tools/ab_schedule.sh / ab_final.sh): generator pass #2, pass #3, then the frozen baseline
                # ("mid") on this stream beside the D step; "1" queues the baseline behind the generator's BPTT instead
                tri_mode = os.environ.get("S2AG_TRI_LATE", "mid") if train else "0"   # A/B: "0" first, "1" after the BPTT
                tri_late = tri_mode != "0"
                if not tri_late:
                    out_tri = run_tri()
                    ev['tdone'] = torch.cuda.Event(); ev['tdone'].record(sb)
                pre_seq.record_stream(sb)
                if use_div and train and cfg.z_type == 'speaker':
                    # the re-parametrisation noise of passes #2 and #3 and the speaker permutation are drawn here, in
                    # the reference's order (:823, :903-909), whatever order the passes are launched in
                    with torch.cuda.stream(draw_s):
                        like = torch.empty(vid_indices.shape[0], G.z_size, device=self.device)
                        eps2, eps3 = en.draw_eps(like), en.draw_eps(like)
                        rand_idx = self.injected_rand_idx if self.injected_rand_idx is not None else \
                            torch.rand(vid_indices.shape[0], device=vid_indices.device).argsort()
                        rand_vids = vid_indices[rand_idx]
                    vid_indices.record_stream(draw_s)
                else:
                    rand_vids = None
                def launch_pass2():
                    # Generator pass #2 (the forward of the G step, :823) does not depend on the D update (only D(out)
                    # after it does).  autograd runs its backward (BPTT) on this stream and orders it against the rest.
                    nonlocal early2
                    sb.wait_event(ev[2])
                    ev_m = torch.cuda.Event(); ev_m.record(main_s)
                    sb.wait_event(ev_m)
                    with torch.cuda.stream(sb):
                        early2 = G(pre_seq, in_text, in_mfcc, vid_indices, shared=shared, text_feat=txt2, eps=eps2)
                    ev['p2done'] = torch.cuda.Event(); ev['p2done'].record(sb)
                    for t_ in tuple(shared) + (txt2, vid_indices, eps2, in_text, in_mfcc):
                        if t_ is not None:
                            t_.record_stream(sb)
                    for t_ in early2:
                        if t_ is not None:
                            t_.record_stream(main_s)

                def launch_pass3():
                    # Generator pass #3 (the no-grad style-diversity pass, :903-910) depends on nothing the D step or
                    # pass #2 produce and is only needed by the loss.
                    nonlocal early3
                    ev_r = torch.cuda.Event(); ev_r.record(main_s)
                    sb.wait_event(ev_r)
                    sb.wait_event(ev[3])
                    with torch.cuda.stream(sb):
                        with torch.no_grad():
                            early3 = G(pre_seq, in_text, in_mfcc, rand_vids, shared=shared_ng, text_feat=txt3, eps=eps3)
                    ev['p3done'] = torch.cuda.Event(); ev['p3done'].record(sb)
                    for t_ in shared_ng + (txt3, rand_vids, eps3, in_text, in_mfcc):
                        if t_ is not None:
                            t_.record_stream(sb)
                    for t_ in early3:
                        if t_ is not None:
                            t_.record_stream(main_s)

                # order on the side stream (measured, tools/ab_final.sh: 11.76 vs 11.87 ms/step): pass #2, pass #3, then the
                # baseline -- D(out) on the main stream can start as soon as the D step is done
                order = os.environ.get("S2AG_PASS_ORDER", "23")
                for which in order:
                    if which == "2" and train:
                        launch_pass2()
                    elif which == "3" and use_div and train:
                        launch_pass3()
                if tri_mode == "mid":
                    out_tri = run_tri()
                elif tri_late:
                    run_tri_late = run_tri
            if ev_real is not None:
                main_s.wait_event(ev_real)
            with torch.set_grad_enabled(train):
                # == D(target), D(out.detach()) (:808-809)
                dis_real, dis_fake = D.forward_pair(target_poses, out_for_d, feat_a=feat_real)
            g_real, g_fake = ops.dis_loss(dis_real, dis_fake, m[M_DIS:M_DIS + 1], want_grads=train)
            if train:
                torch.autograd.backward([dis_real, dis_fake], [g_real, g_fake])
                if use_side:
                    main_s.wait_stream(self._side_stream)  # weight-gradient GEMMs issued on the side streams
                    if ev_real is not None:
                        main_s.wait_stream(self._side_stream_d)   # backward of D(target)'s encoder
                    if self._side_stream_c is not None:
                        main_s.wait_stream(self._side_stream_c)
                self._allreduce_grads(D)
                ops.adam_step(D.flat_params, D.flat_grads, self.dis_m, self.dis_v, self.lr_s2ag_dis, 0.5, 0.999,
                              1e-8, self.dis_step, 1.0 / self.world)

        # ---- train G (processor_v2.py:816-941)
        if train:
            G.zero_grad()
        if use_side and 'tdone' in ev:
            main_s.wait_event(ev['tdone'])
        elif run_tri_late is None and out_tri is None:
            if use_side:
                main_s.wait_event(ev['t'])
            with torch.no_grad():
                out_tri, *_ = Tri(pre_seq, in_text, in_audio, vid_indices, pre=tri_pre)
        if use_side:
            main_s.wait_event(ev[2])
            if 'p2done' in ev:
                main_s.wait_event(ev['p2done'])
        with torch.set_grad_enabled(train):
            if early2 is not None:
                out, z, z_mu, z_log_var = early2
            else:
                out, z, z_mu, z_log_var = G(pre_seq, in_text, in_mfcc, vid_indices, shared=shared, text_feat=txt2,
                                            eps=eps2)
            # D's own parameter gradients from this pass are discarded by the reference (zero_grad at the
            # next D step, :794), so they are not computed; gradients still flow through D into G.
            d_params = [p for p in D.parameters() if p.requires_grad]
            for p in d_params:
                p.requires_grad_(False)
            try:
                dis_out = D(out, in_text)
            finally:
                for p in d_params:
                    p.requires_grad_(True)
        out_rand = z_rand = None
        if early3 is not None:
            main_s.wait_event(ev['p3done'])
            out_rand, z_rand = early3[0], early3[1]
        elif use_div:
            if cfg.z_type == 'speaker':
                # torch.randperm(B) of processor_v2.py:903; drawn as argsort(uniform) on the device so that
                # the draw is capturable in a CUDA graph (CUDA randperm synchronises the host)
                rand_idx = self.injected_rand_idx if self.injected_rand_idx is not None else \
                    torch.rand(vid_indices.shape[0], device=vid_indices.device).argsort()
                rand_vids = vid_indices[rand_idx]
            else:
                rand_vids = None
            if use_side:
                main_s.wait_event(ev[3])
            with torch.no_grad():  # only used detached (:913, :919)
                out_rand, z_rand, _, _ = G(pre_seq, in_text, in_mfcc, rand_vids, shared=shared_ng, text_feat=txt3)
        use_kld = use_div and cfg.z_type == 'speaker'
        weights = (cfg.loss_regression_weight, cfg.loss_kld_weight if use_kld else 0.0,
                   cfg.loss_reg_weight if use_div else 0.0, cfg.loss_gan_weight if gan_on else 0.0)
        g_out, g_dis, g_mu, g_lv = ops.gen_loss(
            out, target_poses, out_rand, z, z_rand, z_mu if use_kld else (z if use_div else None),
            z_log_var if use_kld else (z if use_div else None), dis_out, weights,
            m[M_HUBER:M_HUBER + 5], want_grads=train)
        if train:
            outs, grads = [out], [g_out]
            if gan_on:
                outs.append(dis_out); grads.append(g_dis)
            if use_kld:
                outs += [z_mu, z_log_var]; grads += [g_mu, g_lv]
            torch.autograd.backward(outs, grads)
            if use_side:
                # the text-encoder backward (and its in-kernel parameter-gradient accumulation) ran on the side stream,
                # the generator's own backward on the second one
                main_s.wait_stream(self._side_stream)
                if self._side_stream_c is not None:
                    main_s.wait_stream(self._side_stream_c)
                if self._side_stream_b is not None:
                    main_s.wait_stream(self._side_stream_b)
                if run_tri_late is not None:
                    # the frozen baseline, queued behind the generator's BPTT kernels on the second side stream (two
                    # generator-sized persistent kernels never share the SMs): it runs beside the all-reduce / Adam
                    # tail and is joined below; only the final metric reads its output
                    out_tri = run_tri_late()
                    run_tri_late = None
            self._allreduce_grads(G)
            ops.adam_step(G.flat_params, G.flat_grads, self.gen_m, self.gen_v, self.lr_s2ag_gen, 0.5, 0.999, 1e-8,
                          self.gen_step, 1.0 / self.world)
        if run_tri_late is not None:   # (no backward pass ran: not reachable in training, kept for safety)
            out_tri = run_tri_late()
        if use_side:
            main_s.wait_stream(self._side_stream)  # join (also required before a graph capture ends)
            if ev_real is not None:
                main_s.wait_stream(self._side_stream_d)
            if self._side_stream_c is not None:
                main_s.wait_stream(self._side_stream_c)
            if self._side_stream_b is not None:
                main_s.wait_stream(self._side_stream_b)
            ops.set_side_stream(None)
        ops.l1_mean(out.detach(), target_poses, m[M_L1:M_L1 + 1])
        ops.l1_mean(out_tri, target_poses, m[M_L1_TRI:M_L1_TRI + 1])
        self.last_out, self.last_out_trimodal = out.detach(), out_tri
        return m

    def forward_pass_s2ag(self, in_text, in_audio, in_mfcc, target_poses, vid_indices, train,
                          target_seq=None, words=None, aux_info=None, save_path=None, make_video=False,
                          calculate_metrics=False, losses_all_trimodal=None, joint_mae_trimodal=None,
                          accel_trimodal=None, losses_all=None, joint_mae=None, accel=None):
        """Same signature and return tuple as processor_v2.py:776-779, :956-957."""
        if make_video:
            raise NotImplementedError("video rendering is outside the hot path (SURVEY 8)")
        m = self.gan_step_async(in_text, in_audio, in_mfcc, target_poses, vid_indices, train)
        host = m.tolist()  # the single device->host read of the step
        self.loss_dict = {'loss': self.s2ag_config_args.loss_regression_weight * host[M_HUBER],
                          'KLD': self.s2ag_config_args.loss_kld_weight * host[M_KLD],
                          'DIV_REG': self.s2ag_config_args.loss_reg_weight * host[M_DIV],
                          'gen': self.s2ag_config_args.loss_gan_weight * host[M_GEN], 'dis': host[M_DIS]}
        return host[M_L1] - host[M_L1_TRI], losses_all_trimodal, joint_mae_trimodal, accel_trimodal, losses_all, \
            joint_mae, accel

    # ------------------------------------------------------------------ CUDA-graph step
    def capture_step(self, batch_size, train=True, warmup=3):
        """Capture one whole GAN iteration (all forward/backward/Adam kernels, the NCCL all-reduces
        when distributed, RNG draws) into a CUDA graph over static input buffers."""
        assert self.device.type == "cuda"
        T, P = self.time_steps, self.pose_dim
        dev = self.device
        self.static_in = (torch.zeros(batch_size, T, dtype=torch.int64, device=dev),
                          torch.zeros(batch_size, self.audio_length, device=dev),
                          torch.zeros(batch_size, self.num_mfcc, self.mfcc_length, device=dev),
                          torch.zeros(batch_size, T, P, device=dev),
                          torch.zeros(batch_size, dtype=torch.int64, device=dev))
        self._graph_train = train
        # The capture stream is OUR stream: the warm-up iterations run on it first, so the library scratch buffer of
        # the packed-operand contractions is registered for it before the capture starts (ops._handle never allocates
        # during a capture; a stream first seen while capturing would fall back to the unpacked contraction).
        if getattr(self, "_capture_stream", None) is None:
            self._capture_stream = _stream_for(self.device, "main")
        cs = self._capture_stream
        snap = self._snapshot_state() if train else None   # warm-up iterations must not train the live model
        cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            for _ in range(max(1, warmup)):
                self.gan_step_async(*self.static_in, train)
        torch.cuda.current_stream().wait_stream(cs)
        torch.cuda.synchronize()
        if snap is not None:
            self._restore_state(snap)
        for st in (cs, self._side_stream, self._side_stream_b, self._side_stream_c, getattr(self, "_side_stream_d", None)):
            if st is not None:
                ops._handle(st, dev)   # registers the packed-operand scratch of a stream the warm-up did not launch on
            assert st is None or ops.has_scratch(st, dev), "a stream of the captured step has no registered scratch"
        import gc
        gc.collect()  # no autograd graph of the warm-up passes may survive into the capture
        self._graph = torch.cuda.CUDAGraph()
        bn_mods = [m for n in (self.s2ag_generator, self.s2ag_discriminator, self.trimodal_generator)
                   for m in n.modules() if hasattr(m, "_s2ag_batches")]
        before = [m._s2ag_batches for m in bn_mods]
        with torch.cuda.graph(self._graph, stream=cs):
            self.gan_step_async(*self.static_in, train)
        # BatchNorm's num_batches_tracked is a host-side counter on this path: remember what one iteration adds so
        # that replay_step() advances it; the capture pass itself did not execute
        self._bn_replay_delta = [(m, m._s2ag_batches - b) for m, b in zip(bn_mods, before) if m._s2ag_batches != b]
        for m, b in zip(bn_mods, before):
            m._s2ag_batches = b
        return self._graph

    def _snapshot_state(self):
        """Everything a training iteration mutates: parameters, Adam moments/step counters, BatchNorm buffers and
        their host-side batch counters, the dropout nonce."""
        nets = (self.s2ag_generator, self.s2ag_discriminator, self.trimodal_generator)
        return {
            "flat": [n.flat_params.clone() for n in nets],
            "bufs": [[b.clone() for b in n.buffers()] for n in nets],
            "bnc": [[getattr(m, "_s2ag_batches", 0) for m in n.modules()] for n in nets],
            "adam": [t.clone() for t in (self.gen_m, self.gen_v, self.dis_m, self.dis_v, self.gen_step, self.dis_step)],
            "nonce": ops.seed_nonce(self.device).clone(), "seed": ops.seed_state(),
            "metrics": self.metrics.clone(),
        }

    def _restore_state(self, snap, host_only=False):
        nets = (self.s2ag_generator, self.s2ag_discriminator, self.trimodal_generator)
        for n, c in zip(nets, snap["bnc"]):
            for m, v in zip(n.modules(), c):
                if hasattr(m, "_s2ag_batches"):
                    m._s2ag_batches = v
        if host_only:
            return
        for n, f, bs in zip(nets, snap["flat"], snap["bufs"]):
            n.flat_params.copy_(f)
            n.flat_grads.zero_()
            for b, v in zip(n.buffers(), bs):
                b.copy_(v)
        for t, v in zip((self.gen_m, self.gen_v, self.dis_m, self.dis_v, self.gen_step, self.dis_step), snap["adam"]):
            t.copy_(v)
        ops.seed_nonce(self.device).copy_(snap["nonce"])
        self.metrics.copy_(snap["metrics"])

    def load_static_inputs(self, in_text, in_audio, in_mfcc, target_poses, vid_indices):
        for dst, src in zip(self.static_in, (in_text, in_audio, in_mfcc, target_poses, vid_indices)):
            dst.copy_(src, non_blocking=True)

    def prefetch_inputs(self, in_text, in_audio, in_mfcc, target_poses, vid_indices):
        """Start the host->device copy of the NEXT batch (pinned host tensors) on a copy stream into staging buffers:
        it overlaps the step that is running.  `swap_in_prefetched()` moves it into the graph's static inputs (a
        device-to-device copy, ~30 us for 40 MB) at the start of the next step."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._staging_free = None
        cs = self._copy_stream
        if getattr(self, "_staging", None) is None:
            self._staging = tuple(torch.empty_like(t) for t in self.static_in)
        if self._staging_free is not None:
            cs.wait_event(self._staging_free)   # the previous swap has consumed the staging buffers
        with torch.cuda.stream(cs):
            for dst, src in zip(self._staging, (in_text, in_audio, in_mfcc, target_poses, vid_indices)):
                dst.copy_(src, non_blocking=True)
            self._staging_ready = torch.cuda.Event()
            self._staging_ready.record(cs)
        self._staging_kind = "fp32"

    def swap_in_prefetched(self):
        main = torch.cuda.current_stream()
        main.wait_event(self._staging_ready)
        if getattr(self, "_staging_c", None) is not None and self._staging_kind == "compressed":
            text, a16, amax, m16, tgt, vid = self._staging_c
            self.static_in[0].copy_(text, non_blocking=True)
            ops.expand_inputs(a16, amax, m16, audio_out=self.static_in[1], mfcc_out=self.static_in[2])
            self.static_in[3].copy_(tgt, non_blocking=True)
            self.static_in[4].copy_(vid, non_blocking=True)
        else:
            for dst, src in zip(self.static_in, self._staging):
                dst.copy_(src, non_blocking=True)
        self._staging_free = torch.cuda.Event()
        self._staging_free.record(main)

    def prefetch_inputs_compressed(self, in_text, audio_i16, audio_max, mfcc_f16, target_poses, vid_indices):
        """Like prefetch_inputs, but the batch crosses PCIe in the npz cache's own precision (int16 audio + fp32/fp64
        per-clip scale, fp16 MFCC: processor_v2.py:231, :606-610); swap_in_prefetched() expands it on the device into
        the graph's static inputs (csrc/frontend.cu).  20.6 instead of 40.8 MB per 256-clip step."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._staging_free = None
        if getattr(self, "_staging_c", None) is None:
            dev = self.device
            self._staging_c = (torch.empty_like(self.static_in[0]),
                               torch.empty(self.static_in[1].shape, dtype=torch.int16, device=dev),
                               torch.empty(self.static_in[1].shape[0], dtype=audio_max.dtype, device=dev),
                               torch.empty(self.static_in[2].shape, dtype=torch.float16, device=dev),
                               torch.empty_like(self.static_in[3]), torch.empty_like(self.static_in[4]))
        cs = self._copy_stream
        if self._staging_free is not None:
            cs.wait_event(self._staging_free)
        with torch.cuda.stream(cs):
            for dst, src in zip(self._staging_c, (in_text, audio_i16, audio_max, mfcc_f16, target_poses, vid_indices)):
                dst.copy_(src, non_blocking=True)
            self._staging_ready = torch.cuda.Event()
            self._staging_ready.record(cs)
        self._staging_kind = "compressed"

    def replay_step(self):
        self._graph.replay()
        for m, d in self._bn_replay_delta:
            m._s2ag_batches += d
        return self.metrics

    # ------------------------------------------------------------------ data / epochs
    def _gather(self, samples, keys):
        """npz-cache rows -> device tensors (processor_v2.py:601-611).  The cache holds int16 audio + its per-clip
        scale and fp16 MFCCs; they cross PCIe in that form (2 bytes per value) and `ops.expand_inputs` applies the
        reference's `audio * audio_max / 32767` and the fp16 -> fp32 conversion on the device, bit-identically."""
        dev = self.device
        text = torch.from_numpy(samples['extended_word_seq'][keys]).to(dev)
        vec = torch.from_numpy(samples['vec_seq'][keys]).float().to(dev)
        a16 = np.ascontiguousarray(samples['audio'][keys])
        amax = np.ascontiguousarray(samples['audio_max'][keys])
        m16 = np.ascontiguousarray(samples['mfcc_features'][keys])
        if a16.dtype == np.int16 and m16.dtype == np.float16 and amax.dtype in (np.float32, np.float64):
            audio, mfcc = ops.expand_inputs(torch.from_numpy(a16).to(dev), torch.from_numpy(amax).to(dev),
                                            torch.from_numpy(m16).to(dev))
        else:  # a cache written in another precision: same arithmetic on the host as the reference
            audio = torch.from_numpy(a16 * amax[:, None] / 32767).float().to(dev)
            mfcc = torch.from_numpy(m16.astype(np.float32)).to(dev)
        return text, vec, audio, mfcc, samples['vid_indices'][keys]

    def yield_batch(self, train):
        samples = self.train_samples if train else self.val_samples
        num_data = self.num_train_samples if train else self.num_val_samples
        spk = self.train_speaker_model if train else self.val_speaker_model
        bs = self.args.batch_size
        for _ in range((num_data + bs - 1) // bs):
            keys = np.random.choice(num_data, size=bs, replace=True)
            text, vec, audio, mfcc, cur_vids = self._gather(samples, keys)
            vids = None
            if spk is not None and spk.__class__.__name__ == 'Vocab':
                pool = np.setdiff1d(list(spk.word2index.values()), cur_vids)  # speakers NOT in this batch (:625-630)
                vids = torch.from_numpy(np.random.choice(pool, size=bs)).long().to(self.device)
                self._check_speaker_ids(vids)
            yield text, vec, audio, mfcc, vids

    def per_train_epoch(self):
        self.s2ag_generator.train()
        self.s2ag_discriminator.train()
        total, n = 0., 0
        self.meta_info['iter'] = 0
        num_batches = self.num_train_samples // self.args.batch_size + 1
        for text, vec, audio, mfcc, vids in self.yield_batch(train=True):
            loss, *_ = self.forward_pass_s2ag(text, audio, mfcc, vec, vids, train=True)
            total += loss
            self.iter_info['s2ag_loss'] = loss
            self.meta_info['iter'] += 1
        self.epoch_info['mean_s2ag_loss'] = total / num_batches
        self.io.print_log('\tmean_s2ag_loss: {}'.format(self.epoch_info['mean_s2ag_loss']))

    def per_val_epoch(self):
        self.s2ag_generator.eval()
        self.s2ag_discriminator.eval()
        total = 0.
        num_batches = self.num_val_samples // self.args.batch_size + 1
        for text, vec, audio, mfcc, vids in self.yield_batch(train=False):
            with torch.no_grad():
                loss, *_ = self.forward_pass_s2ag(text, audio, mfcc, vec, vids, train=False)
            total += loss
        self.epoch_info['mean_s2ag_loss'] = total / num_batches
        if self.epoch_info['mean_s2ag_loss'] < self.best_s2ag_loss and self.meta_info['epoch'] > self.min_train_epochs:
            self.best_s2ag_loss = self.epoch_info['mean_s2ag_loss']
            self.best_s2ag_loss_epoch = self.meta_info['epoch']
            self.s2ag_loss_updated = True
        else:
            self.s2ag_loss_updated = False

    def load_model_at_epoch(self, epoch='best'):
        name, e, l = get_epoch_and_loss(self.args.work_dir_s2ag, epoch=epoch)
        if name is None:
            print('Warning! No saved model found.' if epoch == 'best' else
                  'Warning! No saved model found at epoch {}.'.format(epoch))
            return False
        self.best_s2ag_loss_epoch, self.best_s2ag_loss = e, l
        loaded = torch.load(os.path.join(self.args.work_dir_s2ag, name), map_location=self.device)
        self.s2ag_generator.load_state_dict(loaded['gen_model_dict'])
        self.s2ag_discriminator.load_state_dict(loaded['dis_model_dict'])
        return True

    def load_trimodal(self):
        path = os.path.join(self.base_path, 'outputs', 'trimodal_gen.pth.tar')
        if os.path.exists(path):
            ck = torch.load(path, map_location=self.device)
            self.trimodal_generator.load_state_dict(ck['trimodal_gen_dict'])
            return True
        return False  # synthetic runs: random-init frozen baseline

    def train(self):
        """processor_v2.py:1032-1069: frozen baseline from outputs/trimodal_gen.pth.tar (random-init when the file is
        absent: synthetic runs), optional resume, epoch loop with validation and checkpointing (rank 0 writes)."""
        self.load_trimodal()
        if getattr(self.args, "s2ag_load_last_best", False):
            found = self.load_model_at_epoch(epoch=self.args.s2ag_start_epoch)
            if not found and self.args.s2ag_start_epoch != 'best':
                print('Warning! Trying to load best known model for s2ag: ', end='')
                found = self.load_model_at_epoch(epoch='best')
                print('loaded.' if found else 'none found.')
            self.args.s2ag_start_epoch = self.best_s2ag_loss_epoch if found else 0
            if not found:
                print('Warning! Starting at epoch 0')
        else:
            self.args.s2ag_start_epoch = 0
        for epoch in range(self.args.s2ag_start_epoch, self.args.s2ag_num_epoch):
            self.meta_info['epoch'] = epoch
            self.io.print_log('s2ag training epoch: {}'.format(epoch))
            self.per_train_epoch()
            self.io.print_log('Done.')
            if epoch % self.args.val_interval == 0 or epoch + 1 == self.args.s2ag_num_epoch:
                self.io.print_log('s2ag val epoch: {}'.format(epoch))
                self.per_val_epoch()
                self.io.print_log('Done.')
            if self.rank == 0 and (self.s2ag_loss_updated or
                                   (epoch % self.args.save_interval == 0 and epoch > self.min_train_epochs)):
                os.makedirs(self.args.work_dir_s2ag, exist_ok=True)
                torch.save({'gen_model_dict': self.s2ag_generator.state_dict(),
                            'dis_model_dict': self.s2ag_discriminator.state_dict()},
                           os.path.join(self.args.work_dir_s2ag, 'epoch_{:06d}_loss_{:.4f}_model.pth.tar'.format(
                               epoch, self.epoch_info['mean_s2ag_loss'])))

    # ------------------------------------------------------------------ inference entry points
    def generate_gestures(self, samples_to_generate=10, randomized=True, load_saved_model=True,
                          s2ag_epoch='best', make_video=False, calculate_metrics=True):
        """34-frame batched evaluation (processor_v2.py:1071-1142): eval mode, batch 2048, the full step under
        no_grad; L1 / joint MAE / acceleration difference of `push_samples` (:738-774) are reduced on the device
        (csrc/longform.cu) and read back once at the end, and so are the two FGD evaluators when the embedding-net
        checkpoint exists (:750-751, :1116-1136).  Returns the reference's loss_dict (plus 'accel*', 'clips')."""
        if make_video:
            raise NotImplementedError("video rendering is outside the hot path (SURVEY 8)")
        if load_saved_model:
            assert self.load_model_at_epoch(epoch=s2ag_epoch), 'Speech to emotive gestures model not found'
            self.load_trimodal()
        for net in (self.trimodal_generator, self.s2ag_generator, self.s2ag_discriminator):
            net.eval()
        batch_size = 2048
        test = self.data_loader['test_data_s2ag'].samples
        n = min(samples_to_generate, self.num_test_samples)
        cfg = self.s2ag_config_args
        mean = torch.tensor(np.squeeze(np.array(cfg.mean_dir_vec)), dtype=torch.float32, device=self.device)
        acc = torch.zeros(2, 3, device=self.device)   # rows: ours, tri-modal; cols: L1, joint MAE, accel (AverageMeter sums)
        start_time = time.time()
        spk_values = np.array(list(self.test_speaker_model.word2index.values()))
        for s in range(0, n, batch_size):
            keys = np.arange(s, min(n, s + batch_size))
            if randomized:
                keys = np.random.choice(self.num_test_samples, size=len(keys), replace=False)
            text, vec, audio, mfcc, _ = self._gather(test, keys)
            # return_batch draws a random speaker of the test speaker model per sample (:722-724)
            vids = torch.from_numpy(np.random.choice(spk_values, size=len(keys))).long().to(self.device)
            self._check_speaker_ids(vids)
            with torch.no_grad():
                self.gan_step_async(text, audio, mfcc, vec, vids, train=False)
                if calculate_metrics:
                    acc[0] += ops.pose_metrics(self.last_out, vec, mean, cfg.n_pre_poses) * len(keys)
                    acc[1] += ops.pose_metrics(self.last_out_trimodal, vec, mean, cfg.n_pre_poses) * len(keys)
                    if self.evaluator_trimodal:   # push_samples(:750-751): (text, audio, generated, target)
                        self.evaluator_trimodal.push_samples(text, audio, self.last_out_trimodal, vec)
                    if self.evaluator:
                        self.evaluator.push_samples(text, audio, self.last_out, vec)
        (l1, mae, accel), (l1_t, mae_t, accel_t) = (acc / max(n, 1)).tolist()
        loss_dict = {'loss_trimodal': l1_t, 'joint_mae_trimodal': mae_t, 'loss': l1, 'joint_mae': mae,
                     'accel_trimodal': accel_t, 'accel': accel, 'clips': n}
        elapsed = time.time() - start_time
        for tag, ev, sfx, vals in (('[VAL Trimodal]\t', self.evaluator_trimodal, '_trimodal', (l1_t, mae_t, accel_t)),
                                   ('[VAL Ours]\t\t', self.evaluator, '', (l1, mae, accel))):
            if ev and ev.get_no_of_samples() > 0:
                fd, feat_d = ev.get_scores()
                print(tag + 'loss: {:.3f}, joint mae: {:.5f}, accel diff: {:.5f},FGD: {:.3f}, feat_D: {:.3f} / {:.1f}s'
                      .format(vals[0], vals[1], vals[2], fd, feat_d, elapsed))
                loss_dict['frechet' + sfx], loss_dict['feat_dist' + sfx] = fd, feat_d
            else:
                print(tag + 'loss: {:.3f}, joint mae: {:.3f} / {:.1f}s'.format(vals[0], vals[1], elapsed))
        print('Total time taken: {:.2f} seconds.'.format(time.time() - start_time))
        return loss_dict

    def _check_speaker_ids(self, vids):
        """nn.Embedding device-asserts on an out-of-range id (and so does csrc/tcn.cu); fail on the host with a message
        instead: the generator's speaker table is sized from the TRAIN speaker model (:147-150)."""
        emb = getattr(self.s2ag_generator, "speaker_embedding", None)
        if emb is not None and vids is not None and vids.numel():
            rows = emb[0].weight.shape[0]
            hi = int(vids.max())
            if hi >= rows or int(vids.min()) < 0:
                raise IndexError("speaker id %d outside the generator's speaker table (%d rows)" % (hi, rows))

    def render_clip(self, data_params, vid_name, sample_idx, samples_to_generate,
                    clip_poses, clip_audio, sample_rate, clip_words, clip_time,
                    test_samples=None, clip_idx=0, unit_time=None, speaker_vid_idx=0, check_duration=True,
                    fade_out=False, make_video=False, save_pkl=False):
        """Same signature / returns as processor_v2.py:1144-1439: chunked autoregressive synthesis of one clip
        (unit 34 frames, stride 30, seed hand-off of 4 frames, linear blend, optional quadratic fade-out), then
        direction vectors -> joints.  -> (clip_poses_resampled, out_poses_trimodal, out_poses).
        One clip is a lock-step batch of 1; `generate_gestures_by_dataset` renders many clips per batch."""
        from . import longform
        if make_video:
            raise NotImplementedError("video rendering is outside the hot path (SURVEY 8)")
        if test_samples is not None and \
                '{}_{:.2f}_{:.2f}'.format(vid_name, clip_time[0], clip_time[1]) not in test_samples:
            return [], [], [], []
        if check_duration:
            dur = clip_time[1] - clip_time[0]
            if dur < data_params['clip_duration_range'][0] or dur > data_params['clip_duration_range'][1]:
                return None, None, None
        pc = self._prepare_clip(vid_name, clip_poses, clip_audio, sample_rate, clip_words, clip_time, unit_time,
                                speaker_vid_idx, clip_idx)
        print('Sample {} of {}'.format(sample_idx + 1, samples_to_generate))
        (res,) = longform.render_lockstep(self, [pc], fade_out=fade_out, audio_sr=data_params.get('audio_sr', sample_rate))
        if save_pkl:
            self._save_pkl(pc, res, fade_out, data_params.get('audio_sr', sample_rate))
        self.last_render = dict(out_dir_vec_trimodal=res[0], out_dir_vec=res[1])
        return pc.clip_poses_resampled, res[2], res[3]

    def _prepare_clip(self, vid_name, clip_poses, clip_audio, sample_rate, clip_words, clip_time, unit_time,
                      speaker_vid_idx, clip_idx=0):
        from . import longform
        return longform.prepare_clip(self.s2ag_config_args, self.lang_model, self.pose_dim, vid_name, clip_poses,
                                     clip_audio, sample_rate, clip_words, clip_time, unit_time, speaker_vid_idx,
                                     n_speakers=self.s2ag_generator.z_obj.n_words if self.s2ag_generator.z_obj else None,
                                     clip_idx=clip_idx)

    def _save_pkl(self, pc, res, fade_out, audio_sr):
        """processor_v2.py:1418-1437 (two pickles per clip)"""
        import pickle
        mean = np.squeeze(np.array(self.s2ag_config_args.mean_dir_vec))
        prefix = '{}_s{}_{:.2f}_{:.2f}'.format(pc.vid_name, pc.speaker_vid_idx, pc.clip_time[0], pc.clip_time[1])
        sentence = ' '.join(w[0] for w in pc.clip_words)
        os.makedirs(self.args.video_save_path, exist_ok=True)
        for tag, vec, poses in (('trimodal', res[0], res[2]), ('s2ag', res[1], res[3])):
            if vec is None:
                continue
            with open(os.path.join(self.args.video_save_path, '{}_{}.pkl'.format(prefix, tag)), 'wb') as f:
                pickle.dump({'sentence': sentence, 'audio': np.asarray(pc.clip_audio, dtype=np.float32),
                             'out_dir_vec': vec + mean, 'out_poses': poses,
                             'aux_info': '{}_{}_{}'.format(pc.vid_name, pc.speaker_vid_idx, pc.clip_idx),
                             'human_dir_vec': pc.target_dir_vec + mean}, f)

    def generate_gestures_by_dataset(self, dataset, data_params, check_duration=True,
                                     samples=None, randomized=True, fade_out=False,
                                     load_saved_model=True, s2ag_epoch='best',
                                     make_video=False, save_pkl=False):
        """Same signature as processor_v2.py:1441-1444.  dataset 'ted_db': the clip records
        `(words, poses, _, audio, _, _, {'vid','start_frame_no','end_frame_no','start_time','end_time'})` come from
        `data_params['clips']` (any iterable in the LMDB record layout of :1486-1495) or, when the `lmdb` module and
        `data_params['env_file']` are available, from that LMDB; consecutive records of one video are merged exactly
        like :1497-1523 and every merged clip goes through the chunked synthesis of `render_clip`.  B200-first: merged
        clips are queued and rendered `data_params['lockstep_batch']` (default 256) at a time in lock-step
        (longform.render_lockstep) instead of one by one.  Returns the list of per-clip results
        (vid_name, clip_poses_resampled, out_poses_trimodal, out_poses); the reference returns None."""
        from . import longform
        if make_video:
            raise NotImplementedError("video rendering is outside the hot path (SURVEY 8)")
        if load_saved_model:
            assert self.load_model_at_epoch(epoch=s2ag_epoch), 'Speech to emotive gestures model not found'
            self.load_trimodal()
        for net in (self.trimodal_generator, self.s2ag_generator, self.s2ag_discriminator):
            net.eval()
        overall_start_time = time.time()
        if dataset.lower() != 'ted_db':
            raise NotImplementedError("only the 'ted_db' record layout is supported (GENEA BVH I/O is out of scope)")
        if 'clip_duration_range' not in data_params.keys():
            data_params['clip_duration_range'] = [5, 12]
        records = self._clip_records(data_params)
        if randomized:
            records = list(records)
            records = [records[i] for i in np.random.randint(0, len(records), size=len(records))] if records else []
        lock = int(data_params.get('lockstep_batch', 256))
        audio_sr = data_params.get('audio_sr', self.audio_sr)
        queue, results = [], []

        def flush():
            if not queue:
                return
            rendered = longform.render_lockstep(self, queue, fade_out=fade_out, audio_sr=audio_sr)
            for pc, res in zip(queue, rendered):
                if save_pkl:
                    self._save_pkl(pc, res, fade_out, audio_sr)
                results.append((pc.vid_name, pc.clip_poses_resampled, res[2], res[3]))
            queue.clear()

        def submit(vid_name, poses_all, audio_all, words_all, time_all):
            spk = np.random.randint(0, self.test_speaker_model.n_words) if randomized else 0
            if samples is not None and '{}_{:.2f}_{:.2f}'.format(vid_name, time_all[0], time_all[1]) not in samples:
                return
            if check_duration:
                dur = time_all[1] - time_all[0]
                if dur < data_params['clip_duration_range'][0] or dur > data_params['clip_duration_range'][1]:
                    return
            queue.append(self._prepare_clip(vid_name, poses_all, audio_all, audio_sr, words_all, time_all, None, spk))
            if len(queue) >= lock:
                flush()

        clip_vid_name, frames_all = '', [-2, -2]
        poses_all = audio_all = words_all = time_all = None
        for video in records:
            vid_name = video[6]['vid']
            if not (samples is None or any(vid_name in prefix for prefix in samples)):
                continue
            clip_poses, clip_audio, clip_words = video[1], video[3], video[0]
            clip_frames = [video[6]['start_frame_no'], video[6]['end_frame_no']]
            clip_time = [video[6]['start_time'], video[6]['end_time']]
            if vid_name != clip_vid_name or clip_frames[0] - 1 > frames_all[1]:
                if clip_vid_name != '':
                    # (the reference passes the NEW record's vid_name with the accumulated clip here, :1503; the
                    # accumulated clip's own name is used instead)
                    submit(clip_vid_name, poses_all, audio_all, words_all, time_all)
                clip_vid_name = vid_name
                poses_all, audio_all, words_all = clip_poses, clip_audio, list(clip_words)
                frames_all, time_all = list(clip_frames), list(clip_time)
            else:
                last = clip_frames[0] - frames_all[0]
                poses_all = np.concatenate((poses_all[:last], clip_poses), axis=0)
                audio_all = np.concatenate((audio_all[:int((clip_time[0] - time_all[0]) * 16000)], clip_audio))
                for word in clip_words:
                    if word not in words_all:
                        words_all.append(word)
                frames_all[1] = clip_frames[1]
                time_all[1] = clip_time[1]
        if clip_vid_name != '' and data_params.get('render_last', True):
            submit(clip_vid_name, poses_all, audio_all, words_all, time_all)   # (the reference drops the last clip)
        flush()
        print('Total time taken: {:.2f} seconds.'.format(time.time() - overall_start_time))
        return results

    @staticmethod
    def _clip_records(data_params):
        if 'clips' in data_params:
            return data_params['clips']
        try:
            import lmdb
            import pyarrow
        except ImportError as e:
            raise RuntimeError("data_params['clips'] not given and the lmdb/pyarrow reader is unavailable: %s" % e)
        env = lmdb.open(data_params['env_file'], readonly=True, lock=False)
        with env.begin(write=False) as txn:
            return [pyarrow.deserialize(buf) for _, buf in txn.cursor()]

    @torch.no_grad()
    def synthesize_long_form(self, text_chunks, mfcc_chunks, audio_chunks, vid_indices, seed_poses=None,
                             run_trimodal=False):
        """Tensor-level lock-step chunked synthesis over PRE-COMPUTED per-chunk inputs (BASELINE config 5):
        text_chunks [B, n_chunks, 34] int64; mfcc_chunks [B, n_chunks, 37, 71]; audio_chunks [B, n_chunks, audio_len]
        (only read when run_trimodal); vid_indices [B].  Seed hand-off and blend as render_clip (processor_v2.py:
        1282-1327) on the device.  Returns the generator's dir-vec sequence [B, 34 + 30*(n_chunks-1), pose_dim]
        (and the tri-modal baseline's as `self.last_long_form_trimodal` when run_trimodal)."""
        G = self.s2ag_generator
        G.eval()
        self.trimodal_generator.eval()
        B, n_chunks = text_chunks.shape[0], text_chunks.shape[1]
        T, P, n_pre = self.time_steps, self.pose_dim, self.s2ag_config_args.n_pre_poses
        total = T + (T - n_pre) * (n_chunks - 1)
        seed = torch.zeros(B, T, P + 1, device=self.device)
        if seed_poses is not None:
            seed[:, :n_pre, :-1] = seed_poses[:, :n_pre]
            seed[:, :n_pre, -1] = 1
        nets = ((self.trimodal_generator, True),) if run_trimodal else ()
        nets += ((G, False),)
        state = [dict(result=torch.zeros(B, total, P, device=self.device), pre=seed.clone(),
                      nxt=torch.empty_like(seed)) for _ in nets]
        for c in range(n_chunks):
            for (net, is_tri), st in zip(nets, state):
                feat = audio_chunks[:, c] if is_tri else mfcc_chunks[:, c]
                out, *_ = net(st["pre"], text_chunks[:, c], feat, vid_indices)
                ops.longform_blend(out, st["result"], c, n_pre, pre_next=st["nxt"])
                st["pre"], st["nxt"] = st["nxt"], st["pre"]
        if run_trimodal:
            self.last_long_form_trimodal = state[0]["result"]
        return state[-1]["result"]
