"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported with stubbed third-party modules -- SURVEY Appendix A, recipes A and B)
on seeded synthetic inputs, and pins oracle/s2ag_oracle.py against it.

Run in the build container (the reference checkout does not exist on the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py
De-randomisation: every nn.Dropout p=0 and nn.GRU.dropout=0, re_parametrize replaced
by an injected-eps version, torch.randperm patched to the injected permutation.  BatchNorm stays
in train mode.  Weights come from oracle.fill_state_dict (numpy MT19937 => portable), so the
fixtures hold only inputs' seeds and the reference's OUTPUTS.
"""
import os
import sys
import types
from unittest.mock import MagicMock

sys.dont_write_bytecode = True
REF = os.environ.get("S2AG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)


class Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return MagicMock()


for n in ['librosa', 'librosa.feature', 'librosa.display', 'lmdb', 'matplotlib', 'matplotlib.pyplot',
          'matplotlib.ticker', 'matplotlib.animation', 'mpl_toolkits', 'mpl_toolkits.mplot3d',
          'python_speech_features', 'h5py', 'umap', 'soundfile', 'fasttext', 'transforms3d', 'configargparse',
          'pyttsx3', 'nltk', 'nltk.corpus']:
    m = Stub(n)
    m.__path__ = []
    sys.modules[n] = m

import numpy as np  # noqa: E402
import torch  # noqa: E402

if not torch.cuda.is_available():  # AffEncoder hard-codes .cuda() (net/multimodal_context_net_v2.py:106,115,163)
    torch.Tensor.cuda = lambda self, *a, **k: self

import processor_v2 as RP  # noqa: E402  (the reference)
import net.embedding_net as ren  # noqa: E402
import net.ser_att_conv_rnn_v2 as ratt  # noqa: E402
from utils.vocab import Vocab  # noqa: E402
from types import SimpleNamespace as NS  # noqa: E402

import s2ag_oracle as O  # noqa: E402

N_WORDS, N_SPK, B = 64, 24, 4
OUT = os.path.join(HERE, "..", "tests", "golden")


def derand(net):
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.GRU):
            m.dropout = 0.0


def rel(a, b):
    return float((a - b).abs().max() / max(b.abs().max().item(), 1e-9))


def main():
    os.makedirs(OUT, exist_ok=True)
    cfg = NS(**O.CFG)
    spk = Vocab('vid', insert_default_tokens=False)
    [spk.index_word('v%d' % i) for i in range(N_SPK)]
    ctors = (lambda: RP.PoseGenerator(cfg, 27, N_WORDS, 300, None, 71, 37, 34, z_obj=spk),
             lambda: RP.PGT(cfg, 27, N_WORDS, 300, None, z_obj=spk),
             lambda: RP.AffDiscriminator(27),
             lambda: RP.CDT(27))
    G, T, D, C = (c() for c in ctors)
    for i, net in enumerate((G, T, D, C)):
        derand(net)
        O.fill_state_dict(net.state_dict(), 100 + i)

    def clone(i, net):  # (deepcopy fails on old-style weight_norm modules)
        n2 = ctors[i]()
        derand(n2)
        n2.load_state_dict(net.state_dict())
        return n2

    def sd_copy(net):
        return {k: v.detach().clone() for k, v in net.state_dict().items()}
    batch, eps_list, rand_idx = O.synthetic_batch(B, N_WORDS, N_SPK, 36267, seed=1234)
    text, audio, mfcc, target, vid = batch
    pre = target.new_zeros(B, 34, 28)
    pre[:, :4, :-1] = target[:, :4]
    pre[:, :4, -1] = 1

    # ---------------- module level (train-mode BN, fresh copies so running stats start equal)
    fix = {}
    import copy
    eps_it = {'i': 0, 'seq': [eps_list[0]]}

    def inj(mu, lv):
        e = eps_it['seq'][eps_it['i'] % len(eps_it['seq'])]
        eps_it['i'] += 1
        return mu + e * torch.exp(0.5 * lv)
    ren.re_parametrize = inj

    with torch.no_grad():
        g2, t2, d2, c2 = (clone(i, n) for i, n in enumerate((G, T, D, C)))
        ro = g2(pre, text, mfcc, vid)
        oo = O.pose_generator(sd_copy(G), pre, text, mfcc, vid, eps_list[0], True)
        print("G   oracle vs reference rel err", rel(oo[0], ro[0]))
        assert rel(oo[0], ro[0]) < 1e-5
        fix['g_out'], fix['g_z'], fix['g_mu'], fix['g_lv'] = (x.numpy() for x in ro)
        rt = t2(pre, text, audio, vid)
        ot = O.pose_generator_trimodal(sd_copy(T), pre, text, audio, vid, eps_list[0], True)
        print("T   oracle vs reference rel err", rel(ot[0], rt[0]))
        assert rel(ot[0], rt[0]) < 1e-5
        fix['t_out'] = rt[0].numpy()
        rd = d2(target)
        od = O.aff_discriminator(sd_copy(D), target, True)
        print("D   oracle vs reference rel err", rel(od, rd))
        assert rel(od, rd) < 1e-5
        fix['d_out'] = rd.numpy()
        rc = c2(target)
        oc = O.conv_discriminator(sd_copy(C), target, True)
        print("CD  oracle vs reference rel err", rel(oc, rc))
        assert rel(oc, rc) < 1e-5
        fix['c_out'] = rc.numpy()
        # eval-mode generator (running stats as filled)
        g3 = clone(0, G).eval()
        eps_it['i'] = 0
        fix['g_out_eval'] = g3(pre, text, mfcc, vid)[0].numpy()
        oe = O.pose_generator(sd_copy(G), pre, text, mfcc, vid, eps_list[0], False)
        assert rel(oe[0], torch.from_numpy(fix['g_out_eval'])) < 1e-5
        # running statistics after one train-mode pass
        fix['g_rm'] = g2.state_dict()['aff_encoder.batch_norm1.running_mean'].numpy()
        fix['g_rv'] = g2.state_dict()['audio_encoder.batch_norm4.running_var'].numpy()

    # attention (named off-path): net/ser_att_conv_rnn_v2.py:16-34
    att = ratt.Attention(32, 32, False)
    O.fill_state_dict(att.state_dict(), 200, scale=2.0)
    xa = torch.from_numpy(np.random.RandomState(7).normal(0, 1, size=(3, 150, 32)).astype(np.float32))
    with torch.no_grad():
        ao, aa = att(xa)
        bo, ba = O.attention(xa, att.linear1.weight, att.linear1.bias, att.linear2.weight, att.linear2.bias)
    assert rel(bo, ao) < 1e-5 and rel(ba, aa) < 1e-5
    fix['att_out'], fix['att_alpha'] = ao.numpy(), aa.numpy()

    # ---------------- step level: the unmodified Processor.forward_pass_s2ag, two consecutive iterations
    pr = RP.Processor.__new__(RP.Processor)
    pr.s2ag_config_args, pr.meta_info, pr.use_mfcc = cfg, dict(epoch=1, iter=0), True
    pr.trimodal_generator, pr.s2ag_generator, pr.s2ag_discriminator = T, G, D
    pr.s2ag_gen_optimizer = torch.optim.Adam(G.parameters(), lr=cfg.learning_rate, betas=(0.5, 0.999))
    pr.s2ag_dis_optimizer = torch.optim.Adam(D.parameters(), lr=cfg.learning_rate * cfg.discriminator_lr_weight,
                                             betas=(0.5, 0.999))
    g_sd, d_sd, t_sd = O.as_leaves(G.state_dict()), O.as_leaves(D.state_dict()), O.as_leaves(T.state_dict())
    opt_state = {}
    real_randperm = torch.randperm
    torch.randperm = lambda n, *a, **k: rand_idx.clone()
    captured = {}
    G.register_forward_hook(lambda m, i, o: captured.setdefault('g', []).append(o[0].detach().clone()))
    T.register_forward_hook(lambda m, i, o: captured.setdefault('t', []).append(o[0].detach().clone()))
    for it in range(2):
        captured.clear()
        eps_it['i'] = 0
        eps_it['seq'] = eps_list  # G pass 1, T pass, G pass 2, G pass 3
        ret = pr.forward_pass_s2ag(text, audio, mfcc, target, vid, train=True)[0]
        ores = O.gan_step(g_sd, d_sd, t_sd, batch, eps_list, rand_idx, O.CFG, opt_state, train=True)
        out_ref = captured['g'][1]
        print("step %d: ret ref %.6f oracle %.6f | out rel %.2e" % (it, ret, ores['ret'], rel(ores['out_dir_vec'], out_ref)))
        assert abs(ret - ores['ret']) < 1e-5 and rel(ores['out_dir_vec'], out_ref) < 1e-4
        fix['step%d_ret' % it] = np.float32(ret)
        fix['step%d_out' % it] = out_ref.numpy()
        fix['step%d_out_tri' % it] = captured['t'][0].numpy()
        fix['step%d_losses' % it] = np.array([ores[k] for k in ('dis', 'huber', 'gen', 'kld', 'div', 'total')],
                                             dtype=np.float32)
    torch.randperm = real_randperm
    # post-step weights (after two Adam steps): a checksum slice per parameter, G and D
    for name, net, osd in (('g', G, g_sd), ('d', D, d_sd)):
        sd = net.state_dict()
        keys = [k for k in sorted(sd.keys()) if sd[k].dtype.is_floating_point]
        # Adam turns the fp32 rounding noise of mathematically-zero gradients (conv biases feeding a
        # BatchNorm) into +-lr updates, so those tensors are compared to 2*lr*steps absolute instead.
        worst = 0.0
        for k in keys:
            r = rel(osd[k].detach(), sd[k])
            a = float((osd[k].detach() - sd[k]).abs().max())
            if r > 2e-3:
                print("   noisy:", k, "rel %.2e abs %.2e" % (r, a))
                assert a <= 2 * 2 * cfg.learning_rate + 1e-6, k
            else:
                worst = max(worst, r)
        print("post-step %s weights: oracle vs reference worst rel %.2e" % (name, worst))
        fix['post_%s_head' % name] = np.stack([np.resize(sd[k].flatten()[:8].numpy(), 8) for k in keys])
        fix['post_%s_sum' % name] = np.array([sd[k].double().sum().item() for k in keys])
        fix['post_%s_keys' % name] = np.array(keys)
    fix['meta'] = np.array([N_WORDS, spk.n_words, B, 1234, N_SPK])  # spk.n_words = rows of the speaker table
    np.savez_compressed(os.path.join(OUT, "s2ag_reference_golden.npz"), **fix)
    print("wrote", os.path.join(OUT, "s2ag_reference_golden.npz"),
          os.path.getsize(os.path.join(OUT, "s2ag_reference_golden.npz")), "bytes")


def longform():
    """Long-form fixture: the reference's UNMODIFIED Processor.render_clip (processor_v2.py:1144-1439) on a 20 s
    synthetic clip (10 chunks), with and without fade-out.  The only injected piece is `utils.common.mfcc`
    (librosa.feature.mfcc, absent here) = oracle/frontend_oracle.mfcc_librosa; the reference's own
    get_mfcc_features, chunk schedule, word placement, seed hand-off, blend, fade-out and convert_dir_vec_to_pose run
    as shipped.  Also pins push_samples (:738-774)."""
    import copy
    import frontend_oracle as FO
    import utils.common as rcmn
    from utils.vocab import Vocab as RVocab
    from utils.average_meter import AverageMeter
    rcmn.mfcc = FO.mfcc_librosa
    cfg = NS(**O.CFG)
    spk = RVocab('vid', insert_default_tokens=False)
    [spk.index_word('v%d' % i) for i in range(N_SPK)]
    lang = RVocab('words')
    [lang.index_word('w%d' % i) for i in range(4, 56)]   # some clip words (w56..w59) stay unknown -> UNK
    G = RP.PoseGenerator(cfg, 27, N_WORDS, 300, None, 71, 37, 34, z_obj=spk)
    T = RP.PGT(cfg, 27, N_WORDS, 300, None, z_obj=spk)
    for i, net in enumerate((G, T)):
        derand(net)
        O.fill_state_dict(net.state_dict(), 100 + i)
        net.eval()
    eps = torch.from_numpy(np.random.RandomState(77).normal(0, 1, size=(1, 16)).astype(np.float32))
    ren.re_parametrize = lambda mu, lv: mu + eps * torch.exp(0.5 * lv)
    pr = RP.Processor.__new__(RP.Processor)
    pr.s2ag_config_args, pr.pose_dim, pr.device = cfg, 27, torch.device('cpu')
    pr.trimodal_generator, pr.s2ag_generator, pr.lang_model = T, G, lang
    pr.args = NS(train_s2ag=True, video_save_path='/tmp')
    pr.data_loader = {'train_data_s2ag': NS(num_mfcc=14), 'test_data_s2ag': NS(num_mfcc=14)}
    pr.best_s2ag_loss_epoch = 0
    fix = {}
    seen = []
    G.register_forward_pre_hook(lambda m, inp: seen.append((inp[1][0].numpy().copy(), inp[2][0].numpy().copy())))
    for tag, seed, dur in (('a', 1, 20.0), ('b', 2, 9.3)):
        clip = FO.synthetic_clip(seed, duration=dur)
        # the reference's fade-out also re-fits the TARGET over [start_frame, end_frame) (:1362-1367) and raises when the
        # ground-truth motion is shorter than that: give the clip 1 s more motion than audio
        clip[6]['end_time'] += 1.0
        name = '{}_{:.2f}_{:.2f}'.format(clip[6]['vid'], clip[6]['start_time'], clip[6]['end_time'])
        for fade in (False, True):
            del seen[:]
            with torch.no_grad():
                res = pr.render_clip({'audio_sr': 16000, 'clip_duration_range': [5, 12]}, clip[6]['vid'], 0, 1,
                                     clip[1].copy(), clip[3].copy(), 16000, copy.deepcopy(clip[0]),
                                     [clip[6]['start_time'], clip[6]['end_time']], test_samples=[name],
                                     speaker_vid_idx=3, check_duration=False, fade_out=fade)
            k = '%s_fade%d' % (tag, int(fade))
            fix[k + '_resampled'] = np.asarray(res[0], dtype=np.float32)
            fix[k + '_poses_tri'] = np.asarray(res[1], dtype=np.float32)
            fix[k + '_poses'] = np.asarray(res[2], dtype=np.float32)
            print('render_clip', k, 'frames', res[2].shape)
            fix[tag + '_text'] = np.stack([t for t, _ in seen])      # per-chunk word placement the generator saw
            fix[tag + '_mfcc'] = np.stack([m for _, m in seen]).astype(np.float32)   # per-chunk MFCC input
        fix[tag + '_mfcc0'] = rcmn.get_mfcc_features(clip[3][:36266], sr=16000, num_mfcc=14).astype(np.float32)
    fix['eps'] = eps.numpy()
    # push_samples
    rng = np.random.RandomState(5)
    out = torch.from_numpy(rng.normal(0, 0.3, size=(6, 34, 27)).astype(np.float32))
    tgt = torch.from_numpy(rng.normal(0, 0.3, size=(6, 34, 27)).astype(np.float32))
    la, jm, ac = AverageMeter('loss'), AverageMeter('mae_on_joint'), AverageMeter('accel')
    RP.Processor.push_samples(None, tgt.clone(), out.clone(), None, None, la, jm, ac, cfg.mean_dir_vec, 34, 4)
    fix['push_samples'] = np.array([la.avg, jm.avg, ac.avg])
    want = FO.push_samples_metrics(out.numpy(), tgt.numpy(), cfg.mean_dir_vec, 34, 4)
    assert np.allclose(fix['push_samples'], want, rtol=1e-6), (fix['push_samples'], want)
    path = os.path.join(OUT, "s2ag_longform_golden.npz")
    np.savez_compressed(path, **fix)
    print("wrote", path, os.path.getsize(path), "bytes")


def checkpoints():
    """Reference-built state_dicts: (1) key / shape / dtype schema of the four networks in the shipped configuration,
    (2) small-width checkpoints in the reference's file schema (processor_v2.py:1066-1067 {'gen_model_dict',
    'dis_model_dict'}, :1034 {'trimodal_gen_dict'}) with the reference's OWN random initialisation, plus the reference's
    eval-mode outputs for them: load_state_dict(strict=True) into this repo's classes must reproduce those outputs."""
    import json
    cfg = NS(**O.CFG)
    spk = Vocab('vid', insert_default_tokens=False)
    [spk.index_word('v%d' % i) for i in range(N_SPK)]
    nets = dict(gen=RP.PoseGenerator(cfg, 27, N_WORDS, 300, None, 71, 37, 34, z_obj=spk),
                tri=RP.PGT(cfg, 27, N_WORDS, 300, None, z_obj=spk), dis=RP.AffDiscriminator(27), cdis=RP.CDT(27))
    schema = {k: [[n, list(v.shape), str(v.dtype)] for n, v in net.state_dict().items()] for k, net in nets.items()}
    with open(os.path.join(OUT, "ref_state_dict_schema.json"), "w") as f:
        json.dump({"n_words": N_WORDS, "n_speakers": spk.n_words, "schema": schema}, f)
    small = dict(O.CFG)
    small.update(hidden_size=24, hidden_size_s2eg=24, wordembed_dim=24, n_layers=2)
    scfg = NS(**small)
    torch.manual_seed(4242)
    G = RP.PoseGenerator(scfg, 27, 40, 24, None, 71, 37, 34, z_obj=spk)
    T = RP.PGT(scfg, 27, 40, 24, None, z_obj=spk)
    D = RP.AffDiscriminator(27)
    for net in (G, T, D):   # non-trivial running statistics
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
        net.eval()
    ck = os.path.join(OUT, "ref_ckpt_tiny")
    os.makedirs(os.path.join(ck, "outputs"), exist_ok=True)
    os.makedirs(os.path.join(ck, "work"), exist_ok=True)
    torch.save({'gen_model_dict': G.state_dict(), 'dis_model_dict': D.state_dict()},
               os.path.join(ck, "work", 'epoch_{:06d}_loss_{:.4f}_model.pth.tar'.format(22, -0.0123)))
    torch.save({'trimodal_gen_dict': T.state_dict()}, os.path.join(ck, "outputs", "trimodal_gen.pth.tar"))
    batch, eps_list, _ = O.synthetic_batch(3, 40, N_SPK, 36267, seed=99)
    text, audio, mfcc, target, vid = batch
    pre = target.new_zeros(3, 34, 28)
    pre[:, :4, :-1] = target[:, :4]
    pre[:, :4, -1] = 1
    ren.re_parametrize = lambda mu, lv: mu + eps_list[0] * torch.exp(0.5 * lv)
    with torch.no_grad():
        np.savez_compressed(os.path.join(ck, "expected.npz"), g_out=G(pre, text, mfcc, vid)[0].numpy(),
                            t_out=T(pre, text, audio, vid)[0].numpy(), d_out=D(target).numpy())
    print("wrote", ck, sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(ck) for f in fs), "bytes")


def fgd():
    """the UNMODIFIED reference EmbeddingSpaceEvaluator (net/embedding_space_evaluator.py) around a
    reference-built EmbeddingNet(mode='pose') checkpoint in the reference's file schema ('embedding_dict'), fed three
    batches of (generated, real) clips.  The fixture keeps the inputs, the reference's latent features,
    reconstructions, recon_err_diff and get_scores(); the weights are oracle.fill_state_dict(seed) (portable)."""
    import tempfile
    import net.embedding_space_evaluator as rese
    import fgd_oracle as FO
    import inspect
    from scipy import linalg as _sl
    if 'disp' not in inspect.signature(_sl.sqrtm).parameters:
        # harness shim for a third-party API change: scipy >= 1.18 dropped sqrtm's `disp` argument, which the reference
        # passes (net/embedding_space_evaluator.py:138, written against scipy 1.x: disp=False -> (sqrtm, error estimate))
        _sqrtm = _sl.sqrtm
        rese.linalg = NS(sqrtm=lambda a, disp=True: (_sqrtm(a), 0.0) if disp is False else _sqrtm(a))
    cfg = NS(**O.CFG)
    lang = NS(n_words=N_WORDS, word_embedding_weights=None)
    net = ren.EmbeddingNet(cfg, 27, cfg.n_poses, N_WORDS, cfg.wordembed_dim, None, 'pose')
    sd = O.fill_state_dict(net.state_dict(), seed=777)
    net.load_state_dict(sd)
    n_batches, nb = 3, 24
    real, gen = FO.synthetic_pairs(4321, n_batches, nb)
    with tempfile.TemporaryDirectory() as base:
        os.makedirs(os.path.join(base, 'outputs'))
        torch.save({'embedding_dict': net.state_dict()}, os.path.join(base, 'outputs/embedding_net.pth.tar'))
        ev = rese.EmbeddingSpaceEvaluator(base, cfg, 27, lang, torch.device('cpu'))
    recon = []
    for r, g in zip(real, gen):
        rt, gt = torch.from_numpy(r), torch.from_numpy(g)
        ev.push_samples(None, None, gt, rt)
        with torch.no_grad():
            recon.append(ev.net(None, None, rt[:, :cfg.n_pre_poses], gt, 'pose')[6].numpy())
    fd, feat_dist = ev.get_scores()
    gfeat, rfeat = np.vstack(ev.generated_feat_list), np.vstack(ev.real_feat_list)
    # pin the restatement
    sdo = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        o_feat = FO.pose_encoder(sdo, torch.from_numpy(np.concatenate(gen))).numpy()
        o_rec = FO.pose_decoder(sdo, torch.from_numpy(o_feat)).numpy()
    o_fd, o_fdist = FO.get_scores(gfeat, rfeat)
    print("oracle vs reference: feat %.2e recon %.2e fgd %.2e feat_dist %.2e" % (
        np.abs(o_feat - gfeat).max(), np.abs(o_rec - np.concatenate(recon)).max(), abs(o_fd - fd), abs(o_fdist - feat_dist)))
    assert np.abs(o_feat - gfeat).max() < 1e-5 and np.abs(o_rec - np.concatenate(recon)).max() < 1e-5
    assert abs(o_fd - fd) < 1e-9 * max(1, abs(fd)) and abs(o_fdist - feat_dist) < 1e-9
    path = os.path.join(OUT, "s2ag_fgd_golden.npz")
    np.savez_compressed(path, weight_seed=777, pair_seed=4321, n_batches=n_batches, batch=nb, gen_feat=gfeat,
                        real_feat=rfeat, gen_recon_b0=recon[0][:6], recon_err_diff=np.array(ev.recon_err_diff),
                        frechet=np.float64(fd), feat_dist=np.float64(feat_dist),
                        keys=np.array(sorted(net.state_dict().keys())))
    print("wrote", path, os.path.getsize(path), "bytes; FGD %.6f feat_dist %.6f" % (fd, feat_dist))


if __name__ == "__main__":
    what = sys.argv[1:] or ["step", "longform", "checkpoints", "fgd"]
    if "step" in what:
        main()
    if "longform" in what:
        longform()
    if "checkpoints" in what:
        checkpoints()
    if "fgd" in what:
        fgd()
