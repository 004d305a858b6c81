"""Whole-step parity: Processor.forward_pass_s2ag (this repo, CUDA kernels) against the oracle's
restatement of processor_v2.py:776-957 on identical inputs / weights / injected random draws, and
against the fixture recorded from the unmodified reference Processor.  Two consecutive
iterations, so the Adam state, the updated weights and the BatchNorm running statistics are
exercised.  Tolerance: 1e-3 relative on losses and generated poses (north_star)."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from common import O, GOLDEN, cfg_dict, derand, sd_cpu, inject_eps, rel, check_grads
from speech2affective_gestures_b200.processor_v2 import Processor, M_DIS, M_HUBER, M_GEN, M_KLD, M_DIV, M_TOTAL
from speech2affective_gestures_b200.synthetic import make_data_loader, Vocab


def make_processor(kind, n_words, n_spk_rows, dev, batch_size=4):
    c = cfg_dict(kind)
    args = NS(no_cuda=(dev.type != "cuda"), work_dir_s2ag=None, save_log=False, print_log=False, train_s2ag=True,
              batch_size=batch_size, s2ag_num_epoch=1, val_interval=1, save_interval=10)
    dl = make_data_loader(8, 8, 8, n_words=n_words, n_speakers=n_spk_rows)
    pr = Processor("/nonexistent", args, NS(**c), dl, 27, 3, 16000)
    for net, seed in ((pr.s2ag_generator, 100), (pr.trimodal_generator, 101), (pr.s2ag_discriminator, 102)):
        derand(net)
        O.fill_state_dict(net.state_dict(), seed)
    pr.meta_info["epoch"] = 1
    return pr, c


def run_two_steps(pr, c, dev, B, n_words, n_spk, seed, gold=None):
    batch, eps_list, rand_idx = O.synthetic_batch(B, n_words, n_spk, 36267, seed)
    g_sd, d_sd, t_sd = (O.as_leaves(sd_cpu(n)) for n in (pr.s2ag_generator, pr.s2ag_discriminator,
                                                        pr.trimodal_generator))
    ocfg = dict(c)
    # oracle helpers read H / n_layers from the config through keyword defaults: patch via partials
    import functools
    O_pg, O_pt = O.pose_generator, O.pose_generator_trimodal
    O.pose_generator = functools.partial(O_pg, H=c["hidden_size_s2eg"], n_layers=c["n_layers"])
    O.pose_generator_trimodal = functools.partial(O_pt, H=c["hidden_size"], n_layers=c["n_layers"])
    try:
        state = {}
        pr.injected_rand_idx = rand_idx.to(dev)
        pr.s2ag_generator.train(); pr.s2ag_discriminator.train(); pr.trimodal_generator.train()
        dbatch = tuple(x.to(dev) for x in batch)
        for it in range(2):
            inject_eps(eps_list)
            ret = pr.forward_pass_s2ag(*dbatch[:3], dbatch[3], dbatch[4], train=True)[0]
            r = O.gan_step(g_sd, d_sd, t_sd, batch, eps_list, rand_idx, ocfg, state, train=True)
            m = pr.metrics.tolist()
            got = np.array([m[M_DIS], m[M_HUBER], m[M_GEN], m[M_KLD], m[M_DIV], m[M_TOTAL]])
            want = np.array([r[k] for k in ("dis", "huber", "gen", "kld", "div", "total")])
            assert np.allclose(got, want, rtol=1e-3, atol=1e-6), (it, got, want)
            assert abs(ret - r["ret"]) < 1e-3 * max(1.0, abs(r["ret"]))
            assert rel(pr.last_out, r["out_dir_vec"]) < 1e-3
            assert rel(pr.last_out_trimodal, r["out_trimodal"]) < 1e-3
            if gold is not None:
                assert np.allclose(got, gold["step%d_losses" % it], rtol=1e-3, atol=1e-6)
                assert abs(ret - float(gold["step%d_ret" % it])) < 1e-3
                assert rel(pr.last_out, torch.from_numpy(gold["step%d_out" % it])) < 1e-3
                assert rel(pr.last_out_trimodal, torch.from_numpy(gold["step%d_out_tri" % it])) < 1e-3
            if it == 0:  # gradients of the first iteration (identical weights on both sides)
                for name, net, key in (("G", pr.s2ag_generator, "g_grads"), ("D", pr.s2ag_discriminator, "d_grads")):
                    if name == "D":
                        continue  # D.grad was consumed (and is not recomputed in the G step by design)
                    check_grads({n_: p.grad for n_, p in net.named_parameters() if p.grad is not None}, r[key], 2e-3,
                                what="G grads, iteration 0")
        # post-step weights after two Adam steps (Adam amplifies rounding noise of ~zero gradients to
        # +-lr per step, hence the absolute bound; the bulk must agree tightly)
        lr = c["learning_rate"]
        for net, osd in ((pr.s2ag_generator, g_sd), (pr.s2ag_discriminator, d_sd)):
            msd = net.state_dict()
            tot, bad = 0, 0
            for k, v in osd.items():
                if not v.dtype.is_floating_point:
                    continue
                d = (msd[k].cpu() - v.detach()).abs()
                assert d.max().item() <= 2.5 * 2 * lr + 1e-6, k
                tot += d.numel()
                bad += int((d > 1e-3 * v.detach().abs().max().clamp_min(1e-3)).sum())
            assert bad <= 0.01 * tot, (bad, tot)
    finally:
        O.pose_generator, O.pose_generator_trimodal = O_pg, O_pt
        pr.injected_rand_idx = None


def test_gan_step_tiny_vs_oracle(dev):
    pr, c = make_processor("tiny", 40, 12, dev)
    run_two_steps(pr, c, dev, 4, 40, 12, seed=77)


@pytest.mark.gpu
def test_gan_step_full_vs_oracle_and_reference_fixture():
    dev = torch.device("cuda:0")
    gold = np.load(GOLDEN)
    n_words, n_spk_rows, B, seed, n_spk = (int(x) for x in gold["meta"])
    pr, c = make_processor("full", n_words, n_spk_rows, dev)
    run_two_steps(pr, c, dev, B, n_words, n_spk, seed, gold=gold)


@pytest.mark.gpu
def test_gan_step_b128_config2_vs_oracle():
    """BASELINE config 2: batch 128, fp32, parity within 1e-3 on losses and generated poses (one iteration
    through the oracle is ~seconds on the host)."""
    dev = torch.device("cuda:0")
    pr, c = make_processor("full", 2000, 100, dev, batch_size=128)
    batch, eps_list, rand_idx = O.synthetic_batch(128, 2000, 100, 36267, 4321)
    g_sd, d_sd, t_sd = (O.as_leaves(sd_cpu(n)) for n in (pr.s2ag_generator, pr.s2ag_discriminator,
                                                        pr.trimodal_generator))
    pr.injected_rand_idx = rand_idx.to(dev)
    for n in (pr.s2ag_generator, pr.s2ag_discriminator, pr.trimodal_generator):
        n.train()
    inject_eps(eps_list)
    dbatch = tuple(x.to(dev) for x in batch)
    ret = pr.forward_pass_s2ag(*dbatch[:3], dbatch[3], dbatch[4], train=True)[0]
    r = O.gan_step(g_sd, d_sd, t_sd, batch, eps_list, rand_idx, dict(c), {}, train=True)
    m = pr.metrics.tolist()
    got = np.array([m[M_DIS], m[M_HUBER], m[M_GEN], m[M_KLD], m[M_DIV], m[M_TOTAL]])
    want = np.array([r[k] for k in ("dis", "huber", "gen", "kld", "div", "total")])
    assert np.allclose(got, want, rtol=1e-3, atol=1e-6), (got, want)
    assert rel(pr.last_out, r["out_dir_vec"]) < 1e-3 and abs(ret - r["ret"]) < 1e-3


def test_eval_step_and_dropout_train_step(dev):
    """train=False path (per_val_epoch / generate_gestures) changes no weights; a train step with the
    shipped dropout (0.3 / 0.1) runs and yields finite losses."""
    pr, c = make_processor("tiny", 40, 12, dev)
    batch, eps_list, rand_idx = O.synthetic_batch(4, 40, 12, 36267, 9)
    dbatch = tuple(x.to(dev) for x in batch)
    before = pr.s2ag_generator.flat_params.clone()
    pr.s2ag_generator.eval(); pr.s2ag_discriminator.eval()
    with torch.no_grad():
        pr.forward_pass_s2ag(*dbatch[:3], dbatch[3], dbatch[4], train=False)
    assert torch.equal(before, pr.s2ag_generator.flat_params)
    # restore dropout as shipped
    for net in (pr.s2ag_generator, pr.s2ag_discriminator, pr.trimodal_generator):
        for m_ in net.modules():
            if isinstance(m_, torch.nn.GRU):
                m_.dropout = 0.3
        net.train()
    for blk in pr.s2ag_generator.text_encoder.tcn.network:
        blk.dropout1.p = blk.dropout2.p = 0.3
    pr.s2ag_generator.text_encoder.drop.p = 0.1
    from speech2affective_gestures_b200.net import embedding_net as men
    men.eps_source = None
    ret = pr.forward_pass_s2ag(*dbatch[:3], dbatch[3], dbatch[4], train=True)[0]
    assert np.isfinite(ret) and all(np.isfinite(pr.metrics.tolist()))
    assert not torch.equal(before, pr.s2ag_generator.flat_params)


@pytest.mark.gpu
def test_cuda_graph_replay_matches_eager_step():
    """The benchmarked path is a CUDA-graph replay of the whole iteration (three streams, events, autograd across
    streams).  From an identical state (weights, Adam moments, step counters, BatchNorm buffers, injected noise) the
    replay must produce the same metrics, poses and updated weights as the eager call the parity tests exercise."""
    dev = torch.device("cuda:0")
    B, n_words, n_spk = 16, 200, 30
    pr, c = make_processor("full", n_words, n_spk, dev, batch_size=B)
    batch, eps_list, rand_idx = O.synthetic_batch(B, n_words, n_spk, 36267, 2024)
    dbatch = tuple(x.to(dev) for x in batch)
    pr.injected_rand_idx = rand_idx.to(dev)
    for n in (pr.s2ag_generator, pr.s2ag_discriminator, pr.trimodal_generator):
        n.train()
    st = inject_eps([e.to(dev) for e in eps_list])   # device tensors: no host copy inside the capture
    try:
        # one eager pass first (registers the per-stream scratch, creates the side streams), then capture
        pr.gan_step_async(dbatch[0], dbatch[1], dbatch[2], dbatch[3], dbatch[4], True)
        torch.cuda.synchronize()
        i0 = st["i"]
        pr.capture_step(B, train=True, warmup=0)
        nets = (pr.s2ag_generator, pr.s2ag_discriminator, pr.trimodal_generator)
        state = [t for n in nets for t in n.state_dict().values()] + \
                [pr.gen_m, pr.gen_v, pr.dis_m, pr.dis_v, pr.gen_step, pr.dis_step]
        saved = [t.clone() for t in state]
        pr.load_static_inputs(*dbatch)
        pr.replay_step()
        torch.cuda.synchronize()
        m_graph, out_graph = pr.metrics.clone(), pr.last_out.clone()
        w_graph = pr.s2ag_generator.flat_params.clone()
        for t, s_ in zip(state, saved):
            t.copy_(s_)
        st["i"] = i0
        pr.gan_step_async(*dbatch, True)
        torch.cuda.synchronize()
        assert torch.allclose(pr.metrics, m_graph, rtol=1e-5, atol=1e-7), (pr.metrics, m_graph)
        assert rel(out_graph, pr.last_out) < 1e-5
        # updated weights: Adam turns a rounding-level difference of a ~zero gradient (atomic summation order) into
        # up to +-lr, so compare the bulk tightly and bound the rest
        d = (w_graph - pr.s2ag_generator.flat_params).abs()
        assert d.max().item() <= 2.5 * c["learning_rate"] + 1e-6
        assert (d > 1e-6).float().mean().item() < 0.01
    finally:
        pr.injected_rand_idx = None
