"""CPU: the oracle (oracle/s2ag_oracle.py) against the fixtures recorded from the UNMODIFIED reference
(tests/golden/s2ag_reference_golden.npz, written by oracle/gen_golden.py).  This is the parity pin:
the reference ships no tests or golden vectors for this path."""
import numpy as np
import pytest
import torch

from common import O, GOLDEN, rel


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN, allow_pickle=False)


def _ref_state_dicts(n_words, n_spk):
    """state_dicts in the reference naming, built WITHOUT the package under test: shapes are taken
    from the oracle's own knowledge of the reference constructors."""
    from common import build_nets, sd_cpu
    # build_nets only provides names/shapes here (parameter containers); values come from fill_state_dict
    nets = build_nets("full", n_words, n_spk, torch.device("cpu"))
    return [sd_cpu(n) for n in nets]


def test_oracle_modules_match_reference_fixtures(gold):
    n_words, n_spk_rows, B, seed, n_spk = (int(x) for x in gold["meta"])
    g_sd, t_sd, d_sd, c_sd = _ref_state_dicts(n_words, n_spk_rows)
    batch, eps_list, rand_idx = O.synthetic_batch(B, n_words, n_spk, 36267, seed)
    text, audio, mfcc, target, vid = batch
    pre = target.new_zeros(B, 34, 28)
    pre[:, :4, :-1] = target[:, :4]
    pre[:, :4, -1] = 1
    cp = lambda sd: {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        gs = cp(g_sd)
        out, z, mu, lv = O.pose_generator(gs, pre, text, mfcc, vid, eps_list[0], True)
        assert rel(out, torch.from_numpy(gold["g_out"])) < 1e-5
        assert rel(z, torch.from_numpy(gold["g_z"])) < 1e-5 and rel(mu, torch.from_numpy(gold["g_mu"])) < 1e-5
        assert rel(gs["aff_encoder.batch_norm1.running_mean"], torch.from_numpy(gold["g_rm"])) < 1e-5
        assert rel(gs["audio_encoder.batch_norm4.running_var"], torch.from_numpy(gold["g_rv"])) < 1e-5
        oe = O.pose_generator(cp(g_sd), pre, text, mfcc, vid, eps_list[0], False)[0]
        assert rel(oe, torch.from_numpy(gold["g_out_eval"])) < 1e-5
        ot = O.pose_generator_trimodal(cp(t_sd), pre, text, audio, vid, eps_list[0], True)[0]
        assert rel(ot, torch.from_numpy(gold["t_out"])) < 1e-5
        assert rel(O.aff_discriminator(cp(d_sd), target, True), torch.from_numpy(gold["d_out"])) < 1e-5
        assert rel(O.conv_discriminator(cp(c_sd), target, True), torch.from_numpy(gold["c_out"])) < 1e-5


def test_oracle_attention_matches_reference_fixture(gold):
    sd = {"linear1.weight": torch.zeros(32, 32), "linear1.bias": torch.zeros(32), "linear2.weight": torch.zeros(1, 32),
          "linear2.bias": torch.zeros(1)}
    O.fill_state_dict(sd, 200, scale=2.0)
    x = torch.from_numpy(np.random.RandomState(7).normal(0, 1, size=(3, 150, 32)).astype(np.float32))
    o, a = O.attention(x, sd["linear1.weight"], sd["linear1.bias"], sd["linear2.weight"], sd["linear2.bias"])
    assert rel(o, torch.from_numpy(gold["att_out"])) < 1e-5 and rel(a, torch.from_numpy(gold["att_alpha"])) < 1e-5


def test_oracle_gan_step_matches_reference_fixture(gold):
    """two consecutive iterations of the unmodified Processor.forward_pass_s2ag (train=True)"""
    n_words, n_spk_rows, B, seed, n_spk = (int(x) for x in gold["meta"])
    g_sd, t_sd, d_sd, _ = _ref_state_dicts(n_words, n_spk_rows)
    g_sd, t_sd, d_sd = O.as_leaves(g_sd), O.as_leaves(t_sd), O.as_leaves(d_sd)
    batch, eps_list, rand_idx = O.synthetic_batch(B, n_words, n_spk, 36267, seed)
    state = {}
    for it in range(2):
        r = O.gan_step(g_sd, d_sd, t_sd, batch, eps_list, rand_idx, O.CFG, state, train=True)
        assert abs(r["ret"] - float(gold["step%d_ret" % it])) < 1e-5
        assert rel(r["out_dir_vec"], torch.from_numpy(gold["step%d_out" % it])) < 1e-4
        assert rel(r["out_trimodal"], torch.from_numpy(gold["step%d_out_tri" % it])) < 1e-4
        got = np.array([r[k] for k in ("dis", "huber", "gen", "kld", "div", "total")])
        assert np.allclose(got, gold["step%d_losses" % it], rtol=1e-4, atol=1e-6)
    # post-step weights after two Adam steps: bulk within 2e-3 of max|w|, never further than Adam's reach
    lr = O.CFG["learning_rate"]
    for name, sd in (("g", g_sd), ("d", d_sd)):
        keys = [str(k) for k in gold["post_%s_keys" % name]]
        heads = gold["post_%s_head" % name]
        for k, h in zip(keys, heads):
            mine = np.resize(sd[k].detach().flatten()[:8].numpy(), 8)
            assert np.abs(mine - h).max() <= 2.5 * 2 * lr + 1e-6, k
