"""Frechet gesture distance on the device: `EmbeddingNet(mode='pose')` and `EmbeddingSpaceEvaluator`
against
  * the fixture recorded from the UNMODIFIED reference classes (oracle/gen_golden.py fgd): latent features,
    reconstructions, recon_err_diff, get_scores();
  * the numpy / scipy restatement (oracle/fgd_oracle.py) on seeded features of other sizes and conditionings, including
    rank-deficient covariances (fewer samples than dimensions), D < 32, and a checkpoint round trip through the
    reference's file schema.
Tolerances: features / reconstructions 1e-3 relative (fp32 kernels, north_star); the scores themselves are fp64 on
the device and are held to 1e-6 relative on identical features."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from common import O, ROOT, rel
import fgd_oracle as FO  # noqa: E402  (oracle/ is on sys.path through common)
from speech2affective_gestures_b200 import ops
from speech2affective_gestures_b200.net.embedding_net import EmbeddingNet
from speech2affective_gestures_b200.net.embedding_space_evaluator import EmbeddingSpaceEvaluator

FGD_GOLDEN = os.path.join(ROOT, "tests", "golden", "s2ag_fgd_golden.npz")
CFG = NS(**O.CFG)
LANG = NS(n_words=64, word_embedding_weights=None)


def make_evaluator(tmp_path, dev, seed):
    net = EmbeddingNet(CFG, 27, CFG.n_poses, 64, CFG.wordembed_dim, None, 'pose')
    sd = O.fill_state_dict(net.state_dict(), seed=seed)
    os.makedirs(tmp_path / "outputs", exist_ok=True)
    torch.save({'embedding_dict': sd}, tmp_path / "outputs" / "embedding_net.pth.tar")   # the reference's file schema
    return EmbeddingSpaceEvaluator(str(tmp_path), CFG, 27, LANG, dev)


def test_oracle_scores_match_reference_fixture():
    """the restatement reproduces the reference's get_scores on the reference's own features"""
    g = np.load(FGD_GOLDEN)
    fd, feat = FO.get_scores(g["gen_feat"], g["real_feat"])
    assert abs(fd - float(g["frechet"])) <= 1e-9 * max(1.0, abs(float(g["frechet"])))
    assert abs(feat - float(g["feat_dist"])) <= 1e-9


def test_state_dict_keys_match_reference():
    g = np.load(FGD_GOLDEN)
    net = EmbeddingNet(CFG, 27, CFG.n_poses, 64, CFG.wordembed_dim, None, 'pose')
    assert sorted(net.state_dict().keys()) == list(g["keys"])


def test_evaluator_matches_reference_fixture(dev, tmp_path):
    g = np.load(FGD_GOLDEN)
    ev = make_evaluator(tmp_path, dev, int(g["weight_seed"]))
    real, gen = FO.synthetic_pairs(int(g["pair_seed"]), int(g["n_batches"]), int(g["batch"]))
    nb = int(g["batch"])
    for i, (r, q) in enumerate(zip(real, gen)):
        rt, qt = torch.from_numpy(r).to(dev), torch.from_numpy(q).to(dev)
        ev.push_samples(None, None, qt, rt)
        with torch.no_grad():
            _, _, _, feat, mu, _, recon = ev.net(None, None, rt[:, :CFG.n_pre_poses], qt, 'pose')
        assert rel(feat, torch.from_numpy(g["gen_feat"][i * nb:(i + 1) * nb])) < 1e-3
        if i == 0:
            assert rel(recon[:6], torch.from_numpy(g["gen_recon_b0"])) < 1e-3
    assert ev.get_no_of_samples() == int(g["n_batches"])
    fd, feat_dist = ev.get_scores()
    assert abs(fd - float(g["frechet"])) <= 1e-3 * abs(float(g["frechet"])), (fd, float(g["frechet"]))
    assert abs(feat_dist - float(g["feat_dist"])) <= 1e-3 * float(g["feat_dist"])
    assert abs(ev.recon_err_diff - float(np.mean(g["recon_err_diff"]))) <= 1e-3 * abs(float(np.mean(g["recon_err_diff"]))) + 1e-6
    ev.reset()
    assert ev.get_no_of_samples() == 0 and float(ev.acc.abs().sum()) == 0.0


@pytest.mark.parametrize("n,d,kind", [(1000, 32, "iid"), (257, 32, "correlated"), (20, 32, "rank_deficient"),
                                      (300, 7, "small_d"), (33, 32, "shifted")])
def test_scores_vs_oracle_on_identical_features(dev, n, d, kind):
    """the moment accumulation (several ragged pushes) + Jacobi route against np.cov + scipy.linalg.sqrtm"""
    rng = np.random.RandomState(n * 31 + d)
    real = rng.normal(0, 1, size=(n, d))
    gen = rng.normal(0.1, 1.2, size=(n, d))
    if kind == "correlated":
        mix = rng.normal(0, 1, size=(d, d))
        real, gen = real @ mix, gen @ (mix + 0.1 * rng.normal(0, 1, size=(d, d)))
    if kind == "shifted":
        real, gen = real + 50.0, gen + 49.0   # large means: the second-moment route must not cancel
    real, gen = real.astype(np.float32), gen.astype(np.float32)
    acc = ops.fgd_new_accumulator(dev, d)
    cuts = [0, n // 3, n // 3 + 1, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        ops.fgd_accumulate(acc, torch.from_numpy(gen[a:b]).to(dev), torch.from_numpy(real[a:b]).to(dev))
    fd, feat = ops.fgd_scores(acc, d).tolist()
    o_fd, o_feat = FO.get_scores(gen.astype(np.float64), real.astype(np.float64))
    scale = float(np.trace(np.cov(gen, rowvar=False)) + np.trace(np.cov(real, rowvar=False)))
    tol = (1e-6 if kind != "rank_deficient" else 1e-4) * scale   # scipy's sqrtm itself is inexact on singular products
    assert abs(fd - o_fd) <= tol, (fd, o_fd, scale)
    assert abs(feat - o_feat) <= 1e-9 * o_feat


def test_calculate_frechet_distance_static(dev):
    rng = np.random.RandomState(5)
    a, b = rng.normal(0, 1, size=(200, 32)), rng.normal(0.3, 0.8, size=(200, 32))
    m1, s1, m2, s2 = a.mean(0), np.cov(a, rowvar=False), b.mean(0), np.cov(b, rowvar=False)
    got = EmbeddingSpaceEvaluator.calculate_frechet_distance(m1, s1, m2, s2, device=dev)
    want = FO.frechet_distance(m1, s1, m2, s2)
    assert abs(got - want) <= 1e-8 * abs(want)
    same = EmbeddingSpaceEvaluator.calculate_frechet_distance(m1, s1, m1, s1, device=dev)
    assert abs(same) <= 1e-9 * np.trace(s1)


def test_fgd_argument_errors(dev):
    from speech2affective_gestures_b200 import _C
    acc = ops.fgd_new_accumulator(dev, 32)
    with pytest.raises(_C.S2agError):
        ops.fgd_scores(acc, 33)   # D > 32 is refused, not truncated


def test_generate_gestures_reports_fgd(dev, tmp_path):
    """Processor.generate_gestures (processor_v2.py:1071-1142) with the embedding-net checkpoint present: the loss_dict
    carries frechet / feat_dist for both generators, equal to the oracle's scores on the clips it generated"""
    from test_processor_api import make_processor
    net = EmbeddingNet(CFG, 27, CFG.n_poses, 40, CFG.wordembed_dim, None, 'pose')
    sd = O.fill_state_dict(net.state_dict(), seed=31)
    os.makedirs(tmp_path / "outputs", exist_ok=True)
    torch.save({'embedding_dict': sd}, tmp_path / "outputs" / "embedding_net.pth.tar")
    pr, c = make_processor(dev, str(tmp_path))
    assert pr.evaluator is not None and pr.evaluator_trimodal is not None
    np.random.seed(11)
    torch.manual_seed(11)
    r = pr.generate_gestures(samples_to_generate=8, randomized=False, load_saved_model=False)
    assert {'frechet', 'feat_dist', 'frechet_trimodal', 'feat_dist_trimodal'} <= set(r)
    vec = torch.from_numpy(np.asarray(pr.data_loader['test_data_s2ag'].samples['vec_seq'][:8], dtype=np.float32))
    sdo = {k: v.detach().cpu() for k, v in sd.items()}
    with torch.no_grad():
        for key, out in (('', pr.last_out), ('_trimodal', pr.last_out_trimodal)):
            gf = FO.pose_encoder(sdo, out.detach().cpu().float()).numpy().astype(np.float64)
            rf = FO.pose_encoder(sdo, vec).numpy().astype(np.float64)
            fd, feat = FO.get_scores(gf, rf)
            assert abs(r['feat_dist' + key] - feat) <= 1e-3 * feat
            # 8 samples in 32 dimensions: a rank-7 covariance product, compare at the scale of the traces
            scale = float(np.trace(np.cov(gf, rowvar=False)) + np.trace(np.cov(rf, rowvar=False)))
            assert abs(r['frechet' + key] - fd) <= 2e-3 * scale, (r['frechet' + key], fd, scale)
