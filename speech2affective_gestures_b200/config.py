"""Model / training hyper-parameters of the hot path as the reference ships them
(config/multimodal_context_v2.yml:15-46 plus the parse_args.py:39,58 defaults), for drivers that do not parse
the reference's YAML (bench.py, tools/)."""
from types import SimpleNamespace

S2AG_CONFIG = dict(
    n_pre_poses=4, n_poses=34, input_context='both', hidden_size=300, hidden_size_s2eg=300, n_layers=4,
    dropout_prob=0.3, freeze_wordembed=False, wordembed_dim=300, z_type='speaker', learning_rate=5e-4,
    discriminator_lr_weight=0.2, loss_regression_weight=500, loss_gan_weight=5.0, loss_warmup=0,
    loss_kld_weight=0.1, loss_reg_weight=0.05)


def namespace():
    return SimpleNamespace(**S2AG_CONFIG)
