"""Model / training hyper-parameters of the hot path as the reference ships them
(config/multimodal_context_v2.yml:15-46 plus the parse_args.py:39,58 defaults), for drivers that do not parse
the reference's YAML (bench.py, tools/)."""
from types import SimpleNamespace

S2AG_CONFIG = dict(
    n_pre_poses=4, n_poses=34, input_context='both', hidden_size=300, hidden_size_s2eg=300, n_layers=4,
    dropout_prob=0.3, freeze_wordembed=False, wordembed_dim=300, z_type='speaker', learning_rate=5e-4,
    discriminator_lr_weight=0.2, loss_regression_weight=500, loss_gan_weight=5.0, loss_warmup=0,
    loss_kld_weight=0.1, loss_reg_weight=0.05, motion_resampling_framerate=15, num_mfcc=14,
    mean_dir_vec=[0.0154009, -0.9690125, -0.0884354, -0.0022264, -0.8655276, 0.4342174, -0.0035145, -0.8755367, -0.4121039,
                  -0.9236511, 0.3061306, -0.0012415, -0.5155854, 0.8129665, 0.0871897, 0.2348464, 0.1846561, 0.8091402,
                  0.9271948, 0.2960011, -0.013189, 0.5233978, 0.8092403, 0.0725451, -0.2037076, 0.1924306, 0.8196916])


def namespace():
    return SimpleNamespace(**S2AG_CONFIG)
