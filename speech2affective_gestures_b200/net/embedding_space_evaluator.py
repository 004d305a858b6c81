"""Drop-in for the reference's net/embedding_space_evaluator.py `EmbeddingSpaceEvaluator`: the Frechet
gesture distance and the paired feature distance between generated and real 34-frame clips in the latent space of the
pose auto-encoder (`outputs/embedding_net.pth.tar`, key 'embedding_dict').

Same constructor and methods as the reference (:16-101); what differs is where the numbers live.  The reference copies
every batch's features to the host and keeps them in Python lists (:53-56), then stacks them and runs np.mean / np.cov /
scipy.linalg.sqrtm (:73-152).  Here `push_samples` folds the batch into a fp64 moment buffer on the device
(csrc/fgd.cu: count, paired L1 sum, sums, second moments) and `get_scores` is one single-warp kernel (two Jacobi
eigen-decompositions) plus a 16-byte read-back; no feature ever leaves the GPU.
"""
from os.path import join as jn

import numpy as np
import torch

from .. import ops
from .embedding_net import EmbeddingNet

FEAT_DIM = 32   # PoseEncoderConv.fc_mu (net/embedding_net.py:61)


class EmbeddingSpaceEvaluator:
    def __init__(self, base_path, args, pose_dim, lang_model, device):
        self.n_pre_poses = args.n_pre_poses
        checkpoint = torch.load(jn(base_path, 'outputs/embedding_net.pth.tar'), map_location='cpu')
        self.pose_dim = pose_dim
        self.device = torch.device(device)
        self.net = EmbeddingNet(args, self.pose_dim, args.n_poses, lang_model.n_words, args.wordembed_dim,
                                lang_model.word_embedding_weights, 'pose').to(self.device)
        self.net.load_state_dict(checkpoint['embedding_dict'])
        self.net.train(False)
        self.reset()

    def reset(self):
        self.acc = ops.fgd_new_accumulator(self.device, FEAT_DIM)
        self.recon = torch.zeros(2, dtype=torch.float32, device=self.device)   # sums of the two reconstruction L1s
        self._recon_tmp = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.n_batches = 0

    def get_no_of_samples(self):
        """number of pushed BATCHES, like the reference's len(self.real_feat_list) (:42-43)"""
        return self.n_batches

    @torch.no_grad()
    def push_samples(self, context_text, context_spec, generated_poses, real_poses):
        """reference :45-61; the context arguments are unused in mode 'pose' (context_feat is None there too)"""
        pre_poses = real_poses[:, 0:self.n_pre_poses]
        _, _, _, real_feat, _, _, real_recon = self.net(context_text, context_spec, pre_poses, real_poses, 'pose',
                                                        variational_encoding=False)
        _, _, _, generated_feat, _, _, generated_recon = self.net(None, None, pre_poses, generated_poses, 'pose',
                                                                  variational_encoding=False)
        ops.fgd_accumulate(self.acc, generated_feat, real_feat)
        for slot, (a, b) in enumerate(((real_poses, real_recon), (generated_poses, generated_recon))):
            ops.l1_mean(a.contiguous(), b, self._recon_tmp)
            self.recon[slot:slot + 1] += self._recon_tmp
        self.n_batches += 1

    @property
    def recon_err_diff(self):
        """mean over batches of (recon_err_fake - recon_err_real): the reference keeps the per-batch list (:58-61)"""
        r = self.recon.tolist()
        return (r[1] - r[0]) / max(self.n_batches, 1)

    def get_features_for_viz(self):
        raise NotImplementedError("UMAP visualisation is outside the hot path (SURVEY 2: plotting / rendering)")

    def get_scores(self):
        """-> (frechet_dist, feat_dist) as Python floats (reference :73-101)"""
        fd, feat = ops.fgd_scores(self.acc, FEAT_DIM).tolist()
        if not np.isfinite(fd):
            fd = 1e+10   # the reference's ValueError branch (:84-87)
        return fd, feat

    @staticmethod
    def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6, device=None):
        """reference :104-152 for caller-given moments (numpy or tensors); evaluated on the device"""
        dev = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        ts = [torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t, dtype=torch.float64).to(dev)
              for t in (np.atleast_1d(mu1) if not torch.is_tensor(mu1) else mu1,
                        np.atleast_2d(sigma1) if not torch.is_tensor(sigma1) else sigma1,
                        np.atleast_1d(mu2) if not torch.is_tensor(mu2) else mu2,
                        np.atleast_2d(sigma2) if not torch.is_tensor(sigma2) else sigma2)]
        assert ts[0].shape == ts[2].shape, 'Training and test mean vectors have different lengths'
        assert ts[1].shape == ts[3].shape, 'Training and test covariances have different dimensions'
        return float(ops.frechet_distance(*ts)[0])
