"""B200-native drop-in for ``scripts/model/embedding_space_evaluator.py`` (the FGD / diversity evaluator that
``evaluate_testset`` feeds every epoch, scripts/train_expressive.py:394-628).

Same class name, constructor signature and methods.  The feature extractor (``MotionAE`` for TED-Expressive, pose_dim
126) runs on the CUDA kernels of this package, the per-batch reconstruction / cosine side metrics in one kernel
(csrc/losses.cu::recon_metrics_kernel); only the [n, latent_dim] features are copied to the host, where the Frechet
distance (mean / covariance / matrix square root of two 128-d Gaussians) is evaluated in float64 exactly like the
reference does (numpy + scipy.linalg.sqrtm) -- that part is O(latent_dim^3), not a GPU problem.

The TED-Gesture evaluator (pose_dim 27) encodes with ``EmbeddingNet``, a model of the baseline ``multimodal_context``
family; it is not built here and raises.
"""
from __future__ import annotations

import numpy as np
import torch

from ..ops import _c, _call, _chk, _p, _st
from .motion_ae import MotionAE


class EmbeddingSpaceEvaluator:
    def __init__(self, args, embed_net_path, lang_model, device, ckpt=None):
        """``ckpt``: optional already-loaded checkpoint dict (tests); otherwise ``embed_net_path`` is torch.load-ed."""
        self.n_pre_poses = args.n_pre_poses
        if ckpt is None:
            ckpt = torch.load(embed_net_path, map_location=device, weights_only=False)
        self.pose_dim = ckpt["pose_dim"]
        if args.pose_dim == 126:
            self.latent_dim = ckpt["latent_dim"]
            self.net = MotionAE(self.pose_dim, self.latent_dim).to(device)
            self.net.load_state_dict(ckpt["motion_ae"])
        else:
            raise NotImplementedError("the TED-Gesture FGD evaluator encodes with EmbeddingNet (baseline multimodal_context "
                                      "family, scripts/model/embedding_net.py); only the TED-Expressive MotionAE path is built")
        self.net.train(False)
        self.reset()

    def reset(self):
        self.context_feat_list = []
        self.real_feat_list = []
        self.generated_feat_list = []
        self.recon_err_diff = []
        self.cos_err_diff = []

    def get_no_of_samples(self):
        return len(self.real_feat_list)

    @torch.no_grad()
    def push_samples(self, context_text, context_spec, generated_poses, real_poses):
        """embedding_space_evaluator.py:57-102 (pose_dim == 126 branch)."""
        real_poses = _c(real_poses.float())
        generated_poses = _c(generated_poses.float())
        real_recon, real_feat = self.net(real_poses)
        generated_recon, generated_feat = self.net(generated_poses)
        self.real_feat_list.append(real_feat.detach().cpu().numpy())
        self.generated_feat_list.append(generated_feat.detach().cpu().numpy())
        B, T, D = real_poses.shape
        out = torch.empty((4, B), device=real_poses.device, dtype=torch.float32)
        for k, (rec, pose) in enumerate(((real_recon, real_poses), (generated_recon, generated_poses))):
            rec = _c(rec)
            _chk(rec, pose)
            _call("ha2g_recon_metrics", _p(rec), _p(pose), B, T, D, _p(out[2 * k]), _p(out[2 * k + 1]), _st())
        sums = out.sum(dim=1)   # [rec_real, cos_real, rec_fake, cos_fake] summed over the batch (:88,:89,:98,:99)
        self.recon_err_diff.append(sums[2] - sums[0])
        self.cos_err_diff.append(sums[3] - sums[1])

    def get_features_for_viz(self):
        raise NotImplementedError("UMAP visualisation (embedding_space_evaluator.py:104-113) is outside the hot path")

    def get_diversity_scores(self):
        """:115-126 (uses torch.randperm on the host like the reference)."""
        feat1 = np.vstack(self.generated_feat_list[:500])
        random_idx = torch.randperm(len(self.generated_feat_list))[:500]
        shuffle_list = [self.generated_feat_list[x] for x in random_idx]
        feat2 = np.vstack(shuffle_list)
        return np.mean(np.sum(np.absolute(feat1 - feat2), axis=-1))

    def get_scores(self):
        """:128-158 -> (frechet_dist, feat_dist)."""
        generated_feats = np.vstack(self.generated_feat_list)
        real_feats = np.vstack(self.real_feat_list)

        def frechet_distance(samples_A, samples_B):
            A_mu = np.mean(samples_A, axis=0)
            A_sigma = np.cov(samples_A, rowvar=False)
            B_mu = np.mean(samples_B, axis=0)
            B_sigma = np.cov(samples_B, rowvar=False)
            try:
                return self.calculate_frechet_distance(A_mu, A_sigma, B_mu, B_sigma)
            except ValueError:
                return 1e+10

        frechet_dist = frechet_distance(generated_feats, real_feats)
        feat_dist = float(np.mean(np.sum(np.absolute(real_feats - generated_feats), axis=1)))
        return frechet_dist, feat_dist

    @staticmethod
    def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
        """d^2 = ||mu_1 - mu_2||^2 + Tr(C_1 + C_2 - 2 sqrt(C_1 C_2))   (:160-209, the pytorch-fid formulation)."""
        from scipy import linalg
        mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
        sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
        assert mu1.shape == mu2.shape, "Training and test mean vectors have different lengths"
        assert sigma1.shape == sigma2.shape, "Training and test covariances have different dimensions"
        diff = mu1 - mu2
        try:
            covmean, _ = linalg.sqrtm(sigma1.dot(sigma2), disp=False)
        except TypeError:   # scipy >= 1.16 dropped `disp` and returns the square root alone
            covmean = linalg.sqrtm(sigma1.dot(sigma2))
        if not np.isfinite(covmean).all():
            offset = np.eye(sigma1.shape[0]) * eps
            covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
        if np.iscomplexobj(covmean):
            if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
                raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
            covmean = covmean.real
        return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)
