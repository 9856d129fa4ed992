"""B200-native drop-in for the reference's FGD auto-encoder ``scripts/model/motion_ae.py`` (TED-Expressive evaluation:
``EmbeddingSpaceEvaluator`` encodes real and generated 34-frame clips with it, embedding_space_evaluator.py:29-33,67-69).

Same class names, constructor signatures and ``state_dict`` keys as the reference (``encoder.net.0.0.weight`` ...), so
the reference's ``motion_ae`` checkpoints load unchanged.  Inference only (the evaluator runs it under ``train(False)``):
every convolution is an unfold + GEMM, every BatchNorm1d the eval-mode kernel with the activation fused, on channels-last
[B, T, C] activations; ``nn.LeakyReLU(True)`` in the reference means negative_slope = 1.0, i.e. the identity.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_LRELU02, ACT_NONE, ACT_RELU
from .hierarchy_net import _BNP, _Conv1dP, _LinearP, _Slot


def _eval_only(m: nn.Module):
    if m.training:
        raise NotImplementedError("MotionAE here is the evaluator's inference path (train(False)); training it is the "
                                  "reference's train_feature_extractor_expressive.py, outside the hot path")


class _ConvTranspose1dP(nn.Module):
    """nn.ConvTranspose1d(cin, cout, k) parameters: weight [cin, cout, k], bias [cout]."""

    def __init__(self, cin, cout, k):
        super().__init__()
        ref = nn.ConvTranspose1d(cin, cout, k)
        self.weight = nn.Parameter(ref.weight.data.clone())
        self.bias = nn.Parameter(ref.bias.data.clone())

    def forward(self, x):
        # y[t] = sum_k x[t-k] W[:, :, k]  ==  valid Conv1d over x zero-padded by k-1 frames on both sides with the
        # flipped, (in,out)-transposed kernel
        k = self.weight.shape[2]
        B, T, C = x.shape
        xp = torch.zeros((B, T + 2 * (k - 1), C), device=x.device, dtype=torch.float32)
        xp[:, k - 1:k - 1 + T] = x
        w = self.weight.flip(2).permute(1, 0, 2).contiguous()     # [cout, cin, k]: a Conv1d weight
        return ops.conv1d_valid(xp, w, self.bias)


def _conv_norm_relu(cin, cout, downsample=False):
    """ConvNormRelu (motion_ae.py:8-31): Conv1d(k3 s1 | k4 s2) -> BatchNorm1d -> LeakyReLU(0.2)."""
    return nn.ModuleList([_Conv1dP(cin, cout, 4 if downsample else 3), _BNP(cout), _Slot()])


def _run_cnr(block, x, downsample=False):
    y = ops.conv1d_valid(x, block[0].weight, block[0].bias)
    if downsample:   # stride 2 = every second frame of the stride-1 result
        y = y[:, ::2].contiguous()
    return block[1](y, post_act=ACT_LRELU02)


class PoseEncoderConv(nn.Module):
    """motion_ae.py:33-62."""

    def __init__(self, length, pose_dim, latent_dim):
        super().__init__()
        if length != 34:
            raise NotImplementedError("the HA2G configs use 34-frame clips")
        self.net = nn.ModuleList([_conv_norm_relu(pose_dim, 32), _conv_norm_relu(32, 64), _conv_norm_relu(64, 64, True),
                                  _Conv1dP(64, 32, 3)])
        self.out_net = nn.ModuleList([_LinearP(384, 256), _BNP(256), _Slot(), _LinearP(256, 128), _BNP(128), _Slot(),
                                      _LinearP(128, latent_dim)])

    def forward(self, poses):
        _eval_only(self)
        x = _run_cnr(self.net[0], poses)
        x = _run_cnr(self.net[1], x)
        x = _run_cnr(self.net[2], x, downsample=True)
        x = ops.conv1d_valid(x, self.net[3].weight, self.net[3].bias)         # [B, 12, 32] channels-last
        B, T, C = x.shape
        # the reference flattens (B, C, T): feature c*T + t.  Channels-last holds t*C + c, so the first Linear reads its
        # weight columns through that permutation instead of transposing the activations.
        perm = (torch.arange(C, device=x.device).view(1, C) * T + torch.arange(T, device=x.device).view(T, 1)).reshape(-1)
        w0 = ops.gather_cols(self.out_net[0].weight, perm.to(torch.int32))
        h = ops.linear(x.reshape(B, T * C), w0, self.out_net[0].bias)
        h = self.out_net[1](h)                                                # BN eval; LeakyReLU(True) = identity
        h = self.out_net[4](self.out_net[3](h))
        return self.out_net[6](h)


class PoseDecoderConv(nn.Module):
    """motion_ae.py:64-116 (use_pre_poses=False, as MotionAE builds it)."""

    def __init__(self, length, pose_dim, latent_dim, use_pre_poses=False):
        super().__init__()
        if use_pre_poses or length != 34:
            raise NotImplementedError("MotionAE uses PoseDecoderConv(34, pose_dim, latent_dim) without pre-poses")
        self.use_pre_poses = False
        self.pre_net = nn.ModuleList([_LinearP(latent_dim, 64), _BNP(64), _Slot(), _LinearP(64, 136)])
        self.net = nn.ModuleList([_ConvTranspose1dP(4, 32, 3), _BNP(32), _Slot(), _ConvTranspose1dP(32, 32, 3), _BNP(32),
                                  _Slot(), _Conv1dP(32, 32, 3), _Conv1dP(32, pose_dim, 3)])

    def forward(self, feat, pre_poses=None):
        _eval_only(self)
        B = feat.shape[0]
        out = self.pre_net[3](self.pre_net[1](self.pre_net[0](feat)))        # [B, 136] == view(B, 4, 34) channels-first
        idx = (torch.arange(4, device=feat.device).view(1, 4) * 34 + torch.arange(34, device=feat.device).view(34, 1))
        x = ops.gather_cols(out, idx.reshape(-1).to(torch.int32)).reshape(B, 34, 4)   # channels-last [B, 34, 4]
        x = self.net[1](self.net[0](x), post_act=ACT_LRELU02)
        x = self.net[4](self.net[3](x), post_act=ACT_LRELU02)
        x = ops.conv1d_valid(x, self.net[6].weight, self.net[6].bias)
        return ops.conv1d_valid(x, self.net[7].weight, self.net[7].bias)      # [B, 34, pose_dim]


class MotionAE(nn.Module):
    """motion_ae.py:118-130.  forward(pose[B,34,pose_dim]) -> (reconstruction [B,34,pose_dim], z [B,latent_dim])."""

    def __init__(self, pose_dim, latent_dim):
        super().__init__()
        self.encoder = PoseEncoderConv(34, pose_dim, latent_dim)
        self.decoder = PoseDecoderConv(34, pose_dim, latent_dim)

    def forward(self, pose):
        pose = pose.reshape(pose.size(0), pose.size(1), -1).contiguous()
        z = self.encoder(pose)
        pred = self.decoder(z)
        return pred, z
