"""ResNetSE-34 hierarchical audio encoder on channels-last (NHWC) activations.

Follows scripts/model/ResNetSE34V2.py:13-218 (module layout, parameter names, forward) and
scripts/model/ResNetBlocks.py:7-37,81-96 (SEBasicBlock / SELayer), with every op a kernel from
csrc/conv2d.cu, csrc/bn.cu, csrc/audio.cu.  The spectrogram (B,128,70) is treated as an NHWC image
with H = 128 mel bins, W = 70 frames, C = 1; all feature maps stay NHWC so the heads'
``reshape(B, C*F, T).transpose(1, 2)`` becomes one index-remap kernel instead of strided copies.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops, ops_audio
from ..ops import ACT_ELU, ACT_NONE, ACT_RELU
from . import vocab


class _Conv2dP(nn.Module):
    def __init__(self, cin, cout, k, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")  # ResNetSE34V2.py:89-91
        if bias:
            bound = 1 / math.sqrt(cin * k * k)
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)
        self.stride, self.padding = stride, padding

    def forward(self, x):
        return ops_audio.conv2d(x, self.weight, self.bias, self.stride, self.padding)


def _bn(c):
    from .hierarchy_net import _BNP
    return _BNP(c)


def _linear(fin, fout):
    from .hierarchy_net import _LinearP
    return _LinearP(fin, fout)


class _SELayer(nn.Module):
    def __init__(self, channel, reduction=8):
        super().__init__()
        from .hierarchy_net import _Slot
        self.fc = nn.ModuleList([_linear(channel, channel // reduction), _Slot(), _linear(channel // reduction, channel), _Slot()])


class SEBasicBlock(nn.Module):
    """conv1 -> ReLU -> BN1 -> conv2 -> BN2 -> SE -> (+ residual) -> ReLU   (ResNetBlocks.py:21-37)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, reduction=8):
        super().__init__()
        self.conv1 = _Conv2dP(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = _Conv2dP(planes, planes, 3, padding=1, bias=False)
        self.bn2 = _bn(planes)
        self.se = _SELayer(planes, reduction)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        out = self.bn1(self.conv1(x), pre_relu=True)
        out = self.bn2(self.conv2(out))
        residual = x
        if self.downsample is not None:
            residual = self.downsample[1](self.downsample[0](x))
        fc = self.se.fc
        return ops_audio.se_residual_relu(out, residual, fc[0].weight, fc[0].bias, fc[2].weight, fc[2].bias)


class ResNetSE(nn.Module):
    def __init__(self, args, layers, num_filters, nOut, z_obj, pose_level=3):
        super().__init__()
        from .hierarchy_net import _EmbeddingP
        self.pose_level = pose_level
        self.inplanes = num_filters[0]
        self.z_obj = z_obj
        self.conv1 = _Conv2dP(1, num_filters[0], 3, stride=1, padding=1)
        self.bn1 = _bn(num_filters[0])
        self.conv_low = _Conv2dP(64, 64, 2)
        self.bn_low = _bn(64)
        self.fc_low = _linear(63 * 64, nOut)
        self.conv_mid = _Conv2dP(32, 32, 3)
        self.bn_mid = _bn(32)
        self.fc_mid = _linear(62 * 32, nOut)
        self.conv_high = _Conv2dP(16, 16, 3)
        self.bn_high = _bn(16)
        self.fc_high = _linear(62 * 16, nOut)
        self.layer1 = self._make_layer(num_filters[0], layers[0])
        self.layer2 = self._make_layer(num_filters[1], layers[1], stride=2)
        self.layer3 = self._make_layer(num_filters[2], layers[2], stride=2)
        self.layer4 = self._make_layer(num_filters[3], layers[3], stride=2)
        if not vocab.is_vocab(z_obj):
            raise NotImplementedError("the hierarchy audio encoder is speaker-conditioned (z_obj = speaker Vocab)")
        self.speaker_embedding = nn.ModuleList([_EmbeddingP(z_obj.n_words, 16), _linear(16, 16)])
        self.fc1 = _linear(16, 32)
        self.fc2 = _linear(32, self.pose_level * 3)

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.ModuleList([_Conv2dP(self.inplanes, planes, 1, stride=stride, bias=False), _bn(planes)])
        layers = [SEBasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(SEBasicBlock(self.inplanes, planes))
        return nn.ModuleList(layers)

    def _head(self, feat, conv, bn, fc, shuffle):
        if shuffle > 1:
            feat = ops_audio.pixel_shuffle(feat, shuffle)
        f = bn(conv(feat), pre_relu=True)            # conv(+bias) -> ReLU -> BN   (ResNetSE34V2.py:157-159)
        return fc(ops_audio.head_flatten(f))          # [B, 34, nOut]

    def forward(self, x, vid_indices):
        if x.dim() == 4:  # the reference wrapper passes (B,1,128,70)
            x = x.squeeze(1)
        batch_size = x.shape[0]
        x = ops_audio.stem_conv(x, self.conv1.weight, self.conv1.bias)
        x = self.bn1(x, pre_relu=True)
        for blk in self.layer1:
            x = blk(x)
        feat1 = x
        for blk in self.layer2:
            feat1 = blk(feat1)
        feat2 = feat1
        for blk in self.layer3:
            feat2 = blk(feat2)
        feat3 = feat2
        for blk in self.layer4:
            feat3 = blk(feat3)
        feat_low = self._head(feat1, self.conv_low, self.bn_low, self.fc_low, 1)
        feat_mid = self._head(feat2, self.conv_mid, self.bn_mid, self.fc_mid, 2)
        feat_high = self._head(feat3, self.conv_high, self.bn_high, self.fc_high, 4)
        assert vid_indices is not None
        z_context = self.speaker_embedding[1](self.speaker_embedding[0](vid_indices))
        h = ops.act(z_context, ACT_ELU)
        h = self.fc1(h, ACT_ELU)
        logits = self.fc2(h)                          # [B, 3*L] == reshape(B, 3, L)
        weight, blend = ops_audio.speaker_blend(logits, feat_low, feat_mid, feat_high, self.pose_level)
        return weight, feat_low, feat_mid, feat_high, list(blend.unbind(0))
