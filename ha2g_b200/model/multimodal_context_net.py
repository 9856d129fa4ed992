"""B200-native drop-in for the baseline ``scripts/model/multimodal_context_net.py`` (Trimodal-context model: the 1-level
ancestor of the hierarchy path, trained by ``train_eval/train_gan.py``).

Same class names, constructor / ``forward`` signatures, return values and ``state_dict`` keys as the reference:

  WavEncoder          multimodal_context_net.py:9-28    4 x Conv1d(k = 15, strides 5/6/6/6, padding 1600) on RAW 16 kHz audio
  TextEncoderTCN      :31-61                            (returns ``(features, 0)`` here, unlike the hierarchy file)
  PoseGenerator       :64-160
  ConvDiscriminator   :207-252

Every arithmetic op runs on the kernels of this package (ha2g_b200.ops); CUDA tensors only.  ``nn.LeakyReLU(True)`` in
the reference (generator head :102, discriminator :214,217) is negative_slope = 1.0, i.e. the identity: no kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_LRELU03, ACT_NONE, ACT_SIGMOID
from . import vocab
from .hierarchy_net import TextEncoderTCN as _HierTextEncoderTCN
from .hierarchy_net import _BNP, _Conv1dP, _EmbeddingP, _GRUP, _LinearP, _Slot


class WavEncoder(nn.Module):
    """forward(wav_data [B, 36267]) -> [B, 34, 32]   (36267 + 2*1600 samples -> 7891 -> 1313 -> 217 -> 34 frames)."""

    def __init__(self):
        super().__init__()
        self.feat_extractor = nn.ModuleList([_Conv1dP(1, 16, 15), _BNP(16), _Slot(), _Conv1dP(16, 32, 15), _BNP(32), _Slot(),
                                             _Conv1dP(32, 64, 15), _BNP(64), _Slot(), _Conv1dP(64, 32, 15)])

    def forward(self, wav_data):
        fe = self.feat_extractor
        x = wav_data.unsqueeze(2)                                              # channels-last [B, n, 1]
        x = fe[1](ops.conv1d(x, fe[0].weight, fe[0].bias, 5, 1600), post_act=ACT_LRELU03)
        x = fe[4](ops.conv1d(x, fe[3].weight, fe[3].bias, 6, 0), post_act=ACT_LRELU03)
        x = fe[7](ops.conv1d(x, fe[6].weight, fe[6].bias, 6, 0), post_act=ACT_LRELU03)
        return ops.conv1d(x, fe[9].weight, fe[9].bias, 6, 0)                   # (batch x seq x dim), no transpose needed


class TextEncoderTCN(_HierTextEncoderTCN):
    """multimodal_context_net.py:31-61: identical network, but forward returns ``(features, 0)``."""

    def forward(self, input):
        return super().forward(input), 0


class PoseGenerator(nn.Module):
    """multimodal_context_net.py:64-160.  forward(pre_seq, in_text, in_audio, vid_indices) -> (out, z, mu, logvar)."""

    def __init__(self, args, pose_dim, n_words, word_embed_size, word_embeddings, z_obj=None):
        super().__init__()
        self.pre_length = args.n_pre_poses
        self.gen_length = args.n_poses - args.n_pre_poses
        self.z_obj = z_obj
        self.input_context = args.input_context
        if self.input_context != "both":
            raise NotImplementedError("the multimodal_context config uses input_context='both' (config/multimodal_context.yml)")
        self.in_size = 32 + 32 + pose_dim + 1
        self.audio_encoder = WavEncoder()
        self.text_encoder = TextEncoderTCN(args, n_words, word_embed_size, pre_trained_embedding=word_embeddings,
                                           dropout=args.dropout_prob)
        if not vocab.is_vocab(z_obj):
            raise NotImplementedError("z_type must be 'speaker' (z_obj = speaker Vocab), as in config/multimodal_context.yml")
        self.z_size = 16
        self.in_size += self.z_size
        self.speaker_embedding = nn.ModuleList([_EmbeddingP(z_obj.n_words, self.z_size), _LinearP(self.z_size, self.z_size)])
        self.speaker_mu = _LinearP(self.z_size, self.z_size)
        self.speaker_logvar = _LinearP(self.z_size, self.z_size)
        self.hidden_size = args.hidden_size
        self.gru = _GRUP(self.in_size, self.hidden_size, args.n_layers, args.dropout_prob)
        self.out = nn.ModuleList([_LinearP(self.hidden_size, self.hidden_size // 2), _Slot(),
                                  _LinearP(self.hidden_size // 2, pose_dim)])
        self.do_flatten_parameters = False

    def forward(self, pre_seq, in_text, in_audio, vid_indices=None, _eps=None):
        audio_feat_seq = self.audio_encoder(in_audio)
        text_feat_seq, _ = self.text_encoder(in_text)
        assert audio_feat_seq.shape[1] == text_feat_seq.shape[1]
        assert vid_indices is not None
        z_context = self.speaker_embedding[1](self.speaker_embedding[0](vid_indices))
        z_mu = self.speaker_mu(z_context)
        z_logvar = self.speaker_logvar(z_context)
        z_context = ops.reparameterize(z_mu, z_logvar, _eps)
        in_data = ops.concat_seq(pre_seq, audio_feat_seq, text_feat_seq, z_context)
        output, _ = self.gru(in_data, None, sum_dirs=True)
        h = self.out[0](output.reshape(-1, output.shape[2]))          # nn.LeakyReLU(True): slope 1.0 = identity
        o = self.out[2](h)
        return o.reshape(in_data.shape[0], in_data.shape[1], -1), z_context, z_mu, z_logvar


class ConvDiscriminator(nn.Module):
    """multimodal_context_net.py:207-252.  forward(poses [B,34,d], in_text=None) -> [B,1] in (0,1)."""

    def __init__(self, input_size):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = 64
        self.pre_conv = nn.ModuleList([_Conv1dP(input_size, 16, 3), _BNP(16), _Slot(), _Conv1dP(16, 8, 3), _BNP(8), _Slot(),
                                       _Conv1dP(8, 8, 3)])
        self.gru = _GRUP(8, self.hidden_size, 4, 0.3)
        self.out = _LinearP(self.hidden_size, 1)
        self.out2 = _LinearP(28, 1)
        self.do_flatten_parameters = False

    def forward(self, poses, in_text=None):
        pc = self.pre_conv
        x = pc[1](ops.conv1d_valid(poses, pc[0].weight, pc[0].bias))          # BN; LeakyReLU(True) = identity
        x = pc[4](ops.conv1d_valid(x, pc[3].weight, pc[3].bias))
        feat = ops.conv1d_valid(x, pc[6].weight, pc[6].bias)
        output, _ = self.gru(feat, None, sum_dirs=True)
        batch_size = poses.shape[0]
        o = self.out(output.reshape(-1, output.shape[2])).reshape(batch_size, -1)
        return self.out2(o, ACT_SIGMOID)
