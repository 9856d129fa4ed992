"""B200-native drop-in for the reference's ``scripts/model/hierarchy_net.py``.

Same four class names, constructor / ``forward`` signatures, return values and ``state_dict``
key names + shapes (so reference checkpoints load and ours are readable by the reference):

  Hierarchical_WavEncoder      hierarchy_net.py:10-19   (ResNetSE-34: ResNetSE34V2.py:13-218, ResNetBlocks.py)
  TextEncoderTCN               hierarchy_net.py:22-52   (tcn.py:16-64)
  Hierarchical_PoseGenerator   hierarchy_net.py:55-149
  Hierarchical_ConvDiscriminator hierarchy_net.py:197-242

The modules own ordinary ``nn.Parameter``s (callers build ``torch.optim.Adam`` over them), but every
forward/backward arithmetic op is a hand-written sm_100a kernel reached through ha2g_b200.ops.
They run on CUDA only; calling them with CPU tensors raises.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_ELU, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID
from . import vocab
from . import audio_encoder as _audio


# ------------------------------------------------------------------------------------------------
# parameter holders (names chosen so state_dict keys equal the reference's)
# ------------------------------------------------------------------------------------------------
class _LinearP(nn.Module):
    def __init__(self, fin, fout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fout, fin))
        self.bias = nn.Parameter(torch.empty(fout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(fin)
        nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x, act=ACT_NONE):
        return ops.linear(x, self.weight, self.bias, act)


class _EmbeddingP(nn.Module):
    def __init__(self, n, dim, weight=None, freeze=False):
        super().__init__()
        if weight is not None:
            self.weight = nn.Parameter(torch.as_tensor(weight, dtype=torch.float32).clone(), requires_grad=not freeze)
        else:
            self.weight = nn.Parameter(torch.randn(n, dim))

    def forward(self, idx):
        return ops.embedding(self.weight, idx)


class _Slot(nn.Module):
    """Parameter-free placeholder keeping the reference's nn.Sequential child numbering."""

    def forward(self, x):
        return x


class _WNConv1dP(nn.Module):
    """weight_norm(nn.Conv1d(C, C, k)) parameters: bias, weight_g [O,1,1], weight_v [O,I,k] (tcn.py:19-20)."""

    def __init__(self, cin, cout, k):
        super().__init__()
        conv = nn.Conv1d(cin, cout, k)
        v = conv.weight.data
        self.bias = nn.Parameter(conv.bias.data.clone())
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(-1, 1, 1).clone())
        self.weight_v = nn.Parameter(v.clone())


class _Conv1dP(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        conv = nn.Conv1d(cin, cout, k)
        self.weight = nn.Parameter(conv.weight.data.clone())
        self.bias = nn.Parameter(conv.bias.data.clone())


class _BNP(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.eps, self.momentum = 1e-5, 0.1

    def forward(self, x, pre_relu=False, post_act=ACT_NONE):
        y = ops.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var, pre_relu, post_act,
                           self.training, self.eps, self.momentum)
        if self.training:
            self.num_batches_tracked += 1
        return y


class _GRUP(nn.Module):
    """nn.GRU(batch_first=True, bidirectional=True) parameters under the cuDNN names."""

    def __init__(self, input_size, hidden_size, num_layers, dropout):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers, self.dropout = input_size, hidden_size, num_layers, dropout
        self.bidirectional, self.batch_first = True, True
        k = 1.0 / math.sqrt(hidden_size)
        self._names = []
        for l in range(num_layers):
            fin = input_size if l == 0 else 2 * hidden_size
            for suf in ("", "_reverse"):
                for name, shape in ((f"weight_ih_l{l}{suf}", (3 * hidden_size, fin)),
                                    (f"weight_hh_l{l}{suf}", (3 * hidden_size, hidden_size)),
                                    (f"bias_ih_l{l}{suf}", (3 * hidden_size,)),
                                    (f"bias_hh_l{l}{suf}", (3 * hidden_size,))):
                    p = nn.Parameter(torch.empty(shape).uniform_(-k, k))
                    setattr(self, name, p)
                    self._names.append(name)

    def flatten_parameters(self):  # cuDNN artefact of the reference (hierarchy_net.py:101-102); nothing to do
        return None

    def weights(self):
        return [getattr(self, n) for n in self._names]

    def forward(self, x, hx=None, sum_dirs=False):
        if hx is not None:
            raise NotImplementedError("the HA2G path always starts from h0 = 0")
        y = ops.bigru(x, self.weights(), self.hidden_size, self.num_layers, self.dropout, self.training, sum_dirs)
        return y, None


# ------------------------------------------------------------------------------------------------
# TextEncoderTCN
# ------------------------------------------------------------------------------------------------
class _TemporalBlock(nn.Module):
    def __init__(self, c_in, c_out, k, dilation, dropout):
        super().__init__()
        assert c_in == c_out and k == 2, "HA2G uses 300->300, kernel_size 2 blocks (no downsample conv)"
        self.conv1 = _WNConv1dP(c_in, c_out, k)
        self.conv2 = _WNConv1dP(c_out, c_out, k)
        # the reference registers the two convs a second time inside `net` (tcn.py:31-32)
        self.net = nn.ModuleList([self.conv1, _Slot(), _Slot(), _Slot(), self.conv2, _Slot(), _Slot(), _Slot()])
        self.dilation, self.p = dilation, dropout

    def forward(self, x):  # x [B,T,C] channels-last
        y = x
        for conv in (self.conv1, self.conv2):
            w = ops.tcn_weight(conv.weight_g, conv.weight_v)
            y = ops.linear(ops.shift_concat(y, self.dilation), w, conv.bias, ACT_RELU)
            y = ops.dropout(y, self.p, self.training)
        return ops.add_act(y, x, ACT_RELU)


class _TemporalConvNet(nn.Module):
    def __init__(self, num_inputs, num_channels, kernel_size, dropout):
        super().__init__()
        blocks = []
        for i, c in enumerate(num_channels):
            cin = num_inputs if i == 0 else num_channels[i - 1]
            blocks.append(_TemporalBlock(cin, c, kernel_size, 2 ** i, dropout))
        self.network = nn.ModuleList(blocks)

    def forward(self, x):
        for b in self.network:
            x = b(x)
        return x


class TextEncoderTCN(nn.Module):
    """hierarchy_net.py:22-52.  forward(input[B,T] int64) -> [B,T,32]."""

    def __init__(self, args, n_words, embed_size=300, pre_trained_embedding=None, kernel_size=2, dropout=0.3,
                 emb_dropout=0.1):
        super().__init__()
        if pre_trained_embedding is not None:
            assert pre_trained_embedding.shape[0] == n_words
            assert pre_trained_embedding.shape[1] == embed_size
            self.embedding = _EmbeddingP(n_words, embed_size, pre_trained_embedding, freeze=args.freeze_wordembed)
        else:
            self.embedding = _EmbeddingP(n_words, embed_size)
        num_channels = [args.hidden_size] * args.n_layers
        self.tcn = _TemporalConvNet(embed_size, num_channels, kernel_size, dropout)
        self.decoder = _LinearP(num_channels[-1], 32)
        self.emb_dropout = emb_dropout
        self.init_weights()

    def init_weights(self):
        self.decoder.bias.data.fill_(0)
        self.decoder.weight.data.normal_(0, 0.01)

    def forward(self, input):
        emb = ops.dropout(self.embedding(input), self.emb_dropout, self.training)
        y = self.tcn(emb)  # channels-last throughout: no transposes needed
        return self.decoder(y).contiguous()


# ------------------------------------------------------------------------------------------------
# Hierarchical_PoseGenerator
# ------------------------------------------------------------------------------------------------
class Hierarchical_PoseGenerator(nn.Module):
    """hierarchy_net.py:55-149.  forward(pre_seq, in_text, audio_feat_seq, vid_indices)
    -> (out[B,T,pose_dim], z[B,16], mu, logvar)."""

    def __init__(self, args, pose_dim, n_words, word_embed_size, word_embeddings, z_obj=None):
        super().__init__()
        self.pre_length = args.n_pre_poses
        self.gen_length = args.n_poses - args.n_pre_poses
        self.z_obj = z_obj
        self.input_context = args.input_context
        if self.input_context != "both":
            raise NotImplementedError("the hierarchy configs use input_context='both' (config*/hierarchy.yml)")
        self.in_size = 32 + 32 + pose_dim + 1
        self.text_encoder = TextEncoderTCN(args, n_words, word_embed_size, pre_trained_embedding=word_embeddings,
                                           dropout=args.dropout_prob)
        if not vocab.is_vocab(z_obj):
            raise NotImplementedError("z_type must be 'speaker' (z_obj = speaker Vocab), as in the hierarchy configs")
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

    def forward(self, pre_seq, in_text, audio_feat_seq=None, vid_indices=None, _eps=None, _text_feat=None):
        # _eps / _text_feat (not part of the reference signature): reparameterisation noise drawn by the caller, and the
        # output of this generator's own text encoder computed ahead of time (the inference loop batches it over all
        # windows of a clip: it depends on the tokens only)
        text_feat_seq = self.text_encoder(in_text) if _text_feat is None else _text_feat
        assert audio_feat_seq.shape[1] == text_feat_seq.shape[1]
        assert vid_indices is not None
        z_context = self.speaker_embedding[1](self.speaker_embedding[0](vid_indices))
        z_mu = self.speaker_mu(z_context)
        z_logvar = self.speaker_logvar(z_context)
        z_context = ops.reparameterize(z_mu, z_logvar, _eps)
        in_data = ops.concat_seq(pre_seq, audio_feat_seq, text_feat_seq, z_context)
        output, _ = self.gru(in_data, None, sum_dirs=True)
        h = self.out[0](output.reshape(-1, output.shape[2]), ACT_LRELU)
        o = self.out[2](h)
        decoder_outputs = o.reshape(in_data.shape[0], in_data.shape[1], -1)
        return decoder_outputs, z_context, z_mu, z_logvar


# ------------------------------------------------------------------------------------------------
# Hierarchical_ConvDiscriminator
# ------------------------------------------------------------------------------------------------
class Hierarchical_ConvDiscriminator(nn.Module):
    """hierarchy_net.py:197-242.  forward(poses[B,34,d], in_text=None) -> [B,1] in (0,1)."""

    def __init__(self, input_size):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = 64
        self.pre_conv = nn.ModuleList([_Conv1dP(input_size, 16, 3), _BNP(16), _Slot(), _Conv1dP(16, 8, 3), _BNP(8),
                                       _Slot(), _Conv1dP(8, 8, 3)])
        self.gru = _GRUP(8, self.hidden_size, 4, 0.3)
        self.out = _LinearP(self.hidden_size, 1)
        self.out2 = _LinearP(28, 1)
        self.do_flatten_parameters = False

    def forward(self, poses, in_text=None):
        pc = self.pre_conv
        x = ops.conv1d_valid(poses, pc[0].weight, pc[0].bias)  # channels-last: no transpose
        x = pc[1](x, post_act=ACT_LRELU)
        x = ops.conv1d_valid(x, pc[3].weight, pc[3].bias)
        x = pc[4](x, post_act=ACT_LRELU)
        feat = ops.conv1d_valid(x, pc[6].weight, pc[6].bias)
        output, _ = self.gru(feat, None, sum_dirs=True)
        batch_size = poses.shape[0]
        o = self.out(output.reshape(-1, output.shape[2])).reshape(batch_size, -1)
        return self.out2(o, ACT_SIGMOID)


# ------------------------------------------------------------------------------------------------
# Hierarchical_WavEncoder
# ------------------------------------------------------------------------------------------------
class Hierarchical_WavEncoder(nn.Module):
    """hierarchy_net.py:10-19.  forward(audio_spectrum[B,128,70], vid_indices[B]) ->
    (weight[B,3,L], feat_low, feat_mid, feat_high [B,34,32], [L x [B,34,32]])."""

    def __init__(self, args, z_obj, pose_level, nOut=32):
        super().__init__()
        self.feat_extractor = _audio.ResNetSE(args, [3, 4, 6, 3], [32, 64, 128, 256], nOut=nOut, pose_level=pose_level,
                                              z_obj=z_obj)

    def forward(self, audio_spectrum, vid_indices):
        return self.feat_extractor(audio_spectrum, vid_indices)
