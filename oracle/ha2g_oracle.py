"""CPU oracle for the HA2G hierarchical training-step hot path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke() may import it.  The product path
(ha2g_b200/*) never routes through it.

It restates, as *functional* fp32 torch-CPU code over plain ``dict[str, Tensor]`` state dicts,
what the reference nn.Modules and step functions compute.  Every function cites the reference
file:line it follows (paths relative to /root/reference/scripts).  All randomness (reparameterize
noise, dropout masks, speaker permutation) is *injected* so that the CUDA path and the oracle can
be compared on identical draws.

Pinning: tests/golden/*.pt are produced by oracle/make_golden.py, which imports and runs the
UNMODIFIED reference modules / train_iter_* functions in the build container; tests/test_oracle_*.py
check this restatement against those fixtures (fwd, bwd, losses, post-Adam parameters).
Mel-spectrogram (row K) lives in oracle/mel_oracle.py and is "parity unpinned" (librosa absent).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------------------------
# cascade tables (bit-exact index contract)      train_hierarchy.py:86-88,
#                                                train_hierarchy_expressive.py:140-145, 252-310
# --------------------------------------------------------------------------------------------
_HEAD = [37, 38, 39, 40, 41]
BONE_LEVELS = {
    "gesture": [
        [0, 1, 2, 3, 6],
        [0, 1, 2, 3, 4, 6, 7],
        list(range(9)),
    ],
    "expressive": [
        [0, 1, 2] + _HEAD,
        [0, 1, 2, 3, 20] + _HEAD,
        [0, 1, 2, 3, 4, 20, 21] + _HEAD,
        [0, 1, 2, 3, 4, 5, 8, 11, 14, 17, 20, 21, 22, 25, 28, 31, 34] + _HEAD,
        [0, 1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 14, 15, 17, 18, 20, 21, 22, 23, 25, 26, 28, 29, 31, 32, 34, 35] + _HEAD,
        list(range(42)),
    ],
}


def level_channels(variant: str) -> List[np.ndarray]:
    """Channel indices (into the full pose vector) of each level's target, ascending bone order."""
    out = []
    for bones in BONE_LEVELS[variant]:
        out.append(np.array([3 * b + c for b in bones for c in range(3)], dtype=np.int64))
    return out


def cascade_maps(variant: str):
    """For level k>0: (dst columns in pre_seq_k, src columns in out_{k-1}) for frames >= n_pre_poses.

    Quirk kept from the reference: the head bones are copied with ``pre_seq_k[:, n_pre:, -5*3:] = out[:, n_pre:, -5*3:]``
    (train_hierarchy_expressive.py:164,171,...); pre_seq_k is one column wider than the pose (constraint flag
    last), so on the destination side that slice is shifted by +1: the head values land one column to the right
    (the last one in the flag column) and the first head column stays 0."""
    maps = [None]
    levels = BONE_LEVELS[variant]
    for k in range(1, len(levels)):
        prev, cur = levels[k - 1], levels[k]
        dst, src = [], []
        for s_prev, b in enumerate(prev):
            s_cur = cur.index(b)
            shift = 1 if (variant == "expressive" and b in _HEAD) else 0
            for c in range(3):
                dst.append(3 * s_cur + c + shift)
                src.append(3 * s_prev + c)
        maps.append((np.array(dst, dtype=np.int64), np.array(src, dtype=np.int64)))
    return maps


def make_pre_seq(target_k: Tensor, prev_out: Optional[Tensor], cmap, n_pre: int) -> Tensor:
    """train_hierarchy_expressive.py:252-262 (one level)."""
    B, T, d = target_k.shape
    pre = target_k.new_zeros((B, T, d + 1))
    pre[:, :n_pre, :-1] = target_k[:, :n_pre]
    pre[:, :n_pre, -1] = 1
    if prev_out is not None:
        dst, src = cmap
        pre[:, n_pre:, torch.as_tensor(dst)] = prev_out[:, n_pre:, torch.as_tensor(src)]
    return pre


# --------------------------------------------------------------------------------------------
# GRU (nn.GRU semantics, batch_first, bidirectional)            hierarchy_net.py:87-88,144-145
# --------------------------------------------------------------------------------------------
def gru_direction(x: Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> Tensor:
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi = x @ w_ih.t() + b_ih
    h = x.new_zeros((B, H))
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, t, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, dim=1)


def bigru_stack(x: Tensor, sd: SD, prefix: str, n_layers: int,
                drop_masks: Optional[Sequence[Tensor]] = None, p: float = 0.0) -> Tensor:
    """Returns (B,T,2H).  drop_masks[l] (B,T,2H) of {0,1} is applied (scaled 1/(1-p)) to the
    output of layer l for l < n_layers-1 (nn.GRU inter-layer dropout, training only)."""
    for l in range(n_layers):
        outs = []
        for suf, rev in (("", False), ("_reverse", True)):
            outs.append(gru_direction(x, sd[f"{prefix}weight_ih_l{l}{suf}"], sd[f"{prefix}weight_hh_l{l}{suf}"],
                                      sd[f"{prefix}bias_ih_l{l}{suf}"], sd[f"{prefix}bias_hh_l{l}{suf}"], rev))
        x = torch.cat(outs, dim=2)
        if drop_masks is not None and l < n_layers - 1:
            x = x * drop_masks[l] / (1.0 - p)
    return x


# --------------------------------------------------------------------------------------------
# TextEncoderTCN                                             hierarchy_net.py:22-52, tcn.py:16-64
# --------------------------------------------------------------------------------------------
def weight_norm_w(g: Tensor, v: Tensor) -> Tensor:
    """old-style torch.nn.utils.weight_norm, dim=0: w = g * v / ||v||_(in,k)."""
    return g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)


def text_encoder_tcn(tokens: Tensor, sd: SD, prefix: str = "", n_levels: int = 4,
                     emb_mask: Optional[Tensor] = None, emb_p: float = 0.1,
                     tcn_masks: Optional[Sequence[Tensor]] = None, tcn_p: float = 0.3) -> Tensor:
    """tokens (B,T) int64 -> (B,T,32).  emb_mask (B,T,E); tcn_masks[2*i+j] (B,C,T)."""
    emb = sd[prefix + "embedding.weight"][tokens]  # hierarchy_net.py:49
    if emb_mask is not None:
        emb = emb * emb_mask / (1.0 - emb_p)
    x = emb.transpose(1, 2)  # (B,C,T)
    T = x.shape[2]
    for i in range(n_levels):
        dil = 2 ** i
        res = x
        y = x
        for j in (1, 2):
            w = weight_norm_w(sd[f"{prefix}tcn.network.{i}.conv{j}.weight_g"], sd[f"{prefix}tcn.network.{i}.conv{j}.weight_v"])
            y = F.conv1d(y, w, sd[f"{prefix}tcn.network.{i}.conv{j}.bias"], padding=dil, dilation=dil)[:, :, :T]  # chomp
            y = torch.relu(y)
            if tcn_masks is not None:
                y = y * tcn_masks[2 * i + (j - 1)] / (1.0 - tcn_p)
        x = torch.relu(y + res)  # tcn.py:47 (no downsample: 300 -> 300)
    y = x.transpose(1, 2) @ sd[prefix + "decoder.weight"].t() + sd[prefix + "decoder.bias"]
    return y.contiguous()


# --------------------------------------------------------------------------------------------
# Hierarchical_PoseGenerator                                          hierarchy_net.py:99-149
# --------------------------------------------------------------------------------------------
def pose_generator(sd: SD, pre_seq: Tensor, in_text: Tensor, audio_feat: Tensor, vid: Tensor,
                   eps: Tensor, n_layers: int = 4, hidden: int = 300, rng: Optional[dict] = None):
    """rng: optional dict(emb_mask, tcn_masks, gru_masks, p) for train-mode dropout."""
    rng = rng or {}
    text_feat = text_encoder_tcn(in_text, sd, "text_encoder.", n_levels=n_layers,
                                 emb_mask=rng.get("emb_mask"), tcn_masks=rng.get("tcn_masks"),
                                 tcn_p=rng.get("p", 0.3))
    assert audio_feat.shape[1] == text_feat.shape[1]
    zc = sd["speaker_embedding.0.weight"][vid] @ sd["speaker_embedding.1.weight"].t() + sd["speaker_embedding.1.bias"]
    mu = zc @ sd["speaker_mu.weight"].t() + sd["speaker_mu.bias"]
    logvar = zc @ sd["speaker_logvar.weight"].t() + sd["speaker_logvar.bias"]
    z = mu + eps * torch.exp(0.5 * logvar)  # embedding_net.py:10-13
    T = pre_seq.shape[1]
    x = torch.cat((pre_seq, audio_feat, text_feat, z.unsqueeze(1).repeat(1, T, 1)), dim=2)
    y = bigru_stack(x, sd, "gru.", n_layers, rng.get("gru_masks"), rng.get("p", 0.3))
    y = y[:, :, :hidden] + y[:, :, hidden:]
    h = y.reshape(-1, hidden) @ sd["out.0.weight"].t() + sd["out.0.bias"]
    h = F.leaky_relu(h, 0.01)
    o = h @ sd["out.2.weight"].t() + sd["out.2.bias"]
    return o.reshape(pre_seq.shape[0], T, -1), z, mu, logvar


# --------------------------------------------------------------------------------------------
# Hierarchical_ConvDiscriminator                                      hierarchy_net.py:197-242
# --------------------------------------------------------------------------------------------
def _bn(x: Tensor, sd: SD, key: str, training: bool, stats_out: Optional[dict], dims, eps=1e-5, momentum=0.1):
    w, b = sd[key + ".weight"], sd[key + ".bias"]
    shape = [1, -1] + [1] * (x.dim() - 2)
    if training:
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        if stats_out is not None:
            n = x.numel() / x.shape[1]
            stats_out[key + ".running_mean"] = (1 - momentum) * sd[key + ".running_mean"] + momentum * mean.detach()
            stats_out[key + ".running_var"] = (1 - momentum) * sd[key + ".running_var"] + momentum * var.detach() * n / (n - 1)
            stats_out[key + ".num_batches_tracked"] = sd[key + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[key + ".running_mean"], sd[key + ".running_var"]
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps) * w.view(shape) + b.view(shape)


def conv_discriminator(sd: SD, poses: Tensor, training: bool = True, stats_out: Optional[dict] = None,
                       gru_masks=None, p: float = 0.3) -> Tensor:
    x = poses.transpose(1, 2)
    x = F.conv1d(x, sd["pre_conv.0.weight"], sd["pre_conv.0.bias"])
    x = F.leaky_relu(_bn(x, sd, "pre_conv.1", training, stats_out, (0, 2)), 0.01)
    x = F.conv1d(x, sd["pre_conv.3.weight"], sd["pre_conv.3.bias"])
    x = F.leaky_relu(_bn(x, sd, "pre_conv.4", training, stats_out, (0, 2)), 0.01)
    x = F.conv1d(x, sd["pre_conv.6.weight"], sd["pre_conv.6.bias"])
    x = x.transpose(1, 2)
    y = bigru_stack(x, sd, "gru.", 4, gru_masks, p)
    y = y[:, :, :64] + y[:, :, 64:]
    B = poses.shape[0]
    o = (y.reshape(-1, 64) @ sd["out.weight"].t() + sd["out.bias"]).view(B, -1)
    o = o @ sd["out2.weight"].t() + sd["out2.bias"]
    return torch.sigmoid(o)


# --------------------------------------------------------------------------------------------
# Hierarchical_WavEncoder = ResNetSE-34                 ResNetSE34V2.py:118-218, ResNetBlocks.py
# --------------------------------------------------------------------------------------------
def _se_block(x: Tensor, sd: SD, pre: str, stride: int, training: bool, stats_out):
    """conv1 -> ReLU -> BN1 -> conv2 -> BN2 -> SE -> (+res) -> ReLU   (ResNetBlocks.py:21-37)."""
    out = F.conv2d(x, sd[pre + "conv1.weight"], None, stride=stride, padding=1)
    out = _bn(torch.relu(out), sd, pre + "bn1", training, stats_out, (0, 2, 3))
    out = F.conv2d(out, sd[pre + "conv2.weight"], None, stride=1, padding=1)
    out = _bn(out, sd, pre + "bn2", training, stats_out, (0, 2, 3))
    y = out.mean(dim=(2, 3))
    y = torch.relu(y @ sd[pre + "se.fc.0.weight"].t() + sd[pre + "se.fc.0.bias"])
    y = torch.sigmoid(y @ sd[pre + "se.fc.2.weight"].t() + sd[pre + "se.fc.2.bias"])
    out = out * y[:, :, None, None]
    if (pre + "downsample.0.weight") in sd:
        res = F.conv2d(x, sd[pre + "downsample.0.weight"], None, stride=stride)
        res = _bn(res, sd, pre + "downsample.1", training, stats_out, (0, 2, 3))
    else:
        res = x
    return torch.relu(out + res)


def _head(feat: Tensor, sd: SD, pre: str, name: str, shuffle: int, training: bool, stats_out):
    """[PixelShuffle] -> conv(+bias) -> ReLU -> BN -> (B,C*F,T)^T -> Linear  (ResNetSE34V2.py:157-186)."""
    B = feat.shape[0]
    if shuffle > 1:
        feat = F.pixel_shuffle(feat, shuffle)
    f = F.conv2d(feat, sd[f"{pre}conv_{name}.weight"], sd[f"{pre}conv_{name}.bias"])
    f = _bn(torch.relu(f), sd, f"{pre}bn_{name}", training, stats_out, (0, 2, 3))
    f = f.reshape(B, -1, f.shape[-1]).transpose(1, 2)
    f = f.reshape(-1, f.shape[-1])
    o = f @ sd[f"{pre}fc_{name}.weight"].t() + sd[f"{pre}fc_{name}.bias"]
    return o.reshape(B, -1, o.shape[-1])


def wav_encoder(sd: SD, spec: Tensor, vid: Tensor, pose_level: int, training: bool = True,
                stats_out: Optional[dict] = None, layers=(3, 4, 6, 3)):
    pre = "feat_extractor."
    x = spec.unsqueeze(1)
    x = F.conv2d(x, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    x = _bn(torch.relu(x), sd, pre + "bn1", training, stats_out, (0, 2, 3))
    feats = []
    for li, nblk in enumerate(layers, start=1):
        for bi in range(nblk):
            stride = 2 if (li > 1 and bi == 0) else 1
            x = _se_block(x, sd, f"{pre}layer{li}.{bi}.", stride, training, stats_out)
        feats.append(x)
    feat_low = _head(feats[1], sd, pre, "low", 1, training, stats_out)
    feat_mid = _head(feats[2], sd, pre, "mid", 2, training, stats_out)
    feat_high = _head(feats[3], sd, pre, "high", 4, training, stats_out)
    B = spec.shape[0]
    z = sd[pre + "speaker_embedding.0.weight"][vid] @ sd[pre + "speaker_embedding.1.weight"].t() + sd[pre + "speaker_embedding.1.bias"]
    h = F.elu(z)
    h = F.elu(h @ sd[pre + "fc1.weight"].t() + sd[pre + "fc1.bias"])
    w = (h @ sd[pre + "fc2.weight"].t() + sd[pre + "fc2.bias"]).reshape(B, 3, pose_level)
    w = torch.softmax(w, dim=1)
    blend = [feat_low * w[:, 0, i, None, None] + feat_mid * w[:, 1, i, None, None] + feat_high * w[:, 2, i, None, None]
             for i in range(pose_level)]
    return w, feat_low, feat_mid, feat_high, blend


# --------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------
def contrastive_loss(a: Tensor, b: Tensor, variant: str) -> Tensor:
    """SoftmaxContrastiveLoss.forward: train_hierarchy.py:54-68 (gesture: +1e-8, clamp) /
    train_hierarchy_expressive.py:108-121 (expressive: plain 1/D)."""
    a = F.normalize(a, p=2, dim=1)
    b = F.normalize(b, p=2, dim=1)
    D = torch.norm(a[:, None, :] - b[None, :, :], p=2, dim=2)
    if variant == "gesture":
        logits = torch.clamp(1.0 / (D + 1e-8), min=1e-8)
    else:
        logits = 1.0 / D
    return F.cross_entropy(logits, torch.arange(a.shape[0]))


def huber_sum(outs: Sequence[Tensor], tgts: Sequence[Tensor], beta: float = 0.1) -> Tensor:
    """train_hierarchy_expressive.py:312-318."""
    tot = 0
    for o, t in zip(outs, tgts):
        tot = tot + F.smooth_l1_loss(o / beta, t / beta) * beta
    return tot


ANGLE_TABLES = {}  # filled from ha2g constants below (values restated from the reference tables)


def physical_loss(out: Tensor, mean_dir_vec: Tensor, variant: str, pairs, avg, var) -> Tensor:
    """train_hierarchy_expressive.py:426-449 / train_hierarchy.py:242-262."""
    raw = out + mean_dir_vec.view(1, 1, -1)
    if variant == "expressive":
        lp = torch.cross(raw[:, :, 33:36], raw[:, :, 51:54], dim=2)
        rp = torch.cross(raw[:, :, 84:87], raw[:, :, 102:105], dim=2)
        raw = torch.cat((raw, lp, rp), dim=2)
    v = F.normalize(raw.reshape(raw.shape[0], raw.shape[1], -1, 3), dim=-1)
    v = v.reshape(-1, v.shape[2], 3)
    tot = 0
    for i, (p0, p1) in enumerate(pairs):
        ip = torch.clamp((v[:, p0] * v[:, p1]).sum(1), -1 + 1e-7, 1 - 1e-7)
        ang = torch.acos(ip) / math.pi
        tot = tot + torch.mean((ang - avg[i]) ** 2 / (2 * var[i]))
    return tot


def div_reg_loss(out: Tensor, out_rand: Tensor, z: Tensor, z_rand: Tensor, beta: float = 0.05) -> Tensor:
    """train_hierarchy_expressive.py:396-406."""
    l1 = F.smooth_l1_loss(out / beta, out_rand.detach() / beta, reduction="none") * beta
    l1 = l1.sum(dim=1).sum(dim=1)
    zl1 = (z.detach() - z_rand.detach()).abs().mean(1)
    return torch.clamp(-(l1 / (zl1 + 1e-5)), min=-1000).mean()


def kld_loss(mu: Tensor, logvar: Tensor) -> Tensor:
    return -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              b1: float = 0.5, b2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam (no amsgrad, no weight decay) single-tensor update; returns new (p,m,v)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


# --------------------------------------------------------------------------------------------
# whole step                                  train_hierarchy.py:71-293 / ..._expressive.py:124-483
# --------------------------------------------------------------------------------------------
def _leafify(sd: SD) -> SD:
    return {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone())
            for k, v in sd.items()}


def cascade(variant, gens: List[SD], target: Tensor, text: Tensor, blends, vid: Tensor, eps_list, n_pre: int,
            n_layers: int, hidden: int, rngs=None):
    chans = level_channels(variant)
    cmaps = cascade_maps(variant)
    outs, prev, last = [], None, None
    for k, g in enumerate(gens):
        tk = target[:, :, torch.as_tensor(chans[k])]
        pre = make_pre_seq(tk, prev, cmaps[k], n_pre)
        o, z, mu, lv = pose_generator(g, pre, text, blends[k], vid, eps_list[k], n_layers, hidden,
                                      None if rngs is None else rngs[k])
        outs.append(o)
        prev, last = o, (z, mu, lv)
    return outs, last


def train_step(variant: str, args, epoch: int, text: Tensor, spec: Tensor, target: Tensor, vid: Tensor,
               gens: List[SD], dis: SD, audio: SD, textenc: SD, opt_state: dict,
               eps: dict, rand_idx: Tensor, tables: dict):
    """Functional restatement of train_iter_hierarchy{,_expressive} with dropout disabled.

    eps = {'d': [L x (B,16)], 'g': [...], 'r': [...]} reparameterize draws in reference call order.
    opt_state[name] = {'step': int, 'm': {k: T}, 'v': {k: T}} or empty; updated in place.
    Returns (ret_dict, new state dicts: dict(gens=[...], dis=, audio=, text=)).
    """
    L = len(gens)
    n_pre, nl, H = args.n_pre_poses, args.n_layers, args.hidden_size
    gens = [_leafify(g) for g in gens]
    dis, audio, textenc = _leafify(dis), _leafify(audio), _leafify(textenc)
    chans = level_channels(variant)
    tgts = [target[:, :, torch.as_tensor(c)] for c in chans]
    lrs = {"dis": args.learning_rate * args.discriminator_lr_weight}

    def apply_adam(name, sd):
        st = opt_state.setdefault(name, {"step": 0, "m": {}, "v": {}})
        st["step"] += 1
        new = dict(sd)
        for k, p in sd.items():
            if not (torch.is_tensor(p) and p.requires_grad):
                continue
            if p.grad is None:
                continue
            m = st["m"].get(k, torch.zeros_like(p))
            v = st["v"].get(k, torch.zeros_like(p))
            pn, m, v = adam_step(p.detach(), p.grad, m, v, st["step"], lrs.get(name, args.learning_rate))
            st["m"][k], st["v"][k] = m, v
            new[k] = pn
        return new

    bn_a, bn_d = {}, {}
    w, f_low, f_mid, f_high, blends = wav_encoder(audio, spec, vid, L, True, bn_a)
    text_feat = text_encoder_tcn(text, textenc, "", n_levels=nl)

    dis_error = None
    gan_on = epoch > args.loss_warmup and args.loss_gan_weight > 0.0
    if gan_on:
        with torch.no_grad():
            outs_d, _ = cascade(variant, gens, target, text, blends, vid, eps["d"], n_pre, nl, H)
        real = conv_discriminator(dis, target, True, bn_d)
        sd_tmp = dict(dis); sd_tmp.update(bn_d)
        fake = conv_discriminator(sd_tmp, outs_d[-1].detach(), True, bn_d)
        dis_error = torch.sum(-torch.mean(torch.log(real + 1e-8) + torch.log(1 - fake + 1e-8)))
        dis_error.backward()
        dis.update({k: v for k, v in bn_d.items()})
        dis_new = apply_adam("dis", dis)
        dis = _leafify({k: (v.detach() if torch.is_tensor(v) else v) for k, v in dis_new.items()})

    c_pos = contrastive_loss(text_feat.reshape(-1, 32), f_high.reshape(-1, 32), variant)
    c_neg = -contrastive_loss(text_feat.reshape(-1, 32), f_low.reshape(-1, 32), variant)
    outs, (z, mu, lv) = cascade(variant, gens, target, text, blends, vid, eps["g"], n_pre, nl, H)
    huber = huber_sum(outs, tgts)
    bn_d2 = {}
    dis_out = conv_discriminator(dis, outs[-1], True, bn_d2)
    gen_error = -torch.mean(torch.log(dis_out + 1e-8))
    rand_vid = vid[rand_idx]
    with torch.no_grad():
        outs_r, (z_r, _, _) = cascade(variant, gens, target, text, blends, rand_vid, eps["r"], n_pre, nl, H)
    div = div_reg_loss(outs[-1], outs_r[-1], z, z_r)
    kld = kld_loss(mu, lv)
    loss = args.loss_regression_weight * huber + args.loss_kld_weight * kld + args.loss_reg_weight * div
    if epoch > args.loss_warmup:
        loss = loss + args.loss_gan_weight * gen_error
    loss = loss + args.loss_contrastive_pos_weight * c_pos + args.loss_contrastive_neg_weight * c_neg
    mdv = torch.tensor(args.mean_dir_vec, dtype=torch.float32).reshape(-1)
    phy = physical_loss(outs[-1], mdv, variant, tables["pairs"], tables["avg"], tables["var"])
    loss = loss + args.loss_physical_weight * phy
    for p in dis.values():
        if torch.is_tensor(p) and p.requires_grad:
            p.grad = None
    loss.backward()
    dis.update(bn_d2)
    audio.update(bn_a)
    new_gens = [apply_adam(f"g{k + 1}", g) for k, g in enumerate(gens)]
    new_audio = apply_adam("audio", audio)
    new_text = apply_adam("text", textenc)
    ret = {"loss": args.loss_regression_weight * huber.item(), "KLD": args.loss_kld_weight * kld.item(),
           "DIV_REG": args.loss_reg_weight * div.item()}
    if gan_on:
        ret["gen"] = args.loss_gan_weight * gen_error.item()
        ret["dis"] = dis_error.item()
    ret["c_pos"] = args.loss_contrastive_pos_weight * c_pos.item()
    ret["c_neg"] = args.loss_contrastive_neg_weight * c_neg.item()
    ret["phy"] = args.loss_physical_weight * phy.item()
    det = lambda sd: {k: (v.detach() if torch.is_tensor(v) else v) for k, v in sd.items()}
    grads = {"gens": [{k: v.grad for k, v in g.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None} for g in gens],
             "audio": {k: v.grad for k, v in audio.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None},
             "text": {k: v.grad for k, v in textenc.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}}
    return ret, {"gens": [det(g) for g in new_gens], "dis": det(dis), "audio": det(new_audio), "text": det(new_text)}, grads


# --------------------------------------------------------------------------------------------
# Inference loop (row L)                       synthesize_expressive_hierarchy.py:36-259
#                                              (TED-Gesture twin: synthesize_hierarchy.py:36-215)
# --------------------------------------------------------------------------------------------
def generate_gestures(variant: str, args, gens: List[SD], audio_sd: SD, word_index, audio: np.ndarray, words,
                      targets: List[Tensor], vid: int, eps_fn, mel_fn, audio_sr: int = 16000) -> np.ndarray:
    """Sliding-window autoregressive inference restated over state dicts (eval-mode modules, no fade-out).

    word_index(word) -> token id (lang_model.get_word_index);  eps_fn() -> (1,16) reparameterize noise, called once
    per generator per window in cascade order (:132-190);  mel_fn(audio) -> (128, frames) log-mel (:50).
    Pinned by tests/test_oracle_pinned.py against tests/golden/inference.pt, the output of the UNMODIFIED reference loop."""
    n_frames, n_pre, fps = args.n_poses, args.n_pre_poses, args.motion_resampling_framerate
    L = len(gens)
    chans = level_channels(variant)
    spectrogram = torch.as_tensor(np.asarray(mel_fn(audio), dtype=np.float32))             # :50
    clip_length = len(audio) / audio_sr                                                    # :43
    unit_time = n_frames / fps                                                             # :53
    stride_time = (n_frames - n_pre) / fps                                                 # :54
    num_subdivision = 1 if clip_length < unit_time else math.ceil((clip_length - unit_time) / stride_time) + 1   # :55-58
    spec_len = int(round((n_frames / fps * 16000 - 1024) / 512 + 1))                       # data_utils_expressive.py:91-93
    targets = [t.clone().float() for t in targets]
    vid_t = torch.tensor([vid], dtype=torch.int64)
    out_list: List[np.ndarray] = []
    out_prev = None
    for i in range(num_subdivision):                                                       # :76
        start_time = i * stride_time
        end_time = start_time + unit_time
        # quirk kept: the window start is scaled by spectrogram.shape[0] (= 128 mel rows), :84
        a0 = math.floor(start_time / clip_length * spectrogram.shape[0])
        in_spec = spectrogram[:, a0:a0 + spec_len].unsqueeze(0)                            # :85-87
        ext = np.zeros(n_frames)                                                           # :101-111
        frame_duration = (end_time - start_time) / n_frames
        for w in words:
            if w[1] >= end_time:
                break
            if w[2] <= start_time:
                continue
            idx = max(0, int(np.floor((w[1] - start_time) / frame_duration)))
            ext[idx] = word_index(w[0])
        in_text = torch.as_tensor(ext.astype(np.int64)).unsqueeze(0)                       # :112-114
        if i > 0:                                                                          # :117-125 seed frames
            for k in range(L):
                targets[k][0, 0:n_pre, :] = out_prev[-n_pre:][:, torch.as_tensor(chans[k])]
        with torch.no_grad():
            _, _, _, _, blends = wav_encoder(audio_sd, in_spec, vid_t, L, training=False)  # :130
            # the per-level targets are independent inputs in window 0 (:36-41), so the levels are chained here directly
            # (same wiring as cascade(): make_pre_seq of level k from level k-1's output)
            cmaps = cascade_maps(variant)
            outs, prev = [], None
            for k in range(L):                                                             # :132-190
                pre = make_pre_seq(targets[k], prev, cmaps[k], n_pre)
                o, _, _, _ = pose_generator(gens[k], pre, in_text, blends[k], vid_t, eps_fn(), args.n_layers, args.hidden_size)
                outs.append(o)
                prev = o
        out_seq = outs[-1][0].numpy().copy()                                               # :192
        out_prev = torch.as_tensor(out_seq)
        if out_list:                                                                       # :195-203 cross-fade
            last = out_list[-1][-n_pre:]
            out_list[-1] = out_list[-1][:-n_pre]
            n = len(last)
            for j in range(n):
                out_seq[j] = last[j] * (n - j) / (n + 1) + out_seq[j] * (j + 1) / (n + 1)
        out_list.append(out_seq)
    return np.vstack(out_list)                                                             # :206
