"""Drop-in for ``scripts/train_eval/train_gan.py::train_iter_gan`` (the baseline multimodal_context step that
``scripts/train.py:269-272`` calls): same positional signature, same returned dict of python floats.

Follows the reference line by line (train_gan.py:13-104): pre_seq from the target's seed frames, discriminator step on a
detached generator pass (epoch > loss_warmup), generator step with Huber + GAN + diversity regulariser + KLD.  As in the
hierarchy step, the passes whose outputs the reference detaches run without autograd, loss terms are back-propagated
as weighted roots, and all ``.item()`` reads are one packed device->host copy.
"""
from __future__ import annotations

from typing import Dict

import torch

from .. import dp, ops, ops_loss, rng
from ..optim import fused_adam_step, zero_grad


def _w(value: float, like: torch.Tensor) -> torch.Tensor:
    return torch.full((1,), float(value), device=like.device, dtype=torch.float32)


def train_iter_gan(args, epoch, in_text, in_audio, target_poses, vid_indices, pose_decoder, discriminator, pose_dec_optim,
                   dis_optim) -> Dict[str, float]:
    try:
        return _train_iter_gan(args, epoch, in_text, in_audio, target_poses, vid_indices, pose_decoder, discriminator,
                               pose_dec_optim, dis_optim)
    finally:
        ops.end_step()   # closes the step's zero-buffer slab (opened by rng.begin_step)


def _train_iter_gan(args, epoch, in_text, in_audio, target_poses, vid_indices, pose_decoder, discriminator, pose_dec_optim,
                    dis_optim) -> Dict[str, float]:
    warm_up_epochs = args.loss_warmup
    dev = target_poses.device
    rng.begin_step(dev)
    B, T, Dp = target_poses.shape
    # make pre seq input (:19-22): seed frames + constraint bit, zeros elsewhere == level-1 pre_seq without a previous level
    ident = torch.full((Dp + 1,), -1, dtype=torch.int32, device=dev)
    pre_seq = ops.pre_seq(target_poses.contiguous(), None, ident, ident[:1], args.n_pre_poses)
    scalars: Dict[str, torch.Tensor] = {}
    gan_on = epoch > warm_up_epochs and args.loss_gan_weight > 0.0
    use_reg = (args.z_type == "speaker" or args.z_type == "random") and args.loss_reg_weight > 0.0
    if use_reg and args.z_type != "speaker":
        raise NotImplementedError("z_type='random' is not on the multimodal_context config's path")

    # ------------------------------------------------------------------ train D (:27-46)
    if gan_on:
        zero_grad(dis_optim)
        with torch.no_grad():
            out_d, *_ = pose_decoder(pre_seq, in_text, in_audio, vid_indices)
        dis_real = discriminator(target_poses, in_text)
        dis_fake = discriminator(out_d.detach(), in_text)
        l_real, l_fake = ops_loss.neg_mean_log(dis_real), ops_loss.neg_mean_log1m(dis_fake)
        one = _w(1.0, target_poses)
        torch.autograd.backward([l_real, l_fake], [one, one])
        dp.allreduce_grads(dis_optim)
        fused_adam_step(dis_optim)
        scalars["dis_real"], scalars["dis_fake"] = l_real.detach(), l_fake.detach()

    # ------------------------------------------------------------------ train G (:50-96)
    zero_grad(pose_dec_optim)
    roots, root_w = [], []

    def add(name, t, w):
        scalars[name] = t.detach()
        if w != 0.0:
            roots.append(t)
            root_w.append(_w(w, t))

    out_dir_vec, z, z_mu, z_logvar = pose_decoder(pre_seq, in_text, in_audio, vid_indices)
    add("huber", ops_loss.huber(out_dir_vec, target_poses, 0.1), args.loss_regression_weight)
    flags = [(p, p.requires_grad) for p in discriminator.parameters()]
    for p, _ in flags:          # the reference computes, then discards, the discriminator's gradient of this term
        p.requires_grad_(False)
    try:
        dis_output = discriminator(out_dir_vec, in_text)
    finally:
        for p, f in flags:
            p.requires_grad_(f)
    add("gen", ops_loss.neg_mean_log(dis_output), args.loss_gan_weight if epoch > warm_up_epochs else 0.0)
    if use_reg:
        rand_vids = vid_indices[rng.randperm(B, vid_indices.device)]
        with torch.no_grad():
            out_rand, z_rand, _, _ = pose_decoder(pre_seq, in_text, in_audio, rand_vids)
        add("div_reg", ops_loss.div_reg(out_dir_vec, out_rand, z, z_rand, 0.05), args.loss_reg_weight)
        add("kld", ops_loss.kld(z_mu, z_logvar), args.loss_kld_weight)
    torch.autograd.backward(roots, root_w)
    dp.allreduce_grads(pose_dec_optim)
    fused_adam_step(pose_dec_optim)

    names = list(scalars.keys())
    v = dict(zip(names, torch.cat([scalars[n].reshape(1) for n in names]).tolist()))
    ret_dict = {"loss": args.loss_regression_weight * v["huber"]}
    if use_reg:
        if v["kld"]:
            ret_dict["KLD"] = args.loss_kld_weight * v["kld"]
        if v["div_reg"]:
            ret_dict["DIV_REG"] = args.loss_reg_weight * v["div_reg"]
    if gan_on:
        ret_dict["gen"] = args.loss_gan_weight * v["gen"]
        ret_dict["dis"] = v["dis_real"] + v["dis_fake"]
    return ret_dict
