"""One HA2G hierarchical training step on a B200, shared by both dataset variants.

Mirrors ``train_iter_hierarchy`` (scripts/train_eval/train_hierarchy.py:71-293, 3 levels) and
``train_iter_hierarchy_expressive`` (scripts/train_eval/train_hierarchy_expressive.py:124-483, 6 levels):
same call order of the modules, the same losses and weights, the same RNG draws in the same order,
the same optimizer steps and the same returned dict of python floats.  Differences are purely
mechanical and result-preserving:
  * the two cascades whose outputs the reference detaches (D-step, random-speaker pass) run under
    ``torch.no_grad`` (no BPTT state is saved for them);
  * the discriminator's parameters do not accumulate gradient during the generator step (the reference
    computes then discards those gradients: dis_optimizer.zero_grad() precedes the next use);
  * loss terms are back-propagated as several roots with constant weights instead of being summed into
    one scalar first (identical gradients, no scalar arithmetic kernels);
  * those no-grad cascades ride along with the differentiated cascade as extra batch rows: ONE cascade over 3B rows whose
    autograd graph covers B of them (``ops.ride_along``; ``HA2G_BATCH_PASSES=0`` separates the passes again); the random
    draws are made up front in the reference's order;
  * all ``.item()`` reads are packed into a single device->host copy at the end of the step;
  * after two eager calls per (modules, shapes, mode) signature the whole step -- forward, both backwards, the
    gradient all-reduce and the eight Adam updates, ~7 600 kernel launches -- is captured into ONE CUDA graph
    (``ha2g_b200/graph_step.py``) and every later call is: copy the four inputs into the graph's static buffers,
    replay, read the packed scalars.  ``HA2G_CUDA_GRAPH=0`` keeps the eager path.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List

import torch

import os

from .. import cascade, dp, graph_step, ops, ops_loss, rng
from ..optim import fused_adam_step, zero_grad

_BATCH_PASSES = os.environ.get("HA2G_BATCH_PASSES", "1") != "0"


@contextlib.contextmanager
def _frozen(module):
    flags = [(p, p.requires_grad) for p in module.parameters()]
    for p, _ in flags:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p, f in flags:
            p.requires_grad_(f)


def _w(value: float, like: torch.Tensor) -> torch.Tensor:
    return torch.full((1,), float(value), device=like.device, dtype=torch.float32)


def train_step(variant: str, args, epoch, in_text_padded, in_spec, target, vid_indices, gens: List, discriminator,
               audio_encoder, text_encoder, gen_optimizers: List, dis_optimizer, audio_optimizer, text_optimizer
               ) -> Dict[str, float]:
    """The drop-in step: CUDA-graph replay when a captured graph exists for this signature, eager otherwise."""
    world = (variant, args, epoch, in_text_padded, in_spec, target, vid_indices, gens, discriminator, audio_encoder,
             text_encoder, gen_optimizers, dis_optimizer, audio_optimizer, text_optimizer)
    hit = graph_step.run(enqueue_step, world)
    if hit is not None:
        names, vals, flags = hit
    else:
        names, packed, flags = enqueue_step(*world)
        vals = packed.tolist()
    return _finish(args, names, vals, flags)


def _enqueue_step(variant: str, args, epoch, in_text_padded, in_spec, target, vid_indices, gens: List, discriminator,
                 audio_encoder, text_encoder, gen_optimizers: List, dis_optimizer, audio_optimizer, text_optimizer,
                 adam=None):
    """Enqueue one whole step on the current stream without any host synchronisation.
    -> (names, packed device tensor of the step's scalars, flags for _finish)."""
    if adam is None:
        adam = fused_adam_step   # resolved at call time (the eager multi-tensor step; graph capture passes its own)
    warm_up_epochs = args.loss_warmup
    n_pre = args.n_pre_poses
    dev = target.device
    rng.begin_step(dev)   # new dropout stream position for this step (device-side: valid under graph replay)

    scalars: Dict[str, torch.Tensor] = {}
    B, L = target.shape[0], len(gens)
    gan_on = epoch > warm_up_epochs and args.loss_gan_weight > 0.0
    use_reg = (args.z_type == "speaker" or args.z_type == "random") and args.loss_reg_weight > 0.0
    if use_reg and args.z_type != "speaker":
        raise NotImplementedError("z_type='random' is not on the hierarchy configs' path")
    tails = (["d"] if gan_on else []) + (["r"] if use_reg else [])
    ride = _BATCH_PASSES and len(tails) > 0 and rng.fused_dropout()

    # The generators' own text encoders depend on the token batch only: with the ride-along cascade (below) they are
    # enqueued first, on side streams, so that their many small kernels run next to the audio encoder now and -- autograd
    # replays a node on the stream of its forward -- next to the other generators' GRU recurrences in the backward pass.
    text_p = gen_text_feats = None
    if ride:
        text_p = ops.ride_pack(in_text_padded, [in_text_padded] * len(tails))
        if ops.side_streams(dev):
            ops.prime_embedding_heads(text_p)            # shared by every table's scatter-add: computed once, before the fork
            side = ops.fork_side_streams(dev)
            gen_text_feats = [None] * L
            with ops.ride_along(1 + len(tails)):
                for k, g in enumerate(gens):
                    with torch.cuda.stream(side[k % len(side)]):
                        gen_text_feats[k] = g.text_encoder(text_p)

    weight, feat_low, feat_mid, feat_high, linear_blend_feat = audio_encoder(in_spec, vid_indices)
    text_feat = text_encoder(in_text_padded)
    targets = cascade.split_targets(variant, target)

    # The contrastive losses need the encoders' outputs only: on their own stream (with its own scratch arena -- the packed
    # N_local x N_global coefficient matrix of the data-parallel loss is the largest tenant) they run next to the cascade,
    # forward and backward, and so do their all-gather / reduce-scatter under data parallelism.
    c_losses = {}
    loss_stream = ops.loss_stream(dev) if ride else None

    def contrastive_losses():
        tf_ = text_feat.reshape(-1, text_feat.shape[2])
        if args.loss_contrastive_pos_weight > 0.0:
            c_losses["c_pos"] = ops_loss.contrastive(tf_, feat_high.reshape(-1, feat_high.shape[2]), variant)
        if args.loss_contrastive_neg_weight > 0.0:  # text_low_contrastive = -criterion(...)
            c_losses["c_neg"] = ops_loss.contrastive(tf_, feat_low.reshape(-1, feat_low.shape[2]), variant)

    if loss_stream is not None:
        loss_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(loss_stream):
            contrastive_losses()

    # The two cascades whose outputs the reference detaches -- the discriminator-step pass and the mismatched-speaker
    # pass of the diversity loss -- depend only on the (unchanged) generators, so they RIDE ALONG with the differentiated
    # G-step cascade as extra batch rows (ops.ride_along): ONE cascade over 3B rows, a third of the launches, three times
    # the rows per GEMM / GRU step, while autograd, the saved activations and every backward kernel only see the B rows
    # of the G-step pass.  The draws are made up front in the reference's order (D-pass noise x L, G-pass noise x L,
    # randperm, mismatched-pass noise x L).
    eps_g = out_r_last = z_context_rand = out_d_last = ride_result = None
    if ride:
        eps_d = [rng.randn((B, 16), dev) for _ in range(L)] if gan_on else None
        eps_g = [rng.randn((B, 16), dev) for _ in range(L)]
        rand_vids = vid_indices[rng.randperm(B, vid_indices.device)] if use_reg else None
        eps_r = [rng.randn((B, 16), dev) for _ in range(L)] if use_reg else None
        m = 1 + len(tails)
        per_tail = lambda d_val, r_val: [d_val if t == "d" else r_val for t in tails]
        targets_p = [ops.ride_pack(t, per_tail(t, t)) for t in targets]
        blends_p = [ops.ride_pack(f, per_tail(f, f)) for f in linear_blend_feat]
        vid_p = ops.ride_pack(vid_indices, per_tail(vid_indices, rand_vids))
        eps_p = [ops.ride_pack(eps_g[k], per_tail(eps_d[k] if gan_on else None, eps_r[k] if use_reg else None))
                 for k in range(L)]
        if gen_text_feats is not None:
            main = torch.cuda.current_stream(dev)
            for s_ in ops.side_streams(dev):
                main.wait_stream(s_)
        with ops.ride_along(m):
            outs, (z_context, z_mu, z_logvar) = cascade.run_cascade(variant, gens, targets_p, text_p, blends_p, vid_p,
                                                                    n_pre, eps=eps_p, text_feats=gen_text_feats)
        out_tails, z_tails = ops.ride_tails(outs[-1], m), ops.ride_tails(z_context, m)
        if gan_on:
            out_d_last = out_tails[tails.index("d")]
        if use_reg:
            out_r_last, z_context_rand = out_tails[tails.index("r")], z_tails[tails.index("r")]
        ride_result = (outs, z_context, z_mu, z_logvar)

    # ------------------------------------------------------------------ train D
    if gan_on:
        zero_grad(dis_optimizer)
        if not ride:
            with torch.no_grad():
                outs_d, _ = cascade.run_cascade(variant, gens, targets, in_text_padded, linear_blend_feat, vid_indices, n_pre)
            out_d_last = outs_d[-1]
        dis_real = discriminator(target, in_text_padded)
        dis_fake = discriminator(out_d_last.detach(), in_text_padded)
        l_real = ops_loss.neg_mean_log(dis_real)
        l_fake = ops_loss.neg_mean_log1m(dis_fake)
        one = _w(1.0, target)
        torch.autograd.backward([l_real, l_fake], [one, one])
        dp.allreduce_grads(dis_optimizer)
        adam(dis_optimizer)
        scalars["dis_real"], scalars["dis_fake"] = l_real.detach(), l_fake.detach()

    # ------------------------------------------------------------------ train G
    for opt in gen_optimizers:
        zero_grad(opt)
    zero_grad(audio_optimizer)
    zero_grad(text_optimizer)

    roots, root_w = [], []

    def add(name, t, w):
        scalars[name] = t.detach()
        if w != 0.0:
            roots.append(t)
            root_w.append(_w(w, t))

    if loss_stream is not None:
        torch.cuda.current_stream(dev).wait_stream(loss_stream)
    else:
        contrastive_losses()
    if "c_pos" in c_losses:
        add("c_pos", c_losses["c_pos"], args.loss_contrastive_pos_weight)
    if "c_neg" in c_losses:
        add("c_neg", c_losses["c_neg"], -args.loss_contrastive_neg_weight)

    if ride:
        outs, z_context, z_mu, z_logvar = ride_result
    else:
        outs, (z_context, z_mu, z_logvar) = cascade.run_cascade(variant, gens, targets, in_text_padded, linear_blend_feat,
                                                                vid_indices, n_pre)
    out_dir_vec = outs[-1]
    for k, (o, t) in enumerate(zip(outs, targets)):
        add(f"huber{k}", ops_loss.huber(o, t, 0.1), args.loss_regression_weight)

    with _frozen(discriminator):
        dis_output = discriminator(out_dir_vec, in_text_padded)
    add("gen", ops_loss.neg_mean_log(dis_output), args.loss_gan_weight if epoch > warm_up_epochs else 0.0)

    if use_reg:
        if not ride:
            rand_idx = rng.randperm(vid_indices.shape[0], vid_indices.device)
            rand_vids = vid_indices[rand_idx]
            with torch.no_grad():
                outs_r, (z_context_rand, _, _) = cascade.run_cascade(variant, gens, targets, in_text_padded,
                                                                     linear_blend_feat, rand_vids, n_pre)
            out_r_last = outs_r[-1]
        add("div_reg", ops_loss.div_reg(out_dir_vec, out_r_last, z_context, z_context_rand, 0.05), args.loss_reg_weight)
        add("kld", ops_loss.kld(z_mu, z_logvar), args.loss_kld_weight)

    if args.loss_physical_weight > 0.0:
        mdv = [float(v[0]) if isinstance(v, (list, tuple)) else float(v) for v in args.mean_dir_vec]
        add("phy", ops_loss.physical(out_dir_vec, variant, mdv), args.loss_physical_weight)

    g_opts = list(gen_optimizers) + [audio_optimizer, text_optimizer]
    dp.begin_backward(g_opts)     # data parallel: all-reduces are launched from gradient hooks DURING the backward pass
    try:
        torch.autograd.backward(roots, root_w)
    finally:
        dp.end_backward()

    for opt in reversed(g_opts[:len(gen_optimizers)]):   # g_L's gradients (and its all-reduce) finish first
        dp.allreduce_grads(opt)   # no-op on one GPU; otherwise waits for the overlapped all-reduce
        adam(opt)
    for opt in (audio_optimizer, text_optimizer):
        dp.allreduce_grads(opt)
        adam(opt)

    # ------------------------------------------------------------------ one packed device->host read
    names = list(scalars.keys())
    packed = torch.cat([scalars[n].reshape(1) for n in names])
    return names, packed, {"use_reg": use_reg, "gan_on": gan_on, "levels": len(outs)}


def enqueue_step(*a, **kw):
    """Enqueue one whole step on the current stream without any host synchronisation (see _enqueue_step); closes the
    step's zero-buffer slab (ops.begin_step opened it through rng.begin_step) whether or not the step raised."""
    try:
        return _enqueue_step(*a, **kw)
    finally:
        ops.end_step()


def _finish(args, names, vals, flags) -> Dict[str, float]:
    """The reference's returned dict of python floats (train_hierarchy_expressive.py:468-483)."""
    use_reg, gan_on = flags["use_reg"], flags["gan_on"]
    v = dict(zip(names, vals))
    huber_loss = sum(v[f"huber{k}"] for k in range(flags["levels"]))
    ret_dict = {"loss": args.loss_regression_weight * huber_loss}
    if use_reg:
        if v["kld"]:
            ret_dict["KLD"] = args.loss_kld_weight * v["kld"]
        if v["div_reg"]:
            ret_dict["DIV_REG"] = args.loss_reg_weight * v["div_reg"]
    if gan_on:
        ret_dict["gen"] = args.loss_gan_weight * v["gen"]
        ret_dict["dis"] = v["dis_real"] + v["dis_fake"]
    if args.loss_contrastive_pos_weight > 0.0:
        ret_dict["c_pos"] = args.loss_contrastive_pos_weight * v["c_pos"]
    if args.loss_contrastive_neg_weight > 0.0:
        ret_dict["c_neg"] = args.loss_contrastive_neg_weight * (-v["c_neg"])
    if args.loss_physical_weight > 0.0:
        ret_dict["phy"] = args.loss_physical_weight * v["phy"]
    return ret_dict
