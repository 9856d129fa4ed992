"""Test-set evaluation: drop-in for ``evaluate_testset``
(scripts/train_expressive.py:394-628, TED-Gesture twin scripts/train.py).

Per batch: eval-mode audio encoder + L-level cascade on the GPU (the same CUDA modules and cascade wiring as the
training step, seed frames from the target, a random speaker id per clip), then the reference's host-side metrics on
the outputs: L1 loss of the last level, mean absolute joint error after converting direction vectors back to joint
positions (n_pre seed frames excluded), and the acceleration difference; an optional ``embed_space_evaluator`` (any
object with the reference's ``reset / push_samples / get_no_of_samples / get_scores / get_diversity_scores`` methods,
utils/embedding_space_evaluator.py) is fed exactly like the reference feeds its own.  Returns the reference's dict
(``loss``, ``joint_mae`` [+ ``frechet``, ``feat_dist``, ``diversity``, ``bc``]).  The beat-consistency branch is disabled
in the reference (``beat_consistency_score = False``) and therefore reports ``bc = 0``.
"""
from __future__ import annotations

import logging
import random
import time

import numpy as np
import torch

from . import cascade
from . import constants as K


class AverageMeter:
    """utils/average_meter.py."""

    def __init__(self, name):
        self.name, self.val, self.avg, self.sum, self.count = name, 0, 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def convert_dir_vec_to_pose(vec: np.ndarray, variant: str) -> np.ndarray:
    """Direction vectors [..., J*3] or [..., J, 3] -> joint positions [..., J+1, 3] (root at the origin):
    joint[child] = joint[parent] + bone_length * dir_vec (utils/data_utils_expressive.py:127-147, data_utils.py:76-98)."""
    pairs = K.EXPRESSIVE_DIR_VEC_PAIRS if variant == "expressive" else K.GESTURE_DIR_VEC_PAIRS
    vec = np.asarray(vec)
    if vec.shape[-1] != 3:
        vec = vec.reshape(vec.shape[:-1] + (-1, 3))
    if vec.shape[-2] != len(pairs):
        raise ValueError(f"expected {len(pairs)} direction vectors, got {vec.shape[-2]}")
    joint_pos = np.zeros(vec.shape[:-2] + (len(pairs) + 1, 3), dtype=np.float64)
    for j, (parent, child, length) in enumerate(pairs):
        joint_pos[..., child, :] = joint_pos[..., parent, :] + length * vec[..., j, :]
    return joint_pos


def pose_metrics(out_dir_vec: np.ndarray, target_vec: np.ndarray, mean_dir_vec: np.ndarray, n_poses: int, n_pre_poses: int,
                 variant: str):
    """(joint MAE, acceleration difference) of one batch, train_expressive.py:583-602."""
    out_poses = convert_dir_vec_to_pose(out_dir_vec + mean_dir_vec, variant)
    target_poses = convert_dir_vec_to_pose(target_vec + mean_dir_vec, variant)
    if out_poses.shape[1] == n_poses:
        diff = out_poses[:, n_pre_poses:] - target_poses[:, n_pre_poses:]
    else:
        diff = out_poses - target_poses[:, n_pre_poses:]
    mae = float(np.mean(np.absolute(diff)))
    accel = float(np.mean(np.abs(np.diff(target_poses, n=2, axis=1) - np.diff(out_poses, n=2, axis=1))))
    return mae, accel


@torch.no_grad()
def evaluate_testset(test_data_loader, generator, *rest, sigma=0.1, thres=0.03):
    """evaluate_testset(test_data_loader, generator, g1..gL, audio_encoder, loss_fn, embed_space_evaluator, args)."""
    L = len(rest) - 4
    if L not in (3, 6):
        raise TypeError("expected (test_data_loader, generator, g1..gL, audio_encoder, loss_fn, embed_space_evaluator, args)")
    variant = "expressive" if L == 6 else "gesture"
    gens = list(rest[:L])
    audio_encoder, _loss_fn, embed_space_evaluator, args = rest[L:]
    dev = next(gens[0].parameters()).device
    for m in gens + [audio_encoder]:
        m.train(False)
    if embed_space_evaluator:
        embed_space_evaluator.reset()
    losses, joint_mae, accel, bc = AverageMeter("loss"), AverageMeter("mae_on_joint"), AverageMeter("accel"), AverageMeter("bc")
    start = time.time()
    mean_dir_vec = np.array(args.mean_dir_vec).squeeze()
    enc = getattr(audio_encoder, "module", audio_encoder)   # DataParallel-wrapped encoders (train_expressive.py:431-434)
    speaker_model = getattr(getattr(enc, "feat_extractor", enc), "z_obj", None)
    if speaker_model is None:
        raise NotImplementedError("the hierarchy path needs a speaker model (z_type = 'speaker')")

    for data in test_data_loader:
        _in_text, _text_lengths, in_text_padded, _, target_vec, in_audio, in_spec, _aux = data
        batch_size = target_vec.size(0)
        in_text_padded = in_text_padded.to(dev)
        in_spec = in_spec.float().to(dev)
        target = target_vec.to(dev).float()
        ids = list(speaker_model.word2index.values())
        vid_indices = torch.LongTensor([random.choice(ids) for _ in range(batch_size)]).to(dev)

        _, _, _, _, linear_blend_feat = audio_encoder(in_spec, vid_indices)
        targets = cascade.split_targets(variant, target)
        outs, _ = cascade.run_cascade(variant, gens, targets, in_text_padded, linear_blend_feat, vid_indices,
                                      args.n_pre_poses)
        out_dir_vec = outs[-1]
        out_np = out_dir_vec.float().cpu().numpy()
        tgt_np = target_vec.float().cpu().numpy()
        losses.update(float(np.mean(np.abs(out_np - tgt_np))), batch_size)          # F.l1_loss(out_dir_vec, target_6)
        if embed_space_evaluator:
            embed_space_evaluator.push_samples(in_text_padded, in_audio.to(dev), out_dir_vec, target)
        mae, acc = pose_metrics(out_np, tgt_np, mean_dir_vec, args.n_poses, args.n_pre_poses, variant)
        joint_mae.update(mae, batch_size)
        accel.update(acc, batch_size)

    for m in gens + [audio_encoder]:   # the reference switches back to train mode unconditionally (train_expressive.py:601-607)
        m.train(True)
    ret_dict = {"loss": losses.avg, "joint_mae": joint_mae.avg}
    elapsed = time.time() - start
    if embed_space_evaluator and embed_space_evaluator.get_no_of_samples() > 0:
        frechet_dist, feat_dist = embed_space_evaluator.get_scores()
        diversity_score = embed_space_evaluator.get_diversity_scores()
        logging.info("[VAL] loss: %.3f, joint mae: %.5f, accel diff: %.5f, FGD: %.3f, feat_D: %.3f, Diversity: %.3f, "
                     "BC: %.4f / %.1fs", losses.avg, joint_mae.avg, accel.avg, frechet_dist, feat_dist, diversity_score,
                     bc.avg, elapsed)
        ret_dict.update(frechet=frechet_dist, feat_dist=feat_dist, diversity=diversity_score, bc=bc.avg)
    else:
        logging.info("[VAL] loss: %.3f, joint mae: %.3f / %.1fs", losses.avg, joint_mae.avg, elapsed)
    return ret_dict
