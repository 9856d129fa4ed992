"""Drop-in for scripts/train_eval/train_hierarchy.py::train_iter_hierarchy (TED-Gesture, 3 levels).

Same positional signature as the call at scripts/train.py:275-278; returns the same dict of python floats
(keys loss, KLD, DIV_REG, gen, dis, c_pos, c_neg, phy).  The step itself runs on hand-written sm_100a
kernels (see ha2g_b200/train_eval/_step.py and csrc/)."""
from ._step import train_step


def train_iter_hierarchy(args, epoch, in_text_padded, in_spec, target, vid_indices,
                         g1, g2, g3, discriminator, audio_encoder, text_encoder,
                         gen_optimizer_1, gen_optimizer_2, gen_optimizer_3, dis_optimizer,
                         audio_optimizer, text_optimizer):
    return train_step("gesture", args, epoch, in_text_padded, in_spec, target, vid_indices, [g1, g2, g3],
                      discriminator, audio_encoder, text_encoder,
                      [gen_optimizer_1, gen_optimizer_2, gen_optimizer_3], dis_optimizer, audio_optimizer,
                      text_optimizer)
