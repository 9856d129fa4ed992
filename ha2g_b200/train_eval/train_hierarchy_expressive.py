"""Drop-in for scripts/train_eval/train_hierarchy_expressive.py::train_iter_hierarchy_expressive
(TED-Expressive, 6 levels).  Same positional signature as the call at scripts/train_expressive.py:342-346."""
from ._step import train_step


def train_iter_hierarchy_expressive(args, epoch, in_text_padded, in_spec, target, vid_indices,
                                    g1, g2, g3, g4, g5, g6, discriminator, audio_encoder, text_encoder,
                                    gen_optimizer_1, gen_optimizer_2, gen_optimizer_3,
                                    gen_optimizer_4, gen_optimizer_5, gen_optimizer_6, dis_optimizer,
                                    audio_optimizer, text_optimizer):
    return train_step("expressive", args, epoch, in_text_padded, in_spec, target, vid_indices,
                      [g1, g2, g3, g4, g5, g6], discriminator, audio_encoder, text_encoder,
                      [gen_optimizer_1, gen_optimizer_2, gen_optimizer_3, gen_optimizer_4, gen_optimizer_5,
                       gen_optimizer_6], dis_optimizer, audio_optimizer, text_optimizer)
