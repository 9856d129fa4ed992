"""Checkpoint compatibility with the reference.

The reference writes one pickled dict per checkpoint (scripts/train_expressive.py:270-304, scripts/train.py, saved by
``utils.train_utils.save_checkpoint`` = ``torch.save``) and restores it with ``load_checkpoint_hierarchy``
(scripts/utils/train_utils_expressive.py:170-205; TED-Gesture twin: utils/train_utils.py:170-200):

    {'args', 'epoch', 'lang_model', 'speaker_model', 'pose_dim',
     'gen_dict_1' .. 'gen_dict_L', 'dis_dict', 'audio_dict', 'text_dict'}          L = 6 (Expressive) / 3 (Gesture)

``save_checkpoint_hierarchy`` writes exactly that layout from the modules of ``ha2g_b200.model`` (their ``state_dict``
keys, shapes and order are the reference's, tests/test_boundary.py), so the file loads with the reference's own loader;
``load_checkpoint_hierarchy`` reads a file written by either side and returns the reference's tuple
``(args, g1..gL, audio_encoder, loss_fn, lang_model, speaker_model, pose_dim)`` with our CUDA modules in eval mode.

Pickled ``lang_model`` / ``speaker_model`` objects are instances of the reference's ``model.vocab.Vocab``; when that
module is not importable (this package without the reference tree) the loader maps the class onto
``ha2g_b200.model.vocab.Vocab``, which has the same fields.  Optimizer state is not part of the reference format;
``save_checkpoint_hierarchy(..., optimizers=...)`` stores it under the extra key ``'optim_dicts'`` that the reference
loader ignores.
"""
from __future__ import annotations

import contextlib
import importlib
import sys
import types
from typing import List, Optional, Sequence

import torch

from .model import vocab as _vocab
from .model.hierarchy_net import Hierarchical_PoseGenerator, Hierarchical_WavEncoder

LEVEL_DIMS = {3: (5 * 3, 7 * 3, 9 * 3), 6: (8 * 3, 10 * 3, 12 * 3, 22 * 3, 32 * 3, 42 * 3)}


def save_checkpoint_hierarchy(path: str, args, epoch: int, lang_model, speaker_model, pose_dim: int, gens: Sequence,
                              discriminator, audio_encoder, text_encoder, optimizers: Optional[dict] = None) -> dict:
    """Write the reference's checkpoint dict (train_expressive.py:299-304).  Returns the dict that was saved."""
    L = len(gens)
    if L not in LEVEL_DIMS:
        raise ValueError("the hierarchy has 3 (TED-Gesture) or 6 (TED-Expressive) generators")
    cpu = lambda m: {k: v.detach().cpu() for k, v in m.state_dict().items()}
    state = {"args": args, "epoch": epoch, "lang_model": lang_model, "speaker_model": speaker_model, "pose_dim": pose_dim}
    for k, g in enumerate(gens, start=1):
        state[f"gen_dict_{k}"] = cpu(g)
    state["dis_dict"] = cpu(discriminator) if discriminator is not None else None
    state["audio_dict"] = cpu(audio_encoder)
    state["text_dict"] = cpu(text_encoder)
    if optimizers:
        state["optim_dicts"] = {name: opt.state_dict() for name, opt in optimizers.items()}
    torch.save(state, path)
    return state


@contextlib.contextmanager
def _reference_vocab_alias():
    """Let pickle resolve ``model.vocab.Vocab`` when the reference tree is not on sys.path."""
    added = []
    try:
        importlib.import_module("model.vocab")
    except Exception:
        if "model" not in sys.modules:
            sys.modules["model"] = types.ModuleType("model")
            added.append("model")
        sys.modules["model.vocab"] = _vocab
        setattr(sys.modules["model"], "vocab", _vocab)
        added.append("model.vocab")
    try:
        yield
    finally:
        for name in added:
            sys.modules.pop(name, None)


def read_checkpoint(path: str, map_location="cpu") -> dict:
    """The raw checkpoint dict (trusted pickle, like the reference's ``torch.load``)."""
    with _reference_vocab_alias():
        return torch.load(path, map_location=map_location, weights_only=False)


def load_checkpoint_hierarchy(path: str, _device="cuda:0"):
    """Counterpart of ``load_checkpoint_hierarchy`` (utils/train_utils_expressive.py:170-205).
    -> (args, g1, ..., gL, audio_encoder, loss_fn, lang_model, speaker_model, pose_dim), modules in eval mode."""
    ck = read_checkpoint(path)
    args, lang_model, speaker_model, pose_dim = ck["args"], ck["lang_model"], ck["speaker_model"], ck["pose_dim"]
    L = sum(1 for k in ck if k.startswith("gen_dict_"))
    if L not in LEVEL_DIMS:
        raise ValueError(f"checkpoint holds {L} generators; expected 3 or 6")
    gens: List[torch.nn.Module] = []
    for k, d in enumerate(LEVEL_DIMS[L], start=1):
        g = Hierarchical_PoseGenerator(args, d, lang_model.n_words, args.wordembed_dim, lang_model.word_embedding_weights,
                                       z_obj=speaker_model)
        g.load_state_dict(ck[f"gen_dict_{k}"])
        gens.append(g.to(_device).train(False))
    audio_encoder = Hierarchical_WavEncoder(args, speaker_model, pose_level=L, nOut=32)
    audio_encoder.load_state_dict(ck["audio_dict"])
    audio_encoder = audio_encoder.to(_device).train(False)
    loss_fn = torch.nn.L1Loss()   # train_expressive.py:127 (returned for signature compatibility; unused by the step)
    return (args, *gens, audio_encoder, loss_fn, lang_model, speaker_model, pose_dim)


def resume_training(path: str, gens: Sequence, discriminator, audio_encoder, text_encoder,
                    optimizers: Optional[dict] = None) -> int:
    """Restore a training run in place from a checkpoint written by ``save_checkpoint_hierarchy`` (or by the reference's
    ``train_epochs``, which stores no optimizer state: the moments then restart from zero, as they would there):
    module weights and buffers of every generator, the discriminator and both encoders, and -- when the file carries
    ``optim_dicts`` and ``optimizers`` names the same keys -- the Adam moments and step counts.  Returns the stored epoch.
    The captured CUDA graph of the step (if any) notices the re-allocated moment tensors and re-captures itself
    (graph_step._still_valid)."""
    ck = read_checkpoint(path)
    L = sum(1 for k in ck if k.startswith("gen_dict_"))
    if L != len(gens):
        raise ValueError(f"checkpoint holds {L} generators, the run has {len(gens)}")
    for k, g in enumerate(gens, start=1):
        g.load_state_dict(ck[f"gen_dict_{k}"])
    if discriminator is not None and ck.get("dis_dict") is not None:
        discriminator.load_state_dict(ck["dis_dict"])
    audio_encoder.load_state_dict(ck["audio_dict"])
    if text_encoder is not None and ck.get("text_dict") is not None:
        text_encoder.load_state_dict(ck["text_dict"])
    if optimizers and ck.get("optim_dicts"):
        for name, opt in optimizers.items():
            if name in ck["optim_dicts"]:
                opt.load_state_dict(ck["optim_dicts"][name])
    return int(ck["epoch"])
