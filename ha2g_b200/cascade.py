"""Coarse-to-fine cascade wiring of the hierarchical pose generators (bit-exact index contract).

Reference: scripts/train_eval/train_hierarchy.py:86-88,153-170 (TED-Gesture, 3 levels) and
scripts/train_eval/train_hierarchy_expressive.py:140-145,252-310 (TED-Expressive, 6 levels); the same
tables are repeated in train_expressive.py:472-530 and synthesize_expressive_hierarchy.py:132-190.

Every explicit slice assignment there is an instance of one rule (plus one quirk, see host_tables):
pose channel 3*b+c belongs to bone b;
level k owns an ascending list of bones; ``target_k`` gathers those bones; ``pre_seq_k`` carries the
seed frames of ``target_k`` plus, for frames >= n_pre_poses, the previous level's output scattered to
the slots the same bones occupy at level k.  The tables are built from the bone lists below and the
forward/backward copies run in csrc/elementwise.cu (pre_seq_fwd/bwd, gather_cols).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops

_HEAD = [37, 38, 39, 40, 41]  # the five head bones ("-5*3:" in the reference slices)

BONE_LEVELS = {
    # train_hierarchy.py:86-88
    "gesture": [
        [0, 1, 2, 3, 6],
        [0, 1, 2, 3, 4, 6, 7],
        [0, 1, 2, 3, 4, 5, 6, 7, 8],
    ],
    # train_hierarchy_expressive.py:140-145
    "expressive": [
        [0, 1, 2] + _HEAD,
        [0, 1, 2, 3, 20] + _HEAD,
        [0, 1, 2, 3, 4, 20, 21] + _HEAD,
        [0, 1, 2, 3, 4, 5, 8, 11, 14, 17, 20, 21, 22, 25, 28, 31, 34] + _HEAD,
        [0, 1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 14, 15, 17, 18, 20, 21, 22, 23, 25, 26, 28, 29, 31, 32, 34, 35] + _HEAD,
        list(range(42)),
    ],
}


def level_dims(variant: str) -> List[int]:
    return [3 * len(b) for b in BONE_LEVELS[variant]]


def host_tables(variant: str):
    """Per level: (channel gather list, slot_src [d_k+1], src_slot [d_{k-1}]) as Python int lists."""
    levels = BONE_LEVELS[variant]
    out = []
    for k, bones in enumerate(levels):
        chans = [3 * b + c for b in bones for c in range(3)]
        d = len(chans)
        slot_src = [-1] * (d + 1)
        src_slot: List[int] = []
        if k > 0:
            prev = levels[k - 1]
            src_slot = [-1] * (3 * len(prev))
            for s_prev, b in enumerate(prev):
                s_cur = bones.index(b)  # every level contains the previous level's bones
                # Reference quirk (kept for parity): head bones are written through
                # ``pre_seq_k[:, n_pre:, -5*3:] = out_{k-1}[:, n_pre:, -5*3:]`` and pre_seq_k is one column wider
                # than the pose (flag last), so their destination is shifted right by one column -- the last
                # head channel lands in the flag column, the first head column stays 0
                # (train_hierarchy_expressive.py:164,171,178,185,200,215).
                shift = 1 if (variant == "expressive" and b in _HEAD) else 0
                for c in range(3):
                    slot_src[3 * s_cur + c + shift] = 3 * s_prev + c
                    src_slot[3 * s_prev + c] = 3 * s_cur + c + shift
        out.append((chans, slot_src, src_slot))
    return out


_dev_cache: Dict = {}


def device_tables(variant: str, device):
    key = (variant, str(device))
    if key not in _dev_cache:
        tabs = []
        for chans, slot_src, src_slot in host_tables(variant):
            tabs.append((torch.tensor(chans, dtype=torch.int32, device=device),
                         torch.tensor(slot_src, dtype=torch.int32, device=device),
                         torch.tensor(src_slot if src_slot else [-1], dtype=torch.int32, device=device)))
        _dev_cache[key] = tabs
    return _dev_cache[key]


def split_targets(variant: str, target: torch.Tensor) -> List[torch.Tensor]:
    """target_1..target_L (train_hierarchy_expressive.py:140-145)."""
    tabs = device_tables(variant, target.device)
    return [ops.gather_cols(target, t[0]) for t in tabs]


def run_cascade(variant: str, gens, targets: List[torch.Tensor], in_text, blends, vid, n_pre: int, eps=None, text_feats=None):
    """g1 -> ... -> gL with the pre_seq wiring; returns ([out_1..out_L], (z, mu, logvar) of the last level).
    eps: optional per-level reparameterisation noise drawn by the caller (else each generator draws its own);
    text_feats: optional per-level outputs of the generators' own text encoders, computed ahead of time."""
    tabs = device_tables(variant, targets[0].device)
    outs, prev, last = [], None, None
    for k, g in enumerate(gens):
        pre = ops.pre_seq(targets[k], prev, tabs[k][1], tabs[k][2], n_pre)
        kw = {}
        if eps is not None:
            kw["_eps"] = eps[k]
        if text_feats is not None:
            kw["_text_feat"] = text_feats[k]
        out, z, mu, lv = g(pre, in_text, blends[k], vid, **kw)
        outs.append(out)
        prev, last = out, (z, mu, lv)
    return outs, last
