"""Input pipeline on the GPU: from raw clips to the tensors ``train_iter_*`` consumes.

The reference prepares ``in_spec`` offline (librosa log-mel at dataset-build time, dataset_script/script/
make_ted_dataset.py:119-123) and ``in_text_padded`` per sample on DataLoader workers (``extend_word_seq``,
scripts/data_loader/lmdb_data_loader_expressive.py:116-141).  Here both are produced on the device from what a shard
holds -- 16 kHz audio and word timings -- right before the step:

    raw audio [B, 36267]  --csrc/mel.cu-->        in_spec [B, 128, 70]   (fp16-rounded dB, like the stored spectrograms)
    word ids + start times  --place_words-->      in_text_padded [B, 34] int64   (bit-exact frame indices, float64 timing)

so a training batch costs one H2D copy of 145 KB of audio per clip instead of a CPU STFT, and the loader workers only slice.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import mel
from .ops import _call, _p, _st


def pack_words(lang, word_seqs: Sequence[Sequence], start_times: Sequence[float], end_times: Sequence[float]):
    """Host side of the word placement: flat (ids, start times, offsets) arrays for a batch of clips.  ``word_seqs[b]`` is
    the clip's list of (word, start_s, end_s); ids come from ``lang.get_word_index`` (UNK for unknown words)."""
    ids, starts, off = [], [], [0]
    for words in word_seqs:
        for w in words:
            ids.append(lang.get_word_index(w[0]))
            starts.append(float(w[1]))
        off.append(len(ids))
    return (torch.tensor(ids, dtype=torch.int64), torch.tensor(starts, dtype=torch.float64),
            torch.tensor(off, dtype=torch.int32), torch.tensor(list(start_times), dtype=torch.float64),
            torch.tensor(list(end_times), dtype=torch.float64))


def place_words(packed: Tuple[torch.Tensor, ...], n_frames: int, device) -> torch.Tensor:
    """-> in_text_padded [B, n_frames] int64 on ``device`` (``extend_word_seq`` for every clip of the batch)."""
    ids, starts, off, t0, t1 = [t.to(device, non_blocking=True) for t in packed]
    B = t0.numel()
    out = torch.empty((B, n_frames), dtype=torch.int64, device=device)
    _call("ha2g_place_words", _p(ids), _p(starts), _p(off), _p(t0), _p(t1), B, n_frames, _p(out), _st())
    return out


def build_batch(lang, audio: torch.Tensor, word_seqs, start_times, end_times, n_frames: int = 34, fps: float = 15.0,
                device="cuda:0") -> Tuple[torch.Tensor, torch.Tensor]:
    """audio: [B, n_samples] float32 (pinned host or device), ``int(round(n_frames / fps * 16000))`` samples per clip.
    -> (in_text_padded [B, n_frames] int64, in_spec [B, 128, spec_len] float32), both on ``device``."""
    spec_len = mel.calc_spectrogram_length_from_motion_length(n_frames, fps)
    a = audio.to(device, non_blocking=True)
    in_spec = mel.extract_melspectrogram(a, n_out=spec_len)
    in_text = place_words(pack_words(lang, word_seqs, start_times, end_times), n_frames, device)
    return in_text, in_spec
