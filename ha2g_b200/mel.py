"""GPU log-mel spectrogram: drop-in for ``utils.data_utils.extract_melspectrogram`` (scripts/utils/data_utils.py:34-38)
and ``calc_spectrogram_length_from_motion_length`` (:41-43), running on csrc/mel.cu.

The only host-side arithmetic is the constant Slaney mel filterbank table (128 x 513 floats, built once).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .ops import _call, _p, _st

SR, N_FFT, HOP, N_MELS = 16000, 1024, 512, 128
_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = math.log(6.4) / 27.0


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    lin = _F_SP * m
    return np.where(m >= _MIN_LOG_MEL, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), lin)


def _hz_to_mel(f: float) -> float:
    return f / _F_SP if f < _MIN_LOG_HZ else _MIN_LOG_MEL + math.log(f / _MIN_LOG_HZ) / _LOGSTEP


def slaney_filterbank() -> np.ndarray:
    """128 triangular, area-normalised Slaney-scale bands over the 513 rfft bins of a 1024-point FFT at 16 kHz
    (librosa.filters.mel defaults as used by melspectrogram(sr=16000, n_fft=1024))."""
    edges = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(SR / 2.0), N_MELS + 2))
    bins = np.linspace(0.0, SR / 2.0, 1 + N_FFT // 2)
    fb = np.zeros((N_MELS, bins.size), dtype=np.float64)
    for i in range(N_MELS):
        lo, ce, hi = edges[i], edges[i + 1], edges[i + 2]
        up = (bins - lo) / (ce - lo)
        down = (hi - bins) / (hi - ce)
        fb[i] = np.clip(np.minimum(up, down), 0.0, None) * (2.0 / (hi - lo))
    return fb.astype(np.float32)


_tables: Dict[str, tuple] = {}


def _device_tables(device):
    key = str(device)
    if key not in _tables:
        fb = slaney_filterbank()
        start, length = [], []
        for row in fb:
            nz = np.nonzero(row)[0]
            if nz.size == 0:
                start.append(0)
                length.append(0)
            else:
                start.append(int(nz[0]))
                length.append(int(nz[-1] - nz[0] + 1))
        n = np.arange(N_FFT, dtype=np.float64)
        j = np.arange(N_FFT // 2, dtype=np.float64)
        hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)             # periodic Hann (scipy.signal.get_window('hann', 1024))
        tw = np.concatenate([hann, np.cos(2.0 * np.pi * j / N_FFT), -np.sin(2.0 * np.pi * j / N_FFT)]).astype(np.float32)
        _tables[key] = (torch.from_numpy(fb).to(device), torch.tensor(start, dtype=torch.int32, device=device),
                        torch.tensor(length, dtype=torch.int32, device=device), torch.from_numpy(tw).to(device))
    return _tables[key]


def calc_spectrogram_length_from_motion_length(n_frames: int, fps: float) -> int:
    return int(round((n_frames / fps * 16000 - 1024) / 512 + 1))


def extract_melspectrogram(y: torch.Tensor, n_out: int | None = None) -> torch.Tensor:
    """y: CUDA float32 [n_samples] or [B, n_samples] at 16 kHz -> log-mel [128, frames] / [B, 128, frames] in dB,
    values rounded through float16 like the reference's ``.astype('float16')`` (returned as float32, the dtype the
    training loop feeds the encoder: ``in_spec.float()``, scripts/train.py:261)."""
    if not y.is_cuda:
        raise RuntimeError("ha2g_b200.mel runs on CUDA tensors only")
    squeeze = y.dim() == 1
    y2 = (y[None] if squeeze else y).contiguous().float()
    B, n = y2.shape
    frames = 1 + n // HOP
    n_out = frames if n_out is None else n_out
    fb, st, ln, tw = _device_tables(y2.device)
    melpow = torch.empty((B, N_MELS, frames), device=y2.device, dtype=torch.float32)
    cmax = torch.empty((B,), device=y2.device, dtype=torch.int32)
    out = torch.empty((B, N_MELS, n_out), device=y2.device, dtype=torch.float32)
    _call("ha2g_logmel", _p(y2), B, n, _p(fb), _p(st), _p(ln), _p(tw), _p(melpow), _p(cmax), _p(out), n_out, _st())
    return out[0] if squeeze else out
