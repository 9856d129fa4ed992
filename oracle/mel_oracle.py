"""CPU oracle for the log-mel front-end.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED: the arithmetic lives in librosa (third-party, un-vendored, unpinned in the reference's
requirements.txt:6,17; era implies 0.8-0.9) which is absent from /root/reference and from this image, and the
reference holds no golden vector for it.  This file restates librosa's published algorithm for the exact call
    librosa.feature.melspectrogram(y=y, sr=16000, n_fft=1024, hop_length=512, power=2)
    librosa.power_to_db(melspec, ref=np.max)  ->  astype('float16')
made at scripts/utils/data_utils.py:34-38 (= data_utils_expressive.py:84-88, dataset_script/script/make_ted_dataset.py:121-122):
periodic Hann window, center=True with reflect padding (librosa < 0.10), 128 Slaney-scale mel bands with Slaney
area normalisation, fmin 0, fmax sr/2, amin 1e-10, top_db 80.  tests/test_mel.py cross-checks it against
torchaudio.transforms.MelSpectrogram with the matching options (an independent implementation present in the image).
"""
from __future__ import annotations

import numpy as np

SR, N_FFT, HOP, N_MELS = 16000, 1024, 512, 128


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney') -> float32 [n_mels, 1 + n_fft//2]."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def n_frames(n_samples: int) -> int:
    return 1 + n_samples // HOP


def power_mel(y: np.ndarray) -> np.ndarray:
    """|STFT|^2 projected on the mel basis: float32 [128, 1 + len(y)//512]."""
    y = np.asarray(y, dtype=np.float32)
    ypad = np.pad(y, N_FFT // 2, mode="reflect")
    n = np.arange(N_FFT)
    window = (0.5 - 0.5 * np.cos(2 * np.pi * n / N_FFT)).astype(np.float32)  # scipy get_window('hann', fftbins=True)
    nf = n_frames(len(y))
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(nf)[:, None]
    frames = ypad[idx] * window[None, :]
    spec = np.fft.rfft(frames.astype(np.float32), axis=1).astype(np.complex64)
    power = (np.abs(spec) ** 2).astype(np.float32)          # [frames, 513]
    return mel_filterbank() @ power.T                        # [128, frames]


def power_to_db(S: np.ndarray, amin=1e-10, top_db=80.0) -> np.ndarray:
    ref = np.max(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref))
    return np.maximum(log_spec, log_spec.max() - top_db)


def extract_melspectrogram(y: np.ndarray) -> np.ndarray:
    """data_utils.py:34-38 restated: float16 [128, frames]."""
    return power_to_db(power_mel(y)).astype(np.float16)


def calc_spectrogram_length_from_motion_length(n_frames_motion: int, fps: float) -> int:
    """data_utils.py:41-43."""
    return int(round((n_frames_motion / fps * 16000 - 1024) / 512 + 1))
