"""Log-mel front-end (row K).  The oracle (oracle/mel_oracle.py) is 'parity unpinned' (librosa is absent); it is
cross-checked against torchaudio's independent implementation on CPU, and the CUDA kernel against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import mel_oracle as M  # noqa: E402

from ha2g_b200.synthetic import make_audio  # noqa: E402


def test_frame_count_formula():
    # lmdb_data_loader.py:69 / data_utils.py:41-43: 34 frames @ 15 fps -> 36267 samples -> 71 STFT frames, 70 used
    n = int(round(34 / 15 * 16000))
    assert n == 36267 and M.n_frames(n) == 71 and M.calc_spectrogram_length_from_motion_length(34, 15) == 70
    from ha2g_b200 import mel
    assert mel.calc_spectrogram_length_from_motion_length(34, 15) == 70


def test_product_filterbank_equals_oracle():
    from ha2g_b200 import mel
    a, b = mel.slaney_filterbank(), M.mel_filterbank()
    assert a.shape == b.shape == (128, 513)
    assert np.abs(a - b).max() <= 1e-7 * np.abs(b).max()


def test_oracle_vs_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    y = make_audio(36267, 0).numpy()
    tm = torchaudio.transforms.MelSpectrogram(16000, n_fft=1024, hop_length=512, n_mels=128, power=2, center=True,
                                              pad_mode="reflect", norm="slaney", mel_scale="slaney")
    ref = tm(torch.from_numpy(y)).numpy()
    mine = M.power_mel(y)
    assert np.abs(mine - ref).max() <= 1e-4 * np.abs(ref).max()
    db = M.extract_melspectrogram(y)
    assert db.dtype == np.float16 and db.shape == (128, 71) and db.max() == 0 and db.min() >= -80


@pytest.mark.gpu
@pytest.mark.parametrize("n_samples,batch", [(36267, 3), (160000, 1), (513, 2)])
def test_cuda_logmel_vs_oracle(n_samples, batch):
    from ha2g_b200 import mel
    ys = torch.stack([make_audio(n_samples, 10 + i) for i in range(batch)])
    out = mel.extract_melspectrogram(ys.cuda()).cpu().numpy()
    for i in range(batch):
        ref = M.extract_melspectrogram(ys[i].numpy()).astype(np.float32)
        assert out[i].shape == ref.shape
        # values are fp16-rounded dB: allow one fp16 ulp at |x| <= 80 (0.0625) on a handful of rounding ties
        diff = np.abs(out[i] - ref)
        assert diff.max() <= 0.0626, diff.max()
        assert (diff > 0).mean() <= 0.02, (diff > 0).mean()
    # the training loop uses the first 70 of the 71 frames (calc_spectrogram_length_from_motion_length)
    if n_samples == 36267:
        sl = mel.extract_melspectrogram(ys.cuda(), n_out=70).cpu().numpy()
        assert sl.shape == (batch, 128, 70) and np.array_equal(sl, out[:, :, :70])
