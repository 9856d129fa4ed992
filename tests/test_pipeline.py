"""GPU input pipeline: device-side word placement is bit-exact against the host restatement of
SpeechMotionDataset.extend_word_seq (itself pinned to the reference logic in tests/test_data.py), the on-device log-mel
equals the mel path, and a batch built this way drives a training step."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _clips(rs, B, lang_words):
    seqs, t0s, t1s = [], [], []
    for b in range(B):
        t0 = float(rs.uniform(0, 100))
        dur = 34 / 15
        words, t = [], t0 - 0.3            # one word may start before the clip (clamped to frame 0)
        while t < t0 + dur + 0.5:          # and some after its end (dropped)
            words.append([lang_words[rs.randint(len(lang_words))], t, t + 0.2])
            t += float(rs.choice([0.0, 0.01, 0.07, 0.3, 0.9]))   # 0.0 / 0.01: collisions on one frame (last one wins)
        seqs.append(words)
        t0s.append(t0)
        t1s.append(t0 + dur)
    return seqs, t0s, t1s


def test_place_words_bit_exact():
    from ha2g_b200 import data, pipeline
    from ha2g_b200.model.vocab import Vocab
    lang = Vocab("words")
    words = [f"w{i}" for i in range(50)]
    for w in words:
        lang.index_word(w)
    rs = np.random.RandomState(0)
    seqs, t0s, t1s = _clips(rs, 64, words + ["not-in-vocab"])
    seqs[3] = []                                                          # a clip without words
    got = pipeline.place_words(pipeline.pack_words(lang, seqs, t0s, t1s), 34, DEV).cpu()
    ref = torch.stack([data.extend_word_seq(lang, s, a, b, 34) for s, a, b in zip(seqs, t0s, t1s)])
    assert got.dtype == torch.int64 and torch.equal(got, ref)


def test_build_batch_feeds_a_training_step():
    import mel_oracle as M
    from helpers import build_modules
    from ha2g_b200 import pipeline, rng
    from ha2g_b200.model.vocab import Vocab
    from ha2g_b200.synthetic import make_audio, make_batch
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    B = 4
    lang = Vocab("words")
    words = [f"w{i}" for i in range(50)]
    for w in words:
        lang.index_word(w)
    rs = np.random.RandomState(1)
    seqs, t0s, t1s = _clips(rs, B, words)
    n = int(round(34 / 15 * 16000))
    audio = torch.stack([make_audio(n, 20 + i) for i in range(B)]).pin_memory()
    in_text, in_spec = pipeline.build_batch(lang, audio, seqs, t0s, t1s, device=DEV)
    assert in_text.shape == (B, 34) and in_spec.shape == (B, 128, 70) and in_spec.is_cuda
    ref = M.extract_melspectrogram(audio[0].numpy()).astype(np.float32)[:, :70]
    diff = np.abs(in_spec[0].cpu().numpy() - ref)
    assert diff.max() <= 0.0626 and (diff > 0).mean() <= 0.02
    args, gens, D, A, T = build_modules("gesture", lang.n_words, 5, {"gens": 20, "dis": 30, "audio": 31, "text": 32}, DEV)
    lr = args.learning_rate
    mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
    b = make_batch("gesture", B, lang.n_words, 5, seed=3)
    with rng.override(dropout=False):
        ret = train_iter_hierarchy(args, 11, in_text, in_spec, b["target"].to(DEV), b["vid"].to(DEV), *gens, D, A, T,
                                   *[mk(g) for g in gens], mk(D, lr * args.discriminator_lr_weight), mk(A), mk(T))
    assert all(np.isfinite(v) for v in ret.values()) and ret["loss"] > 0
