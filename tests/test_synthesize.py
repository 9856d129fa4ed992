"""Inference loop (row L): ha2g_b200.synthesize.generate_gestures_hierarchy against the fixture produced by the
UNMODIFIED reference loop (scripts/synthesize_expressive_hierarchy.py:36-259, see oracle/make_golden.py)."""
import math
import os

import numpy as np
import pytest
import torch

from ha2g_b200.constants import make_args
from ha2g_b200.model.hierarchy_net import Hierarchical_PoseGenerator, Hierarchical_WavEncoder
from ha2g_b200.model.vocab import Vocab, make_speaker_vocab
from ha2g_b200.synthetic import det_fill, make_audio, make_embedding
from helpers import assert_close, randn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inference.pt")


def test_window_plan_index_contract():
    """10 min of 16 kHz audio -> 300 windows; the spectrogram start index keeps the reference's quirk
    (scaled by the 128 mel rows, synthesize_expressive_hierarchy.py:84)."""
    from ha2g_b200.synthesize import window_plan
    plan = window_plan(9_600_000, 16000, 34, 4, 15)
    assert len(plan) == math.ceil((600 - 34 / 15) / 2.0) + 1 == 300
    assert plan[0] == (0.0, 34 / 15, 0)
    assert plan[150][2] == math.floor(150 * 2.0 / 600 * 128) == 64
    assert len(window_plan(16000, 16000, 34, 4, 15)) == 1  # clip shorter than one unit


def test_word_placement():
    from ha2g_b200.synthesize import place_words
    lang = Vocab("w")
    for w in ("hello", "world"):
        lang.index_word(w)
    words = [["hello", 0.05, 0.2], ["unknown", 1.0, 1.1], ["world", 2.2, 2.4], ["late", 9.0, 9.1]]
    ext = place_words(words, 0.0, 34 / 15, 34, lang)
    assert ext[0] == lang.get_word_index("hello") and ext[15] == Vocab.UNK_token and ext[33] == lang.get_word_index("world")
    assert (ext != 0).sum() == 3


@pytest.mark.gpu
def test_generate_gestures_matches_reference_loop():
    from ha2g_b200 import rng
    from ha2g_b200.synthesize import generate_gestures_hierarchy
    g = torch.load(GOLD, weights_only=False)
    dev = "cuda:0"
    args = make_args("expressive")
    spk = make_speaker_vocab(g["n_spk"])
    emb = make_embedding(g["n_words"], 300, 1).numpy()
    lang = Vocab("words")
    for w in g["vocab_words"]:
        lang.index_word(w)
    dims = (24, 30, 36, 66, 96, 126)
    gens = [det_fill(Hierarchical_PoseGenerator(args, d, g["n_words"], 300, emb, z_obj=spk), g["fill_seeds"]["gens"] + i)
            .to(dev).train(False) for i, d in enumerate(dims)]
    A = det_fill(Hierarchical_WavEncoder(args, spk, pose_level=6, nOut=32), g["fill_seeds"]["audio"]).to(dev).train(False)
    audio = make_audio(g["n_samples"], g["audio_seed"]).numpy()
    targets = [randn((1, 34, d), g["target_seed"], f"t{d}") * 0.1 for d in dims]
    for fade, key in ((False, "out"), (True, "out_fade")):
        feed = rng.ListFeed([randn((1, 16), g["eps_seed"], f"eps{i}") for i in range(18)])
        with rng.override(randn_fn=feed):
            out = generate_gestures_hierarchy(args, *gens, A, lang, audio, g["words"], *[t.clone() for t in targets],
                                              vid=g["vid"], fade_out=fade)
        assert out.shape == tuple(g[key].shape), (out.shape, g[key].shape)
        # 2e-3: the CUDA log-mel may differ from the oracle mel by one fp16 ulp in a few bins (tests/test_mel.py)
        assert_close(torch.from_numpy(out), g[key], f"inference {key}", 2e-3)


@pytest.mark.gpu
def test_window_graph_matches_eager():
    """The CUDA-graph replay of the window body (windows >= 2) must reproduce the eager loop: same clip, same
    (stateless, cyclic) reparameterisation noise, graph on vs off."""
    from ha2g_b200 import rng, synthesize
    dev = "cuda:0"
    args = make_args("expressive")
    spk = make_speaker_vocab(5)
    emb = make_embedding(60, 300, 1).numpy()
    lang = Vocab("words")
    for w in ("alpha", "beta", "gamma"):
        lang.index_word(w)
    dims = (24, 30, 36, 66, 96, 126)
    gens = [det_fill(Hierarchical_PoseGenerator(args, d, 60, 300, emb, z_obj=spk), 20 + i).to(dev).train(False)
            for i, d in enumerate(dims)]
    A = det_fill(Hierarchical_WavEncoder(args, spk, pose_level=6, nOut=32), 31).to(dev).train(False)
    audio = make_audio(16000 * 14, 3).numpy()             # 14 s -> 7 windows
    words = [["alpha", 0.5, 0.9], ["beta", 3.1, 3.4], ["gamma", 7.7, 8.0], ["alpha", 11.0, 11.3]]
    targets = [randn((1, 34, d), 9, f"t{d}") * 0.1 for d in dims]
    noise = [randn((1, 16), 10, f"e{i}").to(dev) for i in range(6)]
    outs = []
    for graph_on in (False, True):
        calls = [0]

        def cyc(shape):
            calls[0] += 1
            return noise[(calls[0] - 1) % 6]
        synthesize._GRAPH = graph_on
        try:
            with rng.override(randn_fn=cyc, graph_safe=True):
                outs.append(synthesize.generate_gestures_hierarchy(args, *gens, A, lang, audio, words,
                                                                   *[t.clone() for t in targets], vid=2))
        finally:
            synthesize._GRAPH = True
    assert outs[0].shape == outs[1].shape and outs[0].shape[0] == 7 * 30 + 4
    assert_close(torch.from_numpy(outs[1]), torch.from_numpy(outs[0]), "window graph vs eager", 1e-5)
