"""Checkpoint compatibility: the reference's checkpoint dict layout, both directions.
CPU only; the direction "written by the reference" needs /root/reference and is skipped where it is absent."""
import argparse
import os
import sys
import types

import numpy as np
import pytest
import torch

from ha2g_b200 import checkpoint
from ha2g_b200.constants import make_args
from ha2g_b200.model.hierarchy_net import (Hierarchical_ConvDiscriminator, Hierarchical_PoseGenerator,
                                           Hierarchical_WavEncoder, TextEncoderTCN)
from ha2g_b200.model.vocab import Vocab, make_speaker_vocab
from ha2g_b200.synthetic import det_fill

REF = "/root/reference/scripts"


def _our_world(variant, n_words=40, n_spk=5):
    args = make_args(variant)
    args.wordembed_dim = 300
    lang = Vocab("words")
    for i in range(n_words - lang.n_words):
        lang.index_word(f"w{i}")
    lang.word_embedding_weights = np.random.RandomState(0).normal(0, 0.05, (lang.n_words, 300)).astype(np.float32)
    spk = make_speaker_vocab(n_spk)
    dims = checkpoint.LEVEL_DIMS[3 if variant == "gesture" else 6]
    gens = [det_fill(Hierarchical_PoseGenerator(args, d, lang.n_words, 300, lang.word_embedding_weights, z_obj=spk), 20 + i)
            for i, d in enumerate(dims)]
    D = det_fill(Hierarchical_ConvDiscriminator(dims[-1]), 30)
    A = det_fill(Hierarchical_WavEncoder(args, spk, pose_level=len(dims), nOut=32), 31)
    T = det_fill(TextEncoderTCN(args, lang.n_words, 300, pre_trained_embedding=lang.word_embedding_weights, dropout=0.3), 32)
    return args, lang, spk, dims, gens, D, A, T


@pytest.mark.parametrize("variant", ["gesture", "expressive"])
def test_layout_and_roundtrip(tmp_path, variant):
    args, lang, spk, dims, gens, D, A, T = _our_world(variant)
    path = str(tmp_path / "ck.bin")
    opt = torch.optim.Adam(D.parameters(), lr=1e-4)
    state = checkpoint.save_checkpoint_hierarchy(path, args, 7, lang, spk, dims[-1], gens, D, A, T, optimizers={"dis": opt})
    want = {"args", "epoch", "lang_model", "speaker_model", "pose_dim", "dis_dict", "audio_dict", "text_dict", "optim_dicts"}
    want |= {f"gen_dict_{k}" for k in range(1, len(dims) + 1)}
    assert set(state) == want                                   # train_expressive.py:299-304 (+ our optimizer extra)
    ck = checkpoint.read_checkpoint(path)
    assert ck["epoch"] == 7 and ck["pose_dim"] == dims[-1] and ck["lang_model"].n_words == lang.n_words
    out = checkpoint.load_checkpoint_hierarchy(path, _device="cpu")
    assert len(out) == len(dims) + 6
    loaded_gens, loaded_audio = out[1:1 + len(dims)], out[1 + len(dims)]
    for g0, g1 in zip(gens, loaded_gens):
        assert not g1.training
        for (k0, v0), (k1, v1) in zip(g0.state_dict().items(), g1.state_dict().items()):
            assert k0 == k1 and torch.equal(v0, v1)
    for (k0, v0), (k1, v1) in zip(A.state_dict().items(), loaded_audio.state_dict().items()):
        assert k0 == k1 and torch.equal(v0, v1)
    assert isinstance(out[-4], torch.nn.L1Loss) and out[-1] == dims[-1]


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")
def test_reads_reference_written_checkpoint_and_vice_versa(tmp_path):
    """A checkpoint written the reference's way from the REFERENCE's modules (pickled reference Vocab objects) loads
    into our modules value for value, without the reference tree on sys.path at load time; and a checkpoint written by
    us loads into the reference's modules with its own load_state_dict calls (train_utils_expressive.py:186-192)."""
    variant = "gesture"
    args, lang, spk, dims, gens, D, A, T = _our_world(variant)
    sys.modules.setdefault("fasttext", types.ModuleType("fasttext"))
    sys.path.insert(0, REF)
    try:
        import importlib
        for m in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[m]
        ref_net = importlib.import_module("model.hierarchy_net")
        ref_vocab = importlib.import_module("model.vocab")
        rlang = ref_vocab.Vocab("words")
        for i in range(lang.n_words - rlang.n_words):
            rlang.index_word(f"w{i}")
        rlang.word_embedding_weights = lang.word_embedding_weights.copy()
        rspk = ref_vocab.Vocab("vid", insert_default_tokens=False)
        for i in range(5):
            rspk.index_word(f"spk{i}")
        rargs = argparse.Namespace(**vars(args))
        torch.manual_seed(1)
        rgens = [ref_net.Hierarchical_PoseGenerator(rargs, d, rlang.n_words, 300, rlang.word_embedding_weights.copy(), rspk)
                 for d in dims]
        rA = ref_net.Hierarchical_WavEncoder(rargs, rspk, pose_level=3, nOut=32)
        rT = ref_net.TextEncoderTCN(rargs, rlang.n_words, 300, pre_trained_embedding=rlang.word_embedding_weights.copy(),
                                    dropout=0.3)
        rD = ref_net.Hierarchical_ConvDiscriminator(dims[-1])
        path = str(tmp_path / "ref.bin")
        torch.save({"args": rargs, "epoch": 3, "lang_model": rlang, "speaker_model": rspk, "pose_dim": dims[-1],
                    "gen_dict_1": rgens[0].state_dict(), "gen_dict_2": rgens[1].state_dict(),
                    "gen_dict_3": rgens[2].state_dict(), "dis_dict": rD.state_dict(), "audio_dict": rA.state_dict(),
                    "text_dict": rT.state_dict()}, path)
        # the other direction while the reference modules exist: our file -> reference modules
        ours = str(tmp_path / "ours.bin")
        checkpoint.save_checkpoint_hierarchy(ours, args, 1, lang, spk, dims[-1], gens, D, A, T)
        ck = torch.load(ours, map_location="cpu", weights_only=False)
        for k, rg in enumerate(rgens, start=1):
            rg.load_state_dict(ck[f"gen_dict_{k}"])
            for (k0, v0), (k1, v1) in zip(gens[k - 1].state_dict().items(), rg.state_dict().items()):
                assert k0 == k1 and torch.equal(v0, v1)
        rA.load_state_dict(ck["audio_dict"]); rT.load_state_dict(ck["text_dict"]); rD.load_state_dict(ck["dis_dict"])
        ref_sd = [dict(g.state_dict()) for g in rgens]   # now equal to ours; reload the reference-written file below
    finally:
        sys.path.remove(REF)
        for m in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[m]
    # reference tree gone from sys.path: the pickled reference Vocab must still resolve
    out = checkpoint.load_checkpoint_hierarchy(path, _device="cpu")
    ck = checkpoint.read_checkpoint(path)
    assert type(out[-3]).__name__ == "Vocab" and out[-3].n_words == lang.n_words and out[-2].n_words == 6
    for k, g in enumerate(out[1:4], start=1):
        for (k0, v0), (k1, v1) in zip(ck[f"gen_dict_{k}"].items(), g.state_dict().items()):
            assert k0 == k1 and torch.equal(v0, v1), (k, k0)
    for (k0, v0), (k1, v1) in zip(ck["audio_dict"].items(), out[4].state_dict().items()):
        assert k0 == k1 and torch.equal(v0, v1), k0


@pytest.mark.gpu
def test_resume_reproduces_the_next_step(tmp_path):
    """Save after two training steps (weights, BatchNorm buffers, Adam moments), resume into freshly built modules and
    optimizers, and take the third step in both worlds: the step is deterministic, so losses and parameters must agree
    exactly -- through the eager path and through a re-captured CUDA graph."""
    from helpers import build_modules
    from ha2g_b200 import checkpoint, graph_step, rng
    from ha2g_b200.model.vocab import make_speaker_vocab
    from ha2g_b200.synthetic import make_batch
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    dev = "cuda:0"
    seeds = {"gens": 20, "dis": 30, "audio": 31, "text": 32}

    def world():
        args, gens, D, A, T = build_modules("gesture", 60, 5, seeds, dev)
        lr = args.learning_rate
        mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
        opts = {f"g{k + 1}": mk(g) for k, g in enumerate(gens)}
        opts.update(dis=mk(D, lr * args.discriminator_lr_weight), audio=mk(A), text=mk(T))
        return args, gens, D, A, T, opts

    gtor = torch.Generator().manual_seed(5)
    noise = [torch.randn((4, 16), generator=gtor).to(dev) for _ in range(9)]
    perm = torch.randperm(4, generator=gtor).to(dev)

    def step(w, i, calls=[0]):
        args, gens, D, A, T, opts = w
        b = {k: v.to(dev) for k, v in make_batch("gesture", 4, 60, 5, seed=700 + i).items()}
        cnt = [0]

        def randn_fn(shape):
            cnt[0] += 1
            return noise[(cnt[0] - 1) % 9]
        with rng.override(randn_fn=randn_fn, randperm_fn=lambda n: perm, dropout=False, graph_safe=True):
            return train_iter_hierarchy(args, 11, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *gens, D, A, T,
                                        *[opts[f"g{k + 1}"] for k in range(3)], opts["dis"], opts["audio"], opts["text"])

    graph_step.reset()
    graph_step.enable(False)
    try:
        w1 = world()
        for i in range(2):
            step(w1, i)
        path = str(tmp_path / "ckpt.bin")
        checkpoint.save_checkpoint_hierarchy(path, w1[0], 7, None, make_speaker_vocab(5), 27, w1[1], w1[2], w1[3], w1[4],
                                             optimizers=w1[5])
        r1 = step(w1, 2)
        w2 = world()
        assert checkpoint.resume_training(path, w2[1], w2[2], w2[3], w2[4], optimizers=w2[5]) == 7
        r2 = step(w2, 2)
        assert r1 == r2, (r1, r2)
        for m1, m2 in zip(w1[1] + [w1[2], w1[3], w1[4]], w2[1] + [w2[2], w2[3], w2[4]]):
            for (n, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
                assert torch.equal(a, b), n
    finally:
        graph_step.enable(True)
        graph_step.reset()
