"""Baseline multimodal_context path: the CUDA ``PoseGenerator`` / ``WavEncoder`` / ``ConvDiscriminator``
and ``train_iter_gan`` against the fixture produced by the UNMODIFIED reference (oracle/make_golden_gan.py ->
tests/golden/step_gan.pt)."""
import os

import pytest
import torch

from ha2g_b200.constants import make_args
from ha2g_b200.model.vocab import make_speaker_vocab
from ha2g_b200.synthetic import _gen, det_fill, make_audio, make_batch, make_embedding
from helpers import assert_close, assert_params_close, assert_summary_close, randn, summary_scale

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_gan.pt")
DEV = "cuda:0"


def _build(g, device):
    from ha2g_b200.model.multimodal_context_net import ConvDiscriminator, PoseGenerator
    args = make_args("gesture")
    spk = make_speaker_vocab(g["n_spk"])
    emb = make_embedding(g["n_words"], 300, 1).numpy()
    G = det_fill(PoseGenerator(args, 27, g["n_words"], 300, emb, z_obj=spk), g["fill_seeds"]["gen"]).to(device)
    D = det_fill(ConvDiscriminator(27), g["fill_seeds"]["dis"]).to(device)
    return args, G, D


def test_state_dict_contract_baseline():
    g = torch.load(GOLD, weights_only=False)
    _, G, D = _build(g, "cpu")
    assert [(k, tuple(v.shape)) for k, v in G.state_dict().items()] == g["gen_keys"]
    assert [(k, tuple(v.shape)) for k, v in D.state_dict().items()] == g["dis_keys"]


@pytest.mark.gpu
def test_wav_encoder_vs_reference():
    """4 x Conv1d(k=15, strides 5/6/6/6, padding 1600) + train-mode BatchNorm1d + LeakyReLU(0.3) on raw audio."""
    from ha2g_b200.model.multimodal_context_net import WavEncoder
    g = torch.load(GOLD, weights_only=False)
    w = g["wav"]
    audio = torch.stack([make_audio(g["n_audio"], w["audio_seed"] + i) for i in range(g["B"])]).to(DEV).requires_grad_(True)
    W = det_fill(WavEncoder(), w["fill"]).to(DEV).train(True)
    y = W(audio)
    assert tuple(y.shape) == (g["B"], 34, 32)
    assert_close(y, w["y"], "WavEncoder forward", 1e-3)
    (y * randn(tuple(y.shape), w["gy_seed"], "gy").to(DEV)).sum().backward()
    floor = summary_scale(w["grads"].values())
    for name, p in W.named_parameters():
        assert_summary_close(p.grad, w["grads"][name], f"WavEncoder grad {name}", 2e-3, floor=floor)
    assert_summary_close(audio.grad, w["dx"], "WavEncoder d(audio)", 2e-3)
    for name, b in W.named_buffers():
        if "num_batches" not in name:
            assert_close(b, w["buffers"][name], f"WavEncoder buffer {name}", 1e-4)


@pytest.mark.gpu
def test_train_iter_gan_vs_reference():
    from ha2g_b200 import rng
    from ha2g_b200.train_eval.train_gan import train_iter_gan
    g = torch.load(GOLD, weights_only=False)
    args, G, D = _build(g, DEV)
    lr = args.learning_rate
    g_opt = torch.optim.Adam(G.parameters(), lr=lr, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(D.parameters(), lr=lr * args.discriminator_lr_weight, betas=(0.5, 0.999))
    B = g["B"]
    for si, rec in enumerate(g["steps"]):
        batch = {k: v.to(DEV) for k, v in make_batch("gesture", B, g["n_words"], g["n_spk"], seed=rec["batch_seed"]).items()}
        aud = torch.stack([make_audio(g["n_audio"], rec["audio_seed"] + i) for i in range(B)]).to(DEV)
        n_draws = 3 if rec["epoch"] > args.loss_warmup else 2
        feed = rng.ListFeed([randn((B, 16), rec["eps_seed"], f"eps{i}") for i in range(n_draws)])
        with rng.override(randn_fn=feed, randperm_fn=lambda n, p=rec["perm"]: p.clone(), dropout=False):
            ret = train_iter_gan(args, rec["epoch"], batch["in_text_padded"], aud, batch["target"], batch["vid"], G, D, g_opt, d_opt)
        assert set(ret) == set(rec["ret"]), (sorted(ret), sorted(rec["ret"]))
        tol = 1e-3 if si == 0 else 1e-2     # the second step starts from Adam-updated parameters (sign-like first update)
        bad = {k: (ret[k], rec["ret"][k]) for k in ret if abs(ret[k] - rec["ret"][k]) > tol * max(1.0, abs(rec["ret"][k]))}
        assert not bad, f"step {si} loss dict mismatch (cuda, reference): {bad}"
        if si == 0:
            floor = summary_scale(rec["grads"].values())
            named = dict(G.named_parameters())
            errs = []
            for name, summ in rec["grads"].items():
                try:
                    assert_summary_close(named[name].grad, summ, f"gan step0 grad {name}", 2e-3, floor=floor)
                except AssertionError as e:
                    errs.append(str(e))
            assert not errs, "\\n".join(errs[:10]) + f"\\n({len(errs)} gradient tensors out of tolerance)"
            sd = G.state_dict()
            for name, summ in rec["gen"].items():
                if "running_" in name or "num_batches" in name:
                    assert_summary_close(sd[name].float(), summ, f"gan step0 {name}", 1e-3)
                else:
                    # a conv bias followed directly by BatchNorm has a mathematically zero gradient: both sides hold
                    # rounding noise there and Adam's sign-like first update moves it by +-lr arbitrarily
                    noise_only = name in ("audio_encoder.feat_extractor.0.bias", "audio_encoder.feat_extractor.3.bias",
                                          "audio_encoder.feat_extractor.6.bias")
                    assert_params_close(sd[name], summ, f"gan step0 {name}", lr, 1, 1.0 if noise_only else 0.02)
