"""TEST INFRASTRUCTURE (not product code).  Fixture for the baseline multimodal_context path: runs the
UNMODIFIED reference ``model.multimodal_context_net`` modules and ``train_eval.train_gan.train_iter_gan`` on the CPU
(same RNG control as oracle/make_golden.py: dropout p = 0, injected reparameterisation noise and speaker permutation,
parameters from ha2g_b200.synthetic.det_fill) and stores module outputs / gradient summaries, the returned loss dicts of
two consecutive steps (epoch 0 and epoch 11) and post-step parameter summaries.  Writes tests/golden/step_gan.pt.

    python oracle/make_golden_gan.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as MG          # noqa: E402  (sets up the reference import path and the third-party stubs)

import torch                      # noqa: E402

import model.embedding_net        # noqa: E402
from model.multimodal_context_net import ConvDiscriminator, PoseGenerator, WavEncoder   # noqa: E402
import train_eval.train_gan as tg  # noqa: E402

from ha2g_b200.constants import make_args                                              # noqa: E402
from ha2g_b200.synthetic import det_fill, make_audio, make_batch, make_embedding, sample_tensor, _gen   # noqa: E402

N_WORDS, N_SPK, B = 60, 5, 3
N_AUDIO = int(round(34 / 15 * 16000))
FILL = {"gen": 120, "dis": 121}


def main():
    args = make_args("gesture")
    spk = MG.speaker_vocab(N_SPK)
    emb = make_embedding(N_WORDS, 300, 1).numpy()
    G = MG.no_dropout(det_fill(PoseGenerator(args, 27, N_WORDS, 300, emb.copy(), z_obj=spk), FILL["gen"]))
    D = MG.no_dropout(det_fill(ConvDiscriminator(27), FILL["dis"]))
    gold = {"n_words": N_WORDS, "n_spk": N_SPK, "B": B, "fill_seeds": FILL, "n_audio": N_AUDIO,
            "gen_keys": [(k, tuple(v.shape)) for k, v in G.state_dict().items()],
            "dis_keys": [(k, tuple(v.shape)) for k, v in D.state_dict().items()]}
    # ---- WavEncoder forward / backward (train-mode BatchNorm) on raw audio
    audio = torch.stack([make_audio(N_AUDIO, 40 + i) for i in range(B)])
    W = det_fill(WavEncoder(), 122).train(True)
    a = audio.clone().requires_grad_(True)
    y = W(a)
    gy = MG.randn(tuple(y.shape), 123, "gy")
    (y * gy).sum().backward()
    gold["wav"] = {"audio_seed": 40, "fill": 122, "gy_seed": 123, "y": y.detach().clone(), "dx": sample_tensor(a.grad),
                   "grads": MG.grads_summary(W), "buffers": {n: b.clone() for n, b in W.named_buffers()}}
    # ---- two training steps
    lr = args.learning_rate
    g_opt = torch.optim.Adam(G.parameters(), lr=lr, betas=(0.5, 0.999))
    d_opt = torch.optim.Adam(D.parameters(), lr=lr * args.discriminator_lr_weight, betas=(0.5, 0.999))
    steps = []
    for si, epoch in enumerate((0, 11)):
        batch = make_batch("gesture", B, N_WORDS, N_SPK, seed=50 + si)
        aud = torch.stack([make_audio(N_AUDIO, 60 + 10 * si + i) for i in range(B)])
        model.embedding_net.reparameterize = MG.EpsFeed(70 + si, B)
        perm = torch.randperm(B, generator=_gen(71 + si, "perm"))
        orig = torch.randperm
        torch.randperm = lambda n, *a_, **k_: perm.clone()
        try:
            ret = tg.train_iter_gan(args, epoch, batch["in_text_padded"], aud, batch["target"], batch["vid"], G, D, g_opt, d_opt)
        finally:
            torch.randperm = orig
        rec = {"epoch": epoch, "batch_seed": 50 + si, "audio_seed": 60 + 10 * si, "eps_seed": 70 + si, "perm": perm,
               "ret": {k: float(v) for k, v in ret.items()}, "grads": MG.grads_summary(G), "gen": MG.params_summary(G),
               "dis": MG.params_summary(D)}
        steps.append(rec)
        print("step", si, "epoch", epoch, rec["ret"])
    gold["steps"] = steps
    torch.save(gold, os.path.join(MG.OUT, "step_gan.pt"))
    print("step_gan.pt written")


if __name__ == "__main__":
    main()
