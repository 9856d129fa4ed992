"""Parity at the BENCHMARK configuration: one full TED-Expressive / TED-Gesture training step at B = 128 clips (epoch 11:
discriminator step + generator step), CUDA path vs the CPU oracle on the same inputs and draws.  Exercises what the
B = 2..5 goldens cannot: M = 128-row GRU tasks, the batched no-grad cascades, 4352-row GEMM tiles, N = 4352 streaming
contrastive loss.  The oracle runs on the host cores of the GPU box (tens of seconds), vocabulary kept small."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from helpers import build_modules, randn, sd_cpu
from ha2g_b200 import constants as K
from ha2g_b200 import rng
from ha2g_b200.synthetic import make_batch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SEEDS = {"gens": 20, "dis": 30, "audio": 31, "text": 32}


def _rel_l2(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm()) / max(float(b.norm()), 1e-30)


@pytest.mark.parametrize("variant", ["expressive", "gesture"])
def test_step_b128_vs_oracle(variant):
    import ha2g_oracle as O
    import ha2g_b200.train_eval._step as S
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
    B, n_words, n_spk, epoch = 128, 60, 5, 11
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    args, gens, D, A, T = build_modules(variant, n_words, n_spk, SEEDS, "cpu")
    state = {"gens": [sd_cpu(m) for m in gens], "dis": sd_cpu(D), "audio": sd_cpu(A), "text": sd_cpu(T)}
    for m in gens + [D, A, T]:
        m.to(DEV)
    L = len(gens)
    batch = make_batch(variant, B, n_words, n_spk, seed=901)
    draws = [randn((B, 16), 902, f"eps{i}") for i in range(3 * L)]
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(9))
    lr = args.learning_rate
    mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
    fn = train_iter_hierarchy if variant == "gesture" else train_iter_hierarchy_expressive
    gb = {k: v.to(DEV) for k, v in batch.items()}
    captured = {}
    orig = S.fused_adam_step

    def spy(opt):
        for grp in opt.param_groups:
            for p in grp["params"]:
                if p.grad is not None:
                    captured[id(p)] = p.grad.detach().clone()
        orig(opt)
    S.fused_adam_step = spy
    try:
        with rng.override(randn_fn=rng.ListFeed(draws), randperm_fn=lambda n: perm.clone(), dropout=False):
            ret = fn(args, epoch, gb["in_text_padded"], gb["in_spec"], gb["target"], gb["vid"], *gens, D, A, T,
                     *[mk(g) for g in gens], mk(D, lr * args.discriminator_lr_weight), mk(A), mk(T))
    finally:
        S.fused_adam_step = orig
    torch.cuda.synchronize()

    tabs = ({"pairs": K.EXPRESSIVE_ANGLE_PAIR, "avg": K.EXPRESSIVE_AVG_ANGLE, "var": K.EXPRESSIVE_VAR_ANGLE} if variant == "expressive"
            else {"pairs": K.GESTURE_ANGLE_PAIR, "avg": K.GESTURE_AVG_ANGLE, "var": K.GESTURE_VAR_ANGLE})
    eps = {"d": draws[:L], "g": draws[L:2 * L], "r": draws[2 * L:]}
    ref, _, ref_grads = O.train_step(variant, args, epoch, batch["in_text_padded"], batch["in_spec"], batch["target"],
                                     batch["vid"], state["gens"], state["dis"], state["audio"], state["text"], {}, eps, perm, tabs)
    assert set(ret) == set(ref), (sorted(ret), sorted(ref))
    bad = {k: (ret[k], ref[k]) for k in ref if abs(ret[k] - ref[k]) > 1e-3 * max(1.0, abs(ref[k]))}
    assert not bad, f"B=128 loss dict mismatch (cuda, oracle): {bad}"

    # pre-Adam gradients, relative L2 per tensor (tensors whose gradient is fp32 noise -- far below the family's scale --
    # are compared against the family scale instead of their own norm)
    report, worst = [], {}
    fams = [(f"g{k + 1}", m, ref_grads["gens"][k], 2e-3) for k, m in enumerate(gens)]
    fams += [("text", T, ref_grads["text"], 2e-3), ("audio", A, ref_grads["audio"], 1e-2)]
    for fam, mod, rg, tol in fams:
        named = dict(mod.named_parameters())
        scale = max(float(g.double().norm()) / g.numel() ** 0.5 for g in rg.values())
        for name, g_ref in rg.items():
            g_cuda = captured[id(named[name])]
            floor = 1e-2 * scale * g_ref.numel() ** 0.5
            e = float((g_cuda.double().cpu() - g_ref.double()).norm()) / max(float(g_ref.double().norm()), floor)
            worst[fam] = max(worst.get(fam, 0.0), e)
            if e > tol:
                report.append(f"{fam}.{name}: rel L2 {e:.3e} > {tol:.0e}")
    print(f"B=128 {variant} worst relative L2 gradient error per family:", {k: f"{v:.2e}" for k, v in worst.items()})
    assert not report, "\n".join(report[:12]) + f"\n({len(report)} gradient tensors out of tolerance)"
