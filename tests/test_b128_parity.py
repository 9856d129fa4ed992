"""Parity at the BENCHMARK configuration: one full TED-Expressive / TED-Gesture training step at B = 128 clips (epoch 11:
discriminator step + generator step) on the CUDA path -- M = 384-row ride-along GRU tasks, 13 056-row GEMMs, N = 4352
streaming contrastive loss -- against the CPU oracle evaluated in fp64 on the same inputs and draws
(tests/golden/b128_<variant>.pt, written by oracle/make_b128_golden.py together with the fp32 oracle's own error against
fp64: the reference's noise floor).  Losses are held to the north-star 1e-3 (measured 1e-5); every pre-Adam gradient
tensor to max(2e-3, 4 x the reference's own fp32 error for that tensor's family) in relative L2 over a strided sample.
The generators' gradients are well conditioned (reference noise 1e-6): with exact-fp32 GEMMs the CUDA path lands at
1e-4..4e-4, with the production tensor-core GEMMs (bf16 hi + lo operands, 2^-18 operand representation, three MMAs per
product) at 2e-4..1.5e-3 -- hence 2e-3, stated here rather than hidden.  The text / audio encoders' gradients are
ill-conditioned even at B = 128 (reference noise 2e-3 / 6e-3..8e-3)."""
import os

import pytest
import torch

from helpers import build_modules, randn
from ha2g_b200 import rng
from ha2g_b200.synthetic import make_batch, sample_tensor

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("variant", ["expressive", "gesture"])
def test_step_b128_vs_fp64_oracle(variant):
    import ha2g_b200.train_eval._step as S
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
    g = torch.load(os.path.join(GOLD, f"b128_{variant}.pt"), weights_only=False)
    B = g["B"]
    args, gens, D, A, T = build_modules(variant, g["n_words"], g["n_spk"], g["fill_seeds"], DEV)
    L = len(gens)
    batch = {k: v.to(DEV) for k, v in make_batch(variant, B, g["n_words"], g["n_spk"], seed=g["batch_seed"]).items()}
    draws = [randn((B, 16), g["eps_seed"], f"eps{i}") for i in range(3 * L)]
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(g["perm_seed"]))
    lr = args.learning_rate
    mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
    fn = train_iter_hierarchy if variant == "gesture" else train_iter_hierarchy_expressive
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
            ret = fn(args, g["epoch"], batch["in_text_padded"], batch["in_spec"], batch["target"], batch["vid"], *gens, D, A, T,
                     *[mk(x) for x in gens], mk(D, lr * args.discriminator_lr_weight), mk(A), mk(T))
    finally:
        S.fused_adam_step = orig
    torch.cuda.synchronize()

    ref = g["ret64"]
    assert set(ret) == set(ref), (sorted(ret), sorted(ref))
    bad = {k: (ret[k], ref[k]) for k in ref if abs(ret[k] - ref[k]) > 1e-3 * max(1.0, abs(ref[k]))}
    assert not bad, f"B=128 loss dict mismatch (cuda, fp64 oracle): {bad}"

    report, worst = [], {}
    mods = {f"g{k + 1}": m for k, m in enumerate(gens)}
    mods.update(text=T, audio=A)
    for fam, entry in g["families"].items():
        named = dict(mods[fam].named_parameters())
        tol = max(2e-3, 4.0 * entry["fp32_ref_worst"])
        for name, rec in entry["tensors"].items():
            summ = rec["summary"]
            mine = sample_tensor(captured[id(named[name])], 512)
            assert mine["numel"] == summ["numel"], (fam, name)
            k = summ["sample"].numel() ** 0.5
            ref_s = summ["sample"].double()
            denom = max(float(ref_s.norm()), 1e-2 * entry["scale_rms"] * k, 1e-30)
            e = float((mine["sample"].double() - ref_s).norm()) / denom
            worst[fam] = max(worst.get(fam, 0.0), e)
            if e > tol:
                report.append(f"{fam}.{name}: rel L2 {e:.3e} > {tol:.1e} (reference fp32's own error {rec['fp32_ref_err']:.1e})")
    print(f"B=128 {variant}: worst relative L2 gradient error vs fp64 per family (cuda | reference fp32):",
          {k: f"{v:.1e} | {g['families'][k]['fp32_ref_worst']:.1e}" for k, v in worst.items()})
    assert not report, "\n".join(report[:12]) + f"\n({len(report)} gradient tensors out of tolerance)"
