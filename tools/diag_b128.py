"""Diagnostic: B=128 step gradients vs the fp64 golden under execution switches; prints the worst tensors per family.
    python tools/diag_b128.py [variant]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_modules, randn
from ha2g_b200 import ops, rng
from ha2g_b200.synthetic import make_batch, sample_tensor
import ha2g_b200.train_eval._step as S

variant = sys.argv[1] if len(sys.argv) > 1 else "expressive"
DEV = "cuda:0"
g = torch.load(os.path.join(ROOT, "tests", "golden", f"b128_{variant}.pt"), weights_only=False)


def run(tag, batch_passes=True, gemm="auto"):
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
    S._BATCH_PASSES = batch_passes
    ops.set_gemm_impl(gemm)
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
    mods = {f"g{k + 1}": m for k, m in enumerate(gens)}
    mods.update(text=T, audio=A)
    print(f"== {tag}")
    for fam, entry in g["families"].items():
        named = dict(mods[fam].named_parameters())
        rows = []
        for name, rec in entry["tensors"].items():
            summ = rec["summary"]
            mine = sample_tensor(captured[id(named[name])], 512)
            k = summ["sample"].numel() ** 0.5
            ref_s = summ["sample"].double()
            denom = max(float(ref_s.norm()), 1e-2 * entry["scale_rms"] * k, 1e-30)
            rows.append((float((mine["sample"].double() - ref_s).norm()) / denom, name))
        rows.sort(reverse=True)
        print(f"   {fam}: " + "; ".join(f"{e:.1e} {n}" for e, n in rows[:3]))


run("ride-along, auto gemm")
run("separate passes, auto gemm", batch_passes=False)
run("ride-along, f32 gemm", gemm="f32")
