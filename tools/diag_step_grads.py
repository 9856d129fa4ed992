"""Diagnostic: first golden step (tests/golden/step_<variant>.pt), per-tensor gradient errors of one module family against
the reference fixture, under different execution switches.   python tools/diag_step_grads.py [variant] [family]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_modules, randn, summary_scale
from ha2g_b200 import ops, rng
from ha2g_b200.synthetic import make_batch, sample_tensor
import ha2g_b200.train_eval._step as S

variant = sys.argv[1] if len(sys.argv) > 1 else "expressive"
fam = sys.argv[2] if len(sys.argv) > 2 else "g_first"
DEV = "cuda:0"
g = torch.load(os.path.join(ROOT, "tests", "golden", f"step_{variant}.pt"), weights_only=False)
rec = g["steps"][0]


def run(tag, batch_passes=True, gemm="auto"):
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
    S._BATCH_PASSES = batch_passes
    ops.set_gemm_impl(gemm)
    args, gens, D, A, T = build_modules(variant, g["n_words"], g["n_spk"], g["fill_seeds"], DEV)
    lr = args.learning_rate
    mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
    fn = train_iter_hierarchy if variant == "gesture" else train_iter_hierarchy_expressive
    L = len(gens)
    batch = {k: v.to(DEV) for k, v in make_batch(variant, g["B"], g["n_words"], g["n_spk"], seed=rec["batch_seed"]).items()}
    n_draws = (3 if rec["epoch"] > args.loss_warmup else 2) * L
    feed = rng.ListFeed([randn((g["B"], 16), rec["eps_seed"], f"eps{i}") for i in range(n_draws)])
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
        with rng.override(randn_fn=feed, randperm_fn=lambda n, p=rec["perm"]: p.clone(), dropout=False):
            ret = fn(args, rec["epoch"], batch["in_text_padded"], batch["in_spec"], batch["target"], batch["vid"], *gens, D, A, T,
                     *[mk(x) for x in gens], mk(D, lr * args.discriminator_lr_weight), mk(A), mk(T))
    finally:
        S.fused_adam_step = orig
    mod = {"g_last": gens[-1], "g_first": gens[0], "audio": A, "text": T}[fam]
    floor = summary_scale(rec["grads"][fam].values())
    named = dict(mod.named_parameters())
    rows = []
    for name, summ in rec["grads"][fam].items():
        mine = sample_tensor(captured[id(named[name])])
        ref_s = summ["sample"].double()
        k = ref_s.numel() ** 0.5
        denom = max(float(ref_s.norm()), summ["norm"] / max(summ["numel"], 1) ** 0.5 * k, floor * k, 1e-12)
        rows.append((float((mine["sample"].double() - ref_s).norm()) / denom, name))
    rows.sort(reverse=True)
    print(f"== {tag}: losses {ret}")
    for e, n in rows[:6]:
        print(f"   {e:.3e}  {n}")


run("ride-along, auto gemm")
run("separate passes, auto gemm", batch_passes=False)
run("ride-along, f32 gemm", gemm="f32")
run("separate passes, f32 gemm", batch_passes=False, gemm="f32")
