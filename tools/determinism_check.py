"""Run the same eager step sequence twice from identical state and report run-to-run differences
(losses per step, per-tensor gradient differences after step 0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import build_modules
from ha2g_b200 import graph_step, rng
from ha2g_b200.synthetic import make_batch

SEEDS = {"gens": 20, "dis": 30, "audio": 31, "text": 32}


def run(variant, n_steps, B, graph=False):
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
    dev = "cuda:0"
    fn = train_iter_hierarchy if variant == "gesture" else train_iter_hierarchy_expressive
    args, gens, D, A, T = build_modules(variant, 60, 5, SEEDS, dev)
    lr = args.learning_rate
    mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
    opts = [mk(g) for g in gens] + [mk(D, lr * args.discriminator_lr_weight), mk(A), mk(T)]
    g = torch.Generator().manual_seed(3)
    # 3 cascade passes x L generators draw reparameterisation noise each step: a fixed cycle of distinct draws (the
    # same in every step and in both runs; identical draws in the G and mismatched-speaker passes would make the
    # diversity loss 0/0)
    noise = [torch.randn((B, 16), generator=g).to(dev) for _ in range(3 * len(gens))]
    calls = [0]

    def randn_fn(shape):
        calls[0] += 1
        return noise[(calls[0] - 1) % len(noise)]
    perm = torch.randperm(B, generator=g).to(dev)
    graph_step.reset(); graph_step.enable(graph)
    rets, grads = [], []
    names = [f"m{mi}.{n}" for mi, m in enumerate(gens + [D, A, T]) for n, p in m.named_parameters()]
    with rng.override(randn_fn=randn_fn, randperm_fn=lambda n: perm, dropout=False, graph_safe=True):
        for i in range(n_steps):
            b = {k: v.to(dev) for k, v in make_batch(variant, B, 60, 5, seed=500 + i).items()}
            rets.append(fn(args, 11, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *gens, D, A, T, *opts))
            grads.append([None if p.grad is None else p.grad.detach().clone() for m in gens + [D, A, T] for p in m.parameters()])
    torch.cuda.synchronize()
    return rets, grads, names


variant = sys.argv[1] if len(sys.argv) > 1 else "gesture"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
a_r, a_g, names = run(variant, 3, B)
b_r, b_g, _ = run(variant, 3, B)
for i in range(3):
    print("step", i, {k: (round(a_r[i][k], 5), round(b_r[i][k], 5)) for k in a_r[i]})
for i in range(2):
    rows = []
    for n, x, y in zip(names, a_g[i], b_g[i]):
        if x is None:
            continue
        d = float((x - y).abs().max()); s = float(x.abs().max())
        rows.append((d / max(s, 1e-30), d, s, n))
    rows.sort(reverse=True)
    nz = sum(1 for r in rows if r[1] > 0)
    print(f"step {i}: {nz}/{len(rows)} gradient tensors differ between the two runs; worst:")
    for r in rows[:12]:
        print(f"   rel {r[0]:.3e} abs {r[1]:.3e} scale {r[2]:.3e}  {r[3]}")
