"""smoke(): one tiny TED-Gesture training step (B=2, epoch 11 = full step with the discriminator) through the
drop-in train_iter_hierarchy on cuda:0, checked against the CPU oracle on the same inputs and draws."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_smoke(variant: str = "gesture", B: int = 2, verbose: bool = True):
    import ha2g_oracle as O
    from ha2g_b200 import constants as K
    from ha2g_b200 import rng
    from ha2g_b200.synthetic import make_batch
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
    from helpers import build_modules, randn, sd_cpu

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    dev = "cuda:0"
    seeds = {"gens": 20, "dis": 30, "audio": 31, "text": 32}
    n_words, n_spk, epoch = 60, 5, 11
    args, gens, D, A, T = build_modules(variant, n_words, n_spk, seeds, "cpu")
    state = {"gens": [sd_cpu(m) for m in gens], "dis": sd_cpu(D), "audio": sd_cpu(A), "text": sd_cpu(T)}
    for m in gens + [D, A, T]:
        m.to(dev)
    L = len(gens)
    batch = make_batch(variant, B, n_words, n_spk, seed=77)
    draws = [randn((B, 16), 78, f"eps{i}") for i in range(3 * L)]
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(5))
    lr = args.learning_rate
    mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
    fn = train_iter_hierarchy if variant == "gesture" else train_iter_hierarchy_expressive
    gb = {k: v.to(dev) for k, v in batch.items()}
    with rng.override(randn_fn=rng.ListFeed(draws), randperm_fn=lambda n: perm.clone(), dropout=False):
        ret = fn(args, epoch, gb["in_text_padded"], gb["in_spec"], gb["target"], gb["vid"], *gens, D, A, T,
                 *[mk(g) for g in gens], mk(D, lr * args.discriminator_lr_weight), mk(A), mk(T))
    torch.cuda.synchronize()
    tabs = ({"pairs": K.EXPRESSIVE_ANGLE_PAIR, "avg": K.EXPRESSIVE_AVG_ANGLE, "var": K.EXPRESSIVE_VAR_ANGLE} if variant == "expressive"
            else {"pairs": K.GESTURE_ANGLE_PAIR, "avg": K.GESTURE_AVG_ANGLE, "var": K.GESTURE_VAR_ANGLE})
    eps = {"d": draws[:L], "g": draws[L:2 * L], "r": draws[2 * L:]}
    ref, _, _ = O.train_step(variant, args, epoch, batch["in_text_padded"], batch["in_spec"], batch["target"], batch["vid"],
                             state["gens"], state["dis"], state["audio"], state["text"], {}, eps, perm, tabs)
    if verbose:
        print("smoke cuda  :", {k: round(v, 6) for k, v in ret.items()})
        print("smoke oracle:", {k: round(v, 6) for k, v in ref.items()})
    assert set(ret) == set(ref), (sorted(ret), sorted(ref))
    for k in ref:
        assert abs(ret[k] - ref[k]) <= 1e-3 * max(1.0, abs(ref[k])), (k, ret[k], ref[k])
    return ret, ref


if __name__ == "__main__":
    run_smoke(sys.argv[1] if len(sys.argv) > 1 else "gesture")
