"""TEST INFRASTRUCTURE (not product code).  fp64 ground truth for the parity test at the benchmark batch size.

Runs the CPU oracle's train_step (oracle/ha2g_oracle.py, pinned to the unmodified reference) at B = 128, epoch 11, twice:
in fp32 (what the reference computes) and in fp64 (what it approximates).  Stores, per parameter tensor, a strided
sample + norm of the fp64 gradient and the fp32 run's own relative-L2 error against fp64 -- the reference's noise floor:
the text / audio encoder gradients are ill-conditioned even at B = 128 (2e-3 / 6e-3), the generators' are not (1e-6).
tests/test_b128_parity.py holds the CUDA step to max(1e-3, 4 x that floor) against the fp64 values.

    python oracle/make_b128_golden.py [expressive gesture]      (~2.5 min per variant on 8 cores)
Writes tests/golden/b128_<variant>.pt
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import ha2g_oracle as O
from ha2g_b200 import constants as K
from ha2g_b200.synthetic import make_batch, sample_tensor
from helpers import build_modules, randn, sd_cpu

B, N_WORDS, N_SPK, EPOCH = 128, 60, 5, 11
SEEDS = {"gens": 20, "dis": 30, "audio": 31, "text": 32}
BATCH_SEED, EPS_SEED, PERM_SEED = 901, 902, 9


def inputs(variant):
    L = 3 if variant == "gesture" else 6
    batch = make_batch(variant, B, N_WORDS, N_SPK, seed=BATCH_SEED)
    draws = [randn((B, 16), EPS_SEED, f"eps{i}") for i in range(3 * L)]
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(PERM_SEED))
    return batch, draws, perm


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for variant in (sys.argv[1:] or ["expressive", "gesture"]):
        args, gens, D, A, T = build_modules(variant, N_WORDS, N_SPK, SEEDS, "cpu")
        state = {"gens": [sd_cpu(m) for m in gens], "dis": sd_cpu(D), "audio": sd_cpu(A), "text": sd_cpu(T)}
        L = len(gens)
        batch, draws, perm = inputs(variant)
        tabs = ({"pairs": K.EXPRESSIVE_ANGLE_PAIR, "avg": K.EXPRESSIVE_AVG_ANGLE, "var": K.EXPRESSIVE_VAR_ANGLE}
                if variant == "expressive" else
                {"pairs": K.GESTURE_ANGLE_PAIR, "avg": K.GESTURE_AVG_ANGLE, "var": K.GESTURE_VAR_ANGLE})
        res = {}
        for dt in (torch.float32, torch.float64):
            torch.set_default_dtype(dt)
            cv = lambda sd: {k: (v.to(dt) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in sd.items()}
            eps = {"d": [d.to(dt) for d in draws[:L]], "g": [d.to(dt) for d in draws[L:2 * L]], "r": [d.to(dt) for d in draws[2 * L:]]}
            t0 = time.time()
            ret, _, grads = O.train_step(variant, args, EPOCH, batch["in_text_padded"], batch["in_spec"].to(dt),
                                         batch["target"].to(dt), batch["vid"], [cv(g) for g in state["gens"]], cv(state["dis"]),
                                         cv(state["audio"]), cv(state["text"]), {}, eps, perm, tabs)
            print(variant, dt, f"{time.time() - t0:.1f} s", ret, flush=True)
            res[dt] = (ret, grads)
        torch.set_default_dtype(torch.float32)
        g32, g64 = res[torch.float32][1], res[torch.float64][1]
        fams = [(f"g{k + 1}", g32["gens"][k], g64["gens"][k]) for k in range(L)] + \
               [("text", g32["text"], g64["text"]), ("audio", g32["audio"], g64["audio"])]
        out = {"variant": variant, "B": B, "n_words": N_WORDS, "n_spk": N_SPK, "epoch": EPOCH, "fill_seeds": SEEDS,
               "batch_seed": BATCH_SEED, "eps_seed": EPS_SEED, "perm_seed": PERM_SEED, "torch": torch.__version__,
               "ret64": {k: float(v) for k, v in res[torch.float64][0].items()},
               "ret32": {k: float(v) for k, v in res[torch.float32][0].items()}, "families": {}}
        for fam, a, b in fams:
            scale = max(float(v.double().norm()) / v.numel() ** 0.5 for v in b.values())
            entry = {"scale_rms": scale, "tensors": {}}
            for n in b:
                floor = 1e-2 * scale * b[n].numel() ** 0.5
                e = float((a[n].double() - b[n].double()).norm()) / max(float(b[n].double().norm()), floor)
                entry["tensors"][n] = {"fp32_ref_err": e, "summary": sample_tensor(b[n], 512)}
            entry["fp32_ref_worst"] = max(t["fp32_ref_err"] for t in entry["tensors"].values())
            out["families"][fam] = entry
            print(f"  {fam}: reference fp32 vs fp64 worst relative L2 {entry['fp32_ref_worst']:.3e}")
        torch.save(out, os.path.join(ROOT, "tests", "golden", f"b128_{variant}.pt"))


if __name__ == "__main__":
    main()
