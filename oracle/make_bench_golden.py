"""TEST INFRASTRUCTURE (not product code).  Golden loss dict of ONE full training step at the benchmark configuration
(B = 128 clips, epoch 11, n_words = 30 000, n_speakers = 1 500, modules initialised exactly as bench.py::build_world
does with seed 0), computed by the CPU oracle (oracle/ha2g_oracle.py, itself pinned to the unmodified reference by
tests/test_oracle_pinned.py) with dropout off and injected reparameterisation noise / speaker permutation.

bench.py re-runs the same step through the CUDA path before its warm-up and refuses to print a number when the losses
differ by more than 1e-3 (north-star tolerance).  Writes tests/golden/bench_step_losses.json.

    python oracle/make_bench_golden.py [expressive gesture]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import bench
import ha2g_oracle as O
from ha2g_b200 import constants as K
from ha2g_b200.synthetic import _gen, make_batch

OUT = os.path.join(ROOT, "tests", "golden", "bench_step_losses.json")
B, EPOCH, BATCH_SEED, EPS_SEED = 128, 11, 4242, 4243


def draws(L):
    return [torch.randn((B, 16), generator=_gen(EPS_SEED, f"eps{i}")) for i in range(3 * L)]


def perm():
    return torch.randperm(B, generator=_gen(EPS_SEED, "perm"))


def main():
    variants = sys.argv[1:] or ["expressive", "gesture"]
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    torch.set_num_threads(os.cpu_count() or 1)
    for variant in variants:
        args, gens, D, A, T, _ = bench.build_world(variant, "cpu", seed=0)
        sd = lambda m: {k: v.detach().clone() for k, v in m.state_dict().items()}
        L = len(gens)
        batch = make_batch(variant, B, bench.N_WORDS, bench.N_SPEAKERS, seed=BATCH_SEED)
        d = draws(L)
        eps = {"d": d[:L], "g": d[L:2 * L], "r": d[2 * L:]}
        tabs = ({"pairs": K.EXPRESSIVE_ANGLE_PAIR, "avg": K.EXPRESSIVE_AVG_ANGLE, "var": K.EXPRESSIVE_VAR_ANGLE}
                if variant == "expressive" else
                {"pairs": K.GESTURE_ANGLE_PAIR, "avg": K.GESTURE_AVG_ANGLE, "var": K.GESTURE_VAR_ANGLE})
        t0 = time.time()
        ret, _, _ = O.train_step(variant, args, EPOCH, batch["in_text_padded"], batch["in_spec"], batch["target"], batch["vid"],
                                 [sd(g) for g in gens], sd(D), sd(A), sd(T), {}, eps, perm(), tabs)
        out[variant] = {"B": B, "epoch": EPOCH, "batch_seed": BATCH_SEED, "eps_seed": EPS_SEED, "n_words": bench.N_WORDS,
                        "n_speakers": bench.N_SPEAKERS, "init_seed": 0, "torch": torch.__version__,
                        "oracle_seconds": round(time.time() - t0, 1), "ret": {k: float(v) for k, v in ret.items()}}
        print(variant, out[variant])
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
