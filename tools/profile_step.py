"""Two full training steps (1 warm + 1 to capture) of the bench workload, for ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py [--batch 128] [--variant expressive] [--steps 2]
A number printed by a run under ncu is never a bench value; this script prints none."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--variant", default="expressive")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--epoch", type=int, default=11)
    a = ap.parse_args()
    import torch
    import bench
    from ha2g_b200 import graph_step, ops
    graph_step.enable(False)   # the same launches, issued one by one: ncu then lists every kernel under its own name
    from ha2g_b200.synthetic import make_batch
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
    dev = torch.device("cuda", 0)
    args, gens, D, A, T, (gopts, dopt, aopt, topt) = bench.build_world(a.variant, dev)
    fn = train_iter_hierarchy if a.variant == "gesture" else train_iter_hierarchy_expressive
    b = {k: v.to(dev) for k, v in make_batch(a.variant, a.batch, bench.N_WORDS, bench.N_SPEAKERS, seed=1).items()}
    for i in range(a.steps):
        n0 = ops.LAUNCHES[0]
        fn(args, a.epoch, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *gens, D, A, T, *gopts, dopt, aopt, topt)
        torch.cuda.synchronize()
        print(f"step {i}: {ops.LAUNCHES[0] - n0} C-ABI calls", file=sys.stderr)


if __name__ == "__main__":
    main()
