"""Is the step host-bound?  Times (a) the host enqueue of one step without any sync inside (the final packed
.tolist() read is the only sync) and (b) the device span, and prints per-section host enqueue times."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from ha2g_b200 import graph_step, ops
    graph_step.enable(False)
    from ha2g_b200.synthetic import make_batch
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive as fn
    dev = torch.device("cuda", 0)
    args, gens, D, A, T, (gopts, dopt, aopt, topt) = bench.build_world("expressive", dev)
    b = {k: v.to(dev) for k, v in make_batch("expressive", 128, bench.N_WORDS, bench.N_SPEAKERS, seed=1).items()}
    call = lambda: fn(args, 11, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *gens, D, A, T, *gopts, dopt, aopt, topt)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    # patch the packed read so that we can see when the host finished enqueueing
    marks = {}
    orig_cat = torch.cat

    def cat_mark(*a, **k):
        marks["enqueued"] = time.perf_counter()
        return orig_cat(*a, **k)
    for i in range(3):
        torch.cuda.synchronize()
        torch.cat = cat_mark
        t0 = time.perf_counter()
        call()
        t1 = time.perf_counter()
        torch.cat = orig_cat
        print(f"step {i}: host enqueue {1e3 * (marks['enqueued'] - t0):.1f} ms, total (incl. final read) {1e3 * (t1 - t0):.1f} ms")
    # audio encoder alone: host enqueue vs device
    for name, f in (("audio fwd", lambda: A(b["in_spec"], b["vid"])), ("text fwd", lambda: T(b["in_text_padded"]))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            f()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{name}: host {1e3 * (t1 - t0):.2f} ms, device done after {1e3 * (t2 - t0):.2f} ms")


if __name__ == "__main__":
    main()
