"""Where do the aten::fill_/zero_ kernels of one eager training step come from?  torch.profiler with Python stacks."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from ha2g_b200 import graph_step
from ha2g_b200.synthetic import make_batch
from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive as fn
graph_step.enable(False)
dev = torch.device("cuda", 0)
args, gens, D, A, T, (gopts, dopt, aopt, topt) = bench.build_world("expressive", dev)
b = {k: v.to(dev) for k, v in make_batch("expressive", 128, bench.N_WORDS, bench.N_SPEAKERS, seed=1).items()}
call = lambda: fn(args, 11, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *gens, D, A, T, *gopts, dopt, aopt, topt)
call(); call()
with profile(activities=[ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
    call()
torch.cuda.synchronize()
hist = collections.Counter()
names = collections.Counter()
for ev in prof.events():
    if ev.name in ("aten::fill_", "aten::zero_"):
        names[ev.name] += 1
        st = [s for s in ev.stack if "ha2g_b200" in s or "torch/autograd" in s or "_step" in s][:2]
        hist[(ev.name, tuple(st) if st else tuple(ev.stack[:2]), str(ev.input_shapes)[:40])] += 1
print(dict(names))
for k, c in hist.most_common(30):
    print(c, k)
