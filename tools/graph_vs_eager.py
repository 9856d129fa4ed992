"""Diagnostic: one synchronised step, CUDA-graph replay vs eager, per-tensor gradient differences (which tensors are not
bit-identical, in module order) and the loss dicts at full precision.   python tools/graph_vs_eager.py [variant] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from ha2g_b200 import graph_step
import test_graph_step as TG

variant = sys.argv[1] if len(sys.argv) > 1 else "gesture"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
graph_step.reset(); graph_step.enable(True)
wg = TG.World(variant, B)
for i in range(4):
    wg.step(i)
graph_step.enable(False)
we = TG.World(variant, B); we.load_from(wg)
we2 = TG.World(variant, B); we2.load_from(wg)
r_e = we.step(4); g_e = TG._grads(we)
r_e2 = we2.step(4); g_e2 = TG._grads(we2)
graph_step.enable(True)
r_g = wg.step(4); g_g = TG._grads(wg)
print("eager :", r_e); print("eager2:", r_e2); print("graph :", r_g)
names = [f"m{mi}.{k}" for mi, m in enumerate(we.mods) for k, _ in m.named_parameters()]
for tag, ga, gb in (("eager vs eager(second world)", g_e, g_e2), ("eager vs graph", g_e, g_g)):
    bad = []
    for n, a, b in zip(names, ga, gb):
        if a is None or b is None:
            if (a is None) != (b is None): bad.append((n, "None mismatch", 0, 0))
            continue
        if not torch.equal(a, b):
            d = float((a - b).abs().max()); s = float(a.abs().max())
            bad.append((n, d / max(s, 1e-30), d, s))
    print(f"== {tag}: {len(bad)} / {len(names)} tensors differ")
    for r in bad[:40]:
        print("   ", r)
