"""How many 8-CTA clusters of the GRU recurrence are resident at once, and what a second wave costs:
forward time over the number of row chunks at NB = 48.   python tools/gru_waves.py"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ha2g_b200._lib import lib
from ha2g_b200 import ops
from ha2g_b200.ops import _p, _st
dev = "cuda:0"
T, H = 34, 300
ops._ensure_workspace()
n = ctypes.c_int(0)
lib.ha2g_gru_max_clusters(ctypes.byref(n))
print("cudaOccupancyMaxActiveClusters (8 CTAs, NB = 48 configuration):", n.value)
w = [torch.randn(3 * H, H, device=dev) * 0.05 for _ in range(2)]
b = [torch.randn(3 * H, device=dev) * 0.05 for _ in range(2)]
for nb in (48, 16):
    for chunks in range(1, 11):
        M = chunks * nb
        gi = torch.randn(M, T, 6 * H, device=dev)
        y = torch.empty(M, T, 2 * H, device=dev)
        ts = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.ha2g_gru_seq_fwd_tc2_dbg(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), None, M, 0, T, H, nb, None, _st())
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"NB={nb} chunks/direction={chunks} ({2 * chunks} clusters, M={M}): {min(ts[2:]):.1f} us")
