"""Phase timing of the tcgen05 cluster GRU recurrence (clock64 samples from cluster 0 / rank 0)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ha2g_b200._lib import lib
from ha2g_b200 import ops
from ha2g_b200.ops import _p, _st
dev = "cuda:0"
M, T, H = 128, 34, 300
gi = torch.randn(M, T, 6 * H, device=dev)
w = [torch.randn(3 * H, H, device=dev) * 0.05 for _ in range(2)]
b = [torch.randn(3 * H, device=dev) * 0.05 for _ in range(2)]
y = torch.empty(M, T, 2 * H, device=dev)
gates = torch.empty(M, T, 8 * H, device=dev)
dbg = torch.zeros(T * 8, dtype=torch.int64, device=dev)
ops._ensure_workspace()
for gts in (gates, None):
    for _ in range(3):
        lib.ha2g_gru_seq_fwd_tc_dbg(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gts), M, T, H, _p(dbg), _st())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.ha2g_gru_seq_fwd_tc_dbg(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gts), M, T, H, _p(dbg), _st())
    e1.record(); torch.cuda.synchronize()
    d = dbg.view(T, 8).cpu()
    names = ["mma issue", "commit->epi wake", "tmem ld + transpose", "gate math + stores", "dsmem push", "cluster.sync"]
    print(f"gates={'saved' if gts is not None else 'none'}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us total, {e0.elapsed_time(e1) * 1e3 / T:.2f} us/step")
    for i, n in enumerate(names):
        seg = (d[5:, i + 1] - d[5:, i]).float()
        print(f"   {n:22s}: {seg.mean():8.0f} cycles")
    print(f"   step period           : {(d[6:, 0] - d[5:-1, 0]).float().mean():8.0f} cycles")
for impl in ("cluster",):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        lib.ha2g_gru_seq_fwd_cluster(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gates), M, T, H, _st())
    e0.record()
    lib.ha2g_gru_seq_fwd_cluster(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gates), M, T, H, _st())
    e1.record(); torch.cuda.synchronize()
    print(f"fp32 cluster kernel: {e0.elapsed_time(e1) * 1e3:.1f} us total, {e0.elapsed_time(e1) * 1e3 / T:.2f} us/step")
