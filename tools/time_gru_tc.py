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
dbg = torch.zeros((T + 1) * 8, dtype=torch.int64, device=dev)
ops._ensure_workspace()
for tc2, gts in ((True, gates), (True, None), (False, gates)):
    fn = lib.ha2g_gru_seq_fwd_tc2_dbg if tc2 else lib.ha2g_gru_seq_fwd_tc_dbg
    for _ in range(3):
        fn(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gts), M, T, H, _p(dbg), _st())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gts), M, T, H, _p(dbg), _st())
    e1.record(); torch.cuda.synchronize()
    dall = dbg.view(T + 1, 8).cpu()
    d = dall[:T]
    names = ["mma issue", "commit->epi wake", "tmem ld + transpose", "gate math + stores", "dsmem push", "cluster.sync"]
    if tc2:
        names = ["mma issue+commit", "commit->epi wake", "tmem ld + transpose", "gate math", "pack + bulk copies", "y/gates stores"]
    print(f"{'tc2' if tc2 else 'tc1'} gates={'saved' if gts is not None else 'none'}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us total, {e0.elapsed_time(e1) * 1e3 / T:.2f} us/step")
    for i, n in enumerate(names):
        seg = (d[5:T - 1, i + 1] - d[5:T - 1, i]).float()
        print(f"   {n:22s}: {seg.mean():8.0f} cycles")
    print(f"   step period           : {(d[6:, 0] - d[5:-1, 0]).float().mean():8.0f} cycles")
    if tc2:
        print(f"   wait for h (mma thread): {(d[5:, 0] - d[5:, 7]).float().mean():8.0f} cycles;  copies issued -> next mma start: {(d[6:, 0] - d[5:-1, 5]).float().mean():8.0f} cycles")
        yref = y.clone()
        pr = dall[T]
        print(f"   prologue: W->smem {int(pr[1] - pr[0])}, smem->TMEM {int(pr[2] - pr[1])}, to loop start {int(pr[3] - pr[2])}; loop {int(pr[4] - pr[3])} cycles; first steps: {[int(x) for x in (d[1:6, 0] - d[0:5, 0])]}")
for impl in ("cluster",):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        lib.ha2g_gru_seq_fwd_cluster(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gates), M, T, H, _st())
    e0.record()
    lib.ha2g_gru_seq_fwd_cluster(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gates), M, T, H, _st())
    e1.record(); torch.cuda.synchronize()
    print(f"max |y_tc2 - y_fp32cluster| = {float((yref - y).abs().max()):.3e} (|y| max {float(y.abs().max()):.3f})")
    print(f"fp32 cluster kernel: {e0.elapsed_time(e1) * 1e3:.1f} us total, {e0.elapsed_time(e1) * 1e3 / T:.2f} us/step")

# ---- backward recurrence: tcgen05 kernel vs the fp32 cluster kernel ------------------------------------------------
dy = torch.randn(M, T, 2 * H, device=dev) * 0.1
lib.ha2g_gru_seq_fwd_tc2(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gates), M, T, H, _st())
dgi = torch.empty(M, T, 6 * H, device=dev); dgh = torch.empty_like(dgi)
dgi2 = torch.empty_like(dgi); dgh2 = torch.empty_like(dgi)
for _ in range(3):
    lib.ha2g_gru_seq_bwd_tc2_dbg(_p(dy), 2 * H, H, _p(y), _p(gates), _p(w[0]), _p(w[1]), _p(dgi), _p(dgh), M, T, H, _p(dbg), _st())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lib.ha2g_gru_seq_bwd_tc2_dbg(_p(dy), 2 * H, H, _p(y), _p(gates), _p(w[0]), _p(w[1]), _p(dgi), _p(dgh), M, T, H, _p(dbg), _st())
e1.record(); torch.cuda.synchronize()
dall = dbg.view(T + 1, 8).cpu(); d = dall[:T]; pr = dall[T]
print(f"bwd tc2: kernel {e0.elapsed_time(e1) * 1e3:.1f} us total, {e0.elapsed_time(e1) * 1e3 / T:.2f} us/step")
for i, nme in enumerate(["wait for partials", "reduce + gate grads + B operand", "MMA (72) + dgi/dgh copy-out", "TMEM->staging->bulk copies"]):
    print(f"   {nme:34s}: {(d[5:T - 1, i + 1] - d[5:T - 1, i]).float().mean():8.0f} cycles")
print(f"   round period                      : {(d[6:T - 1, 0] - d[5:T - 2, 0]).float().mean():8.0f} cycles")
print(f"   prologue: W->smem {int(pr[1] - pr[0])}, smem->TMEM {int(pr[2] - pr[1])}, to loop start {int(pr[3] - pr[2])}; loop {int(pr[4] - pr[3])} cycles")
for _ in range(2):
    lib.ha2g_gru_seq_bwd_cluster(_p(dy), 2 * H, H, _p(y), _p(gates), _p(w[0]), _p(w[1]), _p(dgi2), _p(dgh2), M, T, H, _st())
e0.record()
lib.ha2g_gru_seq_bwd_cluster(_p(dy), 2 * H, H, _p(y), _p(gates), _p(w[0]), _p(w[1]), _p(dgi2), _p(dgh2), M, T, H, _st())
e1.record(); torch.cuda.synchronize()
print(f"bwd fp32 cluster kernel: {e0.elapsed_time(e1) * 1e3:.1f} us total;  max|dgi diff| {float((dgi - dgi2).abs().max()):.3e} (max {float(dgi2.abs().max()):.3e}), "
      f"max|dgh diff| {float((dgh - dgh2).abs().max()):.3e}")
