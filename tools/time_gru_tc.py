"""Phase timing of the tcgen05 cluster GRU recurrence (clock64 samples from cluster 0 / rank 0), forward at every
rows-per-cluster choice and backward.   python tools/time_gru_tc.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ha2g_b200._lib import lib
from ha2g_b200 import ops
from ha2g_b200.ops import _p, _st
dev = "cuda:0"
T, H = 34, 300
ops._ensure_workspace()
w = [torch.randn(3 * H, H, device=dev) * 0.05 for _ in range(2)]
b = [torch.randn(3 * H, device=dev) * 0.05 for _ in range(2)]
dbg = torch.zeros((T + 1) * 8, dtype=torch.int64, device=dev)


def ref_gru(gi, M):
    """fp64 torch recurrence for the forward direction (dir 0) and the reverse one"""
    gi64 = gi.double().view(M, T, 2, 3 * H)
    ys = []
    for d in range(2):
        W, bb = w[d].double(), b[d].double()
        h = torch.zeros(M, H, dtype=torch.float64, device=dev)
        out = [None] * T
        for s in range(T):
            t = s if d == 0 else T - 1 - s
            gh = h @ W.t() + bb
            g = gi64[:, t, d]
            r = torch.sigmoid(g[:, :H] + gh[:, :H]); z = torch.sigmoid(g[:, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(g[:, 2 * H:] + r * gh[:, 2 * H:])
            h = (1 - z) * n + z * h
            out[t] = h
        ys.append(torch.stack(out, 1))
    return torch.cat(ys, 2)


for M, nb, mg in ((112, 16, 112), (128, 20, 128), (128, 0, 128), (224, 32, 0), (224, 0, 0), (336, 48, 128), (336, 0, 128), (384, 56, 128), (384, 0, 128), (1, 0, 0)):
    gi = torch.randn(M, T, 6 * H, device=dev)
    y = torch.empty(M, T, 2 * H, device=dev)
    gates = torch.full((max(mg, 1), T, 8 * H), float("nan"), device=dev)
    gp = _p(gates) if mg > 0 else None
    for _ in range(3):
        rc = lib.ha2g_gru_seq_fwd_tc2_dbg(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), gp, M, mg, T, H, nb, _p(dbg), _st())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.ha2g_gru_seq_fwd_tc2_dbg(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), gp, M, mg, T, H, nb, _p(dbg), _st())
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
    dall = dbg.view(T + 1, 8).cpu(); d = dall[:T]; pr = dall[T]
    err = float((y.double() - ref_gru(gi, M)).abs().max())
    flops = 2.0 * M * T * 2 * 3 * H * H
    print(f"fwd M={M} NB={nb} gates for {mg} rows: {us:.1f} us, {us / T:.2f} us/step, {flops / us * 1e-6:.1f} TFLOP/s fp32-equivalent; "
          f"max|y - fp64| {err:.2e}; gates finite: {bool(torch.isfinite(gates).all()) if mg else None}")
    names = ["mma issue+commit", "commit->epi wake", "tmem ld + transpose", "gate math", "pack + st.async", "y/gates stores"]
    for i, n in enumerate(names):
        print(f"   {n:22s}: {(d[5:T - 1, i + 1] - d[5:T - 1, i]).float().mean():8.0f} cycles")
    print(f"   step period           : {(d[6:, 0] - d[5:-1, 0]).float().mean():8.0f} cycles;  wait for h (mma thread): "
          f"{(d[5:, 0] - d[5:, 7]).float().mean():.0f};  push issued -> next mma start: {(d[6:, 0] - d[5:-1, 5]).float().mean():.0f}")
    print(f"   prologue: W->smem {int(pr[1] - pr[0])}, smem->TMEM {int(pr[2] - pr[1])}, to loop start {int(pr[3] - pr[2])}; loop {int(pr[4] - pr[3])} cycles")

# ---- what-if timings (results are wrong with a flag set): where the off-critical-path work of a step goes -----------------
for M, nb in ((384, 228),):
    gi = torch.randn(M, T, 6 * H, device=dev)
    y = torch.empty(M, T, 2 * H, device=dev)
    gates = torch.empty(128, T, 8 * H, device=dev)
    for flags, what in ((0, "as shipped"), (1, "no gi loads"), (2, "no global y/gate stores"), (4, "no staging, no stores"), (5, "no gi loads, no staging, no stores")):
        lib.ha2g_gru_fwd_xflags(flags)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.ha2g_gru_seq_fwd_tc2_dbg(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gates), M, 128, T, H, nb, _p(dbg), _st())
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        d = dbg.view(T + 1, 8).cpu()[:T]
        print(f"what-if M={M} NB={nb} {what:36s}: {min(ts[1:]):6.1f} us; step period {(d[6:, 0] - d[5:-1, 0]).float().mean():6.0f}, "
              f"push->next mma {(d[6:, 0] - d[5:-1, 5]).float().mean():6.0f}, stores phase {(d[5:T - 1, 6] - d[5:T - 1, 5]).float().mean():6.0f}")
    lib.ha2g_gru_fwd_xflags(0)

# ---- backward recurrence ----------------------------------------------------------------------------------------------
def ref_bwd(gi, dy, M):
    """fp64 autograd through the same recurrence: d(sum(y*dy))/d(gi)"""
    g = gi.double().clone().requires_grad_(True)
    gi64 = g.view(M, T, 2, 3 * H)
    ys = []
    for d in range(2):
        W, bb = w[d].double(), b[d].double()
        h = torch.zeros(M, H, dtype=torch.float64, device=dev)
        out = [None] * T
        for s in range(T):
            t = s if d == 0 else T - 1 - s
            gh = h @ W.t() + bb
            x = gi64[:, t, d]
            r = torch.sigmoid(x[:, :H] + gh[:, :H]); z = torch.sigmoid(x[:, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(x[:, 2 * H:] + r * gh[:, 2 * H:])
            h = (1 - z) * n + z * h
            out[t] = h
        ys.append(torch.stack(out, 1))
    (torch.cat(ys, 2) * dy.double()).sum().backward()
    return g.grad


for M, nb in ((128, 0), (112, 16), (256, 32)):
    gi = torch.randn(M, T, 6 * H, device=dev)
    y = torch.empty(M, T, 2 * H, device=dev); gates = torch.empty(M, T, 8 * H, device=dev)
    dy = torch.randn(M, T, 2 * H, device=dev) * 0.1
    lib.ha2g_gru_seq_fwd_tc2(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gates), M, M, T, H, _st())
    dgi = torch.empty(M, T, 6 * H, device=dev); dgh = torch.empty_like(dgi)
    for _ in range(3):
        lib.ha2g_gru_seq_bwd_tc2_dbg(_p(dy), 2 * H, H, _p(y), _p(gates), _p(w[0]), _p(w[1]), _p(dgi), _p(dgh), M, T, H, nb, _p(dbg), _st())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.ha2g_gru_seq_bwd_tc2_dbg(_p(dy), 2 * H, H, _p(y), _p(gates), _p(w[0]), _p(w[1]), _p(dgi), _p(dgh), M, T, H, nb, _p(dbg), _st())
    e1.record(); torch.cuda.synchronize()
    dall = dbg.view(T + 1, 8).cpu(); d = dall[:T]; pr = dall[T]
    ref = ref_bwd(gi, dy, M)
    err = float((dgi.double() - ref).abs().max() / ref.abs().max())
    print(f"bwd tc2 M={M} NB={nb}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us total, {e0.elapsed_time(e1) * 1e3 / T:.2f} us/step; max|dgi - fp64| / max|dgi| {err:.2e}")
    for i, nme in enumerate(["wait for partials", "reduce + gate grads + B operand", "MMA (72) + dgi/dgh copy-out", "TMEM->staging->bulk copies"]):
        print(f"   {nme:34s}: {(d[5:T - 1, i + 1] - d[5:T - 1, i]).float().mean():8.0f} cycles")
    print(f"   round period                      : {(d[6:T - 1, 0] - d[5:T - 2, 0]).float().mean():8.0f} cycles")
    print(f"   prologue: W->smem {int(pr[1] - pr[0])}, smem->TMEM {int(pr[2] - pr[1])}, to loop start {int(pr[3] - pr[2])}; loop {int(pr[4] - pr[3])} cycles")
