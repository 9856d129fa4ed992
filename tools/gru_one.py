"""A few launches of the fused GRU forward recurrence at the step's shape (M = 384 ride-along rows, gates saved for 128),
for `ncu --set full -k regex:gru_seq_fwd_tc2`:   python tools/gru_one.py [M] [M_gates] [nb]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ha2g_b200._lib import lib
from ha2g_b200 import ops
from ha2g_b200.ops import _p, _st
M = int(sys.argv[1]) if len(sys.argv) > 1 else 384
MG = int(sys.argv[2]) if len(sys.argv) > 2 else 128
NB = int(sys.argv[3]) if len(sys.argv) > 3 else 0
T, H, dev = 34, 300, "cuda:0"
ops._ensure_workspace()
w = [torch.randn(3 * H, H, device=dev) * 0.05 for _ in range(2)]
b = [torch.randn(3 * H, device=dev) * 0.05 for _ in range(2)]
gi = torch.randn(M, T, 6 * H, device=dev)
y = torch.empty(M, T, 2 * H, device=dev)
gates = torch.empty(max(MG, 1), T, 8 * H, device=dev)
for _ in range(5):
    lib.ha2g_gru_seq_fwd_tc2_dbg(_p(gi), _p(w[0]), _p(w[1]), _p(b[0]), _p(b[1]), _p(y), _p(gates) if MG else None, M, MG, T, H, NB, None, _st())
torch.cuda.synchronize()
