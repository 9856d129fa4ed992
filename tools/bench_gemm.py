"""Micro-benchmark of the dense GEMM launchers on the step's shapes (CUDA events, L2-cold via a 256 MB flush)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ha2g_b200 import ops

dev = "cuda:0"
flush = torch.empty(64 << 20, device=dev)
shapes = [  # (M, N, K, tA, tB, split, note)
    (4352, 900, 600, 0, 1, 1, "gi projection l1-3"),
    (4352, 900, 207, 0, 1, 1, "gi projection l0 (unaligned lda)"),
    (900, 600, 4352, 1, 0, 4, "dW_ih"),
    (4352, 600, 900, 0, 0, 1, "dx"),
    (4352, 300, 600, 0, 1, 1, "TCN conv"),
    (4352, 32, 4032, 0, 1, 1, "fc_low"),
    (128, 16, 16, 0, 1, 1, "speaker linear"),
]
for impl in ("f32", "tc", "tc2"):
    ops.set_gemm_impl(impl)
    for (M, N, K, tA, tB, split, note) in shapes:
        A = torch.randn((K, M) if tA else (M, K), device=dev)
        B = torch.randn((N, K) if tB else (K, N), device=dev)
        C = torch.zeros((M, N), device=dev)
        for _ in range(3):
            ops.gemm(A, B, C, None, M, N, K, A.shape[1], B.shape[1], N, tA, tB, 0, 0, split)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(A, B, C, None, M, N, K, A.shape[1], B.shape[1], N, tA, tB, 0, 0, split)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(f"{impl:4s} {M:5d}x{N:4d}x{K:5d} tA={tA} tB={tB} split={split}: {t * 1e3:8.1f} us  {2.0 * M * N * K / t / 1e9:8.2f} TFLOP/s  ({note})")

# packed GEMM alone (operands already packed): isolates the tensor-core main loop from the packing passes
import ctypes
from ha2g_b200._lib import lib
from ha2g_b200.ops import _p, _st
for (M, N, K, note) in [(4352, 900, 600, "gi projection"), (900, 600, 4352, "dW_ih shape"), (4352, 300, 600, "TCN")]:
    rp = ctypes.c_int(); cp = ctypes.c_int(); rpb = ctypes.c_int(); cpb = ctypes.c_int()
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.zeros(M, N, device=dev)
    lib.ha2g_pack_dims(M, K, ctypes.addressof(rp), ctypes.addressof(cp))
    lib.ha2g_pack_dims(N, K, ctypes.addressof(rpb), ctypes.addressof(cpb))
    ah = torch.empty(rp.value * cp.value * 16, dtype=torch.uint8, device=dev); al = torch.empty_like(ah)
    bh = torch.empty(rpb.value * cp.value * 16, dtype=torch.uint8, device=dev); bl = torch.empty_like(bh)
    lib.ha2g_pack_bf16x2(_p(A), K, M, K, 1, 0, 0, _p(ah), _p(al), _st())
    lib.ha2g_pack_bf16x2(_p(B), K, N, K, 1, 0, 0, _p(bh), _p(bl), _st())
    for terms in (3, 1):
        for split in (1, 4):
            ts = []
            for it in range(6):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                lib.ha2g_gemm_packed(_p(ah), _p(al), rp.value, _p(bh), _p(bl), rpb.value, _p(C), None, M, N, cp.value, N, 0, 0, split, terms, _st())
                e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            print(f"packed-only terms={terms} split={split} {M}x{N}x{K}: {t * 1e3:7.1f} us  {2.0 * M * N * K / t / 1e9:8.2f} TFLOP/s fp32-equiv ({note})")
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.ha2g_pack_bf16x2(_p(A), K, M, K, 1, 0, 0, _p(ah), _p(al), _st()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"pack A {M}x{K}: {sorted(ts)[1] * 1e3:.1f} us")
