"""Differentiable ops of the ResNetSE-34 audio encoder (channels-last / NHWC), backed by
csrc/conv2d.cu, csrc/bn.cu and csrc/audio.cu.  See ha2g_b200/ops.py for the conventions."""
from __future__ import annotations

import torch

import ctypes
import os

from . import ops
from ._lib import lib
from .ops import _c, _call, _chk, _p, _st

_CONV_IMPL = os.environ.get("HA2G_CONV_IMPL", "tc")  # "tc": tcgen05 for stride-1 fwd/dgrad; "f32": SIMT everywhere
# operand split on the tensor cores: tf32x3 (default; ~2^-21 product error, needed by the ill-conditioned train-mode-BN
# encoder gradients) or bf16x3 (2^-16; half the tensor time and operand bytes)
_CONV_PREC = 0 if os.environ.get("HA2G_CONV_PRECISION", "tf32x3") == "bf16x3" else 1


_S2_TC = os.environ.get("HA2G_CONV_S2_TC", "1") != "0"   # stride-2 convolutions rewritten as stride-1 ones for tcgen05
_WGRAD_IMPLICIT = os.environ.get("HA2G_WGRAD_IMPLICIT", "1") != "0"  # tap-shifted MN-major tcgen05 kernel for "same" convolutions
_WGRAD_TC_MIN_CIN = int(os.environ.get("HA2G_WGRAD_TC_MIN_CIN", "64"))  # below: SIMT kernel (the stem has its own)
_WGRAD_TC = os.environ.get("HA2G_WGRAD_IMPL", "tc") == "tc"  # weight gradient on the packed tcgen05 GEMM (bf16x3)


def set_conv_precision(mode: str):
    global _CONV_PREC
    _CONV_PREC = 0 if mode == "bf16x3" else 1


def set_conv_impl(impl: str):
    global _CONV_IMPL
    _CONV_IMPL = impl


def _tc_ok(Cs, Cd, stride):
    return _CONV_IMPL == "tc" and stride == 1 and Cs % (4 if _CONV_PREC else 8) == 0 and Cd % 4 == 0


def _conv_tc(src, w, bias, N, H, W, Cs, Cd, pad, KH, KW, dgrad):
    """Stride-1 correlation of the NHWC tensor `src` [N,H,W,Cs] with the OIHW weight `w` on tcgen05 (csrc/conv_tc.cu)."""
    dev = src.device
    guard, rows_p, chunks_p = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.ha2g_conv_tc_dims(N, H, W, Cs, pad, KH, KW, _CONV_PREC, ctypes.addressof(guard), ctypes.addressof(rows_p), ctypes.addressof(chunks_p))
    a = torch.empty((2, rows_p.value * chunks_p.value * 16), dtype=torch.uint8, device=dev)
    _call("ha2g_conv_tc_pack_act", _p(src), N, H, W, Cs, pad, KH, KW, _CONV_PREC, _p(a[0]), _p(a[1]), _st())
    Cout, Cin = w.shape[0], w.shape[1]
    rows_pb = (Cd + 255) // 256 * 256
    b = torch.empty((2, rows_pb * KH * KW * chunks_p.value * 16), dtype=torch.uint8, device=dev)
    _call("ha2g_conv_tc_pack_w", _p(w), Cout, Cin, KH, KW, int(dgrad), _CONV_PREC, _p(b[0]), _p(b[1]), _st())
    Ho, Wo = H + 2 * pad - KH + 1, W + 2 * pad - KW + 1
    out = torch.empty((N, Ho, Wo, Cd), device=dev, dtype=torch.float32)
    _call("ha2g_conv_tc", _p(a[0]), _p(a[1]), _p(b[0]), _p(b[1]), _p(bias), _p(out), N, H, W, Cs, Cd, pad, KH, KW, _CONV_PREC, _st())
    return out


class _Conv2dFn(torch.autograd.Function):
    """nn.Conv2d on NHWC activations; weight stays in the checkpoint layout OIHW."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad):
        x = _c(x)
        _chk(x, w, b)
        N, H, W, Cin = x.shape
        Cout, _, KH, KW = w.shape
        Ho = (H + 2 * pad - KH) // stride + 1
        Wo = (W + 2 * pad - KW) // stride + 1
        ctx.cfg = (N, H, W, Cin, Cout, KH, KW, stride, pad)
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, w)
        if _tc_ok(Cin, Cout, stride):
            return _conv_tc(x, w, b, N, H, W, Cin, Cout, pad, KH, KW, False)
        wf = torch.empty((KH * KW * Cin, Cout), device=x.device, dtype=torch.float32)
        _call("ha2g_conv2d_pack", _p(w), _p(wf), Cout, Cin, KH, KW, 0, _st())
        y = torch.empty((N, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
        _call("ha2g_conv2d_fwd", _p(x), _p(wf), _p(b), _p(y), N, H, W, Cin, Cout, KH, KW, stride, pad, 0, _st())
        ctx.cfg = (N, H, W, Cin, Cout, KH, KW, stride, pad)
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, Cin, Cout, KH, KW, stride, pad = ctx.cfg
        x, w = ctx.saved_tensors
        dy = _c(dy)
        dev = dy.device
        dx = dw = db = None
        if ctx.needs_input_grad[0] and _tc_ok(Cout, Cin, stride):
            Ho, Wo = dy.shape[1], dy.shape[2]
            dx = _conv_tc(dy, w, None, N, Ho, Wo, Cout, Cin, KH - 1 - pad, KH, KW, True)
        elif ctx.needs_input_grad[0]:
            wb = torch.empty((KH * KW * Cout, Cin), device=dev, dtype=torch.float32)
            _call("ha2g_conv2d_pack", _p(w), _p(wb), Cout, Cin, KH, KW, 1, _st())
            dx = torch.empty((N, H, W, Cin), device=dev, dtype=torch.float32)
            _call("ha2g_conv2d_dgrad", _p(dy), _p(wb), _p(dx), N, H, W, Cin, Cout, KH, KW, stride, pad, _st())
        if ctx.needs_input_grad[1]:
            dwf = ops.zeros((KH * KW * Cin, Cout), dev)
            ok2, nb2, ok3 = ctypes.c_int(0), ctypes.c_int64(0), ctypes.c_int(0)
            tc_w = _CONV_IMPL == "tc" and stride == 1 and _WGRAD_TC and _WGRAD_IMPLICIT
            if tc_w and dy.shape[1] == H and dy.shape[2] == W:
                lib.ha2g_conv_wgrad_tc2_workspace(N, H, W, Cin, Cout, KH, KW, pad, ctypes.addressof(ok2), ctypes.addressof(nb2))
            elif tc_w and KH == 2 and KW == 2 and pad == 1 and dy.shape[1] == H + 1 and dy.shape[2] == W + 1:
                lib.ha2g_conv_wgrad_tc2_workspace(N, H + 1, W + 1, Cin, Cout, 3, 3, 1, ctypes.addressof(ok3), ctypes.addressof(nb2))
            if ok3.value:
                # 2x2 / pad 1 (the stride-2 convolutions rewritten over the space-to-depth input): out[i,j] reads x rows
                # i-1, i and columns j-1, j -- exactly taps (r,s) in {0,1}^2 of a 3x3 "same" convolution over the
                # (H+1) x (W+1) output grid with x zero-extended by one row and column.  The implicit kernel computes all nine
                # taps (2.25x the MMAs of the four that are kept, still ~4x faster than im2col + GEMM).
                xp = ops.zeros((N, H + 1, W + 1, Cin), dev)
                xp[:, :H, :W, :] = x
                dwf3 = ops.zeros((9 * Cin, Cout), dev)
                ws = torch.empty((nb2.value,), dtype=torch.uint8, device=dev)
                _call("ha2g_conv_wgrad_tc2", _p(xp), _p(dy), _p(dwf3), N, H + 1, W + 1, Cin, Cout, 3, 3, 1, _p(ws), nb2.value, _st())
                d3 = dwf3.view(3, 3, Cin, Cout)
                dwf = d3[:2, :2].contiguous().view(4 * Cin, Cout)
            elif ok2.value:
                # implicit GEMM over the packed, zero-padded activations: no im2col (csrc/conv_wgrad_tc2.cu)
                ws = torch.empty((nb2.value,), dtype=torch.uint8, device=dev)
                _call("ha2g_conv_wgrad_tc2", _p(x), _p(dy), _p(dwf), N, H, W, Cin, Cout, KH, KW, pad, _p(ws), nb2.value, _st())
            elif _CONV_IMPL == "tc" and stride == 1 and _WGRAD_TC and Cin >= _WGRAD_TC_MIN_CIN:
                nbytes, blk = ctypes.c_int64(), ctypes.c_int()
                lib.ha2g_conv_wgrad_tc_workspace(Cin, Cout, KH, KW, ctypes.addressof(nbytes), ctypes.addressof(blk))
                ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=dev)
                _call("ha2g_conv_wgrad_tc", _p(x), _p(dy), _p(dwf), N, H, W, Cin, Cout, KH, KW, pad, _p(ws), nbytes.value, _st())
            else:
                _call("ha2g_conv2d_wgrad", _p(x), _p(dy), _p(dwf), N, H, W, Cin, Cout, KH, KW, stride, pad, _st())
            dw = torch.empty_like(w)
            _call("ha2g_conv2d_pack", _p(dwf), _p(dw), Cout, Cin, KH, KW, 2, _st())
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.zeros((Cout,), dev)
            rows = dy.numel() // Cout
            _call("ha2g_col_sum", _p(dy), rows, Cout, Cout, _p(db), _st())
        return dx, dw, db, None, None


class _PixelRemapFn(torch.autograd.Function):
    """space-to-depth by 2 (mode 's2d') or every-second-pixel subsampling (mode 'sub') of an NHWC tensor."""

    @staticmethod
    def forward(ctx, x, mode):
        x = _c(x)
        _chk(x)
        N, H, W, C = x.shape
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        fn = "ha2g_space_to_depth2" if mode == "s2d" else "ha2g_subsample2"
        y = torch.empty((N, H2, W2, 4 * C if mode == "s2d" else C), device=x.device, dtype=torch.float32)
        _call(fn, _p(x), _p(y), N, H, W, C, 0, _st())
        ctx.cfg = (N, H, W, C, fn)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, C, fn = ctx.cfg
        dy = _c(dy)
        dx = torch.empty((N, H, W, C), device=dy.device, dtype=torch.float32)
        _call(fn, _p(dy), _p(dx), N, H, W, C, 1, _st())
        return dx, None


class _S2WeightFn(torch.autograd.Function):
    """3x3 stride-2 weight [Cout,C,3,3] -> the equivalent 2x2 stride-1 weight over the space-to-depth input [Cout,4C,2,2]."""

    @staticmethod
    def forward(ctx, w):
        _chk(w)
        Cout, C = w.shape[0], w.shape[1]
        w2 = torch.empty((Cout, 4 * C, 2, 2), device=w.device, dtype=torch.float32)
        _call("ha2g_conv_s2_weight", _p(w), _p(w2), Cout, C, 0, _st())
        ctx.cfg = (Cout, C)
        return w2

    @staticmethod
    def backward(ctx, dw2):
        Cout, C = ctx.cfg
        dw2 = _c(dw2)
        dw = torch.empty((Cout, C, 3, 3), device=dw2.device, dtype=torch.float32)
        _call("ha2g_conv_s2_weight", _p(dw2), _p(dw), Cout, C, 1, _st())
        return dw


class _CropFn(torch.autograd.Function):
    """y[:, :Ho, :Wo, :] of an NHWC tensor as a contiguous tensor (and zero-padding in backward), via ha2g_copy_cols-free
    strided copies of torch (data movement only)."""

    @staticmethod
    def forward(ctx, y, Ho, Wo):
        ctx.shape = y.shape
        return y[:, :Ho, :Wo, :].contiguous()

    @staticmethod
    def backward(ctx, d):
        full = torch.zeros(ctx.shape, device=d.device, dtype=d.dtype)
        full[:, :d.shape[1], :d.shape[2], :] = d
        return full, None, None


def conv2d(x, w, b=None, stride=1, pad=0):
    """nn.Conv2d on NHWC activations.  The stride-2 convolutions of the encoder are rewritten as stride-1 convolutions
    (3x3/pad 1 -> 2x2/pad 1 over the space-to-depth input with a remapped weight; 1x1 -> 1x1 over the subsampled input)
    so that forward, data gradient and weight gradient all run on the tcgen05 kernels."""
    KH = w.shape[2]
    if stride == 2 and _CONV_IMPL == "tc" and _S2_TC and x.shape[3] % 8 == 0 and w.shape[0] % 4 == 0:
        if KH == 1 and pad == 0:
            return _Conv2dFn.apply(_PixelRemapFn.apply(x, "sub"), w, b, 1, 0)
        if KH == 3 and pad == 1 and w.shape[3] == 3:
            Ho, Wo = (x.shape[1] + 1) // 2, (x.shape[2] + 1) // 2
            y = _Conv2dFn.apply(_PixelRemapFn.apply(x, "s2d"), _S2WeightFn.apply(w), b, 1, 1)
            return _CropFn.apply(y, Ho, Wo)
    return _Conv2dFn.apply(x, w, b, stride, pad)


class _StemConvFn(torch.autograd.Function):
    """Conv2d(1, C, 3, padding=1) on the (B,128,70) log-mel image (ResNetSE34V2.py:27,125)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = _c(x)
        _chk(x, w, b)
        N, H, W = x.shape
        Cout = w.shape[0]
        y = torch.empty((N, H, W, Cout), device=x.device, dtype=torch.float32)
        _call("ha2g_stem_conv_fwd", _p(x), _p(w), _p(b), _p(y), N, H, W, Cout, _st())
        ctx.save_for_backward(x)
        ctx.cfg = (N, H, W, Cout)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        N, H, W, Cout = ctx.cfg
        dy = _c(dy)
        dw = ops.zeros((Cout, 1, 3, 3), dy.device)
        db = ops.zeros((Cout,), dy.device)
        _call("ha2g_stem_conv_wgrad", _p(x), _p(dy), _p(dw), _p(db), N, H, W, Cout, _st())
        return None, dw, db  # the spectrogram is an input, never differentiated


def stem_conv(x, w, b):
    return _StemConvFn.apply(x, w, b)


class _SEFn(torch.autograd.Function):
    """out = relu(u * sigmoid(W2 relu(W1 gap(u) + b1) + b2) + res)   (ResNetBlocks.py:29-37,81-96)."""

    @staticmethod
    def forward(ctx, u, res, w1, b1, w2, b2):
        u, res = _c(u), _c(res)
        _chk(u, res, w1, b1, w2, b2)
        N, H, W, C = u.shape
        R = w1.shape[0]
        dev = u.device
        gap = torch.empty((N, C), device=dev, dtype=torch.float32)
        h = torch.empty((N, R), device=dev, dtype=torch.float32)
        s = torch.empty((N, C), device=dev, dtype=torch.float32)
        out = torch.empty_like(u)
        _call("ha2g_se_fwd", _p(u), _p(res), _p(w1), _p(b1), _p(w2), _p(b2), _p(gap), _p(h), _p(s), _p(out), N, H * W, C, R,
              _st())
        ctx.cfg = (N, H * W, C, R)
        ctx.save_for_backward(u, out, gap, h, s, w1, w2)
        return out

    @staticmethod
    def backward(ctx, dout):
        N, HW, C, R = ctx.cfg
        u, out, gap, h, s, w1, w2 = ctx.saved_tensors
        dout = _c(dout)
        dev = dout.device
        dres = torch.empty_like(u)
        du = torch.empty_like(u)
        ds = torch.empty((N, C), device=dev, dtype=torch.float32)
        dgap = torch.empty((N, C), device=dev, dtype=torch.float32)
        dw1 = ops.zeros_like(w1)
        db1 = ops.zeros((R,), dev)
        dw2 = ops.zeros_like(w2)
        db2 = ops.zeros((C,), dev)
        _call("ha2g_se_bwd", _p(dout), _p(out), _p(u), _p(gap), _p(h), _p(s), _p(w1), _p(w2), _p(dres), _p(du), _p(ds),
              _p(dgap), _p(dw1), _p(db1), _p(dw2), _p(db2), N, HW, C, R, _st())
        return du, dres, dw1, db1, dw2, db2


def se_residual_relu(u, res, w1, b1, w2, b2):
    return _SEFn.apply(u, res, w1, b1, w2, b2)


class _PixelShuffleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, r):
        x = _c(x)
        _chk(x)
        N, H, W, C = x.shape
        Co = C // (r * r)
        y = torch.empty((N, H * r, W * r, Co), device=x.device, dtype=torch.float32)
        _call("ha2g_pixel_shuffle", _p(x), _p(y), N, H, W, Co, r, 0, _st())
        ctx.cfg = (N, H, W, Co, r)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, Co, r = ctx.cfg
        dy = _c(dy)
        dx = torch.empty((N, H, W, Co * r * r), device=dy.device, dtype=torch.float32)
        _call("ha2g_pixel_shuffle", _p(dy), _p(dx), N, H, W, Co, r, 1, _st())
        return dx, None


def pixel_shuffle(x, r):
    return _PixelShuffleFn.apply(x, r)


class _HeadFlattenFn(torch.autograd.Function):
    """[N,F,T,C] (NHWC) -> [N,T,C*F] with feature index c*F+f (ResNetSE34V2.py:160-162)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        _chk(x)
        N, F, T, C = x.shape
        y = torch.empty((N, T, C * F), device=x.device, dtype=torch.float32)
        _call("ha2g_head_flatten", _p(x), _p(y), N, F, T, C, 0, _st())
        ctx.cfg = (N, F, T, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, F, T, C = ctx.cfg
        dy = _c(dy)
        dx = torch.empty((N, F, T, C), device=dy.device, dtype=torch.float32)
        _call("ha2g_head_flatten", _p(dy), _p(dx), N, F, T, C, 1, _st())
        return dx


def head_flatten(x):
    return _HeadFlattenFn.apply(x)


class _BlendFn(torch.autograd.Function):
    """softmax over the 3 levels + L weighted sums (ResNetSE34V2.py:202-212)."""

    @staticmethod
    def forward(ctx, logits, f0, f1, f2, L):
        logits, f0, f1, f2 = _c(logits), _c(f0), _c(f1), _c(f2)
        _chk(logits, f0, f1, f2)
        B = f0.shape[0]
        TC = f0.numel() // B
        weight = torch.empty((B, 3, L), device=f0.device, dtype=torch.float32)
        blend = torch.empty((L, *f0.shape), device=f0.device, dtype=torch.float32)
        _call("ha2g_blend_fwd", _p(logits), _p(f0), _p(f1), _p(f2), _p(weight), _p(blend), B, TC, L, _st())
        ctx.cfg = (B, TC, L)
        ctx.save_for_backward(weight, f0, f1, f2)
        return weight, blend

    @staticmethod
    def backward(ctx, dweight, dblend):
        B, TC, L = ctx.cfg
        weight, f0, f1, f2 = ctx.saved_tensors
        dev = f0.device
        dblend = _c(dblend) if dblend is not None else torch.zeros((L, *f0.shape), device=dev, dtype=torch.float32)
        dweight = _c(dweight) if dweight is not None else None
        df0, df1, df2 = torch.empty_like(f0), torch.empty_like(f1), torch.empty_like(f2)
        dlogits = torch.empty((B, 3 * L), device=dev, dtype=torch.float32)
        _call("ha2g_blend_bwd", _p(weight), _p(f0), _p(f1), _p(f2), _p(dweight), _p(dblend), _p(df0), _p(df1), _p(df2),
              _p(dlogits), B, TC, L, _st())
        return dlogits, df0, df1, df2, None


def speaker_blend(logits, f0, f1, f2, L):
    return _BlendFn.apply(logits, f0, f1, f2, L)
