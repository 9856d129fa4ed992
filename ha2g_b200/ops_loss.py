"""Loss ops of the HA2G step (csrc/losses.cu).  Each returns a 1-element device tensor; the kernels emit
the un-scaled input gradient during forward, so backward is one multiply by the upstream device scalar."""
from __future__ import annotations

import ctypes

import torch

from . import constants as K
from ._lib import lib
from .ops import _c, _call, _chk, _p, _st


def _scaled(grad_buf, upstream):
    out = torch.empty_like(grad_buf)
    _call("ha2g_scale_by_scalar", _p(grad_buf), _p(upstream), 1.0, _p(out), grad_buf.numel(), 0, _st())
    return out


class _HuberFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o, t, beta):
        o, t = _c(o), _c(t)
        _chk(o, t)
        loss = torch.zeros((1,), device=o.device, dtype=torch.float32)
        grad = torch.empty_like(o) if ctx.needs_input_grad[0] else None
        _call("ha2g_huber", _p(o), _p(t), _p(grad), o.numel(), beta, _p(loss), _st())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, dl):
        (grad,) = ctx.saved_tensors
        return _scaled(grad, _c(dl)), None, None


def huber(o, t, beta=0.1):
    """F.smooth_l1_loss(o/beta, t/beta) * beta   (train_hierarchy_expressive.py:312-318)."""
    return _HuberFn.apply(o, t, beta)


class _LogLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode):
        x = _c(x)
        _chk(x)
        loss = torch.zeros((1,), device=x.device, dtype=torch.float32)
        grad = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        _call("ha2g_log_loss", _p(x), _p(grad), x.numel(), mode, _p(loss), _st())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, dl):
        (grad,) = ctx.saved_tensors
        return _scaled(grad, _c(dl)), None


def neg_mean_log(x):
    """-mean(log(x + 1e-8))      (generator / real-sample GAN term, :224,:322)."""
    return _LogLossFn.apply(x, 0)


def neg_mean_log1m(x):
    """-mean(log(1 - x + 1e-8))  (fake-sample GAN term, :224)."""
    return _LogLossFn.apply(x, 1)


class _KldFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, lv):
        mu, lv = _c(mu), _c(lv)
        _chk(mu, lv)
        loss = torch.zeros((1,), device=mu.device, dtype=torch.float32)
        dmu, dlv = torch.empty_like(mu), torch.empty_like(lv)
        _call("ha2g_kld", _p(mu), _p(lv), _p(dmu), _p(dlv), mu.numel(), _p(loss), _st())
        ctx.save_for_backward(dmu, dlv)
        return loss

    @staticmethod
    def backward(ctx, dl):
        dmu, dlv = ctx.saved_tensors
        dl = _c(dl)
        return _scaled(dmu, dl), _scaled(dlv, dl)


def kld(mu, logvar):
    return _KldFn.apply(mu, logvar)


class _DivRegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o, r, z, zr, beta):
        o, r, z, zr = _c(o), _c(r), _c(z), _c(zr)
        _chk(o, r, z, zr)
        B = o.shape[0]
        loss = torch.zeros((1,), device=o.device, dtype=torch.float32)
        grad = torch.empty_like(o)
        _call("ha2g_div_reg", _p(o), _p(r), _p(z), _p(zr), _p(grad), B, o.numel() // B, z.shape[1], beta, _p(loss), _st())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, dl):
        (grad,) = ctx.saved_tensors
        return _scaled(grad, _c(dl)), None, None, None, None


def div_reg(out, out_rand, z, z_rand, beta=0.05):
    """train_hierarchy_expressive.py:396-406 (out_rand, z, z_rand are treated as constants, as detached there)."""
    return _DivRegFn.apply(out, out_rand.detach(), z.detach(), z_rand.detach(), beta)


_phy_tables_loaded = {}


def _ensure_phy_tables(variant: str, mean_dir_vec):
    key = (variant, tuple(float(x) for x in mean_dir_vec), torch.cuda.current_device())
    if _phy_tables_loaded.get(variant) == key:
        return
    if variant == "expressive":
        pairs, avg, var, nb, vid = K.EXPRESSIVE_ANGLE_PAIR, K.EXPRESSIVE_AVG_ANGLE, K.EXPRESSIVE_VAR_ANGLE, 42, 1
    else:
        pairs, avg, var, nb, vid = K.GESTURE_ANGLE_PAIR, K.GESTURE_AVG_ANGLE, K.GESTURE_VAR_ANGLE, 9, 0
    n = len(pairs)
    flat = [int(v) for p in pairs for v in p]
    c_pairs = (ctypes.c_int * (2 * n))(*flat)
    c_avg = (ctypes.c_float * n)(*avg)
    c_var = (ctypes.c_float * n)(*var)
    c_mean = (ctypes.c_float * (3 * nb))(*[float(x) for x in mean_dir_vec])
    lib.ha2g_physical_set_tables(vid, ctypes.cast(c_pairs, ctypes.c_void_p), ctypes.cast(c_avg, ctypes.c_void_p),
                                 ctypes.cast(c_var, ctypes.c_void_p), n, ctypes.cast(c_mean, ctypes.c_void_p), nb)
    _phy_tables_loaded[variant] = key


class _PhysicalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, variant_id, nb, npairs):
        out = _c(out)
        _chk(out)
        rows = out.numel() // (3 * nb)
        loss = torch.zeros((1,), device=out.device, dtype=torch.float32)
        grad = torch.empty_like(out) if ctx.needs_input_grad[0] else None
        _call("ha2g_physical", _p(out), _p(grad), rows, variant_id, nb, npairs, _p(loss), _st())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, dl):
        (grad,) = ctx.saved_tensors
        return _scaled(grad, _c(dl)), None, None, None


def physical(out, variant: str, mean_dir_vec):
    """Bone-angle Gaussian NLL (train_hierarchy_expressive.py:426-449 / train_hierarchy.py:242-262)."""
    _ensure_phy_tables(variant, mean_dir_vec)
    if variant == "expressive":
        return _PhysicalFn.apply(out, 1, 42, len(K.EXPRESSIVE_ANGLE_PAIR))
    return _PhysicalFn.apply(out, 0, 9, len(K.GESTURE_ANGLE_PAIR))


class _ContrastiveFn(torch.autograd.Function):
    """Streaming SoftmaxContrastiveLoss.  Under data parallelism (world > 1) the columns b are all-gathered so that the
    loss runs over the GLOBAL batch like the reference's DataParallel step (local rows x global columns, positives at
    rank*N + i), and the gradient wrt the gathered columns is reduce-scattered back to the ranks that own them."""

    @staticmethod
    def forward(ctx, a, b, variant_id, world, rank):
        a, b = _c(a), _c(b)
        _chk(a, b)
        N, C = a.shape
        if C != 32 or b.shape != a.shape:
            raise RuntimeError("contrastive kernel expects two [N,32] feature matrices")
        dev = a.device
        if world > 1:
            import torch.distributed as dist
            b_all = torch.empty((world * N, C), device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(b_all, b)
        else:
            b_all = b
        Nb = b_all.shape[0]
        an, bn = torch.empty_like(a), torch.empty_like(b_all)
        f = lambda n: torch.empty((n,), device=dev, dtype=torch.float32)
        na, nb, sqa, sqb, lse, rowloss = f(N), f(Nb), f(N), f(Nb), f(N), f(N)
        G = torch.empty((N, Nb), device=dev, dtype=torch.float32)   # Gram matrix -> logits -> (backward) coefficients
        loss = torch.zeros((1,), device=dev, dtype=torch.float32)
        _call("ha2g_contrastive_fwd_rect", _p(a), _p(b_all), _p(an), _p(bn), _p(na), _p(nb), _p(sqa), _p(sqb), _p(lse), _p(G),
              _p(rowloss), N, Nb, rank * N, variant_id, _p(loss), _st())
        ctx.cfg = (variant_id, world, rank)
        ctx.save_for_backward(an, bn, na, nb, lse, G)
        return loss

    @staticmethod
    def backward(ctx, dl):
        an, bn, na, nb, lse, G = ctx.saved_tensors
        variant_id, world, rank = ctx.cfg
        N, Nb = an.shape[0], bn.shape[0]
        dev = an.device
        dl = _c(dl)
        da, db_all = torch.empty_like(an), torch.empty_like(bn)
        X, Y = torch.empty_like(an), torch.empty_like(bn)
        rs = torch.empty((N,), device=dev, dtype=torch.float32)
        cs = torch.empty((Nb,), device=dev, dtype=torch.float32)
        # (G is consumed in place: the loss is differentiated once per step)
        _call("ha2g_contrastive_bwd_rect", _p(an), _p(bn), _p(na), _p(nb), _p(lse), _p(dl), _p(G), _p(X), _p(Y), _p(rs), _p(cs),
              _p(da), _p(db_all), N, Nb, rank * N, variant_id, _st())
        if world > 1:
            import torch.distributed as dist
            db = torch.empty_like(an)
            dist.reduce_scatter_tensor(db, db_all, op=dist.ReduceOp.SUM)
        else:
            db = db_all
        return da, db, None, None, None


def contrastive(a, b, variant: str):
    """SoftmaxContrastiveLoss.forward without materialising the N x N x 32 tensor
    (train_hierarchy.py:54-68 gesture: 1/(D+1e-8) clamped; train_hierarchy_expressive.py:108-121: 1/D)."""
    from . import dp
    world, rank = (dp.world_size(), dp.rank()) if dp.global_contrastive() else (1, 0)
    return _ContrastiveFn.apply(a, b, 0 if variant == "gesture" else 1, world, rank)
