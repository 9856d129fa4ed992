"""Multi-tensor Adam step over a caller-owned ``torch.optim.Adam`` (csrc/adam.cu).

The reference's training loop builds ``torch.optim.Adam`` objects itself and hands them to
``train_iter_*`` (scripts/train_expressive.py:212-230,342-346).  To stay a drop-in we step THOSE objects:
hyper-parameters are read from ``param_groups`` and the moments live in ``optimizer.state`` under
torch's own keys (``step``, ``exp_avg``, ``exp_avg_sq``), so ``optimizer.state_dict()`` stays loadable by
stock torch -- but the update itself is one launch of our kernel per param group instead of torch's
foreach kernels.
"""
from __future__ import annotations

from typing import Dict

import torch

from .ops import _call, _p, _st

CHUNK = 16384
_cache: Dict[int, dict] = {}


def zero_grad(optimizer: torch.optim.Optimizer):
    """optimizer.zero_grad(set_to_none=True) without touching torch's profiler hooks."""
    for group in optimizer.param_groups:
        for p in group["params"]:
            p.grad = None


@torch.no_grad()
def fused_adam_step(optimizer: torch.optim.Optimizer):
    for gi, group in enumerate(optimizer.param_groups):
        if group.get("amsgrad", False) or group.get("weight_decay", 0) != 0 or group.get("maximize", False):
            raise NotImplementedError("fused_adam_step covers the reference's configuration (plain Adam)")
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            continue
        dev = params[0].device
        rows, sizes = [], []
        step = None
        for p in params:
            st = optimizer.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] += 1
            s = int(st["step"].item()) if torch.is_tensor(st["step"]) else int(st["step"])
            if step is None:
                step = s
            elif s != step:
                raise NotImplementedError("parameters of one group must share the Adam step count")
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            if not p.is_contiguous():
                raise RuntimeError("fused_adam_step needs contiguous parameters")
            rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()))
            sizes.append(p.numel())
        key = (id(optimizer), gi)
        c = _cache.get(key)
        if c is None or c["sizes"] != sizes or c["dev"] != dev:
            ct, co = [], []
            for k, n in enumerate(sizes):
                for off in range(0, n, CHUNK):
                    ct.append(k)
                    co.append(off)
            c = {"sizes": list(sizes), "dev": dev, "rows": None,
                 "sizes_t": torch.tensor(sizes, dtype=torch.int64, device=dev),
                 "ct": torch.tensor(ct, dtype=torch.int32, device=dev),
                 "co": torch.tensor(co, dtype=torch.int64, device=dev), "n": len(ct),
                 "pin": torch.empty((len(sizes) * 4,), dtype=torch.int64).pin_memory(),
                 "table": torch.empty((len(sizes) * 4,), dtype=torch.int64, device=dev),
                 "keep": None}
            _cache[key] = c
        if c["rows"] != rows:
            # the previous upload (if any) must have been consumed before the pinned buffer is rewritten
            if c.get("evt") is not None:
                c["evt"].synchronize()
            c["pin"].copy_(torch.tensor([a for r in rows for a in r], dtype=torch.int64))
            c["table"].copy_(c["pin"], non_blocking=True)
            c["evt"] = torch.cuda.Event()
            c["evt"].record()
            c["rows"] = rows
        b1, b2 = group["betas"]
        _call("ha2g_adam_multi", _p(c["table"]), _p(c["sizes_t"]), _p(c["ct"]), _p(c["co"]), c["n"], float(group["lr"]),
              float(b1), float(b2), float(group["eps"]), step, _st())
