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
    """Eager step.  Runs the SAME kernels as the captured step (``ha2g_adam_multi_dev``: step counter, hyper-parameters and
    bias corrections in device memory), so an eager step and a CUDA-graph replay of it are bit-identical."""
    for gi, group in enumerate(optimizer.param_groups):
        if group.get("amsgrad", False) or group.get("weight_decay", 0) != 0 or group.get("maximize", False):
            raise NotImplementedError("fused_adam_step covers the reference's configuration (plain Adam)")
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            continue
        dev = params[0].device
        rows, sizes, keep = [], [], []
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
            keep.append(g)   # a contiguous COPY must stay alive until the launch below has been enqueued
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
                 "hyper": torch.zeros((8,), dtype=torch.float64, device=dev), "hyper_host": None,
                 "step": torch.zeros((1,), dtype=torch.int32, device=dev), "step_host": 0,
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
        c["keep"] = keep
        hh = _hyper_of(group)
        if hh != c["hyper_host"]:
            c["hyper"][:4].copy_(torch.tensor(hh, dtype=torch.float64))
            c["hyper_host"] = hh
        if c["step_host"] != step - 1:   # the kernel increments the device counter itself
            c["step"].fill_(step - 1)
        c["step_host"] = step
        _call("ha2g_adam_multi_dev", _p(c["table"]), _p(c["sizes_t"]), _p(c["ct"]), _p(c["co"]), c["n"], _p(c["hyper"]),
              _p(c["step"]), _st())


# ----------------------------------------------------------------------------------------------------------------
# CUDA-graph variant: step count and hyper-parameters in device memory (csrc/adam.cu::ha2g_adam_multi_dev)
# ----------------------------------------------------------------------------------------------------------------
def _hyper_of(group):
    b1, b2 = group["betas"]
    return (float(group["lr"]), float(b1), float(b2), float(group["eps"]))


def graph_prepare(optimizer: torch.optim.Optimizer):
    """Outside capture, after at least one eager step (moments exist, the set of parameters that receive a gradient is
    known): allocate the per-group device tables the captured Adam launch will read."""
    entries = []
    for group in optimizer.param_groups:
        if group.get("amsgrad", False) or group.get("weight_decay", 0) != 0 or group.get("maximize", False):
            raise NotImplementedError("fused_adam_step covers the reference's configuration (plain Adam)")
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            entries.append(None)
            continue
        dev = params[0].device
        sizes = [p.numel() for p in params]
        ct, co = [], []
        for k, n in enumerate(sizes):
            for off in range(0, n, CHUNK):
                ct.append(k)
                co.append(off)
        steps = {int(optimizer.state[p]["step"]) for p in params}
        if len(steps) != 1:
            raise NotImplementedError("parameters of one group must share the Adam step count")
        step0 = steps.pop()
        hh = _hyper_of(group)
        entries.append({
            "params": params, "n": len(ct),
            "sizes_t": torch.tensor(sizes, dtype=torch.int64, device=dev),
            "ct": torch.tensor(ct, dtype=torch.int32, device=dev),
            "co": torch.tensor(co, dtype=torch.int64, device=dev),
            "table": torch.zeros((4 * len(params),), dtype=torch.int64, device=dev),
            "hyper": torch.tensor(list(hh) + [0.0] * 4, dtype=torch.float64, device=dev), "hyper_host": hh,
            "step": torch.tensor([step0], dtype=torch.int32, device=dev), "step_host": step0,
            "rows": None, "grads": None})
    return entries


@torch.no_grad()
def graph_adam_enqueue(optimizer: torch.optim.Optimizer, entries):
    """Inside capture: record the (p, g, m, v) addresses (uploaded by graph_finalize once capture has ended) and
    enqueue the device-hyper Adam launch."""
    for group, e in zip(optimizer.param_groups, entries):
        params = [p for p in group["params"] if p.grad is not None]
        if e is None:
            if params:
                raise RuntimeError("a parameter group without gradients in the warm-up steps received one during capture")
            continue
        if len(params) != len(e["params"]) or any(a is not b for a, b in zip(params, e["params"])):
            raise RuntimeError("the set of parameters receiving gradients changed between warm-up and capture")
        rows = []
        for p in params:
            st = optimizer.state[p]
            if not p.grad.is_contiguous() or not p.is_contiguous():
                raise RuntimeError("fused Adam needs contiguous parameters and gradients")
            rows += [p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()]
        e["rows"], e["grads"] = rows, [p.grad for p in params]
        _call("ha2g_adam_multi_dev", _p(e["table"]), _p(e["sizes_t"]), _p(e["ct"]), _p(e["co"]), e["n"], _p(e["hyper"]),
              _p(e["step"]), _st())


def graph_finalize(entries):
    """After capture: upload the address tables recorded during capture."""
    for e in entries:
        if e is not None:
            e["table"].copy_(torch.tensor(e["rows"], dtype=torch.int64))


def graph_pre_replay(optimizer: torch.optim.Optimizer, entries):
    """Host side of one replay: keep torch's own optimizer state (``step``) and the device copies of the
    hyper-parameters in agreement, and point ``p.grad`` at the graph's static gradient buffers."""
    for group, e in zip(optimizer.param_groups, entries):
        if e is None:
            continue
        hh = _hyper_of(group)
        if hh != e["hyper_host"]:  # e.g. an lr schedule: rare, a small synchronous upload
            e["hyper"][:4].copy_(torch.tensor(hh, dtype=torch.float64))
            e["hyper_host"] = hh
        params = e["params"]
        s = int(optimizer.state[params[0]]["step"])
        if s != e["step_host"]:   # an eager step ran in between
            e["step"].fill_(s)
        for p, g in zip(params, e["grads"]):
            optimizer.state[p]["step"] += 1
            p.grad = g
        e["step_host"] = s + 1
