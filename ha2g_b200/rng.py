"""Random draws of the HA2G step, behind one injectable source.

The reference step consumes randomness in three places: ``reparameterize`` noise
(scripts/model/embedding_net.py:10-13, one (B,16) draw per generator call, also in eval mode),
dropout masks (embedding 0.1, TCN 0.3, GRU inter-layer 0.3) and the speaker ``torch.randperm``
(scripts/train_eval/train_hierarchy_expressive.py:328).  cuDNN's dropout stream is not reproducible,
so the reference is only defined up to these draws; parity tests inject them.

Default source: torch's CUDA generator (device-side Philox; plumbing, not a product kernel).
"""
from __future__ import annotations

import contextlib
from typing import Callable, List, Optional

import torch

_state = {"dropout": True, "randn": None, "mask": None, "randperm": None, "graph_safe": False}


def dropout_enabled() -> bool:
    return _state["dropout"]


def randn(shape, device) -> torch.Tensor:
    if _state["randn"] is not None:
        return _state["randn"](tuple(shape)).to(device=device, dtype=torch.float32).contiguous()
    return torch.randn(tuple(shape), device=device, dtype=torch.float32)


def dropout_mask(shape, p: float, device) -> torch.Tensor:
    """float32 {0,1} keep-mask with P(keep) = 1-p."""
    if _state["mask"] is not None:
        return _state["mask"](tuple(shape), p).to(device=device, dtype=torch.float32).contiguous()
    return (torch.rand(tuple(shape), device=device) >= p).to(torch.float32)


# ---- fused dropout stream (csrc/elementwise.cu::dropout_kernel): Philox keyed by a device-resident {seed, step} pair ----
_dev_state = {}
_calls = [0]


def dropout_state(device) -> torch.Tensor:
    """uint64 {seed, step} in device memory (as int64); the seed is drawn once from torch's CPU generator, so
    torch.manual_seed(...) before the first use makes runs reproducible."""
    key = str(device)
    if key not in _dev_state:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        # data-parallel ranks usually share torch.manual_seed: fold the rank in so that every rank draws its own masks
        # (the reference's DataParallel replicas draw independent dropout streams too)
        from . import dp
        seed = (seed ^ (dp.rank() * 0x9E3779B97F4A7C15)) & (2 ** 62 - 1)
        _dev_state[key] = torch.tensor([seed, 0], dtype=torch.int64, device=device)
    return _dev_state[key]


def next_call_id() -> int:
    _calls[0] = (_calls[0] + 1) & 0x7FFFFFFF
    return _calls[0]


def begin_step(device):
    """Once per training step (inside the captured graph too): tick the device-side step counter and restart the
    host-side call numbering, so a replayed graph draws fresh masks with the same baked-in call ids."""
    from .ops import _call, _p, _st, begin_step as _ops_begin_step
    _ops_begin_step(device)
    _call("ha2g_rng_tick", _p(dropout_state(device)), _st())
    _calls[0] = 0


def fused_dropout() -> bool:
    """True when dropout masks come from the fused Philox kernel (no mask injected by a test)."""
    return _state["mask"] is None


def randperm(n: int, device) -> torch.Tensor:
    if _state["randperm"] is not None:
        return _state["randperm"](n).to(device)
    # rank of n uniform keys (csrc/elementwise.cu::rank_perm_kernel): stream-ordered and CUDA-graph capturable,
    # which torch.randperm's host-side checks are not
    from .ops import _call, _p, _st
    keys = torch.rand((n,), device=device, dtype=torch.float32)
    perm = torch.empty((n,), device=device, dtype=torch.int64)
    _call("ha2g_rank_perm", _p(keys), _p(perm), n, _st())
    return perm


def overridden() -> bool:
    """True while any draw is injected (tests): such steps must not be captured into a replayable CUDA graph, unless
    the injected sources are stateless device-side functions (``override(graph_safe=True)``)."""
    return (not _state["graph_safe"]) and any(_state[k] is not None for k in ("randn", "mask", "randperm"))


@contextlib.contextmanager
def override(randn_fn: Optional[Callable] = None, mask_fn: Optional[Callable] = None,
             randperm_fn: Optional[Callable] = None, dropout: Optional[bool] = None, graph_safe: bool = False):
    """Inject deterministic draws (tests) or switch dropout off (parity with the no-dropout goldens)."""
    old = dict(_state)
    if randn_fn is not None:
        _state["randn"] = randn_fn
    if mask_fn is not None:
        _state["mask"] = mask_fn
    if randperm_fn is not None:
        _state["randperm"] = randperm_fn
    if dropout is not None:
        _state["dropout"] = dropout
    _state["graph_safe"] = bool(graph_safe)
    try:
        yield
    finally:
        _state.update(old)


class ListFeed:
    """randn_fn that replays a fixed list of tensors in call order."""

    def __init__(self, tensors: List[torch.Tensor]):
        self.tensors, self.i = list(tensors), 0

    def __call__(self, shape):
        t = self.tensors[self.i]
        self.i += 1
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t
