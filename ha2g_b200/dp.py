"""Data parallelism for the HA2G step: one process per GPU, gradients all-reduced over NCCL (NVLink 5 / NVSwitch).

The reference wraps every module in single-process ``nn.DataParallel`` (scripts/train_expressive.py:184-197):
scatter the batch, replicate weights, gather outputs, reduce-add gradients onto GPU 0, step there.  Here each
rank owns a full replica and its own B_local clips; after each backward the gradients of the optimizer about to step
are averaged across ranks (bucketed flat all-reduce), so every replica applies the identical Adam update.
BatchNorm statistics stay per-rank, exactly like DataParallel's per-replica statistics.

The contrastive loss runs over the GLOBAL batch like the reference's DataParallel step: each rank
all-gathers the 32-wide audio features, evaluates its own rows against all columns, and reduce-scatters the column
gradients (ops_loss._ContrastiveFn).  ``HA2G_DP_LOCAL_CONTRASTIVE=1`` keeps the loss rank-local (no collective in
the loss; c_pos/c_neg then differ by the log of the world size in their softmax normaliser).
"""
from __future__ import annotations

import os

from typing import List, Optional

import torch

_state = {"world": 1, "enabled": False, "rank": 0,
          "global_contrastive": os.environ.get("HA2G_DP_LOCAL_CONTRASTIVE", "0") != "1"}
BUCKET_BYTES = 64 << 20


def enable(world_size: int, modules: Optional[List[torch.nn.Module]] = None, broadcast: bool = True):
    """Call once after torch.distributed.init_process_group('nccl' | 'gloo').  Broadcasts rank 0's parameters
    and buffers so all replicas start identical."""
    import torch.distributed as dist
    _state["world"], _state["enabled"] = world_size, world_size > 1
    _state["rank"] = dist.get_rank() if world_size > 1 else 0
    if broadcast and modules and world_size > 1:
        for m in modules:
            for t in list(m.parameters()) + list(m.buffers()):
                dist.broadcast(t.data, src=0)


def disable():
    _state["world"], _state["enabled"], _state["rank"] = 1, False, 0


def rank() -> int:
    return _state["rank"]


def global_contrastive() -> bool:
    return _state["enabled"] and _state["global_contrastive"]


def world_size() -> int:
    return _state["world"]


@torch.no_grad()
def allreduce_grads(optimizer: torch.optim.Optimizer):
    """Average the gradients of ``optimizer``'s parameters over all ranks (flat buckets of <= 64 MiB)."""
    if not _state["enabled"]:
        return
    import torch.distributed as dist
    grads = [p.grad for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
    if not grads:
        return
    w = float(_state["world"])
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat)
        flat.div_(w)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * 4
        if size >= BUCKET_BYTES:
            flush()
    flush()
