"""Data parallelism for the HA2G step: one process per GPU, gradients all-reduced over NCCL (NVLink 5 / NVSwitch).

The reference wraps every module in single-process ``nn.DataParallel`` (scripts/train_expressive.py:184-197):
scatter the batch, replicate weights, gather outputs, reduce-add gradients onto GPU 0, step there.  Here each
rank owns a full replica and its own B_local clips; the gradients of each optimizer are averaged across ranks in place
(one coalesced NCCL all-reduce, launched from autograd hooks while the rest of the backward pass still runs), so every
replica applies the identical Adam update.
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
    end_backward()
    _known.clear()
    _inflight.clear()


def rank() -> int:
    return _state["rank"]


def global_contrastive() -> bool:
    return _state["enabled"] and _state["global_contrastive"]


def world_size() -> int:
    return _state["world"]


# ---------------------------------------------------------------------------------------------------------------
# gradient averaging: in place, coalesced, overlapped with the backward pass
# ---------------------------------------------------------------------------------------------------------------
# One coalesced NCCL all-reduce (ncclAvg) per optimizer over the gradient tensors themselves: no flatten copy, no
# divide pass, no copy-back.  During the generator step's backward the all-reduce of an optimizer is LAUNCHED from
# autograd hooks the moment the last gradient of that optimizer has been accumulated (generator k's gradients are final
# when its BPTT ends, while generators k-1..1 and the encoders are still being differentiated): it runs on NCCL's stream
# concurrently with the rest of the backward pass, and ``allreduce_grads`` -- called right before the optimizer's Adam
# launch -- only waits for it.  Under CUDA-graph capture the same fork / join is recorded into the step's graph.
_known = {}        # id(optimizer) -> parameters that received a gradient in the previous step (what the hooks count)
_inflight = {}     # id(optimizer) -> pending work of an all-reduce launched from the hooks
_handles: List = []


def _avg_supported() -> bool:
    import torch.distributed as dist
    return dist.get_backend() == "nccl"


@torch.no_grad()
def _launch(grads, async_op: bool):
    """Average `grads` in place over all ranks with one coalesced collective.  -> work handle (async) or None."""
    import torch.distributed as dist
    if not grads:
        return None
    avg = _avg_supported()
    op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
    dev = grads[0].device if grads[0].is_cuda else None
    with dist._coalescing_manager(device=dev, async_ops=async_op) as cm:
        for g in grads:
            dist.all_reduce(g, op=op)
    if not avg:      # gloo (CPU tests): no AVG reduction
        if async_op:
            cm.wait()
        w = float(_state["world"])
        for g in grads:
            g.div_(w)
        return None
    return cm if async_op else None


def begin_backward(optimizers):
    """Arm the overlap for one backward pass: install per-parameter hooks that launch an optimizer's all-reduce as soon
    as all of its gradients exist.  No-op on one GPU and before an optimizer's gradient set is known (first step)."""
    if not _state["enabled"]:
        return
    for opt in optimizers:
        params = _known.get(id(opt))
        if not params:
            continue
        st = {"left": len(params), "opt": opt, "params": params}

        def hook(_p, st=st):
            st["left"] -= 1
            if st["left"] == 0:
                from . import ops
                ops.join_side_streams()   # gradients of one optimizer may have been produced on several streams
                _inflight[id(st["opt"])] = (_launch([q.grad for q in st["params"]], True), st["params"])
        for p in params:
            _handles.append(p.register_post_accumulate_grad_hook(hook))


def end_backward():
    for h in _handles:
        h.remove()
    _handles.clear()


@torch.no_grad()
def allreduce_grads(optimizer: torch.optim.Optimizer):
    """Make ``optimizer``'s gradients the average over all ranks: wait for the all-reduce launched during backward, or run
    it now (first step, discriminator step, parameters whose gradient set changed)."""
    if not _state["enabled"]:
        return
    params = [p for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
    pending = _inflight.pop(id(optimizer), None)
    if pending is not None:
        work, launched = pending
        if work is not None:
            work.wait()
        if len(launched) != len(params) or any(a is not b for a, b in zip(launched, params)):
            raise RuntimeError("the set of parameters receiving gradients changed while an all-reduce was in flight")
    else:
        _launch([p.grad for p in params], False)
    _known[id(optimizer)] = params
