"""Whole-step CUDA graph for the HA2G training step.

One step of ``train_iter_hierarchy_expressive`` is ~7 600 kernel launches (2 400 C-ABI launcher calls plus the
allocator/fill/RNG plumbing) whose host-side enqueue alone takes ~137 ms at B = 128 -- most of the 188 ms step
(`tools/host_vs_device.py`).  The step has no data-dependent control flow and reads its scalars back once at the
end, so after ``WARMUP`` eager calls for a given signature we capture everything between "inputs are on the device"
and "packed scalars are ready" into a single ``torch.cuda.CUDAGraph`` and replay it:

    static inputs <- copy of the caller's four tensors (device or pinned host, asynchronous)
    replay        (forward, both backward passes, NCCL gradient all-reduce, the eight Adam launches)
    one packed device->host read of the step's scalars

What makes the step capturable:
  * every random draw is device-side and stream-ordered (torch's graph-aware Philox generator for randn / dropout
    keys; ``ha2g_rank_perm`` instead of ``torch.randperm``);
  * Adam's step counter and hyper-parameters live in device memory (``ha2g_adam_multi_dev``), the host only keeps
    torch's own ``optimizer.state[p]['step']`` in agreement and re-uploads ``lr`` when a schedule changes it;
  * the launchers allocate nothing and never synchronise; scratch comes from one arena set before capture.

The graph is keyed by everything baked into it: module / optimizer identities, tensor shapes, train/eval flags,
the loss switches derived from ``args`` and ``epoch``.  Steps with injected randomness (parity tests) and profiled
steps stay eager.  Set ``HA2G_CUDA_GRAPH=0`` to disable.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import dp, optim, rng
from . import ops as _ops

WARMUP = 2          # eager calls per signature before capture (optimizer state, BN buffers, launcher caches)
MAX_GRAPHS = 4      # captured signatures kept alive (each holds one step's activations in its private pool)

_state = {"enabled": os.environ.get("HA2G_CUDA_GRAPH", "1") != "0"}
_graphs: Dict[tuple, dict] = {}
_seen: Dict[tuple, int] = {}
_failed: set = set()
STATS = {"captures": 0, "replays": 0}


def enable(flag: bool = True):
    _state["enabled"] = bool(flag)


def enabled() -> bool:
    return _state["enabled"]


def reset():
    """Drop every captured graph (frees their private memory pools)."""
    _graphs.clear()
    _seen.clear()
    _failed.clear()


_ARG_KEYS = ("loss_warmup", "n_pre_poses", "loss_gan_weight", "loss_contrastive_pos_weight", "loss_contrastive_neg_weight",
             "loss_regression_weight", "loss_reg_weight", "loss_kld_weight", "loss_physical_weight", "z_type")


def _signature(world):
    (variant, args, epoch, in_text, in_spec, target, vid, gens, D, A, T, gopts, dopt, aopt, topt) = world
    mods = list(gens) + [D, A, T]
    gan_on = epoch > args.loss_warmup and args.loss_gan_weight > 0.0
    return (variant, bool(gan_on), bool(epoch > args.loss_warmup),
            tuple(repr(getattr(args, k, None)) for k in _ARG_KEYS), repr(getattr(args, "mean_dir_vec", None)),
            tuple(in_text.shape), tuple(in_spec.shape), tuple(target.shape), tuple(vid.shape), str(target.device),
            tuple(id(m) for m in mods), tuple(m.training for m in mods),
            tuple(id(o) for o in list(gopts) + [dopt, aopt, topt]), dp.world_size(), _ops.config_signature(),
            bool(rng.dropout_enabled()))


def _eligible(world) -> bool:
    if not _state["enabled"] or rng.overridden() or _ops.profiling():
        return False
    target = world[5]
    if not target.is_cuda or not torch.is_grad_enabled():
        return False
    return not torch.cuda.is_current_stream_capturing()


def run(enqueue, world) -> Optional[tuple]:
    """-> (names, values, flags) from a graph replay, or None when the caller should run the step eagerly."""
    if not _eligible(world):
        return None
    key = _signature(world)
    ent = _graphs.get(key)
    if ent is None:
        n = _seen.get(key, 0)
        if n < WARMUP or len(_graphs) >= MAX_GRAPHS or key in _failed:
            _seen[key] = n + 1
            return None
        try:
            ent = _capture(enqueue, world, key)
        except Exception:
            _failed.add(key)   # do not retry the capture on every later step; the caller runs this signature eagerly
            raise
    if not _still_valid(ent):
        # parameters or Adam moments were re-allocated (optimizer.load_state_dict, module.to(...)): the addresses baked
        # into the graph are stale -- drop it and warm up again
        del _graphs[key]
        _seen[key] = 1
        return None
    return _replay(ent, world)


def _still_valid(ent) -> bool:
    for o in ent["opts"]:
        for group, e in zip(o.param_groups, ent["adam"][id(o)]):
            if e is None:
                continue
            rows = e["rows"]
            for i, p in enumerate(e["params"]):
                st = o.state.get(p)
                if st is None or p.data_ptr() != rows[4 * i] or st["exp_avg"].data_ptr() != rows[4 * i + 2] or \
                        st["exp_avg_sq"].data_ptr() != rows[4 * i + 3]:
                    return False
    return True


def _capture(enqueue, world, key) -> dict:
    (variant, args, epoch, in_text, in_spec, target, vid, gens, D, A, T, gopts, dopt, aopt, topt) = world
    dev = target.device
    opts = list(gopts) + [dopt, aopt, topt]
    torch.cuda.synchronize(dev)
    static = {"in_text": torch.empty(in_text.shape, dtype=in_text.dtype, device=dev),
              "in_spec": torch.empty(in_spec.shape, dtype=in_spec.dtype, device=dev),
              "target": torch.empty(target.shape, dtype=target.dtype, device=dev),
              "vid": torch.empty(vid.shape, dtype=vid.dtype, device=dev)}
    for k, src in (("in_text", in_text), ("in_spec", in_spec), ("target", target), ("vid", vid)):
        static[k].copy_(src)
    adam_entries = {id(o): optim.graph_prepare(o) for o in opts}

    def adam(opt):
        optim.graph_adam_enqueue(opt, adam_entries[id(opt)])

    _ops._ensure_workspace()
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize(dev)
    n0 = _ops.LAUNCHES[0]
    with torch.cuda.graph(graph):
        names, packed, flags = enqueue(variant, args, epoch, static["in_text"], static["in_spec"], static["target"],
                                       static["vid"], gens, D, A, T, gopts, dopt, aopt, topt, adam=adam)
    for o in opts:
        optim.graph_finalize(adam_entries[id(o)])
    torch.cuda.synchronize(dev)
    ent = {"graph": graph, "static": static, "names": names, "packed": packed, "flags": flags, "opts": opts,
           "mods": list(gens) + [D, A, T],   # keeps the ids in the key from being recycled
           "adam": adam_entries, "launcher_calls": _ops.LAUNCHES[0] - n0}
    _graphs[key] = ent
    STATS["captures"] += 1
    return ent


def _replay(ent, world):
    (variant, args, epoch, in_text, in_spec, target, vid, *_rest) = world
    st = ent["static"]
    for k, src in (("in_text", in_text), ("in_spec", in_spec), ("target", target), ("vid", vid)):
        if src.data_ptr() != st[k].data_ptr():
            st[k].copy_(src, non_blocking=True)
    for o in ent["opts"]:
        optim.graph_pre_replay(o, ent["adam"][id(o)])
    ent["graph"].replay()
    _ops.LAUNCHES[0] += ent["launcher_calls"]   # the captured launcher calls execute once per replay
    STATS["replays"] += 1
    return ent["names"], ent["packed"].tolist(), ent["flags"]
