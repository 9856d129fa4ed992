"""Differentiable ops of the HA2G step, each backed by hand-written sm_100a kernels (csrc/*.cu).

PyTorch is used here only as the device-memory allocator, stream owner and autograd tape; every
arithmetic operation is a launch through the C ABI (ha2g_b200._lib.lib).  There is no fallback:
CPU tensors or a missing library raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import torch

from ._lib import lib
from . import rng as _rng

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_ELU, ACT_SIGMOID, ACT_TANH, ACT_LRELU02, ACT_LRELU03 = range(8)

LAUNCHES = [0]  # number of C-ABI launcher calls (bench.py reports it as gpu_launches)


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("ha2g_b200 ops need CUDA tensors (no CPU fallback path exists)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"expected float32, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError("expected a contiguous tensor")


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------
# ride-along rows: no-grad companion passes batched INTO the differentiated pass
# ------------------------------------------------------------------------------------------------
# The training step runs the generator cascade three times on the same weights: once with gradient (G step) and twice
# detached (discriminator step, mismatched-speaker pass).  Instead of three cascades -- or one with-grad plus one 2B-row
# no-grad cascade -- the row-wise ops below can carry the detached rows ALONG with the differentiated ones: while
# ``ride_along(m)`` is active, every batch-leading tensor of B rows handed to them is the HEAD of a contiguous buffer of
# m*B rows; forward kernels run over all m*B rows (one launch, m x the rows per GEMM / GRU step), outputs are allocated
# m*B rows long and returned as their B-row head, and autograd only ever sees the heads: saved tensors, gate buffers and
# every backward kernel cover B rows.  ``ride_pack`` builds such buffers, ``ride_tails`` reads the companions' rows back.
_ride = {"mult": 1}


class ride_along:
    def __init__(self, mult: int):
        self.mult, self.prev = int(mult), 1

    def __enter__(self):
        self.prev, _ride["mult"] = _ride["mult"], self.mult
        return self

    def __exit__(self, *exc):
        _ride["mult"] = self.prev
        return False


def _fv(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The whole ride-along buffer behind a head view (identity when no companions ride along)."""
    m = _ride["mult"]
    if m == 1 or t is None:
        return t
    if not t.is_contiguous():
        raise RuntimeError("ride-along tensors must be contiguous heads of their buffers")
    return torch.as_strided(t, (t.shape[0] * m, *t.shape[1:]), t.stride(), t.storage_offset())   # bounds-checked


def _alloc(shape, device, dtype=torch.float32):
    """-> (full buffer with mult x the leading rows, its head view of `shape`)."""
    m = _ride["mult"]
    full = torch.empty((shape[0] * m, *shape[1:]), device=device, dtype=dtype)
    return full, (full if m == 1 else full[:shape[0]])


class _RidePackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, head, *tails):
        full = torch.cat([head, *tails], dim=0)
        return full[:head.shape[0]]

    @staticmethod
    def backward(ctx, g):
        return (g,) + (None,) * (len(ctx.needs_input_grad) - 1)


def ride_pack(head: torch.Tensor, tails: Sequence[torch.Tensor]) -> torch.Tensor:
    """[head; tails...] stacked along the batch axis in one buffer, returned as its head (gradient flows to `head` only)."""
    tails = [t.detach() for t in tails]
    if head.requires_grad and torch.is_grad_enabled():
        return _RidePackFn.apply(head, *tails)
    return torch.cat([head, *tails], dim=0)[:head.shape[0]]


def ride_tails(head: torch.Tensor, mult: int) -> List[torch.Tensor]:
    """The companions' rows of a ride-along result: mult-1 detached [B, ...] views."""
    B = head.shape[0]
    h = head.detach()
    full = torch.as_strided(h, (B * mult, *h.shape[1:]), h.stride(), h.storage_offset())
    return [full[k * B:(k + 1) * B] for k in range(1, mult)]


# ------------------------------------------------------------------------------------------------
# side streams ("lanes")
# ------------------------------------------------------------------------------------------------
# Independent sub-graphs of the step made of small, launch-bound kernels (the six generators' text encoders) run on side
# streams next to the main stream's work -- forward next to the audio encoder, backward (autograd replays a node on the
# stream of its forward) next to the latency-bound GRU recurrences of the other generators.  Every side stream owns a
# scratch arena (ha2g_set_workspace_lane): launchers pick the arena of the stream they are enqueued on.
_lanes: Dict[int, dict] = {}
_LANES_ON = os.environ.get("HA2G_SIDE_STREAMS", "2")


def side_streams(device) -> List["torch.cuda.Stream"]:
    """The side streams of `device` (created and registered with their arenas on first use); [] when disabled."""
    n = int(_LANES_ON)
    dev = torch.device(device)
    if n <= 0 or dev.type != "cuda":
        return []
    i = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _lanes.get(i)
    if st is None:
        _ensure_workspace()
        n = min(n, 2)
        mb = int(os.environ.get("HA2G_LANE_WORKSPACE_MB", "256"))
        st = _lanes[i] = {"streams": [torch.cuda.Stream(device=i) for _ in range(n)],
                          "arenas": [torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{i}") for _ in range(n)], "main": None}
        for k, (s, a) in enumerate(zip(st["streams"], st["arenas"])):
            lib.ha2g_set_workspace_lane(k + 1, ctypes.c_void_p(a.data_ptr()), ctypes.c_int64(a.numel()), ctypes.c_void_p(s.cuda_stream))
    return st.get("text", st["streams"])


def loss_stream(device) -> Optional["torch.cuda.Stream"]:
    """A third side stream for the contrastive losses, with an arena as large as the main one (lane 3); None when side
    streams are disabled."""
    if not side_streams(device) or os.environ.get("HA2G_LOSS_STREAM", "1") == "0":
        return None
    dev = torch.device(device)
    i = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _lanes[i]
    if "loss" not in st:
        s_ = torch.cuda.Stream(device=i)
        a_ = torch.empty(_workspace[i].numel(), dtype=torch.uint8, device=f"cuda:{i}")
        lib.ha2g_set_workspace_lane(3, ctypes.c_void_p(a_.data_ptr()), ctypes.c_int64(a_.numel()), ctypes.c_void_p(s_.cuda_stream))
        st["loss"], st["loss_arena"] = s_, a_
        st["streams"] = st["streams"] + [s_]     # join_side_streams covers it
        st["text"] = st["streams"][:-1]
    return st["loss"]


def fork_side_streams(device) -> List["torch.cuda.Stream"]:
    """Side streams ordered after everything enqueued so far on the current (main) stream."""
    streams = side_streams(device)
    if streams:
        main = torch.cuda.current_stream(device)
        _lanes[main.device.index]["main"] = main
        for s in streams:
            s.wait_stream(main)
    return streams


def join_side_streams(device=None):
    """The current stream waits for the main stream and every side stream of its device (no-op without side streams):
    called before anything that consumes results produced on several of them at once (a collective over all gradients
    of an optimizer launched from a backward hook)."""
    if not _lanes:
        return
    cur = torch.cuda.current_stream(device)
    st = _lanes.get(cur.device.index)
    if st is None:
        return
    for s in ([st["main"]] if st["main"] is not None else []) + st["streams"]:
        if s != cur:
            cur.wait_stream(s)


_prof = {"on": False, "only": None, "flops_fn": None, "recs": {}}


_workspace = {}


def _ensure_workspace():
    """One torch-allocated scratch arena per device for the operand-packing passes of the tensor-core GEMMs and the
    partial planes of the ordered reductions (2 GB: the packed N_local x N_global coefficient matrix of the data-parallel
    contrastive loss is the largest tenant, 0.65 GB at 8 ranks)."""
    dev = torch.cuda.current_device()
    if dev not in _workspace:
        _workspace[dev] = torch.empty(int(os.environ.get("HA2G_WORKSPACE_MB", "2048")) << 20, dtype=torch.uint8, device=f"cuda:{dev}")
        lib.ha2g_set_workspace(_workspace[dev].data_ptr(), _workspace[dev].numel())
    return _workspace[dev]


def _call(name: str, *args):
    if not _workspace:
        _ensure_workspace()
    LAUNCHES[0] += 1
    if _prof["on"] and (_prof["only"] is None or _prof["only"] == name):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        getattr(lib, name)(*args)
        e1.record()
        fl = _prof["flops_fn"](name, args) if _prof["flops_fn"] is not None else 0.0
        _prof["recs"].setdefault(name, []).append((e0, e1, fl))
        return
    getattr(lib, name)(*args)


def profile_begin(all_launchers: bool = False, only: Optional[str] = None, flops_fn=None):
    """Time C-ABI launcher calls with CUDA events on the launching stream (bench.py's roofline leg)."""
    _prof.update(on=bool(all_launchers or only), only=None if all_launchers else only, flops_fn=flops_fn, recs={})


def profile_end():
    """-> {launcher: {"ms": total device time, "calls": n, "flops": algorithmic FLOPs}}"""
    torch.cuda.synchronize()
    out = {}
    for name, recs in _prof["recs"].items():
        out[name] = {"ms": sum(a.elapsed_time(b) for a, b, _ in recs), "calls": len(recs), "flops": sum(f for _, _, f in recs)}
    _prof.update(on=False, only=None, flops_fn=None, recs={})
    return out


# ------------------------------------------------------------------------------------------------
# raw helpers
# ------------------------------------------------------------------------------------------------
def gemm(A, B, C, bias, M, N, K, lda, ldb, ldc, tA=0, tB=0, act=0, acc=0, split=1):
    _call("ha2g_gemm", _p(A), _p(B), _p(C), _p(bias), M, N, K, lda, ldb, ldc, tA, tB, act, acc, split, _st())


_config = {"gemm_impl": "auto", "precision": "fp32x3"}


def set_gemm_impl(impl: str):
    """'auto' (default): packed tcgen05 bf16x3 GEMM for problems big enough to amortise packing, exact fp32 SIMT for
    the small ones; 'f32': SIMT only; 'tc2': packed kernel everywhere."""
    lib.ha2g_set_gemm_impl({"f32": 0, "auto": 1, "tc2": 3}[impl])
    _config["gemm_impl"] = impl


def set_precision(mode: str):
    """'fp32x3' (default): three-term bf16 split, fp32-accurate; 'bf16': plain bf16 operands on the tensor cores."""
    lib.ha2g_set_gemm_terms(1 if mode == "bf16" else 3)
    _config["precision"] = mode


def config_signature() -> tuple:
    """Kernel-selection switches that get baked into a captured CUDA graph (graph_step keys its cache on them)."""
    return (_config["gemm_impl"], _config["precision"])


def profiling() -> bool:
    return bool(_prof["on"])


def col_sum_into(x2d: torch.Tensor, out: torch.Tensor):
    rows, cols = x2d.shape
    _call("ha2g_col_sum", _p(x2d), rows, cols, x2d.stride(0), _p(out), _st())


def _split_for(rows: int, m: int, n: int) -> int:
    """split-K factor for weight-gradient GEMMs: enough CTAs to cover the 148 SMs."""
    tiles = ((m + 127) // 128) * ((n + 63) // 64)
    s = max(1, min(32, (148 + tiles - 1) // tiles))
    while s > 1 and rows // s < 64:
        s //= 2
    return max(1, s)


# ------------------------------------------------------------------------------------------------
# Linear (+ fused activation)
# ------------------------------------------------------------------------------------------------
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        x2 = _c(x.reshape(-1, x.shape[-1]))
        _chk(x2, w, b)
        R, K = x2.shape
        N = w.shape[0]
        yf, y = _alloc((R, N), x.device)
        gemm(_fv(x2), w, yf, b, yf.shape[0], N, K, K, K, N, 0, 1, act, 0, 1)
        ctx.act = act
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        ctx.save_for_backward(x2, w, y if act else None)
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        R, K = x2.shape
        N = w.shape[0]
        g = _c(dy.reshape(R, N))
        if ctx.act:
            g2 = torch.empty_like(g)
            _call("ha2g_act_bwd", _p(g), _p(y), _p(g2), g.numel(), ctx.act, _st())
            g = g2
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((R, K), device=g.device, dtype=torch.float32)
            gemm(g, w, dx, None, R, K, N, N, K, K, 0, 0, 0, 0, 1)
            dx = dx.reshape(ctx.xshape)
        if ctx.needs_input_grad[1]:
            dw = zeros_like(w)
            gemm(g, x2, dw, None, N, K, R, N, K, K, 1, 0, 0, 1, _split_for(R, N, K))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = zeros((N,), g.device)
            col_sum_into(g, db)
        return dx, dw, db, None


def linear(x, w, b=None, act=ACT_NONE):
    return _LinearFn.apply(x, w, b, act)


# ------------------------------------------------------------------------------------------------
# element-wise
# ------------------------------------------------------------------------------------------------
class _ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        x = _c(x)
        _chk(x)
        yf, y = _alloc(x.shape, x.device)
        _call("ha2g_act_fwd", _p(x), _p(yf), yf.numel(), act, _st())
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(dy)
        _call("ha2g_act_bwd", _p(dy), _p(y), _p(dx), dy.numel(), ctx.act, _st())
        return dx, None


def act(x, code):
    return _ActFn.apply(x, code)


class _AddActFn(torch.autograd.Function):
    """y = act(a + b)"""

    @staticmethod
    def forward(ctx, a, b, act):
        a, b = _c(a), _c(b)
        _chk(a, b)
        _fv(a), _fv(b)   # (bounds check: both are heads of ride-along buffers)
        yf, y = _alloc(a.shape, a.device)
        _call("ha2g_add_act_fwd", _p(a), _p(b), _p(yf), yf.numel(), act, _st())
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        if ctx.act:
            dx = torch.empty_like(dy)
            _call("ha2g_act_bwd", _p(dy), _p(y), _p(dx), dy.numel(), ctx.act, _st())
        else:
            dx = dy
        return dx, dx, None


def add_act(a, b, code=ACT_NONE):
    return _AddActFn.apply(a, b, code)


class _MulMaskFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask, scale):
        x = _c(x)
        _chk(x, mask)
        y = torch.empty_like(x)
        _call("ha2g_mul_mask", _p(x), _p(mask), scale, _p(y), x.numel(), _st())
        ctx.scale = scale
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(dy)
        _call("ha2g_mul_mask", _p(dy), _p(mask), ctx.scale, _p(dx), dy.numel(), _st())
        return dx, None, None


class _DropoutFn(torch.autograd.Function):
    """Fused Philox dropout: the mask is a function of (device seed/step, call id, element), regenerated in backward."""

    @staticmethod
    def forward(ctx, x, p, call_id):
        x = _c(x)
        _chk(x)
        _fv(x)
        yf, y = _alloc(x.shape, x.device)
        # the mask is indexed by element: the head's elements come first, so backward regenerates exactly their mask
        _call("ha2g_dropout", _p(x), _p(yf), yf.numel(), p, _p(_rng.dropout_state(x.device)), call_id, _st())
        ctx.cfg = (p, call_id)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, call_id = ctx.cfg
        dy = _c(dy)
        dx = torch.empty_like(dy)
        _call("ha2g_dropout", _p(dy), _p(dx), dy.numel(), p, _p(_rng.dropout_state(dy.device)), call_id, _st())
        return dx, None, None


def dropout(x, p: float, training: bool):
    """nn.Dropout semantics.  Default: one fused Philox kernel per direction (no mask tensor); a test-injected mask
    source (rng.override(mask_fn=...)) switches to explicit masks."""
    if not training or p <= 0.0 or not _rng.dropout_enabled():
        return x
    if _rng.fused_dropout():
        return _DropoutFn.apply(x, float(p), _rng.next_call_id())
    if _ride["mult"] != 1:
        raise RuntimeError("injected dropout masks cannot be combined with ride-along rows")
    mask = _rng.dropout_mask(x.shape, p, x.device)
    return _MulMaskFn.apply(x, mask, 1.0 / (1.0 - p))


class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, idx):
        _chk(table)
        idx = _c(idx)
        if idx.dtype != torch.int64 or not idx.is_cuda:
            raise RuntimeError("embedding indices must be CUDA int64")
        dim = table.shape[1]
        idxf = _fv(idx)
        outf, out = _alloc((*idx.shape, dim), table.device)
        _call("ha2g_embedding_fwd", _p(table), _p(idxf), _p(outf), idxf.numel(), dim, _st())
        ctx.save_for_backward(idx)
        ctx.tshape = table.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        dout = _c(dout)
        dt = zeros(ctx.tshape, dout.device)
        _call("ha2g_embedding_bwd", _p(dout), _p(idx), _p(_embedding_heads(idx)), _p(dt), idx.numel(), ctx.tshape[1], _st())
        return dt, None


_heads_cache = {}


def _embedding_heads(idx: torch.Tensor) -> torch.Tensor:
    """First-occurrence flags of an index tensor (the work list of the deterministic scatter-add).  All text encoders of
    a step see the same token batch, so the flags are computed once per (tensor, version) and shared; the cache is
    cleared at every step start (rng.begin_step)."""
    key = (idx.data_ptr(), idx._version, idx.numel(), idx.device)
    hit = _heads_cache.get(key)
    if hit is None:
        head = torch.empty((idx.numel(),), dtype=torch.uint8, device=idx.device)
        _call("ha2g_embedding_heads", _p(idx), idx.numel(), _p(head), _st())
        hit = _heads_cache[key] = (head, idx)   # holding idx keeps its address from being recycled under the key
    return hit[0]


def prime_embedding_heads(idx: torch.Tensor):
    """Compute (and cache) the first-occurrence flags of `idx` now, on the current stream: backward passes on several
    side streams then find them ready instead of racing to build them."""
    _embedding_heads(_c(idx))


def clear_step_caches():
    _heads_cache.clear()


# ------------------------------------------------------------------------------------------------
# zero-initialised gradient buffers of a step
# ------------------------------------------------------------------------------------------------
# The launchers ACCUMULATE weight / bias gradients into buffers their caller has zeroed.  A training step needs ~1 400 of
# them (most a few hundred bytes); as 1 400 fill kernels they cost ~3 ms per step.  Inside a step (begin_step .. end_step)
# they are carved out of one slab per device that a single memset clears at the step start; outside a step, and for
# whatever does not fit the slab, ``zeros`` is torch.zeros.
_zslab: Dict[int, dict] = {}
_ZSLAB_ALIGN = 256


def begin_step(device):
    """Step start (rng.begin_step; inside the captured graph too): per-step caches dropped, the slab of zero-initialised
    buffers cleared up to the extent the previous steps used."""
    clear_step_caches()
    dev = torch.device(device)
    if dev.type != "cuda":
        return
    i = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _zslab.get(i)
    if st is None:
        cap = int(os.environ.get("HA2G_ZERO_SLAB_MB", "1024")) << 20
        st = _zslab[i] = {"buf": torch.empty(cap, dtype=torch.uint8, device=f"cuda:{i}") if cap > 0 else None, "cap": cap,
                          "off": 0, "hw": 0, "clean": 0, "active": False}
    st["off"] = 0
    st["clean"] = st["hw"]
    st["active"] = st["buf"] is not None
    if st["clean"] > 0:
        st["buf"][:st["clean"]].zero_()


def end_step(device=None):
    for st in _zslab.values():
        st["active"] = False


def zeros(shape, device, dtype=torch.float32):
    """torch.zeros for a buffer that lives at most until the next step start (a gradient on its way to Adam)."""
    dev = torch.device(device)
    st = _zslab.get(dev.index if dev.index is not None else torch.cuda.current_device()) if dev.type == "cuda" else None
    if st is None or not st["active"] or dtype != torch.float32:
        return torch.zeros(shape, device=device, dtype=dtype)
    n = 1
    for d in shape:
        n *= int(d)
    nbytes = (4 * n + _ZSLAB_ALIGN - 1) // _ZSLAB_ALIGN * _ZSLAB_ALIGN
    off = st["off"]
    if n == 0 or off + nbytes > st["cap"]:
        return torch.zeros(shape, device=device, dtype=dtype)
    t = st["buf"][off:off + 4 * n].view(torch.float32).view(tuple(int(d) for d in shape))
    st["off"] = off + nbytes
    st["hw"] = max(st["hw"], st["off"])
    if st["off"] > st["clean"]:   # first steps (and growth): beyond what the step-start memset covered
        t.zero_()
    return t


def zeros_like(x):
    return zeros(tuple(x.shape), x.device, x.dtype)


def embedding(table, idx):
    return _EmbeddingFn.apply(table, idx)


class _ReparamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, eps):
        mu, logvar, eps = _c(mu), _c(logvar), _c(eps)
        _chk(mu, logvar, eps)
        _fv(mu), _fv(logvar), _fv(eps)
        zf, z = _alloc(mu.shape, mu.device)
        _call("ha2g_reparam_fwd", _p(mu), _p(logvar), _p(eps), _p(zf), zf.numel(), _st())
        ctx.save_for_backward(logvar, eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        logvar, eps = ctx.saved_tensors
        dz = _c(dz)
        dmu = torch.empty_like(dz)
        dlv = torch.empty_like(dz)
        _call("ha2g_reparam_bwd", _p(dz), _p(logvar), _p(eps), _p(dmu), _p(dlv), dz.numel(), _st())
        return dmu, dlv, None


def reparameterize(mu, logvar, eps=None):
    """embedding_net.py:10-13; the noise draw comes from ha2g_b200.rng unless the caller already drew it (the step
    pre-draws in the reference's order when it batches two cascade passes into one)."""
    if eps is None:
        eps = _rng.randn(mu.shape, mu.device)
    return _ReparamFn.apply(mu, logvar, eps)


class _ConcatSeqFn(torch.autograd.Function):
    """cat((pre_seq, audio, text, z.repeat over T), dim=2)   (hierarchy_net.py:129,140-142)."""

    @staticmethod
    def forward(ctx, pre, audio, text, z):
        pre, audio, text, z = _c(pre), _c(audio), _c(text), _c(z)
        _chk(pre, audio, text, z)
        B, T, dp = pre.shape
        widths = [dp, audio.shape[2], text.shape[2], z.shape[1]]
        I = sum(widths)
        _fv(pre), _fv(audio), _fv(text), _fv(z)
        xf, x = _alloc((B, T, I), pre.device)
        off = 0
        for src, w, div in ((pre, widths[0], 1), (audio, widths[1], 1), (text, widths[2], 1), (z, widths[3], T)):
            _call("ha2g_copy_cols", _p(src), w, 0, div, _p(xf), I, off, 1, xf.shape[0] * T, w, 0, _st())
            off += w
        ctx.widths, ctx.BT = widths, (B, T)
        return x

    @staticmethod
    def backward(ctx, dx):
        dx = _c(dx)
        B, T = ctx.BT
        I = dx.shape[2]
        outs, off = [], 0
        for i, w in enumerate(ctx.widths):
            if not ctx.needs_input_grad[i]:
                outs.append(None)
            elif i < 3:
                g = torch.empty((B, T, w), device=dx.device, dtype=torch.float32)
                _call("ha2g_copy_cols", _p(dx), I, off, 1, _p(g), w, 0, 1, B * T, w, 0, _st())
                outs.append(g)
            else:
                g = zeros((B, w), dx.device)
                _call("ha2g_copy_cols", _p(dx), I, off, 1, _p(g), w, 0, T, B * T, w, 2, _st())
                outs.append(g)
            off += w
        return tuple(outs)


def concat_seq(pre, audio, text, z):
    return _ConcatSeqFn.apply(pre, audio, text, z)


# ------------------------------------------------------------------------------------------------
# bidirectional multi-layer GRU
# ------------------------------------------------------------------------------------------------
class _BiGRUFn(torch.autograd.Function):
    """nn.GRU(batch_first, bidirectional, num_layers=L, dropout=p) with h0 = 0.

    weights: per layer [w_ih, w_hh, b_ih, b_hh, w_ih_rev, w_hh_rev, b_ih_rev, b_hh_rev]
    (nn.GRU._flat_weights order).  sum_dirs: return y[..., :H] + y[..., H:]  (hierarchy_net.py:145).
    """

    @staticmethod
    def forward(ctx, x, H, L, p, training, sum_dirs, *weights):
        x = _c(x)
        _chk(x, *weights)
        M, T, _ = x.shape
        _fv(x)
        MF = M * _ride["mult"]   # rows the forward kernels process: the head's M rows + the ride-along companions
        need_grad = any(ctx.needs_input_grad)  # (grad mode is off inside Function.forward)
        saved: List[torch.Tensor] = []
        masks = []
        cur = x
        wcats = []
        for l in range(L):
            w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r = weights[8 * l:8 * l + 8]
            I = cur.shape[2]
            # both directions' input weights / biases stacked ([6H, I] / [6H]): the launcher then runs ONE projection GEMM
            # with N = 6H (x packed once) and, in backward, ONE dx GEMM over K = 6H
            wcat, bcat = torch.cat([w_ih, w_ih_r], 0), torch.cat([b_ih, b_ih_r], 0)
            wcats.append(wcat)
            w_ih, w_ih_r, b_ih, b_ih_r = wcat[:3 * H], wcat[3 * H:], bcat[:3 * H], bcat[3 * H:]
            gi = torch.empty((MF, T, 6 * H), device=x.device, dtype=torch.float32)
            yf, y = _alloc((M, T, 2 * H), x.device)
            # the gates are saved for the head rows only: the companions are never differentiated
            gates = torch.empty((M, T, 8 * H), device=x.device, dtype=torch.float32) if need_grad else None
            _call("ha2g_gru_layer_fwd", _p(cur), I, _p(w_ih), _p(w_ih_r), _p(b_ih), _p(b_ih_r), _p(w_hh), _p(w_hh_r),
                  _p(b_hh), _p(b_hh_r), _p(gi), _p(yf), _p(gates), MF, M, T, H, _st())
            if need_grad:
                saved += [cur, y, gates]
            nxt = y
            mask = None
            if l < L - 1 and training and p > 0.0 and _rng.dropout_enabled():
                nf, nxt = _alloc((M, T, 2 * H), x.device)
                if _rng.fused_dropout():
                    mask = _rng.next_call_id()   # the mask is regenerated from this id in backward
                    _call("ha2g_dropout", _p(yf), _p(nf), yf.numel(), float(p), _p(_rng.dropout_state(y.device)), mask, _st())
                else:
                    if MF != M:
                        raise RuntimeError("injected dropout masks cannot be combined with ride-along rows")
                    mask = _rng.dropout_mask(y.shape, p, y.device)
                    _call("ha2g_mul_mask", _p(y), _p(mask), 1.0 / (1.0 - p), _p(nxt), y.numel(), _st())
            masks.append(mask)
            cur = nxt
        if sum_dirs:
            of, out = _alloc((M, T, H), x.device)
            _call("ha2g_copy_cols", _p(cur), 2 * H, 0, 1, _p(of), H, 0, 1, MF * T, H, 0, _st())
            _call("ha2g_copy_cols", _p(cur), 2 * H, H, 1, _p(of), H, 0, 1, MF * T, H, 1, _st())
        else:
            out = cur
        ctx.cfg = (H, L, p, sum_dirs, M, T)
        ctx.masks = masks
        ctx.nw = len(weights)
        ctx.wcats = wcats if need_grad else None
        ctx.save_for_backward(*saved, *weights)
        return out

    @staticmethod
    def backward(ctx, dout):
        H, L, p, sum_dirs, M, T = ctx.cfg
        st = ctx.saved_tensors
        weights = st[len(st) - ctx.nw:]
        saved = st[:len(st) - ctx.nw]
        dev = dout.device
        dy = _c(dout)
        dy_ld, dy_ds = (H, 0) if sum_dirs else (2 * H, H)
        grads: List[Optional[torch.Tensor]] = [None] * ctx.nw
        dgi = torch.empty((M, T, 6 * H), device=dev, dtype=torch.float32)
        dgh = torch.empty((M, T, 6 * H), device=dev, dtype=torch.float32)
        dh = torch.empty((M, 2 * H), device=dev, dtype=torch.float32)
        dx = None
        for l in range(L - 1, -1, -1):
            xin, y, gates = saved[3 * l:3 * l + 3]
            w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r = weights[8 * l:8 * l + 8]
            w_ih, w_ih_r = ctx.wcats[l][:3 * H], ctx.wcats[l][3 * H:]     # the stacked copy made in forward (one dx GEMM)
            I = xin.shape[2]
            # both directions' gradients of one kind live in ONE buffer: the launcher then needs a single GEMM for
            # dW_ih (dgi^T x over all 6H gate columns) and a single column sum per bias pair, and 4 zero-fills, not 8
            gw_ih = zeros((2,) + tuple(w_ih.shape), dev)
            gw_hh = zeros((2,) + tuple(w_hh.shape), dev)
            gb_ih = zeros((2,) + tuple(b_ih.shape), dev)
            gb_hh = zeros((2,) + tuple(b_hh.shape), dev)
            g = [gw_ih[0], gw_hh[0], gb_ih[0], gb_hh[0], gw_ih[1], gw_hh[1], gb_ih[1], gb_hh[1]]
            need_dx = l > 0 or ctx.needs_input_grad[0]
            dx = torch.empty((M, T, I), device=dev, dtype=torch.float32) if need_dx else None
            _call("ha2g_gru_layer_bwd", _p(dy), dy_ld, dy_ds, _p(xin), I, _p(y), _p(gates), _p(w_ih), _p(w_ih_r),
                  _p(w_hh), _p(w_hh_r), _p(dgi), _p(dgh), _p(dh), _p(dx), _p(g[0]), _p(g[4]), _p(g[1]), _p(g[5]),
                  _p(g[2]), _p(g[6]), _p(g[3]), _p(g[7]), M, T, H, _st())
            grads[8 * l:8 * l + 8] = g
            if l > 0:
                mask = ctx.masks[l - 1]
                if isinstance(mask, int):
                    dy = torch.empty_like(dx)
                    _call("ha2g_dropout", _p(dx), _p(dy), dx.numel(), float(p), _p(_rng.dropout_state(dx.device)), mask, _st())
                elif mask is not None:
                    dy = torch.empty_like(dx)
                    _call("ha2g_mul_mask", _p(dx), _p(mask), 1.0 / (1.0 - p), _p(dy), dx.numel(), _st())
                else:
                    dy = dx
                dy_ld, dy_ds = 2 * H, H
        return (dx, None, None, None, None, None, *grads)


def bigru(x, weights: Sequence[torch.Tensor], H: int, L: int, p: float, training: bool, sum_dirs: bool = True):
    return _BiGRUFn.apply(x, H, L, p, training, sum_dirs, *weights)


# ------------------------------------------------------------------------------------------------
# TCN pieces
# ------------------------------------------------------------------------------------------------
class _TcnWeightFn(torch.autograd.Function):
    """weight_norm'd Conv1d weight (g [O,1,1], v [O,I,Kw]) -> packed [O, Kw*I] (tcn.py:19-24)."""

    @staticmethod
    def forward(ctx, g, v):
        _chk(g, v)
        O, I, Kw = v.shape
        w = torch.empty((O, Kw * I), device=v.device, dtype=torch.float32)
        norm = torch.empty((O,), device=v.device, dtype=torch.float32)
        _call("ha2g_tcn_weight_fwd", _p(g), _p(v), _p(w), _p(norm), O, I, Kw, _st())
        ctx.save_for_backward(g, v, norm)
        return w

    @staticmethod
    def backward(ctx, dw):
        g, v, norm = ctx.saved_tensors
        O, I, Kw = v.shape
        dw = _c(dw)
        dg = zeros_like(g)
        dv = zeros_like(v)
        _call("ha2g_tcn_weight_bwd", _p(dw), _p(g), _p(v), _p(norm), _p(dg), _p(dv), O, I, Kw, _st())
        return dg, dv


def tcn_weight(g, v):
    return _TcnWeightFn.apply(g, v)


class _ShiftConcatFn(torch.autograd.Function):
    """[x(t-d) | x(t)] staging of a causal dilated k=2 conv (tcn.py:19-21 + Chomp1d)."""

    @staticmethod
    def forward(ctx, x, d):
        x = _c(x)
        _chk(x)
        B, T, C = x.shape
        _fv(x)
        of, out = _alloc((B, T, 2 * C), x.device)
        _call("ha2g_shift_concat_fwd", _p(x), _p(of), of.shape[0], T, C, d, _st())
        ctx.d = d
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _c(dout)
        B, T, C2 = dout.shape
        dx = torch.empty((B, T, C2 // 2), device=dout.device, dtype=torch.float32)
        _call("ha2g_shift_concat_bwd", _p(dout), _p(dx), B, T, C2 // 2, ctx.d, _st())
        return dx, None


def shift_concat(x, d):
    return _ShiftConcatFn.apply(x, d)


# ------------------------------------------------------------------------------------------------
# cascade glue
# ------------------------------------------------------------------------------------------------
def gather_cols(src: torch.Tensor, idx_i32: torch.Tensor) -> torch.Tensor:
    """src[..., idx] (no gradient: used for the target_k bone subsets)."""
    src = _c(src)
    _chk(src)
    rows = src.numel() // src.shape[-1]
    n = idx_i32.numel()
    out = torch.empty((*src.shape[:-1], n), device=src.device, dtype=torch.float32)
    _call("ha2g_gather_cols", _p(src), src.shape[-1], _p(idx_i32), n, _p(out), n, rows, _st())
    return out


class _PreSeqFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target_k, prev_out, slot_src, src_slot, n_pre):
        target_k = _c(target_k)
        _chk(target_k, prev_out)
        B, T, d = target_k.shape
        _fv(target_k)
        pf, pre = _alloc((B, T, d + 1), target_k.device)
        dp = 0
        if prev_out is not None:
            prev_out = _c(prev_out)
            _fv(prev_out)
            dp = prev_out.shape[2]
        _call("ha2g_pre_seq_fwd", _p(target_k), _p(prev_out), dp, _p(slot_src), _p(pf), pf.shape[0], T, d, n_pre, _st())
        ctx.cfg = (B, T, d, dp, n_pre)
        ctx.src_slot = src_slot
        return pre

    @staticmethod
    def backward(ctx, dpre):
        B, T, d, dp, n_pre = ctx.cfg
        if dp == 0 or not ctx.needs_input_grad[1]:
            return None, None, None, None, None
        dpre = _c(dpre)
        dprev = torch.empty((B, T, dp), device=dpre.device, dtype=torch.float32)
        _call("ha2g_pre_seq_bwd", _p(dpre), d + 1, _p(ctx.src_slot), _p(dprev), B, T, dp, n_pre, _st())
        return None, dprev, None, None, None


def pre_seq(target_k, prev_out, slot_src, src_slot, n_pre):
    return _PreSeqFn.apply(target_k, prev_out, slot_src, src_slot, n_pre)


# ------------------------------------------------------------------------------------------------
# Conv1d (valid) as unfold + GEMM, BatchNorm over [rows, C]
# ------------------------------------------------------------------------------------------------
class _Unfold1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kw):
        x = _c(x)
        _chk(x)
        B, T, C = x.shape
        out = torch.empty((B, T - kw + 1, kw * C), device=x.device, dtype=torch.float32)
        _call("ha2g_unfold1d_fwd", _p(x), _p(out), B, T, C, kw, _st())
        ctx.cfg = (B, T, C, kw)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, T, C, kw = ctx.cfg
        dout = _c(dout)
        dx = torch.empty((B, T, C), device=dout.device, dtype=torch.float32)
        _call("ha2g_unfold1d_bwd", _p(dout), _p(dx), B, T, C, kw, _st())
        return dx, None


class _PackConv1dWFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w):
        _chk(w)
        O, I, Kw = w.shape
        wp = torch.empty((O, Kw * I), device=w.device, dtype=torch.float32)
        _call("ha2g_pack_conv1d_w", _p(w), _p(wp), O, I, Kw, 0, _st())
        ctx.shape = (O, I, Kw)
        return wp

    @staticmethod
    def backward(ctx, dwp):
        O, I, Kw = ctx.shape
        dwp = _c(dwp)
        dw = torch.empty((O, I, Kw), device=dwp.device, dtype=torch.float32)
        _call("ha2g_pack_conv1d_w", _p(dwp), _p(dw), O, I, Kw, 1, _st())
        return dw


class _Unfold1dStridedFn(torch.autograd.Function):
    """im2col of a strided, zero-padded Conv1d on a channels-last sequence: x [B,T,C] -> [B,To,Kw*C]."""

    @staticmethod
    def forward(ctx, x, kw, stride, pad):
        x = _c(x)
        _chk(x)
        B, T, C = x.shape
        To = (T + 2 * pad - kw) // stride + 1
        out = torch.empty((B, To, kw * C), device=x.device, dtype=torch.float32)
        _call("ha2g_unfold1d_strided_fwd", _p(x), _p(out), B, T, C, kw, stride, pad, To, _st())
        ctx.cfg = (B, T, C, kw, stride, pad, To)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, T, C, kw, stride, pad, To = ctx.cfg
        dout = _c(dout)
        dx = torch.empty((B, T, C), device=dout.device, dtype=torch.float32)
        _call("ha2g_unfold1d_strided_bwd", _p(dout), _p(dx), B, T, C, kw, stride, pad, To, _st())
        return dx, None, None, None


def conv1d(x, w, b, stride=1, pad=0, act=ACT_NONE):
    """nn.Conv1d(C_in, C_out, Kw, stride, padding) on a [B,T,C_in] (channels-last) sequence -> [B,To,C_out]
    (the baseline WavEncoder's k = 15, stride 5 / 6 convolutions, multimodal_context_net.py:13-22)."""
    kw = w.shape[2]
    if stride == 1 and pad == 0:
        return conv1d_valid(x, w, b, act)
    return linear(_Unfold1dStridedFn.apply(x, kw, stride, pad), _PackConv1dWFn.apply(w), b, act)


def conv1d_valid(x, w, b, act=ACT_NONE):
    """nn.Conv1d(C_in, C_out, Kw) on a [B,T,C_in] (channels-last) sequence -> [B,T-Kw+1,C_out]."""
    kw = w.shape[2]
    return linear(_Unfold1dFn.apply(x, kw), _PackConv1dWFn.apply(w), b, act)


class _BNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, rm, rv, pre_relu, post_act, training, eps, momentum):
        x = _c(x)
        _chk(x, gamma, beta, rm, rv)
        C = x.shape[-1]
        rows = x.numel() // C
        dev = x.device
        mean = torch.empty((C,), device=dev, dtype=torch.float32)
        invstd = torch.empty((C,), device=dev, dtype=torch.float32)
        sums = torch.empty((2 * C,), device=dev, dtype=torch.float64)
        y = torch.empty_like(x)
        _call("ha2g_bn_fwd", _p(x), rows, C, int(pre_relu), post_act, int(training), _p(gamma), _p(beta), _p(rm), _p(rv),
              eps, momentum, _p(mean), _p(invstd), _p(sums), _p(y), _st())
        ctx.cfg = (rows, C, int(pre_relu), post_act, training)
        ctx.save_for_backward(x, y if post_act else None, gamma, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        rows, C, pre_relu, post_act, training = ctx.cfg
        if not training:
            raise RuntimeError("BatchNorm backward is only implemented for training mode (batch statistics)")
        x, y, gamma, mean, invstd = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(x)
        dg = zeros_like(gamma)
        db = zeros_like(gamma)
        sums = torch.empty((2 * C,), device=dy.device, dtype=torch.float64)
        _call("ha2g_bn_bwd", _p(dy), _p(x), _p(y), rows, C, pre_relu, post_act, _p(gamma), _p(mean), _p(invstd), _p(sums),
              _p(dx), _p(dg), _p(db), _st())
        return dx, dg, db, None, None, None, None, None, None, None


def batch_norm(x, gamma, beta, running_mean, running_var, pre_relu=False, post_act=ACT_NONE, training=True,
               eps=1e-5, momentum=0.1):
    """BatchNorm over the last (channel) axis of a channels-last tensor; see csrc/bn.cu."""
    return _BNFn.apply(x, gamma, beta, running_mean, running_var, pre_relu, post_act, training, eps, momentum)
