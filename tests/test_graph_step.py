"""The whole-step CUDA graph (ha2g_b200/graph_step.py) must be indistinguishable from the eager step:
same losses, same parameter updates, torch's optimizer state kept in agreement, lr changes honoured.

Trajectories of this step are chaotic at tiny batch (Adam's first updates are lr*sign(g)), so the comparison is a single
step from a synchronised state, on pre-Adam quantities (losses, gradients), which the deterministic kernels make exact."""
import copy

import pytest
import torch

from helpers import build_modules
from ha2g_b200 import graph_step, rng
from ha2g_b200.synthetic import make_batch

pytestmark = pytest.mark.gpu

SEEDS = {"gens": 20, "dis": 30, "audio": 31, "text": 32}
DEV = "cuda:0"


class World:
    def __init__(self, variant, B=4):
        from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
        from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
        self.variant, self.B = variant, B
        self.fn = train_iter_hierarchy if variant == "gesture" else train_iter_hierarchy_expressive
        self.args, self.gens, self.D, self.A, self.T = build_modules(variant, 60, 5, SEEDS, DEV)
        lr = self.args.learning_rate
        mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
        self.opts = [mk(g) for g in self.gens] + [mk(self.D, lr * self.args.discriminator_lr_weight), mk(self.A), mk(self.T)]
        g = torch.Generator().manual_seed(3)
        # 3 cascade passes x L generators draw reparameterisation noise each step: a fixed cycle of distinct draws
        # (the same in every step and in every world; stateless, hence safe to bake into a graph)
        self.noise = [torch.randn((B, 16), generator=g).to(DEV) for _ in range(3 * len(self.gens))]
        self.perm = torch.randperm(B, generator=g).to(DEV)
        self.calls = 0

    @property
    def mods(self):
        return self.gens + [self.D, self.A, self.T]

    def _randn(self, shape):
        self.calls += 1
        return self.noise[(self.calls - 1) % len(self.noise)]

    def step(self, i):
        b = {k: v.to(DEV) for k, v in make_batch(self.variant, self.B, 60, 5, seed=500 + i).items()}
        with rng.override(randn_fn=self._randn, randperm_fn=lambda n: self.perm, dropout=False, graph_safe=True):
            return self.fn(self.args, 11, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *self.mods, *self.opts)

    def params(self):
        return [p.detach().clone() for m in self.mods for p in m.parameters()]

    def steps(self):
        return {int(o.state[p]["step"]) for o in self.opts for grp in o.param_groups for p in grp["params"] if p in o.state}

    def load_from(self, other):
        for m, mo in zip(self.mods, other.mods):
            m.load_state_dict(copy.deepcopy(mo.state_dict()))
        for o, oo in zip(self.opts, other.opts):
            o.load_state_dict(copy.deepcopy(oo.state_dict()))
        self.calls = other.calls


@pytest.fixture(autouse=True)
def _fresh_graph_cache():
    graph_step.reset()
    graph_step.enable(True)
    yield
    graph_step.enable(True)
    graph_step.reset()


def _close_losses(a, b, tol, what):
    assert set(a) == set(b), what
    for k in a:
        assert abs(a[k] - b[k]) <= tol * max(1.0, abs(a[k])), (what, k, a[k], b[k])


def _grads(w):
    return [None if p.grad is None else p.grad.detach().clone() for m in w.mods for p in m.parameters()]


@pytest.mark.parametrize("variant", ["gesture", "expressive"])
def test_graph_matches_eager(variant):
    """Sharp and non-chaotic: ONE step from a synchronised state, graph replay vs eager, compared on the quantities the
    step computes BEFORE Adam's sign-like update amplifies anything -- the returned losses and every parameter's
    gradient.  Every reduction of the step is order-fixed (no float atomics: two-stage ordered reductions everywhere),
    so the replayed kernels must reproduce the eager ones; post-Adam parameters then differ at most by the rounding of
    the bias correction (computed on the host for the eager launch, on the device for the captured one)."""
    n = 4
    s0 = dict(graph_step.STATS)
    wg = World(variant)
    for i in range(n):
        wg.step(i)
    assert graph_step.STATS["captures"] - s0["captures"] == 1
    assert graph_step.STATS["replays"] - s0["replays"] == n - graph_step.WARMUP
    assert wg.steps() == {n}

    graph_step.enable(False)
    we = World(variant)
    we.load_from(wg)
    before = we.params()
    r_e = we.step(n)
    g_e = _grads(we)
    graph_step.enable(True)
    r_g = wg.step(n)
    g_g = _grads(wg)
    assert graph_step.STATS["replays"] - s0["replays"] == n - graph_step.WARMUP + 1
    assert wg.steps() == {n + 1} and we.steps() == {n + 1}

    _close_losses(r_e, r_g, 1e-6, "synchronised step")
    names = [f"m{mi}.{k}" for mi, m in enumerate(we.mods) for k, _ in m.named_parameters()]
    inexact = 0
    for name, a, b in zip(names, g_e, g_g):
        assert (a is None) == (b is None), name
        if a is None:
            continue
        if not torch.equal(a, b):
            inexact += 1
            scale = max(float(a.abs().max()), 1e-30)
            assert float((a - b).abs().max()) <= 1e-5 * scale, (name, float((a - b).abs().max()), scale)
    assert inexact == 0, f"{inexact} gradient tensors of the replayed step are not bit-identical to the eager step"
    lr = 5e-4
    for name, p0, pe, pg in zip(names, before, we.params(), wg.params()):
        assert float((pe - pg).abs().max()) <= 1e-3 * lr, (name, float((pe - pg).abs().max()))
    assert any(not torch.equal(p0, pg) for p0, pg in zip(before, wg.params()))


def test_eager_step_is_deterministic():
    """Two eager runs from identical state: bit-identical losses and gradients (tools/determinism_check.py as a test)."""
    graph_step.enable(False)
    outs = []
    for _ in range(2):
        w = World("gesture")
        rets = [w.step(i) for i in range(2)]
        outs.append((rets, _grads(w)))
    (ra, ga), (rb, gb) = outs
    assert ra == rb, (ra, rb)
    for a, b in zip(ga, gb):
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.equal(a, b)


def test_graph_honours_lr_change():
    # lr -> 0 before the 4th call (a replay): parameters must stop moving, losses must keep being computed
    w = World("gesture")
    snaps, rets = [], []
    for i in range(4):
        if i == 3:
            for o in w.opts:
                for grp in o.param_groups:
                    grp["lr"] = 0.0
        rets.append(w.step(i))
        snaps.append(w.params())
    assert graph_step.STATS["replays"] >= 2
    assert any(not torch.equal(a, b) for a, b in zip(snaps[1], snaps[2]))   # the first replay (lr > 0) moved them
    for a, b in zip(snaps[2], snaps[3]):
        assert torch.equal(a, b)
    assert w.steps() == {4} and all(v == v for v in rets[3].values())


def test_graph_and_eager_interleave():
    # an eager step between replays (e.g. an injected-randomness step) must not desynchronise the device-side Adam
    # step counter or leave p.grad pointing at stale buffers
    w = World("gesture")
    for i in range(3):
        w.step(i)
    graph_step.enable(False)
    w.step(3)
    graph_step.enable(True)
    r = w.step(4)
    assert w.steps() == {5} and all(v == v for v in r.values())
    we = World("gesture")
    graph_step.enable(False)
    we.load_from(w)
    r_e = we.step(5)
    graph_step.enable(True)
    r_g = w.step(5)
    _close_losses(r_e, r_g, 1e-3, "after interleaving")
