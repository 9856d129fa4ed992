"""Unit parity of the audio-encoder normalisation / squeeze-excite kernels (csrc/bn.cu, csrc/audio.cu) against fp64
torch on the same inputs: forward, input gradient and parameter gradients at 1e-5 (ResNetBlocks.py:21-37,81-96;
hierarchy_net.py:205-210), and run-to-run bit-reproducibility of the ordered reductions."""
import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("cfg", [  # rows, C, pre_relu, post_act
    (4 * 16 * 9, 256, True, 0), (3 * 128 * 70, 32, True, 0), (2 * 64 * 35, 64, False, 0), (5 * 31, 16, False, 2),
    (7 * 33, 8, False, 2), (3 * 17, 24, True, 0), (1000, 150, False, 0)])
def test_bn_fwd_bwd_vs_fp64(cfg):
    from ha2g_b200 import ops
    rows, C, pre_relu, post_act = cfg
    torch.manual_seed(rows + C)
    x = torch.randn(rows, C) * 1.5 + 0.3
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    rm, rv = 0.1 * torch.randn(C), 1 + 0.2 * torch.rand(C)
    g = torch.randn(rows, C)
    xd, gd, bd = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rmd, rvd = rm.double().clone(), rv.double().clone()
    h = torch.relu(xd) if pre_relu else xd
    y = torch.nn.functional.batch_norm(h, rmd, rvd, gd, bd, True, 0.1, 1e-5)
    if post_act == 2:
        y = torch.nn.functional.leaky_relu(y, 0.01)
    (y * g.double()).sum().backward()

    xg = x.to(DEV).requires_grad_(True)
    gg, bg = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    rmg, rvg = rm.to(DEV), rv.to(DEV)
    yg = ops.batch_norm(xg, gg, bg, rmg, rvg, pre_relu=pre_relu, post_act=post_act, training=True)
    (yg * g.to(DEV)).sum().backward()
    assert_close(yg, y, f"bn fwd {cfg}", 1e-5)
    assert_close(rmg, rmd, f"bn running_mean {cfg}", 1e-5)
    assert_close(rvg, rvd, f"bn running_var {cfg}", 1e-5)
    assert_close(xg.grad, xd.grad, f"bn dx {cfg}", 1e-5)
    assert_close(gg.grad, gd.grad, f"bn dgamma {cfg}", 1e-5)
    assert_close(bg.grad, bd.grad, f"bn dbeta {cfg}", 1e-5)
    # ordered reductions: a second evaluation is bit-identical
    xg2 = x.to(DEV).requires_grad_(True)
    gg2, bg2 = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    yg2 = ops.batch_norm(xg2, gg2, bg2, rm.to(DEV), rv.to(DEV), pre_relu=pre_relu, post_act=post_act, training=True)
    (yg2 * g.to(DEV)).sum().backward()
    assert torch.equal(yg, yg2) and torch.equal(xg.grad, xg2.grad) and torch.equal(gg.grad, gg2.grad)
    # eval mode uses the running statistics
    ye = ops.batch_norm(x.to(DEV), gamma.to(DEV), beta.to(DEV), rm.to(DEV), rv.to(DEV), pre_relu=pre_relu,
                        post_act=post_act, training=False)
    he = torch.relu(x.double()) if pre_relu else x.double()
    yr = torch.nn.functional.batch_norm(he, rm.double(), rv.double(), gamma.double(), beta.double(), False, 0.1, 1e-5)
    if post_act == 2:
        yr = torch.nn.functional.leaky_relu(yr, 0.01)
    assert_close(ye, yr, f"bn eval {cfg}", 1e-5)


@pytest.mark.parametrize("cfg", [(3, 128 * 70, 32), (4, 64 * 35, 64), (2, 32 * 18, 128), (5, 16 * 9, 256), (2, 7 * 5, 24)])
def test_se_fwd_bwd_vs_fp64(cfg):
    """SE tail relu(u * sigmoid(W2 relu(W1 gap(u) + b1) + b2) + res) (ResNetBlocks.py:29-37,81-96)."""
    from ha2g_b200 import ops_audio
    N, HW, C = cfg
    R = max(C // 8, 1)
    torch.manual_seed(N * C)
    u, res = torch.randn(N, HW, 1, C), torch.randn(N, HW, 1, C)
    w1, b1 = torch.randn(R, C) / C ** 0.5, 0.1 * torch.randn(R)
    w2, b2 = torch.randn(C, R) / R ** 0.5, 0.1 * torch.randn(C)
    g = torch.randn(N, HW, 1, C)
    dd = [t.double().requires_grad_(True) for t in (u, res, w1, b1, w2, b2)]
    ud, rd, w1d, b1d, w2d, b2d = dd
    gap = ud.mean(dim=(1, 2))
    s = torch.sigmoid(torch.relu(gap @ w1d.t() + b1d) @ w2d.t() + b2d)
    out = torch.relu(ud * s[:, None, None, :] + rd)
    (out * g.double()).sum().backward()
    gg = [t.to(DEV).requires_grad_(True) for t in (u, res, w1, b1, w2, b2)]
    og = ops_audio.se_residual_relu(*gg)
    (og * g.to(DEV)).sum().backward()
    assert_close(og, out, f"se fwd {cfg}", 1e-5)
    for name, a, b in zip(("du", "dres", "dw1", "db1", "dw2", "db2"), gg, dd):
        assert_close(a.grad, b.grad, f"se {name} {cfg}", 2e-5)
    gg2 = [t.to(DEV).requires_grad_(True) for t in (u, res, w1, b1, w2, b2)]
    og2 = ops_audio.se_residual_relu(*gg2)
    (og2 * g.to(DEV)).sum().backward()
    assert torch.equal(og, og2) and all(torch.equal(a.grad, b.grad) for a, b in zip(gg, gg2))


def test_reductions_bit_reproducible():
    """Split-K GEMM, column sums and the embedding scatter-add give bit-identical results on repeated launches (they use
    per-split partial planes reduced in a fixed order instead of float atomics)."""
    from ha2g_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(4352, 600, device=DEV, requires_grad=True)
    w = (torch.randn(900, 600, device=DEV) / 24).requires_grad_(True)
    b = torch.zeros(900, device=DEV, requires_grad=True)
    g = torch.randn(4352, 900, device=DEV)
    outs = []
    for _ in range(3):
        for t in (x, w, b):
            t.grad = None
        (ops.linear(x, w, b) * g).sum().backward()
        outs.append((x.grad.clone(), w.grad.clone(), b.grad.clone()))
    for o in outs[1:]:
        assert all(torch.equal(a, c) for a, c in zip(outs[0], o))
    wd = (g.double().t() @ x.detach().double())
    assert_close(outs[0][1], wd, "split-K dW", 1e-5)
    assert_close(outs[0][2], g.double().sum(0), "col sum db", 1e-5)
    # embedding: heavy index duplication (PAD-dominated text)
    table = torch.randn(500, 300, device=DEV, requires_grad=True)
    idx = torch.randint(0, 500, (4352,), device=DEV)
    idx[torch.rand(4352, device=DEV) < 0.8] = 0
    ge = torch.randn(4352, 300, device=DEV)
    es = []
    for _ in range(3):
        table.grad = None
        (ops.embedding(table, idx) * ge).sum().backward()
        es.append(table.grad.clone())
    assert torch.equal(es[0], es[1]) and torch.equal(es[0], es[2])
    ref = torch.zeros(500, 300, dtype=torch.float64, device=DEV).index_add_(0, idx, ge.double())
    assert_close(es[0], ref, "embedding scatter-add", 1e-5)
