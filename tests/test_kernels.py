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


@pytest.mark.parametrize("variant", ["expressive", "gesture"])
def test_contrastive_n4352_vs_fp64(variant):
    """SoftmaxContrastiveLoss at the benchmark size N = B*34 = 4352 (train_hierarchy.py:54-68 / ..._expressive.py:108-121):
    streaming CUDA kernels vs the reference's materialised N x N x 32 formulation evaluated in fp64 (on the device: 4.8 GB).
    Text-side rows come from the real TextEncoderTCN on a PAD-dominated synthetic batch (many near-identical rows)."""
    from ha2g_b200 import ops_loss
    from ha2g_b200.constants import make_args
    from ha2g_b200.model.hierarchy_net import TextEncoderTCN
    from ha2g_b200.synthetic import det_fill, make_batch, make_embedding
    torch.manual_seed(1)
    B = 128
    args = make_args("expressive")
    T = det_fill(TextEncoderTCN(args, 60, 300, pre_trained_embedding=make_embedding(60, 300, 1).numpy(), dropout=0.3), 32).to(DEV).train(False)
    text = make_batch("expressive", B, 60, 5, seed=901)["in_text_padded"].to(DEV)
    with torch.no_grad():
        a0 = T(text).reshape(-1, 32).clone()
    b0 = torch.randn(B * 34, 32, device=DEV) * 0.5
    a, b = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    loss = ops_loss.contrastive(a, b, variant)
    loss.backward()
    ad, bd = a0.double().requires_grad_(True), b0.double().requires_grad_(True)
    an, bn = torch.nn.functional.normalize(ad, dim=1), torch.nn.functional.normalize(bd, dim=1)
    D = (an.unsqueeze(1) - bn.unsqueeze(0)).norm(dim=2)
    logits = torch.clamp(1.0 / (D + 1e-8), min=1e-8) if variant == "gesture" else 1.0 / D
    ref = torch.nn.functional.cross_entropy(logits, torch.arange(B * 34, device=DEV))
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    ea = float((a.grad.double() - ad.grad).norm() / ad.grad.norm())
    eb = float((b.grad.double() - bd.grad).norm() / bd.grad.norm())
    print(f"contrastive N=4352 {variant}: rel L2 error da {ea:.2e}, db {eb:.2e}")
    assert ea <= 1e-4 and eb <= 1e-4, (ea, eb)


def test_ride_along_rows_match_separate_passes():
    """ops.ride_along: a generator forward over [G; D; R] rows in one pass (autograd on the G rows only) reproduces three
    separate passes -- outputs of all three, and every gradient of the differentiated one (hierarchy_net.py:99-149 called
    three times per step by train_hierarchy_expressive.py:150-213, 252-310, 336-394)."""
    from ha2g_b200 import ops, rng
    from ha2g_b200.constants import make_args
    from ha2g_b200.model.hierarchy_net import Hierarchical_PoseGenerator
    from ha2g_b200.model.vocab import make_speaker_vocab
    from ha2g_b200.synthetic import det_fill, make_batch, make_embedding
    args = make_args("expressive")
    spk = make_speaker_vocab(5)
    emb = make_embedding(60, 300, 1).numpy()
    for B in (3, 40):
        g = det_fill(Hierarchical_PoseGenerator(args, 30, 60, 300, emb, z_obj=spk), 21).to(DEV)
        torch.manual_seed(B)
        text = make_batch("expressive", B, 60, 5, seed=7)["in_text_padded"].to(DEV)
        pre = [torch.randn(B, 34, 31, device=DEV) * 0.1 for _ in range(3)]
        aud = torch.randn(B, 34, 32, device=DEV, requires_grad=True)
        vids = [torch.randint(1, 6, (B,), device=DEV) for _ in range(3)]
        eps = [torch.randn(B, 16, device=DEV) for _ in range(3)]
        gout = torch.randn(B, 34, 30, device=DEV)
        with rng.override(dropout=False):
            # separate passes
            sep = []
            for k in range(3):
                ctx = torch.enable_grad() if k == 0 else torch.no_grad()
                with ctx:
                    sep.append(g(pre[k], text, aud if k == 0 else aud.detach(), vids[k], _eps=eps[k]))
            (sep[0][0] * gout).sum().backward()
            ref_grads = {n: p.grad.clone() for n, p in g.named_parameters() if p.grad is not None}
            ref_daud = aud.grad.clone()
            g.zero_grad(); aud.grad = None
            # one ride-along pass
            pk = lambda ts: ops.ride_pack(ts[0], ts[1:])
            with ops.ride_along(3):
                out, z, mu, lv = g(pk(pre), pk([text] * 3), pk([aud, aud, aud]), pk(vids), _eps=pk(eps))
            assert out.shape == (B, 34, 30) and z.shape == (B, 16)
            (out * gout).sum().backward()
        tails_o, tails_z = ops.ride_tails(out, 3), ops.ride_tails(z, 3)
        assert_close(out, sep[0][0], f"ride head out B={B}", 1e-5)
        for k in (1, 2):
            assert_close(tails_o[k - 1], sep[k][0], f"ride tail {k} out B={B}", 1e-5)
            assert_close(tails_z[k - 1], sep[k][1], f"ride tail {k} z B={B}", 1e-5)
        assert_close(aud.grad, ref_daud, f"ride d(audio) B={B}", 1e-5)
        for n, p in g.named_parameters():
            if n in ref_grads:
                assert_close(p.grad, ref_grads[n], f"ride grad {n} B={B}", 2e-5)
            else:
                assert p.grad is None
