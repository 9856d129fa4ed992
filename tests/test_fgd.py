"""FGD evaluation path: the CUDA MotionAE + EmbeddingSpaceEvaluator against the fixture produced by the
UNMODIFIED reference classes (oracle/make_fgd_golden.py -> tests/golden/fgd.pt)."""
import argparse
import os
import types

import numpy as np
import pytest
import torch

from ha2g_b200.synthetic import _gen, det_fill
from helpers import assert_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fgd.pt")


def test_frechet_distance_closed_form():
    """Commuting (diagonal) covariances: d^2 = |mu1-mu2|^2 + sum (sqrt(a_i) - sqrt(b_i))^2."""
    from ha2g_b200.model.embedding_space_evaluator import EmbeddingSpaceEvaluator as E
    rs = np.random.RandomState(0)
    a, b = rs.rand(16) + 0.1, rs.rand(16) + 0.1
    m1, m2 = rs.randn(16), rs.randn(16)
    d = E.calculate_frechet_distance(m1, np.diag(a), m2, np.diag(b))
    assert abs(d - (np.sum((m1 - m2) ** 2) + np.sum((np.sqrt(a) - np.sqrt(b)) ** 2))) < 1e-9


def test_motion_ae_state_dict_matches_reference_layout():
    from ha2g_b200.model.motion_ae import MotionAE
    keys = list(MotionAE(126, 128).state_dict().keys())
    assert keys[0] == "encoder.net.0.0.weight" and "decoder.net.7.bias" in keys and len(keys) == 66


@pytest.mark.gpu
def test_evaluator_matches_reference():
    from ha2g_b200.model.embedding_space_evaluator import EmbeddingSpaceEvaluator
    from ha2g_b200.model.motion_ae import MotionAE
    g = torch.load(GOLD, weights_only=False)
    dev = torch.device("cuda:0")
    net = det_fill(MotionAE(126, 128), g["fill_seed"])
    ckpt = {"pose_dim": 126, "latent_dim": 128, "motion_ae": net.state_dict()}
    args = argparse.Namespace(n_pre_poses=4, n_poses=34, pose_dim=126, wordembed_dim=300)
    ev = EmbeddingSpaceEvaluator(args, None, types.SimpleNamespace(), dev, ckpt=ckpt)
    for i in range(g["n_batch"]):
        real = torch.randn((g["B"], 34, 126), generator=_gen(g["data_seed"], f"real{i}")) * 0.3
        gen = real + torch.randn((g["B"], 34, 126), generator=_gen(g["data_seed"], f"gen{i}")) * 0.1
        ev.push_samples(None, None, gen.to(dev), real.to(dev))
        recon, z = ev.net(real.to(dev))
        assert_close(z, g["batches"][i]["z_real"], f"MotionAE z batch {i}", 1e-4)
        assert_close(recon, g["batches"][i]["recon_real"], f"MotionAE reconstruction batch {i}", 1e-4)
    assert ev.get_no_of_samples() == g["n_batch"]
    for mine, ref, what in ((ev.recon_err_diff, g["recon_err_diff"], "recon_err_diff"), (ev.cos_err_diff, g["cos_err_diff"], "cos_err_diff")):
        for a, b in zip(mine, ref):
            assert abs(float(a) - b) <= 1e-3 * max(1.0, abs(b)), (what, float(a), b)
    frechet, feat_dist = ev.get_scores()
    assert abs(frechet - g["frechet"]) <= 1e-3 * max(1.0, abs(g["frechet"])), (frechet, g["frechet"])
    assert abs(feat_dist - g["feat_dist"]) <= 1e-4 * g["feat_dist"], (feat_dist, g["feat_dist"])
    assert ev.get_diversity_scores() > 0
