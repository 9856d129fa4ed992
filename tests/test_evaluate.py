"""Evaluation metrics: host-side part of evaluate_testset against the reference's own conversion
function (CPU; needs /root/reference for the cross-check, hand-computed cases otherwise) and the full loop on the GPU."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from ha2g_b200 import constants as K
from ha2g_b200 import evaluate

REF = "/root/reference/scripts"


def test_convert_dir_vec_to_pose_by_hand():
    vec = np.zeros((9, 3)); vec[0] = (0, 1, 0); vec[1] = (1, 0, 0); vec[3] = (0, 0, 1)
    p = evaluate.convert_dir_vec_to_pose(vec.reshape(-1), "gesture")
    assert p.shape == (10, 3)
    assert np.allclose(p[1], (0, 0.26, 0)) and np.allclose(p[2], (0.18, 0.26, 0)) and np.allclose(p[4], (0, 0.26, 0.22))
    assert len(K.EXPRESSIVE_DIR_VEC_PAIRS) == 42 and len(K.GESTURE_DIR_VEC_PAIRS) == 9
    # every joint except the root is the child of exactly one bone, parents come before children
    for pairs in (K.EXPRESSIVE_DIR_VEC_PAIRS, K.GESTURE_DIR_VEC_PAIRS):
        seen = {0}
        for parent, child, length in pairs:
            assert parent in seen and child not in seen and length > 0
            seen.add(child)
        assert seen == set(range(len(pairs) + 1))


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")
def test_pose_conversion_and_metrics_match_reference():
    for name in ("librosa", "fasttext"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    try:
        for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[m]
        import importlib
        ref_e = importlib.import_module("utils.data_utils_expressive")
        ref_g = importlib.import_module("utils.data_utils")
        assert [tuple(p) for p in ref_e.dir_vec_pairs] == K.EXPRESSIVE_DIR_VEC_PAIRS
        assert [tuple(p) for p in ref_g.dir_vec_pairs] == K.GESTURE_DIR_VEC_PAIRS
        rs = np.random.RandomState(0)
        for variant, ref, d in (("expressive", ref_e, 126), ("gesture", ref_g, 27)):
            out, tgt = rs.randn(3, 34, d).astype(np.float32) * 0.2, rs.randn(3, 34, d).astype(np.float32) * 0.2
            mean = np.array(K.make_args(variant).mean_dir_vec).squeeze()
            ro, rt = ref.convert_dir_vec_to_pose(out + mean), ref.convert_dir_vec_to_pose(tgt + mean)
            ro, rt = np.asarray(ro.cpu() if torch.is_tensor(ro) else ro), np.asarray(rt.cpu() if torch.is_tensor(rt) else rt)
            mine = evaluate.convert_dir_vec_to_pose(out + mean, variant)
            assert np.allclose(mine, ro, atol=1e-5)
            mae, acc = evaluate.pose_metrics(out, tgt, mean, 34, 4, variant)
            assert abs(mae - np.mean(np.absolute(ro[:, 4:] - rt[:, 4:]))) < 1e-6
            assert abs(acc - np.mean(np.abs(np.diff(rt, n=2, axis=1) - np.diff(ro, n=2, axis=1)))) < 1e-6
    finally:
        sys.path.remove(REF)
        for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[m]


@pytest.mark.gpu
def test_evaluate_testset_runs_and_matches_manual_metrics():
    """The loop on the GPU: outputs of the eval-mode cascade (already pinned module by module) -> the returned dict
    equals the metrics recomputed from the same outputs; modules return to training mode."""
    from helpers import build_modules
    from ha2g_b200 import cascade, rng
    from ha2g_b200.synthetic import make_batch
    dev = "cuda:0"
    args, gens, D, A, T = build_modules("gesture", 60, 5, {"gens": 20, "dis": 30, "audio": 31, "text": 32}, dev)
    batches = []
    for i in range(2):
        b = make_batch("gesture", 3, 60, 5, seed=40 + i)
        batches.append((torch.zeros(1), torch.zeros(1), b["in_text_padded"], None, b["target"], torch.zeros(3, 10), b["in_spec"], {}))
    noise = torch.zeros(3, 16, device=dev)
    import random
    random.seed(0)
    with rng.override(randn_fn=lambda s: noise):
        ret = evaluate.evaluate_testset(batches, None, *gens, A, None, None, args)
    assert set(ret) == {"loss", "joint_mae"} and ret["loss"] > 0 and ret["joint_mae"] > 0
    assert all(m.training for m in gens + [A])
    random.seed(0)
    losses = []
    for m in gens + [A]:
        m.train(False)
    with rng.override(randn_fn=lambda s: noise), torch.no_grad():
        for data in batches:
            ids = list(A.feat_extractor.z_obj.word2index.values())
            vid = torch.LongTensor([random.choice(ids) for _ in range(3)]).to(dev)
            _, _, _, _, blends = A(data[6].to(dev), vid)
            tg = data[4].to(dev)
            outs, _ = cascade.run_cascade("gesture", gens, cascade.split_targets("gesture", tg), data[2].to(dev), blends, vid, 4)
            losses.append(float((outs[-1] - tg).abs().mean()))
    for m in gens + [A]:
        m.train(True)
    assert abs(ret["loss"] - np.mean(losses)) < 1e-5


@pytest.mark.gpu
def test_evaluate_testset_with_fgd_evaluator():
    """evaluate_testset on the TED-Expressive cascade with the CUDA FGD evaluator (MotionAE features + Frechet distance):
    the returned dict carries frechet / feat_dist / diversity like the reference's (train_expressive.py:609-627)."""
    import argparse
    import types
    from helpers import build_modules
    from ha2g_b200 import rng
    from ha2g_b200.model.embedding_space_evaluator import EmbeddingSpaceEvaluator
    from ha2g_b200.model.motion_ae import MotionAE
    from ha2g_b200.synthetic import det_fill, make_batch
    dev = "cuda:0"
    args, gens, D, A, T = build_modules("expressive", 60, 5, {"gens": 20, "dis": 30, "audio": 31, "text": 32}, dev)
    ae = det_fill(MotionAE(126, 128), 90)
    ev = EmbeddingSpaceEvaluator(argparse.Namespace(n_pre_poses=4, n_poses=34, pose_dim=126), None, types.SimpleNamespace(),
                                 torch.device(dev), ckpt={"pose_dim": 126, "latent_dim": 128, "motion_ae": ae.state_dict()})
    batches = []
    for i in range(3):
        b = make_batch("expressive", 48, 60, 5, seed=60 + i)
        batches.append((torch.zeros(1), torch.zeros(1), b["in_text_padded"], None, b["target"], torch.zeros(48, 10), b["in_spec"], {}))
    ret = evaluate.evaluate_testset(batches, None, *gens, A, None, ev, args)
    assert {"loss", "joint_mae", "frechet", "feat_dist", "diversity"} <= set(ret)
    assert ev.get_no_of_samples() == 3 and np.isfinite(ret["frechet"]) and ret["frechet"] > 0 and ret["feat_dist"] > 0
    assert all(m.training for m in gens + [A])
