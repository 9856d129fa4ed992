"""TEST INFRASTRUCTURE (not product code).  Fixture for the FGD evaluation path: imports the UNMODIFIED
reference ``model/motion_ae.py`` and ``model/embedding_space_evaluator.py`` (with stub ``umap`` / ``fasttext`` modules:
both are imported at module top and absent from the image), fills MotionAE deterministically, pushes three synthetic
batches of (real, generated) 34-frame TED-Expressive clips through ``push_samples`` in eval mode on the CPU and stores the
encoder features, reconstructions, side metrics and the (FGD, feat_dist) scores.  Writes tests/golden/fgd.pt.

    python oracle/make_fgd_golden.py
"""
import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for name in ("fasttext", "umap"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, "/root/reference/scripts")

import torch
from scipy import linalg as _linalg

# the reference calls scipy.linalg.sqrtm(x, disp=False) -> (sqrt, error estimate); current scipy dropped `disp`
_sqrtm = _linalg.sqrtm


def _sqrtm_compat(a, disp=True, **kw):
    r = _sqrtm(a, **kw)
    return r if disp else (r, 0.0)


_linalg.sqrtm = _sqrtm_compat

from model.embedding_space_evaluator import EmbeddingSpaceEvaluator   # noqa: E402  (the reference's)
from model.motion_ae import MotionAE                                   # noqa: E402
from ha2g_b200.synthetic import _gen, det_fill                         # noqa: E402

FILL_SEED, DATA_SEED, N_BATCH, B = 90, 91, 3, 40


def batches():
    out = []
    for i in range(N_BATCH):
        real = torch.randn((B, 34, 126), generator=_gen(DATA_SEED, f"real{i}")) * 0.3
        gen = real + torch.randn((B, 34, 126), generator=_gen(DATA_SEED, f"gen{i}")) * 0.1
        out.append((real, gen))
    return out


def main():
    net = det_fill(MotionAE(126, 128), FILL_SEED).train(False)
    ckpt = {"pose_dim": 126, "latent_dim": 128, "motion_ae": net.state_dict()}
    path = "/tmp/_fgd_ckpt.pt"
    torch.save(ckpt, path)
    args = argparse.Namespace(n_pre_poses=4, n_poses=34, pose_dim=126, wordembed_dim=300)
    lang = types.SimpleNamespace(word_embedding_weights=None, n_words=10)
    ev = EmbeddingSpaceEvaluator(args, path, lang, torch.device("cpu"))
    recs = []
    with torch.no_grad():
        for real, gen in batches():
            ev.push_samples(None, None, gen, real)
            recon, z = ev.net(real)
            recs.append({"z_real": z.clone(), "recon_real": recon.clone()})
    frechet, feat_dist = ev.get_scores()
    torch.save({"fill_seed": FILL_SEED, "data_seed": DATA_SEED, "n_batch": N_BATCH, "B": B, "batches": recs,
                "recon_err_diff": [float(x) for x in ev.recon_err_diff], "cos_err_diff": [float(x) for x in ev.cos_err_diff],
                "frechet": float(frechet), "feat_dist": float(feat_dist)}, os.path.join(ROOT, "tests", "golden", "fgd.pt"))
    print("fgd.pt written: FGD", frechet, "feat_dist", feat_dist, "recon_err_diff", [float(x) for x in ev.recon_err_diff])


if __name__ == "__main__":
    main()
