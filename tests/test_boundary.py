"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol of
include/ha2g_b200.h, module state_dicts equal the reference's, cascade tables equal the reference's explicit
slices, and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return g.LIB


def test_library_exports_every_header_symbol(built):
    from ha2g_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 40
    dll = ctypes.CDLL(built)
    for name in protos:
        assert hasattr(dll, name), f"{name} declared in include/ha2g_b200.h but not exported"
    text = open(_lib.HEADER_PATH).read()
    assert "extern \"C\"" in text
    for line in text.splitlines():
        if line.startswith("int ha2g_"):  # plain pointers and sizes only: no torch/ATen types in the signatures
            assert not re.search(r"torch|at::|Tensor|c10", line), line


def test_header_cites_reference_for_core_entry_points():
    text = open(os.path.join(ROOT, "include", "ha2g_b200.h")).read()
    for sym, cite in (("ha2g_gru_layer_fwd", "hierarchy_net.py"), ("ha2g_tcn_weight_fwd", "tcn.py"),
                      ("ha2g_conv2d_fwd", "Conv2d"), ("ha2g_contrastive_fwd_rect", "Contrastive")):
        i = text.index(f"int {sym}(")
        assert cite.lower() in text[max(0, i - 1500):i].lower(), f"{sym}: missing reference citation"


def test_state_dict_contract():
    """Key names, shapes and parameter order equal the reference modules' (checkpoint compatibility)."""
    from ha2g_b200.constants import make_args
    from ha2g_b200.model.hierarchy_net import (Hierarchical_ConvDiscriminator, Hierarchical_PoseGenerator,
                                               Hierarchical_WavEncoder, TextEncoderTCN)
    from ha2g_b200.model.vocab import make_speaker_vocab
    from ha2g_b200.synthetic import make_embedding
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    args, spk, emb = make_args("expressive"), make_speaker_vocab(5), make_embedding(60, 300, 1).numpy()
    mods = {"generator126": Hierarchical_PoseGenerator(args, 126, 60, 300, emb, z_obj=spk),
            "discriminator126": Hierarchical_ConvDiscriminator(126),
            "audio6": Hierarchical_WavEncoder(args, spk, pose_level=6, nOut=32),
            "text": TextEncoderTCN(args, 60, 300, pre_trained_embedding=emb, dropout=0.3)}
    for k, m in mods.items():
        sd = {n: list(t.shape) for n, t in m.state_dict().items()}
        assert sd == ref[k], k
        assert [n for n, _ in m.named_parameters()] == ref["_param_order"][k], k
        m2 = type(m).__new__(type(m))  # round trip through load_state_dict
        m.load_state_dict(m.state_dict())


def test_cascade_tables_match_reference_slices():
    """Spot checks copied from the explicit slice assignments (train_hierarchy_expressive.py:252-310,
    train_hierarchy.py:161-169), including the '-5*3:' head-bone shift quirk."""
    from ha2g_b200 import cascade
    tabs = cascade.host_tables("expressive")
    assert [len(t[0]) for t in tabs] == [24, 30, 36, 66, 96, 126]
    # pre_seq_5[:, n_pre:, 19*3:20*3] = out_4[:, n_pre:, 13*3:14*3]
    assert tabs[4][1][19 * 3:20 * 3] == [39, 40, 41]
    # pre_seq_6[:, n_pre:, 20*3:24*3] = out_5[:, n_pre:, 15*3:19*3]
    assert tabs[5][1][20 * 3:24 * 3] == list(range(45, 57))
    # pre_seq_2[:, n_pre:, -5*3:] = out_1[:, n_pre:, -5*3:]  -> columns 16..30 of the 31-wide pre_seq_2
    assert tabs[1][1][16:31] == list(range(9, 24)) and tabs[1][1][15] == -1
    # newly introduced bones stay zero: level 2 adds bones 3 and 20 (slots 3 and 4)
    assert tabs[1][1][9:15] == [-1] * 6
    g = cascade.host_tables("gesture")
    # pre_seq_2[:, n_pre:, 5*3:6*3] = out_1[:, n_pre:, 4*3:5*3];  pre_seq_3[:, n_pre:, 6*3:8*3] = out_2[:, n_pre:, 5*3:7*3]
    assert g[1][1][15:18] == [12, 13, 14] and g[2][1][18:24] == list(range(15, 21)) and g[2][1][27] == -1
    # targets: target_1 = cat(target[:, :, :4*3], target[:, :, 6*3:7*3])
    assert g[0][0] == list(range(12)) + [18, 19, 20]


def test_no_cpu_fallback():
    """Ops refuse CPU tensors: there is no eager/PyTorch fallback behind the modules."""
    from ha2g_b200 import ops
    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros(2, 3), torch.zeros(4, 3), torch.zeros(4))


def test_synthetic_batch_contract():
    from ha2g_b200.synthetic import make_batch
    b = make_batch("expressive", 4, 100, 7, seed=1)
    assert b["in_text_padded"].shape == (4, 34) and b["in_text_padded"].dtype == torch.int64
    assert b["in_spec"].shape == (4, 128, 70) and float(b["in_spec"].min()) >= -80 and float(b["in_spec"].max()) <= 0
    assert b["target"].shape == (4, 34, 126) and b["vid"].min() >= 1 and b["vid"].max() <= 7
    b2 = make_batch("expressive", 4, 100, 7, seed=1)
    assert all(torch.equal(b[k], b2[k]) for k in b)
