"""Synthetic inputs and deterministic parameter fill shared by tests, golden generation and bench.

The batch mirrors what ``SpeechMotionDataset.__getitem__`` / ``default_collate_fn`` deliver to
the training loop (reference: scripts/data_loader/lmdb_data_loader.py:45-55,108-176 and
scripts/train_expressive.py:321-337); distributions follow SURVEY.md section 8(d).
Everything is drawn from CPU ``torch.Generator``s so the same seed gives the same tensors on the
build container and on the GPU box.
"""
from __future__ import annotations

import zlib

import torch


def _gen(seed: int, name: str = "") -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def make_batch(variant: str, batch: int, n_words: int, n_speakers: int, seed: int = 0, n_poses: int = 34):
    """Returns dict(in_text_padded (B,T) i64, in_spec (B,128,70) f32, target (B,T,D) f32, vid (B,) i64)."""
    pose_dim = 126 if variant == "expressive" else 27
    g = _gen(seed, "batch")
    text = torch.zeros((batch, n_poses), dtype=torch.int64)
    for b in range(batch):
        k = int(torch.randint(4, 9, (1,), generator=g))
        pos = torch.randperm(n_poses, generator=g)[:k]
        text[b, pos] = torch.randint(4, n_words, (k,), generator=g)
    spec = (torch.rand((batch, 128, 70), generator=g) * -80.0).half().float()
    target = torch.randn((batch, n_poses, pose_dim), generator=g) * 0.1
    vid = torch.randint(1, n_speakers + 1, (batch,), generator=g)
    return {"in_text_padded": text, "in_spec": spec, "target": target, "vid": vid}


def make_audio(n_samples: int, seed: int = 0) -> torch.Tensor:
    g = _gen(seed, "audio")
    return (torch.randn(n_samples, generator=g) * 0.1).clamp_(-1, 1)


def make_embedding(n_words: int, dim: int = 300, seed: int = 0) -> torch.Tensor:
    """N(0, 1/dim) like Vocab.load_word_vectors' fallback (scripts/model/vocab.py:74-76)."""
    return torch.randn((n_words, dim), generator=_gen(seed, "emb")) / dim ** 0.5


@torch.no_grad()
def det_fill(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Overwrite every parameter/buffer with values that depend only on (seed, tensor name, shape).

    Works on the reference modules and on ours alike (same names => same values), so golden
    fixtures need not store multi-megabyte state dicts.
    """
    for name, p in list(module.named_parameters()) + list(module.named_buffers()):
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            p.zero_()
        elif leaf == "running_mean":
            p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        elif leaf == "running_var":
            p.copy_(1.0 + 0.2 * torch.rand(p.shape, generator=g))
        elif leaf == "weight_g":
            p.copy_(0.5 + torch.rand(p.shape, generator=g))
        elif leaf == "weight_v":
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        elif p.dim() == 1 and leaf == "weight":  # norm scale
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
        elif p.dim() == 1:  # biases
            p.copy_(0.05 * torch.randn(p.shape, generator=g))
        elif "embedding" in name and p.dim() == 2:
            p.copy_(torch.randn(p.shape, generator=g) / p.shape[1] ** 0.5)
        else:
            fan_in = p[0].numel()
            p.copy_(torch.randn(p.shape, generator=g) * (1.0 / fan_in) ** 0.5)
    return module


def sample_tensor(t: torch.Tensor, n: int = 256):
    """(norm, strided sample) summary of a tensor for compact golden storage."""
    f = t.detach().cpu().reshape(-1).double()
    stride = max(1, f.numel() // n)
    return {"norm": float(f.norm()), "numel": f.numel(), "stride": stride, "sample": f[::stride][:n].float().clone()}
