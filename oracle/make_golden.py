"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
Test infrastructure, not product.  RNG control: every nn.Dropout p (and nn.GRU.dropout) is set
to 0 after construction, ``model.embedding_net.reparameterize`` is patched to take its noise from
a fixed list, ``torch.randperm`` inside the step is patched to a fixed permutation.  Every module gets its
OWN copy of the word-embedding matrix: ``nn.Embedding.from_pretrained(torch.FloatTensor(ndarray))``
(hierarchy_net.py:32) aliases the ndarray, so on CPU all seven text encoders would share one table; the
reference's real (GPU) runs break that aliasing in ``.to(device)``, and the copies reproduce that.  Parameters are
overwritten by ha2g_b200.synthetic.det_fill so fixtures only store seeds, outputs and summaries.
"""
import os
import sys
import types
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.modules.setdefault("fasttext", types.ModuleType("fasttext"))
sys.path.insert(0, "/root/reference/scripts")
warnings.filterwarnings("ignore")

# Stub the third-party packages the reference's *drivers* import but this image lacks (none of them is on the
# arithmetic path we pin; librosa's mel is replaced by oracle/mel_oracle.py, which is why row K stays "unpinned").
import importlib.abc  # noqa: E402
import importlib.machinery  # noqa: E402
from unittest import mock  # noqa: E402

_STUBS = ("librosa", "soundfile", "lmdb", "matplotlib", "mpl_toolkits", "configargparse", "umap", "tensorboardX",
          "google", "gentle", "tensorboard", "torch.utils.tensorboard")


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []
        m.__getattr__ = lambda name, _n=spec.name: mock.MagicMock(name=f"{_n}.{name}")
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _STUBS or name in _STUBS:
            return importlib.machinery.ModuleSpec(name, _StubLoader(), is_package=True)
        return None


sys.meta_path.insert(0, _StubFinder())

from model import vocab  # noqa: E402
import model.embedding_net  # noqa: E402
from model.hierarchy_net import (Hierarchical_ConvDiscriminator, Hierarchical_PoseGenerator,  # noqa: E402
                                 Hierarchical_WavEncoder, TextEncoderTCN)
import train_eval.train_hierarchy as th  # noqa: E402
import train_eval.train_hierarchy_expressive as the  # noqa: E402

from ha2g_b200.constants import make_args  # noqa: E402
from ha2g_b200.synthetic import det_fill, make_batch, make_embedding, sample_tensor, _gen  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N_WORDS, N_SPK = 60, 5
torch.set_num_threads(8)


def speaker_vocab(n):
    v = vocab.Vocab("vid", insert_default_tokens=False)
    for i in range(n):
        v.index_word(f"v{i}")
    return v


def no_dropout(m):
    for s in m.modules():
        if isinstance(s, torch.nn.Dropout):
            s.p = 0.0
        if isinstance(s, torch.nn.GRU):
            s.dropout = 0.0
    return m


def grads_summary(m):
    return {n: sample_tensor(p.grad) for n, p in m.named_parameters() if p.grad is not None}


def params_summary(m):
    out = {n: sample_tensor(p) for n, p in m.named_parameters()}
    out.update({n: sample_tensor(b.float()) for n, b in m.named_buffers()})
    return out


def randn(shape, seed, name):
    return torch.randn(shape, generator=_gen(seed, name))


class EpsFeed:
    """Replaces model.embedding_net.reparameterize (embedding_net.py:10-13) with injected noise."""

    def __init__(self, seed, batch):
        self.seed, self.batch, self.n = seed, batch, 0

    def __call__(self, mu, logvar):
        eps = randn((self.batch, 16), self.seed, f"eps{self.n}")
        self.n += 1
        return mu + eps * torch.exp(0.5 * logvar)


def golden_modules():
    args = make_args("expressive")
    spk = speaker_vocab(N_SPK)
    emb = make_embedding(N_WORDS, 300, 1).numpy()
    B = 2
    batch = make_batch("expressive", B, N_WORDS, N_SPK, seed=3)
    gold = {"n_words": N_WORDS, "n_spk": N_SPK, "batch_seed": 3, "B": B}

    # --- TextEncoderTCN
    m = no_dropout(det_fill(TextEncoderTCN(args, N_WORDS, 300, pre_trained_embedding=emb.copy(), dropout=0.3), 11))
    out = m(batch["in_text_padded"])
    gout = randn(out.shape, 5, "gout_text")
    (out * gout).sum().backward()
    gold["text"] = {"fill_seed": 11, "out": out.detach(), "grads": grads_summary(m)}

    # --- generator (expressive level 6 and gesture level 1 input sizes)
    for tag, d in (("gen126", 126), ("gen15", 15)):
        m = no_dropout(det_fill(Hierarchical_PoseGenerator(args, d, N_WORDS, 300, emb.copy(), z_obj=spk), 12))
        pre = randn((B, 34, d + 1), 5, "pre" + tag) * 0.1
        aud = randn((B, 34, 32), 5, "aud" + tag)
        pre.requires_grad_(True); aud.requires_grad_(True)
        model.embedding_net.reparameterize = EpsFeed(7, B)
        out, z, mu, lv = m(pre, batch["in_text_padded"], aud, batch["vid"])
        gout = randn(out.shape, 5, "gout" + tag)
        ((out * gout).sum() + z.sum() * 0.3 + (mu * mu).sum() * 0.2 + lv.sum() * 0.1).backward()
        gold[tag] = {"fill_seed": 12, "out": out.detach(), "z": z.detach(), "mu": mu.detach(), "logvar": lv.detach(),
                     "grads": grads_summary(m), "dpre": pre.grad.clone(), "daud": aud.grad.clone()}

    # --- discriminator (train mode: batch-stat BN + running-stat update)
    m = no_dropout(det_fill(Hierarchical_ConvDiscriminator(126), 13))
    poses = batch["target"].clone().requires_grad_(True)
    out = m(poses)
    gout = randn(out.shape, 5, "gout_dis")
    (out * gout).sum().backward()
    gold["dis"] = {"fill_seed": 13, "out": out.detach(), "grads": grads_summary(m), "dposes": poses.grad.clone(),
                   "buffers": {n: b.clone() for n, b in m.named_buffers()}}
    m.eval()
    gold["dis"]["out_eval"] = m(batch["target"]).detach()

    # --- audio encoder
    m = det_fill(Hierarchical_WavEncoder(args, spk, pose_level=6, nOut=32), 14)
    w, fl, fm, fh, blend = m(batch["in_spec"], batch["vid"])
    loss = 0
    for i, t in enumerate([w, fl, fm, fh] + blend):
        loss = loss + (t * randn(t.shape, 5, f"gout_aud{i}")).sum()
    loss.backward()
    gold["audio"] = {"fill_seed": 14, "weight": w.detach(), "feat_low": fl.detach(), "feat_mid": fm.detach(),
                     "feat_high": fh.detach(), "blend": [b.detach() for b in blend], "grads": grads_summary(m),
                     "buffers": {n: sample_tensor(b.float()) for n, b in m.named_buffers()}}
    m.eval()
    w, fl, fm, fh, blend = m(batch["in_spec"], batch["vid"])
    gold["audio"]["eval"] = {"weight": w.detach(), "feat_low": fl.detach(), "feat_high": fh.detach(), "blend5": blend[5].detach()}

    # --- contrastive, both variants
    for tag, mod in (("gesture", th), ("expressive", the)):
        a = randn((68, 32), 5, "ca").requires_grad_(True)
        b = randn((68, 32), 5, "cb").requires_grad_(True)
        l = mod.SoftmaxContrastiveLoss()(a, b)
        l.backward()
        gold["contrastive_" + tag] = {"loss": l.detach(), "da": a.grad.clone(), "db": b.grad.clone()}
    torch.save(gold, os.path.join(OUT, "modules.pt"))
    print("modules.pt written")


def golden_steps():
    for variant, mod, fn_name, dims in (("gesture", th, "train_iter_hierarchy", (15, 21, 27)),
                                        ("expressive", the, "train_iter_hierarchy_expressive", (24, 30, 36, 66, 96, 126))):
        args = make_args(variant)
        spk = speaker_vocab(N_SPK)
        emb = make_embedding(N_WORDS, 300, 1).numpy()
        B = 3
        gold = {"n_words": N_WORDS, "n_spk": N_SPK, "B": B, "steps": []}
        gens = [no_dropout(det_fill(Hierarchical_PoseGenerator(args, d, N_WORDS, 300, emb.copy(), z_obj=spk), 20 + i))
                for i, d in enumerate(dims)]
        D = no_dropout(det_fill(Hierarchical_ConvDiscriminator(dims[-1]), 30))
        A = det_fill(Hierarchical_WavEncoder(args, spk, pose_level=len(dims), nOut=32), 31)
        T = no_dropout(det_fill(TextEncoderTCN(args, N_WORDS, 300, pre_trained_embedding=emb.copy(), dropout=0.3), 32))
        lr = args.learning_rate
        gopts = [torch.optim.Adam(g.parameters(), lr=lr, betas=(0.5, 0.999)) for g in gens]
        dopt = torch.optim.Adam(D.parameters(), lr=lr * args.discriminator_lr_weight, betas=(0.5, 0.999))
        aopt = torch.optim.Adam(A.parameters(), lr=lr, betas=(0.5, 0.999))
        topt = torch.optim.Adam(T.parameters(), lr=lr, betas=(0.5, 0.999))
        orig_randperm = torch.randperm
        for step, epoch in enumerate((0, 11, 11)):
            batch = make_batch(variant, B, N_WORDS, N_SPK, seed=40 + step)
            model.embedding_net.reparameterize = EpsFeed(50 + step, B)
            perm = orig_randperm(B, generator=_gen(60 + step, "perm"))
            torch.randperm = lambda n, *a, **k: perm.clone()
            try:
                ret = getattr(mod, fn_name)(args, epoch, batch["in_text_padded"], batch["in_spec"], batch["target"],
                                            batch["vid"], *gens, D, A, T, *gopts, dopt, aopt, topt)
            finally:
                torch.randperm = orig_randperm
            rec = {"epoch": epoch, "batch_seed": 40 + step, "eps_seed": 50 + step, "perm": perm, "ret": ret,
                   "gens": [params_summary(g) for g in gens], "dis": params_summary(D),
                   "audio": params_summary(A), "text": params_summary(T)}
            if step == 0:
                rec["grads"] = {"g_last": grads_summary(gens[-1]), "g_first": grads_summary(gens[0]),
                                "audio": grads_summary(A), "text": grads_summary(T)}
            gold["steps"].append(rec)
            print(variant, "step", step, "epoch", epoch, {k: round(v, 5) for k, v in ret.items()})
        gold["fill_seeds"] = {"gens": 20, "dis": 30, "audio": 31, "text": 32}
        torch.save(gold, os.path.join(OUT, f"step_{variant}.pt"))




def golden_keys():
    """state_dict key -> shape of every reference module on the path (checkpoint contract)."""
    import json
    args = make_args("expressive")
    spk = speaker_vocab(N_SPK)
    emb = make_embedding(N_WORDS, 300, 1).numpy()
    mods = {
        "generator126": Hierarchical_PoseGenerator(args, 126, N_WORDS, 300, emb.copy(), z_obj=spk),
        "discriminator126": Hierarchical_ConvDiscriminator(126),
        "audio6": Hierarchical_WavEncoder(args, spk, pose_level=6, nOut=32),
        "text": TextEncoderTCN(args, N_WORDS, 300, pre_trained_embedding=emb.copy(), dropout=0.3),
    }
    out = {k: {n: list(t.shape) for n, t in m.state_dict().items()} for k, m in mods.items()}
    out["_param_order"] = {k: [n for n, _ in m.named_parameters()] for k, m in mods.items()}
    json.dump(out, open(os.path.join(OUT, "state_dict_keys.json"), "w"), indent=0)


def golden_inference():
    """generate_gestures_hierarchy (scripts/synthesize_expressive_hierarchy.py:36-259) UNMODIFIED, eval-mode modules,
    6 s of synthetic audio (3 windows), with only librosa's mel swapped for oracle/mel_oracle.py."""
    import pyarrow
    if not hasattr(pyarrow, "serialize"):
        pyarrow.serialize = pyarrow.deserialize = lambda *a, **k: None
    import synthesize_expressive_hierarchy as S
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mel_oracle
    from ha2g_b200.synthetic import make_audio
    S.device = torch.device("cpu")
    S.extract_melspectrogram = lambda audio, sr=16000: mel_oracle.extract_melspectrogram(audio)
    args = make_args("expressive")
    spk = speaker_vocab(N_SPK)
    emb = make_embedding(N_WORDS, 300, 1).numpy()
    lang = vocab.Vocab("words")
    vocab_words = [f"w{i}" for i in range(N_WORDS - 4)]
    for w in vocab_words:
        lang.index_word(w)
    assert lang.n_words == N_WORDS
    dims = (24, 30, 36, 66, 96, 126)
    gens = [det_fill(Hierarchical_PoseGenerator(args, d, N_WORDS, 300, emb.copy(), z_obj=spk), 70 + i).train(False)
            for i, d in enumerate(dims)]
    A = det_fill(Hierarchical_WavEncoder(args, spk, pose_level=6, nOut=32), 80).train(False)
    audio = make_audio(96000, 5).numpy()
    words = [[vocab_words[(7 * i) % len(vocab_words)], 0.3 + 0.37 * i, 0.3 + 0.37 * i + 0.25] for i in range(15)]
    words.append(["not-in-vocab", 5.7, 5.9])
    targets = [randn((1, 34, d), 81, f"t{d}") * 0.1 for d in dims]
    model.embedding_net.reparameterize = EpsFeed(82, 1)
    out = S.generate_gestures_hierarchy(args, *gens, A, lang, audio, words, *[t.clone() for t in targets], vid=2)
    out_fade = None
    model.embedding_net.reparameterize = EpsFeed(82, 1)
    out_fade = S.generate_gestures_hierarchy(args, *gens, A, lang, audio, words, *[t.clone() for t in targets], vid=2,
                                             fade_out=True)
    torch.save({"out": torch.from_numpy(out), "out_fade": torch.from_numpy(out_fade), "n_words": N_WORDS, "n_spk": N_SPK,
                "words": words, "vocab_words": vocab_words, "audio_seed": 5, "n_samples": 96000, "eps_seed": 82, "vid": 2,
                "fill_seeds": {"gens": 70, "audio": 80}, "target_seed": 81}, os.path.join(OUT, "inference.pt"))
    print("inference.pt written", out.shape, out_fade.shape)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_keys()
    if "--inference" in sys.argv:
        golden_inference()
        sys.exit(0)
    if "--keys" not in sys.argv:
        golden_modules()
        golden_steps()
        golden_inference()
