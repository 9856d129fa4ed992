"""BASELINE.json configs[4]: inference-only, 10 min of synthetic 16 kHz audio through generate_gestures_hierarchy
(TED-Expressive, random-init weights, random word timings) on one GPU -> pose frames per second.
    python tools/bench_infer.py [--minutes 10] [--graph 0|1]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--graph", type=int, default=1)
    a = ap.parse_args()
    from ha2g_b200 import synthesize
    from ha2g_b200.constants import make_args
    from ha2g_b200.model.hierarchy_net import Hierarchical_PoseGenerator, Hierarchical_WavEncoder
    from ha2g_b200.model.vocab import Vocab, make_speaker_vocab
    from ha2g_b200.synthetic import make_audio, make_embedding
    synthesize._GRAPH = bool(a.graph)
    dev = "cuda:0"
    args = make_args("expressive")
    spk = make_speaker_vocab(1500)
    n_words = 30000
    emb = make_embedding(n_words, 300, 1).numpy()
    lang = Vocab("words")
    for i in range(2000):
        lang.index_word(f"w{i}")
    dims = (24, 30, 36, 66, 96, 126)
    torch.manual_seed(0)
    gens = [Hierarchical_PoseGenerator(args, d, n_words, 300, emb, z_obj=spk).to(dev).train(False) for d in dims]
    A = Hierarchical_WavEncoder(args, spk, pose_level=6, nOut=32).to(dev).train(False)
    n = int(a.minutes * 60 * 16000)
    audio = make_audio(n, 1).numpy()
    rs = np.random.RandomState(0)
    t, words = 0.0, []
    while t < a.minutes * 60 - 1:
        t += rs.uniform(0.2, 0.8)
        words.append([f"w{rs.randint(2000)}", t, t + 0.2])
    targets = [torch.randn(1, 34, d) * 0.1 for d in dims]
    run = lambda m: synthesize.generate_gestures_hierarchy(args, *gens, A, lang, audio[:int(m * 60 * 16000)], words,
                                                           *[x.clone() for x in targets], vid=3)
    run(0.2)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = run(a.minutes)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{a.minutes} min audio -> {out.shape[0]} frames in {dt:.3f} s = {out.shape[0] / dt:.0f} pose-frames/s "
          f"(graph={'on' if a.graph else 'off'})")


if __name__ == "__main__":
    main()
