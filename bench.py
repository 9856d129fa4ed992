"""bench.py -- HA2G hierarchical hot path on N B200s: training-step throughput (default) or sliding-window inference.

    python bench.py [--gpus N --steps K --warmup W] [--impl ours|reference] [--variant expressive|gesture]
                    [--mode train|infer]

--mode train (BASELINE.json configs[1..3]): one "step" = one call of the drop-in ``train_iter_hierarchy_expressive``
(full step: discriminator step + generator step, epoch 11 > loss_warmup) on a synthetic batch of B_local = 128 clips of
34 frames per GPU; weak scaling: every rank processes its own 128 clips, gradients are all-reduced over NCCL before the
optimizer steps.  Before the warm-up the SAME path runs one step with injected randomness and its loss dict is checked
against the CPU oracle's value for the benchmark configuration (tests/golden/bench_step_losses.json, 1e-3).

--mode infer (configs[4]): one "step" = ``generate_gestures_hierarchy`` over 10 minutes of synthetic 16 kHz audio (log-mel
kernel + 300 serial windows at batch 1); the window chain is serial, so N GPUs run N independent clips ("replicas only").

--impl reference: the reference's CPU path (oracle port, pinned to the unmodified reference by tests/golden) on the host
cores, on the same workload, each step a bounded sample whose size is stated in the line.
Prints ONE JSON line (see README/DESIGN for the keys).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_WORDS = 30000      # reference vocabulary size is not recorded; 30k stated choice
N_SPEAKERS = 1500
T_FRAMES = 34
GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_step_losses.json")
# make_bench_golden.py constants (the golden step)
G_BATCH_SEED, G_EPS_SEED = 4242, 4243


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--variant", default="expressive", choices=["expressive", "gesture"])
    ap.add_argument("--batch", type=int, default=128, help="clips per GPU")
    ap.add_argument("--epoch", type=int, default=11, help="> loss_warmup (10): full step incl. discriminator")
    ap.add_argument("--minutes", type=float, default=10.0, help="--mode infer: length of the synthetic clip")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds of CPU work the reference arm / cpu_baseline may spend")
    ap.add_argument("--cpu-timeout", type=float, default=400.0, help="seconds allowed for the CPU-baseline subprocess")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-golden-check", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), start=3):
            if any(len(r) > i and r[i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of a kernel from this round's committed ncu capture (profiles/r02_ncu_traffic.json, written
    by tools/summarize_ncu.py from `ncu --set full` reports); None when no capture of this round exists."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        e = d.get(kernel_key)
        if e is None:
            return None, None
        return float(e["dram_bytes"]), f"profiles/r02_ncu_traffic.json[{kernel_key}] <- {e.get('source', '?')}"
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------------------------
# algorithmic work of the hot launchers (FLOPs from the call arguments; SURVEY.md 8(d) formulas)
# ------------------------------------------------------------------------------------------------------------
def launcher_flops(name, a):
    if name in ("ha2g_gemm_f32", "ha2g_gemm"):
        return 2.0 * a[4] * a[5] * a[6]
    if name == "ha2g_gru_layer_fwd":      # (x, I, ..., gi, y, gates, M, M_gates, T, H, stream)
        I, M, T, H = a[1], a[13], a[15], a[16]
        return 2.0 * M * T * 2 * (3 * H * I + 3 * H * H)
    if name == "ha2g_gru_layer_bwd":      # backward = 2x forward GEMM work
        I, M, T, H = a[4], a[23], a[24], a[25]
        return 4.0 * M * T * 2 * (3 * H * I + 3 * H * H)
    if name in ("ha2g_conv2d_fwd", "ha2g_conv2d_dgrad", "ha2g_conv2d_wgrad"):
        off = 4 if name == "ha2g_conv2d_fwd" else 3
        N, H, W, Cin, Cout, KH, KW, stride, pad = a[off:off + 9]
        Ho = (H + 2 * pad - KH) // stride + 1
        Wo = (W + 2 * pad - KW) // stride + 1
        return 2.0 * N * Ho * Wo * Cout * Cin * KH * KW
    if name == "ha2g_conv_tc":            # (a_hi, a_lo, b_hi, b_lo, bias, out, N, H, W, Cs, Cd, pad, KH, KW, prec, stream)
        N, H, W, Cs, Cd, pad, KH, KW = a[6:14]
        return 2.0 * N * (H + 2 * pad - KH + 1) * (W + 2 * pad - KW + 1) * Cs * Cd * KH * KW
    if name in ("ha2g_conv_wgrad_tc2", "ha2g_conv_wgrad_tc"):   # (x, dy, dwf, N, H, W, Cin, Cout, KH, KW, pad, ...)
        N, H, W, Cin, Cout, KH, KW, pad = a[3:11]
        return 2.0 * N * (H + 2 * pad - KH + 1) * (W + 2 * pad - KW + 1) * Cin * Cout * KH * KW
    return 0.0


def step_flops(variant, B, epoch_full=True):
    """Algorithmic GFLOP of one step, for the whole-step arithmetic roofline line."""
    per_sample = (31.3 if epoch_full else 28.2) if variant == "expressive" else (23.7 if epoch_full else 22.1)
    return per_sample * 1e9 * B


# ------------------------------------------------------------------------------------------------------------
def gru_kernel_roofline(dev, M, M_gates, T=T_FRAMES, H=300, iters=24):
    """Live CUDA-event timing of the fused GRU recurrence kernel (the kernel the north star names) at the shape the step
    launches it with (M rows of which the first M_gates save their gates), rotating over buffer sets larger than the
    126 MB L2 between launches.  -> (average launch ms, algorithmic FLOPs per launch, algorithmic bytes per launch, ...)."""
    import torch
    from ha2g_b200._lib import lib
    from ha2g_b200 import ops
    ops._ensure_workspace()
    st = torch.cuda.current_stream().cuda_stream
    per_set = (M * T * (6 * H + 2 * H) + M_gates * T * 8 * H) * 4     # gi read + y written + saved gates written
    nsets = max(2, int(140e6 // per_set) + 1)
    sets = []
    for _ in range(nsets):
        sets.append((torch.randn(M, T, 6 * H, device=dev), torch.empty(M, T, 2 * H, device=dev),
                     torch.empty(max(M_gates, 1), T, 8 * H, device=dev)))
    w = [torch.randn(3 * H, H, device=dev) * 0.05 for _ in range(2)]
    b = [torch.randn(3 * H, device=dev) * 0.05 for _ in range(2)]
    p = lambda t: t.data_ptr()

    def launch(i):
        gi, y, gates = sets[i % len(sets)]
        rc = lib.ha2g_gru_seq_fwd_tc2(p(gi), p(w[0]), p(w[1]), p(b[0]), p(b[1]), p(y), p(gates) if M_gates else None, M, M_gates, T, H, st)
        assert rc is None or rc == 0
    for i in range(4):
        launch(i)
    torch.cuda.synchronize()
    evs = []
    for i in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(i); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b_) for a, b_ in evs) / len(evs)
    return ms, 2.0 * M * T * 2 * (3 * H * H), float(per_set + 2 * 3 * H * H * 4), nsets * per_set


def mel_kernel_roofline(dev, n_samples, iters=12):
    """Live CUDA-event timing of the log-mel launcher on 10-minute clips, buffers rotated over > L2.
    -> (ms per launch, algorithmic bytes per launch: samples in (fp32) + [128, frames] fp32 out)."""
    import torch
    from ha2g_b200 import mel
    per = n_samples * 4
    nsets = max(2, int(140e6 // per) + 1)
    clips = [torch.randn(n_samples, device=dev) * 0.1 for _ in range(nsets)]
    for i in range(3):
        out = mel.extract_melspectrogram(clips[i % nsets])
    torch.cuda.synchronize()
    evs = []
    for i in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = mel.extract_melspectrogram(clips[i % nsets]); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    return ms, float(per + out.numel() * 4)


def build_world(variant, device, seed=0):
    import torch
    from ha2g_b200.constants import make_args
    from ha2g_b200.model.hierarchy_net import (Hierarchical_ConvDiscriminator, Hierarchical_PoseGenerator,
                                               Hierarchical_WavEncoder, TextEncoderTCN)
    from ha2g_b200.model.vocab import make_speaker_vocab
    from ha2g_b200.synthetic import make_embedding
    torch.manual_seed(seed)
    args = make_args(variant)
    spk = make_speaker_vocab(N_SPEAKERS)
    emb = make_embedding(N_WORDS, 300, 1).numpy()
    dims = (15, 21, 27) if variant == "gesture" else (24, 30, 36, 66, 96, 126)
    gens = [Hierarchical_PoseGenerator(args, d, N_WORDS, 300, emb, z_obj=spk).to(device) for d in dims]
    D = Hierarchical_ConvDiscriminator(dims[-1]).to(device)
    A = Hierarchical_WavEncoder(args, spk, pose_level=len(dims), nOut=32).to(device)
    T = TextEncoderTCN(args, N_WORDS, 300, pre_trained_embedding=emb, dropout=args.dropout_prob).to(device)
    lr = args.learning_rate
    mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
    opts = ([mk(g) for g in gens], mk(D, lr * args.discriminator_lr_weight), mk(A), mk(T))
    return args, gens, D, A, T, opts


def infer_inputs(minutes, seed=1):
    """10-minute synthetic clip: N(0, 0.1^2) audio at 16 kHz, a word every 0.2-0.8 s from a 2000-word vocabulary."""
    import numpy as np
    from ha2g_b200.model.vocab import Vocab
    from ha2g_b200.synthetic import make_audio
    lang = Vocab("words")
    for i in range(2000):
        lang.index_word(f"w{i}")
    n = int(minutes * 60 * 16000)
    audio = make_audio(n, seed).numpy()
    rs = np.random.RandomState(seed)
    t, words = 0.0, []
    while t < minutes * 60 - 1:
        t += rs.uniform(0.2, 0.8)
        words.append([f"w{rs.randint(2000)}", t, t + 0.2])
    return lang, audio, words


# ------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------------------
def _oracle_imports():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ha2g_oracle as O
    return O


def run_cpu_train(variant, B_full, epoch, steps, budget_s):
    """The training step restated by the oracle, SAME configuration as the GPU arm (vocabulary, speakers, module init of
    build_world): first one untimed step at the full batch; if `steps` more steps at that batch would not fit the CPU
    budget the timed steps use the largest power-of-two fraction of the batch that does -- and the line says so."""
    import torch
    O = _oracle_imports()
    from ha2g_b200 import constants as K
    from ha2g_b200.synthetic import _gen, make_batch
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    args, gens, D, A, T, _ = build_world(variant, "cpu", seed=0)
    sd = lambda m: {k: v.detach().clone() for k, v in m.state_dict().items()}
    state = {"gens": [sd(g) for g in gens], "dis": sd(D), "audio": sd(A), "text": sd(T)}
    del gens, D, A, T
    tabs = ({"pairs": K.EXPRESSIVE_ANGLE_PAIR, "avg": K.EXPRESSIVE_AVG_ANGLE, "var": K.EXPRESSIVE_VAR_ANGLE} if variant == "expressive"
            else {"pairs": K.GESTURE_ANGLE_PAIR, "avg": K.GESTURE_AVG_ANGLE, "var": K.GESTURE_VAR_ANGLE})
    L = len(state["gens"])
    opt_state = {}

    def one(B, it):
        nonlocal state
        batch = make_batch(variant, B, N_WORDS, N_SPEAKERS, seed=100 + it)
        eps = {k: [torch.randn((B, 16), generator=_gen(it, f"{k}{i}")) for i in range(L)] for k in ("d", "g", "r")}
        perm = torch.randperm(B, generator=_gen(it, "perm"))
        t0 = time.perf_counter()
        _, state, _ = O.train_step(variant, args, epoch, batch["in_text_padded"], batch["in_spec"], batch["target"],
                                   batch["vid"], state["gens"], state["dis"], state["audio"], state["text"], opt_state, eps,
                                   perm, tabs)
        return time.perf_counter() - t0
    t_warm = one(B_full, 0)
    B = B_full
    while B > 16 and t_warm * (B / B_full) * steps > budget_s:
        B //= 2
    times = [one(B, 1 + i) for i in range(steps)]
    mean_t = sum(times) / len(times)
    sample = (f"{steps} timed steps of the oracle train_step ({variant}, epoch {epoch}, n_words={N_WORDS}, n_speakers={N_SPEAKERS}) "
              f"at B={B} clips/step after 1 untimed step at B={B_full} ({t_warm:.1f} s); fp32, torch {torch.__version__} CPU, "
              f"{cores} threads, {mean_t:.2f} s/step")
    return {"value": B * T_FRAMES / mean_t, "unit": "pose-frames/s", "cores": cores, "kind": "port", "sample": sample,
            "sample_batch": B}, mean_t


def run_cpu_infer(variant, minutes_full, steps, budget_s):
    """The inference loop restated by the oracle (O.generate_gestures, pinned to the reference loop by tests/golden/
    inference.pt), on a bounded prefix of the same synthetic clip."""
    import numpy as np
    import torch
    O = _oracle_imports()
    import mel_oracle
    from ha2g_b200.synthetic import _gen
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    args, gens, D, A, T, _ = build_world(variant, "cpu", seed=0)
    sd = lambda m: {k: v.detach().clone() for k, v in m.state_dict().items()}
    gsd, asd = [sd(g) for g in gens], sd(A)
    dims = [g.out[2].weight.shape[0] if hasattr(g, "out") else None for g in gens]
    del gens, D, A, T
    lang, audio, words = infer_inputs(minutes_full)
    targets = [torch.randn((1, 34, d), generator=_gen(7, f"t{d}")) * 0.1 for d in dims]
    cnt = [0]

    def eps():
        cnt[0] += 1
        return torch.randn((1, 16), generator=_gen(8, f"e{cnt[0]}"))

    def one(minutes):
        n = int(minutes * 60 * 16000)
        t0 = time.perf_counter()
        out = O.generate_gestures(variant, args, gsd, asd, lang.get_word_index, audio[:n], words, targets, 3, eps,
                                  lambda a: mel_oracle.extract_melspectrogram(np.asarray(a)))
        return time.perf_counter() - t0, out.shape[0]
    t_warm, f_warm = one(0.2)                     # 6 windows, untimed
    per_min = t_warm / 0.2
    minutes = minutes_full
    while minutes > 0.5 and per_min * minutes * steps > budget_s:
        minutes /= 2
    res = [one(minutes) for _ in range(steps)]
    mean_t = sum(r[0] for r in res) / len(res)
    frames = res[0][1]
    sample = (f"{steps} timed passes of the oracle generate_gestures ({variant}) over the first {minutes:g} min of the "
              f"{minutes_full:g}-min clip ({frames} frames each) after one untimed 0.2-min pass; fp32, torch {torch.__version__} "
              f"CPU, {cores} threads, {mean_t:.2f} s/pass")
    return {"value": frames / mean_t, "unit": "pose-frames/s", "cores": cores, "kind": "port", "sample": sample,
            "sample_minutes": minutes}, mean_t


def _finish_process(world):
    """Tear the process group down cleanly: captured graphs that reference the communicator are dropped first."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch
        import torch.distributed as dist
        from ha2g_b200 import graph_step
        graph_step.reset()
        torch.cuda.synchronize()
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            os._exit(0)


def metric_and_config(a, world):
    ds = "Expressive" if a.variant == "expressive" else "Gesture"
    if a.mode == "train":
        metric = (f"pose-frames/sec training HA2G TED-{ds} "
                  f"({'train_iter_hierarchy_expressive' if a.variant == 'expressive' else 'train_iter_hierarchy'}, full step)")
        config = {"workload": f"config{'_expressive' if a.variant == 'expressive' else ''}/hierarchy.yml TED-{ds} synthetic batch, "
                              f"B_local={a.batch} clips x 34 frames, epoch {a.epoch} (D step + G step), n_words={N_WORDS}, "
                              f"n_speakers={N_SPEAKERS}",
                  "global_batch": a.batch * max(world, 1), "parallelism": f"dp{max(world, 1)}",
                  "l2": "working set (parameters + activations, > 2 GB per step) exceeds the 126 MB L2; no explicit flush"}
    else:
        metric = f"pose-frames/sec inference HA2G TED-{ds} (generate_gestures_hierarchy, {a.minutes:g}-min 16 kHz clip)"
        config = {"workload": f"synthesize_{'expressive_' if a.variant == 'expressive' else ''}hierarchy.py inference-only, "
                              f"{a.minutes:g}-min synthetic 16 kHz audio per GPU, 34-frame windows / 30-frame stride at batch 1, "
                              f"n_words={N_WORDS}, n_speakers={N_SPEAKERS}",
                  "clips": max(world, 1), "parallelism": f"replicas x{max(world, 1)} (the window chain is serial: one clip per GPU)",
                  "l2": "each timed pass streams a fresh 38 MB clip through the mel kernel; the window loop's working set "
                        "(parameters 190 MB) exceeds the 126 MB L2; no explicit flush"}
    return metric, config


def main_reference(a, rank, world):
    if rank != 0:
        return
    metric, config = metric_and_config(a, world)
    steps = max(1, min(a.steps, 5)) if a.steps < 5 else 5
    if a.mode == "train":
        cb, mean_t = run_cpu_train(a.variant, a.batch, a.epoch, steps, a.cpu_budget)
        config["reference_sample"] = f"B={cb['sample_batch']} clips per timed step (same vocabulary / speakers / module init)"
    else:
        cb, mean_t = run_cpu_infer(a.variant, a.minutes, steps, a.cpu_budget)
        config["reference_sample"] = f"first {cb['sample_minutes']:g} min of the clip per timed pass"
    line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "pose-frames/s", "n_gpus": 0,
            "steps": steps, "warmup": 1, "ms_per_step": mean_t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "pose-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def golden_check(a, fn, args, mods, opts, dev):
    """One step with injected randomness through the path the bench times, against the oracle's value for THIS
    configuration.  -> dict for the JSON line; raises SystemExit on a mismatch."""
    import torch
    from ha2g_b200 import rng
    from ha2g_b200.synthetic import _gen, make_batch
    if a.batch != 128 or a.epoch != 11 or not os.path.exists(GOLDEN):
        return {"checked": False, "why": "golden exists for B=128, epoch 11 only"}
    gold = json.load(open(GOLDEN)).get(a.variant)
    if gold is None or gold["n_words"] != N_WORDS or gold["n_speakers"] != N_SPEAKERS:
        return {"checked": False, "why": "no golden for this variant / vocabulary"}
    L = len(mods) - 3
    batch = {k: v.to(dev) for k, v in make_batch(a.variant, a.batch, N_WORDS, N_SPEAKERS, seed=gold["batch_seed"]).items()}
    draws = [torch.randn((a.batch, 16), generator=_gen(gold["eps_seed"], f"eps{i}")) for i in range(3 * L)]
    perm = torch.randperm(a.batch, generator=_gen(gold["eps_seed"], "perm"))
    with rng.override(randn_fn=rng.ListFeed(draws), randperm_fn=lambda n: perm.clone(), dropout=False):
        ret = fn(args, a.epoch, batch["in_text_padded"], batch["in_spec"], batch["target"], batch["vid"], *mods, *opts)
    worst = max(abs(ret[k] - v) / max(1.0, abs(v)) for k, v in gold["ret"].items())
    if set(ret) != set(gold["ret"]) or not worst <= 1e-3:
        raise SystemExit(f"bench.py: the CUDA step does not reproduce the oracle's losses at the benchmark configuration "
                         f"(worst relative error {worst:.3e} > 1e-3): cuda {ret} oracle {gold['ret']}")
    return {"checked": True, "worst_rel_err": worst, "tolerance": 1e-3, "golden": "tests/golden/bench_step_losses.json",
            "losses": {k: round(v, 5) for k, v in ret.items()}}


def main_train(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    from ha2g_b200 import dp, graph_step, ops
    from ha2g_b200.synthetic import make_batch
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
    metric, config = metric_and_config(a, world)
    args, gens, D, A, T, (gopts, dopt, aopt, topt) = build_world(a.variant, dev)
    fn = train_iter_hierarchy if a.variant == "gesture" else train_iter_hierarchy_expressive
    mods, opts = gens + [D, A, T], list(gopts) + [dopt, aopt, topt]
    # parity gate on the benchmark configuration (every rank, before data parallelism is switched on: replicas stay equal)
    gcheck = {"checked": False, "why": "--no-golden-check"} if a.no_golden_check else golden_check(a, fn, args, mods, opts, dev)
    if world > 1:
        dp.enable(world, modules=mods)

    host = [{k: v.pin_memory() for k, v in make_batch(a.variant, a.batch, N_WORDS, N_SPEAKERS, seed=1000 * rank + i).items()}
            for i in range(4)]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]

    def step(i, from_host):
        b = host[i % len(host)]
        if from_host:
            b = {k: v.to(dev, non_blocking=True) for k, v in b.items()}
        else:
            b = resident[i % len(resident)]
        return fn(args, a.epoch, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *mods, *opts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, from_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            ret = step(i, from_host)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ret

    # set-up: the public step runs eagerly twice, then captures itself into one CUDA graph (graph_step.py); these
    # calls are outside the warm-up count so that every warm-up and timed step below is the steady-state path
    for i in range(graph_step.WARMUP + 1 if graph_step.enabled() else 1):
        step(i, False)
    for i in range(max(a.warmup, 3)):
        step(i, False)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = ops.LAUNCHES[0]
    ms, ret = timed(a.steps, False)
    launches = (ops.LAUNCHES[0] - n0)
    ms_e2e, ret = timed(a.steps, True)
    clocks = sampler.stop() if rank == 0 else None

    # roofline leg: the same step launcher by launcher (eager, CUDA events around every C-ABI call on the launching
    # stream): one pass over all launchers, then `steps` passes timing only the dominant one
    ops.profile_begin(all_launchers=True, flops_fn=launcher_flops)
    step(0, False)
    prof = ops.profile_end()
    top = max(prof.items(), key=lambda kv: kv[1]["ms"])[0] if prof else None
    ops.profile_begin(only=top, flops_fn=launcher_flops)
    barrier()
    for i in range(a.steps):
        step(i, False)
    barrier()
    topstat = ops.profile_end().get(top, None)
    # the step's cascade runs once over [G; D; R] rows (ops.ride_along): M = 3B rows per launch (2B before the GAN warm-up
    # ends), gates saved for the B differentiated rows
    gru_M = a.batch * (3 if a.epoch > args.loss_warmup and args.loss_gan_weight > 0 else 2)
    kern = gru_kernel_roofline(dev, gru_M, a.batch) if rank == 0 else None
    if rank != 0:
        _finish_process(world)
        return

    frames = a.batch * T_FRAMES * world
    peaks = load_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    kern_ms, kern_flops, kern_bytes, rot_bytes = kern
    achieved = kern_flops / (kern_ms * 1e-3) / 1e12
    peak_burst = peaks.get("bf16_tflops", 1590.0)   # the kernel is timed alone: burst figure (recipe fallback 1.59 PF)
    peak_burst_src = ("measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback (B200_PROFILING.md 1.59 PF burst)")
    traffic, traffic_src = ncu_traffic(f"gru_seq_fwd_tc2_kernel_M{gru_M}")
    roofline = {"bound": "tensor", "kernel": "gru_seq_fwd_tc2_kernel (csrc/gru_cluster_tc2.cu), one bidirectional GRU layer, "
                                             f"M={gru_M} rows (gates saved for {a.batch}) x T=34 x H=300",
                "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s", "frac": achieved / peak_burst,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": kern_bytes,
                "peak_source": peak_burst_src, "us_per_launch": kern_ms * 1e3, "flops_per_launch": kern_flops,
                "note": "algorithmic fp32-equivalent FLOPs of the recurrence (2*M*T*2*3H*H; the kernel issues 3 bf16 MMAs per "
                        f"product) / average CUDA-event time of 24 launches at the step's shape, buffers rotated over "
                        f"{rot_bytes / 1e6:.0f} MB (> L2); the recurrence is latency-bound by design (34 dependent steps per launch)"}
    top_launcher = None
    if topstat and topstat["calls"]:
        la = topstat["flops"] / (topstat["ms"] * 1e-3) / 1e12 if topstat["ms"] > 0 else 0.0
        top_launcher = {"launcher": top, "achieved_tflops": la, "frac": la / peak_tf, "calls_timed": topstat["calls"],
                        "ms_timed": topstat["ms"], "share_of_step": topstat["ms"] / a.steps / (ms / a.steps) if ms > 0 else None,
                        "note": f"C-ABI launcher with the largest CUDA-event time over {a.steps} eager passes of the same step "
                                "right after the timed region (the timed region replays one CUDA graph)"}
    # launcher-level achieved rates of the other tensor-core engines (one eager pass, CUDA events per C-ABI call)
    others = []
    for name in ("ha2g_conv_tc", "ha2g_conv_wgrad_tc2", "ha2g_gemm", "ha2g_gru_layer_bwd", "ha2g_gru_layer_fwd"):
        v = prof.get(name)
        if v and v["ms"] > 0 and v["flops"] > 0:
            tf = v["flops"] / (v["ms"] * 1e-3) / 1e12
            others.append({"launcher": name, "calls": v["calls"], "ms": round(v["ms"], 3), "achieved_tflops": round(tf, 2),
                           "frac_of_sustained_bf16_peak": round(tf / peak_tf, 4)})
    line = {"metric": metric, "value": frames * a.steps / (ms * 1e-3), "unit": "pose-frames/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": frames * a.steps / (ms_e2e * 1e-3), "unit": "pose-frames/s",
                    "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": 4 * (len(ret) + len(gens) + 2) * world},
            "gpu_launches": launches, "cuda_graph": {"enabled": graph_step.enabled(), **graph_step.STATS},
            "clocks": clocks, "roofline": roofline, "top_launcher": top_launcher, "roofline_launchers": others,
            "step_tflops": step_flops(a.variant, a.batch * world) * a.steps / (ms * 1e-3) / 1e12,
            "parity_gate": gcheck,
            "last_losses": {k: round(v, 5) for k, v in ret.items()},
            "profile_top5": sorted(((k, round(v["ms"], 3), v["calls"]) for k, v in prof.items()), key=lambda r: -r[1])[:5]}
    if os.environ.get("HA2G_BENCH_PROFILE_ALL"):   # every launcher of one eager step (CUDA events, warm caches) -> stderr
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"  {k:34s} {v['ms']:8.3f} ms {v['calls']:5d} calls", file=sys.stderr)
    attach_cpu_baseline(a, line, world)
    print(json.dumps(line))
    _finish_process(world)


def attach_cpu_baseline(a, line, world):
    if a.no_cpu_baseline or world != 1:
        return
    # bounded sample in a subprocess (its own torch thread pool; killed if the host is too slow for the budget)
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--mode", a.mode, "--variant",
                              a.variant, "--batch", str(a.batch), "--epoch", str(a.epoch), "--minutes", str(a.minutes),
                              "--steps", "2", "--cpu-budget", "40"],
                             capture_output=True, text=True, timeout=a.cpu_timeout)
        ref_line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
        line["cpu_baseline"] = ref_line["cpu_baseline"]
    except Exception as e:  # keep the GPU line even if the host is too small for the sample
        line["cpu_baseline"] = {"value": None, "unit": "pose-frames/s", "cores": min(os.cpu_count() or 1, 32),
                                "kind": "port", "sample": f"failed: {type(e).__name__}"}


def main_infer(a, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    from ha2g_b200 import ops, synthesize
    from ha2g_b200.synthetic import _gen
    metric, config = metric_and_config(a, world)
    args, gens, D, A, T, _ = build_world(a.variant, dev)
    for m in gens + [A]:
        m.train(False)
    dims = [g.out[2].weight.shape[0] for g in gens]
    lang, audio, words = infer_inputs(a.minutes, seed=1 + rank)
    clips = [audio, np.roll(audio, 4001)]           # two distinct host clips, alternated
    pinned = [torch.from_numpy(c).pin_memory() for c in clips]
    resident = [p.to(dev) for p in pinned]
    targets = [torch.randn((1, 34, d), generator=_gen(7, f"t{d}")) * 0.1 for d in dims]
    run = lambda aud: synthesize.generate_gestures_hierarchy(args, *gens, A, lang, aud, words, *[x.clone() for x in targets], vid=3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, srcs):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            out = run(srcs[i % len(srcs)])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out
    for i in range(max(a.warmup, 3)):
        run(resident[0][: 16000 * 20]) if i < max(a.warmup, 3) - 1 else run(resident[0])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = ops.LAUNCHES[0]
    ms, out = timed(a.steps, resident)
    launches = ops.LAUNCHES[0] - n0
    ms_e2e, out = timed(a.steps, clips)             # host numpy in, numpy out: the reference-facing call
    clocks = sampler.stop() if rank == 0 else None
    mel_ms, mel_bytes = mel_kernel_roofline(dev, len(audio)) if rank == 0 else (None, None)
    if rank != 0:
        _finish_process(world)
        return
    frames = out.shape[0] * world
    peaks = load_peaks()
    peak_bw = peaks.get("hbm_gbs", 6650.0)
    ach = mel_bytes / (mel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic("logmel")
    roofline = {"bound": "hbm", "kernel": "ha2g_logmel (csrc/mel.cu): frame FFT + Slaney mel + dB over the whole clip",
                "achieved": ach, "peak": peak_bw, "unit": "GB/s", "frac": ach / peak_bw, "traffic": traffic,
                "traffic_source": traffic_src, "us_per_launch": mel_ms * 1e3, "algorithmic_bytes_per_launch": mel_bytes,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md 6.65 TB/s)",
                "note": "algorithmic bytes = clip samples (fp32) read once + [128, frames] fp32 written once; the window loop that "
                        "dominates the pass is a serial chain of batch-1 launches (latency-bound), see DESIGN.md"}
    line = {"metric": metric, "value": frames * a.steps / (ms * 1e-3), "unit": "pose-frames/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": frames * a.steps / (ms_e2e * 1e-3), "unit": "pose-frames/s",
                    "h2d_bytes_per_step": len(audio) * 4 * world, "d2h_bytes_per_step": out.size * 4 * world},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "frames_per_clip": int(out.shape[0]), "windows_per_clip": len(synthesize.window_plan(len(audio), 16000, 34, 4, 15)),
            "realtime_factor": (out.shape[0] / 15.0) / (ms / a.steps * 1e-3)}
    attach_cpu_baseline(a, line, world)
    print(json.dumps(line))
    _finish_process(world)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return main_reference(a, rank, world)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: ha2g_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if a.mode == "infer":
        return main_infer(a, rank, world, local_rank)
    return main_train(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
