"""bench.py -- HA2G hierarchical training-step throughput (pose-frames/s) on N B200s.

    python bench.py [--gpus N --steps K --warmup W] [--impl ours|reference] [--variant expressive|gesture]

One "step" = one call of the drop-in ``train_iter_hierarchy_expressive`` (full step: discriminator step +
generator step, epoch 11 > loss_warmup) on a synthetic batch of B_local = 128 clips of 34 frames per GPU
(BASELINE.json configs[2] / configs[3]); weak scaling: every rank processes its own 128 clips, gradients are
all-reduced over NCCL before the optimizer steps.  Prints ONE JSON line (see README/DESIGN for the keys).
"""
from __future__ import annotations

import argparse
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="expressive", choices=["expressive", "gesture"])
    ap.add_argument("--batch", type=int, default=128, help="clips per GPU")
    ap.add_argument("--epoch", type=int, default=11, help="> loss_warmup (10): full step incl. discriminator")
    ap.add_argument("--cpu-batch", type=int, default=16, help="clips in the CPU-baseline sample")
    ap.add_argument("--cpu-timeout", type=float, default=240.0, help="seconds allowed for the CPU-baseline subprocess")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


# ------------------------------------------------------------------------------------------------------------
# algorithmic work of the hot launchers (FLOPs from the call arguments; SURVEY.md 8(d) formulas)
# ------------------------------------------------------------------------------------------------------------
def launcher_flops(name, a):
    if name in ("ha2g_gemm_f32", "ha2g_gemm"):
        return 2.0 * a[4] * a[5] * a[6]
    if name == "ha2g_gru_layer_fwd":      # (x, I, ..., gi, y, gates, M, T, H, stream)
        I, M, T, H = a[1], a[13], a[14], a[15]
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
    return 0.0


def step_flops(variant, B, epoch_full=True):
    """Algorithmic GFLOP of one step, for the whole-step arithmetic roofline line."""
    per_sample = (31.3 if epoch_full else 28.2) if variant == "expressive" else (23.7 if epoch_full else 22.1)
    return per_sample * 1e9 * B


# ------------------------------------------------------------------------------------------------------------
def gru_kernel_roofline(dev, M, T=T_FRAMES, H=300, iters=24):
    """Live CUDA-event timing of the fused GRU recurrence kernel (gru_seq_fwd_tc2_kernel, the kernel the north star
    names) at the workload's shape, rotating over buffer sets larger than the 126 MB L2 between launches.
    -> (average launch ms, algorithmic FLOPs per launch)."""
    import torch
    from ha2g_b200._lib import lib
    from ha2g_b200 import ops
    ops._ensure_workspace()
    st = torch.cuda.current_stream().cuda_stream
    sets = []
    for _ in range(4):   # 4 x (gi 31 MB + y 10 MB + gates 42 MB) = 334 MB > L2
        sets.append((torch.randn(M, T, 6 * H, device=dev), torch.empty(M, T, 2 * H, device=dev), torch.empty(M, T, 8 * H, device=dev)))
    w = [torch.randn(3 * H, H, device=dev) * 0.05 for _ in range(2)]
    b = [torch.randn(3 * H, device=dev) * 0.05 for _ in range(2)]
    p = lambda t: t.data_ptr()

    def launch(i):
        gi, y, gates = sets[i % len(sets)]
        rc = lib.ha2g_gru_seq_fwd_tc2(p(gi), p(w[0]), p(w[1]), p(b[0]), p(b[1]), p(y), p(gates), M, T, H, st)
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
    return ms, 2.0 * M * T * 2 * (3 * H * H)


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


def run_cpu_port(variant, B, epoch, steps, warmup):
    """The reference's CPU path restated (oracle/ha2g_oracle.py, pinned to the reference by tests/golden) timed on the
    host cores: same step, same synthetic batch distribution, B clips."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ha2g_oracle as O
    from ha2g_b200 import constants as K
    from ha2g_b200.synthetic import make_batch
    from helpers import build_modules, sd_cpu
    # all host cores up to 32 threads: beyond that the step's many small ops (B<=16 GRU steps, 16x9 feature maps)
    # only lose time to oversubscription; the count actually used is what the JSON line reports
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    args, gens, D, A, T = build_modules(variant, 1000, 50, {"gens": 20, "dis": 30, "audio": 31, "text": 32}, "cpu")
    state = {"gens": [sd_cpu(m) for m in gens], "dis": sd_cpu(D), "audio": sd_cpu(A), "text": sd_cpu(T)}
    tabs = ({"pairs": K.EXPRESSIVE_ANGLE_PAIR, "avg": K.EXPRESSIVE_AVG_ANGLE, "var": K.EXPRESSIVE_VAR_ANGLE} if variant == "expressive"
            else {"pairs": K.GESTURE_ANGLE_PAIR, "avg": K.GESTURE_AVG_ANGLE, "var": K.GESTURE_VAR_ANGLE})
    L = len(gens)
    opt_state, times = {}, []
    for it in range(warmup + steps):
        batch = make_batch(variant, B, 1000, 50, seed=100 + it)
        g = torch.Generator().manual_seed(it)
        eps = {k: [torch.randn((B, 16), generator=g) for _ in range(L)] for k in ("d", "g", "r")}
        perm = torch.randperm(B, generator=g)
        t0 = time.perf_counter()
        _, state, _ = O.train_step(variant, args, epoch, batch["in_text_padded"], batch["in_spec"], batch["target"],
                                   batch["vid"], state["gens"], state["dis"], state["audio"], state["text"], opt_state, eps,
                                   perm, tabs)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean_t = sum(times) / len(times)
    return {"value": B * T_FRAMES / mean_t, "unit": "pose-frames/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} steps of the oracle train_step ({variant}, epoch {epoch}) at B={B} clips, fp32, "
                      f"torch {torch.__version__} CPU, {mean_t:.2f} s/step"}, mean_t


def _finish_process(world):
    """Leave without tearing NCCL down: destroy_process_group() blocks for minutes while captured CUDA graphs still
    reference the communicator (observed at N=2), and nothing after the JSON line needs a clean shutdown."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch
        from ha2g_b200 import graph_step
        graph_step.reset()
        torch.cuda.synchronize()
        os._exit(0)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "pose-frames/sec training HA2G TED-Expressive (train_iter_hierarchy_expressive, full step)" \
        if a.variant == "expressive" else "pose-frames/sec training HA2G TED-Gesture (train_iter_hierarchy, full step)"
    config = {"workload": f"config{'_expressive' if a.variant == 'expressive' else ''}/hierarchy.yml TED-"
                          f"{'Expressive' if a.variant == 'expressive' else 'Gesture'} synthetic batch, B_local={a.batch} clips x 34 frames, "
                          f"epoch {a.epoch} (D step + G step), n_words={N_WORDS}, n_speakers={N_SPEAKERS}",
              "global_batch": a.batch * max(world, 1), "parallelism": f"dp{max(world, 1)}",
              "l2": "working set (parameters + activations, > 2 GB per step) exceeds the 126 MB L2; no explicit flush"}

    if a.impl == "reference":
        if rank != 0:
            return
        cb, mean_t = run_cpu_port(a.variant, a.cpu_batch, a.epoch, max(1, min(a.steps, 3)), 1)
        line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "pose-frames/s", "n_gpus": 0,
                "steps": max(1, min(a.steps, 3)), "warmup": 1, "ms_per_step": mean_t * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "pose-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: ha2g_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ha2g_b200 import dp, ops
    from ha2g_b200.synthetic import make_batch
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    from ha2g_b200.train_eval.train_hierarchy_expressive import train_iter_hierarchy_expressive
    args, gens, D, A, T, (gopts, dopt, aopt, topt) = build_world(a.variant, dev)
    if world > 1:
        dp.enable(world, modules=gens + [D, A, T])
    fn = train_iter_hierarchy if a.variant == "gesture" else train_iter_hierarchy_expressive

    host = [{k: v.pin_memory() for k, v in make_batch(a.variant, a.batch, N_WORDS, N_SPEAKERS, seed=1000 * rank + i).items()}
            for i in range(4)]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def step(i, from_host):
        b = host[i % len(host)]
        if from_host:
            b = {k: v.to(dev, non_blocking=True) for k, v in b.items()}
        else:
            b = resident[i % len(resident)]
        return fn(args, a.epoch, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *gens, D, A, T, *gopts, dopt,
                  aopt, topt)

    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]

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

    from ha2g_b200 import graph_step
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
    # stream): one pass picks the dominant launcher, `steps` more passes time only that one
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
    kern_ms, kern_flops = gru_kernel_roofline(dev, a.batch) if rank == 0 else (None, None)
    if rank != 0:
        _finish_process(world)
        return

    frames = a.batch * T_FRAMES * world
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md 1.4 PF sustained)"
    # dominant kernel = the fused GRU recurrence (gru_seq_fwd_tc2_kernel): 36 launches at M = B and 24 at M = 2B per step
    achieved = kern_flops / (kern_ms * 1e-3) / 1e12
    peak_burst = peaks.get("bf16_tflops", 1590.0)   # the kernel is timed alone: burst figure (recipe fallback 1.59 PF)
    peak_burst_src = ("measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback (B200_PROFILING.md 1.59 PF burst)")
    calls_per_step = (len(gens) * 4 * 2 + 4 * 3) if a.epoch > args.loss_warmup else len(gens) * 4 * 2
    roofline = {"bound": "tensor", "kernel": "gru_seq_fwd_tc2_kernel (csrc/gru_cluster_tc2.cu), one bidirectional GRU layer, "
                                             f"M={a.batch} rows x T=34 x H=300",
                "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s", "frac": achieved / peak_burst,
                "traffic": 39.1e6, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r01_f_ncu_full.md "
                                                     "(algorithmic bytes per launch: 85.7e6)",
                "peak_source": peak_burst_src, "us_per_launch": kern_ms * 1e3, "flops_per_launch": kern_flops,
                "launches_per_step": calls_per_step,
                "note": "algorithmic fp32-equivalent FLOPs of the recurrence (2*M*T*2*3H*H; the kernel issues 3 bf16 MMAs per "
                        "product) / average CUDA-event time of 24 launches at the workload's shape, buffers rotated over 334 MB "
                        "(> L2); the recurrence is latency-bound by design (34 dependent steps per launch)"}
    top_launcher = None
    if topstat and topstat["calls"]:
        la = topstat["flops"] / (topstat["ms"] * 1e-3) / 1e12 if topstat["ms"] > 0 else 0.0
        top_launcher = {"launcher": top, "achieved_tflops": la, "frac": la / peak_tf, "calls_timed": topstat["calls"],
                        "ms_timed": topstat["ms"], "share_of_step": topstat["ms"] / a.steps / (ms / a.steps) if ms > 0 else None,
                        "note": f"C-ABI launcher with the largest CUDA-event time over {a.steps} eager passes of the same step "
                                "right after the timed region (the timed region replays one CUDA graph)"}
    line = {"metric": metric, "value": frames * a.steps / (ms * 1e-3), "unit": "pose-frames/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": frames * a.steps / (ms_e2e * 1e-3), "unit": "pose-frames/s",
                    "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": 4 * (len(ret) + len(gens) + 2) * world},
            "gpu_launches": launches, "cuda_graph": {"enabled": graph_step.enabled(), **graph_step.STATS},
            "clocks": clocks, "roofline": roofline, "top_launcher": top_launcher,
            "step_tflops": step_flops(a.variant, a.batch * world) * a.steps / (ms * 1e-3) / 1e12,
            "last_losses": {k: round(v, 5) for k, v in ret.items()},
            "profile_top5": sorted(((k, round(v["ms"], 3), v["calls"]) for k, v in prof.items()), key=lambda r: -r[1])[:5]}
    if not a.no_cpu_baseline and world == 1:
        # bounded sample in a subprocess (its own torch thread pool; killed if the host is too slow for the budget)
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--variant", a.variant,
                                  "--cpu-batch", str(a.cpu_batch), "--epoch", str(a.epoch), "--steps", "2"],
                                 capture_output=True, text=True, timeout=a.cpu_timeout)
            ref_line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
            line["cpu_baseline"] = ref_line["cpu_baseline"]
        except Exception as e:  # keep the GPU line even if the host is too small for the sample
            line["cpu_baseline"] = {"value": None, "unit": "pose-frames/s", "cores": min(os.cpu_count() or 1, 32),
                                    "kind": "port", "sample": f"failed: {type(e).__name__}"}
    print(json.dumps(line))
    _finish_process(world)


if __name__ == "__main__":
    main()
