"""Two-GPU NCCL check of the data-parallel pieces that cannot run on CPU: the global-batch contrastive loss
(all-gather of the columns, reduce-scatter of their gradients) against the single-process loss over the whole batch.
Skipped on a one-GPU box (the driver's `-m gpu` run); exercised with `gpurun --gpus 2`."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from ha2g_b200 import dp, ops_loss
    dp.enable(world)
    g = torch.Generator().manual_seed(7)
    Nl = 136
    a_full, b_full = torch.randn(world * Nl, 32, generator=g), torch.randn(world * Nl, 32, generator=g)
    a = a_full[rank * Nl:(rank + 1) * Nl].to(dev).requires_grad_(True)
    b = b_full[rank * Nl:(rank + 1) * Nl].to(dev).requires_grad_(True)
    loss = ops_loss.contrastive(a, b, "expressive")
    loss.backward()
    # reference: the same loss over the whole batch in one process (dp disabled -> square path)
    dp.disable()
    af, bf = a_full.to(dev).requires_grad_(True), b_full.to(dev).requires_grad_(True)
    full = ops_loss.contrastive(af, bf, "expressive")
    full.backward()
    lsum = loss.detach().clone()
    dist.all_reduce(lsum)
    sl = slice(rank * Nl, (rank + 1) * Nl)
    errs = (abs(float(lsum) / world - float(full)),
            float((a.grad / world - af.grad[sl]).abs().max() / af.grad.abs().max()),
            float((b.grad / world - bf.grad[sl]).abs().max() / bf.grad.abs().max()))
    q.put((rank, errs))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_global_contrastive_two_ranks_nccl():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, (el, ea, eb) in res:
        assert el <= 1e-5 and ea <= 1e-4 and eb <= 1e-4, (rank, el, ea, eb)


def _step_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from helpers import build_modules
    from ha2g_b200 import dp, graph_step
    from ha2g_b200.synthetic import make_batch
    from ha2g_b200.train_eval.train_hierarchy import train_iter_hierarchy
    args, gens, D, A, T = build_modules("gesture", 60, 5, {"gens": 20, "dis": 30, "audio": 31, "text": 32}, dev)
    mods = gens + [D, A, T]
    dp.enable(world, modules=mods)
    lr = args.learning_rate
    mk = lambda m, l=lr: torch.optim.Adam(m.parameters(), lr=l, betas=(0.5, 0.999))
    opts = [mk(g) for g in gens] + [mk(D, lr * args.discriminator_lr_weight), mk(A), mk(T)]
    torch.manual_seed(0)
    rets = []
    for i in range(5):   # 2 eager steps (the second one with hook-launched all-reduces), capture, 2 replays
        b = {k: v.to(dev) for k, v in make_batch("gesture", 8, 60, 5, seed=300 + 10 * i + rank).items()}
        rets.append(train_iter_hierarchy(args, 11, b["in_text_padded"], b["in_spec"], b["target"], b["vid"], *mods, *opts))
    torch.cuda.synchronize()
    flat = torch.cat([p.detach().double().reshape(-1) for m in mods for p in m.parameters()])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    q.put((rank, float((both[0] - both[1]).abs().max()), all(v == v for r in rets for v in r.values()),
           dict(graph_step.STATS), bool(dp._known)))
    graph_step.reset()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dp_training_steps_keep_replicas_identical_nccl():
    """Five data-parallel training steps on two GPUs (different clips per rank): eager steps with the all-reduces launched
    from gradient hooks during backward, then the captured CUDA graph replaying the same fork / join.  Every replica must
    hold bit-identical parameters afterwards, and the process group must shut down cleanly."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 2000
    procs = [ctx.Process(target=_step_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0, "a rank did not shut down cleanly"
    for rank, diff, finite, stats, known in res:
        assert diff == 0.0, (rank, diff)
        assert finite and known and stats["captures"] >= 1 and stats["replays"] >= 2, (rank, finite, known, stats)
