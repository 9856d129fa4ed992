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
