"""world_size-2 gloo test of the data-parallel gradient averaging (ha2g_b200/dp.py) on CPU: after
allreduce_grads every rank holds the mean gradient, and stepping a torch Adam on it keeps replicas identical."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ha2g_b200 import dp
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    if rank == 1:  # start from different weights: enable() must broadcast rank 0's
        for p in model.parameters():
            p.data.add_(1.0)
    dp.enable(world, modules=[model])
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(4, 7, generator=g)
    model(x).pow(2).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    dp.BUCKET_BYTES = 64  # force several buckets
    dp.allreduce_grads(opt)
    opt.step()
    gathered = [torch.zeros_like(torch.cat([l.reshape(-1) for l in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([l.reshape(-1) for l in local]))
    mean = sum(gathered) / world
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    params = torch.cat([p.data.reshape(-1) for p in model.parameters()])
    allp = [torch.zeros_like(params) for _ in range(world)]
    dist.all_gather(allp, params)
    q.put((rank, float((got - mean).abs().max()), float((allp[0] - allp[1]).abs().max())))
    dist.destroy_process_group()


def test_allreduce_grads_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gerr, perr in res:
        assert gerr < 1e-6, (rank, gerr)
        assert perr == 0.0, (rank, perr)
