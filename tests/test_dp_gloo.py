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
    dp.allreduce_grads(opt)
    opt.step()
    gathered = [torch.zeros_like(torch.cat([l.reshape(-1) for l in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([l.reshape(-1) for l in local]))
    mean = sum(gathered) / world
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    params = torch.cat([p.data.reshape(-1) for p in model.parameters()])
    allp = [torch.zeros_like(params) for _ in range(world)]
    dist.all_gather(allp, params)
    # second step: the gradient set is known now, so the all-reduce is launched from the autograd hooks during backward
    opt.zero_grad(set_to_none=True)
    x2 = torch.randn(4, 7, generator=g)
    dp.begin_backward([opt])
    model(x2).pow(2).sum().backward()
    dp.end_backward()
    hooked = id(opt) in dp._inflight
    local2 = torch.cat([p.grad.reshape(-1) for p in model.parameters()])   # gloo path: already averaged in place
    dp.allreduce_grads(opt)
    got2 = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    both = [torch.zeros_like(got2) for _ in range(world)]
    dist.all_gather(both, got2)
    q.put((rank, float((got - mean).abs().max()), float((allp[0] - allp[1]).abs().max()), hooked,
           float((both[0] - both[1]).abs().max()), float((local2 - got2).abs().max())))
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
    for rank, gerr, perr, hooked, g2_rank_diff, g2_changed in res:
        assert gerr < 1e-6, (rank, gerr)
        assert perr == 0.0, (rank, perr)
        assert hooked, "the second backward did not launch the all-reduce from the gradient hooks"
        assert g2_rank_diff == 0.0 and g2_changed == 0.0, (rank, g2_rank_diff, g2_changed)
