"""world_size-2 gloo run of the gradient exchange (host logic of the N>1 path) on CPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from driving_dirty_b200.distributed import GradAllReducer
    torch.manual_seed(0)
    big = torch.nn.Parameter(torch.zeros(1 << 12))
    small = [torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(3, 5))]
    red = GradAllReducer([big] + small, big_numel=1 << 10)
    loss = (big * (rank + 1)).sum() + sum((s * (10 * (rank + 1))).sum() for s in small)
    loss.backward()
    red.finish()
    ok = bool(torch.allclose(big.grad, torch.full_like(big, 1.5)) and
              all(torch.allclose(s.grad, torch.full_like(s, 15.0)) for s in small))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_grad_allreduce_mean_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
