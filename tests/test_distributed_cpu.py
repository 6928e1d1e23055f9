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


# ------------------------------------------------------------------------------------------------
# optim.FusedAdam's data-parallel form: the sharding logic (regions, flat bucket, shard bounds, per-shard state,
# written / launched bookkeeping) runs unchanged; only the three device touch points are replaced -- the symmetric
# buffer becomes a plain CPU tensor with a gloo barrier, and the reduce + Adam + all-gather KERNEL is emulated with
# gloo collectives and torch arithmetic (test infrastructure: the product launch needs CUDA).
# ------------------------------------------------------------------------------------------------
def _adam_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from driving_dirty_b200.optim import FusedAdam, shard_bounds

    class _Handle:
        def barrier(self, channel=0):
            dist.barrier()

    class GlooAdam(FusedAdam):
        def _check_param(self, p):
            assert p.dtype == torch.float32 and p.is_contiguous()

        def _alloc_symmetric(self, numel, device, multicast):
            return torch.empty(numel, dtype=torch.float32), _Handle(), [0] * self.world, 0

        def _alloc_moments(self, numel, device):
            buf = torch.zeros(numel, dtype=torch.float32)

            def peer(rk):          # the product reads a peer-mapped buffer; here every rank calls state_dict(): broadcast
                t = buf.clone()
                dist.broadcast(t, src=rk)
                return t
            return buf, peer

        def _launch_sharded(self, r, ctas_per_sm):
            sy = self._symm
            buf, total, off, lo, hi = sy["buf"], sy["total"], r["off"], r["lo"], r["hi"]
            # 1. reduce: every rank's gradient replica of MY shard
            mine = buf[off + lo: off + hi].clone()
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            for rk in range(self.world):           # rank rk's shard may have another length: one gather per owner
                lo_k, hi_k = shard_bounds(r["n"], self.world, rk)
                piece = buf[off + lo_k: off + hi_k].clone()
                got = [torch.empty_like(piece) for _ in range(self.world)] if rk == self.rank else None
                dist.gather(piece, got, dst=rk)
                if rk == self.rank:
                    parts = got
            g = sum(parts) / self.world
            # 2. Adam on the shard (torch.optim.Adam's arithmetic)
            st = self._shard_state(r)
            st["step"] += 1
            lr, b1, b2, eps, wd = self._hyper()
            w = buf[total + off + lo: total + off + hi]
            if wd:
                g = g + wd * w
            st["exp_avg"].lerp_(g, 1 - b1)
            st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1, bc2 = 1 - b1 ** st["step"], 1 - b2 ** st["step"]
            w.sub_((lr / bc1) * st["exp_avg"] / (st["exp_avg_sq"].sqrt() / bc2 ** 0.5 + eps))
            # 3. all-gather: every owner's new shard into every replica
            for rk in range(self.world):
                lo_k, hi_k = shard_bounds(r["n"], self.world, rk)
                dist.broadcast(buf[total + off + lo_k: total + off + hi_k], src=rk)

        def _launch_local(self, g, p, grad):
            st = self._state(p, p.numel(), p.device)
            st["step"] += 1
            lr, b1, b2, eps, wd = self._hyper()
            gr = grad + wd * p.data if wd else grad
            st["exp_avg"].lerp_(gr.reshape(-1), 1 - b1)
            st["exp_avg_sq"].mul_(b2).addcmul_(gr.reshape(-1), gr.reshape(-1), value=1 - b2)
            bc1, bc2 = 1 - b1 ** st["step"], 1 - b2 ** st["step"]
            p.data.sub_(((lr / bc1) * st["exp_avg"] / (st["exp_avg_sq"].sqrt() / bc2 ** 0.5 + eps)).view_as(p))

    g = torch.Generator().manual_seed(3)
    shapes = [(64, 66), (7,), (3, 5), (130,)]                 # 4224 elements: sharded on its own; the rest: flat bucket
    init = [torch.randn(s, generator=g) for s in shapes]
    grads = [[[torch.randn(s, generator=g) for s in shapes] for _ in range(world)] for _ in range(3)]
    params = [torch.nn.Parameter(t.clone() + (0.0 if rank == 0 else 5.0)) for t in init]     # broadcast_init repairs rank 1
    opt = GlooAdam(params, lr=1e-2, weight_decay=0.01, shard_min_numel=1024)
    assert len(opt._regions) == 1 and opt._flat is not None and len(opt._flat["params"]) == 3
    for step in range(3):
        opt.zero_grad()
        assert params[0].grad is not None                      # the persistent view survives zero_grad
        params[0]._dd_grad_buffer.copy_(grads[step][rank][0])  # what ops.linear's weight-gradient kernel does ...
        params[0]._dd_grad_ready()                             # ... and then signals
        for p, gr in zip(params[1:], grads[step][rank][1:]):
            p.grad = gr.clone()
        opt.step()
    ref = [t.clone().requires_grad_(True) for t in init]
    ropt = torch.optim.Adam(ref, lr=1e-2, weight_decay=0.01)
    for step in range(3):
        for i, p in enumerate(ref):
            p.grad = sum(grads[step][r][i] for r in range(world)) / world
        ropt.step()
    err = max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(params, ref))
    # ---- checkpoint round trip: torch.optim.Adam's layout out and in, at world size 2
    sd = opt.state_dict()                                      # full, parameter-shaped moments on every rank
    rsd = ropt.state_dict()
    sd_err = max(float((sd["state"][i][k] - rsd["state"][i][k]).abs().max()) for i in range(len(shapes))
                 for k in ("exp_avg", "exp_avg_sq"))
    steps_ok = all(int(sd["state"][i]["step"]) == 3 for i in range(len(shapes)))
    torch.optim.Adam([t.clone().requires_grad_(True) for t in init], lr=1e-2).load_state_dict(sd)   # torch accepts it
    params_b = [torch.nn.Parameter(p.detach().clone()) for p in params]
    opt_b = GlooAdam(params_b, lr=1e-2, weight_decay=0.01, shard_min_numel=1024)
    opt_b.load_state_dict(rsd)                                 # ... and we accept torch's (tensor `step`, full moments)
    g4 = [[torch.randn(s, generator=g) for s in shapes] for _ in range(world)]
    opt_b.zero_grad()
    params_b[0]._dd_grad_buffer.copy_(g4[rank][0])
    params_b[0]._dd_grad_ready()
    for p, gr in zip(params_b[1:], g4[rank][1:]):
        p.grad = gr.clone()
    opt_b.step()
    for i, p in enumerate(ref):
        p.grad = sum(g4[r][i] for r in range(world)) / world
    ropt.step()
    err_resume = max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(params_b, ref))
    ckpt_ok = sd_err < 1e-6 and steps_ok and err_resume < 1e-6
    # a second backward without a step must be refused (the gradient buffer is overwritten, not accumulated)
    params[0]._dd_grad_ready()
    try:
        params[0]._dd_grad_ready()
        refused = False
    except RuntimeError:
        refused = True
    opt.step()
    # ---- a partly frozen flat bucket (RoadMapBCE before unfreeze_epoch_no): per-tensor replicated updates, torch's skip rule
    params2 = [torch.nn.Parameter(t.clone()) for t in init]
    opt2 = GlooAdam(params2, lr=1e-2, shard_min_numel=1024)
    ref2 = [t.clone().requires_grad_(True) for t in init]
    ropt2 = torch.optim.Adam(ref2, lr=1e-2)
    for step in range(3):
        frozen = {3} if step < 2 else set()                    # tensor 3 gets its first gradient at the third step
        opt2.zero_grad()
        params2[0]._dd_grad_buffer.copy_(grads[step][rank][0])
        params2[0]._dd_grad_ready()
        for i in (1, 2, 3):
            params2[i].grad = None if i in frozen else grads[step][rank][i].clone()
            ref2[i].grad = None if i in frozen else sum(grads[step][r][i] for r in range(world)) / world
        ref2[0].grad = sum(grads[step][r][0] for r in range(world)) / world
        opt2.step()
        ropt2.step()
    err2 = max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(params2, ref2))
    # the partly frozen history (per-tensor step counts) survives a save / load as well
    sd2 = opt2.state_dict()
    steps2 = [int(sd2["state"][i]["step"]) for i in range(4)]
    params3 = [torch.nn.Parameter(p.detach().clone()) for p in params2]
    opt3 = GlooAdam(params3, lr=1e-2, shard_min_numel=1024)
    opt3.load_state_dict(sd2)
    ckpt_ok = ckpt_ok and steps2 == [3, 3, 3, 1] and opt3._flat_replicated
    q.put((rank, err < 1e-6 and refused and err2 < 1e-6 and opt2._flat_replicated and ckpt_ok,
           (err, err2, sd_err, err_resume, steps2)))
    dist.destroy_process_group()


def test_fused_adam_sharding_logic_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_adam_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)], res
