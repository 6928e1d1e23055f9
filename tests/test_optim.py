"""optim.FusedAdam: shard arithmetic on the CPU; on the GPU the fused kernel against torch.optim.Adam, and (two
GPUs) the sharded reduce + Adam + all-gather kernel against Adam on the averaged gradients."""
import os
import socket

import pytest
import torch


def test_shard_bounds_cover_and_align():
    from driving_dirty_b200.optim import shard_bounds
    for numel in (4, 8, 1000, 81_920_000, 240_648_192, 4 * 12345):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_bounds(numel, world, r)
                assert lo == prev and lo % 4 == 0 and hi % 4 == 0 and hi >= lo
                prev = hi
            assert prev == numel


def _reference_adam(params, grads_per_step, lr, wd=0.0):
    ps = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam(ps, lr=lr, weight_decay=wd)
    for grads in grads_per_step:
        for p, g in zip(ps, grads):
            p.grad = g.clone()
        opt.step()
    return [p.detach() for p in ps]


@pytest.mark.gpu
@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_matches_torch(wd):
    from driving_dirty_b200.optim import FusedAdam
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(5)
    shapes = [(7,), (32,), (1024,), (100003,), (256, 4100), (32, 32, 3, 3)]
    params = [torch.randn(s, generator=g).to(dev) for s in shapes]
    steps = [[(torch.randn(s, generator=g) * 0.1).to(dev) for s in shapes] for _ in range(5)]
    ref = _reference_adam(params, steps, lr=1e-2, wd=wd)
    ours = [torch.nn.Parameter(p.clone()) for p in params]
    opt = FusedAdam(ours, lr=1e-2, weight_decay=wd)
    for grads in steps:
        opt.zero_grad()
        for p, gr in zip(ours, grads):
            p.grad = gr.clone()
        opt.step()
    for a, b in zip(ours, ref):
        err = float((a.detach() - b).abs().max() / b.abs().max())
        assert err < 2e-6, err
    # the lr of a param group is read every step (ReduceLROnPlateau, roadmap_bce_v2.py:156)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=0, factor=0.5)
    sched.step(1.0); sched.step(2.0)
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_worker(rank, world, port, multicast, overlap, out):
    import torch.distributed as dist
    from driving_dirty_b200 import ops
    from driving_dirty_b200.optim import FusedAdam
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g = torch.Generator().manual_seed(11)
        N, K, B = 2048, 1024, 4                      # weight [N,K] = 2 Mi elements: sharded; bias: flat bucket
        w0, b0 = torch.randn(N, K, generator=g) * 0.05, torch.randn(N, generator=g) * 0.05
        xs = [[torch.randn(B, K, generator=g) for _ in range(world)] for _ in range(3)]
        gys = [[torch.randn(B, N, generator=g) for _ in range(world)] for _ in range(3)]
        w = torch.nn.Parameter(w0.clone().to(dev) + (0.0 if rank == 0 else 1.0))     # broadcast_init must repair this
        b = torch.nn.Parameter(b0.clone().to(dev))
        opt = FusedAdam([w, b], lr=1e-2, multicast=multicast, overlap_backward=overlap)
        assert bool(opt.uses_multicast) == bool(multicast)
        for step in range(3):
            opt.zero_grad()
            y = ops.linear(xs[step][rank].to(dev), w, b, impl=1)
            y.backward(gys[step][rank].to(dev))
            opt.step()
        torch.cuda.synchronize()
        # reference: Adam on the rank-averaged gradients (what ddp's all-reduce feeds the optimizer)
        steps = []
        for step in range(3):
            gw = sum(gys[step][r].t() @ xs[step][r] for r in range(world)) / world
            gb = sum(gys[step][r].sum(0) for r in range(world)) / world
            steps.append([gw, gb])
        ref = _reference_adam([w0, b0], steps, lr=1e-2)
        ew = float((w.detach().cpu() - ref[0]).abs().max() / ref[0].abs().max())
        eb = float((b.detach().cpu() - ref[1]).abs().max() / ref[1].abs().max())
        gathered = [torch.empty_like(w.data) for _ in range(world)]
        dist.all_gather(gathered, w.data.contiguous())
        same = all(torch.equal(gathered[0], t) for t in gathered)
        if rank == 0:
            with open(out, "w") as f:
                f.write(f"{ew} {eb} {int(same)}")
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("multicast,overlap", [(False, False), (True, False), (False, True), (True, True)])
def test_fused_adam_sharded_two_gpus(tmp_path, multicast, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.txt")
    mp.spawn(_sharded_worker, args=(2, _free_port(), multicast, overlap, out), nprocs=2, join=True)
    ew, eb, same = open(out).read().split()
    assert float(ew) < 5e-6 and float(eb) < 5e-6 and same == "1", (ew, eb, same)


def test_flat_layout_is_aligned_and_disjoint():
    from driving_dirty_b200.optim import flat_layout
    numels = [640000, 256, 65536, 7, 864, 9216, 32, 1]
    offs, total = flat_layout(numels)
    assert offs[0] == 0 and all(o % 64 == 0 for o in offs) and total % 64 == 0
    for (o, n), o_next in zip(zip(offs, numels), offs[1:] + [total]):
        assert o + n <= o_next
    assert total % 4 == 0          # shard_bounds works on 16-byte granules


@pytest.mark.gpu
def test_fused_adam_step_inside_the_weight_gradient_kernel():
    """World size 1: the wide FC weights take their Adam step in the epilogue of ops.linear's tcgen05 weight-gradient kernel
    (dd_linear_wgrad_adam; the gradient never reaches HBM).  Same weights, moments and bias as the unfused form
    (weight gradient written, then dd_adam_step) over several steps, with weight decay; p.grad stays None for them."""
    from driving_dirty_b200 import _lib, ops
    from driving_dirty_b200.optim import FusedAdam
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(21)
    N, K, B = 4096, 1028, 32                       # N*K >= 2^22: the tensor-core linear kernels apply
    w0, b0 = (torch.randn(N, K, generator=g) * 0.05).to(dev), (torch.randn(N, generator=g) * 0.05).to(dev)
    xs = [torch.randn(B, K, generator=g).to(dev) for _ in range(4)]
    gys = [torch.randn(B, N, generator=g).to(dev) for _ in range(4)]
    res = []
    for fuse in (False, True):
        w, b = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(b0.clone())
        opt = FusedAdam([w, b], lr=1e-2, weight_decay=0.01, fuse_into_backward=fuse)
        for x, gy in zip(xs, gys):
            opt.zero_grad()
            xin = x.clone().requires_grad_(True)
            y = ops.linear(xin, w, b, impl=_lib.IMPL_TCGEN05)
            y.backward(gy)
            assert (w.grad is None) == fuse and b.grad is not None and xin.grad is not None
            opt.step()
        st = opt.state[w]
        res.append((w.detach().clone(), b.detach().clone(), st["exp_avg"].clone(), st["exp_avg_sq"].clone(), st["step"], xin.grad.clone()))
    for a, c in zip(res[0][:4], res[1][:4]):
        assert float((a.reshape(-1) - c.reshape(-1)).abs().max() / a.abs().max()) < 1e-6
    assert res[0][4] == res[1][4] == 4
    assert torch.equal(res[0][5], res[1][5])      # the input gradient used the weight BEFORE its update in both forms
    # checkpoint layout is unchanged
    sd = opt.state_dict()
    assert tuple(sd["state"][0]["exp_avg"].shape) == (N, K) and int(sd["state"][0]["step"]) == 4
