"""Times optim.FusedAdam.step() on the RoadMapBCE parameter set (world 1 or under torchrun), with CUDA events:
whole step, and the pieces of the sharded path (barrier, kernels).  MULTICAST=0/1 selects peer loads vs multimem."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
from driving_dirty_b200.optim import FusedAdam
from driving_dirty_b200._lib import call, stream_ptr
shapes = [(640000, 128), (256, 940032), (640000,), (256,), (256, 256), (128, 256), (32, 32, 3, 3), (32, 32, 3, 3), (32, 3, 3, 3)] + [(256,)] * 8 + [(32,)] * 3
params = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.01) for s in shapes]
mc = os.environ.get("MULTICAST")
opt = FusedAdam(params, lr=1e-3, multicast=None if mc is None else bool(int(mc)))
for p in params:
    if p.grad is None:
        p.grad = torch.randn_like(p) * 0.01
    else:
        p.grad.normal_(0, 0.01)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ctas = int(os.environ.get("CTAS", 0))
if ctas and world > 1:
    orig = opt._launch_sharded
    opt._launch_sharded = lambda info, ctas_per_sm: orig(info, ctas)
def full_step():
    opt._written.update(opt._regions.keys())     # as if this backward pass had written every wide gradient
    opt.step()
t_step = timeit(full_step)
msg = f"rank {rank}/{world} multicast={opt.uses_multicast} ctas/SM={ctas or 8}: opt.step {t_step:.3f} ms"
if world > 1:
    sy = opt._symm
    t_bar = timeit(lambda: sy["hdl"].barrier())
    msg += f" | symm barrier {t_bar:.3f} ms"
    msg += f" | flat bucket only {timeit(opt.step):.3f} ms"
else:
    ref = [torch.nn.Parameter(p.detach().clone()) for p in params]
    for a, b in zip(ref, params): a.grad = b.grad.clone()
    topt = torch.optim.Adam(ref, lr=1e-3, fused=True)
    msg += f" | torch fused Adam {timeit(topt.step):.3f} ms"
nbytes = sum(p.numel() for p in params) * 28
msg += f" | local-form traffic {nbytes/1e9:.2f} GB -> {nbytes / t_step / 1e6:.0f} GB/s if world 1"
print(msg, flush=True)
if world > 1: dist.destroy_process_group()
