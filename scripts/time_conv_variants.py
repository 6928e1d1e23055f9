import os, sys, subprocess
if len(sys.argv) == 1:
    for dbg in (0, 1, 2, 3, 4, 5, 6, 7):
        env = dict(os.environ, DD_CONV_DBG=str(dbg))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("dbg", dbg, "(1 no MMA | 2 no stores | 4 no loads):", out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
    sys.exit(0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from driving_dirty_b200 import _lib
from driving_dirty_b200._lib import call, stream_ptr
B, H, W = 32, 256, 1836
dev = torch.device("cuda")
x = torch.rand(B, H, W, 32, device=dev).bfloat16()
out = torch.empty_like(x)
w = torch.rand(32, 32, 3, 3, device=dev) * 0.1
b = torch.zeros(32, device=dev)
st = stream_ptr()
def run():
    call("dd_conv3x3_c32_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 1, B, H, W, 1, 0, st)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1)/10:.4f} ms")
