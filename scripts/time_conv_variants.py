import os, sys, subprocess
if len(sys.argv) == 1:
    for dbg in [int(a) for a in os.environ.get('DBGS', '0 1 2 3 4 5 6 7').split()]:
        env = dict(os.environ, DD_CONV_DBG=str(dbg))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("dbg", dbg, "(1 no MMA | 2 no stores | 4 no loads):", " || ".join(out.stdout.strip().splitlines()[-2:]) if out.stdout.strip() else out.stderr[-300:])
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

import ctypes
lib = _lib.load()
if hasattr(lib, "dd_debug_s1_prof"):
    buf = (ctypes.c_longlong * (148 * 16))()
    lib.dd_debug_s1_prof(buf)
    import statistics
    cols = list(zip(*[buf[i * 16:(i + 1) * 16] for i in range(148)]))
    med = [statistics.median(c) for c in cols]
    n = med[5] or 1
    print("PROF per input row (cycles, median CTA): waits %.0f (full %.0f, acc %.0f; blocking full %d acc %d of %d each) issue+commit %.0f | total %.0f | rows %d || epi(one warp, per its row): wait %.0f ld %.0f store %.0f total %.0f" % (
        med[0] / n, med[2] / n, med[6] / n, med[3] % 1000000, med[3] // 1000000, n // 4, med[1] / n, med[4] / n, n, med[8] / (n / 2), med[9] / (n / 2), med[10] / (n / 2), med[11] / (n / 2)))
