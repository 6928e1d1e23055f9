"""Runs the c2/c3-shaped tcgen05 conv kernels a few times at the bench shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from driving_dirty_b200 import _lib
from driving_dirty_b200._lib import call, stream_ptr
B, H, W = int(os.environ.get("B", 32)), 256, 1836
dev = torch.device("cuda")
x = torch.rand(B, H, W, 32, device=dev).bfloat16()
dy = (torch.rand(B, H, W, 32, device=dev) - 0.5).bfloat16()
out = torch.empty_like(x)
w = torch.rand(32, 32, 3, 3, device=dev) * 0.1
b = torch.zeros(32, device=dev)
dw, db = torch.empty_like(w), torch.empty_like(b)
n = int(_lib.load().dd_conv_wgrad_workspace_bytes())
ws = torch.empty(n, dtype=torch.uint8, device=dev)
st = stream_ptr()
for _ in range(int(os.environ.get("REPS", 3))):
    call("dd_conv3x3_c32_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 1, B, H, W, 1, 0, st)
    call("dd_conv3x3_c32_dgrad", dy.data_ptr(), w.data_ptr(), x.data_ptr(), out.data_ptr(), 1, B, H, W, 1, 0, st)
    call("dd_conv3x3_c32_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), n, 1, B, H, W, 1, 0, st)
torch.cuda.synchronize()
print("ok")
