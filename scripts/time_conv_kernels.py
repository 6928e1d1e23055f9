"""Times every tcgen05 conv kernel of the encoder at the bench shape with CUDA events (one line each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from driving_dirty_b200 import _lib
from driving_dirty_b200._lib import call, stream_ptr
B, H, W = int(os.environ.get("B", 32)), 256, 1836
dev = torch.device("cuda")
views = torch.rand(B, 6, 3, 256, 306, device=dev)
views_u8 = torch.randint(0, 256, (B, 6, 3, 256, 306), device=dev, dtype=torch.uint8)
x = torch.rand(B, H, W, 32, device=dev).bfloat16()
dy = (torch.rand(B, H, W, 32, device=dev) - 0.5).bfloat16()
out = torch.empty_like(x)
H3, W3 = 128, 918
dy3 = (torch.rand(B, H3, W3, 32, device=dev) - 0.5).bfloat16()
out3 = torch.empty_like(dy3)
w = torch.rand(32, 32, 3, 3, device=dev) * 0.1
w1 = torch.rand(32, 3, 3, 3, device=dev) * 0.1
b = torch.zeros(32, device=dev)
dw, db, dw1 = torch.empty_like(w), torch.empty_like(b), torch.empty_like(w1)
n = int(_lib.load().dd_conv_wgrad_workspace_bytes())
ws = torch.empty(n, dtype=torch.uint8, device=dev)
st = stream_ptr()
GB = 1e9
act = x.numel() * 2
runs = {
    "c1 fwd   (views -> a1)": (lambda: call("dd_conv_c1_fwd", views.data_ptr(), 1, w1.data_ptr(), b.data_ptr(), out.data_ptr(), 1, B, H, W, 0, st), views.numel() * 4 + act),
    "c1 fwd   (u8 views)": (lambda: call("dd_conv_c1_fwd", views_u8.data_ptr(), 3, w1.data_ptr(), b.data_ptr(), out.data_ptr(), 1, B, H, W, 0, st), views.numel() + act),
    "c2 fwd   (s1)": (lambda: call("dd_conv3x3_c32_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 1, B, H, W, 1, 0, st), 2 * act),
    "c3 fwd   (s2)": (lambda: call("dd_conv3x3_c32_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), out3.data_ptr(), 1, B, H, W, 2, 0, st), act + act // 4),
    "c3 dgrad (s2)": (lambda: call("dd_conv3x3_c32_dgrad", dy3.data_ptr(), w.data_ptr(), x.data_ptr(), out.data_ptr(), 1, B, H, W, 2, 0, st), 2 * act + act // 4),
    "c3 wgrad (s2)": (lambda: call("dd_conv3x3_c32_wgrad", x.data_ptr(), dy3.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), n, 1, B, H, W, 2, 0, st), act + act // 4),
    "c2 dgrad (s1)": (lambda: call("dd_conv3x3_c32_dgrad", dy.data_ptr(), w.data_ptr(), x.data_ptr(), out.data_ptr(), 1, B, H, W, 1, 0, st), 3 * act),
    "c2 wgrad (s1)": (lambda: call("dd_conv3x3_c32_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), n, 1, B, H, W, 1, 0, st), 2 * act),
    "c1 wgrad": (lambda: call("dd_conv_c1_wgrad", views.data_ptr(), 1, dy.data_ptr(), 1, dw1.data_ptr(), db.data_ptr(), ws.data_ptr(), n, B, H, W, 0, st), views.numel() * 4 + act),
    "c1 wgrad (u8 views)": (lambda: call("dd_conv_c1_wgrad", views_u8.data_ptr(), 3, dy.data_ptr(), 1, dw1.data_ptr(), db.data_ptr(), ws.data_ptr(), n, B, H, W, 0, st), views.numel() + act),
}
tot = 0.0
for name, (fn, nbytes) in runs.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tot += ms
    print(f"{name:24s} {ms:7.4f} ms   {nbytes / ms / 1e6:7.1f} GB/s ({nbytes / ms / 1e6 / 6553 * 100:4.1f}% of 6553)   HBM floor {nbytes / 6553e6:6.4f} ms")
print(f"total {tot:.3f} ms")
