"""CUDA-event time of the fused inference encoder front (c1 -> c2 in one kernel) against the two kernels it replaces,
at the bench shape (B scenes of six 3 x 256 x 306 views), with the tensor-pipe fraction the convs' flops amount to."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from driving_dirty_b200._lib import call, dtype_code, stream_ptr
B, H, W = int(os.environ.get("B", "32")), 256, 306
Wm = 6 * W
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(1)
views = torch.rand(B, 6, 3, H, W, generator=g).to(dev)
w1 = ((torch.rand(32, 3, 3, 3, generator=g) * 2 - 1) / 27 ** 0.5).to(dev)
w2 = ((torch.rand(32, 32, 3, 3, generator=g) * 2 - 1) / 288 ** 0.5).to(dev)
b1 = ((torch.rand(32, generator=g) * 2 - 1) * 0.1).to(dev)
b2 = ((torch.rand(32, generator=g) * 2 - 1) * 0.1).to(dev)
code, st = dtype_code(torch.bfloat16), stream_ptr()
a1 = torch.empty(B, H, Wm, 32, dtype=torch.bfloat16, device=dev)
a2 = torch.empty_like(a1)
a2f = torch.empty_like(a1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def two():
    call("dd_conv_c1_fwd", views.data_ptr(), 1, w1.data_ptr(), b1.data_ptr(), a1.data_ptr(), code, B, H, Wm, 2, st)
    call("dd_conv3x3_c32_fwd", a1.data_ptr(), w2.data_ptr(), b2.data_ptr(), a2.data_ptr(), code, B, H, Wm, 1, 2, st)

def fused():
    call("dd_encoder_c1c2_fused_fwd", views.data_ptr(), 1, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
         a2f.data_ptr(), B, H, Wm, st)

def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

raw = (views * 255).round().to(torch.uint8)
viewsq = raw.float() / 255
def fused_u8():
    call("dd_encoder_c1c2_fused_fwd", raw.data_ptr(), 3, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
         a2f.data_ptr(), B, H, Wm, st)
t8 = timeit(fused_u8)
print(f"fused from raw bytes {t8:.4f} ms")
t2, tf = timeit(two), timeit(fused)
same = torch.equal(a2.view(torch.int16), a2f.view(torch.int16))
flops = 2.0 * B * H * Wm * 32 * (27 + 288)
print(f"B={B}: c1 + c2 kernels {t2:.4f} ms, fused {tf:.4f} ms ({t2 / tf:.2f}x); bit-identical: {same}; "
      f"fused = {flops / tf / 1e9:.1f} TFLOP/s of conv arithmetic")
