"""Times the HBM-bound kernels (stitch, fused BCE/threat score, BCE backward, pool) at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from driving_dirty_b200 import ops
B = 32
dev = torch.device("cuda")
views = torch.rand(B, 6, 3, 256, 306, device=dev)
logits = torch.randn(B, 800, 800, device=dev) * 0.06
target = (torch.rand(B, 800, 800, device=dev) > 0.5).float()
target_u8 = (target > 0).to(torch.uint8)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
rows = [
    ("stitch f32", lambda: ops.stitch(views), 2 * views.numel() * 4),
    ("bce+ts fwd f32 target, sums only", lambda: ops.bce_threat(logits, target, want_probs=False, want_binary=False), 2 * logits.numel() * 4),
    ("bce+ts fwd u8 target, sums only", lambda: ops.bce_threat(logits, target_u8, want_probs=False, want_binary=False), logits.numel() * 5),
    ("bce+ts fwd f32 target + probs + binary", lambda: ops.bce_threat(logits, target, want_probs=True, want_binary=True), logits.numel() * 13),
]
for name, fn, nbytes in rows:
    ms = t(fn)
    print(f"{name:42s} {ms:7.4f} ms  {nbytes / ms / 1e6:7.1f} GB/s ({nbytes / ms / 1e6 / 6553 * 100:4.1f}% of 6553)")
