"""Launches every hand-written kernel of libdd_b200.so REPS times at its bench shape (for ncu captures and sanitizer runs).
SMALL=1 shrinks the shapes (compute-sanitizer runs 10-50x slower)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from driving_dirty_b200 import _lib, ops
from driving_dirty_b200._lib import call, stream_ptr
SMALL = bool(os.environ.get("SMALL"))
REPS = int(os.environ.get("REPS", 2))
B = 2 if SMALL else 32
VH, VW = (32, 50) if SMALL else (256, 306)
H, W = VH, 6 * VW
H3, W3 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
MAP = 96 if SMALL else 800
dev = torch.device("cuda")
st = stream_ptr()
lib = _lib.load()
views = torch.rand(B, 6, 3, VH, VW, device=dev)
views_u8 = torch.randint(0, 256, (B, 6, 3, VH, VW), device=dev, dtype=torch.uint8)
x = torch.rand(B, H, W, 32, device=dev).bfloat16()
dy = (torch.rand(B, H, W, 32, device=dev) - 0.5).bfloat16()
out = torch.empty_like(x)
dy3 = (torch.rand(B, H3, W3, 32, device=dev) - 0.5).bfloat16()
out3 = torch.empty_like(dy3)
w = torch.rand(32, 32, 3, 3, device=dev) * 0.1
w1 = torch.rand(32, 3, 3, 3, device=dev) * 0.1
b = torch.zeros(32, device=dev)
dw, db, dw1 = torch.empty_like(w), torch.empty_like(b), torch.empty_like(w1)
n = int(lib.dd_conv_wgrad_workspace_bytes())
ws = torch.empty(n, dtype=torch.uint8, device=dev)
nfeat = 8 * H3 * W3
pooled = torch.empty(B, nfeat, device=dev, dtype=torch.bfloat16)
dpooled = torch.rand(B, nfeat, device=dev).bfloat16()
logits = torch.randn(B, MAP, MAP, device=dev) * 0.06
target = (torch.rand(B, MAP, MAP, device=dev) > 0.5).float()
HID, LAT = (16, 8) if SMALL else (256, 128)
xf = torch.rand(B, nfeat, device=dev)
wf = (torch.rand(HID, nfeat, device=dev) - 0.5) * 0.01
bf = torch.zeros(HID, device=dev)
z = torch.rand(B, LAT, device=dev)
wh = (torch.rand(MAP * MAP, LAT, device=dev) - 0.5) * 0.1
bh = torch.zeros(MAP * MAP, device=dev)
p = torch.rand(1 << (16 if SMALL else 26), device=dev)
g = torch.rand_like(p); m = torch.zeros_like(p); v = torch.zeros_like(p)
# one wide dilated layer of the merging CNN
cin, cout, k, d = 96, 64, 7, 7
Hi = 24 if SMALL else 256
Ho = Hi + d * (k - 1)
desc = _lib.ConvDesc(B, cin, cout, Hi, Hi, Ho, Ho, k, k, 1, 1, 0, 0, d, d, 1)
ux = torch.rand(B, Hi, Hi, cin, device=dev).bfloat16()
uy = torch.empty(B, Ho, Ho, cout, device=dev, dtype=torch.bfloat16)
udy = (torch.rand(B, Ho, Ho, cout, device=dev) - 0.5).bfloat16()
udx = torch.empty_like(ux)
uw = (torch.rand(cin, cout, k, k, device=dev) - 0.5) * 0.05
ub = torch.zeros(cout, device=dev)
udw, udb = torch.empty_like(uw), torch.empty_like(ub)
un = int(lib.dd_conv2d_workspace_bytes(ctypes.byref(desc)))
uws = torch.empty(un, dtype=torch.uint8, device=dev)
boxes = torch.rand(20, 2, 4, device=dev) * 10
for _ in range(REPS):
    ops.stitch(views)
    ops.bytes_to_float(views_u8)
    call("dd_conv_c1_fwd", views.data_ptr(), 1, w1.data_ptr(), b.data_ptr(), out.data_ptr(), 1, B, H, W, 0, st)
    call("dd_conv3x3_c32_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 1, B, H, W, 1, 0, st)
    call("dd_conv3x3_c32_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), out3.data_ptr(), 1, B, H, W, 2, 0, st)
    call("dd_pool4_fwd", out3.data_ptr(), pooled.data_ptr(), 1, B, H3, W3, st)
    call("dd_pool4_bwd", out3.data_ptr(), dpooled.data_ptr(), dy3.data_ptr(), 1, B, H3, W3, st)
    call("dd_conv3x3_c32_wgrad", x.data_ptr(), dy3.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), n, 1, B, H, W, 2, 0, st)
    call("dd_conv3x3_c32_dgrad", dy3.data_ptr(), w.data_ptr(), x.data_ptr(), out.data_ptr(), 1, B, H, W, 2, 0, st)
    call("dd_conv3x3_c32_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), n, 1, B, H, W, 1, 0, st)
    call("dd_conv3x3_c32_dgrad", dy.data_ptr(), w.data_ptr(), x.data_ptr(), out.data_ptr(), 1, B, H, W, 1, 0, st)
    call("dd_conv_c1_wgrad", views.data_ptr(), 1, dy.data_ptr(), 1, dw1.data_ptr(), db.data_ptr(), ws.data_ptr(), n, B, H, W, 0, st)
    y1 = ops.linear(xf, wf.requires_grad_(True), bf, _lib.IMPL_TCGEN05 if not SMALL else _lib.IMPL_SIMT)
    y1.backward(torch.ones_like(y1))
    y2 = ops.linear(z.requires_grad_(True), wh.requires_grad_(True), bh, _lib.IMPL_TCGEN05 if not SMALL else _lib.IMPL_SIMT)
    y2.backward(torch.ones_like(y2))
    wf.grad = wh.grad = None
    lg = logits.clone().requires_grad_(True)
    loss, *_ = ops.bce_threat(lg, target, want_probs=True, want_binary=True)
    loss.backward()
    ops.bce_threat(logits, target, want_probs=False, want_binary=False)
    ops.binary_map(logits)
    call("dd_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, st)
    call("dd_conv2d_fwd", ux.data_ptr(), uw.data_ptr(), ub.data_ptr(), uy.data_ptr(), ctypes.byref(desc), 1, 1, uws.data_ptr(), un, st)
    call("dd_conv2d_dgrad", udy.data_ptr(), uw.data_ptr(), ux.data_ptr(), udx.data_ptr(), ctypes.byref(desc), 1, uws.data_ptr(), un, st)
    call("dd_conv2d_wgrad", ux.data_ptr(), udy.data_ptr(), udw.data_ptr(), udb.data_ptr(), ctypes.byref(desc), 1, uws.data_ptr(), un, st)
    from driving_dirty_b200.utils.helper import compute_ats_bounding_boxes
    compute_ats_bounding_boxes(boxes, boxes + 0.1)
torch.cuda.synchronize()
print("ok")
