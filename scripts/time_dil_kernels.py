"""Times the wide stride-1 conv / transposed-conv layers of configs 3 and 4 (dd_conv2d_fwd / dgrad / wgrad, bf16) at full size
with CUDA events: ms, TFLOP/s of the gather formulation (2 * out pixels * N * K * taps) and the fraction of the measured bf16 peak."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from driving_dirty_b200 import _lib
from driving_dirty_b200._lib import call, stream_ptr
B = int(os.environ.get("B", 32))
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    PEAK = 1590.0
def _p(v):
    return v if isinstance(v, tuple) else (v, v)


ALL_BB = [  # the remaining (CUDA-core) layers of config 4: name, transposed, cin, cout, k, stride, pad, dil, Hi, Wi
    ("strip conv 3->32 1x50 s(3,2)", 0, 3, 32, (1, 50), (3, 2), 0, 1, 256, 306),
    ("strip conv 3->32 52x1 s(3,2) p1", 0, 3, 32, (52, 1), (3, 2), 1, 1, 306, 256),
    ("ss_conv 32->32 1x24 s(1,7)", 0, 32, 32, (1, 24), (1, 7), 0, 1, 128, 918),
    ("ss_deconv ConvT 32->32 k2 s2", 1, 32, 32, 2, 2, 0, 1, 128, 128),
    ("rm_conv_1 1->32 k7 s3 d3 p1", 0, 1, 32, 7, 3, 1, 3, 800, 800),
    ("up_conv_4 ConvT 16->8 k7 d3", 1, 16, 8, 7, 1, 0, 3, 382, 382),
    ("up_conv_5 ConvT 8->1 k2 s2", 1, 8, 1, 2, 2, 0, 1, 400, 400),
]
LAYERS = [  # name, transposed, cin, cout, k, pad, dil, Hi, Wi
    ("up_conv_1 ConvT 96->64 k7 d7", 1, 96, 64, 7, 0, 7, 256, 256),
    ("up_conv_2 ConvT 64->32 k7 d7", 1, 64, 32, 7, 0, 7, 298, 298),
    ("up_conv_3 ConvT 32->16 k7 d7", 1, 32, 16, 7, 0, 7, 340, 340),
    ("rm_conv_2 Conv 32->32 k3 d3", 0, 32, 32, 3, 0, 3, 262, 262),
    ("out_conv  Conv 32->32 k3", 0, 32, 32, 3, 0, 1, 258, 258),
    ("dc1 ConvT 64->32 k3 p1 (B=64)", 1, 64, 32, 3, 1, 1, 128, 153),
    ("dc2 ConvT 32->32 k3 p1 (B=64)", 1, 32, 32, 3, 1, 1, 128, 153),
]
dev = torch.device("cuda")
st = stream_ptr()
LAYERS = [(n, t, ci, co, k, 1, p, d, h, w) for (n, t, ci, co, k, p, d, h, w) in LAYERS]
if os.environ.get("ALL_BB"):
    LAYERS = LAYERS + ALL_BB
for name, t, cin, cout, k, s, p, d, Hi, Wi in LAYERS:
    b = 64 if "B=64" in name else B
    (kh, kw), (sh, sw) = _p(k), _p(s)
    if t:
        Ho, Wo = (Hi - 1) * sh - 2 * p + d * (kh - 1) + 1, (Wi - 1) * sw - 2 * p + d * (kw - 1) + 1
    else:
        Ho, Wo = (Hi + 2 * p - d * (kh - 1) - 1) // sh + 1, (Wi + 2 * p - d * (kw - 1) - 1) // sw + 1
    desc = _lib.ConvDesc(b, cin, cout, Hi, Wi, Ho, Wo, kh, kw, sh, sw, p, p, d, d, t)
    x = torch.rand(b, Hi, Wi, cin, device=dev).bfloat16()
    y = torch.empty(b, Ho, Wo, cout, device=dev, dtype=torch.bfloat16)
    dy = (torch.rand(b, Ho, Wo, cout, device=dev) - 0.5).bfloat16()
    dx = torch.empty_like(x)
    w = (torch.rand((cin, cout, kh, kw) if t else (cout, cin, kh, kw), device=dev) - 0.5) * 0.05
    bias = torch.zeros(cout, device=dev)
    dw, db = torch.empty_like(w), torch.empty_like(bias)
    n = int(_lib.load().dd_conv2d_workspace_bytes(ctypes.byref(desc)))
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    runs = {
        "fwd": lambda: call("dd_conv2d_fwd", x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), ctypes.byref(desc), 1, 1, ws.data_ptr(), n, st),
        "dgrad": lambda: call("dd_conv2d_dgrad", dy.data_ptr(), w.data_ptr(), x.data_ptr(), dx.data_ptr(), ctypes.byref(desc), 1, ws.data_ptr(), n, st),
        "wgrad": lambda: call("dd_conv2d_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), ctypes.byref(desc), 1, ws.data_ptr(), n, st),
    }
    for pas, fn in runs.items():
        if os.environ.get("SKIP_WGRAD") and pas == "wgrad":
            continue
        opix = b * (Hi * Wi if t else Ho * Wo)        # MACs: one per (pixel of the non-upsampled side, tap, cin, cout)
        flops = 2.0 * opix * cin * cout * kh * kw
        if t and sh == 1 and pas != "dgrad":
            flops = 2.0 * b * Ho * Wo * cin * cout * kh * kw      # stride-1 transposed conv as a gather over the output
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tc = _lib.load().dd_conv2d_tc_supported(ctypes.byref(desc), 1, {"fwd": 0, "dgrad": 1, "wgrad": 2}[pas])
        print(f"{name:32s} {pas:5s} B={b:3d}  {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s ({flops / ms / 1e9 / PEAK * 100:5.1f}% of {PEAK:.0f})  tcgen05={tc}")
    del x, y, dy, dx
    torch.cuda.empty_cache()
