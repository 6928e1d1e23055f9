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
for name, t, cin, cout, k, p, d, Hi, Wi in LAYERS:
    b = 64 if "B=64" in name else B
    Ho = (Hi - 1) - 2 * p + d * (k - 1) + 1 if t else Hi + 2 * p - d * (k - 1)
    Wo = (Wi - 1) - 2 * p + d * (k - 1) + 1 if t else Wi + 2 * p - d * (k - 1)
    desc = _lib.ConvDesc(b, cin, cout, Hi, Wi, Ho, Wo, k, k, 1, 1, p, p, d, d, t)
    x = torch.rand(b, Hi, Wi, cin, device=dev).bfloat16()
    y = torch.empty(b, Ho, Wo, cout, device=dev, dtype=torch.bfloat16)
    dy = (torch.rand(b, Ho, Wo, cout, device=dev) - 0.5).bfloat16()
    dx = torch.empty_like(x)
    w = (torch.rand((cin, cout, k, k) if t else (cout, cin, k, k), device=dev) - 0.5) * 0.05
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
        opix = b * (Ho * Wo if pas != "dgrad" else Hi * Wi)
        flops = 2.0 * opix * cin * cout * k * k
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
