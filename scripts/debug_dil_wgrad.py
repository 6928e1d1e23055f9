import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from driving_dirty_b200 import _lib
from driving_dirty_b200._lib import call, stream_ptr
torch.manual_seed(0)
for (t, cin, cout, k, p, d, hw, B) in [(1, 96, 64, 7, 0, 7, (20, 150), 2), (1, 64, 32, 7, 0, 7, (30, 140), 2), (0, 32, 32, 3, 0, 1, (50, 258), 2)]:
    Hi, Wi = hw
    Ho = (Hi - 1) - 2 * p + d * (k - 1) + 1 if t else Hi + 2 * p - d * (k - 1)
    Wo = (Wi - 1) - 2 * p + d * (k - 1) + 1 if t else Wi + 2 * p - d * (k - 1)
    desc = _lib.ConvDesc(B, cin, cout, Hi, Wi, Ho, Wo, k, k, 1, 1, p, p, d, d, t)
    x = torch.randn(B, cin, Hi, Wi).bfloat16().float()
    dy = torch.randn(B, cout, Ho, Wo).bfloat16().float()
    w = torch.zeros((cin, cout, k, k) if t else (cout, cin, k, k), requires_grad=True)
    y = F.conv_transpose2d(x, w, dilation=d, padding=p) if t else F.conv2d(x, w, dilation=d, padding=p)
    y.backward(dy)
    ref = w.grad
    xd = x.permute(0, 2, 3, 1).contiguous().bfloat16().cuda()
    dyd = dy.permute(0, 2, 3, 1).contiguous().bfloat16().cuda()
    n = int(_lib.load().dd_conv2d_workspace_bytes(ctypes.byref(desc)))
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    outs = []
    for rep in range(6):
        ws.fill_(0xFF if rep % 2 else 0)
        dw = torch.full(ref.shape, float("nan"), device="cuda")
        db = torch.empty(cout, device="cuda")
        call("dd_conv2d_wgrad", xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), db.data_ptr(), ctypes.byref(desc), 1, ws.data_ptr(), n, stream_ptr())
        torch.cuda.synchronize()
        o = dw.cpu()
        err = (o - ref).abs()
        bad = err > 1e-2 * ref.abs().max()
        print(f"layer cin{cin} cout{cout} k{k} d{d} rep {rep}: rel-max err {float(err.max() / ref.abs().max()):.3e} nan {int(torch.isnan(o).sum())} bad {int(bad.sum())}",
              ("first bad idx " + str(torch.nonzero(bad)[:6].tolist())) if bad.any() else "")
        outs.append(o)
    print("  run-to-run identical:", all(torch.equal(outs[0], o) for o in outs[1:]))
