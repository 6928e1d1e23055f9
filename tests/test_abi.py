"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header
declares, the ctypes table covers the header, and the product refuses CPU tensors."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from driving_dirty_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dd_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.dd_version() >= 100
    assert lib.dd_launch_count() >= 0


def test_argument_errors_are_reported_not_swallowed():
    from driving_dirty_b200 import _lib
    lib = _lib.load()
    assert lib.dd_stitch_f32(None, None, 1, 2, 2, None) == -1
    assert "null" in _lib.last_error()
    assert lib.dd_conv3x3_c32_fwd(1 << 20, 1 << 20, 1 << 20, 1 << 20, 0, 1, 8, 8, 3, 0, None) == -2
    assert "stride" in _lib.last_error()
    with pytest.raises(RuntimeError, match="dd_status -1"):
        _lib.call("dd_bce_ts_fwd", None, None, 0, None, None, None, None, None, 0, 10, None)


def test_no_cpu_fallback():
    from driving_dirty_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.stitch(torch.zeros(1, 6, 3, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.threat_score(torch.zeros(4), torch.zeros(4))


def test_same_seed_same_initial_weights_as_reference(golden):
    """Constructors consume the torch RNG like the reference's: ae_small.pt holds the reference
    BasicAE.state_dict() built under torch.manual_seed(20200505)."""
    from driving_dirty_b200.autoencoder.autoencoder import BasicAE, default_hparams
    g = golden("ae_small")
    torch.manual_seed(20200505)
    ae = BasicAE(default_hparams(hidden_dim=g["hidden"], latent_dim=g["latent"], input_width=6 * g["view_w"],
                                 input_height=g["view_h"], output_width=g["view_w"], output_height=g["view_h"]))
    sd = ae.state_dict()
    assert sorted(sd) == sorted(g["state_dict"])
    for k, v in g["state_dict"].items():
        assert torch.equal(sd[k], v), k
    assert (ae.decoder.deconv_dim_h, ae.decoder.deconv_dim_w) == tuple(g["hw"])


def test_roadmap_state_dict_keys_match_reference(golden):
    from tests.helpers import make_roadmap_model
    from oracle import scene_oracle as so
    g = golden("roadmap_small")
    p = so.init_roadmap_params(g["hidden"], g["latent"], g["view_h"], g["view_w"])
    m = make_roadmap_model(p, g["hidden"], g["latent"], g["view_h"], g["view_w"], device="cpu")
    assert sorted(m.state_dict()) == sorted(g["params_sha"])
    assert m.frozen and not any(q.requires_grad for q in m.ae.parameters()) and m.fc1.weight.requires_grad


def test_c_oracle_matches_numpy_oracle():
    import ctypes
    import numpy as np
    from oracle import scene_oracle as so
    path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = ctypes.CDLL(path)
    views, road = so.synthetic_scene_batch(2, 6, 8, map_hw=16, seed=3)
    v = np.ascontiguousarray(views.numpy())
    out = np.empty((2, 3, 6, 48), np.float32)
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    lib.oracle_stitch(fp(v), fp(out), 2, 6, 8)
    assert np.array_equal(out, so.stitch(views).numpy())
    x, y = np.empty_like(out), np.empty((2, 3, 6, 8), np.float32)
    lib.oracle_stitch_mask(fp(v), fp(x), fp(y), 2, 6, 8, 3)
    xo, yo = so.six_to_one(views, 3)
    assert np.array_equal(x, xo.numpy()) and np.array_equal(y, yo.numpy())
    a3 = torch.relu(torch.randn(2, 32, 3, 6))
    pooled = np.empty((2, 144), np.float32)
    arg = np.empty((2, 144), np.uint8)
    lib.oracle_pool4_flat(fp(np.ascontiguousarray(a3.numpy())), fp(pooled), fp(arg), 2, ctypes.c_longlong(576))
    assert np.array_equal(pooled, so.pool4_flat(a3).numpy()) and np.array_equal(arg, so.pool4_flat_argmax(a3))
    w, b = torch.randn(4, 3, 3, 3), torch.randn(4)
    yc = np.empty((2, 4, 3, 24), np.float32)
    xin = so.stitch(views)
    lib.oracle_conv3x3_relu(fp(np.ascontiguousarray(xin.numpy())), fp(w.numpy()), fp(b.numpy()), fp(yc), 2, 3, 4, 6, 48, 2)
    ref = torch.relu(torch.nn.functional.conv2d(xin, w, b, stride=2, padding=1))
    assert np.allclose(yc, ref.numpy(), atol=1e-5)
    logits = torch.randn(1000) * 1e-7
    bo = np.empty(1000, np.uint8)
    lib.oracle_binarise(fp(logits.numpy()), fp(bo), ctypes.c_longlong(1000))
    assert np.array_equal(bo, torch.sigmoid(logits).round().numpy().astype(np.uint8))
    t = (torch.rand(1000) > 0.5)
    cnt = (ctypes.c_longlong * 3)()
    lib.oracle_ts_counts(fp(t.numpy().astype(np.uint8)), fp(bo), ctypes.c_longlong(1000), cnt)
    tp, nt, nr = so.threat_score_counts(t.float(), torch.from_numpy(bo).float())
    assert (cnt[0], cnt[1], cnt[2]) == (nt, nr, tp)
    lib.oracle_bce_mean.restype = ctypes.c_double
    big = torch.randn(1000)
    bce = lib.oracle_bce_mean(fp(big.numpy()), fp(t.float().numpy()), ctypes.c_longlong(1000))
    assert abs(bce - float(torch.nn.functional.binary_cross_entropy_with_logits(big, t.float()))) < 1e-6
