"""Host layer (modules -> autograd -> torch.library ops) exercised on CPU by injecting the ABI emulation
from oracle/ in place of the CUDA backend (tests only), in float64, against the reference goldens.
This pins (a) the hand-derived backward formulas the CUDA kernels implement and (b) the module wiring."""
import pytest
import torch

import dstd_gcn_b200  # noqa: F401
from dstd_gcn_b200 import _lib
from dstd_gcn_b200.model import dstdgcn as std
from dstd_gcn_b200.model import dstdgcn_fast as fast
from oracle.abi_emul import EmulBackend
from tests.helpers import load_json, load_npz, max_abs, rel_err, split


@pytest.fixture(autouse=True)
def emul_backend():
    prev = _lib._set_backend_for_tests(EmulBackend())
    yield
    _lib._set_backend_for_tests(prev)


def _mod(variant):
    return std if variant == "std" else fast


def _load(module, params):
    with torch.no_grad():
        sd = module.state_dict()
        for k in sd:
            sd[k] = params[k].clone()
    module.load_state_dict(sd, strict=True)
    return module


@pytest.mark.parametrize("variant", ["std", "fast"])
@pytest.mark.parametrize("mode", ["spatial", "temporal"])
def test_operator_module(variant, mode):
    z = load_npz(f"op_{variant}_{mode}.npz")
    shp = z["x"].shape
    n, cin, t, v = (shp if variant == "std" else (shp[0], shp[3], shp[1], shp[2]))
    cout = z["p.conv_f.weight"].shape[0]
    ref_c, kpt = (t, v) if mode == "spatial" else (v, t)
    op = _load(_mod(variant).DSTDGC(cin, cout, ref_c, kpt, mode=mode).double(), split(z, "p."))
    x, A, alpha = (z[k].clone().requires_grad_(True) for k in ("x", "A", "alpha"))
    y = op(x, A, alpha)
    assert y.shape == z["y"].shape and y.is_contiguous()
    assert max_abs(y, z["y"]) < 1e-11
    (y * z["gy"]).sum().backward()
    assert max_abs(x.grad, z["g_x"]) < 1e-10
    assert max_abs(A.grad, z["g_A"]) < 1e-10
    assert max_abs(alpha.grad, z["g_alpha"]) < 1e-9
    for k, g in split(z, "g.").items():
        assert max_abs(dict(op.named_parameters())[k].grad, g) < 1e-9, k


@pytest.mark.parametrize("variant", ["std", "fast"])
@pytest.mark.parametrize("tag,cin,cout", [("in", 6, 8), ("mid", 8, 8), ("out", 8, 3)])
def test_block_module(variant, tag, cin, cout):
    z = load_npz(f"block_{variant}_{tag}.npz")
    blk = _load(_mod(variant).DSTDGCB(cin, cout, 12, 22, "h36m").double(), split(z, "p.")).train()
    x = z["x"].clone().requires_grad_(True)
    y = blk(x)
    assert max_abs(y, z["y"]) < 1e-9
    (y * z["gy"]).sum().backward()
    assert rel_err(x.grad, z["g_x"]) < 1e-9
    named = dict(blk.named_parameters())
    for k, g in split(z, "g.").items():
        assert max_abs(named[k].grad, g) < 1e-8 * max(1.0, float(g.abs().max())), k
    sd = blk.state_dict()
    for k, b in split(z, "after.").items():
        assert max_abs(sd[k], b) < 1e-10, k
    blk.eval()
    with torch.no_grad():
        ye = blk(z["x"])
    assert max_abs(ye, z["y_eval"]) < 1e-9 * max(1.0, float(z["y_eval"].abs().max()))


@pytest.mark.parametrize("name,v,layout", [("std_h36m", 22, "h36m"), ("std_cmu", 25, "cmu"), ("std_3dpw", 23, "3dpw"),
                                           ("fast_h36m", 22, "h36m")])
def test_model_module(name, v, layout):
    z = load_npz(f"model_{name}.npz")
    variant = name.split("_")[0]
    m = _load(_mod(variant).DSTDGCN(6, 4, 6, 0.0, v, 8, 2, layout).double(), split(z, "p.")).train()
    x = z["x"].clone().requires_grad_(True)
    y = m(x)
    assert max_abs(y, z["y"]) < 1e-9
    loss = y.pow(2).mean()
    loss.backward()
    assert rel_err(x.grad, z["g_x"]) < 1e-9
    named = dict(m.named_parameters())
    grads = split(z, "g.")
    for k, g in grads.items():
        assert max_abs(named[k].grad, g) < 1e-8 * max(1.0, float(g.abs().max())), k
    # frozen adjacencies get no gradient, exactly the reference's set of trainable tensors gets one
    assert {k for k, p in named.items() if p.grad is not None} == set(grads.keys())
    sd = m.state_dict()
    for k, b in split(z, "after.").items():
        assert max_abs(sd[k], b) < 1e-10, k
    m.eval()
    with torch.no_grad():
        assert rel_err(m(z["x"]), z["y_eval"]) < 1e-9


@pytest.mark.parametrize("fused", [(True, True), (True, False), (False, False)])
def test_model_module_skip_gradient_paths(monkeypatch, fused):
    """ops.FUSE_SKIP_GRAD routes the layer skip through the BN node (gr_add of the BN backward), ops.FUSE_RES_GRAD also
    the BN residual through the spatial unit's node (gx_add); off = plain autograd sums.  All reproduce the goldens."""
    from dstd_gcn_b200 import ops
    monkeypatch.setattr(ops, "FUSE_SKIP_GRAD", fused[0])
    monkeypatch.setattr(ops, "FUSE_RES_GRAD", fused[1])
    test_model_module("std_h36m", 22, "h36m")
    test_model_module("fast_h36m", 22, "h36m")


@pytest.mark.parametrize("variant", ["std", "fast"])
def test_init_rng_parity_and_appendix_d_anchor(variant):
    """Constructing OUR module under the reference's seed consumes the RNG identically, so the Appendix-D recipe
     run through our modules reproduces the reference's known answers."""
    a = load_json("anchors.json")[variant]
    torch.manual_seed(777)
    m = _mod(variant).DSTDGCN(6, 10, 25, 0.0, 22, 64, 5, "h36m")
    with torch.no_grad():
        for k, p in m.named_parameters():
            leaf = k.split(".")[-1]
            if leaf in ("alpha_sm", "alpha_tm"):
                p.fill_(0.1)
            elif leaf == "W_s":
                p.fill_(0.05)
            elif leaf == "R_t":
                p.fill_(0.01)
    m.double().train()
    x = torch.randn(4, 35, 22, 3, generator=torch.Generator().manual_seed(1234), dtype=torch.float64)
    y = m(x)
    loss = y.pow(2).mean()
    loss.backward()
    gn = torch.sqrt(sum((p.grad ** 2).sum() for p in m.parameters() if p.grad is not None))
    assert abs(float(y.sum()) - a["sum_y"]) < 1e-6
    assert abs(float(loss) - a["loss"]) < 1e-9
    assert abs(float(gn) - a["grad_l2"]) < 1e-6


def test_state_dict_layout_matches_reference():
    keys = load_json("state_keys.json")
    cases = {"std_h36m": std, "fast_h36m": fast, "std_cmu": std, "std_3dpw": std, "std_default_layers": std}
    for name, mod in cases.items():
        spec = keys[name]
        m = mod.DSTDGCN(*spec["args"])
        sd = m.state_dict()
        req = {k: p.requires_grad for k, p in m.named_parameters()}
        got = [[k, list(v.shape), str(v.dtype).replace("torch.", ""), bool(req.get(k, False)), k in req]
               for k, v in sd.items()]
        assert got == spec["keys"], name
        assert sum(p.numel() for p in m.parameters() if p.requires_grad) == spec["n_trainable"]
        assert sum(p.numel() for p in m.parameters()) == spec["n_total"]


def test_registry_and_errors():
    from dstd_gcn_b200.model import get_model
    from dstd_gcn_b200.model.layers.graph import Graph
    opts = {"dstdgcn": dict(input_channels=6, input_time_frame=4, output_time_frame=6, st_gcnn_dropout=0.0,
                            joints_to_consider=22, num_feature=8, num_layers=1, layout="h36m")}
    m = get_model("dstdgcn", **opts)
    assert isinstance(m, std.DSTDGCN)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 9, 22, 3))                       # wrong frame count (dstdgcn.py:296)
    with pytest.raises(AssertionError):
        std.BatchNorm(8, 22, 10)(torch.zeros(1, 8, 9, 22))  # shape assert (dstdgcn.py:46)
    with pytest.raises(AssertionError):
        std.DSTDGC(4, 4, 5, 5, mode="other")
    with pytest.raises(NotImplementedError):
        Graph("nope")
    with pytest.raises(NotImplementedError):
        std.DSTDGC(4, 4, 5, 5, red_channels=3)            # the kernels implement the reference's red_channels=2


def test_deferred_bn_updates_match_torch_batchnorm():
    """`ops.apply_deferred_bn_updates` (the running-statistics update of the second, concurrently running pass of a
    training step) against torch.nn.BatchNorm1d applied to the same two batches in order."""
    import torch
    from dstd_gcn_b200 import ops
    torch.manual_seed(0)
    c, n, t = 12, 6, 9
    bn = torch.nn.BatchNorm1d(c).double().train()
    ref = torch.nn.BatchNorm1d(c).double().train()
    xa, xb = torch.randn(n, c, t, dtype=torch.float64) * 3 + 1, torch.randn(n, c, t, dtype=torch.float64) * 0.5 - 2
    ref(xa)
    ref(xb)
    bn(xa)                                                     # first pass: updated in place
    mean = xb.mean(dim=(0, 2))
    invstd = 1.0 / torch.sqrt(xb.var(dim=(0, 2), unbiased=False) + bn.eps)
    ops.apply_deferred_bn_updates([[bn.running_mean, bn.running_var, bn.num_batches_tracked, bn.momentum, bn.eps,
                                    float(n * t), mean, invstd]])
    assert torch.allclose(bn.running_mean, ref.running_mean, rtol=1e-12, atol=1e-12)
    assert torch.allclose(bn.running_var, ref.running_var, rtol=1e-10, atol=1e-12)
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == 2
