"""The oracle (oracle/dstd_oracle.py) against vectors produced by the reference itself
(tests/golden/make_golden.py) and the Appendix-D anchors.  CPU only."""
import pytest
import torch

from oracle import dstd_oracle as orc
from tests.helpers import load_json, load_npz, max_abs, rel_err, split


def _leafify(p, req_keys):
    out = {}
    for k, v in p.items():
        v = v.clone()
        if k in req_keys:
            v.requires_grad_(True)
        out[k] = v
    return out


@pytest.mark.parametrize("variant", ["std", "fast"])
@pytest.mark.parametrize("mode", ["spatial", "temporal"])
def test_operator(variant, mode):
    z = load_npz(f"op_{variant}_{mode}.npz")
    p = _leafify(split(z, "p."), set(split(z, "g.").keys()))
    x, A, alpha = (z[k].clone().requires_grad_(True) for k in ("x", "A", "alpha"))
    y = orc.dstdgc(x, A, alpha, p, mode, fast=(variant == "fast"))
    assert max_abs(y, z["y"]) < 1e-12
    (y * z["gy"]).sum().backward()
    assert max_abs(x.grad, z["g_x"]) < 1e-11
    assert max_abs(A.grad, z["g_A"]) < 1e-11
    assert max_abs(alpha.grad, z["g_alpha"]) < 1e-10
    for k, g in split(z, "g.").items():
        assert max_abs(p[k].grad, g) < 1e-10, k


@pytest.mark.parametrize("variant", ["std", "fast"])
@pytest.mark.parametrize("tag", ["in", "mid", "out"])
def test_block(variant, tag):
    z = load_npz(f"block_{variant}_{tag}.npz")
    grads = split(z, "g.")
    p = _leafify(split(z, "p."), set(grads.keys()))
    x = z["x"].clone().requires_grad_(True)
    fast = variant == "fast"
    y = orc.dstdgcb(x, p, "", True, fast)
    assert max_abs(y, z["y"]) < 1e-10
    (y * z["gy"]).sum().backward()
    assert rel_err(x.grad, z["g_x"]) < 1e-10
    for k, g in grads.items():
        assert max_abs(p[k].grad, g) < 1e-8 * max(1.0, float(g.abs().max())), k
    for k, b in split(z, "after.").items():
        assert max_abs(p[k], b) < 1e-10, k
    with torch.no_grad():
        ye = orc.dstdgcb(z["x"], p, "", False, fast)
    assert max_abs(ye, z["y_eval"]) < 1e-9 * max(1.0, float(z["y_eval"].abs().max()))


@pytest.mark.parametrize("name", ["std_h36m", "std_cmu", "std_3dpw", "fast_h36m"])
def test_model(name):
    z = load_npz(f"model_{name}.npz")
    fast = name.startswith("fast")
    grads = split(z, "g.")
    p = _leafify(split(z, "p."), set(grads.keys()))
    x = z["x"].clone().requires_grad_(True)
    y = orc.dstdgcn(x, p, True, fast)
    assert max_abs(y, z["y"]) < 1e-9
    loss = y.pow(2).mean()
    assert abs(float(loss) - float(z["loss"])) < 1e-10
    loss.backward()
    assert rel_err(x.grad, z["g_x"]) < 1e-9
    for k, g in grads.items():
        assert max_abs(p[k].grad, g) < 1e-8 * max(1.0, float(g.abs().max())), k
    for k, b in split(z, "after.").items():
        assert max_abs(p[k], b) < 1e-10, k
    with torch.no_grad():
        ye = orc.dstdgcn(z["x"], p, False, fast)
    assert rel_err(ye, z["y_eval"]) < 1e-9


@pytest.mark.parametrize("variant", ["std", "fast"])
def test_engine_steps(variant):
    """Three engine steps (two forwards + one backward + Adam) reproduce the reference trajectory."""
    z = load_npz(f"train_{variant}.npz")
    after = split(z, "after.")
    keys = load_json("state_keys.json")["std_h36m" if variant == "std" else "fast_h36m"]["keys"]
    trainable = {k for k, _, _, req, _ in keys if req}
    p0 = split(z, "p.")
    p = _leafify(p0, {k for k in p0 if _generic(k) in {_generic(t) for t in trainable}})
    params = [v for v in p.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=3e-3, weight_decay=0)
    for step in range(3):
        loss = orc.train_loss(p, z[f"inputs{step}"], z[f"inputs_inv{step}"], z[f"targets{step}"],
                              fast=(variant == "fast"), inverse=True)
        opt.zero_grad()
        loss.backward()
        opt.step()
        assert abs(float(loss) - float(z["losses"][step])) < 1e-9
    for k, v in after.items():
        assert max_abs(p[k], v) < 1e-8, k


def _generic(key):
    """encoders.3.x -> encoders.N.x so a small model's keys match the L=5 key table."""
    parts = key.split(".")
    if parts[0] == "encoders":
        parts[1] = "N"
    return ".".join(parts)


@pytest.mark.parametrize("variant", ["std", "fast"])
def test_appendix_d_anchor_recipe_is_reproducible(variant):
    """anchors.json (recomputed from the reference by make_golden.py) equals SURVEY.md Appendix D."""
    a = load_json("anchors.json")[variant]
    want = {"std": (-2084.2344962196, 16.6128598017, 1719.1769071581),
            "fast": (578.5798513531, 4.3917632569, 209.9816396032)}[variant]
    assert abs(a["sum_y"] - want[0]) < 1e-6
    assert abs(a["loss"] - want[1]) < 1e-8
    assert abs(a["grad_l2"] - want[2]) < 1e-6
