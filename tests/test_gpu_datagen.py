"""On-device synthetic data generator (dpivae_sample_response) vs the reference's `sample_response`
arithmetic (utils/data.py:9-52) evaluated with torch ops on the SAME generator stream: torch.rand per factor, the
surrogate MLP, then torch.randn noise for x, c, y."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case,n", [("bridge", 1000), ("damped_oscillator", 4097), ("simple_beam", 50000)])
def test_sample_response_device_matches_torch_stream(case, n):
    import dpivae_b200 as dpv

    case_mod = importlib.import_module(f"dpivae_b200.cases.{case}")
    d = case_mod.definition
    torch.manual_seed(11)
    x, c, y, z = dpv.sample_response_device(d, n)
    off_dev = torch.cuda.default_generators[0].get_offset()

    torch.manual_seed(11)
    gt = d["dict_gt"]
    cols = []
    for k, v in gt.items():
        lo, hi = float(v["args"]["low"]), float(v["args"]["high"])
        cols.append(torch.distributions.Uniform(torch.tensor(lo, device="cuda"), torch.tensor(hi, device="cuda")).sample((n,)))
    z_ref = torch.stack(cols, dim=1)
    fm = d["full_model"].to("cuda")
    with torch.no_grad():
        # true-fp32 reference of the surrogate (no TF32): fp64 evaluation of the same weights
        zt = (z_ref.double() - fm.input_transform.mean_.to("cuda").double()) / fm.input_transform.scale_.to("cuda").double()
        h = zt
        lin = fm.linear_layers()
        for i, l in enumerate(lin):
            h = h @ l.weight.double().t() + l.bias.double()
            if i < len(lin) - 1:
                h = torch.tanh(h)
    idx_c = [i for i, v in enumerate(gt.values()) if v["type"] == "c"]
    idx_y = [i for i, v in enumerate(gt.values()) if v["type"] == "y"]
    ex = torch.randn(n, d["nd_x"], device="cuda")
    ec = torch.randn(n, len(idx_c), device="cuda")
    ey = torch.randn(n, len(idx_y), device="cuda")
    assert torch.cuda.default_generators[0].get_offset() == off_dev   # same generator consumption
    assert torch.equal(z, z_ref)                                        # bit-identical uniform stream
    sx, sc, sy = float(d["sigma_x"]), float(d["sigma_c"]), float(d["sigma_y"])
    x_ref = h + sx * ex.double()
    assert (x.double() - x_ref).abs().max() <= 1e-5 * max(1.0, float(x_ref.abs().max()))
    assert torch.equal(c, z_ref[:, idx_c] + ec * sc)   # bit-identical normal stream and arithmetic
    assert torch.equal(y, z_ref[:, idx_y] + ey * sy)
    assert tuple(x.shape) == (n, d["nd_x"]) and tuple(c.shape) == (n, d["nd_c"]) and tuple(y.shape) == (n, d["nd_y"])
