"""Bridge population case (reference: cases/bridge/__init__.py:24-286).  7 generative factors,
physics decoder = pretrained Tanh-MLP surrogate 3 -> 64 -> 32 -> 64 -> 64 on (kv1, kv2, delta_xs)."""
import torch
from torch import distributions as dist

from ..utils import device, get_shapes_from_dict
from ._common import SurrogateMLP, load_assets, make_definition, uniform

dict_gt = {
    "kv1": uniform(9.5, 11.5, "x", r"$\log_{10} k_{v,1}$", 11.5),
    "kv2": uniform(9.5, 11.5, "x", r"$\log_{10} k_{v,2}$", 11.5),
    "y1": uniform(0.0, 1.0, "y", r"$y_1$ [-]", 0.1),
    "y2": uniform(0.0, 1.0, "y", r"$y_2$ [-]", 0.1),
    "v": uniform(0.9, 1.1, "c", r"$\delta_{\mathrm{v}}$ [-]", 1.0),
    "delta_xs": uniform(-1.0, 1.0, "c", r"$\delta_\mathrm{s}$ [m]", 0.0, phys=True),
    "f": uniform(0.95, 1.05, "f", r"$\delta_{\mathrm{F}}$ [-]", 1.0),
}
dict_prior_x = {
    k: {"lb": 9.001, "ub": 11.999, "dist": dist.Uniform, "args": {"low": 9.001, "high": 11.999}} for k in ("kv1", "kv2")
}
nd_x = 64
t = torch.linspace(1.0, 21.0, nd_x)
_assets = load_assets("bridge")
full_model = SurrogateMLP(_assets, "full").to(device)
part_model = SurrogateMLP(_assets, "part").to(device)

presets = {
    "vae": {"model_type": "P", "lambda_g0": -1.0, "lambda_x": None, "nz_c": 4, "nz_y": 4},
    "dpivae": {"model_type": "S", "lambda_g0": 1 / 1024, "lambda_x": None, "nz_c": 4, "nz_y": 4},
    "DPIVAE-A": {"name": "DPIVAE-A", "model_type": "P", "lambda_g0": -1.0, "lambda_x": None, "nz_c": 4, "nz_y": 4},
    "DPIVAE-B": {"name": "DPIVAE-B", "model_type": "S", "lambda_g0": 1 / 1024, "lambda_x": None, "nz_c": 4, "nz_y": 4},
}
definition = make_definition(nd_x, dict_gt, dict_prior_x, t, 0.0001, full_model, part_model,
                             get_shapes_from_dict(dict_gt), x_unit="Time [s]", y_unit=r"[$^o/_{oo}$]", ylim=(-1.0, 2.0))
