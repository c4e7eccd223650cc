"""Simple beam case (reference: cases/simple_beam/__init__.py:25-219).  4 factors, physics decoder =
Euler-Bernoulli point-load deflection (simple_beam_model.py:4-30); Normal priors on (E, x_F)."""
import torch
from torch import distributions as dist

from ..utils import device, get_shapes_from_dict
from ._common import EulerBernoulliBeam, SurrogateMLP, load_assets, make_definition, uniform

dict_gt = {
    "E": uniform(2.5, 4.5, "x", r"$E$", 3.0, lb=2.0, ub=6.0),
    "x_F": uniform(0.3, 0.7, "x", r"$x_F$", 0.5, lb=0.01, ub=0.99),
    "log_kv": uniform(6.0, 8.0, "y", r"$\log k_v$", 8.0, lb=5.0, ub=9.0),
    "T": uniform(-11.0, 5.0, "c", r"$T \ [\mathrm{C}^o]$", 5.0, lb=-15.0, ub=15.0),
}
dict_prior_x = {
    "E": {"lb": 2.0, "ub": 6.0, "dist": dist.Normal, "args": {"loc": 4.0, "scale": 1.0}},
    "x_F": {"lb": 0.01, "ub": 0.99, "dist": dist.Normal, "args": {"loc": 0.5, "scale": 0.2}},
}
nd_x = 32
t = torch.linspace(0.00001, 1.0, nd_x)
_assets = load_assets("simple_beam")
full_model = SurrogateMLP(_assets, "full").to(device)
part_model = EulerBernoulliBeam(nd_x)

presets = {
    "vae": {"model_type": "P", "lambda_g0": -1.0, "lambda_x": None, "nz_c": 2, "nz_y": 2},
    "dpivae": {"model_type": "S", "lambda_g0": 1 / 256, "lambda_x": None, "nz_c": 2, "nz_y": 2},
}
definition = make_definition(nd_x, dict_gt, dict_prior_x, t, 0.02, full_model, part_model,
                             get_shapes_from_dict(dict_gt), x_unit="Distance [m]", y_unit="[mm]", ylim=(-25.0, 2.0))
