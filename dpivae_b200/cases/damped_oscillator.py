"""Damped oscillator case (reference: cases/damped_oscillator/__init__.py:25-217).  4 factors,
physics decoder = undamped closed form cos(t / sqrt(m)) (mass_spring.py:8-28)."""
import torch
from torch import distributions as dist

from ..utils import device, get_shapes_from_dict
from ._common import MassSpring, SurrogateMLP, load_assets, make_definition, uniform

dict_gt = {
    "m": uniform(1.2, 1.8, "x", r"$m$ [kg]", 1.5),
    "zeta": uniform(0.0, 2.0, "y", r"$\zeta$ [-]", 0.0),
    "T": uniform(0.01, 39.99, "c", r"$T$", 20.0),
    "x_0": uniform(0.9, 1.1, "f", r"$x_0$ [m]", 1.0),
}
dict_prior_x = {"m": {"lb": 1.0, "ub": 2.0, "dist": dist.Uniform, "args": {"low": 1.0, "high": 2.0}}}
nd_x = 64
_assets = load_assets("damped_oscillator")
t = torch.from_numpy(_assets["t"].copy()).to(device)  # linspace(0, 0.05*199, 64), cases/damped_oscillator/__init__.py:87-91
full_model = SurrogateMLP(_assets, "full").to(device)
part_model = MassSpring(t)

presets = {
    "vae": {"model_type": "P", "lambda_g0": -1.0, "lambda_x": None, "nz_c": 4, "nz_y": 4},
    "dpivae": {"model_type": "S", "lambda_g0": 1 / 128, "lambda_x": None, "nz_c": 4, "nz_y": 4},
}
definition = make_definition(nd_x, dict_gt, dict_prior_x, t, 0.01, full_model, part_model,
                             get_shapes_from_dict(dict_gt), x_unit="Time [s]", y_unit="[m]", ylim=(-2.0, 2.0))
