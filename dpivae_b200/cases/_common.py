"""Building blocks of the case definitions (mirror of the reference's `cases/*/__init__.py`).

A case `definition` dict keeps the keys the training-step path reads (dpivae.py:102-123):
nd_x, nd_c, nd_y, nd_f, nd_p, nz_x, dict_prior_x, dict_gt, sigma_*, n_classes, nk_y, full_model,
part_model, t.  `part_model` objects carry a `physics_kind` tag so that `setup_model` can map them
onto the fused CUDA physics decoder; their torch `__call__` exists for consumers outside the hot
path (synthetic data generation, plotting), not for training.
"""
import os

import numpy as np
import torch
from torch import nn

from ..utils import StandardScaler, device

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def load_assets(case):
    return np.load(os.path.join(ASSETS, f"{case}.npz"))


class SurrogateMLP(nn.Module):
    """Pretrained Tanh-MLP surrogate with a StandardScaler on its input (models/nn.py:29-80,
    cases/bridge/__init__.py:146-185).  state_dict keys: net.{0,2,4,...}.{weight,bias}."""

    physics_kind = "mlp"

    def __init__(self, assets, prefix):
        super().__init__()
        n = int(assets[f"{prefix}_n_layers"])
        mods = []
        for i in range(n):
            w = torch.from_numpy(assets[f"{prefix}_w{i}"].copy())
            lin = nn.Linear(w.shape[1], w.shape[0])
            lin.weight.data.copy_(w)
            lin.bias.data.copy_(torch.from_numpy(assets[f"{prefix}_b{i}"].copy()))
            mods.append(lin)
            if i < n - 1:
                mods.append(nn.Tanh())
        self.net = nn.Sequential(*mods)
        self.input_transform = StandardScaler(assets[f"{prefix}_in_mean"].reshape(1, -1),
                                              assets[f"{prefix}_in_std"].reshape(1, -1))
        self.n_input = self.net[0].in_features
        self.n_output = self.net[-1].out_features
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def linear_layers(self):
        return [m for m in self.net if isinstance(m, nn.Linear)]

    def forward(self, z):
        zt, _ = self.input_transform.forward(z)
        return self.net(zt)


class MassSpring:
    """cases/damped_oscillator/mass_spring.py:8-28 on a fixed time grid."""

    physics_kind = "mass_spring"

    def __init__(self, t):
        self.t = t

    def __call__(self, z):
        t = self.t.to(z.device)
        m = z[..., 0].unsqueeze(-1)
        omega = torch.sqrt(1.0 / m)
        return (0.0 / omega) * torch.sin(omega * t) + 1.0 * torch.cos(omega * t)


class EulerBernoulliBeam:
    """cases/simple_beam/simple_beam_model.py:4-30 with npts = nd_x."""

    physics_kind = "beam"

    def __init__(self, npts, I=2e-6, L=1.0, P=1.0):
        self.npts, self.I, self.L, self.P = npts, I, L, P
        self.t = torch.linspace(0.0, L, npts)

    def __call__(self, z):
        x = self.t.to(z.device)
        E = z[..., 0].unsqueeze(-1) * 1e6
        a = z[..., 1].unsqueeze(-1)
        b = self.L - a
        if (torch.any(a < 0.0)) or (torch.any(a > self.L)):
            raise ValueError("Load position must be between 0 and L")
        w = self.P * b * x * (self.L**2 - b**2 - x**2) / (6 * E * self.I * self.L)
        wb = self.P * ((x - a) ** 3) / (6 * E * self.I)
        return -1000.0 * (w + torch.where(x > a, wb, torch.zeros_like(wb)))


def uniform(lo, hi, typ, label="", val=None, phys=False, lb=None, ub=None):
    """One ground-truth factor: Uniform(lo, hi) with plotting / transform bounds lb, ub."""
    from torch import distributions as dist

    return {"lb": lo if lb is None else lb, "ub": hi if ub is None else ub, "dist": dist.Uniform,
            "args": {"low": lo, "high": hi}, "type": typ, "label": label, "val": hi if val is None else val, "phys": phys}


def make_definition(nd_x, dict_gt, dict_prior_x, t, sigma, full_model, part_model, shapes, **extra):
    nz_x, nd_c, nd_y, nd_f, nd_p = shapes
    d = {
        "nd_x": nd_x, "nd_c": nd_c, "nd_y": nd_y, "nd_f": nd_f, "nd_p": nd_p, "nz_x": nz_x,
        "t_min": float(t.min()), "t_max": float(t.max()), "t": t,
        "dict_prior_x": dict_prior_x, "dict_gt": dict_gt,
        "sigma_x": torch.tensor(sigma).to(device), "sigma_c": torch.tensor(sigma).to(device),
        "sigma_y": torch.tensor(sigma).to(device),
        "n_classes": None, "bins_y": None, "nk_y": None, "logsoftmax_y": False,
        "full_model": full_model, "part_model": part_model,
    }
    d.update(extra)
    return d
