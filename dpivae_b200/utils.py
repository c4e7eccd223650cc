"""Host-side mirror of the reference's `utils/` package for the training-step path.

Same names, argument meaning and error behaviour as the reference (file:line cited per item);
everything here is host logic or small descriptors -- the arithmetic of the step runs in
libdpivae_b200.so.
"""
import argparse
import math

import numpy as np
import torch
from torch import distributions as dist

device = "cuda" if torch.cuda.is_available() else "cpu"  # utils/__init__.py:5


# ------------------------------------------------------------------------------------------------
# flags (utils/__init__.py:19-116): the defaults are the contract
# ------------------------------------------------------------------------------------------------
_ANNEAL = ["cyclical", "sigmoid", "None", None]
_FLAGS = [
    ("name", str, "default"), ("seed", int, 123),
    ("encoder_x", str, "NN"), ("encoder_c", str, "NN"), ("encoder_y", str, "NN"),
    ("n_iter", int, 20_000), ("n_train", int, 1024), ("n_val", int, 512), ("n_test", int, 512),
    ("n_batch", int, 64), ("n_mc_train", int, 16), ("n_mc_val", int, 64), ("n_mc_test", int, 512),
    ("val_freq", int, 10),
    ("lambda_g0", float, 1 / 256), ("beta_x0", float, 1.0), ("beta_c0", float, 1.0), ("beta_y0", float, 1.0),
    ("lambda_x", float, None), ("alpha_x", float, 1.0), ("alpha_c", float, 1.0), ("alpha_y", float, 1.0),
    ("lr", float, 1e-3), ("lr_e", float, 1e-3), ("lr_ex", float, 1e-3), ("lr_ec", float, 1e-3),
    ("lr_ey", float, 1e-3), ("lr_p", float, 1e-3), ("lr_dx", float, 1e-3), ("lr_dc", float, 1e-3),
    ("lr_dy", float, 1e-3), ("lr_sigma", float, 5e-3),
    ("wd_e", float, 0.0), ("wd_p", float, 0.0), ("wd_dx", float, 0.0), ("wd_dc", float, 0.0),
    ("wd_dy", float, 0.0), ("wd_sigma", float, 0.0),
    ("max_grad_norm", float, 1.0), ("patience", int, 200), ("min_delta", float, 0.001),
    ("n_skip_plot_train", int, 0), ("n_skip_plot_val", int, 0), ("n_plot", int, 2000), ("n_interp", int, 5),
    ("ch_in", int, 1), ("ch_out", int, 16), ("ch_latent", int, 64),
]
_SCHEDULES = {"lambda": (5, 0.5, 0.15, 0.15), "beta_x": (5, 0.5, 0.15, 0.15), "beta_c": (5, 0.5, 0.15, 0.15),
              "beta_y": (4, 0.5, 0.2, 0.2)}


def make_parser():
    p = argparse.ArgumentParser("")
    for name, typ, default in _FLAGS:
        p.add_argument(f"--{name}", type=typ, default=default)
    for name in ["use_seed", "full_cov_prior", "clip_gradients"]:
        p.add_argument(f"--{name}", action="store_true", default=False)
    for key, (n_cycles, R, mu, cov) in _SCHEDULES.items():
        p.add_argument(f"--{key}_annealing", type=str, default=None, choices=_ANNEAL)
        p.add_argument(f"--{key}_n_cycles", type=int, default=n_cycles)
        p.add_argument(f"--{key}_R", type=float, default=R)
        p.add_argument(f"--{key}_mu", type=float, default=mu)
        p.add_argument(f"--{key}_cov", type=float, default=cov)
    return p


# ------------------------------------------------------------------------------------------------
# transforms (utils/transforms.py) -- descriptors; the kernels apply them
# ------------------------------------------------------------------------------------------------
class StandardScaler:
    """utils/transforms.py:42-80: mean / POPULATION std over dim 0, kept as (1, d) tensors."""

    def __init__(self, mean=None, scale=None):
        self.mean_ = None if mean is None else torch.as_tensor(mean, dtype=torch.float32)
        self.scale_ = None if scale is None else torch.as_tensor(scale, dtype=torch.float32)

    def fit(self, sample):
        self.mean_ = sample.mean(0, keepdim=True)
        self.scale_ = sample.std(0, unbiased=False, keepdim=True)
        return self

    def forward(self, z):
        z = (z - self.mean_.to(z.device)) / self.scale_.to(z.device)
        log_det = -torch.log(self.scale_).sum().to(z.device) * torch.ones(z.shape[:-1], device=z.device)
        return z, log_det

    def inverse(self, z):
        z = z * self.scale_.to(z.device) + self.mean_.to(z.device)
        log_det = torch.log(self.scale_).sum().to(z.device) * torch.ones(z.shape[:-1], device=z.device)
        return z, log_det


class ShiftScale:
    """utils/transforms.py:83-105: z * (ub - lb) + lb."""

    def __init__(self, lb, ub):
        self.lb, self.ub = lb, ub
        self.a, self.b = ub - lb, lb


class Logistic:
    """utils/transforms.py:108-133: sigmoid(k z)."""

    def __init__(self, k=1):
        self.k = k


class ChainTransform:
    def __init__(self, *args):
        self.lst_transforms = list(args)


class ChainTransformMasked:
    def __init__(self, mask, *args):
        self.mask = mask
        self.lst_transforms = list(args)


class LayerGradRev(torch.nn.Module):
    """utils/transforms.py:222-238.  The effective scale is the CONSTRUCTION-time alpha: the
    reference's train loop assigns `alpha_` but `forward` reads `_alpha`."""

    def __init__(self, revgrad=None, alpha=1.0):
        super().__init__()
        self.revgrad = revgrad
        self._alpha = torch.tensor(alpha, requires_grad=False)


FuncGradRev = None  # the GRL backward (-g * alpha) lives in the fused decoder kernel


# ------------------------------------------------------------------------------------------------
# priors (utils/priors.py)
# ------------------------------------------------------------------------------------------------
class MarginalDistribution:
    """utils/priors.py:7-36: independent per-dimension torch distributions."""

    def __init__(self, distributions):
        self.n_z = len(distributions)
        self.distributions = distributions

    def log_prob(self, z):
        cols = [d.log_prob(z[..., i]) for i, d in enumerate(self.distributions)]
        return torch.stack(cols, dim=-1)

    def sample(self, shape):
        z = torch.zeros((*shape, len(self.distributions)))
        for j, dist_j in enumerate(self.distributions):
            z[..., j] = dist_j.sample(shape).squeeze()
        return z.to(device)


def get_prior_dist(prior):
    return MarginalDistribution([item["dist"](**item["args"]) for item in prior.values()])


def get_shapes_from_dict(dict_gt):
    """utils/priors.py:53-61 -> (nz_x, nd_c, nd_y, nd_f, nd_p)."""
    vals = list(dict_gt.values())
    cnt = lambda t: sum(1 for v in vals if v["type"] == t)  # noqa: E731
    n_p = sum(1 for v in vals if v["phys"] is True and v["type"] == "c")
    return cnt("x"), cnt("c"), cnt("y"), cnt("f"), n_p


# ------------------------------------------------------------------------------------------------
# schedules, early stopping, logger
# ------------------------------------------------------------------------------------------------
class Annealing:
    """utils/annealing.py:6-52: none -> 1 ; cyclical ; sigmoid (Normal CDF)."""

    def __init__(self, type, n_iter, **kwargs):
        self.type = type
        self.n_iter = n_iter
        self.kwargs = kwargs

    def forward(self, iter):
        if (self.type is None) or (self.type == "none") or (self.type == "None"):
            return torch.tensor(1.0)
        if self.type == "cyclical":
            period = self.n_iter / self.kwargs["n_cycles"]
            tau = np.mod(iter, period) / period
            R = self.kwargs["R"]
            return torch.tensor(1.0 * (tau / R) if tau <= R else 1.0)
        if self.type == "sigmoid":
            mu_t = self.kwargs["mu"] * self.n_iter
            sigma_t = mu_t * self.kwargs["cov"]
            return dist.Normal(mu_t, sigma_t).cdf(torch.tensor(iter))
        raise ValueError(f"Invalid type {self.type}")


class EarlyStopping:
    """utils/loss.py:6-25."""

    def __init__(self, patience=1, min_delta=0):
        self.patience = patience
        self.min_delta = min_delta
        self.counter = 0
        self.min_validation_loss = float("inf")

    def early_stop(self, validation_loss):
        if validation_loss < (self.min_validation_loss - self.min_delta):
            self.min_validation_loss = validation_loss
            self.counter = 0
        elif validation_loss > self.min_validation_loss:
            self.counter += 1
            if self.counter >= self.patience:
                return True
        return False


def get_logger_training_curve(logger, label):
    """utils/loss.py:1-4."""
    items = logger.experiment.scalars[label]
    return [it[0] for it in items], [it[1] for it in items]


class _Experiment:
    def __init__(self, owner):
        self._owner = owner

    @property
    def scalars(self):
        self._owner.flush()
        return self._owner._scalars


class ScalarLogger:
    """Surface of torchrl.record.CSVLogger that dpivae.py / utils/loss.py use:
    `log_scalar(name, value, step)` and `experiment.scalars[name] -> [(step, float)]`.

    Device scalars are NOT synchronised when logged: values that are still CUDA tensors are kept
    as (step, tensor-view) and converted in one batched D2H copy when `experiment.scalars` is read
    (the reference pays one host sync per logged scalar, dpivae.py:439-451)."""

    def __init__(self, exp_name="", log_dir=None):
        self._scalars = {}
        self._pending = []  # (name, step, tensor)
        self.experiment = _Experiment(self)

    def log_scalar(self, name, value, step=None):
        if torch.is_tensor(value) and value.is_cuda:
            self._pending.append((name, step, value))
        else:
            self._scalars.setdefault(name, []).append((step, float(value)))

    def flush(self):
        if not self._pending:
            return
        vals = torch.stack([t.reshape(()).float() for _, _, t in self._pending]).cpu().tolist()
        for (name, step, _), v in zip(self._pending, vals):
            self._scalars.setdefault(name, []).append((step, v))
        self._pending = []


# ------------------------------------------------------------------------------------------------
# synthetic data (utils/data.py:9-52)
# ------------------------------------------------------------------------------------------------
def sample_response(definition, n, sample_dist=None, z=None):
    if (sample_dist is None) and (z is None):
        raise ValueError("At least one of `sample_dist` and `z` must not be `None`")
    if z is None:
        z_sample = sample_dist.sample((n,))
    else:
        z_sample = z.unsqueeze(0).repeat(n, 1, 1)
    z_sample = z_sample.to(device)
    dict_gt = definition["dict_gt"]
    z_idx_c = [i for i, v in enumerate(dict_gt.values()) if v["type"] == "c"]
    z_idx_y = [i for i, v in enumerate(dict_gt.values()) if v["type"] == "y"]
    with torch.no_grad():
        x_sample = definition["full_model"](z_sample)
    x_sample = x_sample + dist.Normal(0.0, definition["sigma_x"]).sample(x_sample.shape).to(x_sample.device)
    c_sample = z_sample[..., z_idx_c].reshape((*z_sample.shape[:-1], len(z_idx_c)))
    c_sample = c_sample + dist.Normal(0.0, definition["sigma_c"]).sample(c_sample.shape).to(c_sample.device)
    y_sample = z_sample[..., z_idx_y].reshape((*z_sample.shape[:-1], len(z_idx_y)))
    y_sample = y_sample + dist.Normal(0.0, definition["sigma_y"]).sample(y_sample.shape).to(y_sample.device)
    return x_sample.to(device), c_sample.to(device), y_sample.to(device), z_sample


LOG_2PI = math.log(2 * math.pi)
