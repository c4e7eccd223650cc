"""Load tests/golden/*.npz (outputs of the unmodified reference, see make_golden.py) into the
plain-dict form the oracle and the CUDA parity tests consume."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
ASSETS = os.path.join(HERE, "..", "dpivae_b200", "cases", "assets")
CONFIGS = [("bridge", "P"), ("bridge", "S"), ("damped_oscillator", "P"), ("damped_oscillator", "S"),
           ("simple_beam", "P"), ("simple_beam", "S")]
FW_NAMES = ["xh_p", "xh_d", "ch", "log_sigma_c", "yh", "log_sigma_y", "zx", "zc", "zy", "dens_z"]
L8_NAMES = ["loss", "KLx", "KLc", "KLy", "Rx", "Rc", "Ry", "reg"]


def physics_spec(case):
    a = np.load(os.path.join(ASSETS, f"{case}.npz"))
    if case == "bridge":
        n = int(a["part_n_layers"])
        return {"kind": "mlp", "w": [a[f"part_w{i}"] for i in range(n)], "b": [a[f"part_b{i}"] for i in range(n)],
                "in_mean": a["part_in_mean"], "in_std": a["part_in_std"]}
    if case == "damped_oscillator":
        return {"kind": "mass_spring", "t": a["t"]}
    return {"kind": "beam", "t": torch.linspace(0.0, 1.0, 32).numpy()}


EXT_CONFIGS = [("bridge", "P"), ("bridge", "S"), ("damped_oscillator", "P"), ("simple_beam", "S")]
TRAIN_LOG = ["ELBO", "KLx", "KLc", "KLy", "Rx", "Rc", "Ry", "reg", "lambda_x", "beta_x", "beta_c", "beta_y", "sigma_x"]
VAL_LOG = ["ELBO_val", "KLx_val", "KLc_val", "KLy_val", "Rx_val", "Rc_val", "Ry_val", "reg_val"]


FULLCOV_CONFIGS = [("bridge", "P"), ("simple_beam", "S")]


def load(case, mtype, ext=False):
    """ext=True: the second fixture set (make_golden_ext.py): n_mc = 8, cond / lambda_x / saturation / edge / flagged trajectory.
    ext="fullcov": the `--full_cov_prior True` fixtures (make_golden_fullcov.py)."""
    suffix = "_fullcov" if ext == "fullcov" else ("_ext" if ext else "")
    g = np.load(os.path.join(GOLDEN, f"{case}_{mtype}{suffix}.npz"))
    nz_x, nz_c, nz_y, nd_x, nd_c, nd_y = [int(v) for v in g["spec.dims"]]
    prior = [("uniform" if k == 0.0 else "normal", float(a), float(b)) for k, a, b in g["spec.prior_x"]]
    spec = {
        "case": case, "model_type": str(g["spec.model_type"]),
        "nz_x": nz_x, "nz_c": nz_c, "nz_y": nz_y, "nd_x": nd_x, "nd_c": nd_c, "nd_y": nd_y,
        "idx_c_phys": [int(i) for i in g["spec.idx_c_phys"]], "lambda_g0": float(g["spec.lambda_g0"]),
        "lambda_x": None, "lb": g["spec.lb"], "ub": g["spec.ub"], "prior_x": prior,
        "physics": physics_spec(case), "trainable": [str(k) for k in g["trainable"]],
        "full_cov_prior": bool(int(g["spec.full_cov_prior"])) if "spec.full_cov_prior" in g.files else False,
    }
    for k in ["mean_x", "std_x", "mean_c", "std_c", "mean_y", "std_y"]:
        spec[k] = g[f"spec.{k}"]
    sd = {k: torch.from_numpy(g[f"init.{k}"].copy()) for k in spec["trainable"]}
    return g, spec, sd


def state_of(g, spec, prefix):
    """Alternative weights of a fixture section (e.g. "sat.init", "edge.init")."""
    return {k: torch.from_numpy(g[f"{prefix}.{k}"].copy()) for k in spec["trainable"]}


def traj_flags(g):
    """The non-default flags the ext trajectory was run with -> dict of python values."""
    out = {}
    for item in g["traj.flags"]:
        k, v = str(item).split("=", 1)
        out[k] = True if v == "True" else (False if v == "False" else (v if not v.replace(".", "").replace("e-", "").replace("-", "").isdigit() else (float(v) if ("." in v or "e" in v) else int(v))))
    return out


def eps_of(g, spec, prefix="eps", start=0):
    if spec["model_type"] == "P":
        return tuple(torch.from_numpy(g[f"{prefix}{start + i}"]) for i in range(3))
    return torch.from_numpy(g[f"{prefix}{start}"])


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    den = float(b.norm())
    if den == 0.0:
        return float(a.norm())
    return float((a - b).norm()) / den
