"""Export the reference's pretrained surrogate weights + scaler statistics to small .npz assets.

Container-only (reads /root/reference through tools/ref_harness.py).  The assets are DATA
(pretrained weights `cases/*/full_model`, `cases/bridge/part_model`, and the mean / population
std of `cases/*/X*.pt` that the reference's StandardScaler.fit computes at import time,
cases/bridge/__init__.py:134-171); they are what lets dpivae_b200.cases reproduce the reference's
`definition["full_model"]` / `definition["part_model"]` on a box without /root/reference.

    python tools/export_case_assets.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "dpivae_b200", "cases", "assets")
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402


def mlp_arrays(prefix, mlp):
    out = {}
    lin = [m for m in mlp.net if hasattr(m, "weight")]
    out[f"{prefix}_n_layers"] = np.array(len(lin), dtype=np.int64)
    for i, l in enumerate(lin):
        out[f"{prefix}_w{i}"] = l.weight.detach().cpu().numpy().astype(np.float32)
        out[f"{prefix}_b{i}"] = l.bias.detach().cpu().numpy().astype(np.float32)
    out[f"{prefix}_in_mean"] = mlp.input_transform.mean_.detach().cpu().numpy().astype(np.float32).reshape(-1)
    out[f"{prefix}_in_std"] = mlp.input_transform.scale_.detach().cpu().numpy().astype(np.float32).reshape(-1)
    return out


def main():
    out_dir = os.path.abspath(OUT)
    os.makedirs(out_dir, exist_ok=True)
    for name in ["bridge", "damped_oscillator", "simple_beam"]:
        _, case = ref_harness.load(name)
        d = case.definition
        arrs = mlp_arrays("full", d["full_model"])
        if name == "bridge":
            arrs.update(mlp_arrays("part", d["part_model"]))
        arrs["t"] = d["t"].detach().cpu().numpy().astype(np.float32)
        path = os.path.join(out_dir, f"{name}.npz")
        np.savez_compressed(path, **arrs)
        print(name, {k: v.shape for k, v in arrs.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
