"""Reference-default training run (0_single_run.py shape: n_train 1024, n_batch 64, n_mc 16, val_freq 10) through
train_model: wall time per iteration with the device-resident loop vs the per-iteration host loop, and the ELBO curve.

    python tools/train_default_demo.py [case] [preset] [n_iter]
"""
import contextlib
import importlib
import io
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dpivae_b200 as dpv  # noqa: E402
from helpers import make_args  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "damped_oscillator"
    preset = sys.argv[2] if len(sys.argv) > 2 else "vae"
    n_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    case_mod = importlib.import_module(f"dpivae_b200.cases.{case}")
    d = case_mod.definition
    for dl in (True, False):
        torch.manual_seed(123)
        tr = dpv.sample_response_device(d, 1024)
        va = dpv.sample_response_device(d, 512)
        args = make_args(case_mod, preset, use_seed=True, seed=123, n_iter=n_iter, device_loop=dl)
        with contextlib.redirect_stdout(io.StringIO()):
            vae = dpv.setup_model(args, d, tr)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vae, logger = dpv.train_model(args, vae, d, tr, va)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e = logger.experiment.scalars["ELBO"]
        ev = logger.experiment.scalars["ELBO_val"]
        print(f"{case} {preset} device_loop={dl}: {len(e)} iterations in {dt:.2f} s = {1e6 * dt / len(e):.0f} us/iteration "
              f"({64 * len(e) / dt:.0f} datapoints/s); ELBO {e[0][1]:.3f} -> {e[-1][1]:.3f}; ELBO_val {ev[0][1]:.3f} -> {ev[-1][1]:.3f}")


if __name__ == "__main__":
    main()
