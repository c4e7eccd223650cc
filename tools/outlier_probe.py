"""NaN census of the gradients on rows far outside the scaler's range, per case / model type / math mode.

    python tools/outlier_probe.py
"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_from_golden  # noqa: E402


def main():
    for case in ("bridge", "simple_beam", "damped_oscillator"):
        for mtype in ("P", "S"):
            g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
            eng = vae.engine()
            for scale in (3.0, 30.0, 300.0):
                X, C_, Y = (torch.cat([t] * 40, 0).cuda() for t in (x, c, y))
                sdv = x.std(0, keepdim=True).cuda() + 1e-12
                X = X + scale * sdv * torch.randn(X.shape, generator=torch.Generator().manual_seed(1)).cuda()
                res = []
                for mode in ("fp32", "tc_fp16x3"):
                    eng.set_math_mode(mode)
                    torch.manual_seed(123)
                    rl, s = eng.loss(X, C_, Y, 16, (1.0, 1.0, 1.0, 1.0), True)
                    res.append((mode, int(torch.isnan(eng.grads).sum()), int((~torch.isfinite(eng.grads)).sum()), int((~torch.isfinite(rl)).sum())))
                print(case, mtype, f"{scale:g} sigma:", "; ".join(f"{m}: NaN grads {a}, non-finite grads {b}, non-finite row losses {c_}" for m, a, b, c_ in res), flush=True)


if __name__ == "__main__":
    main()
