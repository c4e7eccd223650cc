"""Per-kernel CUDA-event times of one training step at the reference-default small shape (n_batch 64, n_mc 16).

    python tools/small_batch_kernel_times.py [case] [preset] [rows] [n_mc]     # on a GPU box
"""
import contextlib
import importlib
import io
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dpivae_b200 as dpv  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "damped_oscillator"
    preset = sys.argv[2] if len(sys.argv) > 2 else "dpivae"
    rows = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    case_mod = importlib.import_module(f"dpivae_b200.cases.{case}")
    dev = torch.device("cuda", 0)
    x, c, y = bench.synth(case_mod, 1024, 7, dev)
    args = bench.make_args(case_mod, preset, use_seed=True, n_train=1024, n_batch=rows)
    with contextlib.redirect_stdout(io.StringIO()):
        vae = dpv.setup_model(args, case_mod.definition, (x, c, y))
    eng = vae.engine()
    eng.set_groups(dpv.param_groups(args))
    for mode in ("tc_fp16x3", "fp32"):
        eng.set_math_mode(mode)
        idx = torch.randperm(1024, device=dev)[:rows]
        w = (1.0, 1.0, 1.0, 1.0)
        eng.set_timing(True)
        acc = {}
        for i in range(20):
            eng.loss(x, c, y, n, w, True, idx=idx, adam_step=i + 1)
            torch.cuda.synchronize()
            if i >= 5:
                for k, v in eng.last_kernel_ms().items():
                    acc[k] = acc.get(k, 0.0) + v / 15
        eng.set_timing(False)
        print(mode, {k: round(1e3 * v, 1) for k, v in acc.items()}, "us; sum", round(1e3 * sum(acc.values()), 1))


if __name__ == "__main__":
    main()
