"""lat_fwd (noise pre-pass + pair kernel) time of an unsharded call vs a cyclic / contiguous row shard of a larger global
batch on ONE GPU (the shard's kernels do not depend on the other ranks).   python tools/noise_probe.py [rows]"""
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
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    case_mod = importlib.import_module("dpivae_b200.cases.bridge")
    dev = torch.device("cuda", 0)
    x, c, y = bench.synth(case_mod, rows, 7, dev)
    args = bench.make_args(case_mod, "DPIVAE-A", use_seed=True, n_train=rows, n_batch=rows)
    with contextlib.redirect_stdout(io.StringIO()):
        vae = dpv.setup_model(args, case_mod.definition, (x, c, y))
    eng = vae.engine()
    eng.set_groups(dpv.param_groups(args))
    eng.set_math_mode("tc_fp16x3")
    w = (1.0, 1.0, 1.0, 1.0)
    for name, kw in (("unsharded", {}), ("cyclic S=2", dict(B_global=2 * rows, row_offset=1, row_stride=2)),
                     ("cyclic S=8", dict(B_global=8 * rows, row_offset=5, row_stride=8)),
                     ("contiguous 1/8", dict(B_global=8 * rows, row_offset=5 * rows))):
        eng.set_timing(True)
        acc = {}
        for i in range(12):
            eng.loss(x, c, y, 16, w, True, **kw)
            torch.cuda.synchronize()
            if i >= 4:
                for k, v in eng.last_kernel_ms().items():
                    acc[k] = acc.get(k, 0.0) + v / 8
        eng.set_timing(False)
        print(f"{name:16s} lat_fwd {1e3 * acc['lat_fwd']:7.1f} us  lat_bwd {1e3 * acc['lat_bwd']:7.1f} us")


if __name__ == "__main__":
    main()
