"""Per-phase cycle breakdown of the fused decoder kernel (thread-0 clock64 accounting).

    python tools/phase_profile.py [workload] [rows]     # on a GPU box
"""
import contextlib
import ctypes as C
import importlib
import io
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dpivae_b200 as dpv  # noqa: E402
from dpivae_b200 import _lib  # noqa: E402

TC_NAMES = ["setup", "rowpar_eps", "latent", "aux_fwd", "phys_l0", "aux_bwd", "phys_l1", "fx_hidden", "phys_l2+x_mma", "x_head",
            "bwd1", "bwd2", "bwd3", "bwd4", "latent_bwd", "row_reduce", "row_out", "flush"]
NAMES = ["setup", "rowpar", "eps", "latent_fwd", "aux_fwd", "aux_loss", "aux_bwd", "phys_fwd", "data_fwd", "x_loss",
         "data_bwd", "phys_bwd", "latent_bwd", "row_reduce", "row_out"]


def main():
    wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "bridge_p"])
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    math = sys.argv[3] if len(sys.argv) > 3 else "fp32"
    case_mod = importlib.import_module(f"dpivae_b200.cases.{wl['case']}")
    dev = torch.device("cuda", 0)
    xs, cs, ys = bench.synth(case_mod, 4096, 7, dev)
    args = bench.make_args(case_mod, wl["preset"], use_seed=True, n_train=4096, n_batch=4096)
    with contextlib.redirect_stdout(io.StringIO()):
        vae = dpv.setup_model(args, case_mod.definition, (xs, cs, ys))
    x, c, y = bench.synth(case_mod, rows, 5, dev)
    eng = vae.engine()
    eng.set_groups(dpv.param_groups(args))
    eng.set_math_mode(math)
    w = (1.0, 1.0, 1.0, 1.0)
    for i in range(3):
        eng.loss(x, c, y, wl["n_mc"], w, True, adam_step=i + 1)
    buf = torch.zeros(32, dtype=torch.int64, device=dev)
    _lib.check(eng.lib.dpivae_set_phase_buffer(eng.handle, C.c_void_p(buf.data_ptr())))
    eng.loss(x, c, y, wl["n_mc"], w, True, adam_step=4)
    torch.cuda.synchronize()
    _lib.check(eng.lib.dpivae_set_phase_buffer(eng.handle, C.c_void_p(None)))
    v = buf.cpu().tolist()
    tot = sum(v)
    tile = 64 if math == "fp32" else 128
    nchunks = rows * wl["n_mc"] / tile
    print(f"workload {wl['case']} {wl['preset']} rows {rows} math {math}: {sum(v[:18]) / nchunks:.0f} cycles per {tile}-pair tile")
    if math != "fp32":
        tot = sum(v[:len(TC_NAMES)])
    for nme, c_ in zip(NAMES if math == "fp32" else TC_NAMES, v):
        print(f"  {nme:12s} {100.0 * c_ / tot:5.1f}%  {c_ / nchunks:9.0f} cyc/chunk")
    if math != "fp32":
        print(f"  MMA-issue warp: waiting for stage signals {v[20] / nchunks:9.0f} cyc/tile, issuing {v[21] / nchunks:9.0f} cyc/tile")
        print(f"  aux-decoder warps: waiting (records, operand-buffer hand-overs) {v[22] / nchunks:9.0f} cyc/tile, working {v[23] / nchunks:9.0f} cyc/tile")


if __name__ == "__main__":
    main()
