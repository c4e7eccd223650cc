"""Encode-only kernel time vs batch size (set-up cost vs per-tile cost of enc_fused_kernel), CUDA events over back-to-back calls.

    python tools/encode_probe.py [rows ...]
"""
import contextlib
import importlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ctypes as C  # noqa: E402

import bench  # noqa: E402
from dpivae_b200 import _lib  # noqa: E402
import dpivae_b200 as dpv  # noqa: E402

case_mod = importlib.import_module("dpivae_b200.cases.bridge")
dev = torch.device("cuda:0")
xs, cs, ys = bench.synth(case_mod, 4096, 7, dev)
args = bench.make_args(case_mod, "DPIVAE-A", use_seed=True, seed=123, n_train=4096, n_batch=4096)
with contextlib.redirect_stdout(io.StringIO()):
    vae = dpv.setup_model(args, case_mod.definition, (xs, cs, ys))
eng = vae.engine()
eng.set_math_mode("tc_fp16x3")
rows_list = [int(r) for r in sys.argv[1:]] or [128, 148 * 128, 2 * 148 * 128, 8 * 148 * 128, 524288, 4 * 524288]
for rows in rows_list:
    x, _, _ = bench.synth(case_mod, rows, 2000, dev)
    for _ in range(5):
        eng.encode(x, 1, False)
    torch.cuda.synchronize()
    reps = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.encode(x, 1, False)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    tiles_per_cta = -(-rows // 128) / min(148, -(-rows // 128))
    print(f"rows {rows:8d}: {us:8.1f} us per call, {tiles_per_cta:6.1f} tiles per CTA, {rows / us * 1e-3:6.2f} G rows/s, "
          f"{rows * 300 / us * 1e-3:7.1f} GB/s algorithmic")

# per-role cycle accounting of CTA 0 (enc_fused_kernel's PH counters)
rows = 524288
x, _, _ = bench.synth(case_mod, rows, 2000, dev)
buf = torch.zeros(32, dtype=torch.int64, device=dev)
_lib.check(eng.lib.dpivae_set_phase_buffer(eng.handle, C.c_void_p(buf.data_ptr())))
eng.encode(x, 1, False)
torch.cuda.synchronize()
_lib.check(eng.lib.dpivae_set_phase_buffer(eng.handle, C.c_void_p(None)))
v = buf.cpu().tolist()
names = ["front: wait L1", "front: wait A free", "front: relu", "front: stage x", "latent: wait heads", "latent: load heads", "latent: sample + store",
         "issue: wait x", "issue: L1 MMAs", "issue: wait relu", "issue: wait O free", "issue: head MMAs", "(tiles)",
         "front: stage x: split + st.shared", "front: stage x: proxy fence + arrive", "front: relu: first tcgen05.ld"]
nt = max(v[12], 1)
print(f"CTA 0, {nt} tiles of {rows} rows; cycles per tile:")
for nme, c_ in zip(names, v):
    print(f"  {nme:24s} {c_ / nt:8.0f}")
