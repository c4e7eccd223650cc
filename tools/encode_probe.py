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

def phase_run(dbg):
    os.environ["DPIVAE_ENC_DBG"] = str(dbg)
    rows = 524288
    x, _, _ = bench.synth(case_mod, rows, 2000, dev)
    os.environ["DPIVAE_ENCODE_TRACE"] = "1"
    buf = torch.zeros(32 + 6 * 402, dtype=torch.int64, device=dev)
    _lib.check(eng.lib.dpivae_set_phase_buffer(eng.handle, C.c_void_p(buf.data_ptr())))
    eng.encode(x, 1, False)
    torch.cuda.synchronize()
    _lib.check(eng.lib.dpivae_set_phase_buffer(eng.handle, C.c_void_p(None)))
    v = buf.cpu().tolist()
    names = ["front (unit 0): wait L1", "front: wait A free", "front: relu", "front: wait X free + stage x + fetch", "latent: wait heads", "latent: load heads",
             "latent: sample + store", "issue: wait x", "issue: wait O free", "issue: wait relu (3 units)", "issue: head MMAs", "issue: L1 MMAs"]
    nt = max(v[12], 1)
    print(f"dbg={dbg} (1 skip x staging, 2 skip ReLU epilogue, 4 skip latent math): CTA 0, {nt} tiles of {rows} rows; cycles per tile:")
    for nme, c_ in zip(names, v):
        print(f"  {nme:40s} {c_ / nt:8.0f}")
    return v


for dbg in (1, 2, 3, 4, 7):
    phase_run(dbg)
v = phase_run(0)
# event timeline of CTA 0 (first tiles): per role, (cycle since the first event, event code)
roles = ["relu g0", "relu g1", "relu g2", "latent s0", "latent s1", "issue"]
ev = []
for ri, nme in enumerate(roles):
    t = v[32 + ri * 402: 32 + (ri + 1) * 402]
    n = int(t[400])
    ev += [(t[2 * i], nme, t[2 * i + 1]) for i in range(n)]
ev.sort()
t0 = ev[0][0] if ev else 0
legend = ("front: 1xx L1 done seen, 2xx A free seen, 3xx relu done, 4xx X free seen, 5xx staged, 6xx fetched | latent: 1xx heads seen, 2xx loaded, 3xx done | "
          "issue: 1xx x ready, 2xx O free, u3xx relu_u seen, u4xx head_u issued, u5xx L1_u issued (xx = tile)")
print(legend)
for t, nme, code in ev[:260]:
    print(f"{t - t0:8d}  {nme:10s} {code}")
