#!/usr/bin/env python
"""Headline benchmark: DPI-VAE training step (gather-free fwd + ELBO + bwd + Adam) datapoints/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (BASELINE.json configs):
  bridge_p  (default) bridge case, DPIVAE-A (P) preset, 131,072 rows per GPU x 16 MC samples, fp32.
            At N GPUs the global minibatch is N x 131,072 (N=8 -> config 3's 1,048,576 rows), rows
            sharded, ONE NCCL allreduce of [grads | scalars] per step -> "scaling": "weak".
  beam_s    simple_beam dpivae (S) preset, 65,536 rows x 16 MC (config 2).
  bridge_encode   bridge encode-only inference (transform_inputs -> encode, n = 1), 524,288 rows per GPU (config 5:
            4,194,304 rows over 8 GPUs), communication-free replicas; HBM-bound, 300 B per row.
  ensemble  independent small trainings (damped_oscillator dpivae, seeds x lambda_g0 of 1_disentanglement_metric.py,
            reference-default shape n_train 1024 / n_batch 64 / n_mc 16), 8 members per GPU on 8 CUDA streams, no
            communication (config 4); value = members x 64 datapoints per round / time.
One JSON line on stdout (rank 0).  `--impl reference` times the reference algorithm on the host CPU
(the oracle port, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    "bridge_p": dict(case="bridge", preset="DPIVAE-A", rows=131072, n_mc=16, cpu_rows=8192, ref_rows=65536,
                     flop_step=1_597_440, flop_dec=1_519_616, bytes_row=296),
    "beam_s": dict(case="simple_beam", preset="dpivae", rows=65536, n_mc=16, cpu_rows=16384, ref_rows=65536,
                   flop_step=548_352, flop_dec=2 * 16 * ((4 * 128 + 128 * 32) * 3 + (2 * 64 + 128) * 2 * 3), bytes_row=160),
}
MATH_DOC = {
    "tc_fp16x3": "decoder GEMMs on tcgen05: fp32 operands split into fp16 hi+lo planes, 3 MMAs per GEMM, fp32 TMEM accumulators "
                 "(parity 1e-5 vs the fp64 oracle, tests/test_gpu_tc.py); P-model encoders on tcgen05 the same way, latent math / ELBO / Adam fp32 on the CUDA cores",
    "fp32": "all GEMMs fp32 FFMA on the CUDA cores",
    "tc_fp16": "decoder GEMMs on tcgen05 with plain fp16 operands, fp32 accumulate (tolerance 2e-3 loss / 2e-2 gradients)",
}
# bridge, single-encoder preset (S): same decoders as bridge_p, one 128-wide FullCovarianceNN encoder over all 10 latents
WORKLOADS["bridge_s"] = dict(case="bridge", preset="DPIVAE-B", rows=131072, n_mc=16, cpu_rows=8192, ref_rows=65536,
                             flop_step=1_651_712, flop_dec=1_519_616, bytes_row=296)
WORKLOADS["bridge_encode"] = dict(case="bridge", preset="DPIVAE-A", rows=524288, n_mc=1, bytes_row=300, flop_row=31_744)
WORKLOADS["ensemble"] = dict(case="damped_oscillator", preset="dpivae", rows=64, n_mc=16, members=8, inner_steps=16)
ENCODE_KERNEL_DOC = {
    "fp32": "enc_fwd_kernel (fp32 FFMA) + lat_encode_kernel",
    "tc_fp16x3": "noise_fill_kernel + enc_fused_kernel (tcgen05 encoder MMAs + latent sampling, warp-specialised, one launch per call)",
    "tc_fp16": "noise_fill_kernel + enc_fused_kernel (tcgen05 encoder MMAs + latent sampling, warp-specialised, one launch per call)",
}
METRIC = "ELBO train samples/s (fwd+bwd+Adam)"
UNIT = "datapoints/s"


_JSON_FD = None


def _guard_stdout():
    """Library chatter on fd 1 (e.g. NCCL's version banner) goes to stderr: stdout carries the ONE JSON line only."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def synth(case_mod, n, gen_seed, device):
    """Synthetic minibatch of the case's shape: z ~ ground-truth priors, x = full_model(z) + noise (utils/data.py:9-52).
    On a CUDA device: the on-device generator (dpivae_sample_response: Philox draws + surrogate MLP in hand-written
    kernels, nothing crosses PCIe); on the CPU (the reference arm): the torch path of the mirror."""
    import torch
    from dpivae_b200 import get_prior_dist, sample_response, sample_response_device

    torch.manual_seed(gen_seed)
    d = case_mod.definition
    if torch.device(device).type == "cuda":
        x, c, y, _ = sample_response_device(d, n, device)
        return x, c, y
    x, c, y, _ = sample_response(d, n, sample_dist=get_prior_dist(d["dict_gt"]))
    return x.to(device).float().contiguous(), c.to(device).float().contiguous(), y.to(device).float().contiguous()


def make_args(case_mod, preset, **over):
    from dpivae_b200 import make_parser

    args, _ = make_parser().parse_known_args([])
    for k, v in case_mod.presets[preset].items():
        setattr(args, k, v)
    for k, v in over.items():
        setattr(args, k, v)
    return args


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_throughput(wl, steps, warmup, threads=None, device="cpu", rows=None):
    """Reference algorithm on the host cores: oracle port (fp32 torch CPU tensor algebra) of
    loss -> backward -> Adam on a bounded sample of the workload."""
    import importlib

    import torch
    import golden_util as gu
    from oracle import dpivae_oracle as orc
    import dpivae_b200 as dpv

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    case_mod = importlib.import_module(f"dpivae_b200.cases.{wl['case']}")
    rows, n = rows or wl["cpu_rows"], wl["n_mc"]
    x, c, y = synth(case_mod, rows, 123, "cpu")
    args = make_args(case_mod, wl["preset"], n_train=rows, n_batch=rows)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        vae = dpv.setup_model(args, case_mod.definition, (x, c, y))
    sd = {k: v.detach().clone() for k, v in vae.state_dict().items() if not k.startswith("decoder_x.model.")}
    vec = lambda t: t.detach().reshape(-1).numpy()
    dpx = case_mod.definition["dict_prior_x"]
    prior = [("uniform", float(d.low), float(d.high)) if isinstance(d, torch.distributions.Uniform)
             else ("normal", float(d.loc), float(d.scale)) for d in vae.prior_x.distributions]
    spec = {"model_type": args.model_type, "nz_x": vae.nz_x, "nz_c": vae.nz_c, "nz_y": vae.nz_y, "nd_x": vae.nd_x,
            "nd_c": vae.nd_c, "nd_y": vae.nd_y, "idx_c_phys": list(vae.idx_c_phys), "lambda_g0": args.lambda_g0,
            "lambda_x": None, "lb": [v["lb"] for v in dpx.values()], "ub": [v["ub"] for v in dpx.values()],
            "prior_x": prior, "physics": gu.physics_spec(wl["case"]), "trainable": list(sd.keys()),
            "mean_x": vec(vae.transform_x.mean_), "std_x": vec(vae.transform_x.scale_), "mean_c": vec(vae.transform_c.mean_),
            "std_c": vec(vae.transform_c.scale_), "mean_y": vec(vae.transform_y.mean_), "std_y": vec(vae.transform_y.scale_)}
    spec = orc.cast_spec(spec, torch.float32)
    if device != "cpu":
        # the same torch tensor algebra as eager CUDA ops: the "existing Blackwell path" of SURVEY.md §8(d)
        def mv(o):
            if torch.is_tensor(o):
                return o.to(device)
            if isinstance(o, dict):
                return {k: mv(v) for k, v in o.items()}
            if isinstance(o, (list, tuple)):
                return type(o)(mv(v) for v in o)
            return o
        spec = mv(spec)
        sd = {k: v.to(device).requires_grad_(v.requires_grad) for k, v in sd.items()}
        x, c, y = x.to(device), c.to(device), y.to(device)
    names = spec["trainable"]
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    lr = {k: (5e-3 if k == "log_sigma_x" else 1e-3) for k in names}
    wd = {k: 0.0 for k in names}
    g = torch.Generator(device=device).manual_seed(0)
    widths = (vae.nz_x, vae.nz_c, vae.nz_y)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if device != "cpu":
            torch.cuda.synchronize()
        if args.model_type == "P":
            eps = tuple(torch.randn(n, rows, k, generator=g, device=device) for k in widths)
        else:
            eps = torch.randn(n, rows, sum(widths), generator=g, device=device)
        _, _, _, grads = orc.loss_and_grads(sd, spec, x, c, y, eps)
        orc.adam_step({k: sd[k] for k in names}, grads, m, v, it + 1, lr, wd)
        if device != "cpu":
            torch.cuda.synchronize()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return rows / sec, sec, threads, rows


def reference_available():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_harness

    return ref_harness.available()


def reference_throughput(wl, steps, warmup, rows, threads=None):
    """The UNMODIFIED reference (baseline/_ref, or /root/reference in the build container) on the host cores: its own
    `setup_model` and `train_model` (dpivae.py:89,285) on `rows` rows x n_mc samples per iteration, n_batch = n_train
    (the reference's own multinomial draw then returns a permutation).  Per-iteration wall time is read off the
    reference's one `torch.multinomial` call per iteration (dpivae.py:403); the first `warmup` iterations are dropped.
    Import shims only (tools/ref_harness.py); `torch.cuda.is_available` is masked while the reference's modules are
    imported so that utils/__init__.py:5 picks the CPU."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_harness

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cwd = os.getcwd()
    avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        dp, case = ref_harness.load(wl["case"])
        from utils.data import sample_response as ref_sample_response
        from utils.priors import get_prior_dist as ref_get_prior_dist
    finally:
        torch.cuda.is_available = avail
    try:
        definition = case.definition
        args = ref_harness.make_args(case, wl["preset"], use_seed=True, seed=123, n_train=rows, n_batch=rows, n_val=8,
                                     n_mc_train=wl["n_mc"], n_mc_val=1, n_iter=warmup + steps, val_freq=10 ** 9)
        torch.manual_seed(123)
        prior = ref_get_prior_dist(definition["dict_gt"])
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            data = ref_sample_response(definition, rows, sample_dist=prior)
            data_val = ref_sample_response(definition, 8, sample_dist=prior)
            vae = dp.setup_model(args, definition, data)
            stamps = []
            orig = torch.multinomial

            def tap(*a, **k):
                stamps.append(time.perf_counter())
                return orig(*a, **k)

            torch.multinomial = tap
            try:
                dp.train_model(args, vae, definition, data, data_val)
            finally:
                torch.multinomial = orig
            stamps.append(time.perf_counter())
    finally:
        os.chdir(cwd)
    per = [b - a for a, b in zip(stamps[:-1], stamps[1:])][warmup:]
    sec = sum(per) / len(per)
    return rows / sec, sec, threads, rows


def workload_config(a, wl, n_gpus):
    """`config` of the bench line: shared by both arms (the reference arm times a bounded sample of this workload)."""
    rows = wl["rows"]
    return {"workload": a.workload, "math": MATH_DOC[a.math], "case": wl["case"], "preset": wl["preset"], "rows_per_gpu": rows,
            "global_batch": rows * n_gpus if a.scaling == "weak" else rows, "n_mc": wl["n_mc"], "parallelism": f"dp{n_gpus}",
            "row_sharding": "cyclic (rank k of N owns global rows k, k + N, ...)" if n_gpus > 1 else "none",
            "minibatch_order": "identity (loss is a row sum; the reference's CPU multinomial draw is hoisted)",
            "data_generator": "on-device dpivae_sample_response (torch-stream Philox + surrogate MLP kernels)",
            "l2": "per-step working set (inputs + activations workspace) ~0.3 GB > 126 MB L2, no flush",
            "noise": "in-kernel Philox4x32-10, torch.cuda normal_ stream"}


def run_reference(a, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = a.ref_rows or wl.get("ref_rows", wl["cpu_rows"])
    if reference_available():
        val, sec, threads, rows = reference_throughput(wl, a.steps, a.warmup, rows)
        kind = "reference"
        what = "the reference's own setup_model + train_model (unmodified code, baseline/_ref), fp32 torch CPU"
    else:
        val, sec, threads, rows = cpu_oracle_throughput(wl, a.steps, a.warmup, rows=rows)
        kind = "port"
        what = "oracle port (baseline/_ref not installed), fp32 torch CPU"
    sample = f"{rows} rows x {wl['n_mc']} MC of the {wl['case']} {wl['preset']} step per timed step; {what}"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a, wl, a.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def _dist_setup():
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the host baseline")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local_rank, dev


def _timed_region(fn, steps, world, dev):
    """barrier + synchronize, CUDA events around `steps` calls, max over ranks (ms total)."""
    import torch
    import torch.distributed as dist

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def run_encode(a, wl, ctx):
    """Config 5: encode-only inference, rows sharded as independent replicas (no collective)."""
    import contextlib, importlib, io

    import torch

    import dpivae_b200 as dpv

    world, rank, local_rank, dev = ctx
    case_mod = importlib.import_module(f"dpivae_b200.cases.{wl['case']}")
    rows = wl["rows"]
    xs, cs, ys = synth(case_mod, 4096, 7, dev)
    args = make_args(case_mod, wl["preset"], use_seed=True, seed=123, n_train=4096, n_batch=4096)
    with contextlib.redirect_stdout(io.StringIO()):
        vae = dpv.setup_model(args, case_mod.definition, (xs, cs, ys))
    x, _, _ = synth(case_mod, rows, 2000 + rank, dev)
    eng = vae.engine()
    eng.set_math_mode(a.math)
    torch.manual_seed(5)
    xh = x.cpu().pin_memory()
    xd = torch.empty_like(x)
    out_host = [torch.empty((1, rows, k), dtype=torch.float32).pin_memory() for k in (vae.nz_x, vae.nz_c, vae.nz_y)]

    def step_resident(i):
        eng.encode(x, 1, False)

    def step_e2e(i):
        xd.copy_(xh, non_blocking=True)
        zx, zc, zy, dens = eng.encode(xd, 1, False)
        for h, t in zip(out_host, (zx, zc, zy)):
            h.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(a.warmup):
        step_resident(i)
    l0 = eng.launches
    with ClockSampler(local_rank) as clk:
        ms = _timed_region(step_resident, a.steps, world, dev) / a.steps
        launches = eng.launches - l0
        step_e2e(0)
        ms_e2e = _timed_region(step_e2e, a.steps, world, dev) / a.steps
    line = None
    if rank == 0:
        peaks = load_peaks()
        hbm = peaks.get("hbm_gbs", 6650.0)
        achieved = wl["bytes_row"] * rows / (ms * 1e-3) / 1e9
        line = {"metric": "encode-only datapoints/s (transform_inputs -> encode, n=1)", "value": rows * world / (ms * 1e-3),
                "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": a.workload, "case": wl["case"], "preset": wl["preset"], "rows_per_gpu": rows, "n_mc": 1,
                           "parallelism": f"replicas x{world} (no collective)", "l2": "inputs 134 MB per GPU > 126 MB L2, no flush"},
                "e2e": {"value": rows * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(rows * vae.nd_x * 4), "d2h_bytes_per_step": int(rows * 10 * 4)},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": ENCODE_KERNEL_DOC.get(a.math, ENCODE_KERNEL_DOC["fp32"]), "achieved": achieved, "peak": hbm,
                             "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                             "peak_source": "measured copy bandwidth (MEASURED_PEAKS.json)" if peaks else "fallback 6.65 TB/s"},
                "clocks": clk.summary()}
    del eng, vae
    torch.cuda.empty_cache()
    return line


def run_ensemble(a, wl, ctx):
    """Config 4: independent small trainings, several members per GPU on separate CUDA streams, no communication."""
    import contextlib, importlib, io

    import torch

    import dpivae_b200 as dpv

    world, rank, local_rank, dev = ctx
    case_mod = importlib.import_module(f"dpivae_b200.cases.{wl['case']}")
    M, n, nb = wl["members"], wl["n_mc"], wl["rows"]
    M = int(os.environ.get("DPIVAE_BENCH_MEMBERS", M))
    lambdas = [1e4, 1e3, 1e2, 1e1, 1e0, 0.0, -1e0, -1e1, -1e2, -1e3, -1e4]  # 1_disentanglement_metric.py:56 (/1e4)
    members = []
    for m in range(M):
        gm = rank * M + m
        x, c, y = synth(case_mod, 1024, 100 + gm, dev)
        args = make_args(case_mod, wl["preset"], use_seed=True, seed=gm, n_train=1024, n_batch=nb,
                         lambda_g0=lambdas[gm % len(lambdas)] / 1e4)
        with contextlib.redirect_stdout(io.StringIO()):
            vae = dpv.setup_model(args, case_mod.definition, (x, c, y))
        eng = vae.engine()
        eng.set_groups(dpv.param_groups(args))
        eng.set_math_mode(a.math)
        gcpu = torch.Generator().manual_seed(gm)
        pool = torch.stack([torch.multinomial(torch.ones(1024), nb, False, generator=gcpu) for _ in range(64)]).to(dev)
        members.append(dict(eng=eng, x=x, c=c, y=y, pool=pool, stream=torch.cuda.Stream(dev), step=0))
    w = (1.0, 1.0, 1.0, 1.0)
    inner = wl["inner_steps"]
    # device-resident loop: each member's training step is ONE captured CUDA graph (advance -> fwd -> bwd -> Adam) whose
    # step counter / generator offset / minibatch row live on the device; a bench step replays it `inner` times per member
    torch.cuda.synchronize()
    for mb in members:
        with torch.cuda.stream(mb["stream"]):
            mb["graph"] = mb["eng"].step_graph(mb["x"], mb["c"], mb["y"], n, w, idx_pool=mb["pool"], log_cap=64, unroll=inner)
    torch.cuda.synchronize()

    def round_(i):
        for mb in members:
            with torch.cuda.stream(mb["stream"]):
                mb["graph"].run(inner)

    def round_synced(i):
        cur = torch.cuda.current_stream()
        for mb in members:
            mb["stream"].wait_stream(cur)
        round_(i)
        for mb in members:
            cur.wait_stream(mb["stream"])

    for i in range(a.warmup):
        round_synced(i)
    l0 = sum(mb["eng"].launches for mb in members)
    with ClockSampler(local_rank) as clk:
        ms = _timed_region(round_synced, a.steps, world, dev) / a.steps
    launches = sum(mb["eng"].launches for mb in members) - l0
    line = None
    if rank == 0:
        line = {"metric": "ensemble train samples/s (independent small models, fwd+bwd+Adam)", "value": M * nb * inner * world / (ms * 1e-3),
                "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": a.workload, "case": wl["case"], "preset": wl["preset"], "members_per_gpu": M, "n_batch": nb,
                           "n_mc": n, "n_train": 1024, "math": a.math, "optimizer_steps_per_member_per_bench_step": inner,
                           "loop": "device-resident: one captured CUDA graph per member step, replayed without host arguments",
                           "parallelism": f"{M * world} independent members, no collective"},
                "gpu_launches": launches, "clocks": clk.summary(),
                "elbo": [float(mb["eng"].scalars[0]) for mb in members]}
    for mb in members:
        mb["graph"].close()
    members.clear()
    torch.cuda.empty_cache()
    return line


def rank_check(a, wl, case_mod, vae_factory, world, rank, dev, n, w):
    """N-rank correctness of the data-parallel step on hardware: one optimizer step of one global
    batch, sharded over the ranks with the NCCL allreduce, against the same step of the WHOLE global batch on rank 0
    alone -- once with the fp32 kernels (sharding only changes the order of fp32 partial sums: 1e-6 on the allreduced
    gradient) and once in the bench's arithmetic mode (tensor-core hi/lo split: tile composition and operand scales
    follow the shard, stated tolerance 5e-5).  Parameters must be BITWISE equal across the ranks; against the 1-rank run
    the Adam update is compared (relative L2 of p_after - p_before: Adam's m / sqrt(v) turns round-off on near-zero
    gradients into O(lr) differences on those few elements, so the bar there is 1e-2)."""
    import torch
    import torch.distributed as dist

    import dpivae_b200 as dpv
    from dpivae_b200.parallel import DataParallelStep

    rows = min(wl["rows"], 16384)
    Bg = rows * world
    shards = [synth(case_mod, rows, 5000 + r, dev) for r in ([rank] if rank else range(world))]
    mine = shards[0]
    gen = torch.cuda.default_generators[dev.index]
    out = {"ranks": world, "global_rows": Bg, "steps": 1, "modes": {}}
    ok_all = True
    for mode, gtol in (("fp32", 1e-6), (a.math, 5e-5)):
        if mode in out["modes"]:
            continue
        vae, args = vae_factory()
        dp = DataParallelStep(vae, dpv.param_groups(args))
        dp.eng.set_math_mode(mode)
        p0 = dp.eng.params.clone()
        torch.manual_seed(4242)
        off0 = gen.get_offset()
        dp.step(mine[0], mine[1], mine[2], n, w, Bg, rank, 1, row_stride=world)   # cyclic shards, as in the timed run
        grads_n = dp.eng.gradbuf.clone()
        params_n = dp.eng.params.clone()
        # every rank must hold bitwise identical parameters (same allreduced gradient, same fused Adam)
        lo, hi = params_n.clone(), params_n.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            raise RuntimeError("data-parallel ranks hold different parameters after the allreduced Adam steps")
        cks = torch.tensor([float(params_n.double().sum())], dtype=torch.float64, device=dev)
        allck = [torch.zeros_like(cks) for _ in range(world)]
        dist.all_gather(allck, cks)
        if rank == 0:
            vae1, args1 = vae_factory()
            eng1 = vae1.engine()
            eng1.set_groups(dpv.param_groups(args1))
            eng1.set_math_mode(mode)
            torch.manual_seed(4242)   # (the model factory re-seeded the generator for its scaler sample)
            gen.set_offset(off0)
            X, C_, Y = (torch.stack([sh[i] for sh in shards], dim=1).flatten(0, 1) for i in range(3))   # global row k + N t = rank k, local row t
            eng1.loss(X, C_, Y, n, w, True, B_global=Bg, row_offset=0, adam_step=1)
            g1, p1 = eng1.gradbuf, eng1.params
            gerr = float((grads_n.double() - g1.double()).norm() / g1.double().norm())
            uerr = float(((params_n - p0).double() - (p1 - p0).double()).norm() / (p1 - p0).double().norm())
            ok = bool(gerr < gtol and uerr < 1e-2)
            ok_all = ok_all and ok
            out["modes"][mode] = {"params_bitwise_equal_across_ranks": True, "param_checksums": [float(t) for t in allck],
                                  "grad_rel_l2_vs_1rank": gerr, "adam_update_rel_l2_vs_1rank": uerr,
                                  "tolerance": {"grad_rel_l2": gtol, "adam_update_rel_l2": 1e-2}, "ok": ok}
            del eng1, vae1
        else:
            out["modes"][mode] = None
        del dp, vae
        torch.cuda.empty_cache()
        dist.barrier()
    out["ok"] = ok_all
    return out if rank == 0 else None


def run_train(a, wl, ctx, sub=False):
    """Training-step workloads (bridge_p, beam_s).  Returns the bench line (rank 0) or None."""
    import importlib

    import torch
    import torch.distributed as dist

    import dpivae_b200 as dpv
    from dpivae_b200.parallel import DataParallelStep

    world, rank, local_rank, dev = ctx
    n_gpus = world
    case_mod = importlib.import_module(f"dpivae_b200.cases.{wl['case']}")
    n = wl["n_mc"]
    if a.scaling == "strong" and not sub:
        B_global = wl.get("strong_rows", 1048576)
        rows = B_global // n_gpus
    else:
        rows = wl["rows"]
        B_global = rows * n_gpus
    wl = dict(wl, rows=rows)
    # cyclic row shards: rank k of N owns the global rows k, k + N, ... (each rank then owns whole Philox evaluations of the
    # noise stream); the global batch is the interleaved union of the shards
    row_off, row_stride = (rank, world) if world > 1 else (0, 1)
    import contextlib, io

    def vae_factory():
        # scalers are fitted on a fixed global sample so every rank builds the identical model
        xs, cs, ys = synth(case_mod, 4096, 7, dev)
        args = make_args(case_mod, wl["preset"], use_seed=True, seed=123, n_train=4096, n_batch=4096)
        with contextlib.redirect_stdout(io.StringIO()):
            return dpv.setup_model(args, case_mod.definition, (xs, cs, ys)), args

    vae, args = vae_factory()
    x, c, y = synth(case_mod, rows, 1000 + rank, dev)  # this rank's shard, resident in HBM
    dp = DataParallelStep(vae, dpv.param_groups(args))
    eng = dp.eng
    eng.set_math_mode(a.math)
    w = (1.0, 1.0, 1.0, 1.0)
    torch.manual_seed(99)  # same Philox seed/offset on every rank; noise indexed by global row

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident(i):
        dp.step(x, c, y, n, w, B_global, row_off, i, row_stride=row_stride)

    # End-to-end arm: every step's inputs come from pinned host memory and its 8 loss scalars go back to the host, where
    # the caller waits for them.  Two device input sets: while step i computes, a copy stream uploads the inputs of step
    # i + 1 (the per-step H2D copy stays inside the timed region, it is just not serialised with the kernels).  The host
    # holds TWO different minibatches which alternate, and every step gathers its rows through a fresh int64 index
    # vector (the reference's `x_train[sample_idx]`, dpivae.py:403-404) that is uploaded with the inputs.
    xh, ch, yh = (t.cpu().pin_memory() for t in (x, c, y))
    host_sets = [(xh, ch, yh), tuple(t.flip(0).contiguous().pin_memory() for t in (xh, ch, yh))]
    gcpu = torch.Generator().manual_seed(17 + rank)
    idx_host = [torch.randperm(rows, generator=gcpu).pin_memory() for _ in range(2)]
    scal_host = torch.empty(8, dtype=torch.float32).pin_memory()
    dev_in = [tuple(torch.empty_like(t) for t in (x, c, y)) + (torch.empty(rows, dtype=torch.int64, device=dev),) for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"k": 0, "primed": False}

    def upload(slot, k):
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(dev_in[slot], host_sets[k & 1] + (idx_host[k & 1],)):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e(i):
        k = e2e_state["k"]
        if not e2e_state["primed"]:
            upload(k & 1, k)
            e2e_state["primed"] = True
        main = torch.cuda.current_stream()
        main.wait_event(ready[k & 1])
        xd, cd, yd, idxd = dev_in[k & 1]
        s = dp.step(xd, cd, yd, n, w, B_global, row_off, i, idx=idxd, row_stride=row_stride)
        upload((k + 1) & 1, k + 1)        # step k - 1, the last reader of that slot, was synchronised below
        scal_host.copy_(s, non_blocking=True)
        main.synchronize()         # the user reads the loss every step
        e2e_state["k"] = k + 1

    def timed(fn, steps, first_step):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(first_step + i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    step_no = 1
    for i in range(a.warmup):
        step_resident(step_no)
        step_no += 1
    l0 = eng.launches
    sustained = None
    with ClockSampler(local_rank) as clk:
        ms_total = timed(step_resident, a.steps, step_no)
        launches = eng.launches - l0
        step_no += a.steps
        ms_step = ms_total / a.steps
        value = B_global / (ms_step * 1e-3)

        # end-to-end arm: host buffers in, loss scalars out, every step
        for i in range(2):
            step_e2e(step_no)
            step_no += 1
        ms_e2e = timed(step_e2e, a.steps, step_no) / a.steps
        step_no += a.steps
        e2e_val = B_global / (ms_e2e * 1e-3)

        if not sub and a.sustain_s > 0:
            # the same resident step back to back for >= sustain_s seconds: clocks sampled under sustained load
            k_sus = max(a.steps, int(a.sustain_s * 1e3 / ms_step) + 1)
            ms_sus = timed(step_resident, k_sus, step_no) / k_sus
            step_no += k_sus
            sustained = {"steps": k_sus, "seconds": ms_sus * k_sus * 1e-3, "ms_per_step": ms_sus, "value": B_global / (ms_sus * 1e-3)}

    # per-kernel durations of the dominant kernel, CUDA events on the launching stream
    eng.set_timing(True)
    kms = []
    for i in range(max(3, min(a.steps, 10))):
        step_resident(step_no)
        step_no += 1
        kms.append(eng.last_kernel_ms())
    eng.set_timing(False)
    dec_ms = statistics.mean(k["dec_fused"] for k in kms)
    kshare = {k: statistics.mean(v[k] for v in kms) for k in kms[0]}
    loss_now = float(eng.scalars[0])
    used_tc = eng.used_tensor_cores()

    # the other arithmetic modes of the same step, short runs (reported beside the headline, never as it)
    modes = {}
    if not a.no_other_modes and not sub:
        for m in ("fp32", "tc_fp16x3", "tc_fp16"):
            if m == a.math:
                continue
            eng.set_math_mode(m)
            for i in range(3):
                step_resident(step_no)
                step_no += 1
            k = max(3, min(a.steps, 10))
            ms_m = timed(step_resident, k, step_no) / k
            step_no += k
            modes[m] = {"value": B_global / (ms_m * 1e-3), "ms_per_step": ms_m, "tensor_cores": eng.used_tensor_cores()}
        eng.set_math_mode(a.math)

    check = None
    if world > 1 and not sub and not a.no_rank_check:
        check = rank_check(a, wl, case_mod, vae_factory, world, rank, dev, n, w)

    if rank != 0:
        return None
    peaks = load_peaks()
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured bf16 sustained (MEASURED_PEAKS.json)" if peaks else "fallback 1.4 PFLOP/s sustained"
    ffma_peak = eng.ffma_peak_tflops()
    achieved = wl["flop_dec"] * rows / (dec_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "dec_traffic.json")))
        ent = tr.get(f"{a.workload}:{a.math}", {})
        traffic = ent.get("total")
        traffic_src = f"profiles/dec_traffic.json ({ent.get('source', 'ncu --set full capture')}, commit {ent.get('commit', tr.get('commit', '?'))})"
        if traffic is not None and ent.get("rows") and ent["rows"] != rows:
            traffic = traffic * rows / ent["rows"]
    except Exception:
        pass
    cfg = workload_config(a, wl, n_gpus)
    cfg["global_batch"] = B_global
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": a.scaling if not sub else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(rows * (vae.nd_x + vae.nd_c + vae.nd_y) * 4 + rows * 8), "d2h_bytes_per_step": 32,
                "gather": "per-step int64 row indices uploaded with the inputs (fused gather, dpivae.py:403-404); two host minibatches alternate"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor",
                     "kernel": ("dec_tc_kernel (fused decoders fwd+bwd, tcgen05 kind::f16, TMEM accumulators)" if used_tc
                                else "dec_kernel (fused decoders fwd+bwd, fp32 FFMA)"),
                     "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                     "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_flop_per_launch": wl["flop_dec"] * rows, "launch_ms": dec_ms,
                     "ffma_peak_tflops_measured": ffma_peak, "frac_of_ffma_peak": achieved / ffma_peak if ffma_peak else None,
                     "hbm_gbs_achieved": wl["bytes_row"] * rows / (ms_step * 1e-3) / 1e9,
                     "kernel_ms": kshare},
        "clocks": clk.summary(),
        "elbo": loss_now, "math": a.math, "tensor_cores": used_tc,
    }
    if modes:
        line["other_modes"] = modes
    if sustained:
        line["sustained"] = sustained
    if check is not None:
        line["rank_check"] = check
    del dp, eng, vae
    torch.cuda.empty_cache()
    return line


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bridge_p", choices=list(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--math", default="tc_fp16x3", choices=["fp32", "tc_fp16x3", "tc_fp16"],
                    help="decoder GEMM arithmetic (include/dpivae_b200.h DPIVAE_MATH_*): tc_fp16x3 = tcgen05 with the fp16 hi/lo "
                         "operand split (fp32-accurate, same 1e-5 parity bar as the FFMA kernel), fp32 = CUDA-core FFMA, "
                         "tc_fp16 = tcgen05 with plain fp16 operands (reduced precision, reported separately)")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the short runs of the other math modes")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --rows (131,072) rows per GPU; strong: the global batch is fixed at 1,048,576 rows (config 3) and "
                         "sharded over the ranks (N = 1 runs all of it)")
    ap.add_argument("--ref-rows", type=int, default=0, help="rows per timed step of the reference arm (default 65,536)")
    ap.add_argument("--no-workloads", action="store_true", help="skip the short beam_s / bridge_encode / ensemble sub-runs of the default line")
    ap.add_argument("--no-rank-check", action="store_true", help="skip the N-rank vs 1-rank correctness check (N > 1)")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="extra back-to-back run of the resident step (seconds) for the sustained-clock record")
    a = ap.parse_args()
    _guard_stdout()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    wl = dict(WORKLOADS[a.workload])
    if a.rows:
        wl["rows"] = a.rows
    if a.workload in ("bridge_encode", "ensemble") and a.impl == "reference":
        raise SystemExit("--impl reference covers the training-step workloads (bridge_p, beam_s)")
    if a.impl == "reference":
        return run_reference(a, wl)

    import torch
    import torch.distributed as dist

    ctx = _dist_setup()
    world, rank, local_rank, dev = ctx
    if a.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {a.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    if a.workload == "bridge_encode":
        line = run_encode(a, wl, ctx)
    elif a.workload == "ensemble":
        line = run_ensemble(a, wl, ctx)
    else:
        line = run_train(a, wl, ctx)
        if a.workload == "bridge_p" and not a.no_workloads and a.scaling == "weak":
            # BASELINE.json configs 2, 5 and 4 beside the headline (short runs, same process, same GPUs)
            subs = {}
            sa = argparse.Namespace(**vars(a))
            sa.steps, sa.warmup = max(5, min(a.steps, 10)), 3
            for name, fn in (("beam_s", run_train), ("bridge_encode", run_encode), ("ensemble", run_ensemble)):
                sa.workload = name
                try:
                    sl = fn(sa, dict(WORKLOADS[name]), ctx, True) if fn is run_train else fn(sa, dict(WORKLOADS[name]), ctx)
                except Exception as exc:   # a sub-run never takes the headline down
                    sl = {"error": str(exc)[:300]}
                if rank == 0 and sl is not None:
                    subs[name] = {k: sl[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "roofline", "e2e", "gpu_launches", "error")
                                  if k in sl}
            if rank == 0:
                line["workloads"] = subs
        if rank == 0 and not a.no_cpu_baseline and world == 1:
            n = wl["n_mc"]
            if reference_available():
                crow = 32768
                val, sec, threads, crow = reference_throughput(wl, 3, 1, crow)
                line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                                        "sample": f"{crow} rows x {n} MC, 3 iterations after 1 warm-up of the reference's own train_model "
                                                  "(unmodified code from baseline/_ref, fp32 torch CPU)", "ms_per_step": sec * 1e3}
            else:
                val, sec, threads, crow = cpu_oracle_throughput(wl, 3, 1)
                line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                        "sample": f"{crow} rows x {n} MC, 3 steps after 1 warm-up (oracle port, fp32 torch CPU)",
                                        "ms_per_step": sec * 1e3}
            # baseline leg, second device: the port's torch tensor algebra as eager CUDA ops on this GPU -- the "existing
            # Blackwell path" of SURVEY.md §8(d)
            try:
                torch.cuda.empty_cache()
                ev, esec, _, erow = cpu_oracle_throughput(wl, 3, 2, device=f"cuda:{local_rank}", rows=min(wl["rows"], 65536))
                line["cpu_baseline"]["torch_eager_same_gpu"] = {"value": ev, "unit": UNIT, "kind": "port (torch eager ops, fp32, same GPU)",
                                                                "sample": f"{erow} rows x {n} MC, 3 steps after 2 warm-ups", "ms_per_step": esec * 1e3}
            except Exception as exc:   # never let the informational leg break the bench line
                line["cpu_baseline"]["torch_eager_same_gpu"] = {"unavailable": str(exc)[:200]}
    if rank == 0 and line is not None:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
