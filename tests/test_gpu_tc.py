"""GPU parity of the tensor-core (tcgen05) decoder kernel, through the C ABI.

Two modes (include/dpivae_b200.h DPIVAE_MATH_*), each against the fp64 oracle on the golden inputs,
weights and injected noise:
  tc_fp16x3  fp16 hi/lo operand split, three MMAs per GEMM, fp32 TMEM accumulators.  Stated tolerance:
             1e-5 relative on the per-row loss and the 8 scalars, 2e-5 per-tensor relative L2 on gradients
             (the same bars the fp32 kernel is held to against the reference's fp32 golden outputs).
  tc_fp16    plain fp16 operands.  Stated tolerance: 2e-3 on loss / scalars, 2e-2 on gradients (1e-1 on the
             f_cov tensors, whose gradient is the round-off-dominated residue of an analytically-zero term).
"""
import pytest
import torch

import golden_util as gu
from helpers import build_from_golden

pytestmark = pytest.mark.gpu
TOL = {"tc_fp16x3": (1e-5, 2e-5), "tc_fp16": (2e-3, 2e-2)}


def _dev_eps(g, spec):
    eps = gu.eps_of(g, spec)
    return tuple(e.cuda() for e in eps) if isinstance(eps, tuple) else eps.cuda()


def _tile(t, reps):
    return torch.cat([t] * reps, dim=0)


def _oracle(g, spec, sd, x, c, y, eps, reps, n_rep):
    from oracle import dpivae_oracle as orc

    spec64 = orc.cast_spec(spec, torch.float64)
    sd64 = {k: v.double() for k, v in sd.items()}
    if isinstance(eps, tuple):
        eps64 = tuple(torch.cat([torch.cat([e.double().cpu()] * reps, dim=1)] * n_rep, dim=0) for e in eps)
    else:
        eps64 = torch.cat([torch.cat([eps.double().cpu()] * reps, dim=1)] * n_rep, dim=0)
    return orc.loss_and_grads(sd64, spec64, _tile(x, reps).double(), _tile(c, reps).double(), _tile(y, reps).double(), eps64), eps64


@pytest.mark.parametrize("mode", ["tc_fp16x3", "tc_fp16"])
@pytest.mark.parametrize("case,mtype", gu.CONFIGS)
def test_tc_loss_and_gradients_vs_fp64_oracle(case, mtype, mode):
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    n0 = g["eps0"].shape[0]
    # the tensor-core kernel needs 8 <= n_mc <= 128: repeat the golden MC draws along the MC axis, and the
    # rows so that several 128-pair tiles (and a ragged last tile) are exercised
    n_rep = max(1, -(-8 // n0))
    n = n0 * n_rep
    reps = 3
    eps = gu.eps_of(g, spec)
    (scal, l8, fw, grads), eps64 = _oracle(g, spec, sd, x, c, y, eps, reps, n_rep)
    eng = vae.engine()
    eng.set_math_mode(mode)
    dev_eps = tuple(e.float().cuda() for e in eps64) if isinstance(eps64, tuple) else eps64.float().cuda()
    row_loss, s = eng.loss(_tile(x, reps), _tile(c, reps), _tile(y, reps), n, (1.0, 1.0, 1.0, 1.0), True, eps=dev_eps)
    assert eng.used_tensor_cores(), "tensor-core kernel was not selected"
    tol_l, tol_g = TOL[mode]
    assert gu.rel_l2(row_loss[0].cpu(), l8[0]) < tol_l
    for k in range(8):
        assert abs(float(s[k]) - float(scal[k])) < tol_l * max(1.0, abs(float(scal[k]))), (k, float(s[k]), float(scal[k]))
    bad = {}
    for p, o in eng.slots:
        name = [k for k, q in vae.named_parameters() if q is p][0]
        err = gu.rel_l2(eng.grads[o:o + p.numel()].cpu(), grads[name])
        if err > (1e-1 if mode == "tc_fp16" and ".f_cov." in name else tol_g):
            bad[name] = err
    assert not bad, bad


def test_tc_matches_fp32_kernel_with_philox_noise():
    """Same in-kernel Philox stream in both kernels: per-row losses of the tensor-core path track the fp32
    kernel on a multi-tile batch with a ragged tail."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
    eng = vae.engine()
    reps = 37  # 37 * 24 rows = 888 rows x 16 MC = 111 tiles, last one ragged
    X, C_, Y = _tile(x, reps), _tile(c, reps), _tile(y, reps)
    torch.manual_seed(5)
    rl32, s32 = eng.loss(X, C_, Y, 16, (1.0, 1.0, 1.0, 1.0), True)
    g32 = eng.grads.clone()
    assert not eng.used_tensor_cores()
    eng.set_math_mode("tc_fp16x3")
    torch.manual_seed(5)
    rl, s = eng.loss(X, C_, Y, 16, (1.0, 1.0, 1.0, 1.0), True)
    assert eng.used_tensor_cores()
    assert gu.rel_l2(rl.cpu(), rl32.cpu()) < 1e-5
    assert gu.rel_l2(s.cpu(), s32.cpu()) < 1e-5
    assert gu.rel_l2(eng.grads.cpu(), g32.cpu()) < 2e-5
    # deterministic: same call, same bits
    torch.manual_seed(5)
    rl2, _ = eng.loss(X, C_, Y, 16, (1.0, 1.0, 1.0, 1.0), True)
    assert torch.equal(rl, rl2)


@pytest.mark.parametrize("case,mtype", [("bridge", "P"), ("bridge", "S"), ("simple_beam", "S"), ("damped_oscillator", "P")])
def test_tc_shard_invariance_with_inkernel_philox(case, mtype):
    """Thread-per-pair latent kernels (n_mc = 16), in-kernel Philox: an unsharded call draws its noise with the
    pre-pass (one Philox evaluation per four elements, torch's mapping), row shards draw it per element inside the
    forward; both must be the SAME stream -- per-row losses of the shards bitwise equal to the unsharded call's,
    gradients and scalars summing to it."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    eng = vae.engine()
    eng.set_math_mode("tc_fp16x3")
    reps = 171   # 4104 rows x 16 MC = 513 tiles: above the size where calls use the noise pre-pass; the shards end in ragged tiles
    X, C_, Y = _tile(x, reps).cuda(), _tile(c, reps).cuda(), _tile(y, reps).cuda()
    B = X.shape[0]
    w = (1.0, 1.0, 1.0, 1.0)
    torch.manual_seed(11)
    rl, s = eng.loss(X, C_, Y, 16, w, True)
    assert eng.used_tensor_cores()
    g_full = eng.grads.clone()
    cuts = [0, 1367, 3001, B]
    rls, gsum, ssum = [], torch.zeros_like(g_full), torch.zeros_like(s)
    for a, b in zip(cuts[:-1], cuts[1:]):
        torch.manual_seed(11)
        rl_k, s_k = eng.loss(X[a:b], C_[a:b], Y[a:b], 16, w, True, B_global=B, row_offset=a)
        rls.append(rl_k)
        gsum += eng.grads
        ssum += s_k
    assert torch.equal(torch.cat(rls, dim=1), rl)
    assert gu.rel_l2(gsum.cpu(), g_full.cpu()) < 2e-6
    assert gu.rel_l2(ssum.cpu(), s.cpu()) < 2e-6
    # CYCLIC shards (rank k of S owns rows k, k + S, ...): whole Philox evaluations stay on one rank, the noise comes from
    # the cyclic pre-pass when S * nz divides the generator grid (it does not at this size for every block: both the
    # pre-pass and the per-element fallback are exercised across the parametrisation)
    for S in (2, 4):
        gsum.zero_()
        ssum.zero_()
        for k in range(S):
            torch.manual_seed(11)
            rl_k, s_k = eng.loss(X[k::S].contiguous(), C_[k::S].contiguous(), Y[k::S].contiguous(), 16, w, True,
                                 B_global=B, row_offset=k, row_stride=S)
            assert torch.equal(rl_k, rl[:, k::S])
            gsum += eng.grads
            ssum += s_k
        assert gu.rel_l2(gsum.cpu(), g_full.cpu()) < 2e-6
        assert gu.rel_l2(ssum.cpu(), s.cpu()) < 2e-6


@pytest.mark.parametrize("case,mtype", [("bridge", "P"), ("simple_beam", "S")])
def test_tc_latent_outputs_match_fp32_kernel(case, mtype):
    """Latent outputs of a loss call (zx / zc / zy / dens_z, models/vae.py:161-176) on the tensor-core path -- written by the
    thread-per-pair latent kernel -- against the fp32 kernels on the same in-kernel Philox stream."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    eng = vae.engine()
    reps, n = 7, 16
    X, C_, Y = _tile(x, reps).cuda(), _tile(c, reps).cuda(), _tile(y, reps).cuda()
    B = X.shape[0]
    res = {}
    for mode in ("fp32", "tc_fp16x3"):
        eng.set_math_mode(mode)
        out = {"zx": torch.zeros(n, B, vae.nz_x, device="cuda"), "zc": torch.zeros(n, B, vae.nz_c, device="cuda"),
               "zy": torch.zeros(n, B, vae.nz_y, device="cuda"), "dens_z": torch.zeros(n, B, device="cuda")}
        torch.manual_seed(3)
        eng.loss(X, C_, Y, n, (1.0, 1.0, 1.0, 1.0), False, outputs=out)
        assert eng.used_tensor_cores() == (mode != "fp32")
        res[mode] = out
    for k in ("zx", "zc", "zy", "dens_z"):
        assert gu.rel_l2(res["tc_fp16x3"][k].cpu(), res["fp32"][k].cpu()) < 1e-5, k


# The K-step Adam trajectory of the tensor-core mode against the REFERENCE's own `train_model` run lives in
# tests/test_gpu_ext.py::test_train_model_flagged_run_matches_reference[tc_fp16x3-*] (fixtures with n_mc = 8: the
# first fixture set has n_mc = 4, below the tensor-core kernel's 8 <= n_mc <= 128 window).


def test_tc_training_tracks_fp32_kernel():
    """Ten fused train steps (in-kernel Philox noise, n_mc = 16, several tiles) in tc_fp16x3 mode end at the same
    parameters as the fp32 FFMA kernels: 1e-5 on the logged ELBO of every step, 1e-4 relative L2 on every tensor."""
    from dpivae_b200 import param_groups

    finals, elbos = {}, {}
    for mode in ("fp32", "tc_fp16x3"):
        g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
        eng = vae.engine()
        eng.set_groups(param_groups(args))
        eng.set_math_mode(mode)
        X, C_, Y = _tile(x, 11).cuda(), _tile(c, 11).cuda(), _tile(y, 11).cuda()
        torch.manual_seed(77)
        e = []
        for it in range(10):
            _, scal = eng.loss(X, C_, Y, 16, (1.0, 1.0, 1.0, 1.0), True, adam_step=it + 1)
            e.append(float(scal[0]))
        assert eng.used_tensor_cores() == (mode != "fp32")
        finals[mode] = {k: p.detach().clone() for k, p in vae.named_parameters() if p.requires_grad}
        elbos[mode] = e
    for a, b in zip(elbos["fp32"], elbos["tc_fp16x3"]):
        assert abs(a - b) < 1e-5 * max(1.0, abs(a)), (a, b)
    for k in finals["fp32"]:
        err = gu.rel_l2(finals["tc_fp16x3"][k].cpu(), finals["fp32"][k].cpu())
        assert err < 1e-4, (k, err)


@pytest.mark.parametrize("case,mtype", gu.CONFIGS)
def test_tc_encoder_forward_and_encode(case, mtype):
    """Tensor-core ENCODER kernel (enc_tc_fwd_kernel, used by forward / validation / encode-only calls in the tensor-core
    modes): all 10 forward outputs and the encode-only latents against the reference's golden forward, 1e-5."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    n = g["eps0"].shape[0]
    vae.engine().set_math_mode("tc_fp16x3")
    with vae.inject_noise(_dev_eps(g, spec)):
        fw = vae.forward(x.cuda(), c.cuda(), cond=False, n=n)
        x_t = vae.transform_inputs(x.cuda())[0]
        zx, zc, zy, dens = vae.encode(x_t, n=n)
    for name, t in zip(gu.FW_NAMES, fw):
        err = gu.rel_l2(t.cpu(), g[f"fw.{name}"])
        assert err < 1e-5, (name, err)
    for name, t in (("zx", zx), ("zc", zc), ("zy", zy), ("dens_z", dens)):
        err = gu.rel_l2(t.cpu(), g[f"fw.{name}"])
        assert err < 1e-5, (name, err)


@pytest.mark.parametrize("case,mtype", [("bridge", "P"), ("simple_beam", "S"), ("damped_oscillator", "S")])
@pytest.mark.parametrize("B,n", [(1, 16), (5, 8), (37, 16), (131, 24), (9, 128), (3, 40)])
def test_tc_ragged_shapes_match_fp32_kernel(case, mtype, B, n):
    """Edge shapes of the tile walk: a single row, batches that do not fill the last 128-pair tile, MC counts that are
    not powers of two (rows per tile = 128 // n with idle pair slots), n = 128 (one row per tile), a gathered minibatch.
    The generic and the shape-specialised latent kernels, the tensor-core decoder and the encoder kernels must agree
    with the fp32 FFMA path on the same in-kernel Philox stream (loss rows, scalars, every gradient)."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    eng = vae.engine()
    rows = torch.arange(B) % x.shape[0]
    X, C_, Y = x[rows].contiguous(), c[rows].contiguous(), y[rows].contiguous()
    idx = torch.randperm(B, generator=torch.Generator().manual_seed(B + n))
    out = {}
    for mode in ("fp32", "tc_fp16x3"):
        eng.set_math_mode(mode)
        torch.manual_seed(17)
        rl, s = eng.loss(X, C_, Y, n, (0.7, 1.0, 0.9, 1.1), True, idx=idx)
        assert eng.used_tensor_cores() == (mode != "fp32")
        out[mode] = (rl.cpu().clone(), s.cpu().clone(), eng.grads.cpu().clone())
    assert torch.isfinite(out["fp32"][0]).all() and torch.isfinite(out["fp32"][2]).all()
    assert gu.rel_l2(out["tc_fp16x3"][0], out["fp32"][0]) < 1e-5
    assert gu.rel_l2(out["tc_fp16x3"][1], out["fp32"][1]) < 1e-5
    assert gu.rel_l2(out["tc_fp16x3"][2], out["fp32"][2]) < 3e-5


def test_tc_large_batch_is_deterministic_and_matches_fp32():
    """Many tiles per persistent CTA (double-buffered bulk-copy prefetch in lat_bwd / dec_tc / enc_tc_bwd, register
    prefetch in the encoder and prior kernels): the same call twice gives the same bits for every gradient, and the
    result agrees with the fp32 FFMA path -- a race in the buffer hand-over would show up as either."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
    eng = vae.engine()
    reps = 1500   # 36,000 rows x 16 MC = 4,500 tiles -> ~30 tiles per CTA
    X, C_, Y = _tile(x, reps).cuda(), _tile(c, reps).cuda(), _tile(y, reps).cuda()
    X = X + 2e-5 * torch.randn(X.shape, generator=torch.Generator().manual_seed(1)).cuda()   # ~0.2 sigma of the fitted scaler
    w = (1.0, 1.0, 1.0, 1.0)
    eng.set_math_mode("tc_fp16x3")
    runs = []
    for _ in range(3):
        torch.manual_seed(123)
        rl, s = eng.loss(X, C_, Y, 16, w, True)
        runs.append((rl.clone(), s.clone(), eng.grads.clone()))
    assert eng.used_tensor_cores()
    assert torch.isfinite(runs[0][2]).all() and torch.isfinite(runs[0][0]).all()
    for r in runs[1:]:
        assert torch.equal(r[0], runs[0][0]) and torch.equal(r[1], runs[0][1]) and torch.equal(r[2], runs[0][2])
    eng.set_math_mode("fp32")
    torch.manual_seed(123)
    rl32, s32 = eng.loss(X, C_, Y, 16, w, True)
    assert gu.rel_l2(runs[0][0].cpu(), rl32.cpu()) < 1e-5
    assert gu.rel_l2(runs[0][1].cpu(), s32.cpu()) < 1e-5
    assert gu.rel_l2(runs[0][2].cpu(), eng.grads.cpu()) < 3e-5


def test_tc_outlier_inputs_give_finite_gradients():
    """Rows tens of scaler sigmas away from the training distribution (bad sensors, unit mix-ups): the fp16 hi/lo
    operands keep their range (|standardised input| up to 3750 sigma before they saturate), gradients stay finite and
    the tensor-core path still tracks the fp32 path."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
    eng = vae.engine()
    X = x.clone()
    X[::3] += 0.003    # ~30 sigma of the fitted scaler
    out = {}
    for mode in ("fp32", "tc_fp16x3"):
        eng.set_math_mode(mode)
        torch.manual_seed(3)
        rl, s = eng.loss(X, c, y, 16, (1.0, 1.0, 1.0, 1.0), True)
        assert torch.isfinite(eng.grads).all(), mode
        out[mode] = (rl.cpu().clone(), eng.grads.cpu().clone())
    fin = torch.isfinite(out["fp32"][0][0])
    assert gu.rel_l2(out["tc_fp16x3"][0][0][fin], out["fp32"][0][0][fin]) < 1e-5
    assert gu.rel_l2(out["tc_fp16x3"][1], out["fp32"][1]) < 5e-5


@pytest.mark.gpu
def test_tc_extreme_outliers_do_not_poison_the_gradients():
    """Rows ~100 scaler sigmas out (row losses overflow to +-inf in the fp32 kernels as well): the head gradients of
    those rows exceed the fp16 operand range of the encoder backward; they saturate instead of turning the whole
    weight gradient into NaN (the fp32 kernels give finite gradients on the same data)."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
    eng = vae.engine()
    reps = 40
    X, C_, Y = (torch.cat([t] * reps, 0).cuda() for t in (x, c, y))
    X = X + 0.01 * torch.randn(X.shape, generator=torch.Generator().manual_seed(1)).cuda()
    for mode in ("fp32", "tc_fp16x3"):
        eng.set_math_mode(mode)
        torch.manual_seed(123)
        eng.loss(X, C_, Y, 16, (1.0, 1.0, 1.0, 1.0), True)
        assert not torch.isnan(eng.grads).any(), mode


@pytest.mark.parametrize("case", ["bridge", "damped_oscillator", "simple_beam"])
@pytest.mark.parametrize("B,n", [(1, 1), (127, 1), (128, 3), (129, 1), (1000, 2), (40000, 1)])
def test_fused_encode_kernel_matches_two_kernel_path(case, B, n, monkeypatch):
    """Encode-only calls of the P presets run ONE warp-specialised kernel (enc_fused_kernel: encoder MMAs + latent sampling,
    tiles pipelined through TMEM; its noise comes from noise_fill_kernel, one Philox evaluation per four elements, or --
    DPIVAE_NO_NOISE_PREPASS -- from the per-element generator).  Same arithmetic as enc_tc_fwd_kernel -> headpre ->
    lat_encode_kernel on the same Philox stream, so latents and density must be BITWISE equal: single row, ragged last tile,
    several MC samples, more tiles than CTAs (40,000 rows = 313 tiles on 148 SMs)."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, "P")
    eng = vae.engine()
    eng.set_math_mode("tc_fp16x3")
    rows = torch.arange(B) % x.shape[0]
    gen = torch.Generator().manual_seed(B)
    X = (x[rows] + 0.05 * torch.randn(B, x.shape[1], generator=gen)).contiguous().cuda()
    out = {}
    for mode in ("fused", "fused_inline_noise", "two_kernels"):
        monkeypatch.delenv("DPIVAE_NO_FUSED_ENCODE", raising=False)
        monkeypatch.delenv("DPIVAE_NO_NOISE_PREPASS", raising=False)
        if mode == "two_kernels":
            monkeypatch.setenv("DPIVAE_NO_FUSED_ENCODE", "1")
        if mode == "fused_inline_noise":
            monkeypatch.setenv("DPIVAE_NO_NOISE_PREPASS", "1")
        torch.manual_seed(5)
        l0 = eng.launches
        out[mode] = [t.cpu().clone() for t in eng.encode(X, n, False)]
        assert eng.launches - l0 == (1 if mode == "fused_inline_noise" else 2)
    for mode in ("fused", "fused_inline_noise"):
        for a, b, name in zip(out[mode], out["two_kernels"], ("zx", "zc", "zy", "dens_z")):
            assert a.shape == b.shape and torch.isfinite(a).all(), (mode, name)
            assert torch.equal(a, b), (mode, name, (a - b).abs().max())
