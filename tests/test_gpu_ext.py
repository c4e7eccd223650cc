"""GPU parity of the branches and shapes the first fixture set does not reach, through the C ABI, against the second
fixture set of the unmodified reference (tests/golden/make_golden_ext.py, n_mc = 8 so that the tensor-core kernels run
them too) and against the fp64 oracle at a BASELINE.json shape:

  * `train_model` with weight decay, gradient clipping, annealed loss weights and validation passes: all 13 training
    and 8 validation scalars the reference logs + final parameters, fp32 kernels AND the tc_fp16x3 tensor-core mode
    (the validation pass is the tensor-core forward-only loss path);
  * `forward(cond=True)`, the `lambda_x` regulariser, clamp saturation, the Uniform-prior -inf edge;
  * `DPIVAE.sample`: deterministic members equal the forward outputs, the three draws have the stated moments;
  * 8,192 rows x 16 MC (bridge P, simple_beam S) with IN-KERNEL Philox noise against the fp64 oracle fed the torch
    draws of the same generator state.
Tolerances: 1e-5 relative on losses / scalars / latents, 2e-5 relative L2 per gradient tensor, 1e-4 on parameters after
a trajectory -- the same bars for the fp32 kernels and for tc_fp16x3."""
import pytest
import torch

import golden_util as gu
from helpers import build_from_golden

pytestmark = pytest.mark.gpu
TOL = 1e-5
NAMES8 = ["ELBO", "KLx", "KLc", "KLy", "Rx", "Rc", "Ry", "reg"]


def _dev(e):
    return tuple(t.cuda() for t in e) if isinstance(e, tuple) else e.cuda()


@pytest.mark.parametrize("mode", ["fp32", "tc_fp16x3"])
@pytest.mark.parametrize("case,mtype", gu.EXT_CONFIGS)
def test_train_model_flagged_run_matches_reference(case, mtype, mode, monkeypatch):
    """dpivae.py:285-524 through the mirror's `train_model`: the reference's recorded minibatch indices and noise, its
    non-default flags (weight decay, clip_gradients, annealing, validation every 2 iterations)."""
    import dpivae_b200 as dpv

    g0 = gu.load(case, mtype, ext=True)[0]
    flags = gu.traj_flags(g0)
    K = int(g0["traj.K"])
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype, ext=True, n_val=16, n_mc_train=8, n_mc_val=8,
                                                                     n_iter=K, val_freq=2, math_mode=mode, **flags)
    per = 3 if mtype == "P" else 1
    draws = iter([torch.from_numpy(g["traj.idx"][it]) for it in range(K)])
    monkeypatch.setattr(torch, "multinomial", lambda *a, **k: next(draws))
    used_tc = []

    def provider(kind, it):
        used_tc.append(vae.engine().used_tensor_cores())
        if kind == "train":
            return _dev(gu.eps_of(g, spec, prefix="traj.eps", start=per * it))
        return _dev(gu.eps_of(g, spec, prefix="traj.val_eps", start=per * (it // 2)))

    args.eps_provider = provider
    xv, cv, yv = (torch.from_numpy(g[k]) for k in ("x_val", "c_val", "y_val"))
    vae2, logger = dpv.train_model(args, vae, case_mod.definition, (x, c, y), (xv, cv, yv))
    assert vae.engine().used_tensor_cores() == (mode != "fp32")
    sc = logger.experiment.scalars
    for nme in gu.TRAIN_LOG + gu.VAL_LOG:
        got = sc[nme]
        ref_v, ref_i = g[f"traj.log.{nme}"], g[f"traj.log_iter.{nme}"]
        assert [int(s) for s, _ in got] == [int(i) for i in ref_i], nme
        for (it, v), r in zip(got, ref_v):
            assert abs(float(v) - float(r)) < TOL * max(1.0, abs(float(r))), (nme, it, float(v), float(r))
    for k, p in vae.named_parameters():
        if p.requires_grad:
            err = gu.rel_l2(p.detach().cpu(), g[f"traj.final.{k}"])
            assert err < 1e-4, (k, err)


@pytest.mark.parametrize("case,mtype", gu.EXT_CONFIGS)
def test_forward_cond_true(case, mtype):
    """models/vae.py:165-167 (`forward(cond=True)`): zc comes from the conditional prior net, drawn right after the encoder's."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype, ext=True)
    per = 3 if mtype == "P" else 1
    eps = gu.eps_of(g, spec, prefix="cond.eps")
    eps = (tuple(eps) if isinstance(eps, tuple) else (eps,)) + (torch.from_numpy(g[f"cond.eps{per}"]),)
    if mtype == "S":
        eps = (eps[0], None, None, eps[1])
    for mode in ("fp32", "tc_fp16x3"):
        vae.engine().set_math_mode(mode)
        with vae.inject_noise(tuple(None if e is None else e.cuda() for e in eps)):
            fw = vae.forward(x.cuda(), c.cuda(), cond=True, n=8)
        for name, t in zip(gu.FW_NAMES, fw):
            err = gu.rel_l2(t.cpu(), g[f"cond.fw.{name}"])
            assert err < TOL, (mode, name, err)


def _loss_grads_vs_golden(g, spec, vae, x, c, y, section, grad_tol, TOL=TOL, grads_ref=None):
    """grads_ref: compare the gradients with these (fp64 oracle) instead of the reference's fp32 autograd values."""
    eng = vae.engine()
    eps = _dev(gu.eps_of(g, spec, prefix=f"{section}.eps"))
    row_loss, scal = eng.loss(x, c, y, 8, (1.0, 1.0, 1.0, 1.0), True, eps=eps)
    names6 = ["loss", "KLx", "Rx", "Rc", "Ry", "reg"]
    for i, nme in enumerate(names6):
        err = gu.rel_l2(row_loss[i].cpu(), g[f"{section}.loss8.{nme}"])
        assert err < TOL, (section, nme, err)
    for k in range(8):
        ref = float(g[f"{section}.scalars"][k])
        assert abs(float(scal[k]) - ref) < TOL * max(1.0, abs(ref)), (section, k, float(scal[k]), ref)
    names = {id(p): k for k, p in vae.named_parameters()}
    bad = {}
    for p, o in eng.slots:
        ref = g[f"{section}.grad.{names[id(p)]}"] if grads_ref is None else grads_ref[names[id(p)]]
        err = gu.rel_l2(eng.grads[o:o + p.numel()].cpu(), ref)
        if err > grad_tol:
            bad[names[id(p)]] = err
    assert not bad, (section, sorted(bad.items(), key=lambda kv: -kv[1])[:6])
    return eng


@pytest.mark.parametrize("case,mtype", gu.EXT_CONFIGS)
def test_lambda_x_regulariser(case, mtype):
    """models/vae.py:217-219: reg = mean_n sum_d log N(xh_d; 0, lambda_x) enters the loss and every decoder_x gradient."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype, ext=True, lambda_x=0.7)
    eng = _loss_grads_vs_golden(g, spec, vae, x, c, y, "lamx", 2e-5)
    assert float(eng.scalars[7]) != 0.0


@pytest.mark.parametrize("mode", ["fp32", "tc_fp16x3"])
@pytest.mark.parametrize("case,mtype", gu.EXT_CONFIGS)
def test_clamp_saturation(case, mtype, mode):
    """models/encoders.py:35-39,123-124: heads on / beyond +-50, [-7, 3], +-20 -- values clamp, gradients through a
    saturated clamp are exactly zero.  Per-row losses and scalars against the reference's golden values; gradients
    against the fp64 oracle (the reference's own fp32 gradients carry ~5e-4 of round-off noise in this regime, see
    tests/test_oracle_golden_ext.py), 2e-5 per tensor in both modes."""
    from oracle import dpivae_oracle as orc

    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype, ext=True)
    sat = gu.state_of(g, spec, "sat.init")
    vae.load_state_dict(sat, strict=False)
    vae.engine().set_math_mode(mode)
    eps64 = gu.eps_of(g, spec, prefix="sat.eps")
    eps64 = tuple(e.double() for e in eps64) if isinstance(eps64, tuple) else eps64.double()
    _, _, _, o_grads = orc.loss_and_grads({k: v.double() for k, v in sat.items()}, orc.cast_spec(spec, torch.float64),
                                          x.double(), c.double(), y.double(), eps64)
    eng = _loss_grads_vs_golden(g, spec, vae, x, c, y, "sat", 2e-5, grads_ref=o_grads)
    assert eng.used_tensor_cores() == (mode != "fp32")
    names = {k: p for k, p in vae.named_parameters()}
    off = {id(p): o for p, o in eng.slots}
    pre = "encoder_y" if mtype == "P" else "encoder"
    nz = names[f"{pre}.net.f_mean.bias"].numel()
    assert float(eng.grads[off[id(names[f"{pre}.net.f_mean.bias"])] + nz - 1]) == 0.0
    assert float(eng.grads[off[id(names[f"{pre}.net.f_sigma.bias"])] + nz - 1]) == 0.0
    assert float(eng.grads[off[id(names["prior_net_c.net.f_mean.bias"])]]) == 0.0


@pytest.mark.parametrize("mode", ["fp32", "tc_fp16x3"])
@pytest.mark.parametrize("case,mtype", [cfg for cfg in gu.EXT_CONFIGS if cfg[0] != "simple_beam"])
def test_uniform_prior_edge_gives_inf(case, mtype, mode):
    """sigmoid -> 1.0f puts zx exactly on `high` of its half-open Uniform prior: log p = -inf, KL = +inf on the same rows
    as the reference, the reconstruction terms stay finite and equal."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype, ext=True)
    vae.load_state_dict(gu.state_of(g, spec, "edge.init"), strict=False)
    eng = vae.engine()
    eng.set_math_mode(mode)
    row_loss, _ = eng.loss(x, c, y, 8, (1.0, 1.0, 1.0, 1.0), False, eps=_dev(gu.eps_of(g, spec, prefix="edge.eps")))
    ref_kl = torch.from_numpy(g["edge.loss8.KLx"])
    kl = row_loss[1].cpu()
    assert torch.isinf(ref_kl).any()
    assert torch.equal(torch.isinf(kl) & (kl > 0), torch.isinf(ref_kl) & (ref_kl > 0))
    for i, nme in ((2, "Rx"), (3, "Rc"), (4, "Ry")):
        assert gu.rel_l2(row_loss[i].cpu(), g[f"edge.loss8.{nme}"]) < TOL, nme


def test_sample_members_and_moments():
    """models/vae.py:233-255: the 9-tuple's deterministic members are the forward outputs on the same noise; the three
    Normal draws are centred on the decoder means with the decoder scales."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P", ext=True)
    eps = _dev(gu.eps_of(g, spec, prefix="lamx.eps"))
    n = 8
    from oracle import dpivae_oracle as orc

    fw = orc.forward({k: v.double() for k, v in sd.items()}, orc.cast_spec(spec, torch.float64), x.double(), c.double(),
                     tuple(e.double().cpu() for e in eps))
    torch.manual_seed(3)
    with vae.inject_noise(eps):
        out = vae.sample(x.cuda(), c.cuda(), cond=False, n=n)
    x_s, xh_p, xh_d, c_s, y_s, zx, zc, zy, dens = out
    for t, ref in ((xh_p, fw[0]), (xh_d, fw[1]), (zx, fw[6]), (zc, fw[7]), (zy, fw[8]), (dens, fw[9])):
        assert gu.rel_l2(t.cpu(), ref) < TOL
    sx = float(vae.log_sigma_x.detach().exp())
    rx = (x_s - (xh_p + xh_d)).double().cpu() / sx
    m = rx.numel()
    assert abs(float(rx.mean())) < 5.0 / m ** 0.5 and abs(float(rx.std()) - 1.0) < 5.0 / (2 * m) ** 0.5
    for s, mean, ls in ((c_s, fw[2], fw[3]), (y_s, fw[4], fw[5])):
        r = ((s.double().cpu() - mean) / ls.exp())
        k = r.numel()
        assert abs(float(r.mean())) < 5.0 / k ** 0.5 and abs(float(r.std()) - 1.0) < 5.0 / (2 * k) ** 0.5


@pytest.mark.parametrize("case,mtype,mode", [("bridge", "P", "tc_fp16x3"), ("simple_beam", "S", "tc_fp16x3"), ("bridge", "P", "fp32")])
def test_baseline_shape_inkernel_philox_vs_fp64_oracle(case, mtype, mode):
    """8,192 rows x 16 MC samples (the BASELINE.json shapes' per-row work, 1,024 tiles of 128 pairs): the kernels draw
    their own Philox noise; the fp64 oracle gets the `torch.randn`-equivalent CUDA draws of the same generator state
    (bitwise the same stream: tests/test_gpu_parity.py::test_philox_reproduces_torch_cuda_stream)."""
    from oracle import dpivae_oracle as orc

    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype, ext=True)
    reps = 8192 // x.shape[0] + 1
    X, C_, Y = (torch.cat([t] * reps)[:8192].contiguous() for t in (x, c, y))
    # de-duplicate the rows: a row-dependent perturbation of the size of the data noise
    gen = torch.Generator().manual_seed(5)
    X = X + 1e-3 * X.std(0, keepdim=True) * torch.randn(X.shape, generator=gen)
    n, B = 16, 8192
    eng = vae.engine()
    eng.set_math_mode(mode)
    torch.manual_seed(2024)
    widths = (vae.nz_x, vae.nz_c, vae.nz_y) if mtype == "P" else (vae.nz_x + vae.nz_c + vae.nz_y,)
    draws = [torch.empty((n, B, k), device="cuda").normal_() for k in widths]
    torch.manual_seed(2024)
    row_loss, scal = eng.loss(X, C_, Y, n, (1.0, 1.0, 1.0, 1.0), True)
    assert eng.used_tensor_cores() == (mode != "fp32")
    eps64 = tuple(d.double().cpu() for d in draws) if mtype == "P" else draws[0].double().cpu()
    o_scal, o_l8, _, o_grads = orc.loss_and_grads({k: v.double() for k, v in sd.items()}, orc.cast_spec(spec, torch.float64),
                                                  X.double(), C_.double(), Y.double(), eps64)
    assert gu.rel_l2(row_loss[0].cpu(), o_l8[0]) < TOL
    assert gu.rel_l2(row_loss[1].cpu(), o_l8[1]) < TOL
    for k in range(8):
        assert abs(float(scal[k]) - float(o_scal[k])) < TOL * max(1.0, abs(float(o_scal[k]))), (k, float(scal[k]), float(o_scal[k]))
    names = {id(p): k for k, p in vae.named_parameters()}
    bad = {}
    for p, o in eng.slots:
        err = gu.rel_l2(eng.grads[o:o + p.numel()].cpu(), o_grads[names[id(p)]])
        if err > 2e-5:
            bad[names[id(p)]] = err
    assert not bad, bad
