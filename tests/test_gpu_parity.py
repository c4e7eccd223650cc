"""GPU parity: the CUDA path (through the C ABI, via the dpivae_b200 Python mirror) against
(a) the golden outputs of the unmodified reference and (b) the fp64 oracle, on the same inputs,
weights and injected noise.  Tolerance: 1e-5 relative (north_star, fp32)."""
import pytest
import torch

import golden_util as gu
from helpers import build_from_golden

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev_eps(g, spec):
    eps = gu.eps_of(g, spec)
    return tuple(e.cuda() for e in eps) if isinstance(eps, tuple) else eps.cuda()


@pytest.mark.parametrize("case,mtype", gu.CONFIGS)
def test_forward_latents_and_decoders(case, mtype):
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    n = g["eps0"].shape[0]
    with vae.inject_noise(_dev_eps(g, spec)):
        fw = vae.forward(x.cuda(), c.cuda(), cond=False, n=n)
    for name, t in zip(gu.FW_NAMES, fw):
        err = gu.rel_l2(t.cpu(), g[f"fw.{name}"])
        assert err < TOL, (name, err)


@pytest.mark.parametrize("case,mtype", gu.CONFIGS)
def test_loss_and_gradients(case, mtype):
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    n = g["eps0"].shape[0]
    B = x.shape[0]
    with vae.inject_noise(_dev_eps(g, spec)):
        loss8 = vae.loss(x.cuda(), c.cuda(), y.cuda(), n=n)
    for name, t in zip(gu.L8_NAMES, loss8):
        err = gu.rel_l2(t.detach().cpu(), g[f"loss8.{name}"])
        assert err < TOL, (name, err)
    elbo = loss8[0].sum() / (B * (vae.nd_x + vae.nd_c + vae.nd_y))
    elbo.backward()
    assert abs(float(elbo) - float(g["scalars"][0])) < TOL * max(1.0, abs(float(g["scalars"][0])))
    bad = {}
    for k, p in vae.named_parameters():
        if not p.requires_grad:
            continue
        err = gu.rel_l2(p.grad.cpu(), g[f"grad.{k}"])
        if err > 2e-5:  # the reference's own autograd noise floor on f_cov, see test_oracle_golden.py
            bad[k] = err
    assert not bad, bad


@pytest.mark.parametrize("case,mtype", gu.CONFIGS)
def test_against_fp64_oracle(case, mtype):
    """Error of the CUDA fp32 path measured against the fp64 oracle (tighter than fp32-vs-fp32)."""
    from oracle import dpivae_oracle as orc

    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    n = g["eps0"].shape[0]
    spec64 = orc.cast_spec(spec, torch.float64)
    sd64 = {k: v.double() for k, v in sd.items()}
    eps = gu.eps_of(g, spec)
    eps64 = tuple(e.double() for e in eps) if isinstance(eps, tuple) else eps.double()
    scal, l8, fw, grads = orc.loss_and_grads(sd64, spec64, x.double(), c.double(), y.double(), eps64)
    eng = vae.engine()
    row_loss, s = eng.loss(x, c, y, n, (1.0, 1.0, 1.0, 1.0), True, eps=_dev_eps(g, spec))
    assert gu.rel_l2(row_loss[0].cpu(), l8[0]) < TOL
    for k in range(8):
        assert abs(float(s[k]) - float(scal[k])) < TOL * max(1.0, abs(float(scal[k]))), (k, float(s[k]), float(scal[k]))
    bad = {}
    for p, o in eng.slots:
        name = [k for k, q in vae.named_parameters() if q is p][0]
        err = gu.rel_l2(eng.grads[o:o + p.numel()].cpu(), grads[name])
        if err > TOL:
            bad[name] = err
    assert not bad, bad


@pytest.mark.parametrize("case,mtype", [("bridge", "P"), ("damped_oscillator", "P"), ("simple_beam", "S")])
def test_adam_trajectory(case, mtype):
    """K fused train steps (gather + fwd + bwd + Adam) vs the reference's train_model run."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    from dpivae_b200 import param_groups

    K = int(g["traj.K"])
    n = g["traj.eps0"].shape[0]
    per = 3 if mtype == "P" else 1
    eng = vae.engine()
    eng.set_groups(param_groups(args))
    xd, cd, yd = x.cuda(), c.cuda(), y.cuda()
    for it in range(K):
        eps = gu.eps_of(g, spec, prefix="traj.eps", start=per * it)
        eps = tuple(e.cuda() for e in eps) if isinstance(eps, tuple) else eps.cuda()
        idx = torch.from_numpy(g["traj.idx"][it])
        # injected noise is indexed by minibatch position, like the reference's (n, B, nz) draw
        _, scal = eng.loss(xd, cd, yd, n, (1.0, 1.0, 1.0, 1.0), True, eps=eps, idx=idx, adam_step=it + 1)
        ref = float(g["traj.log.ELBO"][it])
        assert abs(float(scal[0]) - ref) < 1e-5 * max(1.0, abs(ref)), (it, float(scal[0]), ref)
    for k, p in vae.named_parameters():
        if p.requires_grad:
            err = gu.rel_l2(p.detach().cpu(), g[f"traj.final.{k}"])
            assert err < 1e-4, (k, err)


def test_philox_reproduces_torch_cuda_stream():
    """rng mode 1 must consume torch's CUDA generator exactly like the reference's rsample calls:
    same seed/offset -> bitwise the same latents as injecting torch.randn draws."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
    for n, rows in ((16, 24), (3, 17)):
        xd, cd = x[:rows].cuda(), c[:rows].cuda()
        torch.manual_seed(2024)
        eps = tuple(torch.randn(n, rows, k, device="cuda") for k in (vae.nz_x, vae.nz_c, vae.nz_y))
        after_torch = torch.cuda.default_generators[0].get_offset()
        with vae.inject_noise(eps):
            ref = vae.forward(xd, cd, n=n)
        torch.manual_seed(2024)
        out = vae.forward(xd, cd, n=n)
        assert torch.cuda.default_generators[0].get_offset() == after_torch
        for a, b in zip(out, ref):
            assert torch.equal(a, b)


def test_shard_invariance_and_determinism():
    """Data-parallel contract: two row shards of one global batch (global normaliser, noise indexed
    by global row) give per-row losses identical to, and gradients summing to, the 1-shard result."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
    n, B = 4, x.shape[0]
    eng = vae.engine()
    xd, cd, yd = x.cuda(), c.cuda(), y.cuda()
    w = (1.0, 1.0, 1.0, 1.0)
    eps = _dev_eps(g, spec)
    rl, s = eng.loss(xd, cd, yd, n, w, True, eps=eps)
    g_full = eng.grads.clone()
    rl2, s2 = eng.loss(xd, cd, yd, n, w, True, eps=eps)
    assert torch.equal(rl, rl2) and torch.equal(g_full, eng.grads)  # bitwise deterministic
    h = 10
    rl_a, s_a = eng.loss(xd[:h], cd[:h], yd[:h], n, w, True, eps=eps, B_global=B, row_offset=0)
    g_a = eng.grads.clone()
    rl_b, s_b = eng.loss(xd[h:], cd[h:], yd[h:], n, w, True, eps=eps, B_global=B, row_offset=h)
    g_b = eng.grads.clone()
    assert torch.equal(torch.cat([rl_a, rl_b], dim=1), rl)
    assert gu.rel_l2((g_a + g_b).cpu(), g_full.cpu()) < 1e-6
    assert gu.rel_l2((s_a + s_b).cpu(), s.cpu()) < 1e-6


@pytest.mark.parametrize("case,mtype", gu.CONFIGS)
def test_encode_only(case, mtype):
    """transform_inputs -> encode (models/vae.py:161-162, 125-151; dpivae_encode): latents and density of the
    encode-only inference path against the reference's forward outputs on the same noise."""
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    n = g["eps0"].shape[0]
    x_t = vae.transform_inputs(x.cuda())[0]
    with vae.inject_noise(_dev_eps(g, spec)):
        zx, zc, zy, dens = vae.encode(x_t, n=n)
    for name, t in (("zx", zx), ("zc", zc), ("zy", zy), ("dens_z", dens)):
        err = gu.rel_l2(t.cpu(), g[f"fw.{name}"])
        assert err < TOL, (name, err)


def test_sample_and_evaluate_model():
    """DPIVAE.sample 9-tuple (models/vae.py:233-255) and evaluate_model (dpivae.py:527-559) on the fused path."""
    import dpivae_b200 as dpv

    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
    out = vae.sample(x.cuda(), c.cuda(), cond=False, n=3)
    assert len(out) == 9
    B = x.shape[0]
    assert out[0].shape == (3, B, vae.nd_x) and out[4].shape == (3, B, vae.nd_y) and out[8].shape == (3, B)
    assert all(torch.isfinite(t).all() for t in out)
    args.n_mc_test = 8
    metrics, pred = dpv.evaluate_model(args, case_mod.definition, vae, (x.cuda(), c.cuda(), y))
    assert set(metrics[args.name]) == {"R2", "MSE", "MAE"} and pred[args.name].shape == (B, vae.nd_y)
    assert all(v.shape == (vae.nd_y,) for v in metrics[args.name].values())   # per-output raw values (utils/metrics.py:29-31)
