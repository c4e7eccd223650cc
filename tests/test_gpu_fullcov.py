"""GPU parity of `--full_cov_prior True` (dpivae.py:151-153: FullCovarianceNN conditional prior nets; models/vae.py:202-203:
MultivariateNormal log-density with a full lower-triangular factor) through the C ABI, against fixtures of the unmodified
reference (tests/golden/make_golden_fullcov.py): loss 8-tuple, scalars, every gradient (incl. the priors' f_cov heads),
`forward(cond=True)`, `prior_net`, and a `train_model` trajectory -- fp32 kernels and the tc_fp16x3 mode (whose decoder side
runs the fp32 kernel for this flag, the encoders stay on the tensor cores).  Tolerances as everywhere: 1e-5 / 2e-5 / 1e-4."""
import pytest
import torch

import golden_util as gu
from helpers import build_from_golden

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev(e):
    return tuple(t.cuda() for t in e) if isinstance(e, tuple) else e.cuda()


@pytest.mark.parametrize("mode", ["fp32", "tc_fp16x3"])
@pytest.mark.parametrize("case,mtype", gu.FULLCOV_CONFIGS)
def test_loss_and_gradients(case, mtype, mode):
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype, ext="fullcov", full_cov_prior=True)
    eng = vae.engine()
    eng.set_math_mode(mode)
    eps = _dev(gu.eps_of(g, spec, prefix="loss.eps"))
    row_loss, scal = eng.loss(x, c, y, 8, (1.0, 1.0, 1.0, 1.0), True, eps=eps)
    for i, nme in enumerate(["loss", "KLx", "Rx", "Rc", "Ry", "reg"]):
        assert gu.rel_l2(row_loss[i].cpu(), g[f"loss.loss8.{nme}"]) < TOL, nme
    for k in range(8):
        ref = float(g["loss.scalars"][k])
        assert abs(float(scal[k]) - ref) < TOL * max(1.0, abs(ref)), (k, float(scal[k]), ref)
    names = {id(p): k for k, p in vae.named_parameters()}
    bad = {}
    for p, o in eng.slots:
        err = gu.rel_l2(eng.grads[o:o + p.numel()].cpu(), g[f"loss.grad.{names[id(p)]}"])
        if err > 2e-5:
            bad[names[id(p)]] = err
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:6]
    assert "prior_net_c.net.f_cov.weight" in names.values()


@pytest.mark.parametrize("case,mtype", gu.FULLCOV_CONFIGS)
def test_forward_cond_and_prior_net(case, mtype):
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype, ext="fullcov", full_cov_prior=True)
    per = 3 if mtype == "P" else 1
    eps = gu.eps_of(g, spec, prefix="cond.eps")
    eps = (tuple(eps) if isinstance(eps, tuple) else (eps,)) + (torch.from_numpy(g[f"cond.eps{per}"]),)
    if mtype == "S":
        eps = (eps[0], None, None, eps[1])
    with vae.inject_noise(tuple(None if e is None else e.cuda() for e in eps)):
        fw = vae.forward(x.cuda(), c.cuda(), cond=True, n=8)
    for name, t in zip(gu.FW_NAMES, fw):
        assert gu.rel_l2(t.cpu(), g[f"cond.fw.{name}"]) < TOL, name
    for nme, t in zip(("loc_c", "tril_c", "loc_y", "tril_y"), vae.prior_net(c.cuda(), y.cuda())):
        assert gu.rel_l2(t.cpu(), g[f"pnet.{nme}"]) < TOL, nme


@pytest.mark.parametrize("mode", ["fp32", "tc_fp16x3"])
@pytest.mark.parametrize("case,mtype", gu.FULLCOV_CONFIGS)
def test_train_model_matches_reference(case, mtype, mode, monkeypatch):
    import dpivae_b200 as dpv

    g0 = gu.load(case, mtype, ext="fullcov")[0]
    K = int(g0["traj.K"])
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype, ext="fullcov", full_cov_prior=True, n_val=24,
                                                                     n_mc_train=8, n_mc_val=8, n_iter=K, val_freq=1000, math_mode=mode)
    per = 3 if mtype == "P" else 1
    draws = iter([torch.from_numpy(g["traj.idx"][it]) for it in range(K)])
    monkeypatch.setattr(torch, "multinomial", lambda *a, **k: next(draws))

    def provider(kind, it):
        if kind == "train":
            return _dev(gu.eps_of(g, spec, prefix="traj.eps", start=per * it))
        return _dev(gu.eps_of(g, spec, prefix="traj.val_eps", start=0))

    args.eps_provider = provider
    vae2, logger = dpv.train_model(args, vae, case_mod.definition, (x, c, y), (x, c, y))
    sc = logger.experiment.scalars
    for nme in gu.TRAIN_LOG:
        for (it, v), r in zip(sc[nme], g[f"traj.log.{nme}"]):
            assert abs(float(v) - float(r)) < TOL * max(1.0, abs(float(r))), (nme, it, float(v), float(r))
    for k, p in vae.named_parameters():
        if p.requires_grad:
            err = gu.rel_l2(p.detach().cpu(), g[f"traj.final.{k}"])
            assert err < 1e-4, (k, err)
