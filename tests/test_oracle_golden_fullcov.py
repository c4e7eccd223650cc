"""Pin the oracle's full-covariance conditional priors (`--full_cov_prior True`, dpivae.py:151-153, models/vae.py:202-203)
against fixtures of the unmodified reference (tests/golden/make_golden_fullcov.py): loss 8-tuple, scalars, every gradient,
`forward(cond=True)` and a `train_model` trajectory."""
import pytest
import torch

import golden_util as gu
from oracle import dpivae_oracle as orc

TOL = 1e-5


def _setup(case, mtype, dtype=torch.float32):
    g, spec, sd = gu.load(case, mtype, ext="fullcov")
    assert spec["full_cov_prior"]
    spec = orc.cast_spec(spec, dtype)
    sd = {k: v.to(dtype) for k, v in sd.items()}
    x, c, y = (torch.from_numpy(g[k]).to(dtype) for k in "xcy")
    return g, spec, sd, x, c, y


def _eps(g, spec, prefix, start=0):
    return gu.eps_of(g, spec, prefix=prefix, start=start)


@pytest.mark.parametrize("case,mtype", gu.FULLCOV_CONFIGS)
def test_loss_and_grads(case, mtype):
    g, spec, sd, x, c, y = _setup(case, mtype)
    assert "prior_net_c.net.f_cov.weight" in spec["trainable"]
    scal, loss8, fw, grads = orc.loss_and_grads(sd, spec, x, c, y, _eps(g, spec, "loss.eps"))
    for name, t in zip(gu.L8_NAMES, loss8):
        assert gu.rel_l2(t, g[f"loss.loss8.{name}"]) < TOL, name
    for a, b in zip(scal, g["loss.scalars"]):
        assert abs(float(a) - float(b)) <= TOL * max(1.0, abs(float(b)))
    bad = {k: gu.rel_l2(grads[k], g[f"loss.grad.{k}"]) for k in spec["trainable"]}
    bad = {k: v for k, v in bad.items() if v > 2e-5}
    assert not bad, bad
    # the strict lower triangle of the prior factor carries gradient, the rest of f_cov none
    gc = torch.from_numpy(g["loss.grad.prior_net_c.net.f_cov.bias"]).reshape(spec["nz_c"], spec["nz_c"])
    assert float(gc.tril(-1).abs().sum()) > 0.0 and float(gc.triu(0).abs().sum()) == 0.0


@pytest.mark.parametrize("case,mtype", gu.FULLCOV_CONFIGS)
def test_cond_forward_and_prior_net(case, mtype):
    g, spec, sd, x, c, y = _setup(case, mtype)
    per = 3 if mtype == "P" else 1
    fw = orc.forward(sd, spec, x, c, _eps(g, spec, "cond.eps"), cond=True, eps_cond=torch.from_numpy(g[f"cond.eps{per}"]))
    for name, t in zip(gu.FW_NAMES, fw):
        assert gu.rel_l2(t, g[f"cond.fw.{name}"]) < TOL, name
    c_t = orc.standardise(c, spec["mean_c"], spec["std_c"])
    loc, tril = orc.prior_heads(sd, spec, "prior_net_c", c_t, spec["nz_c"])
    assert gu.rel_l2(loc, g["pnet.loc_c"]) < TOL and gu.rel_l2(tril, g["pnet.tril_c"]) < TOL
    assert float(torch.from_numpy(g["pnet.tril_c"]).tril(-1).abs().sum()) > 0.0


@pytest.mark.parametrize("case,mtype", gu.FULLCOV_CONFIGS)
def test_trajectory(case, mtype):
    g, spec, sd, x, c, y = _setup(case, mtype)
    K = int(g["traj.K"])
    per = 3 if mtype == "P" else 1
    batches, eps_list = [], []
    for it in range(K):
        idx = torch.from_numpy(g["traj.idx"][it])
        batches.append((x[idx], c[idx], y[idx]))
        eps_list.append(_eps(g, spec, "traj.eps", per * it))
    lr = {k: (5e-3 if k == "log_sigma_x" else 1e-3) for k in spec["trainable"]}
    wd = {k: 0.0 for k in spec["trainable"]}
    final, hist = orc.train_steps(sd, spec, batches, eps_list, lr, wd)
    for it in range(K):
        for j, nme in enumerate(["ELBO", "KLx", "KLc", "KLy", "Rx", "Rc", "Ry", "reg"]):
            ref = g[f"traj.log.{nme}"][it]
            assert abs(hist[it][j] - ref) < TOL * max(1.0, abs(ref)), (it, nme, hist[it][j], ref)
    for k in spec["trainable"]:
        assert gu.rel_l2(final[k], g[f"traj.final.{k}"]) < 1e-4, k
