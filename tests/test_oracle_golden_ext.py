"""Pin the oracle's non-default branches against the second fixture set of the unmodified reference
(tests/golden/make_golden_ext.py): cond=True forward, the lambda_x regulariser, clamp saturation, the Uniform-prior
-inf edge, and a `train_model` run with weight decay, gradient clipping, annealed weights and validation passes."""
import argparse
import math

import pytest
import torch

import golden_util as gu
from oracle import dpivae_oracle as orc

TOL = 1e-5


def _setup(case, mtype, dtype=torch.float32, weights=None):
    g, spec, sd = gu.load(case, mtype, ext=True)
    if weights is not None:
        sd = gu.state_of(g, spec, weights)
    spec = orc.cast_spec(spec, dtype)
    sd = {k: v.to(dtype) for k, v in sd.items()}
    x, c, y = (torch.from_numpy(g[k]).to(dtype) for k in "xcy")
    return g, spec, sd, x, c, y


def _eps(g, spec, prefix, start=0, dtype=torch.float32):
    e = gu.eps_of(g, spec, prefix=prefix, start=start)
    return tuple(t.to(dtype) for t in e) if isinstance(e, tuple) else e.to(dtype)


@pytest.mark.parametrize("case,mtype", gu.EXT_CONFIGS)
def test_cond_forward(case, mtype):
    """models/vae.py:165-167: zc drawn from the conditional prior net p(zc | c) with the draw that follows the encoder's."""
    g, spec, sd, x, c, y = _setup(case, mtype)
    per = 3 if mtype == "P" else 1
    eps = _eps(g, spec, "cond.eps")
    fw = orc.forward(sd, spec, x, c, eps, cond=True, eps_cond=torch.from_numpy(g[f"cond.eps{per}"]))
    for name, t in zip(gu.FW_NAMES, fw):
        assert gu.rel_l2(t, g[f"cond.fw.{name}"]) < TOL, name


@pytest.mark.parametrize("case,mtype", gu.EXT_CONFIGS)
@pytest.mark.parametrize("section", ["lamx", "sat"])
def test_loss_and_grads_sections(case, mtype, section):
    g, spec, sd, x, c, y = _setup(case, mtype, weights="sat.init" if section == "sat" else None)
    if section == "lamx":
        spec["lambda_x"] = 0.7
    scal, loss8, fw, grads = orc.loss_and_grads(sd, spec, x, c, y, _eps(g, spec, f"{section}.eps"))
    for name, t in zip(gu.L8_NAMES, loss8):
        assert gu.rel_l2(t, g[f"{section}.loss8.{name}"]) < TOL, name
    for a, b in zip(scal, g[f"{section}.scalars"]):
        assert abs(float(a) - float(b)) <= TOL * max(1.0, abs(float(b)))
    bad = {}
    for k in spec["trainable"]:
        ref = torch.from_numpy(g[f"{section}.grad.{k}"])
        err = gu.rel_l2(grads[k], ref)
        # sat: sigma = e^-7 next to |L| = 20 makes L^-1 huge, and the reference's analytically-zero Mahalanobis gradient
        #  turns into fp32 round-off noise of ~5e-4 relative on the loc / sigma heads: its own fp32
        # gradient is that far from the fp64 value, so this section pins the oracle to 1e-3 (and to EXACT zeros through
        # the saturated clamps, below); the CUDA kernels are compared with the fp64 oracle, which has no such noise
        if err > (1e-3 if section == "sat" else 2e-5):
            bad[k] = err
    assert not bad, bad
    if section == "sat":
        # the saturated heads receive exactly zero gradient (torch clamp backward), in the reference and here
        pre = "encoder_y" if mtype == "P" else "encoder"
        nz = sd[f"{pre}.net.f_mean.bias"].numel()
        assert float(g[f"sat.grad.{pre}.net.f_mean.bias"][nz - 1]) == 0.0 and float(grads[f"{pre}.net.f_mean.bias"][nz - 1]) == 0.0
        assert float(grads[f"{pre}.net.f_sigma.bias"][nz - 1]) == 0.0
        assert float(g["sat.grad.prior_net_c.net.f_mean.bias"][0]) == 0.0 and float(grads["prior_net_c.net.f_mean.bias"][0]) == 0.0


@pytest.mark.parametrize("case,mtype", [c for c in gu.EXT_CONFIGS if c[0] != "simple_beam"])
def test_uniform_prior_edge_gives_inf(case, mtype):
    """sigmoid saturates to 1.0f => zx == high => Uniform.log_prob = -inf => KL = +inf."""
    g, spec, sd, x, c, y = _setup(case, mtype, weights="edge.init")
    loss8, _ = orc.loss(sd, spec, x, c, y, _eps(g, spec, "edge.eps"))
    ref_kl = torch.from_numpy(g["edge.loss8.KLx"])
    assert torch.isinf(ref_kl).any()
    assert torch.equal(torch.isinf(loss8[1]), torch.isinf(ref_kl))
    fin = ~torch.isinf(ref_kl)
    if fin.any():
        assert gu.rel_l2(loss8[1][fin], ref_kl[fin]) < TOL
    for name, t in zip(gu.L8_NAMES, loss8):
        if name in ("Rx", "Rc", "Ry"):
            assert gu.rel_l2(t, g[f"edge.loss8.{name}"]) < TOL, name


def schedule(g, K):
    """Host-side schedule of the flagged run (dpivae_b200.utils.Annealing == utils/annealing.py) -> per-step weights."""
    from dpivae_b200.utils import Annealing

    f = gu.traj_flags(g)
    mk = lambda key, dflt: Annealing(f.get(f"{key}_annealing"), K, n_cycles=f.get(f"{key}_n_cycles", dflt["n_cycles"]),  # noqa: E731
                                     R=f.get(f"{key}_R", 0.5), mu=f.get(f"{key}_mu", dflt["mu"]), cov=f.get(f"{key}_cov", dflt["cov"]))
    d5 = dict(n_cycles=5, mu=0.15, cov=0.15)
    d4 = dict(n_cycles=4, mu=0.2, cov=0.2)
    ann = {"lambda": mk("lambda", d5), "beta_x": mk("beta_x", d5), "beta_c": mk("beta_c", d5), "beta_y": mk("beta_y", d4)}
    return f, ann


@pytest.mark.parametrize("case,mtype", gu.EXT_CONFIGS)
def test_flagged_trajectory_and_validation(case, mtype):
    g, spec, sd, x, c, y = _setup(case, mtype)
    K = int(g["traj.K"])
    per = 3 if mtype == "P" else 1
    f, ann = schedule(g, K)
    batches, eps_list, weights = [], [], []
    for it in range(K):
        idx = torch.from_numpy(g["traj.idx"][it])
        batches.append((x[idx], c[idx], y[idx]))
        eps_list.append(_eps(g, spec, "traj.eps", per * it))
        weights.append(dict(beta_x=float(1.0 * ann["beta_x"].forward(it))))
        # the annealers reproduce the reference's logged schedule
        assert abs(weights[-1]["beta_x"] - g["traj.log.beta_x"][it]) < 1e-6
        assert abs(float(ann["beta_c"].forward(it)) - g["traj.log.beta_c"][it]) < 1e-6
        assert abs(float(ann["beta_y"].forward(it)) - g["traj.log.beta_y"][it]) < 1e-6
        assert abs(float(ann["lambda"].forward(it)) * spec["lambda_g0"] - g["traj.log.lambda_x"][it]) < 1e-6 * max(1.0, abs(spec["lambda_g0"]))
    wdmap = {"encoder": "wd_e", "prior_net": "wd_p", "decoder_x": "wd_dx", "decoder_c": "wd_dc", "decoder_y": "wd_dy", "log_sigma_x": "wd_sigma"}
    wd = {k: f[[v for p, v in wdmap.items() if k.startswith(p)][0]] for k in spec["trainable"]}
    lr = {k: (5e-3 if k == "log_sigma_x" else 1e-3) for k in spec["trainable"]}
    xv, cv, yv = (torch.from_numpy(g[k]) for k in ("x_val", "c_val", "y_val"))
    val, sig = {}, {}
    nd_sum = spec["nd_x"] + spec["nd_c"] + spec["nd_y"]

    def on_step(it, cur):
        sig[it] = float(cur["log_sigma_x"].exp())
        if it % 2 == 0:
            l8, _ = orc.loss(cur, spec, xv, cv, yv, _eps(g, spec, "traj.val_eps", per * (it // 2)), **weights[it])
            val[it] = [float(s) for s in orc.normalise(l8, xv.shape[0], nd_sum)]

    final, hist = orc.train_steps(sd, spec, batches, eps_list, lr, wd, max_grad_norm=f["max_grad_norm"], weights=weights, on_step=on_step)
    names8 = ["ELBO", "KLx", "KLc", "KLy", "Rx", "Rc", "Ry", "reg"]
    for it in range(K):
        for j, nme in enumerate(names8):
            ref = g[f"traj.log.{nme}"][it]
            assert abs(hist[it][j] - ref) < TOL * max(1.0, abs(ref)), (it, nme, hist[it][j], ref)
        assert abs(sig[it] - g["traj.log.sigma_x"][it]) < 1e-5
    for j, it in enumerate(range(0, K, 2)):
        assert int(g["traj.log_iter.ELBO_val"][j]) == it
        for q, nme in enumerate(names8):
            ref = g[f"traj.log.{nme}_val"][j]
            assert abs(val[it][q] - ref) < TOL * max(1.0, abs(ref)), (it, nme, val[it][q], ref)
    for k in spec["trainable"]:
        assert gu.rel_l2(final[k], g[f"traj.final.{k}"]) < 1e-4, k
    # the clip was active: the unclipped first-step gradient norm exceeds max_grad_norm
    _, _, _, g0 = orc.loss_and_grads(sd, spec, *batches[0], eps_list[0], **weights[0])
    assert orc.clip_grad_norm(g0, f["max_grad_norm"])[1] > f["max_grad_norm"]
