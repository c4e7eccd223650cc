"""Second set of golden vectors from the UNMODIFIED reference: the branches and shapes make_golden.py does not reach.

    python tests/golden/make_golden_ext.py        (container-only: imports /root/reference through tools/ref_harness.py)

Per (case, preset) -> tests/golden/<case>_<P|S>_ext.npz, all with B = 24 rows and n_mc = 8 Monte-Carlo samples (the
tensor-core kernels need 8 <= n_mc <= 128, so every fixture here also runs on the tcgen05 path):

  traj.*    K = 6 iterations of the reference's own `train_model` (dpivae.py:285-524) with the flags its default run
            never touches: weight decay on every group, `clip_gradients` with a max norm small enough to be active,
            cyclical beta_x / sigmoid beta_c,y / cyclical lambda annealing, validation every 2 iterations
            (n_val = 16 rows x n_mc_val = 8).  Recorded: minibatch indices, training and validation noise, all 13
            training and 8 validation scalars the logger holds, final parameters.
  cond.*    `DPIVAE.forward(x, c, cond=True, n=8)` (models/vae.py:160-175): zc from the conditional prior net.
  lamx.*    `DPIVAE.loss` with `lambda_x = 0.7` (models/vae.py:217-219): 8-tuple, scalars, every gradient.
  sat.*     clamp saturation (models/encoders.py:35-39): head biases pushed so that loc / log-sigma / L entries sit
            on and beyond +-50 / -7 / 3 / +-20 for some latent dimensions (decoder first layers scaled by 0.02 so that the
            loss stays ~1e2 .. 1e4): 8-tuple and gradients (zero through a saturated clamp).
  edge.*    sigmoid -> 1.0f: a physics latent lands exactly on the upper bound of its Uniform prior, whose density
            is half-open: log p = -inf, KL = +inf; 8-tuple only.
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
from make_golden import NoiseTap, spec_from  # noqa: E402

CONFIGS = [("bridge", "DPIVAE-A"), ("bridge", "DPIVAE-B"), ("damped_oscillator", "vae"), ("simple_beam", "dpivae")]
B, N_MC, K, N_VAL, N_MC_VAL = 24, 8, 6, 16, 8
TRAIN_NAMES = ["ELBO", "KLx", "KLc", "KLy", "Rx", "Rc", "Ry", "reg", "lambda_x", "beta_x", "beta_c", "beta_y", "sigma_x"]
VAL_NAMES = ["ELBO_val", "KLx_val", "KLc_val", "KLy_val", "Rx_val", "Rc_val", "Ry_val", "reg_val"]
L8 = ["loss", "KLx", "KLc", "KLy", "Rx", "Rc", "Ry", "reg"]
FW = ["xh_p", "xh_d", "ch", "log_sigma_c", "yh", "log_sigma_y", "zx", "zc", "zy", "dens_z"]
TRAJ_FLAGS = dict(wd_e=1e-2, wd_p=2e-2, wd_dx=5e-3, wd_dc=1e-2, wd_dy=1.5e-2, wd_sigma=1e-3, clip_gradients=True,
                  max_grad_norm=0.05, beta_x_annealing="cyclical", beta_x_n_cycles=2, beta_x_R=0.5,
                  beta_c_annealing="sigmoid", beta_y_annealing="sigmoid", lambda_annealing="cyclical", lambda_n_cycles=3)


def tapped(fn):
    import torch.distributions.multivariate_normal as mvn_mod

    tap = NoiseTap(4321)
    orig = mvn_mod._standard_normal
    mvn_mod._standard_normal = tap
    try:
        res = fn()
    finally:
        mvn_mod._standard_normal = orig
    return res, tap.draws


def loss_and_grads(vae, x, c, y, out, prefix, with_grads=True):
    nd_sum = vae.nd_x + vae.nd_c + vae.nd_y
    loss8, eps = tapped(lambda: vae.loss(x, c, y, n=N_MC))
    for i, e in enumerate(eps):
        out[f"{prefix}.eps{i}"] = e.numpy()
    for nme, t in zip(L8, loss8):
        out[f"{prefix}.loss8.{nme}"] = t.detach().numpy().astype(np.float32)
    elbo = loss8[0].sum() / (x.shape[0] * nd_sum)
    out[f"{prefix}.scalars"] = np.array([float(elbo)] + [float(t.sum() / x.shape[0]) for t in loss8[1:]], dtype=np.float64)
    if with_grads:
        vae.zero_grad()
        elbo.backward()
        for k, p in vae.named_parameters():
            if p.requires_grad:
                out[f"{prefix}.grad.{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    return float(elbo)


def main():
    import torch.distributions.multivariate_normal as mvn_mod

    for case_name, preset in CONFIGS:
        dp, case = ref_harness.load(case_name)
        from utils.data import sample_response
        from utils.priors import get_prior_dist

        definition = case.definition
        args = ref_harness.make_args(case, preset, use_seed=True, seed=321, n_train=B, n_batch=B, n_val=N_VAL,
                                     n_mc_train=N_MC, n_mc_val=N_MC_VAL, n_iter=K, val_freq=2, **TRAJ_FLAGS)
        torch.manual_seed(17)
        prior = get_prior_dist(definition["dict_gt"])
        data = sample_response(definition, B, sample_dist=prior)
        data_val = sample_response(definition, N_VAL, sample_dist=prior)
        vae = dp.setup_model(args, definition, data)
        for p in getattr(vae.decoder_x.model, "parameters", lambda: [])():
            p.requires_grad = False
        trainable = [k for k, p in vae.named_parameters() if p.requires_grad]
        out = {f"spec.{k}": v for k, v in spec_from(vae, definition, args).items()}
        out["trainable"] = np.array(trainable)
        init = {k: v.detach().clone() for k, v in vae.state_dict().items()}
        for k in trainable:
            out[f"init.{k}"] = init[k].numpy().astype(np.float32).copy()
        x, c, y = (t.clone() for t in data[:3])
        out["x"], out["c"], out["y"] = x.numpy(), c.numpy(), y.numpy()
        out["x_val"], out["c_val"], out["y_val"] = (t.numpy().copy() for t in data_val[:3])

        # ---- cond=True forward ----------------------------------------------------------------------------
        def fw_cond():
            with torch.no_grad():
                return vae.forward(x, c, cond=True, n=N_MC)

        fw, eps = tapped(fw_cond)
        for i, e in enumerate(eps):
            out[f"cond.eps{i}"] = e.numpy()
        for nme, t in zip(FW, fw):
            out[f"cond.fw.{nme}"] = t.detach().numpy().astype(np.float32)

        # ---- lambda_x regulariser -------------------------------------------------------------------------
        vae.lambda_x = 0.7
        loss_and_grads(vae, x, c, y, out, "lamx")
        vae.lambda_x = None

        # ---- clamp saturation: head biases beyond the clamp bounds on some latent dimensions ----------------
        encs = [vae.encoder] + ([vae.encoder_c, vae.encoder_y] if args.model_type == "P" else [])
        with torch.no_grad():
            for e_i, enc in enumerate(encs):
                nz = enc.net.f_mean.bias.numel()
                # the physics latents keep their (bounded) range: saturation there is the `edge` fixture below
                lo = vae.nz_x if (args.model_type == "S" or e_i == 0) else 0
                if nz - lo >= 1:
                    enc.net.f_mean.bias[nz - 1] += 70.0          # loc clamps at +50
                    enc.net.f_sigma.bias[nz - 1] -= 12.0         # log sigma clamps at -7
                if nz - lo >= 2:
                    enc.net.f_mean.bias[nz - 2] -= 70.0          # loc clamps at -50
                    enc.net.f_sigma.bias[nz - 2] += 6.0          # log sigma clamps at 3
                    enc.net.f_cov.bias[(nz - 1) * nz + nz - 2] += 40.0   # L[nz-1][nz-2] clamps at +20
                if nz - lo >= 3:
                    enc.net.f_cov.bias[(nz - 1) * nz + nz - 3] -= 40.0   # L[nz-1][nz-3] clamps at -20
            for pn in (vae.prior_net_c, vae.prior_net_y):
                pn.net.f_mean.bias[0] += 80.0                    # prior loc clamps at +50
                pn.net.f_sigma.bias[0] += 9.0                    # prior log sigma clamps at 3
            # latents pinned at +-50 (and spread by sigma = e^3, |L| = 20) would drive the heteroscedastic decoders'
            # log-sigma outputs to +-20 and the loss to ~1e21: the decoders' first layers are scaled down so that the
            # saturated regime stays numerically meaningful (losses ~1e2 .. 1e4)
            for lin in (vae.decoder_x.fx0, vae.decoder_c.net[0], vae.decoder_y.net[0]):
                lin.weight *= 0.02
        sat_sd = {k: v.detach().clone() for k, v in vae.state_dict().items()}
        for k in trainable:
            out[f"sat.init.{k}"] = sat_sd[k].numpy().astype(np.float32).copy()
        loss_and_grads(vae, x, c, y, out, "sat")
        vae.load_state_dict(init)

        # ---- sigmoid -> 1.0f: a physics latent exactly on the upper bound of a Uniform prior ---------------
        if isinstance(vae.prior_x.distributions[0], torch.distributions.Uniform):
            with torch.no_grad():
                vae.encoder.net.f_mean.bias[0] += 45.0
                vae.encoder.net.f_sigma.bias[0] -= 12.0
            edge_sd = {k: v.detach().clone() for k, v in vae.state_dict().items()}
            for k in trainable:
                out[f"edge.init.{k}"] = edge_sd[k].numpy().astype(np.float32).copy()
            loss_and_grads(vae, x, c, y, out, "edge", with_grads=False)
            vae.load_state_dict(init)

        # ---- K-step trajectory through the reference's train_model ------------------------------------------
        tap3 = NoiseTap(77)
        idx_log = []
        orig_mult = torch.multinomial
        orig = mvn_mod._standard_normal

        def mult(*a, **k):
            r = orig_mult(*a, **k)
            idx_log.append(r.clone())
            return r

        mvn_mod._standard_normal = tap3
        torch.multinomial = mult
        try:
            torch.manual_seed(23)
            vae2, logger = dp.train_model(args, vae, definition, data, data_val)
        finally:
            mvn_mod._standard_normal = orig
            torch.multinomial = orig_mult
        per = 3 if args.model_type == "P" else 1
        tr_eps = [d for d in tap3.draws if d.shape[1] == B]
        va_eps = [d for d in tap3.draws if d.shape[1] == N_VAL]
        n_val_passes = len(range(0, K, 2))
        assert len(tr_eps) == per * K and len(va_eps) == per * n_val_passes, (len(tr_eps), len(va_eps))
        for i, e in enumerate(tr_eps):
            out[f"traj.eps{i}"] = e.numpy()
        for i, e in enumerate(va_eps):
            out[f"traj.val_eps{i}"] = e.numpy()
        out["traj.idx"] = torch.stack(idx_log).numpy()
        out["traj.K"] = np.array(K)
        for k in trainable:
            out[f"traj.final.{k}"] = vae2.state_dict()[k].detach().numpy().astype(np.float32).copy()
        for nme in TRAIN_NAMES + VAL_NAMES:
            out[f"traj.log.{nme}"] = np.array([v for _, v in logger.experiment.scalars[nme]], dtype=np.float64)
            out[f"traj.log_iter.{nme}"] = np.array([s for s, _ in logger.experiment.scalars[nme]], dtype=np.int64)
        out["traj.flags"] = np.array([f"{k}={v}" for k, v in sorted(TRAJ_FLAGS.items())])

        path = os.path.join(HERE, f"{case_name}_{args.model_type}_ext.npz")
        np.savez_compressed(path, **out)
        print(case_name, preset, args.model_type, "->", os.path.basename(path), os.path.getsize(path),
              "ELBO log", out["traj.log.ELBO"], "edge KL", out.get("edge.scalars", [None, None])[1])


if __name__ == "__main__":
    main()
