"""Golden vectors from the UNMODIFIED reference for `--full_cov_prior True` (dpivae.py:151-153: FullCovarianceNN conditional
prior nets; models/vae.py:202-203: MultivariateNormal(loc, scale_tril).log_prob with a full lower-triangular factor).

    python tests/golden/make_golden_fullcov.py    (container-only: imports /root/reference through tools/ref_harness.py)

Per (case, preset) -> tests/golden/<case>_<P|S>_fullcov.npz, B = 24 rows, n_mc = 8:
  loss.*   `DPIVAE.loss`: injected noise, 8-tuple, normalised scalars, autograd gradient of every trainable tensor
  cond.*   `DPIVAE.forward(x, c, cond=True, n=8)`: zc = prior loc + prior scale_tril eps
  pnet.*   `DPIVAE.prior_net(c, y)`: loc / scale_tril of both conditional priors
  traj.*   K = 5 iterations of the reference's own `train_model` (indices, noise, logged scalars, final parameters)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
from make_golden import NoiseTap, spec_from  # noqa: E402
from make_golden_ext import FW, TRAIN_NAMES, loss_and_grads, tapped  # noqa: E402

CONFIGS = [("bridge", "DPIVAE-A"), ("simple_beam", "dpivae")]
B, N_MC, K = 24, 8, 5


def main():
    import torch.distributions.multivariate_normal as mvn_mod

    for case_name, preset in CONFIGS:
        dp, case = ref_harness.load(case_name)
        from utils.data import sample_response
        from utils.priors import get_prior_dist

        definition = case.definition
        args = ref_harness.make_args(case, preset, use_seed=True, seed=99, n_train=B, n_batch=B, n_val=B, n_mc_train=N_MC,
                                     n_mc_val=N_MC, n_iter=K, val_freq=1000, full_cov_prior=True)
        torch.manual_seed(41)
        prior = get_prior_dist(definition["dict_gt"])
        data = sample_response(definition, B, sample_dist=prior)
        vae = dp.setup_model(args, definition, data)
        for p in getattr(vae.decoder_x.model, "parameters", lambda: [])():
            p.requires_grad = False
        # the off-diagonal factor entries start near zero at the default init: spread them so the triangular solve matters
        with torch.no_grad():
            g = torch.Generator().manual_seed(5)
            for pn in (vae.prior_net_c, vae.prior_net_y):
                pn.net.f_cov.bias += 0.6 * torch.randn(pn.net.f_cov.bias.shape, generator=g)
                pn.net.f_sigma.bias += 0.3 * torch.randn(pn.net.f_sigma.bias.shape, generator=g)
        trainable = [k for k, p in vae.named_parameters() if p.requires_grad]
        out = {f"spec.{k}": v for k, v in spec_from(vae, definition, args).items()}
        out["spec.full_cov_prior"] = np.array(1)
        out["trainable"] = np.array(trainable)
        init = {k: v.detach().clone() for k, v in vae.state_dict().items()}
        for k in trainable:
            out[f"init.{k}"] = init[k].numpy().astype(np.float32).copy()
        x, c, y = (t.clone() for t in data[:3])
        out["x"], out["c"], out["y"] = x.numpy(), c.numpy(), y.numpy()

        loss_and_grads(vae, x, c, y, out, "loss")

        def fw_cond():
            with torch.no_grad():
                return vae.forward(x, c, cond=True, n=N_MC)

        fw, eps = tapped(fw_cond)
        for i, e in enumerate(eps):
            out[f"cond.eps{i}"] = e.numpy()
        for nme, t in zip(FW, fw):
            out[f"cond.fw.{nme}"] = t.detach().numpy().astype(np.float32)
        with torch.no_grad():
            for nme, t in zip(("loc_c", "tril_c", "loc_y", "tril_y"), vae.prior_net(c, y)):
                out[f"pnet.{nme}"] = t.numpy().astype(np.float32)

        tap3 = NoiseTap(78)
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
            torch.manual_seed(29)
            vae2, logger = dp.train_model(args, vae, definition, data, data)
        finally:
            mvn_mod._standard_normal = orig
            torch.multinomial = orig_mult
        # draw order (dpivae.py:390-470): training step 0, the validation pass of iteration 0 (0 % val_freq == 0), then the
        # training steps 1 .. K-1; validation set = training set here, so the draws have the same shape
        per = 3 if args.model_type == "P" else 1
        assert len(tap3.draws) == per * (K + 1), len(tap3.draws)
        tr_eps = tap3.draws[:per] + tap3.draws[2 * per:]
        for i, e in enumerate(tr_eps):
            out[f"traj.eps{i}"] = e.numpy()
        for i, e in enumerate(tap3.draws[per:2 * per]):
            out[f"traj.val_eps{i}"] = e.numpy()
        out["traj.idx"] = torch.stack(idx_log).numpy()
        out["traj.K"] = np.array(K)
        for k in trainable:
            out[f"traj.final.{k}"] = vae2.state_dict()[k].detach().numpy().astype(np.float32).copy()
        for nme in TRAIN_NAMES:
            out[f"traj.log.{nme}"] = np.array([v for _, v in logger.experiment.scalars[nme]], dtype=np.float64)
        path = os.path.join(HERE, f"{case_name}_{args.model_type}_fullcov.npz")
        np.savez_compressed(path, **out)
        print(case_name, preset, args.model_type, "->", os.path.basename(path), os.path.getsize(path), "ELBO log", out["traj.log.ELBO"],
              "n_eps", len(tr_eps))


if __name__ == "__main__":
    main()
