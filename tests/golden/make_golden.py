"""Generate golden vectors by RUNNING THE UNMODIFIED REFERENCE in the build container.

    python tests/golden/make_golden.py

Container-only: imports /root/reference through tools/ref_harness.py (CPU, fp32).  For each
case x model type it records, from the reference's own `setup_model`, `DPIVAE.loss`,
`DPIVAE.forward`, autograd and `train_model`:

  * the spec (dims, bounds, scaler statistics fitted by setup_model, prior kinds),
  * the initial state_dict (trainable tensors) produced with `--use_seed` seed 123,
  * a seeded minibatch (x, c, y) from the reference's `sample_response`,
  * the injected reparameterisation noise eps (captured from torch's `_standard_normal`),
  * the 8-tuple per-row loss, the 10-tuple forward outputs, the normalised scalars,
  * autograd gradients of the normalised ELBO for every trainable tensor,
  * for the trajectory configs: K `train_model` iterations (reference Adam, param groups,
    minibatch indices from torch.multinomial recorded) -> final parameters + logged scalars.

The fixtures pin oracle/dpivae_oracle.py (tests/test_oracle_golden.py) and are the reference
side of the CUDA parity tests (tests/test_gpu_parity.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
import ref_harness  # noqa: E402

CONFIGS = [
    # (case, preset, B, n_mc, K trajectory steps (0 = none))
    ("bridge", "DPIVAE-A", 24, 4, 5),
    ("bridge", "DPIVAE-B", 24, 4, 0),
    ("damped_oscillator", "vae", 24, 4, 5),
    ("damped_oscillator", "dpivae", 24, 4, 0),
    ("simple_beam", "vae", 24, 4, 0),
    ("simple_beam", "dpivae", 24, 4, 5),
]


class NoiseTap:
    """Replace torch.distributions' _standard_normal with a recorded, seeded source."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.draws = []

    def __call__(self, shape, dtype, device):
        e = torch.randn(tuple(shape), generator=self.gen, dtype=dtype)
        self.draws.append(e.clone())
        return e


def spec_from(vae, definition, args):
    def vec(t):
        return t.detach().cpu().numpy().astype(np.float32).reshape(-1)

    prior = []
    for d in vae.prior_x.distributions:
        if isinstance(d, torch.distributions.Uniform):
            prior.append((0.0, float(d.low), float(d.high)))
        else:
            prior.append((1.0, float(d.loc), float(d.scale)))
    dpx = definition["dict_prior_x"]
    return {
        "model_type": np.array(args.model_type),
        "dims": np.array([vae.nz_x, vae.nz_c, vae.nz_y, vae.nd_x, vae.nd_c, vae.nd_y], dtype=np.int64),
        "idx_c_phys": np.array(vae.idx_c_phys, dtype=np.int64),
        "lambda_g0": np.array(args.lambda_g0, dtype=np.float64),
        "lb": np.array([v["lb"] for v in dpx.values()], dtype=np.float32),
        "ub": np.array([v["ub"] for v in dpx.values()], dtype=np.float32),
        "prior_x": np.array(prior, dtype=np.float64),
        "mean_x": vec(vae.transform_x.mean_), "std_x": vec(vae.transform_x.scale_),
        "mean_c": vec(vae.transform_c.mean_), "std_c": vec(vae.transform_c.scale_),
        "mean_y": vec(vae.transform_y.mean_), "std_y": vec(vae.transform_y.scale_),
    }


def main():
    import torch.distributions.multivariate_normal as mvn_mod
    import torch.distributions.utils as dutils

    for case_name, preset, B, n_mc, K in CONFIGS:
        dp, case = ref_harness.load(case_name)
        from utils.data import sample_response
        from utils.priors import get_prior_dist

        definition = case.definition
        n_train = B
        args = ref_harness.make_args(case, preset, use_seed=True, seed=123, n_train=n_train, n_batch=B,
                                     n_val=8, n_mc_train=n_mc, n_mc_val=2, n_iter=max(K, 1), val_freq=1000)
        torch.manual_seed(7)
        data = sample_response(definition, n_train, sample_dist=get_prior_dist(definition["dict_gt"]))
        data_val = sample_response(definition, 8, sample_dist=get_prior_dist(definition["dict_gt"]))
        vae = dp.setup_model(args, definition, data)
        for p in getattr(vae.decoder_x.model, "parameters", lambda: [])():
            p.requires_grad = False
        trainable = [k for k, p in vae.named_parameters() if p.requires_grad]
        out = {f"spec.{k}": v for k, v in spec_from(vae, definition, args).items()}
        out["trainable"] = np.array(trainable)
        for k in trainable:
            out[f"init.{k}"] = vae.state_dict()[k].detach().numpy().astype(np.float32).copy()
        x, c, y = data[0].clone(), data[1].clone(), data[2].clone()
        out["x"], out["c"], out["y"] = x.numpy(), c.numpy(), y.numpy()

        # ---- single step: loss / forward / grads with tapped noise -------------------------
        tap = NoiseTap(1234)
        orig = mvn_mod._standard_normal
        mvn_mod._standard_normal = tap
        try:
            loss8 = vae.loss(x, c, y, n=n_mc)
        finally:
            mvn_mod._standard_normal = orig
        eps = tap.draws
        for i, e in enumerate(eps):
            out[f"eps{i}"] = e.numpy()
        nd_sum = vae.nd_x + vae.nd_c + vae.nd_y
        elbo = loss8[0].sum() / (B * nd_sum)
        vae.zero_grad()
        elbo.backward()
        names8 = ["loss", "KLx", "KLc", "KLy", "Rx", "Rc", "Ry", "reg"]
        for nme, t in zip(names8, loss8):
            out[f"loss8.{nme}"] = t.detach().numpy().astype(np.float32)
        out["scalars"] = np.array([float(elbo)] + [float(t.sum() / B) for t in loss8[1:]], dtype=np.float64)
        for k, p in vae.named_parameters():
            if p.requires_grad:
                out[f"grad.{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
        # forward 10-tuple with the same noise
        tap2 = NoiseTap(1234)
        mvn_mod._standard_normal = tap2
        try:
            with torch.no_grad():
                fw = vae.forward(x, c, cond=False, n=n_mc)
        finally:
            mvn_mod._standard_normal = orig
        for nme, t in zip(["xh_p", "xh_d", "ch", "log_sigma_c", "yh", "log_sigma_y", "zx", "zc", "zy", "dens_z"], fw):
            out[f"fw.{nme}"] = t.detach().numpy().astype(np.float32)

        # ---- K-step trajectory through the reference's train_model --------------------------
        if K > 0:
            tap3 = NoiseTap(99)
            idx_log = []
            orig_mult = torch.multinomial

            def mult(*a, **k):
                r = orig_mult(*a, **k)
                idx_log.append(r.clone())
                return r

            mvn_mod._standard_normal = tap3
            torch.multinomial = mult
            try:
                torch.manual_seed(11)
                vae2, logger = dp.train_model(args, vae, definition, data, data_val)
            finally:
                mvn_mod._standard_normal = orig
                torch.multinomial = orig_mult
            # draws: per iteration P:3 / S:1 training draws; iteration 0 also validates (after the
            # optimizer step) with P:3 / S:1 extra draws of shape (n_mc_val, n_val, .)
            per = 3 if args.model_type == "P" else 1
            tr_eps = [d for d in tap3.draws if d.shape[0] == n_mc and d.shape[1] == B]
            assert len(tr_eps) == per * K, (len(tr_eps), per, K)
            for i, e in enumerate(tr_eps):
                out[f"traj.eps{i}"] = e.numpy()
            out["traj.idx"] = torch.stack(idx_log).numpy()
            out["traj.K"] = np.array(K)
            for k in trainable:
                out[f"traj.final.{k}"] = vae2.state_dict()[k].detach().numpy().astype(np.float32).copy()
            for nme in ["ELBO", "KLx", "Rx", "Rc", "Ry", "sigma_x"]:
                out[f"traj.log.{nme}"] = np.array([v for _, v in logger.experiment.scalars[nme]], dtype=np.float64)
            out["traj.lr"] = np.array([args.lr_ex, args.lr_ec, args.lr_ey, args.lr_e, args.lr_p, args.lr_dx,
                                       args.lr_dc, args.lr_dy, args.lr_sigma])

        path = os.path.join(HERE, f"{case_name}_{args.model_type}.npz")
        np.savez_compressed(path, **out)
        print(case_name, preset, args.model_type, "->", os.path.basename(path), os.path.getsize(path),
              "ELBO", float(elbo))


if __name__ == "__main__":
    main()
