"""CPU oracle for the DPI-VAE training step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A closed-form restatement, in plain torch tensor algebra on the CPU (fp32 or fp64), of the
reference's `DPIVAE.loss -> backward -> Adam.step` path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this
module; nothing under `dpivae_b200/` does (the product path is CUDA only and fails loudly when
its extension is missing).

Parity pin: the reference ships no tests / golden vectors for this path.
This restatement is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the
build container by `tests/golden/make_golden.py` (imports the unmodified reference through
`tools/ref_harness.py`) and committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py`
checks every function here against those fixtures.

Reference lines each function follows (paths relative to the reference root):
  standardise            utils/transforms.py:64-73, models/vae.py:72-97
  full_cov_heads         models/encoders.py:33-44
  factorized_heads       models/encoders.py:121-128
  sample_latent          models/encoders.py:73-93  (+ torch MultivariateNormal.rsample/log_prob)
  logistic_shift_scale   utils/transforms.py:97-100,124-133,145-150,170-177
  physics_*              models/nn.py:67-80, cases/damped_oscillator/mass_spring.py:8-28,
                         cases/simple_beam/simple_beam_model.py:4-30
  decoders               models/decoders.py:36-49,79-92 ; GRL utils/transforms.py:202-238
  log_prior_zx           utils/priors.py:19-23 (+ torch Uniform/Normal.log_prob)
  loss                   models/vae.py:160-231
  normalise              dpivae.py:419-426
  adam_step              dpivae.py:335-373,436 (+ torch.optim.Adam, non-amsgrad, L2 weight decay)
  clip_grad_norm         dpivae.py:432-433 (+ torch.nn.utils.clip_grad_norm_)

A second fixture set (`tests/golden/make_golden_ext.py` -> `*_ext.npz`, `tests/test_oracle_golden_ext.py`) pins the
branches the reference's default run never takes: `cond=True`, `lambda_x`, weight decay, gradient clipping, annealed
loss weights, the validation pass, clamp saturation and the Uniform-prior `-inf` edge.
"""
import math

import torch

LOG_2PI = math.log(2.0 * math.pi)
LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


# ----------------------------------------------------------------------------------------------
# small pieces
# ----------------------------------------------------------------------------------------------
def standardise(v, mean, std):
    """utils/transforms.py:70-73 -- (v - mean) / population-std."""
    return (v - mean) / std


def linear(sd, name, v):
    """nn.Linear: v @ W^T + b with W (out,in)."""
    return v @ sd[name + ".weight"].T + sd[name + ".bias"]


def full_cov_heads(sd, prefix, x_t, nz, jitter=1e-8):
    """models/encoders.py:33-44.  Returns loc (B,nz), scale_tril (B,nz,nz), eps-independent."""
    h = torch.relu(linear(sd, prefix + ".net.net.encoder_linear_0", x_t))
    loc = linear(sd, prefix + ".net.f_mean", h).clamp(-50.0, 50.0)
    sigma = torch.exp(linear(sd, prefix + ".net.f_sigma", h).clamp(-7.0, 3.0))
    L = torch.tril(linear(sd, prefix + ".net.f_cov", h).clamp(-20.0, 20.0).reshape(-1, nz, nz), diagonal=-1)
    scale_tril = L + torch.diag_embed(sigma + jitter)
    return loc, scale_tril


def factorized_heads(sd, prefix, v_t, jitter=1e-8):
    """models/encoders.py:121-128.  Returns loc (B,nz) and the DIAGONAL sigma+jitter (B,nz)."""
    h = torch.relu(linear(sd, prefix + ".net.net.encoder_linear_0", v_t))
    loc = linear(sd, prefix + ".net.f_mean", h).clamp(-50.0, 50.0)
    sigma = torch.exp(linear(sd, prefix + ".net.f_sigma", h).clamp(-7.0, 3.0))
    return loc, sigma + jitter


def prior_heads(sd, spec, prefix, v_t, nz):
    """Conditional prior net -> (loc (B,nz), scale_tril (B,nz,nz)): FactorizedNN by default (dpivae.py:155-157),
    FullCovarianceNN with `--full_cov_prior True` (dpivae.py:151-153)."""
    if spec.get("full_cov_prior"):
        return full_cov_heads(sd, prefix, v_t, nz)
    loc, sig = factorized_heads(sd, prefix, v_t)
    return loc, torch.diag_embed(sig)


def tril_mvn_log_prob(z, loc, tril):
    """MultivariateNormal(loc, scale_tril=tril).log_prob(z) (models/vae.py:202-203): z (n,B,nz), loc (B,nz), tril (B,nz,nz).
    t = L^-1 (z - loc) by forward substitution; log p = -1/2 (nz log 2 pi + |t|^2) - sum log diag L."""
    nz = loc.shape[-1]
    d = (z - loc.unsqueeze(0)).unsqueeze(-1)                                   # (n,B,nz,1)
    t = torch.linalg.solve_triangular(tril.unsqueeze(0).expand(z.shape[0], -1, -1, -1), d, upper=False).squeeze(-1)
    hld = torch.diagonal(tril, dim1=-2, dim2=-1).log().sum(-1)
    return -0.5 * (nz * LOG_2PI + (t * t).sum(-1)) - hld.unsqueeze(0)


def sample_latent(loc, scale_tril, eps):
    """models/encoders.py:84-86 with injected eps (n,B,nz).

    z = loc + L eps;  log q(z) = -1/2 ||L^-1 (z-loc)||^2 - sum log diag L - nz/2 log 2pi.
    L^-1 (z - loc) == eps identically, so the Mahalanobis term is -1/2 ||eps||^2 and carries no
    gradient (the reference's autograd produces the same up to round-off noise)."""
    nz = loc.shape[-1]
    z = loc.unsqueeze(0) + torch.einsum("bij,nbj->nbi", scale_tril, eps)
    half_log_det = torch.diagonal(scale_tril, dim1=-2, dim2=-1).log().sum(-1)  # (B,)
    log_q = -0.5 * (nz * LOG_2PI + (eps**2).sum(-1)) - half_log_det.unsqueeze(0)
    return z, log_q


def logistic_shift_scale(z, lb, ub, k=1.0):
    """Logistic(k) then ShiftScale(lb,ub): utils/transforms.py:124-133,97-100.

    Returns transformed z and the summed log|det J| (n,B)."""
    ld1 = (k * z - 2.0 * torch.nn.functional.softplus(k * z) + math.log(k)).sum(-1)
    u = torch.sigmoid(k * z)
    a = ub - lb
    out = u * a + lb
    ld2 = (torch.log(torch.abs(a)) * torch.ones_like(out)).sum(-1)
    return out, ld1 + ld2


def log_prior_zx(zx, prior_x):
    """utils/priors.py:19-23 summed over the last dim (models/vae.py:201).

    prior_x: list of ("uniform", low, high) | ("normal", loc, scale)."""
    tot = torch.zeros(zx.shape[:-1], dtype=zx.dtype, device=zx.device)
    for i, (kind, a, b) in enumerate(prior_x):
        zi = zx[..., i]
        if kind == "uniform":
            inside = (zi >= a) & (zi < b)  # torch Uniform.log_prob: lb.mul(ub).log() - log(high-low)
            tot = tot + torch.where(inside, torch.zeros_like(zi), torch.full_like(zi, -math.inf)) - math.log(b - a)
        elif kind == "normal":
            tot = tot + (-((zi - a) ** 2) / (2.0 * b * b) - math.log(b) - LOG_SQRT_2PI)
        else:
            raise ValueError(kind)
    return tot


def normal_log_prob(value, loc, log_scale):
    """torch Normal(loc, exp(log_scale)).log_prob(value)."""
    var = torch.exp(log_scale) ** 2
    return -((value - loc) ** 2) / (2.0 * var) - log_scale - LOG_SQRT_2PI


def diag_mvn_log_prob(z, loc, sigma):
    """MultivariateNormal(loc, scale_tril=diag(sigma)).log_prob(z) (models/vae.py:202-203)."""
    nz = loc.shape[-1]
    m = (((z - loc) / sigma) ** 2).sum(-1)
    return -0.5 * (nz * LOG_2PI + m) - sigma.log().sum(-1)


class _GradRev(torch.autograd.Function):
    """utils/transforms.py:202-219: identity forward, grad_in = -grad_out * alpha."""

    @staticmethod
    def forward(ctx, v, alpha):
        ctx.alpha = alpha
        return v.view_as(v)

    @staticmethod
    def backward(ctx, g):
        return -g * ctx.alpha, None


# ----------------------------------------------------------------------------------------------
# physics decoders
# ----------------------------------------------------------------------------------------------
def physics_mlp(phys, zx_in):
    """models/nn.py:67-80 with StandardScaler input and Tanh hidden layers (bridge part_model)."""
    h = (zx_in - phys["in_mean"]) / phys["in_std"]
    nl = len(phys["w"])
    for i in range(nl):
        h = h @ phys["w"][i].T + phys["b"][i]
        if i < nl - 1:
            h = torch.tanh(h)
    return h


def physics_mass_spring(phys, zx_in):
    """cases/damped_oscillator/mass_spring.py:8-28: x = (0/w) sin(w t) + 1*cos(w t), w = sqrt(1/m)."""
    t = phys["t"]
    m = zx_in[..., 0].unsqueeze(-1)
    omega = torch.sqrt(1.0 / m)
    B = 0.0 / omega
    return B * torch.sin(omega * t) + 1.0 * torch.cos(omega * t)


def physics_beam(phys, zx_in, I=2e-6, L=1.0, P=1.0):
    """cases/simple_beam/simple_beam_model.py:4-30 (grid = linspace(0, L, nd_x))."""
    x = phys["t"]
    E = zx_in[..., 0].unsqueeze(-1) * 1e6
    a = zx_in[..., 1].unsqueeze(-1)
    b = L - a
    if bool(torch.any(a < 0.0)) or bool(torch.any(a > L)):
        raise ValueError("Load position must be between 0 and L")
    mask = x > a
    w = P * b * x * (L**2 - b**2 - x**2) / (6 * E * I * L)
    wb = P * ((x - a) ** 3) / (6 * E * I)
    w = w + torch.where(mask, wb, torch.zeros_like(wb))
    return -1000.0 * w


PHYSICS = {"mlp": physics_mlp, "mass_spring": physics_mass_spring, "beam": physics_beam}


# ----------------------------------------------------------------------------------------------
# the step
# ----------------------------------------------------------------------------------------------
def cast_spec(spec, dtype):
    """Tensor-ise the numeric members of a spec dict in `dtype`."""
    out = dict(spec)
    for k in ["mean_x", "std_x", "mean_c", "std_c", "mean_y", "std_y", "lb", "ub"]:
        out[k] = torch.as_tensor(spec[k]).to(dtype)
    phys = dict(spec["physics"])
    for k in ["in_mean", "in_std", "t"]:
        if k in phys and phys[k] is not None:
            phys[k] = torch.as_tensor(phys[k]).to(dtype)
    if "w" in phys:
        phys["w"] = [torch.as_tensor(w).to(dtype) for w in phys["w"]]
        phys["b"] = [torch.as_tensor(b).to(dtype) for b in phys["b"]]
    out["physics"] = phys
    return out


def forward(sd, spec, x, c, eps, cond=False, eps_cond=None):
    """models/vae.py:160-175 (DPIVAE.forward) with injected noise.

    eps: P -> tuple (eps_x (n,B,nz_x), eps_c, eps_y); S -> tensor (n,B,nz_x+nz_c+nz_y).
    Returns the reference's 10-tuple."""
    nz_x, nz_c, nz_y = spec["nz_x"], spec["nz_c"], spec["nz_y"]
    x_t = standardise(x, spec["mean_x"], spec["std_x"])
    if spec["model_type"] == "S":
        Z = nz_x + nz_c + nz_y
        loc, tril = full_cov_heads(sd, "encoder", x_t, Z)
        z, log_q = sample_latent(loc, tril, eps)
        zx_t, ld = logistic_shift_scale(z[..., :nz_x], spec["lb"], spec["ub"])  # ChainTransformMasked
        zx, zc, zy = zx_t, z[..., nz_x:nz_x + nz_c], z[..., nz_x + nz_c:]
        dens_z = log_q - ld
    else:
        eps_x, eps_c, eps_y = eps
        loc_x, tril_x = full_cov_heads(sd, "encoder", x_t, nz_x)
        loc_c, tril_c = full_cov_heads(sd, "encoder_c", x_t, nz_c)
        loc_y, tril_y = full_cov_heads(sd, "encoder_y", x_t, nz_y)
        zx0, lq_x = sample_latent(loc_x, tril_x, eps_x)
        zc, lq_c = sample_latent(loc_c, tril_c, eps_c)
        zy, lq_y = sample_latent(loc_y, tril_y, eps_y)
        zx, ld = logistic_shift_scale(zx0, spec["lb"], spec["ub"])
        dens_z = (lq_x - ld) + lq_c + lq_y
    if cond:
        c_t = standardise(c, spec["mean_c"], spec["std_c"])
        ploc, ptril = prior_heads(sd, spec, "prior_net_c", c_t, nz_c)
        zc = ploc.unsqueeze(0) + torch.einsum("bij,nbj->nbi", ptril, eps_cond)
    n = zx.shape[0]
    c_phys = c[..., spec["idx_c_phys"]].unsqueeze(0).repeat(n, 1, 1)
    zx_in = torch.cat((zx, c_phys), dim=-1)

    # decode (models/vae.py:153-158, models/decoders.py)
    z_rev = torch.cat((zc, zy), dim=-1)
    z_d = _GradRev.apply(z_rev, spec["lambda_g0"])
    xh_d = linear(sd, "decoder_x.fx1", torch.relu(linear(sd, "decoder_x.fx0", z_d)))
    xh_p = PHYSICS[spec["physics"]["kind"]](spec["physics"], zx_in)
    oy = linear(sd, "decoder_y.net.2", torch.relu(linear(sd, "decoder_y.net.0", zy)))
    oc = linear(sd, "decoder_c.net.2", torch.relu(linear(sd, "decoder_c.net.0", zc)))
    nd_c, nd_y = spec["nd_c"], spec["nd_y"]
    return xh_p, xh_d, oc[..., :nd_c], oc[..., nd_c:], oy[..., :nd_y], oy[..., nd_y:], zx, zc, zy, dens_z


def loss(sd, spec, x, c, y, eps, beta_x=1.0, alpha_x=1.0, alpha_c=1.0, alpha_y=1.0):
    """models/vae.py:177-231 (DPIVAE.loss) -> the reference's 8-tuple, plus the forward 10-tuple."""
    fw = forward(sd, spec, x, c, eps)   # loss always calls forward with cond=False (models/vae.py:194)
    xh_p, xh_d, ch, lsc, yh, lsy, zx, zc, zy, dens_z = fw
    xh = xh_p + xh_d
    c_t = standardise(c, spec["mean_c"], spec["std_c"])
    y_t = standardise(y, spec["mean_y"], spec["std_y"])
    if spec.get("full_cov_prior"):
        ploc_c, ptril_c = prior_heads(sd, spec, "prior_net_c", c_t, spec["nz_c"])
        ploc_y, ptril_y = prior_heads(sd, spec, "prior_net_y", y_t, spec["nz_y"])
        log_prior_z = log_prior_zx(zx, spec["prior_x"]) + tril_mvn_log_prob(zc, ploc_c, ptril_c) \
            + tril_mvn_log_prob(zy, ploc_y, ptril_y)
    else:
        ploc_c, psig_c = factorized_heads(sd, "prior_net_c", c_t)
        ploc_y, psig_y = factorized_heads(sd, "prior_net_y", y_t)
        log_prior_z = log_prior_zx(zx, spec["prior_x"]) + diag_mvn_log_prob(zc, ploc_c, psig_c) \
            + diag_mvn_log_prob(zy, ploc_y, psig_y)
    KL_x = torch.mean(dens_z - log_prior_z, dim=0)
    R_x = normal_log_prob(x, xh, sd["log_sigma_x"]).sum(-1).mean(0)
    R_c = normal_log_prob(c, ch, lsc).sum(-1).mean(0)
    R_y = normal_log_prob(y, yh, lsy).sum(-1).mean(0)
    reg = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
    if spec.get("lambda_x") is not None:
        lam = float(spec["lambda_x"])
        reg = reg + (-(xh_d**2) / (2.0 * lam * lam) - math.log(lam) - LOG_SQRT_2PI).sum(-1).mean(0)
    total = beta_x * KL_x - alpha_x * R_x - alpha_c * R_c - alpha_y * R_y - reg
    zero = torch.tensor(0.0, dtype=x.dtype, device=x.device)
    return (total, KL_x, zero, zero, R_x, R_c, R_y, reg), fw


def normalise(loss8, n_batch, nd_sum):
    """dpivae.py:419-426: ELBO / (n_batch*(nd_x+nd_y+nd_c)), the rest / n_batch."""
    out = [loss8[0].sum() / (n_batch * nd_sum)]
    for t in loss8[1:]:
        out.append(t.sum() / n_batch)
    return out


def loss_and_grads(sd, spec, x, c, y, eps, n_batch=None, **kw):
    """One forward+backward.  sd tensors listed in spec['trainable'] get .requires_grad.

    Returns (scalars[8], per-row 8-tuple, forward 10-tuple, grads dict)."""
    names = spec["trainable"]
    sd = {k: (v.detach().clone().requires_grad_(k in names)) for k, v in sd.items()}
    loss8, fw = loss(sd, spec, x, c, y, eps, **kw)
    nb = x.shape[0] if n_batch is None else n_batch
    scal = normalise(loss8, nb, spec["nd_x"] + spec["nd_c"] + spec["nd_y"])
    scal[0].backward()
    grads = {k: (sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])) for k in names}
    return [s.detach() for s in scal], [t.detach() for t in loss8], [t.detach() for t in fw], grads


def adam_step(params, grads, m, v, step, lr, wd, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (non-amsgrad, L2 decay), in place.

    step is 1-based.  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""
    for k in params:
        g = grads[k]
        if wd[k] != 0.0:
            g = g + wd[k] * params[k]
        m[k].mul_(beta1).add_(g, alpha=1.0 - beta1)
        v[k].mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        bc1 = 1.0 - beta1**step
        bc2 = 1.0 - beta2**step
        denom = (v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].addcdiv_(m[k], denom, value=-(lr[k] / bc1))


def clip_grad_norm(grads, max_norm, norm_eps=1e-6):
    """torch.nn.utils.clip_grad_norm_ (dpivae.py:432-433): one global L2 norm over every gradient tensor,
    coefficient max_norm / (norm + 1e-6) clamped to 1.  Returns (clipped grads, total norm)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(next(iter(grads.values())).dtype)
    coef = torch.clamp(max_norm / (total + norm_eps), max=1.0)
    return {k: g * coef for k, g in grads.items()}, float(total)


def train_steps(sd, spec, batches, eps_list, lr, wd, n_batch=None, max_grad_norm=None, weights=None, on_step=None, **kw):
    """K Adam steps from `sd` over fixed minibatches / injected noise (dpivae.py:390-436).

    weights: optional per-step list of dict(beta_x=, alpha_x=, alpha_c=, alpha_y=) (annealing, dpivae.py:397-400);
    max_grad_norm: clip_grad_norm_ before the step (dpivae.py:432-433); on_step(it, sd) runs after the optimizer step
    (the validation pass of dpivae.py:454-501 reads the updated parameters)."""
    names = spec["trainable"]
    sd = {k: v.detach().clone() for k, v in sd.items()}
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    hist = []
    for it, ((x, c, y), eps) in enumerate(zip(batches, eps_list)):
        kw_it = dict(kw)
        if weights is not None:
            kw_it.update(weights[it])
        scal, _, _, grads = loss_and_grads(sd, spec, x, c, y, eps, n_batch=n_batch, **kw_it)
        if max_grad_norm is not None:
            grads, _ = clip_grad_norm(grads, max_grad_norm)
        adam_step({k: sd[k] for k in names}, grads, m, v, it + 1, lr, wd)
        hist.append([float(s) for s in scal])
        if on_step is not None:
            on_step(it, sd)
    return sd, hist
