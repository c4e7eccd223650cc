"""DPIVAE.decode / prior_net / sample_prior (models/vae.py:99-123,153-158) through the C ABI vs the oracle."""
import pytest
import torch

import golden_util as gu
from helpers import build_from_golden
from oracle import dpivae_oracle as orc

pytestmark = pytest.mark.gpu
CASES = [("bridge", "P"), ("damped_oscillator", "P"), ("simple_beam", "S"), ("bridge", "S")]


@pytest.mark.parametrize("case,mtype", CASES)
def test_decode_on_given_latents(case, mtype):
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    spec = orc.cast_spec(spec, torch.float32)
    eps = gu.eps_of(g, spec)
    fw = orc.forward(sd, spec, x, c, eps)
    xh_p, xh_d, ch, lsc, yh, lsy, zx, zc, zy, _ = fw
    n = zx.shape[0]
    c_phys = c[..., spec["idx_c_phys"]].unsqueeze(0).repeat(n, 1, 1)
    zx_in = torch.cat((zx, c_phys), dim=-1)
    out = vae.decode(zx_in.cuda(), zc.cuda(), zy.cuda())
    for nm, a, b in zip(("xh_p", "xh_d", "ch", "log_sigma_c", "yh", "log_sigma_y"), out, (xh_p, xh_d, ch, lsc, yh, lsy)):
        assert tuple(a.shape) == tuple(b.shape), nm
        assert gu.rel_l2(a.cpu(), b) < 1e-5, (nm, gu.rel_l2(a.cpu(), b))
    # 2-D latents (B, .) -> 2-D outputs, same numbers as the first MC sample
    out2 = vae.decode(zx_in[0].cuda(), zc[0].cuda(), zy[0].cuda())
    for a, b in zip(out2, out):
        assert torch.equal(a, b[0])
    with pytest.raises(ValueError):
        vae.decode(zx_in[..., :-1].cuda() if zx_in.shape[-1] > 1 else zx_in.cuda()[..., :0], zc.cuda(), zy.cuda())


@pytest.mark.parametrize("case,mtype", CASES[:3])
def test_prior_net_and_sample_prior(case, mtype):
    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden(case, mtype)
    spec = orc.cast_spec(spec, torch.float32)
    c_t = orc.standardise(c, spec["mean_c"], spec["std_c"])
    y_t = orc.standardise(y, spec["mean_y"], spec["std_y"])
    oloc_c, osig_c = orc.factorized_heads(sd, "prior_net_c", c_t)
    oloc_y, osig_y = orc.factorized_heads(sd, "prior_net_y", y_t)
    loc_c, tril_c, loc_y, tril_y = vae.prior_net(c.cuda(), y.cuda())
    assert gu.rel_l2(loc_c.cpu(), oloc_c) < 1e-5 and gu.rel_l2(loc_y.cpu(), oloc_y) < 1e-5
    assert gu.rel_l2(tril_c.cpu(), torch.diag_embed(osig_c)) < 1e-5 and gu.rel_l2(tril_y.cpu(), torch.diag_embed(osig_y)) < 1e-5
    lc, tc_, ly, ty = vae.prior_net(c.cuda())
    assert ly is None and ty is None and torch.equal(lc, loc_c) and torch.equal(tc_, tril_c)

    n, B = 5, c.shape[0]
    torch.manual_seed(31)
    zc, dzc, zy, dzy = vae.sample_prior(c.cuda(), y.cuda(), n=n)
    torch.manual_seed(31)   # the reference's draw order: (n, B, nz_c) then (n, B, nz_y) on the CUDA generator
    eps_c = torch.empty((n, B, spec["nz_c"]), device="cuda").normal_().cpu()
    eps_y = torch.empty((n, B, spec["nz_y"]), device="cuda").normal_().cpu()
    ozc, odc = orc.sample_latent(oloc_c, torch.diag_embed(osig_c), eps_c)
    ozy, ody = orc.sample_latent(oloc_y, torch.diag_embed(osig_y), eps_y)
    assert tuple(zc.shape) == (n, B, spec["nz_c"]) and tuple(dzc.shape) == (n, B)
    for a, b in ((zc, ozc), (dzc, odc), (zy, ozy), (dzy, ody)):
        assert gu.rel_l2(a.cpu(), b) < 1e-5
