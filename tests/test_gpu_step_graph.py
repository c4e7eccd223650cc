"""Device-resident training loop (dpivae_step_graph_*): a replayed captured step must be the same computation as the
per-step C-ABI calls -- same minibatch rows, same Philox noise stream, same Adam bias corrections."""
import pytest
import torch

import golden_util as gu
from helpers import build_from_golden

pytestmark = pytest.mark.gpu


def _tile(t, k):
    return torch.cat([t] * k, dim=0)


def _run(mode, use_graph, K=7, n=16, clip=0.0):
    from dpivae_b200 import param_groups

    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("bridge", "P")
    eng = vae.engine()
    eng.set_groups(param_groups(args))
    eng.set_math_mode(mode)
    X, C_, Y = _tile(x, 9).cuda(), _tile(c, 9).cuda(), _tile(y, 9).cuda()
    N, B = X.shape[0], 64
    gen = torch.Generator().manual_seed(5)
    pool = torch.stack([torch.multinomial(torch.ones(N), B, False, generator=gen) for _ in range(4)])
    torch.manual_seed(99)
    w = (1.0, 1.0, 1.0, 1.0)
    elbo = []
    if use_graph:
        sg = eng.step_graph(X, C_, Y, n, w, idx_pool=pool, max_grad_norm=clip, log_cap=16, unroll=2)
        sg.run(3)
        sg.run(K - 3)   # a second run re-bases the device state from the host bookkeeping
        rows = sg.log_rows(1, K)
        elbo = rows[:, 0].tolist()
        lsx = rows[:, 8].tolist()
        torch.cuda.synchronize()
        assert eng.step_count == K
        assert abs(lsx[-1] - float(vae.log_sigma_x)) < 1e-7
        sg.close()
    else:
        for it in range(K):
            _, scal = eng.loss(X, C_, Y, n, w, True, idx=pool[it % 4], adam_step=it + 1, max_grad_norm=clip)
            elbo.append(float(scal[0]))
    off = torch.cuda.default_generators[torch.cuda.current_device()].get_offset()
    return {k: p.detach().clone().cpu() for k, p in vae.named_parameters() if p.requires_grad}, elbo, off


@pytest.mark.parametrize("mode,clip", [("fp32", 0.0), ("tc_fp16x3", 0.0), ("tc_fp16x3", 0.5)])
def test_step_graph_matches_per_step_calls(mode, clip):
    pa, ea, offa = _run(mode, False, clip=clip)
    pb, eb, offb = _run(mode, True, clip=clip)
    assert offa == offb   # the torch CUDA generator ends at the same offset
    for a, b in zip(ea, eb):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (ea, eb)
    for k in pa:
        err = gu.rel_l2(pb[k], pa[k])
        assert err < 1e-6, (k, err)


def test_train_model_device_loop_matches_host_loop():
    """train_model (dpivae.py:285-524 mirror): the chunked device-resident loop logs the same curves and ends at the
    same parameters as the per-iteration host loop, including the validation passes in between."""
    import importlib

    import dpivae_b200 as dpv
    from helpers import make_args

    case_mod = importlib.import_module("dpivae_b200.cases.damped_oscillator")
    out = {}
    for dl in (False, True):
        torch.manual_seed(3)
        d = case_mod.definition
        tr = dpv.sample_response(d, 256, sample_dist=dpv.get_prior_dist(d["dict_gt"]))
        va = dpv.sample_response(d, 128, sample_dist=dpv.get_prior_dist(d["dict_gt"]))
        args = make_args(case_mod, "vae", use_seed=True, seed=11, n_train=256, n_val=128, n_batch=64, n_iter=23, val_freq=5,
                         n_mc_train=8, n_mc_val=8, device_loop=dl)
        vae = dpv.setup_model(args, d, tr)
        vae, logger = dpv.train_model(args, vae, d, tr, va)
        sc = logger.experiment.scalars
        out[dl] = ({k: p.detach().clone().cpu() for k, p in vae.named_parameters() if p.requires_grad},
                   {k: list(v) for k, v in sc.items()})
    pa, la = out[False]
    pb, lb = out[True]
    assert set(la) == set(lb)
    for name in la:
        assert [s for s, _ in la[name]] == [s for s, _ in lb[name]], name
        for (_, a), (_, b) in zip(la[name], lb[name]):
            assert abs(a - b) <= 2e-6 * max(1.0, abs(a)), (name, a, b)
    for k in pa:
        assert gu.rel_l2(pb[k], pa[k]) < 2e-6, k


def test_checkpoint_resume_is_exact(tmp_path):
    """dpivae_b200.checkpoint: weights + scaler statistics + Adam state + generator positions;
    a run resumed from the file ends bit-identical to the uninterrupted one."""
    import importlib

    import dpivae_b200 as dpv
    from helpers import make_args

    case_mod = importlib.import_module("dpivae_b200.cases.simple_beam")
    d = case_mod.definition

    def data():
        torch.manual_seed(4)
        return (dpv.sample_response(d, 256, sample_dist=dpv.get_prior_dist(d["dict_gt"])),
                dpv.sample_response(d, 128, sample_dist=dpv.get_prior_dist(d["dict_gt"])))

    def mk(n_iter, start=0):
        return make_args(case_mod, "dpivae", use_seed=True, seed=21, n_train=256, n_val=128, n_batch=64, n_iter=n_iter,
                         val_freq=5, n_mc_train=8, n_mc_val=8, start_iter=start, device_loop=True)

    tr, va = data()
    vae = dpv.setup_model(mk(20), d, tr)
    vae, _ = dpv.train_model(mk(20), vae, d, tr, va)
    full = {k: p.detach().clone().cpu() for k, p in vae.named_parameters() if p.requires_grad}

    tr, va = data()
    vae1 = dpv.setup_model(mk(11), d, tr)
    vae1, _ = dpv.train_model(mk(11), vae1, d, tr, va)
    path = tmp_path / "ckpt.pt"
    dpv.save_checkpoint(path, vae1)
    st = torch.load(path, map_location="cpu", weights_only=False)
    assert st["optim"]["step"] == 11 and set(st["scalers"]) == {"x", "c", "y"}
    assert set(st["optim"]["exp_avg"]) == {k for k, p in vae1.named_parameters() if p.requires_grad}

    tr2, va2 = data()
    torch.manual_seed(12345)   # scramble: everything must come back from the file
    vae2 = dpv.setup_model(mk(20, 11), d, (tr2[0] * 3.0 + 1.0, tr2[1], tr2[2], tr2[3]))   # deliberately wrong scaler fit
    dpv.load_checkpoint(path, vae2)
    vae2, _ = dpv.train_model(mk(20, 11), vae2, d, tr, va)
    for k, p in vae2.named_parameters():
        if p.requires_grad:
            assert torch.equal(p.detach().cpu(), full[k]), k
