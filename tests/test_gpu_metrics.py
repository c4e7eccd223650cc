"""Device-side post-processing (dpivae_mc_mean / dpivae_regression_metrics / dpivae_linreg_r2) vs numpy / sklearn, the
reference's host implementation of the same steps (utils/metrics.py:11-32, dpivae.py:672-690)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_mc_mean_and_regression_metrics_match_sklearn():
    from sklearn import metrics

    from dpivae_b200.metrics import mc_mean, regression_metrics_device

    g = torch.Generator().manual_seed(0)
    v = torch.randn(37, 1000, 2, generator=g)
    m = mc_mean(v.cuda())
    assert torch.allclose(m.cpu(), v.mean(0), rtol=1e-5, atol=1e-6)
    y = torch.randn(1000, 2, generator=g) * torch.tensor([1.0, 30.0]) + torch.tensor([5.0, -100.0])
    p = y + 0.3 * torch.randn(1000, 2, generator=g)
    out = regression_metrics_device(y, p.cuda())
    # the reference's keys and per-output arrays (utils/metrics.py:29-31, multioutput="raw_values")
    assert set(out) == {"R2", "MSE", "MAE"} and all(v.shape == (2,) for v in out.values())
    ref = {"R2": metrics.r2_score(y.numpy(), p.numpy(), multioutput="raw_values"),
           "MSE": metrics.mean_squared_error(y.numpy(), p.numpy(), multioutput="raw_values"),
           "MAE": metrics.mean_absolute_error(y.numpy(), p.numpy(), multioutput="raw_values")}
    for k in ref:
        assert np.allclose(out[k], ref[k], rtol=1e-5, atol=1e-6), (k, out[k], ref[k])
    # a single column
    out1 = regression_metrics_device(y[:, :1], p[:, :1].cuda())
    assert out1["R2"].shape == (1,) and abs(out1["R2"][0] - metrics.r2_score(y[:, :1].numpy(), p[:, :1].numpy())) < 1e-5


@pytest.mark.parametrize("k", [1, 2, 4, 8])
def test_linreg_r2_matches_sklearn(k):
    from sklearn.linear_model import LinearRegression

    from dpivae_b200.metrics import linreg_r2

    rng = np.random.default_rng(k)
    ntr, nte, f = 2048, 777, 5
    scale = np.array([1.0, 100.0, 0.01, 7.0, 1.0, 3.0, 0.5, 20.0])[:k]
    ztr = (rng.normal(size=(ntr, k)) * scale + 50.0 * scale).astype(np.float32)   # offsets: the intercept matters
    zte = (rng.normal(size=(nte, k)) * scale + 50.0 * scale).astype(np.float32)
    W = rng.normal(size=(k, f)) / scale[:, None]
    ttr = (ztr @ W + 0.5 * rng.normal(size=(ntr, f)) + 3.0).astype(np.float32)
    tte = (zte @ W + 0.5 * rng.normal(size=(nte, f)) + 3.0).astype(np.float32)
    ttr[:, -1] = rng.normal(size=ntr)   # a factor the latents do not explain: R2 ~ 0 or negative
    tte[:, -1] = rng.normal(size=nte)
    got = linreg_r2(torch.from_numpy(ztr).cuda(), torch.from_numpy(ttr), torch.from_numpy(zte).cuda(), torch.from_numpy(tte)).cpu().numpy()
    for i in range(f):
        ref = LinearRegression().fit(ztr, ttr[:, i]).score(zte, tte[:, i])
        assert abs(got[i] - ref) < 1e-4, (i, got[i], ref)


def test_disentanglement_metric_linear_on_device():
    """dpivae.py:618-703 through the mirror: ordering [zx, zc, zy] x factors and values equal to the sklearn path."""
    import importlib

    from sklearn.linear_model import LinearRegression

    import dpivae_b200 as dpv
    from helpers import make_args

    case_mod = importlib.import_module("dpivae_b200.cases.damped_oscillator")
    d = case_mod.definition
    torch.manual_seed(1)
    tr = dpv.sample_response(d, 512, sample_dist=dpv.get_prior_dist(d["dict_gt"]))
    te = dpv.sample_response(d, 256, sample_dist=dpv.get_prior_dist(d["dict_gt"]))
    args = make_args(case_mod, "vae", use_seed=True, seed=2, n_train=512, n_batch=64, n_mc_test=8)
    vae = dpv.setup_model(args, d, tr)
    torch.manual_seed(7)
    scores = dpv.disentanglement_metric(args, vae, d, tr, te, regressor="linear", cond=False, use_mean=True)
    factors = list(d["dict_gt"].keys())
    assert [(s[0], s[1]) for s in scores] == [(t, f) for f in factors for t in ("zx", "zc", "zy")]
    # same latents again (same seed) -> sklearn on the host
    torch.manual_seed(7)
    lat = {}
    for tag, data in (("train", tr), ("test", te)):
        out = vae.sample(data[0], data[1], cond=False, n=8)
        lat[tag] = [out[k].mean(0).cpu().numpy() for k in (5, 6, 7)]
    ztr, zte = tr[3].squeeze(0).cpu().numpy(), te[3].squeeze(0).cpu().numpy()
    it = iter(scores)
    for i in range(len(factors)):
        for gi in range(3):
            ref = LinearRegression().fit(lat["train"][gi], ztr[:, i]).score(lat["test"][gi], zte[:, i])
            got = next(it)[2]
            assert abs(got - ref) < 1e-4 * max(1.0, abs(ref)), (i, gi, got, ref)
