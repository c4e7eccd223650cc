"""Device-side post-processing of the hot path's outputs (include/dpivae_b200.h dpivae_mc_mean /
dpivae_regression_metrics / dpivae_linreg_r2): what `evaluate_model` (dpivae.py:527-559, utils/metrics.py:11-32) and
`disentanglement_metric` with the linear regressor (dpivae.py:618-703) compute with numpy / sklearn on the host after
a device->host copy in the reference."""
import ctypes as C

import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _prep(t, dev):
    return t.detach().to(dev, torch.float32).contiguous()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def mc_mean(v):
    """(n, B, d) -> (B, d): mean over the Monte-Carlo axis."""
    lib = _lib.load()
    v = _prep(v, v.device)
    n, B, d = (int(s) for s in v.shape)
    out = torch.empty((B, d), dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(lib.dpivae_mc_mean(_ptr(v), n, B, d, _ptr(out), _stream(v.device)))
    return out


def regression_metrics_device(y_true, y_pred):
    """utils/metrics.py:11-32 -> {"R2", "MSE", "MAE"}, each a numpy array with one value per output column (sklearn
    `multioutput="raw_values"`, the reference's keys and shapes)."""
    lib = _lib.load()
    dev = y_pred.device
    y_pred = _prep(y_pred, dev)
    y_true = _prep(torch.as_tensor(y_true), dev).reshape(y_pred.shape)
    N, d = int(y_pred.shape[0]), int(y_pred.shape[1])
    scratch = torch.empty(4 * d, dtype=torch.float64, device=dev)
    out = torch.empty(3 * d, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dpivae_regression_metrics(_ptr(y_true), _ptr(y_pred), N, d, _ptr(scratch), _ptr(out), _stream(dev)))
    vals = out.cpu().numpy().astype("float64").reshape(3, d)
    return {"R2": vals[0], "MSE": vals[1], "MAE": vals[2]}


def linreg_r2(z_train, t_train, z_test, t_test):
    """R2 on the test set of the least-squares fit t ~ [1, z] for EVERY target column: z_* (N, k <= 8), t_* (N, f)
    -> tensor (f,) on the device (sklearn LinearRegression().fit(z_train, t[:, i]).score(z_test, t_test[:, i]))."""
    lib = _lib.load()
    dev = z_train.device
    z_train, z_test = _prep(z_train, dev), _prep(z_test, dev)
    t_train, t_test = _prep(torch.as_tensor(t_train), dev), _prep(torch.as_tensor(t_test), dev)
    k, f = int(z_train.shape[1]), int(t_train.shape[1])
    out = torch.empty(f, dtype=torch.float32, device=dev)
    scratch = torch.empty((f, 80), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = _stream(dev)
        for i in range(f):
            _lib.check(lib.dpivae_linreg_r2(_ptr(z_train), C.c_void_p(t_train.data_ptr() + 4 * i), f, int(z_train.shape[0]),
                                            _ptr(z_test), C.c_void_p(t_test.data_ptr() + 4 * i), f, int(z_test.shape[0]), k,
                                            _ptr(scratch[i]), C.c_void_p(out.data_ptr() + 4 * i), st))
    return out
