"""On-device synthetic data generator (include/dpivae_b200.h dpivae_sample_response):
`sample_response` (utils/data.py:9-52) for the reference's cases, whose `full_model` is a Tanh-MLP surrogate behind a
StandardScaler.  Draws come from the torch CUDA generator's Philox stream (same numbers as torch.rand / torch.randn
called in the reference's order on that generator); the generator offset is advanced accordingly."""
import ctypes as C

import torch

from . import _lib


def _desc(definition):
    fm = definition["full_model"]
    if getattr(fm, "physics_kind", None) != "mlp":
        raise ValueError("the device generator needs a Tanh-MLP `full_model` (SurrogateMLP)")
    gt = definition["dict_gt"]
    d = _lib.DataGenDesc()
    keys = list(gt.keys())
    d.n_factors = len(keys)
    for j, k in enumerate(keys):
        if gt[k]["dist"] is not torch.distributions.Uniform:
            raise ValueError("the device generator supports Uniform ground-truth factors")
        d.lo[j], d.hi[j] = float(gt[k]["args"]["low"]), float(gt[k]["args"]["high"])
    lin = fm.linear_layers()
    d.n_layers = len(lin)
    d.dims[0] = lin[0].in_features
    for i, l in enumerate(lin):
        d.dims[i + 1] = l.out_features
    mean = fm.input_transform.mean_.reshape(-1).tolist()
    std = fm.input_transform.scale_.reshape(-1).tolist()
    for j in range(d.n_factors):
        d.in_mean[j], d.in_std[j] = mean[j], std[j]
    idx_c = [i for i, v in enumerate(gt.values()) if v["type"] == "c"]
    idx_y = [i for i, v in enumerate(gt.values()) if v["type"] == "y"]
    d.nd_x, d.nd_c, d.nd_y = int(definition["nd_x"]), len(idx_c), len(idx_y)
    for j, v in enumerate(idx_c):
        d.idx_c[j] = v
    for j, v in enumerate(idx_y):
        d.idx_y[j] = v
    d.sigma_x, d.sigma_c, d.sigma_y = (float(definition[k]) for k in ("sigma_x", "sigma_c", "sigma_y"))
    return d, lin


def sample_response_device(definition, n, device=None):
    """-> (x (n, nd_x), c (n, nd_c), y (n, nd_y), z (n, n_factors)) on the CUDA device, like
    `sample_response(definition, n, get_prior_dist(definition["dict_gt"]))`."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    d, lin = _desc(definition)
    with torch.cuda.device(dev):
        w = torch.cat([l.weight.detach().reshape(-1).to(dev, torch.float32) for l in lin]).contiguous()
        b = torch.cat([l.bias.detach().reshape(-1).to(dev, torch.float32) for l in lin]).contiguous()
        z = torch.empty((n, d.n_factors), dtype=torch.float32, device=dev)
        x = torch.empty((n, d.nd_x), dtype=torch.float32, device=dev)
        c = torch.empty((n, d.nd_c), dtype=torch.float32, device=dev)
        y = torch.empty((n, d.nd_y), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.dpivae_datagen_workspace_bytes(C.byref(d), n), dtype=torch.uint8, device=dev)
        gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        props = torch.cuda.get_device_properties(dev)
        off_out = C.c_uint64(0)
        _lib.check(lib.dpivae_sample_response(
            C.byref(d), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), int(n), gen.initial_seed(), gen.get_offset(),
            props.multi_processor_count, props.max_threads_per_multi_processor, C.c_void_p(z.data_ptr()), C.c_void_p(x.data_ptr()),
            C.c_void_p(c.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), C.byref(off_out)))
        gen.set_offset(off_out.value)
    return x, c, y, z
