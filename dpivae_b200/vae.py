"""`DPIVAE`: drop-in mirror of the reference's `models/vae.py` class whose arithmetic runs in the
hand-written sm_100a kernels of libdpivae_b200.so (no eager PyTorch path, no CPU fallback).

Same constructor, attributes, method signatures and tuple orders as the reference
(models/vae.py:9-255); parameters keep their `state_dict` names but are views into one flat
device buffer, with gradients / Adam moments in parallel flat buffers (the fused optimizer and the
single NCCL allreduce of the data-parallel path operate on those).
"""
import ctypes as C
from contextlib import contextmanager

import numpy as np
import torch
from torch import nn

from . import _lib
from .modules import FactorizedNN, FullCovarianceNN
from .utils import ChainTransform, ChainTransformMasked, Logistic, ShiftScale

_L8 = ("loss", "KLx", "KLc", "KLy", "Rx", "Rc", "Ry", "reg")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


class _Engine:
    """Owns the C handle, the flat buffers and the workspace of one DPIVAE on one CUDA device."""

    def __init__(self, vae, dev):
        self.lib = _lib.load()
        self.dev = torch.device(dev)
        if self.dev.type != "cuda":
            raise RuntimeError("DPIVAE runs on CUDA only (sm_100a kernels; there is no CPU fallback)")
        self.vae = vae
        self.handle = C.c_void_p(None)
        self.workspace = None
        self.step_count = 0
        self._build(vae)

    # -- flat layout --------------------------------------------------------------------------------
    def _build(self, vae):
        units = []  # (name, [(param, ...)] ) in flat order

        def enc_unit(mod):  # GaussianEncoder(FullCovarianceNN | FactorizedNN)
            net = mod.net
            l0 = net.net.encoder_linear_0
            heads = [net.f_mean, net.f_sigma] + ([net.f_cov] if isinstance(net, FullCovarianceNN) else [])
            return dict(in_dim=l0.in_features, hid=l0.out_features, out_dim=sum(h.out_features for h in heads),
                        w0=[l0.weight], b0=[l0.bias], w1=[h.weight for h in heads], b1=[h.bias for h in heads])

        def dec_unit(l0, l1):
            return dict(in_dim=l0.in_features, hid=l0.out_features, out_dim=l1.out_features,
                        w0=[l0.weight], b0=[l0.bias], w1=[l1.weight], b1=[l1.bias])

        if isinstance(vae.prior_net_c.net, FullCovarianceNN) != isinstance(vae.prior_net_y.net, FullCovarianceNN):
            raise ValueError("the two conditional prior nets must be of the same kind (both FactorizedNN or both FullCovarianceNN)")
        enc_mods = [vae.encoder] + ([vae.encoder_c, vae.encoder_y] if vae.model_type == "P" else [])
        groups = []  # (name, unit dict)
        for name, m in zip(["encoder", "encoder_c", "encoder_y"], enc_mods):
            groups.append((name, enc_unit(m)))
        groups.append(("prior_net_c", enc_unit(vae.prior_net_c)))
        groups.append(("prior_net_y", enc_unit(vae.prior_net_y)))
        groups.append(("decoder_x", dec_unit(vae.decoder_x.fx0, vae.decoder_x.fx1)))
        groups.append(("decoder_c", dec_unit(vae.decoder_c.net[0], vae.decoder_c.net[2])))
        groups.append(("decoder_y", dec_unit(vae.decoder_y.net[0], vae.decoder_y.net[2])))

        slots, ranges, off = [], {}, 0
        for name, u in groups:
            beg = off
            for key in ("w0", "b0", "w1", "b1"):
                u[key + "_off"] = off
                for p in u[key]:
                    slots.append((p, off))
                    off += p.numel()
            ranges[name] = (beg, off)
        slots.append((vae.log_sigma_x, off))
        ranges["log_sigma_x"] = (off, off + 1)
        lsx_off = off
        off += 1
        self.n_params = off
        self.ranges = ranges
        self.slots = slots

        dev = self.dev
        self.params = torch.empty(off, dtype=torch.float32, device=dev)
        # gradients and the 8 loss scalars share one buffer: ONE allreduce covers both (data parallel)
        self.gradbuf = torch.zeros(off + 8, dtype=torch.float32, device=dev)
        self.grads = self.gradbuf[:off]
        self.scalars = self.gradbuf[off:]
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in slots:
                self.params[o:o + p.numel()].copy_(p.detach().reshape(-1).to(dev, torch.float32))
                p.data = self.params[o:o + p.numel()].view(p.shape)
                p.grad = None

        # -- descriptor ------------------------------------------------------------------------------
        d = _lib.ModelDesc()
        d.model_type = _lib.MODEL_P if vae.model_type == "P" else _lib.MODEL_S
        d.nz_x, d.nz_c, d.nz_y = vae.nz_x, vae.nz_c, vae.nz_y
        d.nd_x, d.nd_c, d.nd_y = vae.nd_x, vae.nd_c, vae.nd_y
        d.nd_p = len(vae.idx_c_phys)
        for i, v in enumerate(vae.idx_c_phys):
            d.idx_c_phys[i] = int(v)

        def fill(dst, u):
            dst.in_dim, dst.hid, dst.out_dim = u["in_dim"], u["hid"], u["out_dim"]
            dst.w0, dst.b0, dst.w1, dst.b1 = u["w0_off"], u["b0_off"], u["w1_off"], u["b1_off"]

        gd = dict(groups)
        fill(d.enc[0], gd["encoder"])
        if vae.model_type == "P":
            fill(d.enc[1], gd["encoder_c"])
            fill(d.enc[2], gd["encoder_y"])
        fill(d.prior[0], gd["prior_net_c"])
        fill(d.prior[1], gd["prior_net_y"])
        fill(d.fx, gd["decoder_x"])
        fill(d.dec_c, gd["decoder_c"])
        fill(d.dec_y, gd["decoder_y"])
        d.log_sigma_x = lsx_off
        d.n_params = off

        def put(dst, t, n):
            v = t.detach().reshape(-1).cpu().float().tolist()
            if len(v) != n:
                raise ValueError("scaler statistics do not match the data dimension")
            for i, val in enumerate(v):
                dst[i] = val

        for nm, tr, n in (("x", vae.transform_x, vae.nd_x), ("c", vae.transform_c, vae.nd_c), ("y", vae.transform_y, vae.nd_y)):
            if tr is None:
                put(getattr(d, "mean_" + nm), torch.zeros(n), n)
                put(getattr(d, "std_" + nm), torch.ones(n), n)
            else:
                put(getattr(d, "mean_" + nm), tr.mean_, n)
                put(getattr(d, "std_" + nm), tr.scale_, n)

        # bijector Logistic(k=1) -> ShiftScale(lb, ub) on the physics latents (dpivae.py:184-187,237-238)
        ot = vae.encoder.output_transform
        if not isinstance(ot, (ChainTransform, ChainTransformMasked)) or len(ot.lst_transforms) != 2 \
                or not isinstance(ot.lst_transforms[0], Logistic) or not isinstance(ot.lst_transforms[1], ShiftScale) \
                or float(ot.lst_transforms[0].k) != 1.0:
            raise ValueError("encoder output transform must be Logistic(k=1) -> ShiftScale(lb, ub)")
        if isinstance(ot, ChainTransformMasked) and list(ot.mask) != list(range(vae.nz_x)):
            raise ValueError("S-model bijector mask must be the leading nz_x latent columns")
        lb = ot.lst_transforms[1].lb.detach().cpu().float().tolist()
        ub = ot.lst_transforms[1].ub.detach().cpu().float().tolist()
        for i in range(vae.nz_x):
            d.lb[i], d.ub[i] = lb[i], ub[i]
        for i, dist_i in enumerate(vae.prior_x.distributions):
            if isinstance(dist_i, torch.distributions.Uniform):
                d.prior_kind[i], d.prior_a[i], d.prior_b[i] = _lib.PRIOR_UNIFORM, float(dist_i.low), float(dist_i.high)
            elif isinstance(dist_i, torch.distributions.Normal):
                d.prior_kind[i], d.prior_a[i], d.prior_b[i] = _lib.PRIOR_NORMAL, float(dist_i.loc), float(dist_i.scale)
            else:
                raise ValueError(f"unsupported prior over zx: {type(dist_i).__name__}")
        gr = vae.decoder_x.grad_reverse
        d.lambda_g0 = float(gr._alpha) if gr is not None else -1.0  # no GRL == plain gradient (x +1)
        d.has_lambda_x = 0 if vae.lambda_x is None else 1
        d.lambda_x = 0.0 if vae.lambda_x is None else float(vae.lambda_x)

        model = vae.decoder_x.model
        kind = getattr(model, "physics_kind", None)
        frozen = None
        if kind == "mlp":
            lin = model.linear_layers()
            d.phys_kind, d.phys_n_layers = _lib.PHYS_MLP, len(lin)
            d.phys_dims[0] = lin[0].in_features
            for i, l in enumerate(lin):
                d.phys_dims[i + 1] = l.out_features
            frozen = (
                np.concatenate([l.weight.detach().cpu().float().numpy().reshape(-1) for l in lin]),
                np.concatenate([l.bias.detach().cpu().float().numpy().reshape(-1) for l in lin]),
                model.input_transform.mean_.detach().cpu().float().numpy().reshape(-1).copy(),
                model.input_transform.scale_.detach().cpu().float().numpy().reshape(-1).copy(),
            )
        elif kind in ("mass_spring", "beam"):
            d.phys_kind = _lib.PHYS_MASS_SPRING if kind == "mass_spring" else _lib.PHYS_BEAM
            grid = model.t.detach().cpu().float().tolist()
            if len(grid) != vae.nd_x:
                raise ValueError("physics grid length must equal nd_x")
            for i, v in enumerate(grid):
                d.phys_grid[i] = v
        else:
            raise ValueError("part_model must carry physics_kind in {'mlp', 'mass_spring', 'beam'}: arbitrary Python "
                             "callables cannot run inside the fused decoder kernel")
        self.desc = d
        with torch.cuda.device(dev):
            _lib.check(self.lib.dpivae_create(C.byref(d), C.byref(self.handle)))
            if frozen is not None:
                arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in frozen]
                _lib.check(self.lib.dpivae_set_physics_mlp(self.handle, *[a.ctypes.data_as(C.c_void_p) for a in arrs]))
            _lib.check(self.lib.dpivae_bind(self.handle, _ptr(self.params), _ptr(self.grads), _ptr(self.exp_avg),
                                            _ptr(self.exp_avg_sq)))
        props = torch.cuda.get_device_properties(dev)
        self.sm_count = props.multi_processor_count
        self.max_threads_per_sm = props.max_threads_per_multi_processor
        self.launches = 0

    def close(self):
        if self.handle:
            self.lib.dpivae_destroy(self.handle)
            self.handle = C.c_void_p(None)

    # -- optimizer groups ----------------------------------------------------------------------------
    def set_groups(self, groups):
        """groups: list of (range name, lr, weight_decay) covering every range (dpivae.py:335-363)."""
        n = len(groups)
        beg = (C.c_int64 * n)(*[self.ranges[g[0]][0] for g in groups])
        end = (C.c_int64 * n)(*[self.ranges[g[0]][1] for g in groups])
        lr = (C.c_float * n)(*[float(g[1]) for g in groups])
        wd = (C.c_float * n)(*[float(g[2]) for g in groups])
        _lib.check(self.lib.dpivae_set_groups(self.handle, n, beg, end, lr, wd))

    # -- calls -----------------------------------------------------------------------------------------
    def _workspace(self, B, n):
        need = self.lib.dpivae_workspace_bytes(self.handle, B, n)
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.dev)
        return self.workspace

    def _rng(self, B_global, n, cond, eps):
        r = _lib.Rng()
        if eps is not None:
            r.mode = 0
            eps = list(eps) if isinstance(eps, (tuple, list)) else [eps]
            self._keep = [None if e is None else e.to(self.dev, torch.float32).contiguous() for e in eps]
            for k, e in enumerate(self._keep):
                if e is not None:
                    r.eps[k] = e.data_ptr()
            return r
        r.mode = 1
        gen = torch.cuda.default_generators[self.dev.index if self.dev.index is not None else torch.cuda.current_device()]
        r.seed = gen.initial_seed()
        new_off = self.lib.dpivae_philox_plan(self.handle, B_global, n, int(bool(cond)), gen.get_offset(), self.sm_count,
                                              self.max_threads_per_sm, C.byref(r))
        gen.set_offset(new_off)
        return r

    def _batch(self, x, c, y, n, cond=False, idx=None, B_global=None, row_offset=0, row_stride=1):
        def prep(t):
            return None if t is None else t.to(self.dev, torch.float32).contiguous()

        x, c, y = prep(x), prep(c), prep(y)
        if idx is not None:
            idx = idx.to(self.dev, torch.int64).contiguous()
        B = int(idx.shape[0]) if idx is not None else int(x.shape[0])
        b = _lib.Batch()
        b.x, b.c, b.y, b.idx = _ptr(x), _ptr(c), _ptr(y), _ptr(idx)
        b.B, b.B_global, b.row_offset = B, int(B_global if B_global is not None else B), int(row_offset)
        b.n_mc, b.cond = int(n), int(bool(cond))
        b.row_stride = int(row_stride)   # global row of local row r = row_offset + r * row_stride (cyclic shards: rank, world)
        return b, (x, c, y, idx), B

    def loss(self, x, c, y, n, weights, with_grad, outputs=None, eps=None, idx=None, B_global=None, row_offset=0,
             adam_step=None, max_grad_norm=0.0, row_stride=1):
        """One dpivae_loss / dpivae_train_step call.  Returns (row_loss (6,B), scalars (8,))."""
        with torch.cuda.device(self.dev):
            b, keep, B = self._batch(x, c, y, n, False, idx, B_global, row_offset, row_stride)
            rng = self._rng(b.B_global, n, False, eps)
            w = _lib.LossWeights(*[float(v) for v in weights])
            out = _lib.Outputs()
            row_loss = torch.empty((6, B), dtype=torch.float32, device=self.dev)
            scalars = self.scalars if with_grad or adam_step is not None else torch.empty(8, dtype=torch.float32, device=self.dev)
            out.row_loss, out.scalars = row_loss.data_ptr(), scalars.data_ptr()
            for k, t in (outputs or {}).items():
                setattr(out, k, t.data_ptr())
            ws = self._workspace(B, n)
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            if adam_step is not None:
                _lib.check(self.lib.dpivae_train_step(self.handle, C.byref(b), C.byref(rng), C.byref(w), int(adam_step),
                                                      float(max_grad_norm), C.byref(out), _ptr(ws), ws.numel(), stream))
            else:
                _lib.check(self.lib.dpivae_loss(self.handle, C.byref(b), C.byref(rng), C.byref(w), int(with_grad),
                                                C.byref(out), _ptr(ws), ws.numel(), stream))
            self.launches += self.lib.dpivae_last_launch_count(self.handle)
            del keep
            return row_loss, (scalars.clone() if scalars is self.scalars else scalars)

    def forward(self, x, c, n, cond, eps=None):
        with torch.cuda.device(self.dev):
            b, keep, B = self._batch(x, c, None, n, cond)
            rng = self._rng(B, n, cond, eps)
            v = self.vae
            shp = {"xh_p": v.nd_x, "xh_d": v.nd_x, "ch": v.nd_c, "log_sigma_c": v.nd_c, "yh": v.nd_y, "log_sigma_y": v.nd_y,
                   "zx": v.nz_x, "zc": v.nz_c, "zy": v.nz_y}
            outs = {k: torch.empty((n, B, d), dtype=torch.float32, device=self.dev) for k, d in shp.items()}
            outs["dens_z"] = torch.empty((n, B), dtype=torch.float32, device=self.dev)
            out = _lib.Outputs()
            for k, t in outs.items():
                setattr(out, k, t.data_ptr())
            ws = self._workspace(B, n)
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(self.lib.dpivae_loss(self.handle, C.byref(b), C.byref(rng), None, 0, C.byref(out), _ptr(ws),
                                            ws.numel(), stream))
            self.launches += self.lib.dpivae_last_launch_count(self.handle)
            del keep
            return outs

    def encode(self, x, n, standardised, eps=None):
        with torch.cuda.device(self.dev):
            b, keep, B = self._batch(x, None, None, n)
            rng = self._rng(B, n, False, eps)
            v = self.vae
            zx = torch.empty((n, B, v.nz_x), dtype=torch.float32, device=self.dev)
            zc = torch.empty((n, B, v.nz_c), dtype=torch.float32, device=self.dev)
            zy = torch.empty((n, B, v.nz_y), dtype=torch.float32, device=self.dev)
            dens = torch.empty((n, B), dtype=torch.float32, device=self.dev)
            ws = self._workspace(B, n)
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(self.lib.dpivae_encode(self.handle, C.byref(b), C.byref(rng), int(bool(standardised)), _ptr(zx),
                                              _ptr(zc), _ptr(zy), _ptr(dens), _ptr(ws), ws.numel(), stream))
            self.launches += self.lib.dpivae_last_launch_count(self.handle)
            del keep
            return zx, zc, zy, dens

    def decode(self, zx_in, zc, zy):
        with torch.cuda.device(self.dev):
            v = self.vae
            squeeze = zx_in.dim() == 2
            zs = [t.to(self.dev, torch.float32).contiguous() for t in (zx_in, zc, zy)]
            if squeeze:
                zs = [t.unsqueeze(0) for t in zs]
            n, B = int(zs[0].shape[0]), int(zs[0].shape[1])
            want = (v.nz_x + len(v.idx_c_phys), v.nz_c, v.nz_y)
            for t, d in zip(zs, want):
                if tuple(t.shape) != (n, B, d):
                    raise ValueError(f"decode: expected latents of shape ({n}, {B}, {d}), got {tuple(t.shape)}")
            shp = {"xh_p": v.nd_x, "xh_d": v.nd_x, "ch": v.nd_c, "log_sigma_c": v.nd_c, "yh": v.nd_y, "log_sigma_y": v.nd_y}
            outs = {k: torch.empty((n, B, d), dtype=torch.float32, device=self.dev) for k, d in shp.items()}
            out = _lib.Outputs()
            for k, t in outs.items():
                setattr(out, k, t.data_ptr())
            ws = self._workspace(B, n)
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(self.lib.dpivae_decode(self.handle, _ptr(zs[0]), _ptr(zs[1]), _ptr(zs[2]), B, n, C.byref(out), _ptr(ws),
                                              ws.numel(), stream))
            self.launches += self.lib.dpivae_last_launch_count(self.handle)
            res = tuple(outs[k] for k in ("xh_p", "xh_d", "ch", "log_sigma_c", "yh", "log_sigma_y"))
            return tuple(t.squeeze(0) for t in res) if squeeze else res

    def prior_net(self, c, y=None):
        with torch.cuda.device(self.dev):
            v = self.vae
            c = c.to(self.dev, torch.float32).contiguous()
            y = None if y is None else y.to(self.dev, torch.float32).contiguous()
            B = int(c.shape[0])
            loc_c = torch.empty((B, v.nz_c), dtype=torch.float32, device=self.dev)
            tril_c = torch.empty((B, v.nz_c, v.nz_c), dtype=torch.float32, device=self.dev)
            loc_y = tril_y = None
            if y is not None:
                loc_y = torch.empty((B, v.nz_y), dtype=torch.float32, device=self.dev)
                tril_y = torch.empty((B, v.nz_y, v.nz_y), dtype=torch.float32, device=self.dev)
            ws = self._workspace(B, 1)
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(self.lib.dpivae_prior_net(self.handle, _ptr(c), _ptr(y), B, _ptr(loc_c), _ptr(tril_c), _ptr(loc_y),
                                                 _ptr(tril_y), _ptr(ws), ws.numel(), stream))
            self.launches += self.lib.dpivae_last_launch_count(self.handle)
            return loc_c, tril_c, loc_y, tril_y

    def gaussian_sample(self, loc, tril, eps):
        with torch.cuda.device(self.dev):
            n, B, nz = (int(d) for d in eps.shape)
            z = torch.empty((n, B, nz), dtype=torch.float32, device=self.dev)
            dens = torch.empty((n, B), dtype=torch.float32, device=self.dev)
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(self.lib.dpivae_gaussian_sample(_ptr(loc.contiguous()), _ptr(tril.contiguous()), _ptr(eps.contiguous()), n, B, nz,
                                                       _ptr(z), _ptr(dens), stream))
            self.launches += 1
            return z, dens

    def set_math_mode(self, mode):
        """'fp32' (FFMA, default) | 'tc_fp16x3' (tcgen05, fp16 hi/lo split, fp32-accurate) | 'tc_fp16' (tcgen05, plain
        fp16 operands) -- include/dpivae_b200.h DPIVAE_MATH_*."""
        code = {"fp32": _lib.MATH_FP32, "tc_fp16x3": _lib.MATH_TC_FP16X3, "tc_fp16": _lib.MATH_TC_FP16}[mode]
        _lib.check(self.lib.dpivae_set_math_mode(self.handle, code))
        self.math_mode = mode

    def used_tensor_cores(self):
        return bool(self.lib.dpivae_last_used_tensor_cores(self.handle))

    def set_timing(self, enable):
        _lib.check(self.lib.dpivae_set_timing(self.handle, int(bool(enable))))

    def last_kernel_ms(self):
        out = (C.c_float * 7)()
        _lib.check(self.lib.dpivae_last_kernel_ms(self.handle, out))
        return dict(zip(("enc_fwd", "dec_fused", "enc_bwd", "reduce", "adam", "lat_fwd", "lat_bwd"), [float(v) for v in out]))

    def ffma_peak_tflops(self):
        v = C.c_float(0.0)
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.dpivae_ffma_peak_tflops(C.byref(v), C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)))
        return float(v.value)

    def step_graph(self, x, c, y, n, weights, idx_pool=None, max_grad_norm=0.0, log_cap=1024, unroll=1):
        return StepGraph(self, x, c, y, n, weights, idx_pool, max_grad_norm, log_cap, unroll)

    def adam_step(self, step, max_grad_norm=0.0):
        with torch.cuda.device(self.dev):
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(self.lib.dpivae_adam_step(self.handle, int(step), float(max_grad_norm), stream))
            self.launches += self.lib.dpivae_last_launch_count(self.handle)


class StepGraph:
    """Device-resident training loop (include/dpivae_b200.h dpivae_step_graph_*): one captured training step --
    minibatch gather from a resident index pool, loss, backward, clip, Adam -- whose step counter, Adam bias
    corrections and generator offset live on the device.  `run(k)` enqueues k replays without any host
    synchronisation; `log` is the (log_cap, 9) device ring of per-step [8 loss scalars | log_sigma_x]."""

    def __init__(self, eng, x, c, y, n, weights, idx_pool=None, max_grad_norm=0.0, log_cap=1024, unroll=1):
        self.eng, self.n = eng, int(n)
        with torch.cuda.device(eng.dev):
            self.pool = None if idx_pool is None else idx_pool.to(eng.dev, torch.int64).contiguous()
            B = int(self.pool.shape[1]) if self.pool is not None else int(x.shape[0])
            self.idx_cur = torch.zeros(B, dtype=torch.int64, device=eng.dev) if self.pool is not None else None
            b, self._keep, _ = eng._batch(x, c, y, n, False, self.idx_cur, None, 0)
            self.B = B
            self.log = torch.zeros((int(log_cap), 9), dtype=torch.float32, device=eng.dev)
            # private workspace: the graph holds its address for as long as it lives
            self.ws = torch.empty(eng.lib.dpivae_workspace_bytes(eng.handle, B, n), dtype=torch.uint8, device=eng.dev)
            # captured on a private stream (the legacy default stream cannot be captured); replays go to the caller's stream
            self._cap_stream = torch.cuda.Stream(eng.dev)
            self._cap_stream.wait_stream(torch.cuda.current_stream(eng.dev))
            rng, self.inc = self._plan()
            w = _lib.LossWeights(*[float(v) for v in weights])
            self.handle = C.c_void_p(None)
            st = C.c_void_p(self._cap_stream.cuda_stream)
            _lib.check(eng.lib.dpivae_step_graph_create(
                eng.handle, C.byref(b), C.byref(rng), self.inc, C.byref(w), eng.step_count + 1, float(max_grad_norm),
                _ptr(self.pool), 0 if self.pool is None else int(self.pool.shape[0]), _ptr(self.idx_cur), _ptr(eng.scalars),
                _ptr(self.log), int(log_cap), int(unroll), _ptr(self.ws), self.ws.numel(), st, C.byref(self.handle)))
            torch.cuda.current_stream(eng.dev).wait_stream(self._cap_stream)

    def _gen(self):
        dev = self.eng.dev
        return torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]

    def _plan(self):
        """Philox plan of the next step at the torch CUDA generator's current offset (generator left untouched)."""
        eng, gen = self.eng, self._gen()
        r = _lib.Rng()
        r.mode, r.seed = 1, gen.initial_seed()
        off = gen.get_offset()
        new_off = eng.lib.dpivae_philox_plan(eng.handle, self.B, self.n, 0, off, eng.sm_count, eng.max_threads_per_sm, C.byref(r))
        return r, int(new_off - off)

    def set_pool(self, idx_pool):
        """New minibatch rows for the coming steps (same shape; step t reads row (t - 1) % pool_rows).  Host tensors
        go through a pinned staging buffer with an asynchronous copy (no host wait for the device)."""
        if idx_pool.is_cuda:
            self.pool.copy_(idx_pool.to(torch.int64), non_blocking=True)
            return
        if getattr(self, "_pin", None) is None:
            self._pin = torch.empty(tuple(self.pool.shape), dtype=torch.int64, pin_memory=True)
            self._pin_free = torch.cuda.Event()
            self._pin_used = False
        if self._pin_used:
            self._pin_free.synchronize()   # the previous asynchronous copy has read the staging buffer
        self._pin.copy_(idx_pool)
        with torch.cuda.device(self.eng.dev):
            self.pool.copy_(self._pin, non_blocking=True)
            self._pin_free.record(torch.cuda.current_stream(self.eng.dev))
        self._pin_used = True

    def run(self, n_steps):
        """Enqueue n_steps training steps; keeps eng.step_count and the torch CUDA generator in step with the device."""
        eng = self.eng
        with torch.cuda.device(eng.dev):
            st = C.c_void_p(torch.cuda.current_stream(eng.dev).cuda_stream)
            rng, _ = self._plan()
            _lib.check(eng.lib.dpivae_step_graph_reset(self.handle, C.byref(rng), eng.step_count + 1, st))
            _lib.check(eng.lib.dpivae_step_graph_launch(self.handle, int(n_steps), st))
            gen = self._gen()
            gen.set_offset(gen.get_offset() + self.inc * int(n_steps))
            eng.launches += eng.lib.dpivae_last_launch_count(eng.handle)
            eng.step_count += int(n_steps)

    def log_rows_device(self, first_step, n_steps):
        """Device copy (stream-ordered, no host synchronisation) of the log rows of optimizer steps
        first_step .. first_step + n_steps - 1 (1-based); must be taken before the ring wraps around."""
        cap = self.log.shape[0]
        a = (first_step - 1) % cap
        if a + n_steps <= cap:
            return self.log[a:a + n_steps].clone()
        return torch.cat([self.log[a:], self.log[:a + n_steps - cap]], dim=0)

    def log_rows(self, first_step, n_steps):
        """Host copy of the log rows of optimizer steps first_step .. first_step + n_steps - 1 (1-based)."""
        return self.log_rows_device(first_step, n_steps).cpu()

    def close(self):
        if self.handle:
            self.eng.lib.dpivae_step_graph_destroy(self.handle)
            self.handle = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _LossFn(torch.autograd.Function):
    """Keeps `vae.loss(...)[0].sum().backward()` (the reference loop, dpivae.py:407-429) working:
    the fused kernel already produced d(sum_b loss_b / (B*D)) / d(params); a uniform upstream
    gradient rescales it, anything else cannot be expressed by the fused backward and raises."""

    @staticmethod
    def forward(ctx, engine, row_loss, scale, *params):
        ctx.engine, ctx.scale = engine, scale
        ctx.flat = engine.grads.clone()
        return tuple(row_loss[i] for i in range(6))

    @staticmethod
    def backward(ctx, g_loss, *g_rest):
        for g in g_rest:
            if g is not None and bool((g != 0).any()):
                raise RuntimeError("only the total loss (first element of DPIVAE.loss) is differentiable in the fused path")
        if g_loss is None:
            return (None, None, None) + tuple(None for _ in ctx.engine.slots)
        g0 = g_loss.reshape(-1)[0]
        if not bool(torch.all(g_loss == g0)):
            raise RuntimeError("the fused backward supports a uniform upstream gradient (loss.sum() / const) only")
        flat = ctx.flat * (g0 * ctx.scale)
        outs = [flat[o:o + p.numel()].view(p.shape) for p, o in ctx.engine.slots]
        return (None, None, None) + tuple(outs)


class DPIVAE(nn.Module):
    """models/vae.py:9-255."""

    def __init__(self, prior_x, prior_net_c, prior_net_y, encoder, decoder_x, decoder_c, decoder_y, nz_x, nz_c, nz_y,
                 nd_x, nd_c, nd_y, idx_c_phys, model_type=None, encoder_c=None, encoder_y=None, lambda_x=None,
                 transform_x=None, transform_c=None, transform_y=None, jitter=1e-6):
        super().__init__()
        self.prior_x = prior_x
        self.prior_net_c = prior_net_c
        self.prior_net_y = prior_net_y
        self.encoder = encoder
        self.model_type = model_type
        self.encoder_c = encoder_c
        self.encoder_y = encoder_y
        self.decoder_x = decoder_x
        self.decoder_c = decoder_c
        self.decoder_y = decoder_y
        self.nz_x, self.nz_c, self.nz_y = nz_x, nz_c, nz_y
        self.nd_x, self.nd_c, self.nd_y = nd_x, nd_c, nd_y
        self.idx_c_phys = idx_c_phys
        self.lambda_x = lambda_x
        self.transform_x, self.transform_c, self.transform_y = transform_x, transform_c, transform_y
        self.jitter = jitter
        if (self.model_type != "P") and (self.model_type != "S"):
            raise ValueError(f"Invalid model_type {self.model_type}")
        if self.model_type == "S" and ((self.encoder_c is not None) or (self.encoder_y is not None)):
            raise ValueError("encoder_c and encoder_y must NOT be defined for model type S")
        if self.model_type == "P" and ((self.encoder_c is None) or (self.encoder_y is None)):
            raise ValueError("encoder_c and encoder_y must be defined for model type P")
        self.log_sigma_x = nn.Parameter(torch.tensor(0.0), requires_grad=True)
        self._engine = None
        self._eps = None

    # -- engine plumbing ---------------------------------------------------------------------------
    def engine(self, dev=None):
        if self._engine is None:
            if dev is None:
                dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
            if dev is None:
                raise RuntimeError("DPIVAE needs a CUDA device: the step runs in sm_100a kernels, there is no CPU path")
            self._engine = _Engine(self, dev)
        return self._engine

    def _apply(self, fn, *a, **k):
        # .to()/.cuda() would replace the flat-buffer views; re-flatten lazily afterwards
        if self._engine is not None:
            self._engine.close()
            self._engine = None
        return super()._apply(fn, *a, **k)

    def compile(self, *a, **k):  # the reference's vae.compile() is inert
        return self

    @contextmanager
    def inject_noise(self, eps):
        """Use the given reparameterisation noise instead of the in-kernel Philox stream.
        P: (eps_x, eps_c, eps_y[, eps_cond]) each (n, B, nz_k); S: tensor (n, B, Z)."""
        self._eps = eps
        try:
            yield self
        finally:
            self._eps = None

    # -- reference API -----------------------------------------------------------------------------------
    def transform_inputs(self, x=None, c=None, y=None):
        out = []
        for v, tr in ((x, self.transform_x), (c, self.transform_c), (y, self.transform_y)):
            if v is None:
                out.append(torch.nan)
            else:
                out.append(tr.forward(v)[0] if tr is not None else v)
        return tuple(out)

    def encode(self, x, n=1):
        """models/vae.py:125-151 -- `x` is the already standardised input."""
        return self.engine().encode(x, n, True, self._eps)

    def forward(self, x, c, cond=False, n=1):
        o = self.engine().forward(x, c, n, cond, self._eps)
        return (o["xh_p"], o["xh_d"], o["ch"], o["log_sigma_c"], o["yh"], o["log_sigma_y"], o["zx"], o["zc"], o["zy"],
                o["dens_z"])

    def loss(self, x, c, y, n=1, beta_x=1.0, beta_c=1.0, beta_y=1.0, alpha_x=1.0, alpha_c=1.0, alpha_y=1.0):
        """models/vae.py:177-231 -> (loss (B,), KL_x (B,), 0, 0, R_x, R_c, R_y, reg)."""
        eng = self.engine()
        with_grad = torch.is_grad_enabled() and any(p.requires_grad for p, _ in eng.slots)
        w = (float(beta_x), float(alpha_x), float(alpha_c), float(alpha_y))
        row_loss, _ = eng.loss(x, c, y, n, w, with_grad, eps=self._eps)
        zero = torch.tensor(0.0)
        if with_grad:
            B = row_loss.shape[1]
            scale = float(B * (self.nd_x + self.nd_c + self.nd_y))
            rl = _LossFn.apply(eng, row_loss, scale, *[p for p, _ in eng.slots])
        else:
            rl = tuple(row_loss[i] for i in range(6))
        return rl[0], rl[1], zero, zero.clone(), rl[2], rl[3], rl[4], rl[5]

    def sample(self, x, c, cond=False, n=1):
        """models/vae.py:233-255: forward + three Gaussian noise draws (taken from torch's CUDA
        generator right after the in-kernel Philox draws, i.e. at the reference's stream positions)."""
        xh_p, xh_d, ch, lsc, yh, lsy, zx, zc, zy, dens_z = self.forward(x, c, cond=cond, n=n)
        with torch.no_grad():
            sx = self.log_sigma_x.detach().exp()
            x_sample = torch.normal(xh_p + xh_d, sx.expand_as(xh_p))
            c_sample = torch.normal(ch, lsc.exp())
            y_sample = torch.normal(yh, lsy.exp())
        return x_sample, xh_p, xh_d, c_sample, y_sample, zx, zc, zy, dens_z

    def decode(self, zx, zc, zy):
        """models/vae.py:153-158: `zx` is the physics-decoder input [zx | c_phys]; (n, B, .) or (B, .) latents."""
        return self.engine().decode(zx, zc, zy)

    def prior_net(self, c, y=None):
        """models/vae.py:99-110 -> (loc_c, scale_tril_c, loc_y | None, scale_tril_y | None)."""
        return self.engine().prior_net(c, y)

    def sample_prior(self, c, y, n=1):
        """models/vae.py:112-123: draws (n, B, nz_c) then (n, B, nz_y) from torch's CUDA generator (the reference's
        order), z = loc + sigma * eps and its log-density in the fused sampling kernel."""
        eng = self.engine()
        loc_c, tril_c, loc_y, tril_y = eng.prior_net(c, y)
        with torch.cuda.device(eng.dev):
            B = loc_c.shape[0]
            eps_c = torch.empty((n, B, self.nz_c), dtype=torch.float32, device=eng.dev).normal_()
            eps_y = torch.empty((n, B, self.nz_y), dtype=torch.float32, device=eng.dev).normal_()
        zc, dens_zc = eng.gaussian_sample(loc_c, tril_c, eps_c)
        zy, dens_zy = eng.gaussian_sample(loc_y, tril_y, eps_y)
        return zc, dens_zc, zy, dens_zy
