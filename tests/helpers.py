"""Shared helpers of the parity tests: build a dpivae_b200 model from a golden fixture."""
import importlib

import torch

import golden_util as gu

PRESET = {("bridge", "P"): "DPIVAE-A", ("bridge", "S"): "DPIVAE-B", ("damped_oscillator", "P"): "vae",
          ("damped_oscillator", "S"): "dpivae", ("simple_beam", "P"): "vae", ("simple_beam", "S"): "dpivae"}


def make_args(case_mod, preset, **over):
    from dpivae_b200 import make_parser

    args, _ = make_parser().parse_known_args([])
    for k, v in case_mod.presets[preset].items():
        setattr(args, k, v)
    for k, v in over.items():
        setattr(args, k, v)
    return args


def build_from_golden(case, mtype, device="cuda", ext=False, **over):
    """setup_model on the golden minibatch (so the scalers are fitted exactly like the reference's),
    then load the reference's initial weights.  ext=True: the second fixture set (tests/golden/make_golden_ext.py)."""
    import dpivae_b200 as dpv

    g, spec, sd = gu.load(case, mtype, ext=ext)
    case_mod = importlib.import_module(f"dpivae_b200.cases.{case}")
    x, c, y = (torch.from_numpy(g[k].copy()) for k in "xcy")
    B = x.shape[0]
    args = make_args(case_mod, PRESET[(case, mtype)], use_seed=True, seed=123, n_train=B, n_batch=B, **over)
    vae = dpv.setup_model(args, case_mod.definition, (x, c, y))
    missing = vae.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if not k.startswith("decoder_x.model.")], missing
    assert not missing.unexpected_keys, missing
    return g, spec, sd, args, case_mod, vae, (x, c, y)
