"""Import the UNMODIFIED reference (JanKoune/DPI-VAE) in this container.

Container-only tooling (reads /root/reference, which does not exist on the GPU box).
Used by tests/golden/make_golden.py and tools/export_case_assets.py.  Nothing under
dpivae_b200/, bench.py or the gpu tests imports this file.

Recipe; none of the shims touches hot-path arithmetic:
  1. working copy of /root/reference under /tmp (the reference tree is read-only and its
     case modules use cwd-relative paths such as ./cases/bridge/),
  2. stub packages for the missing third-party imports (pytorch_lightning, torchrl,
     matplotlib, seaborn) that only provide base classes / loggers / plotting,
  3. torch.load(map_location="cpu") (surrogate checkpoints were pickled from CUDA tensors),
  4. zero placeholder y*.pt blobs (simulator outputs nothing on the hot path reads).
"""
import os
import shutil
import sys
import types

REF_SRC = "/root/reference"
WORK = "/tmp/dpivae_ref_work"
# git-ignored in-repo copy made by tools/install_reference.py: travels to the GPU box with the snapshot, so that
# `bench.py --impl reference` can time the reference's OWN code there (nothing else reads it)
INSTALLED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "baseline", "_ref")
PLACEHOLDERS = [
    ("cases/damped_oscillator/y.pt", (20000, 200)),
    ("cases/simple_beam/y.pt", (20000, 200)),
    ("cases/bridge/y.pt", (5000, 200)),
    ("cases/bridge/y_partial.pt", (5000, 200)),
]


def copy_reference(dst):
    """Unmodified copy of the reference tree + zero placeholders for the four missing simulator blobs."""
    import torch

    if not os.path.exists(dst):
        shutil.copytree(REF_SRC, dst, ignore=shutil.ignore_patterns("output", "figures", ".git"))
        os.system(f"chmod -R u+w {dst}")
    for rel, shape in PLACEHOLDERS:
        p = os.path.join(dst, rel)
        if not os.path.exists(p):
            # stride-0 view: the file holds one row, the loaded tensor has the full shape (only the shape is read)
            torch.save(torch.zeros((1, shape[1])).expand(*shape), p)
    return dst


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, k):
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())


def _module_getattr(k):
    if k.startswith("__"):
        raise AttributeError(k)
    return _Anything()


class _ScalarStore(dict):
    def __missing__(self, k):
        self[k] = []
        return self[k]


class _CSVLogger:
    """Surface of torchrl.record.CSVLogger used by dpivae.py:377,439-451,505-519."""

    def __init__(self, exp_name="", log_dir=None, **k):
        self.experiment = types.SimpleNamespace(scalars=_ScalarStore())

    def log_scalar(self, name, value, step=None):
        self.experiment.scalars[name].append((step, float(value)))


def available():
    return os.path.isdir(REF_SRC) or os.path.isdir(INSTALLED)


def prepare(root=None):
    """root: None = the installed copy when /root/reference is absent (GPU box), else a /tmp working copy."""
    import torch

    global WORK
    if root is None:
        root = WORK if os.path.isdir(REF_SRC) else os.path.abspath(INSTALLED)
    if root == WORK:
        copy_reference(WORK)
    elif not os.path.isdir(root):
        raise RuntimeError(f"reference copy not found at {root} (run tools/install_reference.py in the build container)")
    WORK = root

    # stubs
    import torch.nn as nn

    pl = _stub("pytorch_lightning", LightningModule=nn.Module)
    plu = _stub("pytorch_lightning.utilities")
    plm = _stub("pytorch_lightning.utilities.model_summary", ModelSummary=lambda *a, **k: "<summary>")
    pl.utilities = plu
    plu.model_summary = plm
    tr = _stub("torchrl")
    trr = _stub("torchrl.record", CSVLogger=_CSVLogger)
    tr.record = trr
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm",
                 "matplotlib.lines", "matplotlib.patches", "matplotlib.ticker", "matplotlib.gridspec", "seaborn"]:
        m = _stub(name)
        m.__getattr__ = _module_getattr  # type: ignore
        m.rcParams = {}
        m.colormaps = {}
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    if not getattr(torch.load, "_dpivae_wrapped", False):
        _orig = torch.load

        def _load(*a, **k):
            k.setdefault("map_location", "cpu")
            return _orig(*a, **k)

        _load._dpivae_wrapped = True
        torch.load = _load

    os.chdir(WORK)
    if WORK not in sys.path:
        sys.path.insert(0, WORK)
    sys.argv = sys.argv[:1]


def load(case_name):
    """Return (dpivae module, case module) of the reference."""
    prepare()
    import importlib

    dpivae = importlib.import_module("dpivae")
    case = importlib.import_module(f"cases.{case_name}")
    return dpivae, case


def make_args(case, preset, **over):
    from utils import make_parser

    args, _ = make_parser().parse_known_args([])
    for k, v in case.presets[preset].items():
        setattr(args, k, v)
    for k, v in over.items():
        setattr(args, k, v)
    return args
