"""The C-ABI library loads without a GPU and exports every symbol include/dpivae_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dpivae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpivae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dpivae_b200 import _lib, build

    lib_path = build.build_library()
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    # the ctypes binding lists the same set
    assert sorted(_lib.EXPORTS) == syms
    assert lib.dpivae_abi_version() == 3


def test_struct_sizes_match_header():
    """ctypes mirrors of the ABI structs have the C layout (sizes computed from the header's field list)."""
    from dpivae_b200 import _lib

    assert ctypes.sizeof(_lib.Mlp2) == 48
    assert ctypes.sizeof(_lib.Batch) == 4 * 8 + 3 * 8 + 2 * 4 + 8   # ... + row_stride (ABI 3)
    assert ctypes.sizeof(_lib.Rng) == 8 + 4 * 8 + 8 + 4 * 8 + 4 * 4
    assert ctypes.sizeof(_lib.LossWeights) == 16
    assert ctypes.sizeof(_lib.Outputs) == 12 * 8
    # ModelDesc: 8 int32 + 4 int32 + 8 mlp2 + 2 int64 + floats/ints as declared
    n = 8 * 4 + 4 * 4 + 8 * 48 + 16 + (64 * 2 + 4 * 4) * 4 + (4 + 4) * 4 + 4 * 4 + 8 * 4 + 12 + 8 + 7 * 4 + 64 * 4
    assert ctypes.sizeof(_lib.ModelDesc) == (n + 7) // 8 * 8
    # dpivae_datagen_desc_t: 6 int32 + dims[8] + 4 x float[16] + idx_c[4] + idx_y[4] + 4 floats
    assert ctypes.sizeof(_lib.DataGenDesc) == 6 * 4 + 8 * 4 + 4 * 16 * 4 + 4 * 4 + 4 * 4 + 4 * 4


def test_no_cpu_fallback():
    """Creating a model without CUDA must fail loudly (no eager / CPU path exists)."""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from helpers import build_from_golden

    g, spec, sd, args, case_mod, vae, (x, c, y) = build_from_golden("simple_beam", "S", device="cpu")
    with pytest.raises(RuntimeError):
        vae.loss(x, c, y, n=2)
    with pytest.raises(RuntimeError):
        vae.encoder(x)  # sub-module forwards are fused-only
