"""ctypes binding of libdpivae_b200.so (include/dpivae_b200.h).

There is NO CPU or PyTorch fallback: importing this module without the built library, or
creating a model without a CUDA device, raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DPIVAE_B200_LIB: load another build of the same ABI (A/B timing of kernel variants); never set in tests / bench runs
LIB_PATH = os.environ.get("DPIVAE_B200_LIB") or os.path.join(HERE, "libdpivae_b200.so")

MAX_ZX, MAX_ZCY, MAX_Z, MAX_NDX, MAX_NDCY, MAX_PHYS_LAYERS = 4, 8, 16, 64, 4, 6
MODEL_P, MODEL_S = 0, 1
PHYS_MLP, PHYS_MASS_SPRING, PHYS_BEAM = 0, 1, 2
PRIOR_UNIFORM, PRIOR_NORMAL = 0, 1


class Mlp2(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("hid", C.c_int32), ("out_dim", C.c_int32), ("_pad", C.c_int32),
                ("w0", C.c_int64), ("b0", C.c_int64), ("w1", C.c_int64), ("b1", C.c_int64)]


class ModelDesc(C.Structure):
    _fields_ = [
        ("model_type", C.c_int32),
        ("nz_x", C.c_int32), ("nz_c", C.c_int32), ("nz_y", C.c_int32),
        ("nd_x", C.c_int32), ("nd_c", C.c_int32), ("nd_y", C.c_int32), ("nd_p", C.c_int32),
        ("idx_c_phys", C.c_int32 * MAX_NDCY),
        ("enc", Mlp2 * 3), ("prior", Mlp2 * 2),
        ("fx", Mlp2), ("dec_c", Mlp2), ("dec_y", Mlp2),
        ("log_sigma_x", C.c_int64), ("n_params", C.c_int64),
        ("mean_x", C.c_float * MAX_NDX), ("std_x", C.c_float * MAX_NDX),
        ("mean_c", C.c_float * MAX_NDCY), ("std_c", C.c_float * MAX_NDCY),
        ("mean_y", C.c_float * MAX_NDCY), ("std_y", C.c_float * MAX_NDCY),
        ("lb", C.c_float * MAX_ZX), ("ub", C.c_float * MAX_ZX),
        ("prior_kind", C.c_int32 * MAX_ZX), ("prior_a", C.c_float * MAX_ZX), ("prior_b", C.c_float * MAX_ZX),
        ("lambda_g0", C.c_float), ("has_lambda_x", C.c_int32), ("lambda_x", C.c_float),
        ("phys_kind", C.c_int32), ("phys_n_layers", C.c_int32),
        ("phys_dims", C.c_int32 * (MAX_PHYS_LAYERS + 1)),
        ("phys_grid", C.c_float * MAX_NDX),
    ]


class Batch(C.Structure):
    _fields_ = [("x", C.c_void_p), ("c", C.c_void_p), ("y", C.c_void_p), ("idx", C.c_void_p),
                ("B", C.c_int64), ("B_global", C.c_int64), ("row_offset", C.c_int64),
                ("n_mc", C.c_int32), ("cond", C.c_int32), ("row_stride", C.c_int64)]


class Rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("_pad", C.c_int32), ("eps", C.c_void_p * 4), ("seed", C.c_uint64),
                ("offset", C.c_uint64 * 4), ("grid_threads", C.c_uint32 * 4)]


class LossWeights(C.Structure):
    _fields_ = [("beta_x", C.c_float), ("alpha_x", C.c_float), ("alpha_c", C.c_float), ("alpha_y", C.c_float)]


class Outputs(C.Structure):
    _fields_ = [("row_loss", C.c_void_p), ("scalars", C.c_void_p),
                ("xh_p", C.c_void_p), ("xh_d", C.c_void_p), ("ch", C.c_void_p), ("log_sigma_c", C.c_void_p),
                ("yh", C.c_void_p), ("log_sigma_y", C.c_void_p), ("zx", C.c_void_p), ("zc", C.c_void_p),
                ("zy", C.c_void_p), ("dens_z", C.c_void_p)]


class DataGenDesc(C.Structure):
    _fields_ = [("n_factors", C.c_int32), ("n_layers", C.c_int32), ("nd_x", C.c_int32), ("nd_c", C.c_int32), ("nd_y", C.c_int32),
                ("_pad", C.c_int32), ("dims", C.c_int32 * (MAX_PHYS_LAYERS + 2)),
                ("lo", C.c_float * 16), ("hi", C.c_float * 16), ("in_mean", C.c_float * 16), ("in_std", C.c_float * 16),
                ("idx_c", C.c_int32 * MAX_NDCY), ("idx_y", C.c_int32 * MAX_NDCY),
                ("sigma_x", C.c_float), ("sigma_c", C.c_float), ("sigma_y", C.c_float), ("_pad2", C.c_float)]


EXPORTS = [
    "dpivae_create", "dpivae_destroy", "dpivae_last_error", "dpivae_abi_version", "dpivae_sizeof_model_desc", "dpivae_set_physics_mlp",
    "dpivae_bind", "dpivae_set_groups", "dpivae_workspace_bytes", "dpivae_loss", "dpivae_adam_step",
    "dpivae_train_step", "dpivae_encode", "dpivae_philox_plan", "dpivae_last_launch_count",
    "dpivae_set_timing", "dpivae_last_kernel_ms", "dpivae_ffma_peak_tflops", "dpivae_set_phase_buffer",
    "dpivae_set_math_mode", "dpivae_last_used_tensor_cores",
    "dpivae_decode", "dpivae_prior_net", "dpivae_gaussian_sample",
    "dpivae_mc_mean", "dpivae_regression_metrics", "dpivae_linreg_r2",
    "dpivae_datagen_workspace_bytes", "dpivae_sample_response",
    "dpivae_step_graph_create", "dpivae_step_graph_reset", "dpivae_step_graph_launch", "dpivae_step_graph_destroy",
]
MATH_FP32, MATH_TC_FP16X3, MATH_TC_FP16 = 0, 1, 2
ABI_VERSION = 3   # include/dpivae_b200.h DPIVAE_ABI_VERSION

_lib = None


def load():
    """Load the shared library, (re)building it first when it is missing or older than its sources (build.py's
    staleness check; a box without nvcc uses the shipped binary as it is).  The ABI version the binding was written
    against and the size of the model descriptor are checked against the binary: a stale library fails here."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    if not os.environ.get("DPIVAE_B200_LIB") and (os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) or not os.path.exists(LIB_PATH)):
        try:
            _build.build_library()   # no-op unless a source is newer than the binary
        except Exception:
            if not os.path.exists(LIB_PATH):
                raise
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m dpivae_b200.build` (no fallback path exists)")
    lib = C.CDLL(LIB_PATH)
    lib.dpivae_abi_version.restype = C.c_int
    if lib.dpivae_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.dpivae_abi_version()} != binding version {ABI_VERSION} "
                          "(stale binary: run `python -m dpivae_b200.build --force`)")
    lib.dpivae_sizeof_model_desc.restype = C.c_size_t
    if lib.dpivae_sizeof_model_desc() != C.sizeof(ModelDesc):
        raise ImportError(f"{LIB_PATH}: dpivae_model_desc_t is {lib.dpivae_sizeof_model_desc()} bytes in the binary, "
                          f"{C.sizeof(ModelDesc)} in the binding")
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    lib.dpivae_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
    lib.dpivae_destroy.argtypes = [vp]
    lib.dpivae_last_error.restype = C.c_char_p
    lib.dpivae_set_physics_mlp.argtypes = [vp, vp, vp, vp, vp]
    lib.dpivae_bind.argtypes = [vp, vp, vp, vp, vp]
    lib.dpivae_set_groups.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(f32), C.POINTER(f32)]
    lib.dpivae_workspace_bytes.argtypes = [vp, i64, i32]
    lib.dpivae_workspace_bytes.restype = C.c_size_t
    lib.dpivae_loss.argtypes = [vp, C.POINTER(Batch), C.POINTER(Rng), C.POINTER(LossWeights), i32, C.POINTER(Outputs),
                                vp, C.c_size_t, vp]
    lib.dpivae_adam_step.argtypes = [vp, i64, f32, vp]
    lib.dpivae_train_step.argtypes = [vp, C.POINTER(Batch), C.POINTER(Rng), C.POINTER(LossWeights), i64, f32,
                                      C.POINTER(Outputs), vp, C.c_size_t, vp]
    lib.dpivae_encode.argtypes = [vp, C.POINTER(Batch), C.POINTER(Rng), i32, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.dpivae_philox_plan.argtypes = [vp, i64, i32, i32, u64, i32, i32, C.POINTER(Rng)]
    lib.dpivae_philox_plan.restype = u64
    lib.dpivae_last_launch_count.argtypes = [vp]
    lib.dpivae_set_timing.argtypes = [vp, i32]
    lib.dpivae_set_phase_buffer.argtypes = [vp, vp]
    lib.dpivae_last_kernel_ms.argtypes = [vp, C.POINTER(f32)]
    lib.dpivae_ffma_peak_tflops.argtypes = [C.POINTER(f32), vp]
    lib.dpivae_set_math_mode.argtypes = [vp, i32]
    lib.dpivae_last_used_tensor_cores.argtypes = [vp]
    lib.dpivae_step_graph_create.argtypes = [vp, C.POINTER(Batch), C.POINTER(Rng), u64, C.POINTER(LossWeights), i64, f32,
                                             vp, i64, vp, vp, vp, i64, i32, vp, C.c_size_t, vp, C.POINTER(vp)]
    lib.dpivae_decode.argtypes = [vp, vp, vp, vp, i64, i32, C.POINTER(Outputs), vp, C.c_size_t, vp]
    lib.dpivae_prior_net.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.dpivae_gaussian_sample.argtypes = [vp, vp, vp, i32, i64, i32, vp, vp, vp]
    lib.dpivae_mc_mean.argtypes = [vp, i32, i64, i32, vp, vp]
    lib.dpivae_regression_metrics.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    lib.dpivae_linreg_r2.argtypes = [vp, vp, i64, i64, vp, vp, i64, i64, i32, vp, vp, vp]
    lib.dpivae_datagen_workspace_bytes.argtypes = [C.POINTER(DataGenDesc), i64]
    lib.dpivae_datagen_workspace_bytes.restype = C.c_size_t
    lib.dpivae_sample_response.argtypes = [C.POINTER(DataGenDesc), vp, vp, i64, u64, u64, i32, i32, vp, vp, vp, vp, vp, C.c_size_t, vp,
                                           C.POINTER(u64)]
    lib.dpivae_step_graph_reset.argtypes = [vp, C.POINTER(Rng), i64, vp]
    lib.dpivae_step_graph_launch.argtypes = [vp, i32, vp]
    lib.dpivae_step_graph_destroy.argtypes = [vp]
    for name in EXPORTS:
        getattr(lib, name)  # fail loudly on a stale library
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libdpivae_b200: " + load().dpivae_last_error().decode())
