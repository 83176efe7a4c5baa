"""ctypes binding of libtib.so - the C ABI declared in include/tib.h.  Fails loudly when the
library is missing (the product path has no CPU or eager-PyTorch fallback)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtib.so")
ABI_VERSION = 6

VARIANT_AMBIENT, VARIANT_LATENT_MULTI_T, VARIANT_LATENT_SINGLE_T = 0, 1, 2
MATH_FP32_SIMT, MATH_F16X3_TC, MATH_F16_TC, MATH_F16X3_LAYERED = 0, 1, 2, 3
METHOD_EULER, METHOD_MIDPOINT, METHOD_RK4 = 0, 1, 2
KERNEL_KINDS = ("embed", "edge_init", "message", "update", "readout", "step", "train_gemm", "train_other")
N_KERNEL_KINDS = len(KERNEL_KINDS)
MATH_NAMES = {0: "fp32_simt", 1: "f16x3_tcgen05", 2: "f16_tcgen05", 3: "f16x3_tcgen05_layered"}
GAMMA_BROWNIAN, GAMMA_SIN2 = 0, 1
GAMMAS = {"brownian": GAMMA_BROWNIAN, "sin2": GAMMA_SIN2}
GEMM_STORE, GEMM_ATOMIC, GEMM_ACCUM = 0, 1, 2
METHODS = {"euler": METHOD_EULER, "midpoint": METHOD_MIDPOINT, "rk4": METHOD_RK4}


class ModelDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("variant", C.c_int32), ("n_features", C.c_int32),
                ("n_layers", C.c_int32), ("n_types", C.c_int32), ("n_edge_types", C.c_int32),
                ("temp_length", C.c_float), ("time_length", C.c_float), ("length_scale", C.c_float),
                ("temp_mean", C.c_float), ("temp_range", C.c_float)]


class Batch(C.Structure):
    _fields_ = [("n_mol", C.c_int32), ("n_nodes", C.c_int32), ("n_edges", C.c_int64), ("max_atoms", C.c_int32),
                ("mol_ptr", C.c_void_p), ("edge_ptr", C.c_void_p), ("atom_id", C.c_void_p),
                ("edge_type", C.c_void_p), ("temp0", C.c_void_p), ("temp1", C.c_void_p),
                ("n_embed_rows", C.c_int32), ("embed_index", C.c_void_p), ("embed_atom_id", C.c_void_p),
                ("embed_temp0", C.c_void_p), ("embed_temp1", C.c_void_p),
                ("n_tiles", C.c_int32), ("tile_node_ptr", C.c_void_p)]


class FixedOpts(C.Structure):
    _fields_ = [("method", C.c_int32), ("n_times", C.c_int32), ("t_grid", C.POINTER(C.c_float)),
                ("save_frames", C.c_int32), ("eps", C.c_float), ("noise", C.c_void_p),
                ("score_model", C.c_void_p)]


NORM_ALLREDUCE = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_void_p)


class Dopri5Opts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("n_times", C.c_int32),
                ("t_grid", C.POINTER(C.c_float)), ("save_frames", C.c_int32), ("max_attempts", C.c_int32),
                ("norm_allreduce", NORM_ALLREDUCE), ("norm_user", C.c_void_p)]


class Dopri5Stats(C.Structure):
    _fields_ = [("nfe", C.c_int32), ("attempts", C.c_int32), ("accepted", C.c_int32), ("last_dt", C.c_double)]


class Interpolant(C.Structure):
    _fields_ = [("gamma_kind", C.c_int32), ("a", C.c_float)]


class TrainBatch(C.Structure):
    _fields_ = [("n_mol", C.c_int32), ("n_nodes", C.c_int32), ("n_edges", C.c_int64),
                ("mol_ptr", C.c_void_p), ("edge_ptr", C.c_void_p), ("atom_id", C.c_void_p), ("edge_type", C.c_void_p),
                ("temp0", C.c_void_p), ("temp1", C.c_void_p), ("x0", C.c_void_p), ("x1", C.c_void_p),
                ("t", C.c_void_p), ("z", C.c_void_p)]


# every symbol include/tib.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("tib_packed_weight_count", C.c_size_t, [C.POINTER(ModelDesc)]),
    ("tib_model_create", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(ModelDesc), C.c_void_p, C.c_size_t, C.c_int]),
    ("tib_model_destroy", None, [C.c_void_p]),
    ("tib_model_set_math", C.c_int, [C.c_void_p, C.c_int]),
    ("tib_model_status", C.c_int, [C.c_void_p, C.c_void_p]),
    ("tib_debug_counters", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    ("tib_selftest_gemm", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("tib_workspace_bytes", C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64]),
    ("tib_drift", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("tib_div_workspace_bytes", C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    ("tib_drift_div", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("tib_div_rollout_workspace_bytes", C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    ("tib_rollout_fixed_dlogp", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(FixedOpts), C.c_float, C.c_float,
                                          C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("tib_rollout_dopri5_dlogp", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(Dopri5Opts), C.c_float, C.c_float,
                                           C.c_float, C.c_void_p, C.POINTER(Dopri5Stats), C.c_void_p, C.c_size_t, C.c_void_p]),
    ("tib_zmatrix", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("tib_tica_project", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_void_p]),
    ("tib_step_euler", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("tib_rollout_fixed", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(FixedOpts), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("tib_rollout_dopri5", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(Dopri5Opts), C.c_void_p, C.POINTER(Dopri5Stats), C.c_void_p, C.c_size_t, C.c_void_p]),
    ("tib_reweight_stats", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    ("tib_adw_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_int]),
    ("tib_adw_destroy", None, [C.c_void_p]),
    ("tib_adw_drift_div", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("tib_train_workspace_bytes", C.c_size_t, [C.POINTER(ModelDesc), C.c_int32, C.c_int32, C.c_int64]),
    ("tib_train_loss_grad", C.c_int, [C.POINTER(ModelDesc), C.c_void_p, C.POINTER(TrainBatch), C.POINTER(Interpolant), C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("tib_train_status", C.c_int, [C.c_void_p]),
    ("tib_train_gemm_flops", C.c_double, [C.c_int]),
    ("tib_gemm_debug", C.c_int, [C.c_int, C.c_void_p]),
    ("tib_adam_step", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    ("tib_gemm_f16x3", C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_float, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    ("tib_last_error", C.c_char_p, []),
    ("tib_abi_version", C.c_int, []),
    ("tib_launch_count", C.c_uint64, [C.c_int]),
    ("tib_profile_begin", C.c_int, []),
    ("tib_profile_end", C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
]

_lib = None


def load() -> C.CDLL:
    """dlopen libtib.so (once) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing. Build it with `python -m thermodynamic_interpolation_b200.build` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)      # AttributeError if the .so does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.tib_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libtib.so ABI {lib.tib_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = "libtib"):
    if rc != 0:
        msg = load().tib_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed: {msg}")
