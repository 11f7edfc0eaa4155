"""ctypes mirror of include/kanode.h and loader of libkanode_b200.so.

The product path has no CPU fallback: `load_library()` raises if the CUDA
library has not been built, and every compute entry point of the library
returns KANODE_ERR_NO_DEVICE when no sm_100-class GPU is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

KANODE_MAX_LAYERS = 8

# kanode_normalizer (kdense.jl:25,41-47)
NORM_TANH, NORM_SOFTSIGN, NORM_SIGMOID = 0, 1, 2
# kanode_basis (utils.jl:8-62)
BASIS_RBF, BASIS_RSWAF, BASIS_IQF = 0, 1, 2
# kanode_rhs_kind
RHS_CHAIN, RHS_SOURCE_LAPLACIAN, RHS_MAP = 0, 1, 2
LAYER_KDENSE, LAYER_DENSE = 0, 1
ACT_IDENTITY, ACT_TANH = 0, 1
# kanode_retcode
RET_SUCCESS, RET_MAXITERS, RET_DT_LESS_THAN_MIN, RET_UNSTABLE, RET_RECORD_OVERFLOW = range(5)
RETCODE_NAMES = ("Success", "MaxIters", "DtLessThanMin", "Unstable", "RecordOverflow")

ERR_NAMES = {0: "OK", -1: "INVALID", -2: "NO_DEVICE", -3: "CUDA", -4: "NOMEM", -5: "UNSUPPORTED", -6: "SOLVER"}
ERR_SOLVER = -6


class LayerDesc(C.Structure):
    _fields_ = [
        ("in_dims", C.c_int32), ("out_dims", C.c_int32), ("grid_len", C.c_int32),
        ("normalizer", C.c_int32), ("basis", C.c_int32), ("use_base_act", C.c_int32),
        ("grid_lo", C.c_float), ("grid_hi", C.c_float), ("denominator", C.c_float),
        ("kind", C.c_int32), ("dense_act", C.c_int32),
    ]


class Desc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("layers", LayerDesc * KANODE_MAX_LAYERS),
        ("rhs_kind", C.c_int32),
        ("n_state", C.c_int32),
        ("lap_coef", C.c_double),
        ("dx", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [("naccept", C.c_int32), ("nreject", C.c_int32), ("nf", C.c_int32), ("retcode", C.c_int32)]


REPO_ROOT = Path(__file__).resolve().parent.parent
LIB_NAME = "libkanode_b200.so"
LIB_PATH = Path(__file__).resolve().parent / "csrc" / LIB_NAME

# every symbol include/kanode.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = (
    "kanode_version", "kanode_last_error", "kanode_param_count", "kanode_create", "kanode_destroy",
    "kanode_sync", "kanode_set_params", "kanode_set_params_dev", "kanode_rhs", "kanode_rhs_dev",
    "kanode_vjp", "kanode_solve", "kanode_solve_dev", "kanode_loss_grad", "kanode_loss_grad_dev",
    "kanode_launch_count", "kanode_set_record_capacity",
    "kanode_set_params_f64", "kanode_rhs_f64", "kanode_vjp_f64", "kanode_solve_f64", "kanode_loss_grad_f64",
    "kanode_loss_grad_dev_f64", "kanode_last_timing", "kanode_adam_step_dev", "kanode_last_gpass_timing",
    "kanode_loss_grad_replay", "kanode_loss_grad_replay_f64",
    "kanode_solve_adjoint", "kanode_solve_adjoint_f64", "kanode_solve_adjoint_dev",
    "kanode_edge_activations", "kanode_edge_activations_f64", "kanode_set_regularizer", "kanode_reg_loss",
    "kanode_pack_sums_dev", "kanode_pack_sums_dev_f64", "kanode_train_apply_packed_dev", "kanode_create_multi", "kanode_device_count", "kanode_train_begin", "kanode_train_step_dev", "kanode_train_apply_dev", "kanode_train_params",
    "kanode_peer_export", "kanode_peer_attach", "kanode_pack_allreduce_dev", "kanode_pack_allreduce_dev_f64", "kanode_peer_status",
)

_lib = None


class KanodeError(RuntimeError):
    pass


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """Load libkanode_b200.so and declare its prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else Path(os.environ.get("KANODE_B200_LIB", LIB_PATH))
    if not p.exists():
        raise KanodeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the KAN-ODE hot path)")
    lib = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    vp, i64, f32p, f64p = C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_double)
    lib.kanode_version.restype = C.c_char_p
    lib.kanode_last_error.restype = C.c_char_p
    lib.kanode_last_error.argtypes = [vp]
    lib.kanode_param_count.restype = C.c_size_t
    lib.kanode_param_count.argtypes = [C.POINTER(Desc)]
    lib.kanode_create.argtypes = [C.POINTER(Desc), C.c_int, vp, C.POINTER(vp)]
    lib.kanode_destroy.argtypes = [vp]
    lib.kanode_sync.argtypes = [vp]
    lib.kanode_set_params.argtypes = [vp, vp, C.c_size_t]
    lib.kanode_set_params_dev.argtypes = [vp, vp, C.c_size_t]
    lib.kanode_rhs.argtypes = [vp, vp, vp, i64]
    lib.kanode_rhs_dev.argtypes = [vp, vp, vp, i64]
    lib.kanode_vjp.argtypes = [vp, vp, vp, vp, vp, i64]
    solve_args = [vp, vp, i64, C.c_double, C.c_double, vp, C.c_int32, C.c_float, C.c_float, vp, vp]
    lib.kanode_solve.argtypes = solve_args
    lib.kanode_solve_dev.argtypes = solve_args
    lg_args = [vp, vp, i64, C.c_double, C.c_double, vp, C.c_int32, vp, C.c_float, C.c_float, vp, vp, vp, vp, vp]
    lib.kanode_loss_grad.argtypes = lg_args
    lib.kanode_loss_grad_dev.argtypes = lg_args
    lib.kanode_set_record_capacity.argtypes = [vp, C.c_int32]
    lib.kanode_set_record_capacity.restype = C.c_int
    lg64 = [vp, vp, i64, C.c_double, C.c_double, vp, C.c_int32, vp, C.c_double, C.c_double, vp, vp, vp, vp, vp]
    lib.kanode_loss_grad_dev_f64.argtypes = lg64
    lib.kanode_loss_grad_dev_f64.restype = C.c_int
    lib.kanode_adam_step_dev.argtypes = [vp, vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
    lib.kanode_adam_step_dev.restype = C.c_int
    lib.kanode_last_timing.argtypes = [vp, vp]
    lib.kanode_last_timing.restype = C.c_int
    lib.kanode_last_gpass_timing.argtypes = [vp, vp, vp]
    lib.kanode_last_gpass_timing.restype = C.c_int
    lib.kanode_peer_export.argtypes = [vp, vp]
    lib.kanode_peer_attach.argtypes = [vp, C.c_int32, C.c_int32, vp]
    lib.kanode_pack_allreduce_dev.argtypes = [vp, vp, vp, i64, vp]
    lib.kanode_pack_allreduce_dev_f64.argtypes = [vp, vp, vp, i64, vp]
    lib.kanode_peer_status.argtypes = [vp]
    for name in ("kanode_peer_export", "kanode_peer_attach", "kanode_pack_allreduce_dev", "kanode_pack_allreduce_dev_f64", "kanode_peer_status"):
        getattr(lib, name).restype = C.c_int
    lib.kanode_launch_count.restype = C.c_int64
    lib.kanode_launch_count.argtypes = [vp]
    for name in ("kanode_create", "kanode_destroy", "kanode_sync", "kanode_set_params", "kanode_set_params_dev",
                 "kanode_rhs", "kanode_rhs_dev", "kanode_vjp", "kanode_solve", "kanode_solve_dev",
                 "kanode_loss_grad", "kanode_loss_grad_dev"):
        getattr(lib, name).restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def check(lib, handle, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.kanode_last_error(handle)
        raise KanodeError(f"{what} failed: {ERR_NAMES.get(rc, rc)}: {msg.decode() if msg else ''}")
