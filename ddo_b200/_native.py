"""ctypes loader for libddo_b200.so (the C ABI of include/ddo_b200.h).  Fails loudly when the CUDA extension is missing."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libddo_b200.so"

OK, CUTOFF = 0, 1
ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED, ERR_NO_DEVICE = -1, -2, -3, -4, -5
EXACT, RELAXED, RESTRICTED = 0, 1, 2
LAST_EXACT_LAYER, FRONTIER = 1, 2
WIDTH_FIXED, WIDTH_NB_UNASSIGNED, WIDTH_TIMES_NB_UNASSIGNED, WIDTH_DIVBY_NB_UNASSIGNED = 0, 1, 2, 3
I64_MIN, I64_MAX = -(1 << 63), (1 << 63) - 1


class Decision(C.Structure):
    _fields_ = [("variable", C.c_int32), ("value", C.c_int32)]


class Completion(C.Structure):
    _fields_ = [("is_exact", C.c_int32), ("has_best_value", C.c_int32), ("best_value", C.c_int64), ("has_best_exact", C.c_int32),
                ("cutset_size", C.c_int32), ("best_exact_value", C.c_int64), ("lel_depth", C.c_int32), ("n_layers", C.c_int32),
                ("expanded", C.c_uint64), ("transitions", C.c_uint64)]


# every symbol include/ddo_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "ddo_last_error": (C.c_char_p, []),
    "ddo_device_count": (C.c_int, []),
    "ddo_kernel_launches": (C.c_uint64, []),
    "ddo_model_create_misp": (C.c_int, [C.c_int32, _P, C.c_int64, _P, _P, C.c_int, C.POINTER(_P)]),
    "ddo_model_create_max2sat": (C.c_int, [C.c_int32, C.c_int64, _P, C.c_int, C.POINTER(_P)]),
    "ddo_model_kind": (C.c_int32, [_P]),
    "ddo_model_destroy": (None, [_P]),
    "ddo_model_nb_variables": (C.c_int32, [_P]),
    "ddo_model_state_words": (C.c_int32, [_P]),
    "ddo_model_initial_state": (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    "ddo_mdd_create": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "ddo_mdd_destroy": (None, [_P]),
    "ddo_mdd_compile": (C.c_int, [_P, C.c_int32, C.c_uint64, _P, C.c_int64, C.c_int32, C.c_int64, _P, C.POINTER(Completion)]),
    "ddo_mdd_compile_batch": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_int64, _P, _P]),
    "ddo_mdd_best_solution": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.POINTER(C.c_int32)]),
    "ddo_mdd_drain_cutset": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, C.POINTER(C.c_int32)]),
    "ddo_mdd_drain_cutset_batch": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "ddo_mdd_drain_layer_index": (C.c_int, [_P, _P, C.c_int64]),
    "ddo_mdd_set_profiling": (C.c_int, [_P, C.c_int32]),
    "ddo_mdd_kernel_times": (C.c_int, [_P, C.POINTER(C.c_double * 6), C.POINTER(C.c_uint64 * 6)]),
    "ddo_mdd_layer_trace": (C.c_int, [_P, C.c_int32, _P, _P, C.c_int32]),
    "ddo_mdd_stage_roots": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P]),
    "ddo_mdd_compile_staged": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_float)]),
    "ddo_mdd_fetch_completions": (C.c_int, [_P, C.c_int32, _P]),
    "ddo_solver_create": (C.c_int, [_P, _P, C.c_int32, C.c_uint64, C.c_int32, C.POINTER(_P)]),
    "ddo_solver_destroy": (None, [_P]),
    "ddo_solver_maximize": (C.c_int, [_P, C.c_double, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "ddo_solver_init": (C.c_int, [_P, C.c_int32]),
    "ddo_solver_wave": (C.c_int, [_P, _P, C.POINTER(C.c_int64 * 3)]),
    "ddo_solver_set_lower_bound": (C.c_int, [_P, C.c_int64]),
    "ddo_solver_retain_share": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "ddo_solver_node_words": (C.c_int32, [_P]),
    "ddo_solver_export_open": (C.c_int, [_P, C.c_int32, _P, C.POINTER(C.c_int32)]),
    "ddo_solver_import_open": (C.c_int, [_P, C.c_int32, _P]),
    "ddo_solver_maximize_sharded": (C.c_int, [_P, _P, C.c_double, C.c_uint64, C.c_int32, C.POINTER(C.c_int64 * 8)]),
    "ddo_solver_finish": (C.c_int, [_P]),
    "ddo_solver_best_lower_bound": (C.c_int64, [_P]),
    "ddo_solver_best_upper_bound": (C.c_int64, [_P]),
    "ddo_solver_best_value": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "ddo_solver_best_solution": (C.c_int, [_P, _P, C.POINTER(C.c_int32)]),
    "ddo_solver_explored": (C.c_uint64, [_P]),
    "ddo_solver_fringe_len": (C.c_uint64, [_P]),
    "ddo_solver_stats": (C.c_int, [_P, C.POINTER(C.c_double * 8)]),
    "ddo_comm_unique_id": (C.c_int, [_P]),
    "ddo_comm_init": (C.c_int, [C.c_int32, C.c_int32, _P, C.c_int, C.POINTER(_P)]),
    "ddo_comm_destroy": (None, [_P]),
    "ddo_comm_size": (C.c_int32, [_P]),
    "ddo_comm_rank": (C.c_int32, [_P]),
    "ddo_comm_allreduce_max": (C.c_int, [_P, _P, C.c_int32]),
    "ddo_comm_allgather": (C.c_int, [_P, _P, C.c_int32, _P]),
    "ddo_comm_send": (C.c_int, [_P, _P, C.c_int64, C.c_int32]),
    "ddo_comm_recv": (C.c_int, [_P, _P, C.c_int64, C.c_int32]),
}

_lib = None


def build(force: bool = False) -> Path:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(PKG / "csrc")] + (["-B"] if force else []) + ["all"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libddo_b200.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        import os
        path = os.environ.get("DDO_B200_LIB", str(LIB_PATH))  # development: an alternative build of the same library (kernel tuning A/B runs)
        _lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


class DdoError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().ddo_last_error()
        super().__init__(f"{where}: status {code}: {msg.decode() if msg else ''}")


def check(code: int, where: str) -> int:
    if code < 0:
        raise DdoError(code, where)
    return code
