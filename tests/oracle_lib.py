"""ctypes binding of the CPU oracle (oracle/build/liboracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB = ORACLE_DIR / "build" / "liboracle.so"
SELFTEST = ORACLE_DIR / "build" / "selftest"

I64_MIN = -(1 << 63)
I64_MAX = (1 << 63) - 1
EXACT, RELAXED, RESTRICTED = 0, 1, 2
LEL, FRONTIER = 1, 2


def build(force: bool = False) -> None:
    srcs = [ORACLE_DIR / f for f in ("oracle_capi.cpp", "ddo_oracle.hpp", "models.hpp", "selftest.cpp")]
    stale = force or not LIB.exists() or not SELFTEST.exists() or any(s.stat().st_mtime > min(LIB.stat().st_mtime, SELFTEST.stat().st_mtime) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "all"], check=True, capture_output=True)


class DDResult(C.Structure):
    _fields_ = [("has_best", C.c_int32), ("is_exact", C.c_int32), ("has_best_exact", C.c_int32), ("lel", C.c_int32),
                ("best_value", C.c_int64), ("best_exact_value", C.c_int64), ("expanded", C.c_uint64), ("transitions", C.c_uint64),
                ("n_layers", C.c_int32), ("cutset_size", C.c_int32), ("cutoff", C.c_int32), ("pad", C.c_int32)]


class SolveResult(C.Structure):
    _fields_ = [("has_value", C.c_int32), ("is_exact", C.c_int32), ("best_value", C.c_int64), ("best_lb", C.c_int64), ("best_ub", C.c_int64),
                ("explored", C.c_uint64), ("expanded", C.c_uint64), ("transitions", C.c_uint64), ("compilations", C.c_uint64), ("waves", C.c_uint64),
                ("seconds", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.oracle_misp_new.restype = C.c_void_p
        _lib.oracle_misp_new.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.oracle_misp_free.argtypes = [C.c_void_p]
        _lib.oracle_misp_words.argtypes = [C.c_void_p]
        _lib.oracle_misp_dd_new.restype = C.c_void_p
        _lib.oracle_misp_dd_new.argtypes = [C.c_void_p, C.c_int32]
        _lib.oracle_misp_dd_free.argtypes = [C.c_void_p]
        _lib.oracle_misp_dd_compile.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_int32, C.POINTER(DDResult)]
        _lib.oracle_misp_dd_layers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _lib.oracle_misp_dd_cutset.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int32, C.c_int32]
        _lib.oracle_misp_dd_solution.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        _lib.oracle_misp_solve.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_double, C.c_uint64,
                                           C.POINTER(SolveResult), C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        _lib.oracle_misp_compile_many.restype = C.c_uint64
        _lib.oracle_misp_compile_many.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        _lib.oracle_misp_stepper_new.restype = C.c_void_p
        _lib.oracle_misp_stepper_new.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64]
        for f in ("free", "finish"):
            getattr(_lib, f"oracle_misp_stepper_{f}").argtypes = [C.c_void_p]
        _lib.oracle_misp_stepper_init.argtypes = [C.c_void_p, C.c_int32]
        _lib.oracle_misp_stepper_wave.argtypes = [C.c_void_p, C.POINTER(C.c_int64 * 3)]
        _lib.oracle_misp_stepper_set_lb.argtypes = [C.c_void_p, C.c_int64]
        _lib.oracle_misp_stepper_retain_share.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        _lib.oracle_misp_stepper_state.argtypes = [C.c_void_p, C.POINTER(C.c_int64 * 6)]
        for pre in ("misp", "m2s"):
            getattr(_lib, f"oracle_{pre}_stepper_node_words").argtypes = [C.c_void_p]
            getattr(_lib, f"oracle_{pre}_stepper_export").argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
            getattr(_lib, f"oracle_{pre}_stepper_import").argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _lib.oracle_misp_stepper_solution.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _lib.oracle_misp_stepper_sol_value.restype = C.c_int64
        _lib.oracle_misp_stepper_sol_value.argtypes = [C.c_void_p]
        _lib.oracle_m2s_new.restype = C.c_void_p
        _lib.oracle_m2s_new.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
        _lib.oracle_m2s_free.argtypes = [C.c_void_p]
        _lib.oracle_m2s_words.argtypes = [C.c_void_p]
        _lib.oracle_m2s_initial_value.restype = C.c_int64
        _lib.oracle_m2s_initial_value.argtypes = [C.c_void_p]
        _lib.oracle_m2s_order.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oracle_m2s_transition.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _lib.oracle_m2s_dd_new.restype = C.c_void_p
        _lib.oracle_m2s_dd_new.argtypes = [C.c_void_p, C.c_int32]
        _lib.oracle_m2s_dd_free.argtypes = [C.c_void_p]
        _lib.oracle_m2s_dd_compile.argtypes = _lib.oracle_misp_dd_compile.argtypes
        _lib.oracle_m2s_dd_layers.argtypes = _lib.oracle_misp_dd_layers.argtypes
        _lib.oracle_m2s_dd_cutset.argtypes = _lib.oracle_misp_dd_cutset.argtypes
        _lib.oracle_m2s_dd_solution.argtypes = _lib.oracle_misp_dd_solution.argtypes
        _lib.oracle_m2s_solve.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_double, C.c_uint64,
                                          C.POINTER(SolveResult), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        _lib.oracle_m2s_compile_many.restype = C.c_uint64
        _lib.oracle_m2s_compile_many.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        _lib.oracle_m2s_stepper_new.restype = C.c_void_p
        _lib.oracle_m2s_stepper_new.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64]
        _lib.oracle_m2s_stepper_free.argtypes = [C.c_void_p]
        _lib.oracle_m2s_stepper_init.argtypes = [C.c_void_p, C.c_int32]
        _lib.oracle_m2s_stepper_wave.argtypes = [C.c_void_p, C.POINTER(C.c_int64 * 3)]
        _lib.oracle_m2s_stepper_state.argtypes = [C.c_void_p, C.POINTER(C.c_int64 * 6)]
        _lib.oracle_m2s_stepper_set_lb.argtypes = [C.c_void_p, C.c_int64]
        _lib.oracle_m2s_stepper_retain_share.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        _lib.oracle_m2s_stepper_finish.argtypes = [C.c_void_p]
        _lib.oracle_tsptw_new.restype = C.c_void_p
        _lib.oracle_tsptw_new.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
        _lib.oracle_tsptw_free.argtypes = [C.c_void_p]
        _lib.oracle_tsptw_words.argtypes = [C.c_void_p]
        _lib.oracle_tsptw_initial_state.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oracle_tsptw_dd_new.restype = C.c_void_p
        _lib.oracle_tsptw_dd_new.argtypes = [C.c_void_p, C.c_int32]
        _lib.oracle_tsptw_dd_free.argtypes = [C.c_void_p]
        _lib.oracle_tsptw_dd_compile.argtypes = _lib.oracle_misp_dd_compile.argtypes
        _lib.oracle_tsptw_dd_layers.argtypes = _lib.oracle_misp_dd_layers.argtypes
        _lib.oracle_tsptw_dd_cutset.argtypes = _lib.oracle_misp_dd_cutset.argtypes
        _lib.oracle_tsptw_dd_solution.argtypes = _lib.oracle_misp_dd_solution.argtypes
        _lib.oracle_knapsack_solve.argtypes = [C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, C.c_int32,
                                               C.POINTER(SolveResult), C.c_void_p]
        _lib.oracle_locbounds_dump.argtypes = [C.c_int32, C.c_int64, C.c_char_p, C.c_int32]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _OracleModel:
    """Entry points shared by the models (oracle_<prefix>_dd_* of oracle/oracle_capi.cpp)."""

    PREFIX = ""

    def _f(self, name):
        return getattr(lib(), f"oracle_{self.PREFIX}_{name}")

    def __del__(self):
        try:
            self._f("free")(self.h)
        except Exception:
            pass

    def compile(self, comp_type, max_width, root_state=None, root_value=None, root_depth=0, best_lb=I64_MIN, cutset_type=LEL, cutoff=False, want_paths=False):
        class _L:
            pass
        L = _L()
        for nm in ("dd_new", "dd_free", "dd_compile", "dd_layers", "dd_cutset", "dd_solution"):
            setattr(L, f"oracle_misp_{nm}", self._f(nm))
        if root_value is None:
            root_value = self.initial_value()
        dd = L.oracle_misp_dd_new(self.h, cutset_type)
        try:
            if root_state is None:
                root_state = self.inst.initial_state()
            root_state = np.ascontiguousarray(root_state, dtype=np.uint64)
            res = DDResult()
            rc = L.oracle_misp_dd_compile(dd, comp_type, max_width, _p(root_state), root_value, root_depth, best_lb, int(cutoff), C.byref(res))
            out = {k: getattr(res, k) for k, _ in DDResult._fields_ if k != "pad"}
            out["rc"] = rc
            if rc != 0:
                return out
            n = self.inst.n
            vars_ = np.zeros(n + 1, dtype=np.int32)
            widths = np.zeros(n + 1, dtype=np.int32)
            nl = L.oracle_misp_dd_layers(dd, _p(vars_), _p(widths), n + 1)
            out["layer_vars"] = vars_[:nl].copy()
            out["layer_widths"] = widths[:nl].copy()
            cs = res.cutset_size
            states = np.zeros((max(cs, 1), self.words), dtype=np.uint64)
            values = np.zeros(max(cs, 1), dtype=np.int64)
            ubs = np.zeros(max(cs, 1), dtype=np.int64)
            depths = np.zeros(max(cs, 1), dtype=np.int32)
            plens = np.zeros(max(cs, 1), dtype=np.int32)
            stride = n
            paths = np.zeros((max(cs, 1), stride, 2), dtype=np.int32) if want_paths else None
            L.oracle_misp_dd_cutset(dd, _p(states), _p(values), _p(ubs), _p(depths), _p(plens), _p(paths), cs, stride)
            out["cutset_states"] = states[:cs]
            out["cutset_values"] = values[:cs]
            out["cutset_ubs"] = ubs[:cs]
            out["cutset_depths"] = depths[:cs]
            if want_paths:
                out["cutset_paths"] = [paths[i, : plens[i]].copy() for i in range(cs)]
            for key, ex in (("best_solution", 0), ("best_exact_solution", 1)):
                sv = np.zeros(n + 1, dtype=np.int32)
                sx = np.zeros(n + 1, dtype=np.int32)
                ln = L.oracle_misp_dd_solution(dd, ex, _p(sv), _p(sx), n + 1)
                out[key] = None if ln < 0 else list(zip(sv[:ln].tolist(), sx[:ln].tolist()))
            return out
        finally:
            L.oracle_misp_dd_free(dd)



class OracleMisp(_OracleModel):
    """CPU oracle for one MISP instance."""

    PREFIX = "misp"

    def __init__(self, inst):
        self.inst = inst
        self.h = lib().oracle_misp_new(inst.n, _p(inst.weights), len(inst.src), _p(inst.src), _p(inst.dst))
        self.words = lib().oracle_misp_words(self.h)

    def initial_value(self):
        return 0

    def solve(self, mode="sequential", k=1, width=None, cutset_type=LEL, time_budget_s=0.0, max_waves=0, trace_cap=0, width_kind=None):
        """mode: sequential | wave | parallel.  width None -> NbUnassignedWidth (the reference CLI default, misp/main.rs:322-328);
        width_kind 2 / 3: Times(width, NbUnassignedWidth) / DivBy(width, NbUnassignedWidth)."""
        m = {"sequential": 0, "wave": 1, "parallel": 2}[mode]
        wk = width_kind if width_kind is not None else (0 if width is not None else 1)
        res = SolveResult()
        sol = np.zeros(self.inst.n, dtype=np.int32)
        sl = C.c_int32(0)
        trace = np.zeros((max(trace_cap, 1), 4), dtype=np.int64)
        tl = C.c_int32(0)
        lib().oracle_misp_solve(self.h, m, k, wk, width or 0, cutset_type, time_budget_s, max_waves, C.byref(res),
                                _p(sol), C.byref(sl), _p(trace) if trace_cap else None, trace_cap, C.byref(tl))
        out = {k_: getattr(res, k_) for k_, _ in SolveResult._fields_}
        out["solution"] = sorted(sol[: sl.value].tolist())
        out["trace"] = trace[: tl.value].copy()
        return out

    def compile_many(self, roots_states, roots_values, roots_depths, widths, best_lb, threads, cutset_type=LEL):
        n = len(roots_values)
        rs = np.ascontiguousarray(roots_states, dtype=np.uint64)
        rv = np.ascontiguousarray(roots_values, dtype=np.int64)
        rd = np.ascontiguousarray(roots_depths, dtype=np.int32)
        w = np.ascontiguousarray(widths, dtype=np.uint64)
        rb = np.zeros(n, dtype=np.int64)
        xb = np.zeros(n, dtype=np.int64)
        cs = np.zeros(n, dtype=np.int32)
        tr = C.c_uint64(0)
        sec = C.c_double(0)
        exp = lib().oracle_misp_compile_many(self.h, threads, n, _p(rs), _p(rv), _p(rd), _p(w), best_lb, cutset_type, _p(rb), _p(xb), _p(cs), C.byref(tr), C.byref(sec))
        return {"expanded": int(exp), "transitions": int(tr.value), "seconds": float(sec.value), "restricted_best": rb, "relaxed_best": xb, "cutset_sizes": cs}


class OracleM2s(_OracleModel):
    """CPU oracle for one MAX2SAT instance (states: n int32 benefits packed in uint64 words)."""

    PREFIX = "m2s"

    def __init__(self, inst):
        self.inst = inst
        self.h = lib().oracle_m2s_new(inst.n, len(inst.clauses), _p(inst.clauses))
        self.words = lib().oracle_m2s_words(self.h)

    def initial_value(self):
        return int(lib().oracle_m2s_initial_value(self.h))

    def order(self):
        o = np.zeros(self.inst.n, dtype=np.int32)
        lib().oracle_m2s_order(self.h, _p(o))
        return o

    def transition(self, sub, depth, var, value):
        sub = np.ascontiguousarray(sub, dtype=np.int32)
        out = np.zeros(self.inst.n, dtype=np.int32)
        cost, rank, rub = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        lib().oracle_m2s_transition(self.h, _p(sub), depth, var, value, _p(out), C.byref(cost), C.byref(rank), C.byref(rub))
        return out, int(cost.value), int(rank.value), int(rub.value)

    def solve(self, mode="sequential", k=1, width=None, cutset_type=LEL, time_budget_s=0.0, max_waves=0, trace_cap=0, width_kind=None):
        """mode: sequential | wave | parallel.  width None -> NbUnassignedWidth (max2sat/main.rs:87-93)."""
        m = {"sequential": 0, "wave": 1, "parallel": 2}[mode]
        wk = width_kind if width_kind is not None else (0 if width is not None else 1)
        res = SolveResult()
        sv = np.zeros(self.inst.n + 1, dtype=np.int32)
        sx = np.zeros(self.inst.n + 1, dtype=np.int32)
        sl = C.c_int32(0)
        trace = np.zeros((max(trace_cap, 1), 4), dtype=np.int64)
        tl = C.c_int32(0)
        lib().oracle_m2s_solve(self.h, m, k, wk, width or 0, cutset_type, time_budget_s, max_waves, C.byref(res),
                               _p(sv), _p(sx), C.byref(sl), _p(trace) if trace_cap else None, trace_cap, C.byref(tl))
        out = {k_: getattr(res, k_) for k_, _ in SolveResult._fields_}
        out["solution"] = list(zip(sv[: sl.value].tolist(), sx[: sl.value].tolist()))
        out["trace"] = trace[: tl.value].copy()
        return out


class OracleTsptw(_OracleModel):
    """CPU oracle for one TSPTW instance at the DD level (states: 16 uint64 words, see oracle_capi.cpp::TsptwHandle) -- the checker of the
    TSPTW device model to come; `compile` is the shared _OracleModel.compile."""

    PREFIX = "tsptw"

    def __init__(self, inst):
        self.inst = inst
        self.h = lib().oracle_tsptw_new(inst.n, _p(np.ascontiguousarray(inst.dist)), _p(np.ascontiguousarray(inst.tw)))
        self.words = lib().oracle_tsptw_words(self.h)
        inst.initial_state = self.initial_state  # what _OracleModel.compile asks the instance for

    def initial_state(self):
        out = np.zeros(self.words, dtype=np.uint64)
        lib().oracle_tsptw_initial_state(self.h, _p(out))
        return out

    def initial_value(self):
        return 0


def _m2s_compile_many(self, roots_states, roots_values, roots_depths, widths, best_lb, threads, cutset_type=LEL, time_budget_s=0.0):
    """`threads` workers compile restricted + relaxed DDs of independent roots (parallel.rs:391-437 without the fringe), optionally time-boxed."""
    n = len(roots_values)
    rs = np.ascontiguousarray(roots_states, dtype=np.uint64)
    rv = np.ascontiguousarray(roots_values, dtype=np.int64)
    rd = np.ascontiguousarray(roots_depths, dtype=np.int32)
    w = np.ascontiguousarray(widths, dtype=np.uint64)
    rb = np.zeros(n, dtype=np.int64)
    xb = np.zeros(n, dtype=np.int64)
    cs = np.zeros(n, dtype=np.int32)
    tr = C.c_uint64(0)
    sec = C.c_double(0)
    exp = lib().oracle_m2s_compile_many(self.h, threads, n, _p(rs), _p(rv), _p(rd), _p(w), best_lb, cutset_type, time_budget_s, _p(rb), _p(xb), _p(cs),
                                        C.byref(tr), C.byref(sec))
    return {"expanded": int(exp), "transitions": int(tr.value), "seconds": float(sec.value), "restricted_best": rb, "relaxed_best": xb, "cutset_sizes": cs}


OracleM2s.compile_many = _m2s_compile_many


class OracleStepper:
    """Stepwise CPU wave solver with the interface ddo_b200.sharded.sharded_maximize expects (stand-in for the device solver)."""

    def __init__(self, oracle, wave_size: int, width=None):
        self.o = oracle
        self.p = oracle.PREFIX
        self.h = self._f("stepper_new")(oracle.h, wave_size, 0 if width is not None else 1, width or 0)

    def _f(self, name):
        return getattr(lib(), f"oracle_{self.p}_{name}")

    def __del__(self):
        try:
            self._f("stepper_free")(self.h)
        except Exception:
            pass

    def init(self, push_root=True):
        self._f("stepper_init")(self.h, int(push_root))

    def wave(self):
        out = (C.c_int64 * 3)()
        rc = self._f("stepper_wave")(self.h, C.byref(out))
        assert rc == 0
        return int(out[0]), int(out[1]), int(out[2])

    def set_lower_bound(self, lb):
        self._f("stepper_set_lb")(self.h, lb)

    def retain_share(self, rank, nranks):
        self._f("stepper_retain_share")(self.h, rank, nranks)

    def finish(self):
        self._f("stepper_finish")(self.h)

    def _state(self):
        out = (C.c_int64 * 6)()
        self._f("stepper_state")(self.h, C.byref(out))
        return [int(x) for x in out]

    def fringe_len(self):
        return self._state()[2]

    def best_lower_bound(self):
        return self._state()[0]

    def explored(self):
        return self._state()[3]

    def expanded(self):
        return self._state()[4]

    # work hand-off / solution gather of ddo_b200.sharded (same packed rows as the device solver's export_open / import_open)
    def node_words(self):
        return int(self._f("stepper_node_words")(self.h))

    def export_open(self, max_nodes):
        rows = np.zeros((max(max_nodes, 1), self.node_words()), dtype=np.int64)
        k = self._f("stepper_export")(self.h, max_nodes, _p(rows))
        return rows[:k].copy()

    def import_open(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, self.node_words())
        if rows.shape[0]:
            self._f("stepper_import")(self.h, rows.shape[0], _p(rows))

    def best_value(self):
        if self.p != "misp":
            return None
        v = int(lib().oracle_misp_stepper_sol_value(self.h))
        return None if v == I64_MIN else v

    def best_solution(self):
        if self.p != "misp":
            return None
        n = self.o.inst.n
        sv = np.zeros(n + 1, dtype=np.int32); sx = np.zeros(n + 1, dtype=np.int32)
        ln = lib().oracle_misp_stepper_solution(self.h, _p(sv), _p(sx), n + 1)
        if ln < 0:
            return None
        from ddo_b200 import Decision
        return [Decision(int(a), int(b)) for a, b in zip(sv[:ln], sx[:ln])]


def knapsack_solve(inst, solver="sequential", k=1, width=None, cutset_type=FRONTIER, caching=True):
    res = SolveResult()
    taken = np.zeros(len(inst.profit), dtype=np.int32)
    lib().oracle_knapsack_solve(len(inst.profit), inst.capacity, _p(np.ascontiguousarray(inst.profit, dtype=np.int64)),
                                _p(np.ascontiguousarray(inst.weight, dtype=np.int64)), 0 if solver == "sequential" else 2, k,
                                0 if width is not None else 1, width or 0, cutset_type, int(caching), C.byref(res), _p(taken))
    out = {k_: getattr(res, k_) for k_, _ in SolveResult._fields_}
    out["taken"] = taken
    return out


def TsptwInstance(text: str):
    """examples/tsptw/instance.rs:51-108 through the host-side parser of the package (ddo_b200/instances.py::parse_tsptw)."""
    from ddo_b200.instances import parse_tsptw

    return parse_tsptw(text)


def tsptw_solve(inst, factor=1, solver="sequential", k=1, cutset_type=FRONTIER, caching=True, time_budget_s=0.0):
    """examples/tsptw/tests.rs:33-57 (`solve(instance, width, threads)`); returns the result fields plus `cost` = -(best_value) / 10000 as f32."""
    res = SolveResult()
    perm = np.zeros(inst.n, dtype=np.int32)
    lib().oracle_tsptw_solve.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                         C.POINTER(SolveResult), C.c_void_p]
    lib().oracle_tsptw_solve(inst.n, _p(np.ascontiguousarray(inst.dist)), _p(np.ascontiguousarray(inst.tw)), factor, 0 if solver == "sequential" else 2, k,
                             cutset_type, int(caching), time_budget_s, C.byref(res), _p(perm))
    out = {k_: getattr(res, k_) for k_, _ in SolveResult._fields_}
    out["perm"] = perm
    out["cost"] = float(-(np.float32(res.best_value)) / np.float32(10000.0)) if res.has_value else -1.0
    return out


def locbounds_dump(cutset_type=FRONTIER, best_lb=0) -> str:
    buf = C.create_string_buffer(1 << 16)
    n = lib().oracle_locbounds_dump(cutset_type, best_lb, buf, len(buf))
    assert n >= 0
    return buf.value.decode()


def run_selftest() -> subprocess.CompletedProcess:
    build()
    return subprocess.run([str(SELFTEST)], capture_output=True, text=True)


def read_clq(golden_dir, name: str) -> str:
    """Text of the DIMACS fixture `name` under tests/golden/misp (the two n = 500 instances are stored gzip-compressed)."""
    import gzip
    from pathlib import Path
    p = Path(golden_dir) / "misp" / f"{name}.clq"
    if p.exists():
        return p.read_text()
    return gzip.decompress((Path(golden_dir) / "misp" / f"{name}.clq.gz").read_bytes()).decode()
