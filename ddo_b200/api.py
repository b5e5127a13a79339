"""Host-side mirror of the reference's plugin interface for the hot path, over the C ABI.

Reference items mirrored (paths relative to /root/reference/ddo/):
  Problem / Relaxation / StateRanking for MISP   examples/misp/main.rs:37-209    -> ``Misp`` (one declarative device model)
  Max2Sat / Max2SatRelax / Max2SatRanking        examples/max2sat/{model,relax,heuristics}.rs -> ``Max2Sat``
  CompilationInput / CompilationType             src/abstraction/mdd.rs:40-71    -> ``GpuMdd.compile(...)`` keyword arguments
  DecisionDiagram                                src/abstraction/mdd.rs:75-114   -> ``GpuMdd``
  SubProblem / Decision / Completion / Reason    src/common.rs:58-121            -> ``SubProblem`` / ``Decision`` / ``Completion`` / ``CutoffOccurred``
  FixedWidth / NbUnassignedWidth                 src/implementation/heuristics/width.rs:166-170,397-401
  Solver::maximize & getters                     src/abstraction/solver.rs:32-97 -> ``ParNoCachingSolverLel`` (= DefaultSolver, solver/mod.rs:29)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N
from .instances import Max2SatInstance, MispInstance

LAST_EXACT_LAYER = N.LAST_EXACT_LAYER
FRONTIER = N.FRONTIER


class CompilationType:  # src/abstraction/mdd.rs:40-47
    Exact = N.EXACT
    Relaxed = N.RELAXED
    Restricted = N.RESTRICTED


class CutoffOccurred(Exception):
    """Err(Reason::CutoffOccurred), src/common.rs:108-111."""


@dataclass
class Decision:  # src/common.rs:58-64
    variable: int
    value: int


@dataclass
class SubProblem:  # src/common.rs:75-87
    state: np.ndarray
    value: int
    path: List[Decision] = field(default_factory=list)
    ub: int = N.I64_MAX
    depth: int = 0


@dataclass
class Completion:  # src/common.rs:115-121 + DecisionDiagram getters
    is_exact: bool
    best_value: Optional[int]
    best_exact_value: Optional[int]
    cutset_size: int
    lel_depth: int
    n_layers: int
    expanded: int
    transitions: int


def device_count() -> int:
    return N.lib().ddo_device_count()


def kernel_launches() -> int:
    return int(N.lib().ddo_kernel_launches())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class FixedWidth:  # width.rs:166-170
    def __init__(self, w: int):
        self.w = int(w)

    def max_width(self, sp: SubProblem) -> int:
        return self.w


class NbUnassignedWidth:  # width.rs:397-401
    def __init__(self, nb_vars: int):
        self.nb_vars = int(nb_vars)

    def max_width(self, sp: SubProblem) -> int:
        return self.nb_vars - sp.depth


class Times:  # width.rs:636-641
    def __init__(self, k: int, inner):
        self.k, self.inner = int(k), inner

    def max_width(self, sp: SubProblem) -> int:
        return max(1, self.k * self.inner.max_width(sp))


class DivBy:  # width.rs:875-880
    def __init__(self, k: int, inner):
        self.k, self.inner = int(k), inner

    def max_width(self, sp: SubProblem) -> int:
        return max(1, self.inner.max_width(sp) // self.k)


def solver_gap(lb: int, ub: int) -> float:
    """Solver::gap (abstraction/solver.rs:80-93): 1.0 while a bound is missing, else (max(|ub|, |lb|) - min(|ub|, |lb|)) / max(|ub|, |lb|) in f32
    (0 / 0 is NaN, as in the reference)."""
    if ub == N.I64_MAX or lb == N.I64_MIN:
        return 1.0
    u, l = max(abs(ub), abs(lb)), min(abs(ub), abs(lb))
    if u == 0:
        return float("nan")
    return float(np.float32(u - l) / np.float32(u))


def _width_spec(width, nb_vars: int):
    """(kind, parameter, largest width it can ask for) of a WidthHeuristic for the C ABI."""
    if isinstance(width, FixedWidth):
        return N.WIDTH_FIXED, width.w, width.w
    if isinstance(width, NbUnassignedWidth):
        return N.WIDTH_NB_UNASSIGNED, 0, nb_vars
    if isinstance(width, (Times, DivBy)) and isinstance(width.inner, FixedWidth):  # a constant: fold it
        w = width.max_width(SubProblem(None, 0))
        return N.WIDTH_FIXED, w, w
    if isinstance(width, Times) and isinstance(width.inner, NbUnassignedWidth):
        return N.WIDTH_TIMES_NB_UNASSIGNED, width.k, width.k * nb_vars
    if isinstance(width, DivBy) and isinstance(width.inner, NbUnassignedWidth):
        return N.WIDTH_DIVBY_NB_UNASSIGNED, width.k, nb_vars
    raise TypeError("unsupported WidthHeuristic: FixedWidth, NbUnassignedWidth, Times(k, .) and DivBy(k, .) of those are available")


class Misp:
    """MISP as a device model: Problem + MispRelax + MispRanking of examples/misp/main.rs:37-209, resident in HBM."""

    def __init__(self, inst: MispInstance, device: int = 0):
        self.inst = inst
        self.device = device
        h = C.c_void_p()
        N.check(N.lib().ddo_model_create_misp(inst.n, _ptr(inst.weights), len(inst.src), _ptr(inst.src), _ptr(inst.dst), device, C.byref(h)),
                "ddo_model_create_misp")
        self.h = h
        self.words = N.lib().ddo_model_state_words(h)

    def nb_variables(self) -> int:  # dp.rs:39
        return N.lib().ddo_model_nb_variables(self.h)

    def initial_state(self) -> np.ndarray:  # dp.rs:41 / main.rs:69-71
        s = np.zeros(self.words, dtype=np.uint64)
        N.check(N.lib().ddo_model_initial_state(self.h, _ptr(s), None), "ddo_model_initial_state")
        return s

    def initial_value(self) -> int:  # dp.rs:43
        return 0

    def close(self):
        if getattr(self, "h", None):
            N.lib().ddo_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Max2Sat:
    """MAX2SAT as a device model: Max2Sat + Max2SatRelax + Max2SatRanking of examples/max2sat/{model,relax,heuristics}.rs, resident in HBM.
    A state is n int32 marginal benefits packed two per uint64 word (``unpack_state`` gives the benefits); its depth is SubProblem.depth."""

    def __init__(self, inst: Max2SatInstance, device: int = 0):
        self.inst = inst
        self.device = device
        h = C.c_void_p()
        N.check(N.lib().ddo_model_create_max2sat(inst.n, len(inst.clauses), _ptr(inst.clauses), device, C.byref(h)), "ddo_model_create_max2sat")
        self.h = h
        self.words = N.lib().ddo_model_state_words(h)

    def nb_variables(self) -> int:  # dp.rs:39
        return N.lib().ddo_model_nb_variables(self.h)

    def _initial(self):
        s = np.zeros(self.words, dtype=np.uint64)
        v = C.c_int64(0)
        N.check(N.lib().ddo_model_initial_state(self.h, _ptr(s), C.byref(v)), "ddo_model_initial_state")
        return s, int(v.value)

    def initial_state(self) -> np.ndarray:  # model.rs:259-264
        return self._initial()[0]

    def initial_value(self) -> int:  # model.rs:266-269: sum of the tautologies
        return self._initial()[1]

    def unpack_state(self, state: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(state, dtype=np.uint64).view(np.int32)[: self.inst.n].copy()

    def pack_state(self, benefits) -> np.ndarray:
        b = np.zeros(2 * self.words, dtype=np.int32)
        b[: self.inst.n] = benefits
        return b.view(np.uint64).copy()

    def close(self):
        if getattr(self, "h", None):
            N.lib().ddo_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _completion(c: N.Completion) -> Completion:
    return Completion(bool(c.is_exact), int(c.best_value) if c.has_best_value else None, int(c.best_exact_value) if c.has_best_exact else None,
                      int(c.cutset_size), int(c.lel_depth), int(c.n_layers), int(c.expanded), int(c.transitions))


class GpuMdd:
    """DecisionDiagram (mdd.rs:75-114) backed by a batch of device DD workspaces (`D::default()`, parallel.rs:580)."""

    def __init__(self, problem: Misp, max_width_cap: int, batch_cap: int = 1, cutset_type: int = LAST_EXACT_LAYER):
        self.problem = problem
        self.batch_cap = batch_cap
        self.cutset_type = cutset_type
        h = C.c_void_p()
        N.check(N.lib().ddo_mdd_create(problem.h, problem.device, max_width_cap, batch_cap, cutset_type, C.byref(h)), "ddo_mdd_create")
        self.h = h
        self._last: List[Completion] = []
        self._roots: List[SubProblem] = []

    # -- DecisionDiagram::compile -------------------------------------------------------------------------------
    def compile(self, comp_type: int, max_width: int, residual: SubProblem, best_lb: int = N.I64_MIN, cutoff: Optional[np.ndarray] = None) -> Completion:
        return self.compile_batch(comp_type, [max_width], [residual], best_lb, cutoff)[0]

    def compile_batch(self, comp_type: int, max_widths: Sequence[int], residuals: Sequence[SubProblem], best_lb: int = N.I64_MIN,
                      cutoff: Optional[np.ndarray] = None) -> List[Completion]:
        n = len(residuals)
        widths = np.asarray(max_widths, dtype=np.uint64)
        states = np.ascontiguousarray(np.stack([np.asarray(r.state, dtype=np.uint64) for r in residuals]))
        values = np.asarray([r.value for r in residuals], dtype=np.int64)
        depths = np.asarray([r.depth for r in residuals], dtype=np.int32)
        out = (N.Completion * n)()
        rc = N.check(N.lib().ddo_mdd_compile_batch(self.h, n, comp_type, _ptr(widths), _ptr(states), _ptr(values), _ptr(depths), best_lb,
                                                  _ptr(cutoff) if cutoff is not None else None, out), "ddo_mdd_compile_batch")
        if rc == N.CUTOFF:
            raise CutoffOccurred()
        self._last = [_completion(c) for c in out]
        self._roots = list(residuals)
        return self._last

    # -- getters (valid for DD `index` of the last batch) ---------------------------------------------------------
    def is_exact(self, index: int = 0) -> bool:
        return self._last[index].is_exact

    def best_value(self, index: int = 0) -> Optional[int]:
        return self._last[index].best_value

    def best_exact_value(self, index: int = 0) -> Optional[int]:
        return self._last[index].best_exact_value

    def _solution(self, index: int, exact: int) -> Optional[List[Decision]]:
        c = self._last[index]
        if (c.best_exact_value if exact else c.best_value) is None:
            return None
        cap = self.problem.nb_variables() + 1
        buf = (N.Decision * cap)()
        ln = C.c_int32(cap)
        N.check(N.lib().ddo_mdd_best_solution(self.h, index, exact, buf, C.byref(ln)), "ddo_mdd_best_solution")
        return list(self._roots[index].path) + [Decision(buf[i].variable, buf[i].value) for i in range(ln.value)]  # clean.rs:335

    def best_solution(self, index: int = 0):
        return self._solution(index, 0)

    def best_exact_solution(self, index: int = 0):
        return self._solution(index, 1)

    def drain_cutset(self, index: int = 0, ub_cap: int = N.I64_MAX, lb_filter: int = N.I64_MIN, with_paths: bool = True) -> List[SubProblem]:
        """drain_cutset (clean.rs:417-445): the MARKED cutset nodes as SubProblems (path = root path ++ decisions, clean.rs:329-343)."""
        if self.cutset_type == N.FRONTIER:
            return self._drain_frontier(index, ub_cap, lb_filter, with_paths)
        c = self._last[index]
        cap = max(c.cutset_size, 1)
        W = self.problem.words
        states = np.zeros((cap, W), dtype=np.uint64)
        values = np.zeros(cap, dtype=np.int64)
        ubs = np.zeros(cap, dtype=np.int64)
        depth = C.c_int32(0)
        plen = C.c_int32(0)
        count = C.c_int32(cap)
        pcap = self.problem.nb_variables()
        paths = (N.Decision * (cap * pcap))() if with_paths else None
        N.check(N.lib().ddo_mdd_drain_cutset(self.h, index, ub_cap, lb_filter, _ptr(states), _ptr(values), _ptr(ubs), C.byref(depth), C.byref(plen),
                                             paths, C.byref(count)), "ddo_mdd_drain_cutset")
        root = self._roots[index]
        out = []
        for i in range(count.value):
            p = list(root.path)
            if with_paths:
                p += [Decision(paths[i * plen.value + j].variable, paths[i * plen.value + j].value) for j in range(plen.value)]
            out.append(SubProblem(states[i].copy(), int(values[i]), p, int(ubs[i]), int(depth.value)))
        return out

    def _drain_frontier(self, index: int, ub_cap: int, lb_filter: int, with_paths: bool) -> List[SubProblem]:
        """FRONTIER cutset (clean.rs:586-606) of DD `index`: the nodes sit in different layers, so every record carries its own layer
        (ddo_mdd_drain_layer_index); canonical order = layer descending, position ascending."""
        n = len(self._last)
        cap = max(self._last[index].cutset_size, 1)
        W = self.problem.words
        caps = np.zeros(n, dtype=np.int64)
        lbs = np.full(n, N.I64_MAX, dtype=np.int64)
        caps[index] = ub_cap
        lbs[index] = lb_filter
        pwmax = (self.problem.nb_variables() + 64) // 64
        out = dict(states=np.zeros((cap, W), dtype=np.uint64), values=np.zeros(cap, dtype=np.int64), ubs=np.zeros(cap, dtype=np.int64),
                   dd=np.zeros(cap, dtype=np.int32), bits=np.zeros(cap * pwmax, dtype=np.uint64))
        total, pw = self.drain_cutset_batch(n, caps, lbs, out)
        layers = np.zeros(max(total, 1), dtype=np.int32)
        N.check(N.lib().ddo_mdd_drain_layer_index(self.h, _ptr(layers), len(layers)), "ddo_mdd_drain_layer_index")
        root = self._roots[index]
        vars_, _ = self.layer_trace(index) if (with_paths and total) else (None, None)
        bits = out["bits"][: total * pw].reshape(total, pw) if total else None
        res = []
        for i in range(total):
            tt = int(layers[i])
            p = list(root.path)
            if with_paths:  # terminal -> root order like clean.rs:329-343
                for t in range(tt - 1, -1, -1):
                    b = int((int(bits[i, t >> 6]) >> (t & 63)) & 1)
                    p.append(Decision(int(vars_[t]), (b if isinstance(self.problem, Misp) else 2 * b - 1)))  # MISP: YES = 1 / NO = 0; MAX2SAT: T = 1 / F = -1
            res.append(SubProblem(out["states"][i].copy(), int(out["values"][i]), p, int(out["ubs"][i]), root.depth + tt))
        return res

    def drain_cutset_batch(self, count: int, ub_caps, lb_filters, out=None):
        """Batched drain (ddo_mdd_drain_cutset_batch) into caller-provided numpy buffers `out` = dict(states, values, ubs, dd, bits).
        Returns (total, path_words)."""
        caps = np.asarray(ub_caps, dtype=np.int64)
        lbs = np.asarray(lb_filters, dtype=np.int64)
        o = out or {}
        total = C.c_int64(len(o["values"]) if "values" in o else (len(o["states"]) if "states" in o else 0))
        pw = C.c_int32(0)
        N.check(N.lib().ddo_mdd_drain_cutset_batch(self.h, count, _ptr(caps), _ptr(lbs), _ptr(o.get("states")), _ptr(o.get("values")), _ptr(o.get("ubs")),
                                                   _ptr(o.get("dd")), _ptr(o.get("bits")), C.byref(pw), C.byref(total)), "ddo_mdd_drain_cutset_batch")
        return int(total.value), int(pw.value)

    def set_profiling(self, on: bool):
        N.lib().ddo_mdd_set_profiling(self.h, int(on))

    def kernel_times(self):
        ms = (C.c_double * 6)()
        ln = (C.c_uint64 * 6)()
        N.lib().ddo_mdd_kernel_times(self.h, C.byref(ms), C.byref(ln))
        names = ["k_expand", "k_finish", "k_compact", "k_finalize_bottomup", "k_drain", "k_small"]  # MAX2SAT engine: slot "k_small" = merge kernels
        return {n: {"ms": ms[i], "launches": int(ln[i])} for i, n in enumerate(names)}

    def layer_trace(self, index: int = 0):
        n = self.problem.nb_variables() + 1
        v = np.zeros(n, dtype=np.int32)
        w = np.zeros(n, dtype=np.int32)
        ln = N.check(N.lib().ddo_mdd_layer_trace(self.h, index, _ptr(v), _ptr(w), n), "ddo_mdd_layer_trace")
        return v[:ln].copy(), w[:ln].copy()

    # -- device-resident variants (bench) ---------------------------------------------------------------------------
    def stage_roots(self, max_widths, residuals):
        n = len(residuals)
        widths = np.asarray(max_widths, dtype=np.uint64)
        states = np.ascontiguousarray(np.stack([np.asarray(r.state, dtype=np.uint64) for r in residuals]))
        values = np.asarray([r.value for r in residuals], dtype=np.int64)
        depths = np.asarray([r.depth for r in residuals], dtype=np.int32)
        N.check(N.lib().ddo_mdd_stage_roots(self.h, n, _ptr(widths), _ptr(states), _ptr(values), _ptr(depths)), "ddo_mdd_stage_roots")
        self._roots = list(residuals)

    def compile_staged(self, count: int, comp_type: int, best_lb: int = N.I64_MIN) -> float:
        ms = C.c_float(0)
        N.check(N.lib().ddo_mdd_compile_staged(self.h, count, comp_type, best_lb, C.byref(ms)), "ddo_mdd_compile_staged")
        return float(ms.value)

    def fetch_completions(self, count: int) -> List[Completion]:
        out = (N.Completion * count)()
        N.check(N.lib().ddo_mdd_fetch_completions(self.h, count, out), "ddo_mdd_fetch_completions")
        self._last = [_completion(c) for c in out]
        return self._last

    def close(self):
        if getattr(self, "h", None):
            N.lib().ddo_mdd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ParNoCachingSolverLel:
    """`ParNoCachingSolverLel` (solver/mod.rs:29-47; the solver examples/misp/main.rs:354 builds): branch-and-bound whose workers are the
    DDs of one device batch.  ``custom(problem, width, cutoff_seconds, wave_size)`` mirrors ``ParallelSolver::custom`` (parallel.rs:319-358)
    with ``nb_threads`` replaced by the number of DDs compiled in lock-step."""

    CUTSET_TYPE = LAST_EXACT_LAYER

    def __init__(self, problem: Misp, width, wave_size: int = 128, max_width_cap: Optional[int] = None, mdd: Optional[GpuMdd] = None,
                 batch_cap: Optional[int] = None):
        self.problem = problem
        kind, w, need = _width_spec(width, problem.nb_variables())
        cap = max_width_cap or need
        self.mdd = mdd or GpuMdd(problem, cap, min(wave_size, batch_cap) if batch_cap else wave_size, cutset_type=self.CUTSET_TYPE)
        h = C.c_void_p()
        N.check(N.lib().ddo_solver_create(problem.h, self.mdd.h, kind, w, wave_size, C.byref(h)), "ddo_solver_create")
        self.h = h
        self._is_exact = None

    @classmethod
    def custom(cls, problem, width, wave_size=128, **kw):
        return cls(problem, width, wave_size, **kw)

    # Solver::maximize, solver.rs:56
    def maximize(self, time_budget_s: float = 0.0, max_waves: int = 0):
        ex, has, val = C.c_int32(0), C.c_int32(0), C.c_int64(0)
        N.check(N.lib().ddo_solver_maximize(self.h, time_budget_s, max_waves, C.byref(ex), C.byref(has), C.byref(val)), "ddo_solver_maximize")
        self._is_exact = bool(ex.value)
        return Completion(bool(ex.value), int(val.value) if has.value else None, int(val.value) if has.value else None, 0, -1, 0,
                          int(self.stats()["expanded"]), int(self.stats()["transitions"]))

    # stepwise form (multi-GPU driver)
    def init(self, push_root: bool = True):
        N.check(N.lib().ddo_solver_init(self.h, int(push_root)), "ddo_solver_init")

    def wave(self, cutoff: Optional[np.ndarray] = None):
        out = (C.c_int64 * 3)()
        rc = N.check(N.lib().ddo_solver_wave(self.h, _ptr(cutoff) if cutoff is not None else None, C.byref(out)), "ddo_solver_wave")
        if rc == N.CUTOFF:
            raise CutoffOccurred()
        return int(out[0]), int(out[1]), int(out[2])

    def set_lower_bound(self, lb: int):
        N.lib().ddo_solver_set_lower_bound(self.h, lb)

    def retain_share(self, rank: int, nranks: int):
        N.check(N.lib().ddo_solver_retain_share(self.h, rank, nranks), "ddo_solver_retain_share")

    # work hand-off between ranks (ddo_b200/sharded.py): packed open nodes, ddo_solver_node_words() int64 words each
    def node_words(self) -> int:
        return int(N.lib().ddo_solver_node_words(self.h))

    def export_open(self, max_nodes: int) -> np.ndarray:
        rows = np.zeros((max(max_nodes, 1), self.node_words()), dtype=np.int64)
        cnt = C.c_int32(0)
        N.check(N.lib().ddo_solver_export_open(self.h, max_nodes, _ptr(rows), C.byref(cnt)), "ddo_solver_export_open")
        return rows[:cnt.value].copy()

    def import_open(self, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, self.node_words())
        if rows.shape[0]:
            N.check(N.lib().ddo_solver_import_open(self.h, rows.shape[0], _ptr(rows)), "ddo_solver_import_open")

    def maximize_sharded(self, comm, time_budget_s: float = 0.0, max_waves: int = 0, rebalance: bool = True):
        """`Solver::maximize` of a fringe-sharded search: ONE native call per rank (ddo_solver_maximize_sharded); `comm` is a
        ddo_b200.sharded.NativeComm.  Returns the dict ddo_b200.sharded.sharded_maximize returns."""
        out = (C.c_int64 * 8)()
        N.check(N.lib().ddo_solver_maximize_sharded(self.h, comm.h, time_budget_s, max_waves, int(rebalance), C.byref(out)), "ddo_solver_maximize_sharded")
        sol = self.best_solution()
        return {"best_lb": int(out[0]), "best_ub": int(out[1]), "is_exact": bool(out[2]), "waves": int(out[3]), "collectives": int(out[4]), "handoffs": int(out[5]),
                "nodes_sent": int(out[6]), "nodes_received": int(out[7]), "best_value": self.best_value(),
                "solution": None if sol is None else [(d.variable, d.value) for d in sol]}

    def finish(self):
        N.lib().ddo_solver_finish(self.h)

    def best_value(self):  # solver.rs:74
        has, val = C.c_int32(0), C.c_int64(0)
        N.lib().ddo_solver_best_value(self.h, C.byref(has), C.byref(val))
        return int(val.value) if has.value else None

    def best_solution(self):  # solver.rs:71
        if self.best_value() is None:
            return None
        cap = self.problem.nb_variables() + 1
        buf = (N.Decision * cap)()
        ln = C.c_int32(cap)
        N.check(N.lib().ddo_solver_best_solution(self.h, buf, C.byref(ln)), "ddo_solver_best_solution")
        return [Decision(buf[i].variable, buf[i].value) for i in range(ln.value)]

    def best_lower_bound(self) -> int:  # solver.rs:83
        return int(N.lib().ddo_solver_best_lower_bound(self.h))

    def best_upper_bound(self) -> int:  # solver.rs:86
        return int(N.lib().ddo_solver_best_upper_bound(self.h))

    def gap(self) -> float:  # solver.rs:80-93
        return solver_gap(self.best_lower_bound(), self.best_upper_bound())

    def explored(self) -> int:  # solver.rs:96
        return int(N.lib().ddo_solver_explored(self.h))

    def fringe_len(self) -> int:
        return int(N.lib().ddo_solver_fringe_len(self.h))

    def stats(self):
        s = (C.c_double * 8)()
        N.lib().ddo_solver_stats(self.h, C.byref(s))
        return dict(expanded=s[0], transitions=s[1], compilations=s[2], waves=s[3], device_ms=s[4], fringe_ms=s[5], bytes_h2d=s[6], bytes_d2h=s[7])

    def close(self):
        if getattr(self, "h", None):
            N.lib().ddo_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ParNoCachingSolverFc(ParNoCachingSolverLel):
    """`ParNoCachingSolverFc` (solver/mod.rs:33): the same solver over DDs with the FRONTIER cutset (clean.rs:586-606), both device models.
    A FRONTIER engine keeps 12 B per (layer, position) per DD slot on top of the engine's arenas: pass ``batch_cap`` for wide DDs
    (e.g. 512 slots at W = 10 000, n = 500 are 33 + 31 GB)."""
    CUTSET_TYPE = FRONTIER


DefaultSolver = ParNoCachingSolverLel  # solver/mod.rs:29
