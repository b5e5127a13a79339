"""Shared helpers of the GPU parity tests: run the same compilation through the C ABI and through the CPU oracle and compare everything
the DecisionDiagram trait exposes (mdd.rs:75-114) plus the per-layer trace."""
from __future__ import annotations

import numpy as np

import oracle_lib as O
from ddo_b200 import CompilationType, GpuMdd, Max2Sat, Misp, SubProblem
from ddo_b200 import _native as N

def _sha(*arrays) -> str:
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def oracle_digest(ref: dict, comp_type: int) -> dict:
    """Digest of one oracle DD (tests/golden/make_dd_goldens.py); device_digest() builds the same from the C ABI's outputs."""
    d = {"has_best": int(ref["has_best"]), "best_value": int(ref["best_value"]) if ref["has_best"] else None, "is_exact": int(ref["is_exact"]),
         "has_best_exact": int(ref["has_best_exact"]), "best_exact_value": int(ref["best_exact_value"]) if ref["has_best_exact"] else None,
         "expanded": int(ref["expanded"]), "transitions": int(ref["transitions"]), "lel": int(ref["lel"]), "n_layers": int(len(ref["layer_vars"])),
         "layers_sha": _sha(np.asarray(ref["layer_vars"], dtype=np.int32), np.asarray(ref["layer_widths"], dtype=np.int32))}
    if comp_type == O.RELAXED:
        d["cutset_size"] = int(ref["cutset_size"])
        d["cutset_sha"] = _sha(np.asarray(ref["cutset_states"], dtype=np.uint64), np.asarray(ref["cutset_values"], dtype=np.int64),
                               np.asarray(ref["cutset_ubs"], dtype=np.int64), np.asarray(ref["cutset_depths"], dtype=np.int32))
    sol = ref["best_exact_solution"]
    d["best_exact_solution_sha"] = None if sol is None else _sha(np.asarray(sol, dtype=np.int32))
    return d


def device_digest(mdd: GpuMdd, index: int, comp_type: int, root_depth: int = 0) -> dict:
    c = mdd._last[index]
    v, w = mdd.layer_trace(index)
    d = {"has_best": int(c.best_value is not None), "best_value": c.best_value, "is_exact": int(c.is_exact),
         "has_best_exact": int(c.best_exact_value is not None), "best_exact_value": c.best_exact_value,
         "expanded": int(c.expanded), "transitions": int(c.transitions), "lel": (c.lel_depth - root_depth) if c.lel_depth >= 0 else -1,
         "n_layers": int(len(v)), "layers_sha": _sha(np.asarray(v, dtype=np.int32), np.asarray(w, dtype=np.int32))}
    if comp_type == O.RELAXED:
        cs = mdd.drain_cutset(index, with_paths=False)
        d["cutset_size"] = len(cs)
        words = len(cs[0].state) if cs else 0
        d["cutset_sha"] = _sha(np.asarray([sp.state for sp in cs], dtype=np.uint64).reshape(len(cs), words) if cs else np.zeros((0, 0), dtype=np.uint64),
                               np.asarray([sp.value for sp in cs], dtype=np.int64), np.asarray([sp.ub for sp in cs], dtype=np.int64),
                               np.asarray([sp.depth for sp in cs], dtype=np.int32))
    sol = mdd.best_exact_solution(index) if c.best_exact_value is not None else None
    d["best_exact_solution_sha"] = None if sol is None else _sha(np.asarray([(x.variable, x.value) for x in sol], dtype=np.int32))
    return d


CT = {O.EXACT: CompilationType.Exact, O.RELAXED: CompilationType.Relaxed, O.RESTRICTED: CompilationType.Restricted}


def compare_dd(oracle: O.OracleMisp, mdd: GpuMdd, index: int, comp_type: int, width: int, root_state, root_value, root_depth, best_lb, check_paths=True,
               cutset_type=O.LEL):
    """Assert device DD `index` of the last batch == oracle DD on the same CompilationInput. Returns the oracle result."""
    ref = oracle.compile(comp_type, width, root_state, root_value, root_depth, best_lb, cutset_type=cutset_type, want_paths=check_paths)
    c = mdd._last[index]
    ctx = f"comp={comp_type} W={width} depth={root_depth} lb={best_lb}"
    assert (c.best_value is not None) == bool(ref["has_best"]), ctx
    if ref["has_best"]:
        assert c.best_value == ref["best_value"], ctx
    assert c.is_exact == bool(ref["is_exact"]), ctx
    assert (c.best_exact_value is not None) == bool(ref["has_best_exact"]), ctx
    if ref["has_best_exact"]:
        assert c.best_exact_value == ref["best_exact_value"], ctx
    assert c.expanded == ref["expanded"], (ctx, c.expanded, ref["expanded"])
    assert c.transitions == ref["transitions"], ctx
    v, w = mdd.layer_trace(index)
    assert v.tolist() == ref["layer_vars"].tolist(), ctx
    assert w.tolist() == ref["layer_widths"].tolist(), ctx
    assert c.lel_depth == (root_depth + ref["lel"] if ref["lel"] >= 0 else -1), ctx
    if comp_type == O.RELAXED:
        assert c.cutset_size == ref["cutset_size"], (ctx, c.cutset_size, ref["cutset_size"])
        cs = mdd.drain_cutset(index, with_paths=check_paths)
        assert len(cs) == ref["cutset_size"]
        for i, sp in enumerate(cs):  # identical ORDER, states, values, bounds, depths (exact-cutset node set, bit-exact)
            assert np.array_equal(sp.state, ref["cutset_states"][i]), (ctx, i)
            assert sp.value == ref["cutset_values"][i] and sp.ub == ref["cutset_ubs"][i] and sp.depth == ref["cutset_depths"][i], (ctx, i)
            if check_paths:
                assert [(d.variable, d.value) for d in sp.path] == [tuple(x) for x in ref["cutset_paths"][i].tolist()], (ctx, i)
    # solutions: exact ones are compared decision by decision; the best path of an inexact relaxed DD only by its value
    if ref["best_exact_solution"] is not None:
        sol = mdd.best_exact_solution(index)
        assert [(d.variable, d.value) for d in sol] == ref["best_exact_solution"], ctx
    if ref["best_solution"] is not None and (comp_type != O.RELAXED or ref["is_exact"]):
        sol = mdd.best_solution(index)
        assert [(d.variable, d.value) for d in sol] == ref["best_solution"], ctx
    return ref


def check_instance(inst, widths, comp_types=(O.RESTRICTED, O.RELAXED), best_lbs=(N.I64_MIN,), batch=None, roots=None, check_paths=True, model="misp",
                   cutset_type=O.LEL):
    """Compile the given roots (default: the problem root) for every width / type / best_lb, batched on the device, and compare each DD."""
    oracle = O.OracleMisp(inst) if model == "misp" else O.OracleM2s(inst)
    pb = Misp(inst) if model == "misp" else Max2Sat(inst)
    roots = roots or [SubProblem(inst.initial_state(), pb.initial_value(), [], N.I64_MAX, 0)]
    jobs = [(w, r) for w in widths for r in roots]
    mdd = GpuMdd(pb, max(max(widths), 2), batch or len(jobs), cutset_type=cutset_type)
    n = 0
    try:
        for ct in comp_types:
            for lb in best_lbs:
                for s in range(0, len(jobs), mdd.batch_cap):
                    chunk = jobs[s : s + mdd.batch_cap]
                    mdd.compile_batch(CT[ct], [w for w, _ in chunk], [r for _, r in chunk], lb)
                    for i, (w, r) in enumerate(chunk):
                        compare_dd(oracle, mdd, i, ct, w, r.state, r.value, r.depth, lb, check_paths, cutset_type)
                        n += 1
    finally:
        mdd.close()
        pb.close()
    return n
