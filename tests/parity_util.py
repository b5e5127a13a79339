"""Shared helpers of the GPU parity tests: run the same compilation through the C ABI and through the CPU oracle and compare everything
the DecisionDiagram trait exposes (mdd.rs:75-114) plus the per-layer trace."""
from __future__ import annotations

import numpy as np

import oracle_lib as O
from ddo_b200 import CompilationType, GpuMdd, Max2Sat, Misp, SubProblem
from ddo_b200 import _native as N

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
