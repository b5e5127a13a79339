"""GPU parity tests of the FRONTIER cutset (clean.rs:586-606): the device sweeps (ddo_b200/csrc/frontier.cuh, m2s_frontier.cuh) against the CPU
oracle, bit-exact, for both device models.

Compared per relaxed DD: the frontier node set in canonical order (layer descending, position ascending), every node's state (re-derived on
the device by replaying its best path), value_top, upper bound min(value_top + rub, value_top + value_bot, best_value), depth and path;
per solver run: the optimum, the proven bound and the whole branch-and-bound trajectory of `ParNoCachingSolverFc` (solver/mod.rs:33)."""
import json

import numpy as np
import pytest

import oracle_lib as O
from ddo_b200 import CompilationType, FixedWidth, GpuMdd, Max2Sat, Misp, NbUnassignedWidth, ParNoCachingSolverFc, SubProblem, gnp, parse_dimacs, random_max2sat, read_wcnf
from ddo_b200 import _native as N
from parity_util import check_instance

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,seed", [(12, 0.3, 1), (30, 0.5, 2), (64, 0.2, 3), (65, 0.5, 4), (100, 0.1, 5), (129, 0.3, 7), (200, 0.5, 8)])
def test_frontier_dd_parity_random_graphs_all_widths(n, p, seed):
    inst = gnp(n, p, seed)
    widths = [1, 2, 3, 5, 8, 13, 50, 400]
    cnt = check_instance(inst, widths, best_lbs=(N.I64_MIN, 2, 6), cutset_type=O.FRONTIER)
    assert cnt == len(widths) * 2 * 3


def test_frontier_nodes_come_from_several_layers():
    """The point of the frontier cutset: its nodes sit in different layers (a LEL cutset is one layer), none below the first merge."""
    inst = gnp(100, 0.1, 5)
    pb = Misp(inst)
    mdd = GpuMdd(pb, 8, 1, cutset_type=N.FRONTIER)
    c = mdd.compile(CompilationType.Relaxed, 8, SubProblem(pb.initial_state(), 0))
    cs = mdd.drain_cutset(0)
    assert not c.is_exact and len(cs) == c.cutset_size > 0
    depths = [sp.depth for sp in cs]
    assert len(set(depths)) > 1 and depths == sorted(depths, reverse=True)
    assert all(len(sp.path) == sp.depth for sp in cs)
    # the solver-side filter of parallel.rs:460-461 applied by the drain
    lb = sorted(sp.ub for sp in cs)[len(cs) // 2]
    kept = mdd.drain_cutset(0, lb_filter=lb)
    assert [sp.ub for sp in kept] == [sp.ub for sp in cs if sp.ub > lb]
    capped = mdd.drain_cutset(0, ub_cap=lb)
    assert [sp.ub for sp in capped] == [min(sp.ub, lb) for sp in cs]
    with pytest.raises(N.DdoError) as e:  # the single-DD ABI call reports one depth per cutset
        N.check(N.lib().ddo_mdd_drain_cutset(mdd.h, 0, N.I64_MAX, N.I64_MIN, None, None, None, None, None, None, None), "ddo_mdd_drain_cutset")
    mdd.close()
    pb.close()


def test_frontier_dd_parity_weighted_instance():
    inst = gnp(90, 0.3, 11)
    rng = np.random.default_rng(5)
    inst.weights[:] = rng.integers(1, 50, size=inst.n)
    check_instance(inst, [1, 2, 4, 9, 30, 200], best_lbs=(N.I64_MIN, 100), cutset_type=O.FRONTIER)


def test_frontier_dd_parity_subproblem_roots():
    """Roots taken from a real frontier cutset (different depths, non-zero values), compiled as one batch."""
    inst = gnp(120, 0.4, 21)
    oracle = O.OracleMisp(inst)
    ref = oracle.compile(O.RELAXED, 20, cutset_type=O.FRONTIER)
    roots = [SubProblem(ref["cutset_states"][i].copy(), int(ref["cutset_values"][i]), [], int(ref["cutset_ubs"][i]), int(ref["cutset_depths"][i]))
             for i in range(min(ref["cutset_size"], 24))]
    assert len(roots) >= 5 and len({r.depth for r in roots}) > 1
    check_instance(inst, [3, 20], roots=roots, best_lbs=(N.I64_MIN, 10), check_paths=False, cutset_type=O.FRONTIER)


def test_frontier_dd_parity_wide_states_and_layers():
    """n = 500 (BASELINE config 2's state size, 8 words) and n = 1000 (config 5, 16 words) at widths the oracle finishes in seconds."""
    check_instance(gnp(500, 0.5, 1), [100, 1000], comp_types=(O.RELAXED,), cutset_type=O.FRONTIER)
    check_instance(gnp(1000, 0.5, 1), [300], comp_types=(O.RELAXED,), check_paths=False, cutset_type=O.FRONTIER)


@pytest.mark.parametrize("name", ["johnson8-2-4", "hamming6-4", "MANN_a9", "johnson8-4-4", "c-fat200-5", "brock200_2"])
def test_frontier_solver_known_optima_and_trajectory(golden_dir, name):
    """ParNoCachingSolverFc on the reference's DIMACS fixtures: asserted optimum (misp/tests.rs) and the same trajectory as the oracle."""
    exp = json.loads((golden_dir / "expected.json").read_text())["misp"][name]["optimum"]
    inst = parse_dimacs((golden_dir / "misp" / f"{name}.clq").read_text(), name)
    pb = Misp(inst)
    K = 16
    s = ParNoCachingSolverFc(pb, NbUnassignedWidth(inst.n), wave_size=K)
    comp = s.maximize()
    assert comp.is_exact and comp.best_value == exp
    assert s.best_lower_bound() == s.best_upper_bound() == exp
    ref = O.OracleMisp(inst).solve("wave", k=K, cutset_type=O.FRONTIER)
    st = s.stats()
    assert (s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"]), int(st["waves"])) == \
           (ref["explored"], ref["expanded"], ref["transitions"], ref["compilations"], ref["waves"])
    sol = sorted(d.variable for d in s.best_solution() if d.value == 1)
    assert sol == ref["solution"]
    adj = set(zip(inst.src.tolist(), inst.dst.tolist())) | set(zip(inst.dst.tolist(), inst.src.tolist()))
    assert all((a, b) not in adj for a in sol for b in sol if a != b) and len(sol) == exp


def test_frontier_solver_fixed_width_matches_oracle_trace():
    inst = gnp(120, 0.4, 31)
    pb = Misp(inst)
    s = ParNoCachingSolverFc(pb, FixedWidth(10), wave_size=32)
    comp = s.maximize()
    ref = O.OracleMisp(inst).solve("wave", k=32, width=10, cutset_type=O.FRONTIER)
    assert comp.is_exact and comp.best_value == ref["best_value"]
    assert (s.explored(), int(s.stats()["expanded"])) == (ref["explored"], ref["expanded"])
    sol = sorted(d.variable for d in s.best_solution() if d.value == 1)
    assert sol == ref["solution"]


# ---- MAX2SAT device model (m2s_frontier.cuh) ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,m,seed", [(3, 4, 1), (9, 30, 2), (16, 60, 3), (33, 150, 4), (61, 400, 5), (130, 900, 6)])
def test_frontier_max2sat_dd_parity_all_widths(n, m, seed):
    inst = random_max2sat(n, m, seed)
    widths = [1, 2, 3, 5, 8, 13, 50, 300]
    lo = O.OracleM2s(inst).compile(O.RESTRICTED, 3)["best_value"]
    cnt = check_instance(inst, widths, best_lbs=(N.I64_MIN, lo - 3, lo + 2), model="m2s", cutset_type=O.FRONTIER)
    assert cnt == len(widths) * 2 * 3


def test_frontier_max2sat_ties_and_reference_fixtures(golden_dir):
    for seed in (7, 8):
        check_instance(random_max2sat(24, 40, seed, max_weight=1), [1, 2, 4, 7, 16, 64], model="m2s", cutset_type=O.FRONTIER)
    for name, widths in (("debug2", [1, 2, 8]), ("pass", [1, 2, 3, 50]), ("negative_wt", [1, 4]), ("frb10-6-1", [1, 4, 25, 150])):
        check_instance(read_wcnf(golden_dir / "max2sat" / f"{name}.wcnf"), widths, model="m2s", cutset_type=O.FRONTIER)


def test_frontier_max2sat_subproblem_roots():
    inst = random_max2sat(70, 500, 21)
    oracle = O.OracleM2s(inst)
    ref = oracle.compile(O.RELAXED, 12, cutset_type=O.FRONTIER)
    roots = [SubProblem(ref["cutset_states"][i].copy(), int(ref["cutset_values"][i]), [], int(ref["cutset_ubs"][i]), int(ref["cutset_depths"][i]))
             for i in range(min(ref["cutset_size"], 24))]
    assert len(roots) >= 5 and len({r.depth for r in roots}) > 1
    lo = oracle.compile(O.RESTRICTED, 5)["best_value"]
    check_instance(inst, [3, 12], roots=roots, best_lbs=(N.I64_MIN, lo), check_paths=False, model="m2s", cutset_type=O.FRONTIER)


def test_frontier_max2sat_full_state_size():
    """BASELINE config 3's state size (500 variables, 2 KB rows) at a width the oracle finishes in seconds."""
    check_instance(random_max2sat(500, 3000, 1), [8], comp_types=(O.RELAXED,), model="m2s", check_paths=False, cutset_type=O.FRONTIER)


@pytest.mark.parametrize("n,m,seed,width,K", [(25, 100, 3, 4, 8), (35, 200, 5, 10, 16), (40, 250, 6, 16, 16)])
def test_frontier_max2sat_solver_trajectory(n, m, seed, width, K):
    inst = random_max2sat(n, m, seed)
    s = ParNoCachingSolverFc(Max2Sat(inst), FixedWidth(width), wave_size=K)
    comp = s.maximize()
    ref = O.OracleM2s(inst).solve("wave", k=K, width=width, cutset_type=O.FRONTIER)
    st = s.stats()
    assert comp.is_exact and comp.best_value == ref["best_value"] == O.OracleM2s(inst).solve("wave", k=K, width=width)["best_value"]
    assert s.best_lower_bound() == s.best_upper_bound() == ref["best_value"]
    assert (s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"]), int(st["waves"])) == \
           (ref["explored"], ref["expanded"], ref["transitions"], ref["compilations"], ref["waves"])
    assert [(d.variable, d.value) for d in s.best_solution()] == ref["solution"]
