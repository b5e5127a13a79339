"""GPU parity tests of the MAX2SAT device model (BASELINE config 3): the CUDA path through the C ABI against the CPU oracle, bit-exact
(integer work).  The oracle's MAX2SAT model is pinned to the reference's known optima and unit vectors in tests/test_oracle_golden.py."""
import json

import numpy as np
import pytest

import oracle_lib as O
from ddo_b200 import CompilationType, FixedWidth, GpuMdd, Max2Sat, NbUnassignedWidth, ParNoCachingSolverLel, SubProblem, random_max2sat, read_wcnf
from ddo_b200 import _native as N
from parity_util import check_instance

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,seed", [(3, 4, 1), (9, 30, 2), (16, 60, 3), (33, 150, 4), (61, 400, 5), (130, 900, 6)])
def test_dd_parity_random_instances_all_widths(n, m, seed):
    """Every observable of restricted / relaxed DDs (values, exactness, expanded / transition counts, layer widths, LEL depth, cutset order /
    states / values / upper bounds / paths, exact solutions) over a sweep of widths and incumbents."""
    inst = random_max2sat(n, m, seed)
    widths = [1, 2, 3, 5, 8, 13, 50, 300]
    lo = O.OracleM2s(inst).compile(O.RESTRICTED, 3)["best_value"]
    cnt = check_instance(inst, widths, best_lbs=(N.I64_MIN, lo - 3, lo + 2), model="m2s")
    assert cnt == len(widths) * 2 * 3


def test_dd_parity_duplicate_states_and_ranking_ties():
    """Few distinct weights -> many equal states (dedup, value_top = max, best parent) and many equal (value_top, rank) keys at the cut
    (canonical lexicographic tie-break)."""
    for seed in (7, 8, 9):
        inst = random_max2sat(24, 40, seed, max_weight=1)
        check_instance(inst, [1, 2, 4, 7, 16, 64, 1000], best_lbs=(N.I64_MIN,), model="m2s")


def test_dd_parity_reference_fixtures(golden_dir):
    """The reference's own instances (resources/max2sat): negative weights, unit clauses, tautologies, and a 60-variable frb instance."""
    for name, widths in (("debug", [1, 2, 8]), ("debug2", [1, 2, 8]), ("pass", [1, 2, 3, 50]), ("tautology", [1, 4]), ("unit", [1, 4]),
                         ("negative_wt", [1, 4]), ("frb10-6-1", [1, 4, 25, 150])):
        inst = read_wcnf(golden_dir / "max2sat" / f"{name}.wcnf")
        check_instance(inst, widths, model="m2s")


def test_dd_parity_exact_compilation():
    inst = random_max2sat(11, 40, 12)
    check_instance(inst, [4096], comp_types=(O.EXACT,), model="m2s")


def test_dd_parity_subproblem_roots_deep_in_the_search():
    """Roots taken from a real cutset (non-zero value, depth, partially assigned benefit vectors), compiled as one batch."""
    inst = random_max2sat(70, 500, 21)
    oracle = O.OracleM2s(inst)
    ref = oracle.compile(O.RELAXED, 12)
    roots = [SubProblem(ref["cutset_states"][i].copy(), int(ref["cutset_values"][i]), [], int(ref["cutset_ubs"][i]), int(ref["cutset_depths"][i]))
             for i in range(ref["cutset_size"])]
    assert len(roots) >= 5
    lo = oracle.compile(O.RESTRICTED, 5)["best_value"]
    check_instance(inst, [3, 12], roots=roots, best_lbs=(N.I64_MIN, lo), check_paths=False, model="m2s")


@pytest.mark.parametrize("name", ["debug", "debug2", "pass", "tautology", "unit", "negative_wt"])
def test_solver_known_optima_small(golden_dir, name):
    exp = json.loads((golden_dir / "expected.json").read_text())["max2sat"][name]["optimum"]
    inst = read_wcnf(golden_dir / "max2sat" / f"{name}.wcnf")
    for width in (NbUnassignedWidth(inst.n), FixedWidth(1), FixedWidth(2)):
        s = ParNoCachingSolverLel(Max2Sat(inst), width, wave_size=4)
        comp = s.maximize()
        assert comp.is_exact and comp.best_value == exp, (name, width)
        assert s.best_lower_bound() == s.best_upper_bound() == exp


@pytest.mark.parametrize("name,width,K", [("frb10-6-1", 100, 16), ("frb10-6-3", 40, 32)])
def test_solver_known_optimum_and_wave_trace_parity(golden_dir, name, width, K):
    """Solver::maximize on the reference's frb10-6-* fixtures: asserted optimum (max2sat/tests.rs:92-106), proven bound, and the exact same
    branch-and-bound trajectory as the oracle's wave solver (explored / expanded / transitions / compilations / waves, solution)."""
    exp = json.loads((golden_dir / "expected.json").read_text())["max2sat"][name]["optimum"]
    inst = read_wcnf(golden_dir / "max2sat" / f"{name}.wcnf")
    s = ParNoCachingSolverLel(Max2Sat(inst), FixedWidth(width), wave_size=K)
    comp = s.maximize()
    assert comp.is_exact and comp.best_value == exp
    assert s.best_lower_bound() == s.best_upper_bound() == exp
    ref = O.OracleM2s(inst).solve("wave", k=K, width=width)
    st = s.stats()
    assert ref["best_value"] == exp
    assert (s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"]), int(st["waves"])) == \
           (ref["explored"], ref["expanded"], ref["transitions"], ref["compilations"], ref["waves"])
    sol = [(d.variable, d.value) for d in s.best_solution()]
    assert sol == ref["solution"]
    model = dict(sol)
    uniq = {}
    for w, x, y in inst.clauses.tolist():
        uniq[(min(x, y), max(x, y))] = w
    assert sum(w for (x, y), w in uniq.items() if model[abs(x) - 1] * x > 0 or model[abs(y) - 1] * y > 0) == exp


def test_full_size_config3_root_dds_and_properties():
    """BASELINE config 3 at full size: 500 variables / 3000 clauses, W = 5000.  The oracle cannot finish a DD of this size in test time,
    so: size-independent properties of the root restricted / relaxed DDs (restricted value <= relaxed bound; the restricted solution's
    clause weight equals its value; cutset upper bounds bounded by the DD bound; cutset values consistent with their paths), and
    bit-exact parity with the oracle on the same instance at W = 8 for the first layers (cutoff by depth is not available, so a narrow DD)."""
    inst = random_max2sat(500, 3000, 1)
    pb = Max2Sat(inst)
    W = 5000
    mdd = GpuMdd(pb, W, 1)
    root = SubProblem(pb.initial_state(), pb.initial_value())
    lo = mdd.compile(CompilationType.Restricted, W, root)
    sol = mdd.best_exact_solution(0) if lo.best_exact_value is not None else mdd.best_solution(0)
    model = {d.variable: d.value for d in sol}
    assert len(model) == 500
    uniq = {}
    for w, x, y in inst.clauses.tolist():
        uniq[(min(x, y), max(x, y))] = w
    sat = sum(w for (x, y), w in uniq.items() if model[abs(x) - 1] * x > 0 or model[abs(y) - 1] * y > 0)
    assert sat == lo.best_value
    hi = mdd.compile(CompilationType.Relaxed, W, root, best_lb=lo.best_value)
    assert hi.best_value >= lo.best_value and not hi.is_exact and lo.expanded > 1_000_000
    cs = mdd.drain_cutset(0, lb_filter=lo.best_value)
    assert 0 < len(cs) <= W
    total = sum(abs(w) for w in uniq.values())
    for sp in cs[:50]:
        assert lo.best_value < sp.ub <= hi.best_value and sp.value <= total and sp.depth == len(sp.path)
        b = pb.unpack_state(sp.state)
        assigned = [d.variable for d in sp.path]
        assert np.all(b[assigned] == 0)
    mdd.close()
    check_instance(inst, [8], model="m2s", check_paths=False)


def test_full_size_config3_root_dds_bit_exact(golden_dir):
    """BASELINE config 3 at full size, bit-exact: MAX2SAT 500 variables / 3000 clauses seed 1, W = 5000 -- the restricted root DD and the
    relaxed root DD (compiled against the restricted value) against the oracle's digests (tests/golden/dd_digests.json, made by
    tests/golden/make_dd_goldens.py in ~3 minutes of CPU): every scalar, the per-layer trace, the cutset and the best exact solution."""
    import json
    from parity_util import device_digest
    g = json.loads((golden_dir / "dd_digests.json").read_text())["config3_max2sat_n500_c3000_seed1_w5000"]
    inst = random_max2sat(500, 3000, 1)
    pb = Max2Sat(inst)
    mdd = GpuMdd(pb, g["width"], 1)
    root = SubProblem(pb.initial_state(), pb.initial_value())
    mdd.compile(CompilationType.Restricted, g["width"], root)
    assert device_digest(mdd, 0, O.RESTRICTED) == g["restricted"]
    mdd.compile(CompilationType.Relaxed, g["width"], root, best_lb=g["relaxed_best_lb"])
    assert device_digest(mdd, 0, O.RELAXED) == g["relaxed"]
    mdd.close()
