"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit-exact (integer / bitset work)."""
import json

import numpy as np
import pytest

import oracle_lib as O
from ddo_b200 import CompilationType, FixedWidth, GpuMdd, Misp, NbUnassignedWidth, ParNoCachingSolverLel, SubProblem, gnp, parse_dimacs
from ddo_b200 import _native as N
from parity_util import check_instance, compare_dd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,seed", [(12, 0.3, 1), (30, 0.5, 2), (64, 0.2, 3), (65, 0.5, 4), (100, 0.1, 5), (128, 0.7, 6), (129, 0.3, 7), (200, 0.5, 8)])
def test_dd_parity_random_graphs_all_widths(n, p, seed):
    """Every observable of restricted / relaxed DDs over a sweep of widths (1 = everything merged ... wide = exact), several best_lb."""
    inst = gnp(n, p, seed)
    widths = [1, 2, 3, 5, 8, 13, 50, 400]
    cnt = check_instance(inst, widths, best_lbs=(N.I64_MIN, 2, 6))
    assert cnt == len(widths) * 2 * 3


def test_dd_parity_weighted_instance():
    inst = gnp(90, 0.3, 11)
    rng = np.random.default_rng(5)
    inst.weights[:] = rng.integers(1, 50, size=inst.n)
    check_instance(inst, [1, 2, 4, 9, 30, 200], best_lbs=(N.I64_MIN, 100))


def test_dd_parity_exact_compilation_and_capacity_error():
    inst = gnp(40, 0.6, 12)
    check_instance(inst, [4000], comp_types=(O.EXACT,))
    pb = Misp(gnp(60, 0.1, 13))
    mdd = GpuMdd(pb, 8, 1)
    with pytest.raises(N.DdoError) as e:
        mdd.compile(CompilationType.Exact, 8, SubProblem(pb.initial_state(), 0))
    assert e.value.code == N.ERR_CAPACITY


def test_dd_parity_subproblem_roots_deep_in_the_search():
    """Roots taken from a real cutset (non-zero value, depth, partial states), compiled as one batch."""
    inst = gnp(120, 0.4, 21)
    oracle = O.OracleMisp(inst)
    ref = oracle.compile(O.RELAXED, 20)
    roots = [SubProblem(ref["cutset_states"][i].copy(), int(ref["cutset_values"][i]), [], int(ref["cutset_ubs"][i]), int(ref["cutset_depths"][i]))
             for i in range(ref["cutset_size"])]
    assert len(roots) >= 5
    check_instance(inst, [3, 20], roots=roots, best_lbs=(N.I64_MIN, 10), check_paths=False)


def test_error_behaviour_matches_reference():
    pb = Misp(gnp(30, 0.5, 3))
    mdd = GpuMdd(pb, 16, 2)
    root = SubProblem(pb.initial_state(), 0)
    with pytest.raises(N.DdoError):  # max_width == 0 panics in the reference (clean.rs:827)
        mdd.compile(CompilationType.Relaxed, 0, root)
    with pytest.raises(N.DdoError):  # width above the arena
        mdd.compile(CompilationType.Restricted, 17, root)
    flag = np.ones(1, dtype=np.int32)
    from ddo_b200 import CutoffOccurred
    with pytest.raises(CutoffOccurred):  # Err(Reason::CutoffOccurred), clean.rs:352-354
        mdd.compile(CompilationType.Restricted, 4, root, cutoff=flag)
    # infeasible / fully pruned: best_lb above every bound -> Ok(Completion{best_value: None}) (clean.rs:1669-1749)
    c = mdd.compile(CompilationType.Relaxed, 4, root, best_lb=1000)
    assert c.best_value is None and c.expanded == 0
    with pytest.raises(N.DdoError):  # only LAST_EXACT_LAYER = 1 and FRONTIER = 2 exist (mdd.rs:24-28; the reference panics at clean.rs:559)
        GpuMdd(pb, 16, 1, cutset_type=3)


FAST = ["johnson8-2-4", "hamming6-4", "hamming6-2", "MANN_a9", "johnson8-4-4", "c-fat200-5"]


@pytest.mark.parametrize("name", FAST + ["brock200_2", "hamming8-2", "c-fat500-1", "c-fat500-2"])
def test_solver_known_optima_and_wave_trace_parity(golden_dir, name):
    """Solver::maximize on the reference's DIMACS fixtures: asserted optimum (misp/tests.rs), proven bound, and the exact same
    branch-and-bound trajectory as the oracle's wave solver (explored / expanded / transitions / compilations)."""
    exp = json.loads((golden_dir / "expected.json").read_text())["misp"][name]["optimum"]
    inst = parse_dimacs(O.read_clq(golden_dir, name), name)
    pb = Misp(inst)
    K = 16
    s = ParNoCachingSolverLel(pb, NbUnassignedWidth(inst.n), wave_size=K)
    comp = s.maximize()
    assert comp.is_exact and comp.best_value == exp
    assert s.best_lower_bound() == s.best_upper_bound() == exp
    ref = O.OracleMisp(inst).solve("wave", k=K)
    st = s.stats()
    assert (s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"]), int(st["waves"])) == \
           (ref["explored"], ref["expanded"], ref["transitions"], ref["compilations"], ref["waves"])
    sol = sorted(d.variable for d in s.best_solution() if d.value == 1)
    assert sol == ref["solution"]
    adj = set(zip(inst.src.tolist(), inst.dst.tolist())) | set(zip(inst.dst.tolist(), inst.src.tolist()))
    assert all((a, b) not in adj for a in sol for b in sol if a != b) and len(sol) == exp


def test_solver_fixed_width_matches_oracle_trace():
    inst = gnp(150, 0.3, 31)
    pb = Misp(inst)
    s = ParNoCachingSolverLel(pb, FixedWidth(10), wave_size=32)
    comp = s.maximize()
    ref = O.OracleMisp(inst).solve("wave", k=32, width=10)
    assert comp.is_exact and comp.best_value == ref["best_value"]
    assert (s.explored(), int(s.stats()["expanded"])) == (ref["explored"], ref["expanded"])


def test_full_size_config2_solver_trajectory_matches_oracle_golden(golden_dir):
    """BASELINE config 2 solved to proven optimality: objective, bound and the whole branch-and-bound trajectory (sub-problems explored,
    nodes expanded, transitions, compilations, waves) equal the CPU oracle's wave solver, whose 15-minute run is committed as
    tests/golden/config2_trajectory_k512.json (made by tests/golden/make_trajectory.py)."""
    g = json.loads((golden_dir / "config2_trajectory_k512.json").read_text())
    inst = gnp(500, 0.5, 1)
    pb = Misp(inst)
    s = ParNoCachingSolverLel(pb, FixedWidth(g["width"]), wave_size=g["wave_size"])
    comp = s.maximize()
    st = s.stats()
    assert comp.is_exact and comp.best_value == g["best_value"] == 13
    assert s.best_lower_bound() == g["best_lb"] and s.best_upper_bound() == g["best_ub"]
    got = (s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"]), int(st["waves"]))
    assert got == (g["explored"], g["expanded"], g["transitions"], g["compilations"], g["waves"])
    sol = sorted(d.variable for d in s.best_solution() if d.value == 1)
    assert sol == g["solution"]


def test_small_dd_fast_path_equals_general_engine(monkeypatch):
    """The shared-memory fast path (k_small) and the general layer-by-layer engine give the same search: same optimum, bounds and counters."""
    inst = gnp(220, 0.4, 17)
    out = []
    for ws in ("256", "0"):
        monkeypatch.setenv("DDO_SMALL_WS", ws)
        s = ParNoCachingSolverLel(Misp(inst), FixedWidth(300), wave_size=64)
        c = s.maximize()
        st = s.stats()
        out.append((c.best_value, c.is_exact, s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"]), int(st["waves"]),
                    sorted(d.variable for d in s.best_solution() if d.value == 1)))
    assert out[0] == out[1]
    ref = O.OracleMisp(inst).solve("wave", k=64, width=300)
    assert out[0][:4] == (ref["best_value"], bool(ref["is_exact"]), ref["explored"], ref["expanded"])


def test_dual_mode_fork_equals_sequential_restricted_then_relaxed(monkeypatch):
    """Forking the relaxed twin on the device at the first cut of the restricted DD (dual mode) gives the same search as compiling the
    two DDs one after the other; instance and widths chosen so that many sub-problems are inexact and the incumbent improves mid-wave."""
    inst = gnp(110, 0.25, 23)
    out = []
    for dual in ("1", "0"):
        monkeypatch.setenv("DDO_DUAL", dual)
        s = ParNoCachingSolverLel(Misp(inst), FixedWidth(24), wave_size=48, batch_cap=32)
        c = s.maximize()
        st = s.stats()
        out.append((c.best_value, c.is_exact, s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"]), int(st["waves"]),
                    sorted(d.variable for d in s.best_solution() if d.value == 1)))
    assert out[0] == out[1]
    ref = O.OracleMisp(inst).solve("wave", k=48, width=24)
    assert out[0][:6] == (ref["best_value"], bool(ref["is_exact"]), ref["explored"], ref["expanded"], ref["transitions"], ref["compilations"])
    assert out[0][7] == ref["solution"]


def test_full_size_config2_root_dd_bit_exact():
    """BASELINE config 2 at full size: G(500, 0.5), W = 10 000 -- root restricted + relaxed DD against the oracle (a few seconds of CPU)."""
    inst = gnp(500, 0.5, 1)
    check_instance(inst, [10000], check_paths=False)


def test_full_size_properties_batch():
    """Size-independent properties at full size on a batch of real sub-problems: relaxed bound >= restricted value, cutset ubs bounded
    by the DD bound, children states are subsets of the root state, a batch equals the same DDs compiled one by one."""
    inst = gnp(500, 0.5, 2)
    pb = Misp(inst)
    W = 2000
    mdd = GpuMdd(pb, W, 8)
    root = SubProblem(pb.initial_state(), 0)
    mdd.compile(CompilationType.Relaxed, W, root)
    cs = mdd.drain_cutset(0, with_paths=False)[:8]
    rel = mdd.compile_batch(CompilationType.Relaxed, [W] * len(cs), cs)
    cuts = [mdd.drain_cutset(i, with_paths=False) for i in range(len(cs))]
    res = mdd.compile_batch(CompilationType.Restricted, [W] * len(cs), cs)
    for i, sp in enumerate(cs):
        assert res[i].best_value <= rel[i].best_value <= sp.ub
        for ch in cuts[i]:
            assert ch.ub <= rel[i].best_value and ch.value >= sp.value
            assert not np.any(ch.state & ~sp.state)
    single = [mdd.compile(CompilationType.Relaxed, W, sp) for sp in cs[:3]]
    for i, c in enumerate(single):
        assert (c.best_value, c.expanded, c.cutset_size) == (rel[i].best_value, rel[i].expanded, rel[i].cutset_size)


def test_solver_times_and_divby_width_heuristics_match_oracle():
    """Times(k, NbUnassignedWidth) / DivBy(k, NbUnassignedWidth) (heuristics/width.rs:636-641,875-880) evaluated per sub-problem on the host."""
    from ddo_b200 import DivBy, Times

    inst = gnp(90, 0.3, 41)
    for wh, kind, k in ((Times(2, NbUnassignedWidth(inst.n)), 2, 2), (DivBy(4, NbUnassignedWidth(inst.n)), 3, 4)):
        s = ParNoCachingSolverLel(Misp(inst), wh, wave_size=16)
        comp = s.maximize()
        ref = O.OracleMisp(inst).solve("wave", k=16, width=k, width_kind=kind)
        assert comp.is_exact and comp.best_value == ref["best_value"]
        assert (s.explored(), int(s.stats()["expanded"])) == (ref["explored"], ref["expanded"])
    assert Times(3, FixedWidth(5)).max_width(SubProblem(None, 0)) == 15 and DivBy(10, FixedWidth(5)).max_width(SubProblem(None, 0)) == 1


@pytest.mark.parametrize("env", [{"DDO_FINISH_CL_MAX": "0"}, {"DDO_EXPAND1_MIN": "1000000000", "DDO_COMPACT1_MIN": "1000000000"}, {"DDO_SMALL_WS_FIRST": "0"},
                                 {"DDO_FINISH_CL_MAX": "0", "DDO_EXPAND1_MIN": "1000000000", "DDO_COMPACT1_MIN": "1000000000", "DDO_SMALL_WS_FIRST": "0", "DDO_DUAL": "0"}])
def test_kernel_variants_give_the_same_search(monkeypatch, env):
    """Every alternative kernel path (one-CTA finish instead of the cluster finish, lane-group expansion / compaction instead of thread-per-node,
    single-tier fast path, no dual mode) compiles the same DDs: same optimum, bounds, counters and solution as the default build and the oracle."""
    inst = gnp(160, 0.3, 29)
    ref = O.OracleMisp(inst).solve("wave", k=32, width=40)

    def run():
        s = ParNoCachingSolverLel(Misp(inst), FixedWidth(40), wave_size=32, batch_cap=32)
        c = s.maximize()
        st = s.stats()
        return (c.best_value, c.is_exact, s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"]), int(st["waves"]),
                sorted(d.variable for d in s.best_solution() if d.value == 1))

    base = run()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    alt = run()
    assert base == alt
    assert base[:6] == (ref["best_value"], bool(ref["is_exact"]), ref["explored"], ref["expanded"], ref["transitions"], ref["compilations"])
    assert base[7] == ref["solution"]


def test_full_size_config5_root_dds_bit_exact_at_width_20000(golden_dir):
    """BASELINE config 5 beyond W ~ 11 000, where the cut keys of the one-CTA kernels leave shared memory: MISP G(1000, 0.5) seed 1, W = 20 000,
    restricted + relaxed root DDs against the oracle's digests (two minutes of CPU, committed in tests/golden/dd_digests.json by
    tests/golden/make_dd_goldens.py): every scalar of the DecisionDiagram trait, the per-layer trace, the cutset (order, states, values,
    upper bounds, depths) and the best exact solution."""
    from parity_util import device_digest
    g = json.loads((golden_dir / "dd_digests.json").read_text())["config5_misp_n1000_p0.5_seed1_w20000"]
    inst = gnp(1000, 0.5, 1)
    pb = Misp(inst)
    mdd = GpuMdd(pb, g["width"], 1)
    root = SubProblem(pb.initial_state(), 0)
    mdd.compile(CompilationType.Restricted, g["width"], root)
    assert device_digest(mdd, 0, O.RESTRICTED) == g["restricted"]
    mdd.compile(CompilationType.Relaxed, g["width"], root, best_lb=g["relaxed_best_lb"])
    assert device_digest(mdd, 0, O.RELAXED) == g["relaxed"]
    mdd.close()


def test_bench_configuration_trajectory_matches_oracle_golden(golden_dir):
    """The bench's own configuration (wave 2048, batch cap 512) on BASELINE config 2: the trajectory of tests/golden/config2_trajectory_k2048.json
    (the oracle's wave solver, tests/golden/make_trajectory.py 2048)."""
    g = json.loads((golden_dir / "config2_trajectory_k2048.json").read_text())
    s = ParNoCachingSolverLel(Misp(gnp(500, 0.5, 1)), FixedWidth(g["width"]), wave_size=2048, batch_cap=512)
    comp = s.maximize()
    st = s.stats()
    assert comp.is_exact and comp.best_value == g["best_value"] == 13
    assert (s.best_lower_bound(), s.best_upper_bound()) == (g["best_lb"], g["best_ub"])
    got = (s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"]), int(st["waves"]))
    assert got == (g["explored"], g["expanded"], g["transitions"], g["compilations"], g["waves"])
    assert sorted(d.variable for d in s.best_solution() if d.value == 1) == g["solution"]


@pytest.mark.parametrize("cs", ["1", "4", "16"])
def test_persistent_whole_dd_kernel_matches_oracle(monkeypatch, cs):
    """The opt-in persistent whole-DD kernel (dd_kernel.cuh, DDO_DD=1: one thread-block cluster compiles a DD from root to terminal layer, node
    and candidate records in distributed shared memory) at several cluster sizes: DD-level parity over widths / best_lb (cuts, pruning,
    merges, the recycled corner), and a dual-mode solve against the oracle's wave solver."""
    monkeypatch.setenv("DDO_DD", "1")
    monkeypatch.setenv("DDO_DD_CS", cs)
    check_instance(gnp(100, 0.3, 5), [1, 2, 3, 13, 50, 400], best_lbs=(N.I64_MIN, 6))
    check_instance(gnp(129, 0.5, 7), [8, 400], best_lbs=(N.I64_MIN, 5))
    inst = gnp(110, 0.25, 23)
    s = ParNoCachingSolverLel(Misp(inst), FixedWidth(24), wave_size=48, batch_cap=32)
    c = s.maximize()
    st = s.stats()
    ref = O.OracleMisp(inst).solve("wave", k=48, width=24)
    assert (c.best_value, c.is_exact, s.explored(), int(st["expanded"]), int(st["transitions"]), int(st["compilations"])) == \
        (ref["best_value"], bool(ref["is_exact"]), ref["explored"], ref["expanded"], ref["transitions"], ref["compilations"])
    assert sorted(d.variable for d in s.best_solution() if d.value == 1) == ref["solution"]
