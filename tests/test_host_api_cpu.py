"""Host-side mirror of the reference interface (ddo_b200/api.py) on CPU: the width heuristics against the vectors of
ddo/src/implementation/heuristics/width.rs:884-1075, their mapping onto the C ABI, and Solver::gap (abstraction/solver.rs:80-93)."""
import pytest

from ddo_b200 import Decision, DivBy, FixedWidth, NbUnassignedWidth, SubProblem, Times
from ddo_b200 import _native as N
from ddo_b200.api import _width_spec, solver_gap


def _sub(decided):
    return SubProblem("a", 10, [Decision(i, i) for i in range(decided)], 100, decided)


def test_width_heuristics_match_the_reference_vectors():
    nb = NbUnassignedWidth(5)
    assert (nb.max_width(_sub(1)), nb.max_width(_sub(0)), nb.max_width(_sub(5))) == (4, 5, 0)            # width.rs:890-933
    f5 = FixedWidth(5)
    assert (f5.max_width(_sub(1)), f5.max_width(_sub(0)), f5.max_width(_sub(5))) == (5, 5, 5)            # width.rs:942-985
    assert [Times(k, f5).max_width(_sub(5)) for k in (2, 3, 1, 10)] == [10, 15, 5, 50]                   # width.rs:995-1014
    assert [DivBy(k, FixedWidth(w)).max_width(_sub(5)) for k, w in ((2, 4), (3, 9), (1, 10))] == [2, 3, 10]  # width.rs:1017-1035
    assert Times(0, FixedWidth(10)).max_width(_sub(5)) == 1 and Times(10, FixedWidth(0)).max_width(_sub(5)) == 1  # width.rs:1038-1055
    with pytest.raises(ZeroDivisionError):                                                               # width.rs:1057-1074 (#[should_panic])
        DivBy(0, FixedWidth(0)).max_width(_sub(5))


def test_width_heuristics_map_onto_the_abi():
    assert _width_spec(FixedWidth(100), 500) == (N.WIDTH_FIXED, 100, 100)
    assert _width_spec(NbUnassignedWidth(500), 500) == (N.WIDTH_NB_UNASSIGNED, 0, 500)
    assert _width_spec(Times(3, NbUnassignedWidth(500)), 500) == (N.WIDTH_TIMES_NB_UNASSIGNED, 3, 1500)
    assert _width_spec(DivBy(4, NbUnassignedWidth(500)), 500) == (N.WIDTH_DIVBY_NB_UNASSIGNED, 4, 500)
    assert _width_spec(Times(2, FixedWidth(7)), 500) == (N.WIDTH_FIXED, 14, 14)  # a constant folds into FixedWidth
    assert _width_spec(DivBy(2, FixedWidth(7)), 500) == (N.WIDTH_FIXED, 3, 3)
    with pytest.raises(TypeError):
        _width_spec(object(), 5)


def test_solver_gap_is_the_reference_formula():
    """abstraction/solver.rs:80-93 and the expectations of parallel.rs:1152-1254 (gap is 1 before the search, 0 once the bounds meet)."""
    import math

    assert solver_gap(N.I64_MIN, N.I64_MAX) == 1.0 and solver_gap(N.I64_MIN, 220) == 1.0 and solver_gap(5, N.I64_MAX) == 1.0
    assert solver_gap(220, 220) == 0.0 and solver_gap(-7, -7) == 0.0
    assert solver_gap(13, 16) == pytest.approx(3 / 16) and solver_gap(-16, -13) == pytest.approx(3 / 16)
    assert math.isnan(solver_gap(0, 0))


def test_misp_dd_never_has_more_layers_than_its_root_state_has_vertices():
    """The log pool of the device engine sizes a batch by Engine::layers_bound: a MISP DD rooted at a state with p vertices has at most p
    layers below its root (next_variable only returns a vertex that some state of the layer still holds, misp/main.rs:109-143; every state
    is a subset of the root state; a branched vertex leaves all descendants).  Checked on the oracle: restricted and relaxed DDs of
    sub-problems (cutset nodes of a narrow relaxed root DD) of random graphs of different densities."""
    import numpy as np

    import oracle_lib as O
    from ddo_b200.instances import gnp

    checked = 0
    for n, p, seed in ((60, 0.2, 3), (90, 0.5, 5), (120, 0.1, 7), (70, 0.8, 11)):
        o = O.OracleMisp(gnp(n, p, seed))
        root = o.compile(O.RELAXED, 8)
        assert len(root["layer_vars"]) <= n
        for st, val, dep in list(zip(root["cutset_states"], root["cutset_values"], root["cutset_depths"]))[:12]:
            pc = sum(bin(int(w)).count("1") for w in st)
            assert pc <= n - int(dep)
            for comp, width in ((O.RESTRICTED, 5), (O.RELAXED, 5), (O.RELAXED, 1000)):
                r = o.compile(comp, width, root_state=np.asarray(st, dtype=np.uint64), root_value=int(val), root_depth=int(dep))
                assert r["rc"] == 0
                assert len(r["layer_vars"]) <= pc, (n, p, seed, pc, len(r["layer_vars"]))
                assert len(set(int(v) for v in r["layer_vars"])) == len(r["layer_vars"])  # every vertex is branched on at most once
                checked += 1
    assert checked >= 60
