"""The oracle against the reference's own golden vectors (CPU only).

* oracle/selftest.cpp restates the reference's unit tests (clean.rs:1097-2668, node_flags.rs, no_duplicate.rs, solver tests).
* tests/golden/*.dot are the reference's graphviz goldens (resources/visualisation_tests, compared verbatim by clean.rs:2401-2546):
  every node's val / locb / rub / theta, every edge with decision, cost and best-edge pen width, cutset / relaxed / deleted styling.
* tests/golden/misp, tests/golden/knapsack + expected.json: instance optima asserted by examples/{misp,knapsack}/tests.rs.
"""
import json
import re

import numpy as np
import pytest

import oracle_lib as O
from ddo_b200.instances import parse_dimacs, parse_knapsack, random_knapsack, read_wcnf


def test_selftest_restated_reference_unit_tests():
    r = O.run_selftest()
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SELFTEST OK" in r.stdout


# ----------------------------------------------------------------------------------------------------------------
# graphviz goldens
# ----------------------------------------------------------------------------------------------------------------
_NODE = re.compile(r'^(\d+) \[shape=(\w+),style=filled,color=("?[#\w]+"?),peripheries=(\d),group="(\w+)",label="\'(.)\'(.*)"\];$')
_EDGE = re.compile(r'^(\d+) -> (\d+) \[penwidth=(\d),label="\(x(\d+) = (-?\d+)\)\\ncost = (-?\d+)"\];$')


def parse_dot(text):
    nodes, edges = {}, []
    for raw in text.splitlines():
        line = raw.strip()
        m = _NODE.match(line)
        if m:
            nid, shape, color, periph, _group, label, rest = m.groups()
            attrs = dict(re.findall(r"\\n(\w+): ([-+\w]+)", rest))
            nodes[int(nid)] = {"label": label, "shape": shape, "color": color.strip('"'), "peripheries": int(periph), **attrs}
            continue
        m = _EDGE.match(line)
        if m:
            a, b, pen, var, val, cost = m.groups()
            edges.append((int(a), int(b), int(pen), int(var), int(val), int(cost)))
    return nodes, edges


def parse_dump(text):
    nodes, edges = {}, []
    for line in text.splitlines():
        t = line.split()
        if t[0] == "N":
            nodes[t[1]] = dict(val=t[2], locb=t[3], rub=t[4], theta=t[5], exact=t[6] == "1", relaxed=t[7] == "1", cutset=t[8] == "1", deleted=t[9] == "1")
        else:
            edges.append((t[1], t[2], int(t[3]), int(t[4]), int(t[5]), t[6] == "1"))
    return nodes, edges


def test_default_viz_golden_pins_every_node_and_edge(golden_dir):
    """clean.rs:2401-2431 (test_default_visualisation): FC cutset, relaxed, W=3, best_lb=0."""
    gnodes, gedges = parse_dot((golden_dir / "default_viz.dot").read_text())
    onodes, oedges = parse_dump(O.locbounds_dump(O.FRONTIER, 0))
    by_label = {v["label"]: v for v in gnodes.values()}
    # default config hides deleted nodes: c and d are absent from the golden, present (deleted) in the oracle
    assert set(by_label) == {k for k, v in onodes.items() if not v["deleted"]}
    assert {k for k, v in onodes.items() if v["deleted"]} == {"c", "d"}
    for label, g in by_label.items():
        o = onodes[label]
        assert g["val"] == o["val"], label
        assert g["rub"] == o["rub"], label
        assert g["theta"] == ("+inf" if o["theta"] == "none" else o["theta"]) or (g["theta"] == o["theta"]), label
        # locb is printed for every node; the terminal's is 0
        assert g["locb"] == o["locb"], label
        assert (g["color"] == "red" and g["peripheries"] == 4) == o["cutset"], label  # cutset styling
        assert (g["shape"] == "square") == o["relaxed"], label                      # merged nodes are squares
    id2label = {k: v["label"] for k, v in gnodes.items()}
    gset = sorted((id2label[a], id2label[b], var, val, cost, pen == 3) for a, b, pen, var, val, cost in gedges)
    oset = sorted(e for e in oedges if not onodes[e[1]]["deleted"])
    assert gset == oset


def test_deleted_viz_golden_shows_merged_away_nodes(golden_dir):
    """clean.rs:2471-2507 (show_deleted): the merged-away nodes c,d and their original edges are still in the DD."""
    gnodes, gedges = parse_dot((golden_dir / "deleted_viz.dot").read_text())
    onodes, oedges = parse_dump(O.locbounds_dump(O.FRONTIER, 0))
    assert {v["label"] for v in gnodes.values()} == set(onodes)
    id2label = {k: v["label"] for k, v in gnodes.items()}
    gset = sorted((id2label[a], id2label[b], var, val, cost, pen == 3) for a, b, pen, var, val, cost in gedges)
    assert gset == sorted(oedges)
    # creation order of the nodes (ids in the golden) is reproduced up to the hash-order of the last two layers
    order = [id2label[i] for i in sorted(id2label)]
    assert order[:8] == ["r", "a", "b", "c", "d", "e", "f", "M"]
    assert list(onodes)[:8] == order[:8]


# ----------------------------------------------------------------------------------------------------------------
# instance-level known answers
# ----------------------------------------------------------------------------------------------------------------
def _expected(golden_dir):
    return json.loads((golden_dir / "expected.json").read_text())


FAST_MISP = ["johnson8-2-4", "hamming6-4", "hamming6-2", "MANN_a9", "johnson8-4-4", "hamming8-2", "brock200_2", "c-fat200-5", "c-fat200-1", "c-fat200-2",
             "p_hat300-1", "c-fat500-1", "c-fat500-2"]
SLOW_MISP = ["keller4", "brock200_3"]


@pytest.mark.parametrize("name", FAST_MISP + SLOW_MISP)
def test_misp_known_optima_parallel_solver(golden_dir, name):
    """examples/misp/tests.rs: DefaultSolver (parallel, LEL, NbUnassignedWidth) proves the asserted optimum."""
    exp = _expected(golden_dir)["misp"][name]
    inst = parse_dimacs(O.read_clq(golden_dir, name), name)
    r = O.OracleMisp(inst).solve("parallel", k=8)
    assert r["is_exact"] and r["best_value"] == exp["optimum"], exp["source"]
    assert r["best_lb"] == r["best_ub"] == exp["optimum"]
    _check_independent(inst, r["solution"], exp["optimum"])


def _check_independent(inst, sol, value):
    adj = set(zip(inst.src.tolist(), inst.dst.tolist())) | set(zip(inst.dst.tolist(), inst.src.tolist()))
    assert all((a, b) not in adj for a in sol for b in sol if a != b)
    assert int(inst.weights[sol].sum()) == value


@pytest.mark.parametrize("name", FAST_MISP)
def test_misp_sequential_wave_and_parallel_agree(golden_dir, name):
    """Objective and proven bound are schedule independent; wave K=1 is exactly the sequential solver."""
    exp = _expected(golden_dir)["misp"][name]["optimum"]
    inst = parse_dimacs(O.read_clq(golden_dir, name), name)
    o = O.OracleMisp(inst)
    seq = o.solve("sequential")
    w1 = o.solve("wave", k=1)
    w16 = o.solve("wave", k=16)
    for r in (seq, w1, w16):
        assert r["is_exact"] and r["best_value"] == exp and r["best_lb"] == exp and r["best_ub"] == exp
    # same DDs compiled in the same order; the wave solver (like parallel.rs:531-535) clears the fringe at the first popped node
    # with ub <= best_lb whereas sequential.rs:337 pops and discards them one by one, hence explored may only be smaller
    assert (seq["expanded"], seq["transitions"], seq["compilations"]) == (w1["expanded"], w1["transitions"], w1["compilations"])
    assert w1["explored"] <= seq["explored"]
    assert seq["solution"] == w1["solution"]
    # fixed width and frontier cutset variants reach the same optimum
    assert o.solve("sequential", width=5)["best_value"] == exp
    assert o.solve("sequential", cutset_type=O.FRONTIER)["best_value"] == exp


def test_knapsack_known_optima(golden_dir):
    """examples/knapsack/tests.rs:65-206 (SeqCachingSolverFc-equivalent: FC cutset + SimpleCache + KPDominance) -- BASELINE config 1 plumbing."""
    exp = _expected(golden_dir)["knapsack"]
    assert len(exp) >= 12
    for name, e in exp.items():
        inst = parse_knapsack((golden_dir / "knapsack" / name).read_text(), name)
        for caching, cutset in ((True, O.FRONTIER), (False, O.LEL)):
            if not caching and e["items"] > 23:
                continue
            r = O.knapsack_solve(inst, caching=caching, cutset_type=cutset)
            assert r["is_exact"] and r["best_value"] == e["optimum"], (name, e["source"])
            assert int(inst.profit[r["taken"] == 1].sum()) == e["optimum"]
            assert int(inst.weight[r["taken"] == 1].sum()) <= inst.capacity


def test_knapsack_config1_50_items_matches_dp():
    """BASELINE config 1: 50-item instance, SequentialSolver on CPU; checked against a textbook DP."""
    inst = random_knapsack(50, seed=1)
    best = np.zeros(inst.capacity + 1, dtype=np.int64)
    for p, w in zip(inst.profit.tolist(), inst.weight.tolist()):
        best[w:] = np.maximum(best[w:], best[:-w] + p) if w <= inst.capacity else best[w:]
    r = O.knapsack_solve(inst, solver="sequential", caching=True, cutset_type=O.FRONTIER)
    assert r["is_exact"] and r["best_value"] == int(best[-1])
    r2 = O.knapsack_solve(inst, solver="parallel", k=4, caching=True, cutset_type=O.FRONTIER)
    assert r2["best_value"] == int(best[-1])


# ----------------------------------------------------------------------------------------------------------------
# MAX2SAT (BASELINE config 3)
# ----------------------------------------------------------------------------------------------------------------
def test_max2sat_model_unit_vectors(golden_dir):
    """examples/max2sat/model.rs:388-448 (test_initial_value, test_next_state, test_rank) and data.rs:118-125 (4 clauses, 3 vars)."""
    inst = read_wcnf(golden_dir / "max2sat" / "debug2.wcnf")
    assert inst.n == 3 and len(inst.clauses) == 4
    o = O.OracleM2s(inst)
    assert o.initial_value() == 0
    root = np.zeros(3, dtype=np.int32)
    nod_f, _, _, _ = o.transition(root, 0, 0, -1)
    assert nod_f.tolist() == [0, -4, 3]
    nod_t, _, _, _ = o.transition(root, 0, 0, 1)
    assert nod_t.tolist() == [0, 0, 0]
    benef = [-183, -122, -61, -183, -183, -183, -122, -122, -183, -122, -61, -122, 0, -122, -122, -122, -122, -122, -183, -122, -61, 0, -122, -61, 0, 0, 0, 0, 0,
             -244, -61, -183, 0, -122, -244, -183, -61, -61, -122, -122, -122, -183, -122, 0, -183, -61, -183, -122, -122, -183, -183, -61, -61, -122, 0, 0, 0, 0, 0, 0]
    assert int(np.abs(np.array(benef)).sum()) == 5917  # model.rs:434-448: rank() is the sum of absolute benefits


def test_max2sat_known_optima(golden_dir):
    """examples/max2sat/tests.rs:65-106: the ten instances the reference solves in its (non-ignored) tests."""
    exp = _expected(golden_dir)["max2sat"]
    assert len(exp) == 10
    for name, e in exp.items():
        inst = read_wcnf(golden_dir / "max2sat" / f"{name}.wcnf")
        o = O.OracleM2s(inst)
        if inst.n > 10:  # frb10-6-*: 60 variables; one default-width sequential solve each is ~7 s -> use the threaded solver
            r = o.solve("parallel", k=8)
        else:
            r = o.solve("sequential")
            assert o.solve("sequential", width=2)["best_value"] == e["optimum"]
            assert o.solve("wave", k=4, width=3)["best_value"] == e["optimum"]
        assert r["is_exact"] and r["best_value"] == e["optimum"], (name, e["source"])
        # the decision vector really has that value: count the satisfied clause weights
        model = dict(r["solution"])
        assert len(model) == inst.n
        uniq = {}
        for w, x, y in inst.clauses.tolist():
            uniq[(min(x, y), max(x, y))] = w
        sat = sum(w for (x, y), w in uniq.items() if model[abs(x) - 1] * x > 0 or model[abs(y) - 1] * y > 0)
        assert sat == e["optimum"], name


def test_max2sat_restricted_relaxed_bracket_the_optimum(golden_dir):
    """A restricted DD gives a lower bound, a relaxed DD an upper bound, the exact DD the optimum (clean.rs:345-381 on the MAX2SAT model)."""
    inst = read_wcnf(golden_dir / "max2sat" / "pass.wcnf")
    o = O.OracleM2s(inst)
    ex = o.compile(O.EXACT, 1 << 20)
    assert ex["is_exact"] and ex["best_value"] == 54
    for w in (1, 2, 3):
        lo = o.compile(O.RESTRICTED, w)
        hi = o.compile(O.RELAXED, w)
        assert lo["best_value"] <= 54 <= hi["best_value"]


# ----------------------------------------------------------------------------------------------------------------
# TSPTW (BASELINE config 4): oracle only so far -- the device model is the next row of the scope table
# ----------------------------------------------------------------------------------------------------------------
def _tour_cost(inst, perm):
    """Travel + waiting time of the tour depot -> perm[0] -> ... -> perm[n-1] (= depot), None when a time window is missed."""
    t, pos, total = 0, 0, 0
    for city in perm.tolist():
        arrive = t + int(inst.dist[pos, city])
        if arrive > int(inst.tw[city, 1]):
            return None
        start = max(arrive, int(inst.tw[city, 0]))
        total += start - t
        t, pos = start, city
    return total


def test_tsptw_known_optima(golden_dir):
    """examples/tsptw/tests.rs:33-57,80-683: `solve(instance, width = 1, threads = 1)` with the DefaultCachingSolver stack (FRONTIER cutset,
    SimpleCache, SimpleDominanceChecker(TsptwDominance), TsptwWidth, NoDupFringe(MaxUB(TsptwRanking))); the asserted value is
    -(best_value) / 10000 as f32.  The solution is re-scored as a tour: every city once, every time window met, same cost."""
    exp = _expected(golden_dir)["tsptw"]
    assert len(exp) == 15
    for name, e in exp.items():
        inst = O.TsptwInstance((golden_dir / "tsptw" / name).read_text())
        r = O.tsptw_solve(inst)
        assert r["is_exact"] and r["has_value"], name
        assert np.float32(r["cost"]) == np.float32(float(e["optimum"])), (name, r["cost"], e["optimum"], e["source"])
        assert r["best_lb"] == r["best_ub"] == r["best_value"]
        perm = r["perm"]
        assert sorted(perm.tolist()) == list(range(inst.n)) and perm[-1] == 0, name
        assert _tour_cost(inst, perm) == -r["best_value"], name


def test_tsptw_solver_variants_agree(golden_dir):
    """The optimum does not depend on the cutset type, the cache / dominance filters, the width factor or the number of threads."""
    for name in ("SolomonPotvinBengio/rc_203.4.txt", "SolomonPotvinBengio/rc_205.1.txt", "Langevin/N40ft403.dat"):
        inst = O.TsptwInstance((golden_dir / "tsptw" / name).read_text())
        ref = O.tsptw_solve(inst)["best_value"]
        assert O.tsptw_solve(inst, factor=3)["best_value"] == ref
        assert O.tsptw_solve(inst, solver="parallel", k=4)["best_value"] == ref
        assert O.tsptw_solve(inst, cutset_type=O.LEL)["best_value"] == ref
        r = O.tsptw_solve(inst, caching=False, cutset_type=O.LEL, time_budget_s=30)
        assert (not r["is_exact"]) or r["best_value"] == ref


def test_tsptw_instance_scaling_is_f32():
    """instance.rs:86-87: `(distance * 10000.0) as usize` in f32 -- 79.4 becomes 794000 exactly but 70.3 becomes 703000 only through f32 rounding."""
    inst = O.TsptwInstance("2\n0 70.3\n79.4 0\n0 1000.5\n10.25 408\n")
    assert inst.dist.tolist() == [[0, int(np.float32(70.3) * np.float32(10000.0))], [794000, 0]]
    assert inst.tw.tolist() == [[0, 10005000], [102500, 4080000]]


def test_tsptw_dd_level_checker(golden_dir):
    """The DD-level entry points the TSPTW device model will be compared against (oracle_capi.cpp::TsptwHandle, 16-word packed states):
    restricted <= optimum <= relaxed, the frontier cutset is a superset in size of nothing less than the last exact layer's marked nodes'
    parents, states survive the ABI round trip, and branching on a relaxed DD's exact cutset recovers the optimum."""
    from ddo_b200 import parse_tsptw

    recovered = 0
    for name in ("SolomonPotvinBengio/rc_207.4.txt", "SolomonPotvinBengio/rc_203.4.txt", "SolomonPotvinBengio/rc_205.1.txt"):
        inst = parse_tsptw((golden_dir / "tsptw" / name).read_text())
        opt = O.tsptw_solve(inst)["best_value"]
        o = O.OracleTsptw(inst)
        for cutset in (O.LEL, O.FRONTIER):
            width = 2 if inst.n <= 6 else 5
            lo = o.compile(O.RESTRICTED, width, cutset_type=cutset)
            hi = o.compile(O.RELAXED, width, cutset_type=cutset, want_paths=True)
            assert (not lo["has_best"]) or lo["best_value"] <= opt
            assert hi["has_best"] and hi["best_value"] >= opt
            if hi["is_exact"]:
                assert hi["best_value"] == opt
                continue
            assert hi["cutset_size"] > 0
            assert all(len(p) == d for p, d in zip(hi["cutset_paths"], hi["cutset_depths"].tolist()))
            assert all(int(u) <= hi["best_value"] for u in hi["cutset_ubs"].tolist())
            # every cutset node is an exact sub-problem: the best of their own optima (wide restricted DDs are exact here) is the optimum
            best, all_exact = None, True
            for i in range(hi["cutset_size"]):
                if int(hi["cutset_ubs"][i]) < opt:
                    continue
                sub = o.compile(O.RESTRICTED, 2000, root_state=hi["cutset_states"][i], root_value=int(hi["cutset_values"][i]),
                                root_depth=int(hi["cutset_depths"][i]))
                all_exact = all_exact and bool(sub["is_exact"])
                if sub["has_best"]:
                    best = sub["best_value"] if best is None else max(best, sub["best_value"])
            assert best is None or best <= opt
            if all_exact:
                assert best == opt, (name, cutset, best, opt)
                recovered += 1
        # the layer-1 width of a relaxed DD is never cut (clean.rs:789), the others are bounded by max_width
        w = o.compile(O.RELAXED, 5)["layer_widths"].tolist()
        assert w[0] == 1 and all(x <= 5 for x in w[2:])
    assert recovered >= 1


def test_frontier_and_lel_solvers_agree_with_brute_force():
    """The oracle's FRONTIER path is the checker of the device's frontier kernels: on small random instances the sequential, wave and
    parallel solvers with either cutset type must all prove the brute-force optimum (MISP: maximum weight independent set by enumeration;
    MAX2SAT: best assignment by enumeration)."""
    import itertools

    from ddo_b200 import gnp, random_max2sat

    for seed in range(6):
        inst = gnp(16, 0.3, 100 + seed)
        rng = np.random.default_rng(seed)
        inst.weights[:] = rng.integers(1, 9, size=inst.n)
        adj = np.zeros((inst.n, inst.n), dtype=bool)
        adj[inst.src, inst.dst] = True
        adj[inst.dst, inst.src] = True
        best = 0
        for mask in range(1 << inst.n):
            vs = [v for v in range(inst.n) if mask >> v & 1]
            if all(not adj[a, b] for a, b in itertools.combinations(vs, 2)):
                best = max(best, int(inst.weights[vs].sum()))
        o = O.OracleMisp(inst)
        for width in (1, 2, 5):
            for cutset in (O.LEL, O.FRONTIER):
                for mode, k in (("sequential", 1), ("wave", 4), ("parallel", 3)):
                    r = o.solve(mode, k=k, width=width, cutset_type=cutset)
                    assert r["is_exact"] and r["best_value"] == best, (seed, width, cutset, mode)
    for seed in range(4):
        inst = random_max2sat(10, 40, 200 + seed)
        uniq = {}
        for w, x, y in inst.clauses.tolist():
            uniq[(min(x, y), max(x, y))] = w
        best = max(sum(w for (x, y), w in uniq.items() if (a >> (abs(x) - 1) & 1) == (x > 0) or (a >> (abs(y) - 1) & 1) == (y > 0)) for a in range(1 << inst.n))
        o = O.OracleM2s(inst)
        for width in (1, 3):
            for cutset in (O.LEL, O.FRONTIER):
                for mode, k in (("sequential", 1), ("wave", 4)):
                    r = o.solve(mode, k=k, width=width, cutset_type=cutset)
                    assert r["is_exact"] and r["best_value"] == best, (seed, width, cutset, mode)
