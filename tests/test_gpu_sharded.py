"""The fringe-sharded driver (ddo_b200/sharded.py) over the DEVICE solver's stepwise interface (ddo_solver_init / wave / retain_share /
set_lower_bound / finish): `world` ranks run as threads on one GPU with an in-process allreduce(max), and every rank must follow exactly
the trajectory of the CPU oracle's stepper run through the same protocol (the gloo tests cover the protocol with real process groups)."""
import threading

import pytest

import oracle_lib as O
from ddo_b200 import FixedWidth, Max2Sat, Misp, ParNoCachingSolverLel, gnp, random_max2sat
from ddo_b200.sharded import sharded_maximize

pytestmark = pytest.mark.gpu


class _ThreadAllreduce:
    """allreduce(max) between `world` threads: two barriers per call."""

    def __init__(self, world):
        self.world = world
        self.slots = [None] * world
        self.bar = threading.Barrier(world)

    def for_rank(self, rank):
        def f(vals):
            self.slots[rank] = list(vals)
            self.bar.wait()
            out = [max(s[i] for s in self.slots) for i in range(len(vals))]
            self.bar.wait()
            return out
        return f


def _run(world, make_stepper):
    ar = _ThreadAllreduce(world)
    res = [None] * world
    err = []

    def work(rank):
        try:
            st = make_stepper()
            r = sharded_maximize(st, rank, world, ar.for_rank(rank))
            r.update(explored=st.explored(), expanded=int(st.stats()["expanded"]) if hasattr(st, "stats") else st.expanded())
            res[rank] = r
        except Exception as e:  # a dead rank would leave the others at the barrier
            err.append(e)
            ar.bar.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not err, err
    return res


@pytest.mark.parametrize("world", [2, 3])
def test_device_ranks_follow_the_oracle_ranks_misp(world):
    inst = gnp(120, 0.4, 31)
    pb = Misp(inst)
    dev = _run(world, lambda: ParNoCachingSolverLel(pb, FixedWidth(10), wave_size=32))
    oracle = O.OracleMisp(inst)
    ref = _run(world, lambda: O.OracleStepper(oracle, 32, 10))
    single = oracle.solve("wave", k=32, width=10)
    for d, r in zip(dev, ref):
        assert d["is_exact"] and d["best_lb"] == d["best_ub"] == single["best_value"]
        assert (d["waves"], d["collectives"], d["explored"], d["expanded"]) == (r["waves"], r["collectives"], r["explored"], r["expanded"])
    assert max(d["explored"] for d in dev) < single["explored"] + world


def test_device_ranks_follow_the_oracle_ranks_max2sat():
    inst = random_max2sat(35, 200, 5)
    pb = Max2Sat(inst)
    dev = _run(2, lambda: ParNoCachingSolverLel(pb, FixedWidth(10), wave_size=16))
    oracle = O.OracleM2s(inst)
    ref = _run(2, lambda: O.OracleStepper(oracle, 16, 10))
    for d, r in zip(dev, ref):
        assert d["is_exact"] and d["best_lb"] == d["best_ub"] == oracle.solve("wave", k=16, width=10)["best_value"]
        assert (d["waves"], d["explored"], d["expanded"]) == (r["waves"], r["explored"], r["expanded"])
