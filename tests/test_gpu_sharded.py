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


class _ThreadComm:
    """The comm interface of ddo_b200.sharded (allgather / send / recv) between `world` threads of one process."""

    def __init__(self, world):
        import queue
        self.world = world
        self.slots = [None] * world
        self.bar = threading.Barrier(world)
        self.q = {(a, b): queue.Queue() for a in range(world) for b in range(world)}

    def for_rank(self, rank):
        import numpy as np
        outer = self

        class C:
            def allgather(self, vals):
                outer.slots[rank] = list(vals)
                outer.bar.wait()
                out = np.asarray(outer.slots, dtype=np.int64)
                outer.bar.wait()
                return out

            def send(self, arr, peer):
                outer.q[(rank, peer)].put(np.array(arr, dtype=np.int64).reshape(-1).copy())

            def recv(self, count, peer):
                a = outer.q[(peer, rank)].get(timeout=120)
                assert a.size == count
                return a
        return C()


def _run(world, make_stepper, new_protocol=False):
    ar = _ThreadComm(world) if new_protocol else _ThreadAllreduce(world)
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


@pytest.mark.parametrize("world", [2, 4])
def test_rebalanced_device_ranks_follow_the_oracle_ranks(world):
    """The round-2 protocol (one all-gather per wave, open nodes handed from loaded to idle ranks through ddo_solver_export_open /
    ddo_solver_import_open, solution gathered from the rank that holds it): device ranks and oracle ranks exchange the same nodes and
    follow the same trajectories; every rank returns the optimum with an independent set of that size."""
    inst = gnp(130, 0.35, 11)
    pb = Misp(inst)
    dev = _run(world, lambda: ParNoCachingSolverLel(pb, FixedWidth(6), wave_size=8), new_protocol=True)
    oracle = O.OracleMisp(inst)
    ref = _run(world, lambda: O.OracleStepper(oracle, 8, 6), new_protocol=True)
    single = oracle.solve("wave", k=8, width=6)
    assert sum(d["nodes_sent"] for d in dev) == sum(d["nodes_received"] for d in dev) > 0
    for d, r in zip(dev, ref):
        assert d["is_exact"] and d["best_lb"] == d["best_ub"] == single["best_value"] == d["best_value"]
        assert (d["waves"], d["collectives"], d["handoffs"], d["nodes_sent"], d["nodes_received"], d["explored"], d["expanded"]) == \
            (r["waves"], r["collectives"], r["handoffs"], r["nodes_sent"], r["nodes_received"], r["explored"], r["expanded"])
        chosen = [v for v, x in d["solution"] if x == 1]
        assert len(chosen) == single["best_value"] and all(not inst.has_edge(a, b) for i, a in enumerate(chosen) for b in chosen[i + 1:])


def test_asynchronous_board_protocol_over_device_solvers():
    """The opt-in asynchronous protocol (ddo_b200.sharded.sharded_maximize_async: status board in shared memory, no collective) with two
    and three device solvers as ranks (threads of this process): every rank proves the optimum and returns one common independent set."""
    from ddo_b200.sharded import StatusBoard, sharded_maximize_async

    inst = gnp(130, 0.35, 11)
    pb = Misp(inst)
    single = O.OracleMisp(inst).solve("wave", k=8, width=6)
    for world in (2, 3):
        box, bar = [None], threading.Barrier(world)
        res, err = [None] * world, []

        def bootstrap_for(rank):
            def f(obj):
                if rank == 0:
                    box[0] = obj
                bar.wait()
                out = box[0]
                bar.wait()
                return out
            return f

        def work(rank):
            try:
                st = ParNoCachingSolverLel(pb, FixedWidth(6), wave_size=8)
                board = StatusBoard(rank, world, st.node_words(), inst.n, bootstrap_for(rank), mail_nodes=256)
                res[rank] = sharded_maximize_async(st, rank, world, board)
                bar.wait()
                board.close()
            except Exception as e:
                err.append(e)
                bar.abort()

        ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not err, err
        assert all(r["is_exact"] and r["best_lb"] == single["best_value"] == r["best_value"] and r["collectives"] == 0 for r in res)
        assert sum(r["nodes_sent"] for r in res) == sum(r["nodes_received"] for r in res)
        chosen = [v for v, x in res[0]["solution"] if x == 1]
        assert len(chosen) == single["best_value"] and all(not inst.has_edge(a, b) for i, a in enumerate(chosen) for b in chosen[i + 1:])
        assert all(r["solution"] == res[0]["solution"] for r in res)
