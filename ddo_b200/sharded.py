"""Fringe-sharded branch-and-bound over several GPUs of one box (one process per GPU).

SURVEY.md section 8(e): the open sub-problems are independent, so the path shards with NO data-path collective.  Every rank compiles the
root DD (identical, deterministic), keeps its share of the open nodes of the common MaxUB order (`retain_share`: a rotating deal), and then runs waves on its
own fringe.  After each wave ONE collective -- an all-gather of four int64 per rank: [best_lb, ub of the best open node, fringe length,
objective of the locally held solution] -- synchronises the incumbent lower bound, the global proven upper bound, termination AND tells
every rank how loaded the others are.  This replaces the mutex-protected `Critical` block of ddo/src/implementation/solver/parallel.rs:32-81
(best_lb read at :398/:426, written at :446-453; termination test at :512).

The reference's workers pull from ONE fringe (parallel.rs:500-559), so its load balances itself.  Here every rank derives the same
hand-off plan from the gathered fringe lengths: a rank whose fringe has run (nearly) dry receives packed open nodes -- state, value, bound,
depth, full decision path -- from the most loaded one, point to point (`export_open` / `import_open`, ncclSend / ncclRecv).  No transfer
happens while the fringes are within a factor of each other, so the steady state is one 32-byte-per-rank collective per wave.

When the search ends, the rank whose local solution reaches the global optimum broadcasts its decisions: every rank returns the optimum
WITH a solution of that value (the reference's set_primal keeps value and solution together, solver.rs:77).

`stepper` is anything with the stepwise solver interface (init / wave / set_lower_bound / retain_share / export_open / import_open /
finish / getters): the device solver (`ParNoCachingSolverLel`) in production, a CPU stand-in in the gloo tests.  `comm` provides
allgather / send / recv: `NativeComm` (ddo_comm_* of the C ABI: NCCL from C++, what a Rust host would bind) or `TorchComm`
(torch.distributed: gloo on CPU for the tests).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

I64_MIN = -(1 << 63)
MAX_HANDOFF = 8192      # open nodes per hand-off
MIN_DONOR = 64          # a rank keeps at least this many nodes for itself


class TorchComm:
    """Collectives over torch.distributed (NCCL when `device` is a CUDA device, gloo on CPU)."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allgather(self, vals: List[int]) -> np.ndarray:
        t = self.torch.tensor(vals, dtype=self.torch.int64, device=self.device)
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t)
        return self.torch.stack(parts).cpu().numpy()

    def send(self, arr: np.ndarray, peer: int):
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64).reshape(-1))
        self.dist.send(t.to(self.device) if self.device is not None else t, peer)

    def recv(self, count: int, peer: int) -> np.ndarray:
        t = self.torch.empty(count, dtype=self.torch.int64, device=self.device)
        self.dist.recv(t, peer)
        return t.cpu().numpy()

    def close(self):
        pass


class NativeComm:
    """ddo_comm_* of include/ddo_b200.h: NCCL driven from the C ABI.  `bootstrap(obj) -> obj` hands rank 0's unique id to the others
    (any out-of-band channel; bench.py uses torch.distributed's store)."""

    def __init__(self, rank: int, world: int, device: int, bootstrap):
        from . import _native as N
        self.N, self.rank, self.world = N, rank, world
        ident = (C.c_char * 128)()
        if rank == 0:
            N.check(N.lib().ddo_comm_unique_id(ident), "ddo_comm_unique_id")
        raw = bootstrap(bytes(ident.raw))
        ident = (C.c_char * 128).from_buffer_copy(raw)
        h = C.c_void_p()
        N.check(N.lib().ddo_comm_init(world, rank, ident, device, C.byref(h)), "ddo_comm_init")
        self.h = h

    def allgather(self, vals: List[int]) -> np.ndarray:
        v = np.asarray(vals, dtype=np.int64)
        out = np.zeros(self.world * len(vals), dtype=np.int64)
        self.N.check(self.N.lib().ddo_comm_allgather(self.h, v.ctypes.data_as(C.c_void_p), len(vals), out.ctypes.data_as(C.c_void_p)), "ddo_comm_allgather")
        return out.reshape(self.world, len(vals))

    def allreduce_max(self, vals: List[int]) -> List[int]:
        v = np.asarray(vals, dtype=np.int64)
        self.N.check(self.N.lib().ddo_comm_allreduce_max(self.h, v.ctypes.data_as(C.c_void_p), len(vals)), "ddo_comm_allreduce_max")
        return [int(x) for x in v]

    def send(self, arr: np.ndarray, peer: int):
        a = np.ascontiguousarray(arr, dtype=np.int64).reshape(-1)
        self.N.check(self.N.lib().ddo_comm_send(self.h, a.ctypes.data_as(C.c_void_p), a.nbytes, peer), "ddo_comm_send")

    def recv(self, count: int, peer: int) -> np.ndarray:
        a = np.zeros(count, dtype=np.int64)
        self.N.check(self.N.lib().ddo_comm_recv(self.h, a.ctypes.data_as(C.c_void_p), a.nbytes, peer), "ddo_comm_recv")
        return a

    def close(self):
        if getattr(self, "h", None):
            self.N.lib().ddo_comm_destroy(self.h)
            self.h = None


def torch_allreduce_max(device=None):
    """(kept for callers of the round-1 interface) allreduce(max) over torch.distributed."""
    import torch
    import torch.distributed as dist

    def f(vals: List[int]) -> List[int]:
        t = torch.tensor(vals, dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [int(x) for x in t.tolist()]

    return f


def handoff_plan(lens: List[int]) -> List[tuple]:
    """(donor, receiver, count) triples, the same on every rank: the emptiest ranks are refilled by the fullest ones.  A rank receives when
    it holds less than a quarter of the mean; a donor gives half of what it has above the mean, never below MIN_DONOR of its own."""
    world = len(lens)
    total = sum(lens)
    if world < 2 or total < world * 2:
        return []
    mean = total / world
    left = list(lens)
    order_lo = sorted(range(world), key=lambda r: (left[r], r))
    order_hi = sorted(range(world), key=lambda r: (-left[r], r))
    plan, used = [], set()
    for dst in order_lo:
        if left[dst] * 4 >= mean:
            break
        for src in order_hi:
            if src == dst or src in used or left[src] <= mean or left[src] < 2 * MIN_DONOR:
                continue
            cnt = int(min(MAX_HANDOFF, (left[src] - mean) / 2 + 1, (left[src] - MIN_DONOR) // 2, max(mean - left[dst], 1)))
            if cnt > 0:
                plan.append((src, dst, cnt))
                used.add(src)
                left[src] -= cnt
                left[dst] += cnt
            break
    return plan


def sharded_maximize(stepper, rank: int, world: int, comm, max_waves: int = 0, rebalance: bool = True):
    """Returns dict(best_lb, best_ub, waves, collectives, handoffs, nodes_sent, nodes_received, is_exact, best_value, solution)."""
    if callable(comm) and not hasattr(comm, "allgather"):  # round-1 interface: an allreduce(max) function (no hand-off, no solution gather)
        return _sharded_maximize_allreduce(stepper, rank, world, comm, max_waves)
    stepper.init(True)
    lb, top, more = stepper.wave()  # the root DD: identical on every rank
    stepper.retain_share(rank, world)
    waves, colls, handoffs, sent, received = 1, 0, 0, 0, 0
    best_ub: Optional[int] = None
    aborted = False
    nw = stepper.node_words() if rebalance and hasattr(stepper, "node_words") else 0
    top = I64_MIN  # (the root's bound says nothing about the shares)
    while True:
        sv = stepper.best_value() if hasattr(stepper, "best_value") else None
        rows = comm.allgather([lb, top, stepper.fringe_len(), I64_MIN if sv is None else sv])  # ---- the ONE collective of the wave
        colls += 1
        g_lb = int(rows[:, 0].max())
        if g_lb > lb:
            stepper.set_lower_bound(g_lb)
            lb = g_lb
        g_top = int(rows[:, 1].max())
        if g_top != I64_MIN:
            best_ub = g_top  # the best open node anywhere before the last wave: a valid global bound while the search runs
        lens = [int(x) for x in rows[:, 2]]
        if sum(lens) == 0:
            break
        if max_waves and waves >= max_waves:
            aborted = True
            break
        if nw:
            for src, dst, cnt in handoff_plan(lens):
                if rank == src:
                    out = stepper.export_open(cnt)
                    comm.send(np.asarray([out.shape[0]], dtype=np.int64), dst)
                    if out.shape[0]:
                        comm.send(out, dst)
                    handoffs += 1
                    sent += int(out.shape[0])
                elif rank == dst:
                    k = int(comm.recv(1, src)[0])
                    if k:
                        stepper.import_open(comm.recv(k * nw, src).reshape(k, nw))
                    handoffs += 1
                    received += k
        lb, top, more = stepper.wave()  # a rank with an empty fringe returns immediately (top = INT64_MIN)
        waves += 1
    if not aborted:
        stepper.finish()
        best_ub = lb
    # the solution travels once, from the lowest rank that holds one of the optimal value
    solution, value = None, None
    if hasattr(stepper, "best_solution"):
        owners = [r for r in range(world) if int(rows[r, 3]) == lb]
        if owners:
            owner = owners[0]
            value = lb
            if rank == owner:
                sol = stepper.best_solution() or []
                packed = np.asarray([(d.variable << 32) | (d.value & 0xFFFFFFFF) for d in sol], dtype=np.int64)
                for r in range(world):
                    if r != owner:
                        comm.send(np.asarray([packed.size], dtype=np.int64), r)
                        if packed.size:
                            comm.send(packed, r)
                solution = [(d.variable, d.value) for d in sol]
            else:
                k = int(comm.recv(1, owner)[0])
                packed = comm.recv(k, owner) if k else np.zeros(0, dtype=np.int64)
                solution = [(int(x >> 32), int(np.int32(np.uint32(x & 0xFFFFFFFF)))) for x in packed.tolist()]
    return {"best_lb": lb, "best_ub": best_ub, "waves": waves, "collectives": colls, "handoffs": handoffs, "nodes_sent": sent, "nodes_received": received,
            "is_exact": not aborted, "best_value": value, "solution": solution}


def _sharded_maximize_allreduce(stepper, rank: int, world: int, allreduce_max, max_waves: int = 0):
    """Round-1 protocol (kept for comparison runs): static deal, allreduce(max) of three int64 before and after every wave."""
    stepper.init(True)
    lb, top, more = stepper.wave()
    stepper.retain_share(rank, world)
    more = 1 if stepper.fringe_len() > 0 else 0
    waves, colls = 1, 0
    best_ub = None
    aborted = False
    while True:
        g_lb, g_top, g_more = allreduce_max([lb, I64_MIN, more])
        colls += 1
        if g_lb > lb:
            stepper.set_lower_bound(g_lb)
            lb = g_lb
        if not g_more:
            break
        if max_waves and waves >= max_waves:
            aborted = True
            break
        lb, top, more = stepper.wave()
        waves += 1
        g = allreduce_max([lb, top, more])
        colls += 1
        lb = max(lb, g[0])
        stepper.set_lower_bound(lb)
        best_ub = g[1] if g[1] != I64_MIN else best_ub
        more = 1 if stepper.fringe_len() > 0 else 0
    if not aborted:
        stepper.finish()
        best_ub = lb
    return {"best_lb": lb, "best_ub": best_ub, "waves": waves, "collectives": colls, "is_exact": not aborted}
