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


# ----------------------------------------------------------------------------------------------------------------------------------
# Asynchronous variant (opt-in; `bench.py --async-shards`): no per-wave collective at all.
#
# The ranks of ONE box share a status board in POSIX shared memory -- the closest thing to the reference's `Mutex<Critical>`
# (parallel.rs:32-81) between processes: every rank publishes its incumbent, the size of its fringe and whether it is idle, and reads
# the others' words whenever it likes.  Nobody waits for a slower rank (profiles/r02_rank_timeline_*: with the blocking gather a rank
# spends up to two thirds of a solve waiting at it).  A rank whose fringe ran dry asks the fullest rank for work; the donor answers
# between two of its waves by writing packed open nodes (export_open) into the requester's mailbox.  Termination (parallel.rs:512): all
# ranks idle, no request or mail pending, nodes sent == nodes received, seen twice in a row by rank 0.
# ----------------------------------------------------------------------------------------------------------------------------------
class StatusBoard:
    """int64 words in shared memory.  header: [done, owner of the solution, solution length]; per rank RW words:
    [lb, open nodes, idle, sent, received, objective of the local solution, asks rank (-1 none), mailbox state (0 empty / 1 full), mailbox count];
    then one mailbox of `mail_nodes * node_words` words per rank and one solution area."""
    HDR, RW = 8, 16
    LB, LEN, IDLE, SENT, RECV, SOL, ASK, MAIL, MCNT, WAVES = range(10)

    def __init__(self, rank: int, world: int, node_words: int, n_vars: int, bootstrap, mail_nodes: int = 2048):
        import mmap
        import os
        self.rank, self.world, self.nw, self.mail_nodes = rank, world, node_words, mail_nodes
        self.sol_words = n_vars + 1
        words = self.HDR + world * self.RW + world * mail_nodes * node_words + self.sol_words
        # a file in /dev/shm mapped by every rank (plain mmap: multiprocessing.shared_memory would hand the segment to a resource tracker
        # that unlinks it when the first attached process exits)
        if rank == 0:
            self.path = f"/dev/shm/ddo_board_{os.getpid()}_{int.from_bytes(os.urandom(4), 'little'):08x}"
            fd = os.open(self.path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, words * 8)
            self.path = bootstrap(self.path)
        else:
            self.path = bootstrap(None)
            fd = os.open(self.path, os.O_RDWR)
        self.mm = mmap.mmap(fd, words * 8)
        os.close(fd)
        self.w = np.frombuffer(self.mm, dtype=np.int64)
        self.mail0 = self.HDR + world * self.RW
        self.sol0 = self.mail0 + world * mail_nodes * node_words
        if rank == 0:
            for r in range(world):
                self.set(r, self.LB, I64_MIN); self.set(r, self.SOL, I64_MIN); self.set(r, self.ASK, -1)
        bootstrap(None) if rank else bootstrap("ready")  # nobody reads the board before rank 0 initialised it

    def get(self, r, f): return int(self.w[self.HDR + r * self.RW + f])
    def set(self, r, f, v): self.w[self.HDR + r * self.RW + f] = v
    def col(self, f): return [self.get(r, f) for r in range(self.world)]
    def mailbox(self, r): return self.w[self.mail0 + r * self.mail_nodes * self.nw: self.mail0 + (r + 1) * self.mail_nodes * self.nw]

    def close(self):
        import os
        self.w = None
        try:
            self.mm.close()
        except BufferError:  # a view of the mailbox is still alive somewhere: the mapping goes with the process
            pass
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except FileNotFoundError:
                pass


def sharded_maximize_async(stepper, rank: int, world: int, board: StatusBoard, max_waves: int = 0, poll_s: float = 1e-4):
    """Returns dict(best_lb, best_ub, waves, handoffs, nodes_sent, nodes_received, is_exact, best_value, solution); no collectives."""
    import time
    B, me = board, rank
    stepper.init(True)
    lb, top, more = stepper.wave()  # the root DD: identical on every rank
    stepper.retain_share(rank, world)
    waves, handoffs, sent, received = 1, 0, 0, 0
    aborted = False
    nw = stepper.node_words()

    def publish(idle):
        sv = stepper.best_value()
        B.set(me, B.SOL, I64_MIN if sv is None else sv)
        B.set(me, B.LB, lb); B.set(me, B.LEN, stepper.fringe_len()); B.set(me, B.WAVES, waves)
        B.set(me, B.IDLE, 1 if idle else 0)

    def adopt():
        nonlocal lb
        g = max(B.col(B.LB))
        if g > lb:
            stepper.set_lower_bound(g); lb = g

    def serve():
        """answer the ranks that ask this one for work: half of what this rank holds above MIN_DONOR, or nothing"""
        nonlocal handoffs, sent
        for r in range(world):
            if r == me or B.get(r, B.ASK) != me or B.get(r, B.MAIL) != 0:
                continue
            have = stepper.fringe_len()
            cnt = min(B.mail_nodes, (have - MIN_DONOR) // 2) if have >= 2 * MIN_DONOR else 0
            k = 0
            if cnt > 0:
                out = stepper.export_open(cnt)
                k = int(out.shape[0])
                if k:
                    B.mailbox(r)[:k * nw] = np.ascontiguousarray(out, dtype=np.int64).reshape(-1)
            B.set(r, B.MCNT, k)
            sent += k
            B.set(me, B.SENT, sent)      # before the mail becomes visible: sent >= received at all times
            B.set(r, B.MAIL, 1)
            handoffs += 1 if k else 0

    publish(False)
    while True:
        if B.w[0]:
            break
        adopt()
        serve()
        if stepper.fringe_len() == 0:
            if B.get(me, B.MAIL) == 1:  # the answer to this rank's request
                k = B.get(me, B.MCNT)
                if k:
                    B.set(me, B.IDLE, 0)  # active BEFORE the nodes count as received (termination test below)
                    rows = np.array(B.mailbox(me)[:k * nw], dtype=np.int64).reshape(k, nw)
                    stepper.import_open(rows)
                    received += k
                    B.set(me, B.RECV, received)
                    handoffs += 1
                B.set(me, B.ASK, -1)
                B.set(me, B.MAIL, 0)
                if k:
                    publish(False)
                    continue
            publish(True)
            if B.get(me, B.ASK) == -1 and B.get(me, B.MAIL) == 0:
                lens = B.col(B.LEN)
                donor = max((r for r in range(world) if r != me and not B.get(r, B.IDLE) and lens[r] >= 2 * MIN_DONOR), key=lambda r: (lens[r], -r), default=-1)
                if donor >= 0:
                    B.set(me, B.ASK, donor)
            if me == 0:  # termination: two identical quiet scans in a row (counters first, flags after)
                def scan():
                    s, r = B.col(B.SENT), B.col(B.RECV)
                    quiet = sum(s) == sum(r) and all(B.col(B.IDLE)) and all(a == -1 for a in B.col(B.ASK)) and not any(B.col(B.MAIL))
                    return quiet, (tuple(s), tuple(r))
                q1, c1 = scan()
                if q1:
                    q2, c2 = scan()
                    if q2 and c1 == c2:
                        B.w[0] = 1
                        break
            time.sleep(poll_s)
            continue
        publish(False)
        if max_waves and waves >= max_waves:
            aborted = True
            B.w[0] = 2  # a rank that ran out of budget stops everybody
            break
        lb, top, more = stepper.wave()
        waves += 1
        publish(False)
    aborted = aborted or int(B.w[0]) == 2
    adopt()
    if not aborted:
        stepper.finish()
    publish(True)
    # the solution: the lowest rank holding one of the optimal value writes it to the board
    solution, value = None, None
    deadline = time.time() + 60
    while True:  # every rank has published its final words once all of them show the final bound or are idle
        if all(B.col(B.IDLE)) or time.time() > deadline:
            break
        time.sleep(poll_s)
    g_lb = max(B.col(B.LB))
    owners = [r for r in range(world) if B.get(r, B.SOL) == g_lb]
    if owners and hasattr(stepper, "best_solution"):
        owner, value = owners[0], g_lb
        if me == owner:
            sol = stepper.best_solution() or []
            packed = np.asarray([(d.variable << 32) | (d.value & 0xFFFFFFFF) for d in sol], dtype=np.int64)
            B.w[B.sol0 + 1: B.sol0 + 1 + packed.size] = packed
            B.w[B.sol0] = packed.size
            B.w[2] = 1  # published
            solution = [(d.variable, d.value) for d in sol]
        else:
            while not B.w[2] and time.time() < deadline:
                time.sleep(poll_s)
            k = int(B.w[B.sol0])
            solution = [(int(x >> 32), int(np.int32(np.uint32(x & 0xFFFFFFFF)))) for x in B.w[B.sol0 + 1: B.sol0 + 1 + k].tolist()]
    B.set(me, B.WAVES, -waves)  # (this rank is done with the board)
    return {"best_lb": g_lb, "best_ub": g_lb if not aborted else None, "waves": waves, "collectives": 0, "handoffs": handoffs, "nodes_sent": sent,
            "nodes_received": received, "is_exact": not aborted, "best_value": value, "solution": solution}
