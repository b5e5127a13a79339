"""Fringe-sharded branch-and-bound over several GPUs of one box (one process per GPU).

SURVEY.md section 8(e): the open sub-problems are independent, so the path shards with NO data-path collective.  Every rank compiles the
root DD (identical, deterministic), keeps every `world`-th open node of the common MaxUB order (`retain_share`), and then runs waves on its
own fringe; after each wave ONE allreduce(max) of three int64 -- [best_lb, ub of the best open node, has_work] -- synchronises the incumbent
lower bound, the global proven upper bound and termination.  This replaces the mutex-protected `Critical` block of
ddo/src/implementation/solver/parallel.rs:32-81 (best_lb read at :398/:426, written at :446-453; termination test at :512).

`stepper` is anything with the stepwise solver interface (init / wave / set_lower_bound / retain_share / finish / getters): the device solver
(`ParNoCachingSolverLel`) in production, a CPU stand-in in the gloo tests.  `allreduce_max(list[int]) -> list[int]` is the only collective.
"""
from __future__ import annotations

from typing import Callable, List

I64_MIN = -(1 << 63)


def torch_allreduce_max(device=None) -> Callable[[List[int]], List[int]]:
    """allreduce(max) over torch.distributed (NCCL over NVLink when `device` is a CUDA device, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    def f(vals: List[int]) -> List[int]:
        t = torch.tensor(vals, dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [int(x) for x in t.tolist()]

    return f


def sharded_maximize(stepper, rank: int, world: int, allreduce_max: Callable[[List[int]], List[int]], max_waves: int = 0):
    """Returns dict(best_lb, best_ub, waves, collectives, is_exact)."""
    stepper.init(True)
    lb, top, more = stepper.wave()  # the root DD: identical on every rank
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
        lb, top, more = stepper.wave()  # a rank with an empty fringe returns immediately
        waves += 1
        # proven bound: the best open node anywhere before this wave (valid global ub while the search runs)
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
