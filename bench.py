#!/usr/bin/env python
"""bench.py -- MDD nodes expanded / s on BASELINE.json's headline configuration (MISP G(500, 0.5), width 10 000).

One STEP = one `Solver::maximize()` of the instance to proven optimality (solver.rs:56): the root restricted + relaxed DD (width 10 000,
500 layers) and then every open sub-problem of the branch-and-bound, in waves of `--wave` DDs compiled in lock-step on the device.
`value` = nodes expanded (the rough-upper-bound test of clean.rs:365 passed) / WALL CLOCK of ddo_solver_maximize, max over ranks (SURVEY.md
section 8(d): instance upload excluded, fringe and collectives included; the instance is resident in HBM when the timed region starts);
`device_value` = the same count / device time of the compilations (CUDA events on the engine's stream); `e2e` = through the reference-
facing call with host buffers (for this path the same call: every wave's roots go H2D, completions and cutsets D2H inside the timed
region).  N > 1: the fringe is sharded over the ranks (ddo_b200/sharded.py).  Before printing, the line's objective / bound / explored /
expanded are checked against the committed oracle trajectory of this exact configuration (tests/golden/config2_trajectory_k<wave>.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--wave B] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "MDD nodes expanded/s, MISP n=500 width=10000"
UNIT = "nodes/s"
N_VERT, P_EDGE, SEED, WIDTH = 500, 0.5, 1, 10000
# --workload max2sat: BASELINE config 3 (a secondary line; the headline metric is config 2 above)
M2S_METRIC = "MDD nodes expanded/s, MAX2SAT 500 vars / 3000 clauses width=5000"
M2S_VARS, M2S_CLAUSES, M2S_WIDTH = 500, 3000, 5000


class Workload:
    """What differs between the two measured configurations: instance, device model, oracle, the work of one step."""

    def __init__(self, args):
        self.kind = args.workload
        if self.kind == "misp":
            from ddo_b200.instances import gnp
            self.inst = gnp(N_VERT, P_EDGE, SEED)
            self.metric, self.width = METRIC, WIDTH
            self.desc = f"MISP G({N_VERT},{P_EDGE}) seed {SEED}, FixedWidth({WIDTH}), LEL cutset, NoDupFringe/MaxUB"
            self.step_desc = "Solver::maximize to proven optimality"
            self.max_waves = 0
            self.scaling = "strong"  # the whole search is fixed: N GPUs share it
            self.dtype = "u64 bitset / i32 value"
        else:
            from ddo_b200.instances import random_max2sat
            self.inst = random_max2sat(M2S_VARS, M2S_CLAUSES, SEED)
            self.metric, self.width = M2S_METRIC, M2S_WIDTH
            self.desc = f"MAX2SAT random {M2S_VARS} vars / {M2S_CLAUSES} clauses seed {SEED}, FixedWidth({M2S_WIDTH}), LEL cutset, NoDupFringe/MaxUB"
            self.max_waves = args.max_waves or 2
            self.step_desc = f"Solver::maximize cut off after {self.max_waves} waves (the instance is far beyond proof of optimality; every step repeats the same deterministic search prefix)"
            self.scaling = "weak"  # every rank runs max_waves waves of its own shard of the fringe: per-GPU work is fixed as N grows
            self.dtype = "i32 benefit vector / i32 value"

    def problem(self, device):
        from ddo_b200 import Max2Sat, Misp
        return Misp(self.inst, device=device) if self.kind == "misp" else Max2Sat(self.inst, device=device)

    def oracle(self):
        import oracle_lib as O
        return O.OracleMisp(self.inst) if self.kind == "misp" else O.OracleM2s(self.inst)

    def state_bytes(self, pb):
        return pb.words * 8 if self.kind == "misp" else 4 * self.inst.n


def load_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ddo_b200 import FixedWidth, ParNoCachingSolverLel, kernel_launches
    from ddo_b200.sharded import NativeComm, sharded_maximize

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = Workload(args)
    inst = wl.inst
    pb = wl.problem(local_rank)
    solver = ParNoCachingSolverLel(pb, FixedWidth(wl.width), wave_size=args.wave, batch_cap=args.batch_cap)
    sampler = ClockSampler(local_rank)
    comm = None

    def bootstrap(raw):
        box = [raw]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    if world > 1:  # the search's own collectives go through the C ABI (ddo_comm_*: NCCL from C++); torch.distributed only bootstraps the id and times
        comm = NativeComm(rank, world, local_rank, bootstrap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """one STEP = Solver::maximize() of the instance to proven optimality (every rank works on its shard of the fringe)."""
        s0 = solver.stats()
        t0 = time.perf_counter()
        if world == 1:
            comp = solver.maximize(max_waves=wl.max_waves)
            res = {"best_lb": solver.best_lower_bound(), "best_ub": solver.best_upper_bound(), "is_exact": comp.is_exact}
        elif args.async_shards:  # opt-in: status board in shared memory, no per-wave collective (ddo_b200/sharded.py::sharded_maximize_async)
            from ddo_b200.sharded import StatusBoard, sharded_maximize_async
            board = StatusBoard(rank, world, solver.node_words(), pb.nb_variables(), bootstrap)
            res = sharded_maximize_async(solver, rank, world, board, max_waves=wl.max_waves)
            board.close()
            res = {k: res[k] for k in ("best_lb", "best_ub", "is_exact", "handoffs", "nodes_sent", "collectives")}
        else:
            res = solver.maximize_sharded(comm, max_waves=wl.max_waves)  # ddo_solver_maximize_sharded: the whole protocol in one native call per rank
            res = {k: res[k] for k in ("best_lb", "best_ub", "is_exact", "handoffs", "nodes_sent", "collectives")}
        wall = time.perf_counter() - t0
        s1 = solver.stats()
        # init() resets the counters, so s1 holds this step only (bytes are cumulative engine counters)
        return {"wall": wall, "dev_ms": s1["device_ms"], "expanded": s1["expanded"], "transitions": s1["transitions"], "waves": s1["waves"],
                "compilations": s1["compilations"], "fringe_ms": s1["fringe_ms"], "h2d": s1["bytes_h2d"] - s0["bytes_h2d"], "d2h": s1["bytes_d2h"] - s0["bytes_d2h"],
                "explored": solver.explored(), **res}

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = kernel_launches()
    sampler.start()
    t0 = time.perf_counter()
    steps = [step() for _ in range(args.steps)]
    barrier()
    wall_total = time.perf_counter() - t0
    launches = kernel_launches() - launches0
    clocks = sampler.stop()
    last = steps[-1]
    # per-kernel profile pass (CUDA events around every launch; separate from the timed passes), rank 0 / single GPU only
    kt = None
    if world == 1:
        solver.mdd.set_profiling(True)
        step()
        kt = solver.mdd.kernel_times()
        solver.mdd.set_profiling(False)

    def allred_f(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=op); return float(t.item())

    dev_ms = allred_f(sum(s["dev_ms"] for s in steps), dist.ReduceOp.MAX if world > 1 else None)
    wall_max = allred_f(wall_total, dist.ReduceOp.MAX if world > 1 else None)
    expanded_all = allred_f(sum(s["expanded"] for s in steps), dist.ReduceOp.SUM if world > 1 else None)
    root_expanded = 0
    if world > 1:  # every rank compiles the (identical) root DD pair before the deal: redundant work, counted ONCE in the throughput
        solver.maximize(max_waves=1)
        root_expanded = int(solver.stats()["expanded"])
        expanded_all -= args.steps * (world - 1) * root_expanded
    root_transitions = int(solver.stats()["transitions"]) if world > 1 else 0
    transitions_all = allred_f(sum(s["transitions"] for s in steps), dist.ReduceOp.SUM if world > 1 else None) - args.steps * (world - 1) * root_transitions
    launches_all = allred_f(launches, dist.ReduceOp.SUM if world > 1 else None)
    explored_all = allred_f(last["explored"], dist.ReduceOp.SUM if world > 1 else None)
    h2d_all = allred_f(last["h2d"], dist.ReduceOp.SUM if world > 1 else None)
    d2h_all = allred_f(last["d2h"], dist.ReduceOp.SUM if world > 1 else None)
    cfg3 = None
    if wl.kind == "misp" and not args.no_config3:  # every rank takes part (the fringe of config 3 is sharded like the main workload's)
        cfg3 = second_workload_line(args, local_rank, world, comm)
    if rank != 0:
        dist.destroy_process_group()
        return
    golden = check_against_golden(wl, args, world, last, explored_all, expanded_all / args.steps)
    cbar = transitions_all / max(expanded_all, 1)
    S_bytes = wl.state_bytes(pb)
    b_node = (S_bytes + 8) + cbar * (S_bytes + 16)  # read the parent once, write each child once
    peak, peak_src = load_peaks()
    line = {
        "metric": wl.metric, "value": expanded_all / wall_max, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall_max * 1e3 / args.steps, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": wl.dtype,
        "data": "synthetic", "impl": "ddo_b200",
        "config": {"workload": f"{wl.step_desc}, {wl.desc}",
                   "wave_size": args.wave, "batch_cap": args.batch_cap, "objective": int(last["best_lb"]), "proven_upper_bound": int(last["best_ub"]), "is_exact": bool(last["is_exact"]),
                   "explored_subproblems": int(explored_all), "handoffs_rank0": last.get("handoffs"), "nodes_sent_rank0": last.get("nodes_sent"), "collectives_rank0": last.get("collectives"), "expanded_nodes_per_step": int(expanded_all / args.steps), "replicated_root_nodes_not_counted": (world - 1) * root_expanded, "waves_per_step_rank0": int(last["waves"]),
                   "l2": "no L2 flush: every step re-runs the whole search (thousands of launches over >10 GB of arenas), far beyond the 126 MB L2",
                   "parallelism": (f"fringe sharded over {world} GPU(s); asynchronous status board in shared memory, no per-wave collective, idle ranks ask the fullest rank for open nodes" if args.async_shards else f"fringe sharded over {world} GPU(s); one all-gather of 4 x int64 per rank per wave (ddo_comm_allgather, NCCL from the C ABI), open nodes handed from loaded to idle ranks point to point")},
        "device_value": expanded_all / (dev_ms * 1e-3), "device_ms_per_step": dev_ms / args.steps, "golden_check": golden,
        "e2e": {"value": expanded_all / wall_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                "note": "wall clock of ddo_solver_maximize (host fringe, H2D of every wave's roots, D2H of completions and cutsets included)"},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
        "host_fringe_ms_per_step": last["fringe_ms"],
    }
    if kt is not None:
        if wl.kind == "misp" and kt["k_finish"]["launches"] == 0:  # the persistent whole-DD kernel times in the first slot
            kt["k_dd"] = kt.pop("k_expand"); kt.pop("k_finish"); kt.pop("k_compact")
        if wl.kind != "misp":
            kt["m2_merge"] = kt.pop("k_small")  # the MAX2SAT engine times its merge kernels in that slot
        dom = max((k for k in kt if k not in ("k_finalize_bottomup", "k_drain")), key=lambda k: kt[k]["ms"])
        dom_gbs = last["expanded"] * b_node / (kt[dom]["ms"] * 1e-3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / ("r02_traffic.json" if wl.kind == "misp" else "r01_traffic_max2sat.json")
        if tf.exists():  # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full captures
            tj = json.loads(tf.read_text())
            for t in tj.get("entries", [tj]):
                if t["kernel"].startswith(dom) or t["kernel"].startswith(dom.replace("k_", "m2_")):
                    traffic = {"kernel": t["kernel"], "bytes_per_launch": t["dram_bytes_read"] + t["dram_bytes_write"], "context": t["context"]}
                    if "algorithmic_bytes" in t:
                        traffic["algorithmic_bytes_per_launch"] = t["algorithmic_bytes"]
                    break
        line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak, "traffic": traffic,
                            "peak_source": peak_src, "bytes_per_node": b_node, "mean_out_degree": cbar,
                            "kernel_ms_per_step": {k: round(v["ms"], 3) for k, v in kt.items()},
                            "kernel_launches_per_step": {k: v["launches"] for k, v in kt.items()},
                            "achieved_by_kernel": {k: round(last["expanded"] * b_node / (v["ms"] * 1e-3) / 1e9, 1) for k, v in kt.items() if v["ms"] > 0 and k not in ("k_finalize_bottomup", "k_drain")},
                            "whole_step_frac": (expanded_all / (dev_ms * 1e-3)) * b_node / 1e9 / peak, "whole_step_frac_wall": (expanded_all / wall_max) * b_node / 1e9 / peak,
                            "note": "achieved = expanded nodes of one step x bytes_per_node / summed CUDA-event duration of the dominant kernel's launches"}
    if cfg3 is not None:
        line["configs"] = {"config3_max2sat": cfg3}
    if not args.no_cpu_baseline and world == 1:  # (the CPU leg is reported at N = 1 only)
        line["cpu_baseline"] = cpu_baseline_subprocess(args)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def check_against_golden(wl, args, world, last, explored_all, expanded_per_step):
    """The bench's own configuration against the committed oracle trajectory (tests/golden/config2_trajectory_k<wave>.json, made by
    tests/golden/make_trajectory.py): objective and proven bound at every N; sub-problems explored and nodes expanded at N = 1, where the
    search is the oracle wave solver's.  A mismatch is an error, not a number."""
    if wl.kind != "misp":
        return {"checked": False, "why": "config 3 is far beyond a proof: parity is tested on DD digests (tests/golden/dd_digests.json)"}
    f = ROOT / "tests" / "golden" / f"config2_trajectory_k{args.wave}.json"
    if not f.exists():
        return {"checked": False, "why": f"no golden for wave {args.wave}"}
    g = json.loads(f.read_text())
    if (int(last["best_lb"]), int(last["best_ub"]), bool(last["is_exact"])) != (g["best_lb"], g["best_ub"], True):
        raise RuntimeError(f"bench result {last['best_lb']} / {last['best_ub']} differs from the oracle golden {g['best_lb']} / {g['best_ub']}")
    if world == 1 and (int(explored_all), int(expanded_per_step)) != (g["explored"], g["expanded"]):
        raise RuntimeError(f"bench trajectory (explored {explored_all}, expanded {expanded_per_step}) differs from the oracle golden ({g['explored']}, {g['expanded']})")
    return {"checked": True, "file": f.name, "objective": g["best_lb"], "explored": g["explored"] if world == 1 else None, "expanded": g["expanded"] if world == 1 else None}


def second_workload_line(args, local_rank, world=1, comm=None):
    """BASELINE config 3 (MAX2SAT 500 vars / 3000 clauses, W = 5000, "1->8 B200 fringe-sharded") measured in the same run, so that the driver's
    default invocation carries a number for it at every N: same definitions as the main line (value = expanded / wall clock of maximize(),
    device_value = / device time).  At N > 1 every rank compiles the root DD pair, keeps its share of the root's cutset and runs its own
    waves on it (ddo_solver_maximize_sharded): the per-GPU work is fixed, the total grows with N -- weak scaling.  The root pair is compiled
    by every rank but counted once."""
    import copy
    import torch
    import torch.distributed as dist
    from ddo_b200 import FixedWidth, ParNoCachingSolverLel
    a = copy.copy(args)
    a.workload, a.wave, a.batch_cap, a.max_waves = "max2sat", 148, 148, 0
    wl = Workload(a)
    pb = wl.problem(local_rank)
    solver = ParNoCachingSolverLel(pb, FixedWidth(wl.width), wave_size=a.wave, batch_cap=a.batch_cap)

    def run(max_waves):
        t0 = time.perf_counter()
        if world == 1:
            solver.maximize(max_waves=max_waves)
        else:
            solver.maximize_sharded(comm, max_waves=max_waves)
        return time.perf_counter() - t0, solver.stats()

    solver.maximize(max_waves=1)  # warm-up; also the size of the root DD pair (the same on every rank, no collective involved)
    st_root = solver.stats()
    run(wl.max_waves)
    exp = wall = dev = 0.0
    if world > 1:
        dist.barrier()
    for _ in range(2):
        w, st = run(wl.max_waves)
        wall += w; exp += st["expanded"]; dev += st["device_ms"]
    if world > 1:
        t = torch.tensor([exp], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM)
        exp = float(t.item()) - 2 * (world - 1) * st_root["expanded"]
        t = torch.tensor([wall, dev], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev = float(t[0].item()), float(t[1].item())
    S_bytes = wl.state_bytes(pb)
    b_node = (S_bytes + 8) + 2 * (S_bytes + 16)
    peak, _ = load_peaks()
    out = {"metric": wl.metric, "value": exp / wall, "unit": UNIT, "device_value": exp / (dev * 1e-3), "n_gpus": world, "scaling": "weak", "steps": 2, "warmup": 2,
           "ms_per_step": wall * 1e3 / 2, "expanded_nodes_per_step": exp / 2, "workload": f"{wl.step_desc}, {wl.desc}",
           "whole_step_frac": exp / (dev * 1e-3) * b_node / 1e9 / (peak * world), "bytes_per_node": b_node}
    solver.close() if hasattr(solver, "close") else None
    return out


def cpu_baseline_subprocess(args):
    """The CPU leg runs in its own process AFTER the timed region (it would otherwise dilute every GPU-busy sample of this process)."""
    cmd = [sys.executable, str(Path(__file__).resolve()), "--cpu-baseline-only", "--workload", args.workload, "--cpu-seconds", str(args.cpu_seconds)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return {"error": (r.stderr or r.stdout)[-300:]}


def cpu_baseline(wl, budget_s: float):
    """The CPU oracle's ParallelSolver ("restated reference": per-node heap states, hash-map dedup, full comparison sort for the width
    cut, one mutex-protected fringe) on all host cores, same instance and width, under a TimeBudget (heuristics/cutoff.rs:302-323) --
    a BOUNDED sample of the same search; the rate is expanded nodes / elapsed."""
    cores = os.cpu_count() or 1
    if wl.kind == "max2sat":
        return cpu_sample_max2sat(wl, cores, budget_s)
    r = wl.oracle().solve("parallel", k=cores, width=wl.width, time_budget_s=budget_s)
    return {"value": r["expanded"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"ParallelSolver({cores} threads) maximize() with TimeBudget({budget_s:g} s): {r['explored']} sub-problems explored, "
                      f"{r['expanded']} nodes expanded in {r['seconds']:.1f} s, lb={r['best_lb']} (finished={bool(r['is_exact'])})",
            "expanded": int(r["expanded"]), "seconds": r["seconds"]}


def cpu_sample_max2sat(wl, cores: int, budget_s: float):
    """MAX2SAT at W = 5000 takes the CPU path more than a minute per DD, and a ParallelSolver run starts with ONE open sub-problem (its other
    threads idle until the root DDs are done), so a bounded sample of `maximize()` would time a single thread.  Instead every thread gets work
    from the start: the exact cutset of a narrow relaxed root DD (W = cores) gives >= cores independent sub-problems of the same search,
    and `cores` workers compile their restricted + relaxed DDs at the full width under a TimeBudget (the ParallelSolver worker body,
    parallel.rs:391-437, without the fringe); layers expanded before the cutoff count."""
    import oracle_lib as O

    o = wl.oracle()
    r0 = o.compile(O.RELAXED, max(cores, 2))
    n = int(r0["cutset_size"])
    r = o.compile_many(r0["cutset_states"], r0["cutset_values"], r0["cutset_depths"], [wl.width] * n, O.I64_MIN, cores, time_budget_s=budget_s)
    return {"value": r["expanded"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} worker threads compiling restricted + relaxed DDs (W={wl.width}) of {n} independent depth-{int(r0['cutset_depths'][0])} sub-problems under "
                      f"TimeBudget({budget_s:g} s): {r['expanded']} nodes expanded in {r['seconds']:.1f} s",
            "expanded": int(r["expanded"]), "seconds": r["seconds"]}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  ddo is Rust and cannot be built in this image (no cargo /
    rustc), so this arm times the oracle port of ParallelSolver (oracle/), all host threads; each step is a time-boxed maximize()."""
    if rank != 0:
        return
    import oracle_lib as O
    from ddo_b200.instances import gnp

    wl = Workload(args)
    cores = os.cpu_count() or 1
    if wl.kind == "max2sat":
        exp, sec, smp = 0, 0.0, None
        budget = max(10.0, min(args.cpu_seconds, 200.0 / max(args.steps + args.warmup, 1)))
        for i in range(args.warmup + args.steps):
            smp = cpu_sample_max2sat(wl, cores, budget)
            if i >= args.warmup:
                exp += smp["expanded"]; sec += smp["seconds"]
        v = exp / sec
        line = {"metric": wl.metric, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i64 benefit vector / i64 value", "data": "synthetic", "impl": "reference",
                "config": {"workload": f"time-boxed DD compilations, {wl.desc}", "note": "oracle port of ddo's DD compilation (Rust toolchain absent); CPU only, rank 0 only"},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": smp["sample"]},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return
    o = wl.oracle()
    # every step is maximize() of the SAME instance and width under a TimeBudget of >= 30 s, long enough that the single-threaded root DDs
    # (W = 10 000) are a small part of it; on a box that finishes the proof inside the box the arm runs the same configuration as ours
    budget = max(30.0, min(90.0, 240.0 / max(args.steps, 1)))
    small = gnp(200, 0.5, SEED)
    os_ = O.OracleMisp(small)
    for _ in range(args.warmup):  # untimed: a small instance, warms the allocator and the thread pool
        os_.solve("parallel", k=cores, width=100, time_budget_s=2.0)
    exp, sec, last, finished = 0, 0.0, None, True
    for _ in range(args.steps):
        last = o.solve("parallel", k=cores, width=wl.width, time_budget_s=budget)
        exp += last["expanded"]; sec += last["seconds"]; finished = finished and bool(last["is_exact"])
    v = exp / sec
    sample = f"ParallelSolver({cores} threads) maximize() under TimeBudget({budget:g} s) per step ({'proof completed' if finished else 'cut off before the proof'})"
    line = {"metric": wl.metric, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": wl.dtype.replace("i32 value", "i64 value"), "data": "synthetic", "impl": "reference",
            "config": {"workload": f"Solver::maximize ({'to proven optimality' if finished else 'time-boxed'}), {wl.desc}",
                       "note": "C++ port of ddo's ParallelSolver (oracle/; the Rust toolchain is absent) -- kind 'port', never real ddo; CPU only, rank 0 only", "lb_at_cutoff": int(last["best_lb"])},
            "same_config": finished,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


_JSON_FD = None


def emit(obj):
    """The ONE JSON line of the run, on the process's original stdout."""
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # Libraries write to file descriptor 1 behind Python's back (NCCL prints its version banner there): everything but the JSON line goes to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--wave", type=int, default=2048, help="open sub-problems popped per wave and per GPU")
    ap.add_argument("--batch-cap", type=int, default=2048, help="DD slots of the general (layer-by-layer) engine: how many DDs a batch may hold in lock-step when the log pool allows (it holds 512 full-depth DDs; deeper sub-problems log fewer layers)")
    ap.add_argument("--cpu-seconds", type=float, default=30.0, help="TimeBudget of the CPU baseline sample")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="misp", choices=["misp", "max2sat"], help="misp = BASELINE config 2 (the headline metric); max2sat = config 3")
    ap.add_argument("--max-waves", type=int, default=0, help="max2sat: waves per step (default 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--async-shards", action="store_true", help="N > 1: the asynchronous status-board protocol instead of one all-gather per wave (opt-in, see DESIGN.md section 6)")
    ap.add_argument("--no-config3", action="store_true", help="skip the secondary MAX2SAT (config 3) measurement carried under 'configs'")
    ap.add_argument("--cpu-baseline-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.workload == "max2sat":  # 2 KB states: fewer, larger DDs per wave -- one DD per SM (m2_finish is one CTA per DD)
        if args.wave == 2048: args.wave = 148
        if args.batch_cap == 2048: args.batch_cap = 148
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.cpu_baseline_only:
        emit(cpu_baseline(Workload(args), args.cpu_seconds))
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
