import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O
from ddo_b200 import random_max2sat, Max2Sat, GpuMdd, SubProblem, CompilationType
n, m, seed, W, lb = [int(x) for x in sys.argv[1:6]]
inst = random_max2sat(n, m, seed)
o = O.OracleM2s(inst)
pb = Max2Sat(inst)
mdd = GpuMdd(pb, max(W, 2), 1)
root = SubProblem(pb.initial_state(), pb.initial_value())
c = mdd.compile(CompilationType.Relaxed, W, root, best_lb=lb)
r = o.compile(O.RELAXED, W, best_lb=lb, want_paths=True)
print("dev", c)
print("ref", {k: r[k] for k in ("has_best", "best_value", "is_exact", "has_best_exact", "best_exact_value", "lel", "expanded", "transitions", "cutset_size")})
print("dev layers", mdd.layer_trace(0))
print("ref layers", r["layer_vars"].tolist(), r["layer_widths"].tolist())
print("dev best", [(d.variable, d.value) for d in mdd.best_solution(0)] if c.best_value is not None else None)
print("ref best", r["best_solution"], "exact", r["best_exact_solution"])
