"""Debug: device layer trace of one DD vs the oracle (first differing layer)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle_lib as O
from ddo_b200 import CompilationType, GpuMdd, Misp, SubProblem, gnp
n, p, seed, W = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lb = int(sys.argv[6]) if len(sys.argv) > 6 else -(1 << 63)
ct = {"r": (CompilationType.Restricted, O.RESTRICTED), "x": (CompilationType.Relaxed, O.RELAXED)}[sys.argv[5] if len(sys.argv) > 5 else "r"]
inst = gnp(n, p, seed)
pb = Misp(inst); mdd = GpuMdd(pb, max(W, 400), 1)
c = mdd.compile(ct[0], W, SubProblem(pb.initial_state(), 0), best_lb=lb)
v, w = mdd.layer_trace(0)
ref = O.OracleMisp(inst).compile(ct[1], W, best_lb=lb)
rv, rw = ref["layer_vars"].tolist(), ref["layer_widths"].tolist()
print(f"W={W} {sys.argv[5] if len(sys.argv) > 5 else chr(114)} lb={lb}", "device expanded", c.expanded, "oracle", ref["expanded"], "best", c.best_value, ref["best_value"], "layers", len(v), len(rv))
for t in range(min(len(v), len(rv))):
    if v[t] != rv[t] or w[t] != rw[t]:
        print("first difference at layer", t, "device (var,width)", v[t], w[t], "oracle", rv[t], rw[t]); break
else:
    print("traces equal")
