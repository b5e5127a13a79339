"""Profiling target: the root restricted + relaxed DD of BASELINE config 2 (one DD, width 10 000, 501 layers each).
Used under ncu (see profiles/README.md); prints the device time of both compilations when run alone."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddo_b200 import CompilationType, GpuMdd, Misp, SubProblem, gnp  # noqa: E402

n_dd = int(sys.argv[1]) if len(sys.argv) > 1 else 1
inst = gnp(500, 0.5, 1)
pb = Misp(inst)
mdd = GpuMdd(pb, 10000, n_dd)
roots = [SubProblem(pb.initial_state(), 0)] * n_dd
mdd.stage_roots([10000] * n_dd, roots)
for ct, name in ((CompilationType.Restricted, "restricted"), (CompilationType.Relaxed, "relaxed")):
    ms = mdd.compile_staged(n_dd, ct, 12 if ct == CompilationType.Relaxed else -(1 << 63))
    c = mdd.fetch_completions(n_dd)
    print(f"{name}: {ms:.2f} ms, {sum(x.expanded for x in c)} nodes expanded, {sum(x.expanded for x in c) / ms / 1e3:.1f} M nodes/s, best {c[0].best_value}")
