"""BASELINE config 3 (MAX2SAT 500 vars / 3000 clauses, W = 5000): time-boxed Solver::maximize, per-kernel CUDA-event times."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddo_b200 import FixedWidth, Max2Sat, ParNoCachingSolverLel, random_max2sat  # noqa: E402

wave = int(sys.argv[1]) if len(sys.argv) > 1 else 32
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
width = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
prof = int(sys.argv[4]) if len(sys.argv) > 4 else 0
inst = random_max2sat(500, 3000, 1)
pb = Max2Sat(inst)
s = ParNoCachingSolverLel(pb, FixedWidth(width), wave_size=wave)
if prof:
    s.mdd.set_profiling(True)
t0 = time.perf_counter()
c = s.maximize(time_budget_s=budget)
dt = time.perf_counter() - t0
st = s.stats()
print(f"wall {dt:.2f}s dev {st['device_ms']:.0f} ms fringe {st['fringe_ms']:.0f} ms waves {int(st['waves'])} explored {s.explored()} compilations {int(st['compilations'])} "
      f"expanded {int(st['expanded'])} -> {st['expanded'] / dt / 1e6:.1f} M nodes/s e2e, {st['expanded'] / st['device_ms'] / 1e3:.1f} M nodes/s device; lb {s.best_lower_bound()} ub {s.best_upper_bound()} "
      f"h2d {st['bytes_h2d'] / 1e6:.1f} MB d2h {st['bytes_d2h'] / 1e6:.1f} MB")
if prof:
    print(s.mdd.kernel_times())
