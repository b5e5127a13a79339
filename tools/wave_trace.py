"""Per-wave trace of one warm Solver::maximize of BASELINE config 2 (DDO_WAVE_TRACE): where the wall time of a solve goes.
Columns: wave, popped, general DDs, inexact, fast-path ms, general ms, layer steps, expanded, fringe size, pop ms, wave wall ms."""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddo_b200 import FixedWidth, Misp, ParNoCachingSolverLel, gnp  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/wave_trace.txt"
pb = Misp(gnp(500, 0.5, 1))
s = ParNoCachingSolverLel(pb, FixedWidth(10000), wave_size=2048, batch_cap=512)
s.maximize()  # warm-up
os.environ["DDO_WAVE_TRACE"] = out
s2 = ParNoCachingSolverLel(pb, FixedWidth(10000), wave_size=2048, mdd=s.mdd)
t0 = time.perf_counter()
s2.maximize()
print(f"solve {time.perf_counter() - t0:.3f}s lb {s2.best_lower_bound()} explored {s2.explored()} stats {s2.stats()}")
