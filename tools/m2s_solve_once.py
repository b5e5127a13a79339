"""A fixed number of waves of Solver::maximize on BASELINE config 3 (no warm-up: deterministic launch indices) -- target of the ncu captures."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddo_b200 import FixedWidth, Max2Sat, ParNoCachingSolverLel, random_max2sat  # noqa: E402

wave = int(sys.argv[1]) if len(sys.argv) > 1 else 64
max_waves = int(sys.argv[2]) if len(sys.argv) > 2 else 2
s = ParNoCachingSolverLel(Max2Sat(random_max2sat(500, 3000, 1)), FixedWidth(5000), wave_size=wave)
t0 = time.perf_counter()
s.maximize(max_waves=max_waves)
print(f"{time.perf_counter() - t0:.2f}s stats {s.stats()}")
