"""One Solver::maximize of BASELINE config 2 (no warm-up: kernel launch indices are deterministic) -- the target of the ncu captures."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddo_b200 import FixedWidth, Misp, ParNoCachingSolverLel, gnp  # noqa: E402

wave = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 512
max_waves = int(sys.argv[3]) if len(sys.argv) > 3 else 0
s = ParNoCachingSolverLel(Misp(gnp(500, 0.5, 1)), FixedWidth(10000), wave_size=wave, batch_cap=cap)
t0 = time.perf_counter()
if max_waves:
    s.init(True)
    for _ in range(max_waves):
        s.wave()
else:
    s.maximize()
print(f"solve {time.perf_counter() - t0:.2f}s lb {s.best_lower_bound()} explored {s.explored()} stats {s.stats()}")
