"""Timing of wave 1 of BASELINE config 2 (root restricted DD + its relaxed twin, dual mode) -- warm."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddo_b200 import FixedWidth, Misp, ParNoCachingSolverLel, gnp
s = ParNoCachingSolverLel(Misp(gnp(500, 0.5, 1)), FixedWidth(10000), wave_size=2048, batch_cap=512)
for rep in range(3):
    s.init(True); t0 = time.perf_counter(); s.wave(); print("wave1 wall ms %.2f device ms %.2f" % ((time.perf_counter() - t0) * 1e3, s.stats()["device_ms"]))
