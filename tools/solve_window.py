"""Waves [first, first + n) of one config-2 solve inside a cudaProfilerStart/Stop range (ncu --profile-from-start off): the window of the
batched regime (wave 5: 2048 sub-problems in lock-step) for the launch list and the --set full captures.
usage: solve_window.py [wave] [cap] [first_wave] [n_waves]"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddo_b200 import FixedWidth, Misp, ParNoCachingSolverLel, gnp  # noqa: E402

wave = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
first = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n = int(sys.argv[4]) if len(sys.argv) > 4 else 1
s = ParNoCachingSolverLel(Misp(gnp(500, 0.5, 1)), FixedWidth(10000), wave_size=wave, batch_cap=cap)
rt = ctypes.CDLL("libcudart.so.12")
s.init(True)
for _ in range(first - 1):
    s.wave()
e0 = int(s.stats()["expanded"])
rt.cudaProfilerStart()
for _ in range(n):
    s.wave()
rt.cudaProfilerStop()
print(f"window waves {first}..{first + n - 1}: expanded {int(s.stats()['expanded']) - e0} lb {s.best_lower_bound()} fringe {s.fringe_len()}")
