"""Per-wave, per-kernel device time of one warm Solver::maximize of BASELINE config 2 (CUDA events around every launch: the launches are
serialised, so the shares -- not the absolute times -- are what to read).  usage: python tools/wave_kernel_profile.py [wave] [batch_cap]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddo_b200 import FixedWidth, Misp, ParNoCachingSolverLel, gnp  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
CAP = int(sys.argv[2]) if len(sys.argv) > 2 else 512
s = ParNoCachingSolverLel(Misp(gnp(500, 0.5, 1)), FixedWidth(10000), wave_size=K, batch_cap=CAP)
s.maximize()  # warm-up
s.mdd.set_profiling(True)
s.init(True)
prev = {k: v["ms"] for k, v in s.mdd.kernel_times().items()}
prev_st = s.stats()
w = 0
tot = {}
while True:
    lb, top, more = s.wave()
    w += 1
    kt = {k: v["ms"] for k, v in s.mdd.kernel_times().items()}
    st = s.stats()
    d = {k: kt[k] - prev[k] for k in kt}
    exp = st["expanded"] - prev_st["expanded"]
    for k in d:
        tot[k] = tot.get(k, 0) + d[k]
    if sum(d.values()) > 2.0:
        print(f"wave {w:3d} expanded {int(exp):10d} dev {st['device_ms'] - prev_st['device_ms']:7.1f} ms | " + " ".join(f"{k} {v:6.1f}" for k, v in d.items()))
    prev, prev_st = kt, st
    if not more:
        break
print("total", " ".join(f"{k} {v:6.1f}" for k, v in tot.items()))
