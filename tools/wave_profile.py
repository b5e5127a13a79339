"""Per-wave profile of Solver::maximize on BASELINE config 2 (stepwise API): device ms, host fringe ms, expanded nodes, compilations.
Writes gpurun_out/wave_profile.json."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ddo_b200 import FixedWidth, Misp, ParNoCachingSolverLel, gnp  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
CAP = int(sys.argv[2]) if len(sys.argv) > 2 else min(K, 64)
pb = Misp(gnp(500, 0.5, 1))
s = ParNoCachingSolverLel(pb, FixedWidth(10000), wave_size=K, batch_cap=CAP)
for rep in range(2):
    s.init(True)
    rows, prev = [], s.stats()
    t0 = time.perf_counter()
    while True:
        tw = time.perf_counter()
        lb, top, more = s.wave()
        st = s.stats()
        rows.append({"wall_ms": (time.perf_counter() - tw) * 1e3, "dev_ms": st["device_ms"] - prev["device_ms"], "fringe_ms": st["fringe_ms"] - prev["fringe_ms"],
                     "expanded": st["expanded"] - prev["expanded"], "comps": st["compilations"] - prev["compilations"], "lb": lb, "top": top, "fringe_len": s.fringe_len()})
        prev = st
        if not more:
            break
    s.finish()
    total = time.perf_counter() - t0
print(f"total {total:.2f}s waves {len(rows)} explored {s.explored()} expanded {int(prev['expanded'])} dev {prev['device_ms']:.0f} ms fringe {prev['fringe_ms']:.0f} ms")
tiny = [r for r in rows if r["expanded"] < 2e6]
wide = [r for r in rows if r["expanded"] >= 2e6]
for name, grp in (("tiny(<2M nodes)", tiny), ("wide(>=2M nodes)", wide)):
    print(name, "waves", len(grp), "dev ms", round(sum(r["dev_ms"] for r in grp)), "wall ms", round(sum(r["wall_ms"] for r in grp)), "expanded", int(sum(r["expanded"] for r in grp)),
          "fringe ms", round(sum(r["fringe_ms"] for r in grp)))
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/wave_profile.json").write_text(json.dumps(rows))
