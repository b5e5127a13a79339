"""BASELINE config 5: MISP G(1000, 0.5), width sweep 1k..100k, time-boxed Solver::maximize per width (HBM-roofline scan).
One JSON line per width: nodes expanded / s (device and end-to-end), algorithmic GB/s of the whole step against the measured HBM peak."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ddo_b200 import FixedWidth, Misp, ParNoCachingSolverLel, gnp  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
widths = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1000, 2000, 5000, 10000, 20000, 50000, 100000]
mem_budget = float(sys.argv[3]) * 1e9 if len(sys.argv) > 3 else 6e9  # device memory the DD arenas may take (GB)
peak = 6530.6
pk = ROOT / "MEASURED_PEAKS.json"
if pk.exists():
    peak = float(json.loads(pk.read_text())["hbm_gbs"])
inst = gnp(1000, 0.5, 1)
pb = Misp(inst)
for w in widths:
    cap = max(2, min(256, int(mem_budget // (w * 1001 * 20))))  # ~20 B of logs per node and layer per DD slot
    s = ParNoCachingSolverLel(pb, FixedWidth(w), wave_size=2048, batch_cap=cap)
    s.maximize(max_waves=2)  # warm-up (allocations, first launches)
    t0 = time.perf_counter()
    c = s.maximize(time_budget_s=budget)
    dt = time.perf_counter() - t0
    st = s.stats()
    cbar = st["transitions"] / max(st["expanded"], 1)
    b_node = (128 + 8) + cbar * (128 + 16)
    print(json.dumps({"width": w, "batch_cap": cap, "finished": bool(c.is_exact), "best_lb": s.best_lower_bound(), "best_ub": s.best_upper_bound(),
                      "explored": s.explored(), "expanded": int(st["expanded"]), "wall_s": round(dt, 3), "device_s": round(st["device_ms"] / 1e3, 3),
                      "nodes_per_s_e2e": st["expanded"] / dt, "nodes_per_s_device": st["expanded"] / (st["device_ms"] / 1e3),
                      "bytes_per_node": round(b_node, 1), "algorithmic_GBs_device": st["expanded"] * b_node / (st["device_ms"] / 1e3) / 1e9,
                      "hbm_frac": st["expanded"] * b_node / (st["device_ms"] / 1e3) / 1e9 / peak}), flush=True)
    s.close()
    s.mdd.close()
