"""Generates tests/golden/config2_trajectory_k<K>.json: the branch-and-bound trajectory of the CPU oracle's wave solver on BASELINE config 2
(MISP G(500, 0.5) seed 1, FixedWidth(10000)) for a given wave size.  ~15 minutes of single-thread CPU per run.

  python tests/golden/make_trajectory.py 512            # LAST_EXACT_LAYER cutset -> config2_trajectory_k512.json
  python tests/golden/make_trajectory.py 2048 frontier   # FRONTIER cutset         -> config2_trajectory_fc_k2048.json
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle_lib as O  # noqa: E402
from ddo_b200.instances import gnp  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
FC = len(sys.argv) > 2 and sys.argv[2] == "frontier"
inst = gnp(500, 0.5, 1)
r = O.OracleMisp(inst).solve("wave", k=K, width=10000, trace_cap=1 << 16, cutset_type=O.FRONTIER if FC else O.LEL)
out = {"instance": "gnp(500, 0.5, seed=1)", "width": 10000, "wave_size": K, "cutset": "frontier" if FC else "last exact layer",
       **{k: int(r[k]) for k in ("best_value", "best_lb", "best_ub", "is_exact", "explored", "expanded", "transitions", "compilations", "waves")},
       "solution": r["solution"], "oracle_seconds": r["seconds"],
       "trace_best_lb_fringe_len": [[int(t[1]), int(t[2])] for t in r["trace"]]}
(Path(__file__).resolve().parent / f"config2_trajectory_{'fc_' if FC else ''}k{K}.json").write_text(json.dumps(out) + "\n")
print("written", {k: out[k] for k in ("explored", "expanded", "transitions", "compilations", "waves", "best_value")})
