"""Tiny MISP parity run incl. dual mode, the shared-memory fast path and the FRONTIER-cutset kernels (target of compute-sanitizer)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402
from ddo_b200 import FixedWidth, Misp, ParNoCachingSolverFc, ParNoCachingSolverLel, gnp  # noqa: E402
from parity_util import check_instance  # noqa: E402

n = check_instance(gnp(70, 0.3, 3), [1, 4, 16], check_paths=True)
inst = gnp(48, 0.25, 9)
s = ParNoCachingSolverLel(Misp(inst), FixedWidth(6), wave_size=8)
c = s.maximize()
ref = O.OracleMisp(inst).solve("wave", k=8, width=6)
assert c.best_value == ref["best_value"] and s.explored() == ref["explored"]
nf = check_instance(gnp(70, 0.3, 3), [1, 4, 16], check_paths=True, cutset_type=O.FRONTIER)  # k_fc_sweep / k_fc_eval / k_fc_count / k_fc_write
sf = ParNoCachingSolverFc(Misp(inst), FixedWidth(6), wave_size=8)
cf = sf.maximize()
reff = O.OracleMisp(inst).solve("wave", k=8, width=6, cutset_type=O.FRONTIER)
assert cf.best_value == reff["best_value"] and sf.explored() == reff["explored"]
print("tiny misp parity ok", n, nf, c.best_value)
