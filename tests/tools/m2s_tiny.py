"""Tiny MAX2SAT parity run incl. the FRONTIER-cutset kernels (target of compute-sanitizer)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from ddo_b200 import random_max2sat  # noqa: E402
from parity_util import check_instance  # noqa: E402

n = check_instance(random_max2sat(16, 60, 3), [1, 3, 8], model="m2s")
import oracle_lib as O  # noqa: E402

nf = check_instance(random_max2sat(16, 60, 3), [1, 3, 8], model="m2s", cutset_type=O.FRONTIER)  # m2_fc_sweep / m2_fc_eval / m2_fc_write
print("tiny m2s parity ok", n, nf)
