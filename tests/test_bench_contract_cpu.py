"""bench.py's contract, as far as it can be checked without a GPU: the golden check refuses a wrong trajectory, the CPU leg prints ONE JSON
line on stdout (libraries that write to file descriptor 1 are redirected), the product arm fails loudly without a CUDA device."""
import json
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_golden_check_accepts_the_oracle_trajectory_and_refuses_anything_else():
    import bench

    g = json.loads((ROOT / "tests" / "golden" / "config2_trajectory_k2048.json").read_text())
    wl, args = SimpleNamespace(kind="misp"), SimpleNamespace(wave=2048)
    good = {"best_lb": g["best_lb"], "best_ub": g["best_ub"], "is_exact": True}
    out = bench.check_against_golden(wl, args, 1, good, g["explored"], g["expanded"])
    assert out["checked"] and out["explored"] == g["explored"] and out["expanded"] == g["expanded"]
    with pytest.raises(RuntimeError):  # one node off
        bench.check_against_golden(wl, args, 1, good, g["explored"], g["expanded"] + 1)
    with pytest.raises(RuntimeError):  # wrong objective
        bench.check_against_golden(wl, args, 1, dict(good, best_lb=g["best_lb"] - 1), g["explored"], g["expanded"])
    with pytest.raises(RuntimeError):  # not proven
        bench.check_against_golden(wl, args, 1, dict(good, is_exact=False), g["explored"], g["expanded"])
    # several GPUs: objective and bound are checked, the rank-count dependent counters are not
    assert bench.check_against_golden(wl, args, 8, good, g["explored"] + 13, g["expanded"] * 1.1)["checked"]
    assert not bench.check_against_golden(SimpleNamespace(kind="misp"), SimpleNamespace(wave=77), 1, good, 0, 0)["checked"]


def test_cpu_leg_prints_one_json_line_on_stdout():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--cpu-baseline-only", "--cpu-seconds", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["kind"] == "port" and d["unit"] == "nodes/s" and d["cores"] >= 1 and d["value"] > 0


def test_product_arm_fails_loudly_without_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-config3"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr or "CUDA" in r.stderr
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]  # no number without the device
