"""The reference's example executables rebuilt over the C ABI (ddo_b200/csrc/cli.cpp -> ddo_b200/bin/misp, ddo_b200/bin/max2sat):
argument handling, instance parsing and loud failure without a device on CPU; known optima and the report format on the GPU."""
import json
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "ddo_b200" / "bin"


def _run(exe, *args, timeout=300):
    import __graft_entry__ as g

    g.build()
    return subprocess.run([str(BIN / exe), *map(str, args)], capture_output=True, text=True, timeout=timeout)


def test_cli_usage_and_parse_errors(tmp_path):
    assert _run("misp").returncode == 2                                   # clap: the file name is required
    assert _run("max2sat", "-w", "3").returncode == 2                     # --file is required (max2sat/main.rs:37-39)
    assert _run("misp", "x.clq", "-w", "abc").returncode == 2
    r = _run("misp", tmp_path / "missing.clq")
    assert r.returncode == 2 and "io error" in r.stderr                   # Error::Io, misp/main.rs:245-246
    bad = tmp_path / "bad.clq"
    bad.write_text("c a comment\np edge 3 1\ne 1 2\nthis is not an instance\n")
    r = _run("misp", bad)
    assert r.returncode == 2 and "ill formed instance" in r.stderr        # Error::Format, misp/main.rs:314


def test_cli_fails_loudly_without_a_device(golden_dir):
    from ddo_b200 import _native as N

    if N.lib().ddo_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run("misp", golden_dir / "misp" / "keller4.clq", "-w", "100")
    assert r.returncode == 2 and "no CUDA device" in r.stderr and "Objective" not in r.stdout
    r = _run("max2sat", "-f", golden_dir / "max2sat" / "pass.wcnf")
    assert r.returncode == 2 and "no CUDA device" in r.stderr


def _report(stdout):
    keys = [ln.split(":")[0] for ln in stdout.strip().splitlines()]
    vals = {ln.split(":", 1)[0]: ln.split(":", 1)[1].strip() for ln in stdout.strip().splitlines()}
    return keys, vals


@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [("brock200_2", ["-w", "100"]), ("johnson8-4-4", []), ("hamming6-4", ["-t", "4", "--cutset", "frontier"]), ("MANN_a9", ["-w", "20", "-d", "60"])])
def test_cli_misp_known_optima_and_report(golden_dir, name, args):
    """examples/misp/tests.rs optima through the executable; the report is the reference's seven lines (misp/main.rs:391-397)."""
    exp = json.loads((golden_dir / "expected.json").read_text())["misp"][name]["optimum"]
    r = _run("misp", golden_dir / "misp" / f"{name}.clq", *args)
    assert r.returncode == 0, r.stderr
    keys, vals = _report(r.stdout)
    assert keys == ["Duration", "Objective", "Upper Bnd", "Lower Bnd", "Gap", "Aborted", "Solution"]
    assert re.fullmatch(r"\d+\.\d{3} seconds", vals["Duration"])
    assert int(vals["Objective"]) == int(vals["Upper Bnd"]) == int(vals["Lower Bnd"]) == exp
    assert vals["Gap"] == "0.000" and vals["Aborted"] == "false"
    sol = json.loads(vals["Solution"])
    assert len(sol) == exp and sol == sorted(sol)
    edges = set()
    for ln in (golden_dir / "misp" / f"{name}.clq").read_text().splitlines():
        if ln.startswith("e "):
            a, b = map(int, ln.split()[1:3])
            edges.add((a - 1, b - 1)); edges.add((b - 1, a - 1))
    assert all((a, b) not in edges for a in sol for b in sol if a != b)


@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [("pass", []), ("debug2", ["-w", "2"]), ("frb10-6-1", ["-w", "100", "-t", "120"])])
def test_cli_max2sat_known_optima_and_report(golden_dir, name, args):
    """examples/max2sat/tests.rs optima; eight-line report with the cost of the violated clauses (max2sat/main.rs:66-73)."""
    exp = json.loads((golden_dir / "expected.json").read_text())["max2sat"][name]["optimum"]
    path = golden_dir / "max2sat" / f"{name}.wcnf"
    r = _run("max2sat", "--file", path, *args)
    assert r.returncode == 0, r.stderr
    keys, vals = _report(r.stdout)
    assert keys == ["Duration", "Objective", "Upper Bnd", "Lower Bnd", "Gap", "Aborted", "Cost", "Solution"]
    assert int(vals["Objective"]) == int(vals["Upper Bnd"]) == int(vals["Lower Bnd"]) == exp and vals["Aborted"] == "false"
    lits = json.loads(vals["Solution"])
    assert [abs(x) for x in lits] == list(range(1, len(lits) + 1))
    # objective + cost = total weight of the (deduplicated) clauses
    from ddo_b200 import read_wcnf

    uniq = {}
    for w, x, y in read_wcnf(path).clauses.tolist():
        uniq[(min(x, y), max(x, y))] = w
    assert exp + int(vals["Cost"]) == sum(uniq.values())
