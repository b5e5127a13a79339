"""Provenance script for tests/golden (run in the build container, where /root/reference is mounted).

Copies the *data* fixtures the reference's own tests use for the hot path (no reference source code):
  * resources/visualisation_tests/*.dot  -- golden graphviz dumps compared verbatim by clean.rs:2401-2546
  * resources/misp/*.clq (small ones)    -- DIMACS instances whose optima are asserted in ddo/examples/misp/tests.rs:66-193
  * resources/knapsack/* (small ones)    -- instances whose optima are asserted in ddo/examples/knapsack/tests.rs:65-206
  * resources/max2sat/*.wcnf (small)     -- instances whose optima are asserted in ddo/examples/max2sat/tests.rs:65-103
  * resources/tsptw/{Langevin,SolomonPotvinBengio}/* (a selection) -- instances whose optima are asserted in ddo/examples/tsptw/tests.rs:80-683
and writes expected.json with the asserted optima (transcribed from those test files, with their line numbers).
"""
import json
import re
import shutil
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

MISP = ["johnson8-2-4", "hamming6-4", "hamming6-2", "MANN_a9", "johnson8-4-4", "keller4", "hamming8-2", "hamming8-4",
        "brock200_2", "brock200_3", "brock200_4", "c-fat200-5", "c-fat200-1", "c-fat200-2", "p_hat300-1"]
MISP_GZ = ["c-fat500-1", "c-fat500-2"]  # the two n = 500 instances of the non-ignored tests (1.1 MB of edges each): stored gzip-compressed
KNAPSACK_MAX_ITEMS = 500  # (the 1000- and 2000-item instances of the non-ignored tests take the CPU oracle 20 - 150 s each: checked once, not committed)
TSPTW = ["Langevin/N20ft301.dat", "Langevin/N20ft405.dat", "Langevin/N40ft403.dat", "Langevin/N60ft406.dat", "Langevin/N60ft410.dat",
         "SolomonPotvinBengio/rc_201.1.txt", "SolomonPotvinBengio/rc_201.3.txt", "SolomonPotvinBengio/rc_202.2.txt", "SolomonPotvinBengio/rc_203.1.txt",
         "SolomonPotvinBengio/rc_203.4.txt", "SolomonPotvinBengio/rc_205.1.txt", "SolomonPotvinBengio/rc_205.2.txt", "SolomonPotvinBengio/rc_205.4.txt",
         "SolomonPotvinBengio/rc_206.3.txt", "SolomonPotvinBengio/rc_207.4.txt"]
MAX2SAT = ["debug", "debug2", "pass", "tautology", "unit", "negative_wt", "frb10-6-1", "frb10-6-2", "frb10-6-3", "frb10-6-4"]


def asserted(tests_rs: Path):
    """{instance id: (value, line)} from `assert_eq!(solve_id("<id>"), <value>);` lines (ignored tests included)."""
    out = {}
    for ln, line in enumerate(tests_rs.read_text().splitlines(), 1):
        m = re.search(r'assert_eq!\(solve_id\("([^"]+)"\),\s*(-?[\d_]+)\)', line)
        if m:
            out[m.group(1)] = (int(m.group(2).replace("_", "")), ln)
    return out


def main():
    (OUT / "misp").mkdir(exist_ok=True)
    (OUT / "knapsack").mkdir(exist_ok=True)
    for f in (REF / "resources/visualisation_tests").glob("*.dot"):
        shutil.copy(f, OUT / f.name)
    (OUT / "max2sat").mkdir(exist_ok=True)
    expected = {"misp": {}, "knapsack": {}, "max2sat": {}}
    x = asserted(REF / "ddo/examples/max2sat/tests.rs")
    for name in MAX2SAT:
        shutil.copy(REF / "resources/max2sat" / f"{name}.wcnf", OUT / "max2sat" / f"{name}.wcnf")
        v, ln = x[f"{name}.wcnf"]
        expected["max2sat"][name] = {"optimum": v, "source": f"ddo/examples/max2sat/tests.rs:{ln}"}
    m = asserted(REF / "ddo/examples/misp/tests.rs")
    for name in MISP:
        shutil.copy(REF / "resources/misp" / f"{name}.clq", OUT / "misp" / f"{name}.clq")
        v, ln = m[f"{name}.clq"]
        expected["misp"][name] = {"optimum": v, "source": f"ddo/examples/misp/tests.rs:{ln}"}
    for name in MISP_GZ:
        import gzip
        (OUT / "misp" / f"{name}.clq.gz").write_bytes(gzip.compress((REF / "resources/misp" / f"{name}.clq").read_bytes(), 9, mtime=0))
        v, ln = m[f"{name}.clq"]
        expected["misp"][name] = {"optimum": v, "source": f"ddo/examples/misp/tests.rs:{ln}"}
    k = asserted(REF / "ddo/examples/knapsack/tests.rs")
    for name, (v, ln) in sorted(k.items()):
        src = REF / "resources/knapsack" / name
        n_items = int(src.read_text().split()[0])
        if n_items <= KNAPSACK_MAX_ITEMS:
            shutil.copy(src, OUT / "knapsack" / name)
            expected["knapsack"][name] = {"optimum": v, "items": n_items, "source": f"ddo/examples/knapsack/tests.rs:{ln}"}
    # TSPTW: `assert_eq!(<f32 literal>, solve_langevin("<id>"))` / solve_solomon_potvin_bengio; the literal is kept as text (compared as f32)
    expected["tsptw"] = {}
    lines = (REF / "ddo/examples/tsptw/tests.rs").read_text().splitlines()
    for ln, line in enumerate(lines, 1):
        mm = re.search(r'assert_eq!\(([\d.]+), solve_(langevin|solomon_potvin_bengio)\("([^"]+)"\)\)', line)
        if not mm:
            continue
        name = ("Langevin/" if mm.group(2) == "langevin" else "SolomonPotvinBengio/") + mm.group(3)
        if name in TSPTW:
            ignored = "#[ignore]" in lines[ln - 3]
            assert not ignored, name
            (OUT / "tsptw" / name).parent.mkdir(parents=True, exist_ok=True)
            shutil.copy(REF / "resources/tsptw" / name, OUT / "tsptw" / name)
            expected["tsptw"][name] = {"optimum": mm.group(1), "source": f"ddo/examples/tsptw/tests.rs:{ln}"}
    assert len(expected["tsptw"]) == len(TSPTW)
    (OUT / "expected.json").write_text(json.dumps(expected, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
