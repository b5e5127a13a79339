"""Host logic of the solver (no GPU): ddo::NoDupFringe's threaded burst push against its sequential push, checked by a small C++ program
(tests/host/fringe_check.cpp) linked against the in-tree libddo_b200.so.  Only host code of the library runs here."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_push_many_equals_sequential_pushes(tmp_path):
    import __graft_entry__ as g
    from ddo_b200 import _native as N

    g.build()
    exe = tmp_path / "fringe_check"
    libdir = Path(N.LIB_PATH).parent
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-I/usr/local/cuda/include", str(ROOT / "tests" / "host" / "fringe_check.cpp"), "-o", str(exe),
           f"-L{libdir}", f"-l:{Path(N.LIB_PATH).name}", "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK"), r.stdout
