"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a, loads, exports every symbol include/ddo_b200.h declares, and
fails loudly (no CPU fallback) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "ddo_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ddo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    import __graft_entry__ as g
    from ddo_b200 import _native as N

    g.build()
    lib = C.CDLL(str(N.LIB_PATH))
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ddo_b200.h but not exported"
    assert sorted(N.SYMBOLS) == declared, "ddo_b200/_native.py must bind exactly the declared ABI"


def test_library_is_sm100a_only():
    import subprocess
    from ddo_b200 import _native as N

    out = subprocess.run(["cuobjdump", "-lelf", str(N.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device():
    import torch
    from ddo_b200 import Misp, gnp
    from ddo_b200 import _native as N

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert N.lib().ddo_device_count() == 0
    with pytest.raises(N.DdoError) as e:
        Misp(gnp(10, 0.5, 1))
    assert e.value.code == N.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_product_code_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load oracle/."""
    for f in list((ROOT / "ddo_b200").rglob("*.py")) + list((ROOT / "ddo_b200" / "csrc").glob("*")):
        if f.is_file() and f.suffix in {".py", ".cu", ".cuh", ".hpp", ".h"}:
            txt = f.read_text()
            assert "oracle_lib" not in txt and "liboracle" not in txt and "ddo_oracle" not in txt, f


def test_instance_generator_and_dimacs_roundtrip():
    from ddo_b200.instances import SplitMix64, gnp, parse_dimacs

    r = SplitMix64(1)
    assert [r.next() for _ in range(2)] == [10451216379200822465, 13757245211066428519]  # SplitMix64 reference vectors (seed 1)
    g = gnp(50, 0.5, 3)
    h = parse_dimacs(g.to_dimacs())
    assert h.n == g.n and np.array_equal(h.src, g.src) and np.array_equal(h.dst, g.dst) and np.array_equal(h.weights, g.weights)
    assert 450 < len(g.src) < 800
    with pytest.raises(ValueError):
        parse_dimacs("p edge 3 1\nthis is not an instance\n")
    w = parse_dimacs("c comment\np edge 3 2\nn 2 7\ne 1 2\ne 2 3\n")
    assert w.weights.tolist() == [1, 7, 1] and w.src.tolist() == [0, 1] and w.dst.tolist() == [1, 2]
    assert g.initial_state().tolist() == [(1 << 50) - 1]


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/ddo_b200.h must compile as C99 (no C++ types, no torch types) and as C++."""
    import subprocess

    src = tmp_path / "hdr.c"
    src.write_text('#include "ddo_b200.h"\nint main(void) { ddo_completion c; ddo_decision d; (void)c; (void)d; return 0; }\n')
    inc = str(ROOT / "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)], check=True, capture_output=True)
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)], check=True, capture_output=True)
