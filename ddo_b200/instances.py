"""Instance readers and synthetic generators (host side, numpy only).

* DIMACS ``p edge`` reader: same grammar as the reference's
  ``read_instance`` (ddo/examples/misp/main.rs:258-317): ``c`` comments,
  ``p edge <n> <m>``, optional ``n <node> <weight>`` lines, ``e <src> <dst>``
  lines (1-based); any other non-empty line is a format error.
* Knapsack reader: ddo/examples/knapsack/main.rs:267-299 (``n cap`` header then
  ``profit weight`` per item).
* ``gnp``: the synthetic G(n, p) family of BASELINE.json configs 2 and 5
  (SURVEY.md section 8d): SplitMix64(seed), ``u01 = (next() >> 11) * 2**-53``,
  edge (i, j), i < j in row-major order, iff ``u01 < p``; unit weights.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

_MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _MASK

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)

    def u01(self) -> float:
        return (self.next() >> 11) * (1.0 / (1 << 53))


@dataclass
class MispInstance:
    """A MISP instance: n vertices, integer weights, undirected edge list (0-based)."""

    n: int
    weights: np.ndarray  # int64[n]
    src: np.ndarray  # int32[m]
    dst: np.ndarray  # int32[m]
    name: str = ""
    words: int = field(init=False)

    def __post_init__(self):
        self.words = (self.n + 63) // 64
        self.weights = np.ascontiguousarray(self.weights, dtype=np.int64)
        self.src = np.ascontiguousarray(self.src, dtype=np.int32)
        self.dst = np.ascontiguousarray(self.dst, dtype=np.int32)

    def initial_state(self) -> np.ndarray:
        """All vertices present (misp/main.rs:69-71), packed little-endian in uint64 words."""
        s = np.zeros(self.words, dtype=np.uint64)
        for w in range(self.words):
            bits = min(64, self.n - 64 * w)
            s[w] = np.uint64(_MASK if bits == 64 else (1 << bits) - 1)
        return s

    def to_dimacs(self) -> str:
        lines = [f"p edge {self.n} {len(self.src)}"]
        for i, w in enumerate(self.weights):
            if int(w) != 1:
                lines.append(f"n {i + 1} {int(w)}")
        lines += [f"e {int(a) + 1} {int(b) + 1}" for a, b in zip(self.src, self.dst)]
        return "\n".join(lines) + "\n"


_PB = re.compile(r"^p\s+edge\s+(\d+)\s+(\d+)$")
_ND = re.compile(r"^n\s+(\d+)\s+(-?\d+)")
_ED = re.compile(r"^e\s+(\d+)\s+(\d+)")
_CM = re.compile(r"^c\s.*$")


def parse_dimacs(text: str, name: str = "") -> MispInstance:
    n = 0
    weights = None
    src, dst = [], []
    for raw in text.splitlines():
        line = raw.strip()
        if not line or _CM.match(line):
            continue
        m = _PB.match(line)
        if m:
            n = int(m.group(1))
            weights = np.ones(n, dtype=np.int64)
            continue
        m = _ND.match(line)
        if m:
            weights[int(m.group(1)) - 1] = int(m.group(2))
            continue
        m = _ED.match(line)
        if m:
            src.append(int(m.group(1)) - 1)
            dst.append(int(m.group(2)) - 1)
            continue
        raise ValueError("ill formed instance")
    if weights is None:
        raise ValueError("ill formed instance")
    return MispInstance(n, weights, np.array(src, dtype=np.int32), np.array(dst, dtype=np.int32), name)


def read_dimacs(path: str) -> MispInstance:
    with open(path) as f:
        return parse_dimacs(f.read(), name=str(path))


def gnp(n: int, p: float, seed: int) -> MispInstance:
    rng = SplitMix64(seed)
    src, dst = [], []
    for i in range(n):
        for j in range(i + 1, n):
            if rng.u01() < p:
                src.append(i)
                dst.append(j)
    return MispInstance(n, np.ones(n, dtype=np.int64), np.array(src, dtype=np.int32), np.array(dst, dtype=np.int32), f"gnp_{n}_{p}_{seed}")


@dataclass
class KnapsackInstance:
    capacity: int
    profit: np.ndarray
    weight: np.ndarray
    name: str = ""


def parse_knapsack(text: str, name: str = "") -> KnapsackInstance:
    """knapsack/main.rs:267-299: first line ``n capacity``; then ``profit weight``; stops after n items."""
    lines = [ln.strip() for ln in text.splitlines() if ln.strip()]
    n, cap = (int(x) for x in lines[0].split()[:2])
    profit, weight = [], []
    for ln in lines[1 : 1 + n]:
        a, b = ln.split()[:2]
        profit.append(int(a))
        weight.append(int(b))
    return KnapsackInstance(cap, np.array(profit, dtype=np.int64), np.array(weight, dtype=np.int64), name)


def random_knapsack(n: int, seed: int) -> KnapsackInstance:
    """BASELINE config 1 stand-in (no 50-item file exists in the reference's resources)."""
    rng = SplitMix64(seed)
    profit = np.array([1 + rng.next() % 1000 for _ in range(n)], dtype=np.int64)
    weight = np.array([1 + rng.next() % 1000 for _ in range(n)], dtype=np.int64)
    return KnapsackInstance(int(weight.sum() // 2), profit, weight, f"kp_{n}_{seed}")
