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

    def has_edge(self, a: int, b: int) -> bool:
        return bool(np.any(((self.src == a) & (self.dst == b)) | ((self.src == b) & (self.dst == a))))

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


@dataclass
class TsptwInstance:
    """ddo/examples/tsptw/instance.rs:51-59: nb_nodes (depot included), distance matrix and one time window per node, as integers."""
    n: int
    dist: np.ndarray  # [n][n] int64
    tw: np.ndarray    # [n][2] int64 (earliest, latest)
    name: str = ""


def parse_tsptw(text: str, name: str = "") -> TsptwInstance:
    """ddo/examples/tsptw/instance.rs:61-108: `#` comment lines and empty lines are skipped; the first line holds the number of nodes, the next
    n lines the distance matrix, the last n lines `earliest latest`.  Every number is parsed as f32, multiplied by 10000.0 in f32 and truncated
    to an integer (instance.rs:86-87,97-98) -- 70.3 becomes 702999 or 703000 depending on that rounding, so the arithmetic is done in float32."""
    rows = [ln.strip() for ln in text.splitlines()]
    rows = [ln for ln in rows if ln and not ln.startswith("#")]
    n = int(rows[0].split()[0])
    scale = np.float32(10000.0)

    def conv(tok: str) -> int:
        return int(np.float32(tok) * scale)

    dist = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        for j, tok in enumerate(rows[1 + i].split()):
            dist[i, j] = conv(tok)
    tw = np.array([[conv(t) for t in rows[1 + n + i].split()[:2]] for i in range(n)], dtype=np.int64)
    return TsptwInstance(n, dist, tw, name)


def read_tsptw(path: str) -> TsptwInstance:
    with open(path) as f:
        return parse_tsptw(f.read(), str(path))


@dataclass
class Max2SatInstance:
    """A weighted MAX2SAT instance: ``clauses`` is int64[m, 3] = (weight, literal x, literal y); the literal of variable i (0-based) is
    +-(i + 1); x == y encodes a unit clause.  Duplicated clauses keep the LAST weight (the reference inserts into a hash map,
    ddo/examples/max2sat/data.rs:99,106)."""

    n: int
    clauses: np.ndarray
    name: str = ""
    words: int = field(init=False)

    def __post_init__(self):
        self.words = (self.n + 1) // 2  # int32 benefits packed two per uint64 word at the ABI
        self.clauses = np.ascontiguousarray(self.clauses, dtype=np.int64).reshape(-1, 3)

    def initial_state(self) -> np.ndarray:
        return np.zeros(self.words, dtype=np.uint64)  # model.rs:259-264: all benefits zero

    def to_wcnf(self) -> str:
        lines = [f"p wcnf {self.n} {len(self.clauses)}"]
        for w, x, y in self.clauses.tolist():
            lines.append(f"{w} {x} 0" if x == y else f"{w} {x} {y} 0")
        return "\n".join(lines) + "\n"


_WC_COMMENT = re.compile(r"^c\s.*$")
_WC_PB = re.compile(r"^p\s+wcnf\s+(?P<vars>\d+)\s+(?P<clauses>\d+)")
_WC_BIN = re.compile(r"^(?P<w>-?\d+)\s+(?P<x>-?\d+)\s+(?P<y>-?\d+)\s+0")
_WC_UNIT = re.compile(r"^(?P<w>-?\d+)\s+(?P<x>-?\d+)-?\s+0")


def parse_wcnf(text: str, name: str = "") -> Max2SatInstance:
    """Same grammar as the reference's reader (ddo/examples/max2sat/data.rs:66-110): the four regular expressions are tried in the same
    order; lines matching none of them are ignored."""
    n = 0
    clauses = []
    for raw in text.splitlines():
        line = raw.strip()
        if not line or _WC_COMMENT.match(line):
            continue
        m = _WC_PB.match(line)
        if m:
            n = int(m.group("vars"))
            continue
        m = _WC_BIN.match(line)
        if m:
            clauses.append((int(m.group("w")), int(m.group("x")), int(m.group("y"))))
            continue
        m = _WC_UNIT.match(line)
        if m:
            clauses.append((int(m.group("w")), int(m.group("x")), int(m.group("x"))))
            continue
    return Max2SatInstance(n, np.array(clauses, dtype=np.int64).reshape(-1, 3), name)


def read_wcnf(path: str) -> Max2SatInstance:
    with open(path) as f:
        return parse_wcnf(f.read(), name=str(path))


def random_max2sat(n: int, m: int, seed: int, max_weight: int = 10) -> Max2SatInstance:
    """The synthetic family of BASELINE.json config 3: m clauses over two distinct uniform variables, each literal
    negated with probability 1/2, integer weight uniform in [1, max_weight]; SplitMix64(seed)."""
    rng = SplitMix64(seed)
    clauses = []
    for _ in range(m):
        a = rng.next() % n
        b = rng.next() % (n - 1)
        if b >= a:
            b += 1
        x = (a + 1) * (-1 if rng.next() & 1 else 1)
        y = (b + 1) * (-1 if rng.next() & 1 else 1)
        w = 1 + rng.next() % max_weight
        clauses.append((w, x, y))
    return Max2SatInstance(n, np.array(clauses, dtype=np.int64), f"max2sat_{n}_{m}_{seed}")
