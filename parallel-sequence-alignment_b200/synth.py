"""Seeded synthetic workloads of the shapes named in BASELINE.json (host logic, numpy only).

Distribution = the reference's own generator (sequences_generator, main.c:58-86): independent
uniform letters 'A' + r % 26.  The reference seeds with time(); here the stream is splitmix64 so
every box generates byte-identical inputs.  Seq1 is stream 0 of the workload seed, query q is
stream q+1; each stream starts from an independently mixed state (see _stream_state).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

GAMMA = 0x9E3779B97F4A7C15
_M64 = (1 << 64) - 1


def _mix(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _stream_state(seed, stream) -> np.ndarray:
    """Independent splitmix64 start state per (seed, stream): the finaliser is applied to the pair first,
    so streams are not shifted copies of one another (seed + k*GAMMA alone would make query q+1 equal
    query q moved by one letter)."""
    with np.errstate(over="ignore"):
        s = np.asarray(stream, dtype=np.uint64)
        return _mix(_mix(np.uint64(seed & _M64) * np.uint64(GAMMA) + np.uint64(0xD1B54A32D192ED03)) ^ _mix(s + np.uint64(1)))


def letters(seed: int, n: int, alphabet: int = 26, stream: int = 0) -> bytes:
    """n uniform letters: 'A' + (mix(state + (k+1)*GAMMA) >> 33) % alphabet, state = _stream_state(seed, stream)."""
    with np.errstate(over="ignore"):
        k = np.arange(1, n + 1, dtype=np.uint64)
        z = _mix(_stream_state(seed, stream) + k * np.uint64(GAMMA))
    return (np.uint8(65) + ((z >> np.uint64(33)) % np.uint64(alphabet)).astype(np.uint8)).tobytes()


def query_seed(seed: int, q: int) -> int:
    """kept for callers that want a scalar seed per query: stream q+1 of `seed`"""
    return int(_stream_state(seed, q + 1))


def query_matrix(seed: int, nq: int, len2: int) -> np.ndarray:
    """nq x len2 uint8 matrix of query letters; row q is stream q+1 of `seed` (Seq1 is stream 0)."""
    with np.errstate(over="ignore"):
        states = _stream_state(seed, np.arange(1, nq + 1, dtype=np.uint64))
        k = np.arange(1, len2 + 1, dtype=np.uint64)
        z = _mix(states[:, None] + k[None, :] * np.uint64(GAMMA))
    return (np.uint8(65) + ((z >> np.uint64(33)) % np.uint64(26)).astype(np.uint8))


@dataclass
class Workload:
    name: str
    weights: List[float]
    is_max: bool
    seq1: bytes
    queries: List[bytes] = field(default_factory=list)
    note: str = ""

    @property
    def pair_evals(self) -> int:
        n1 = len(self.seq1)
        return sum((n1 - len(q) + 1) * len(q) for q in self.queries)


def workload(name: str, nq: int | None = None, weights=None, seed_shift: int = 0) -> Workload:
    """BASELINE.json configs 2-5 (config 1 is the reference's input.txt block 1, a fixture under
    tests/golden/).  nq overrides the query count (bounded samples for CPU baselines);
    seed_shift decorrelates ranks in the weak-scaling bench."""
    name = name.lower()
    if name == "c2":      # single pair 3000/2000 MIN (tie-breaking)
        s = 2 + seed_shift
        return Workload("c2", list(weights or [1, 1, 1, 1]), False, letters(s, 3000), [letters(s, 2000, stream=1)],
                        "len1=3000 len2=2000 MIN")
    if name == "c3":      # 1024 queries len2=500 vs len1=3000 MAX
        s = 3 + seed_shift
        n = 1024 if nq is None else nq
        m = query_matrix(s, n, 500)
        return Workload("c3", list(weights or [1, 3, 4, 2]), True, letters(s, 3000), [m[i].tobytes() for i in range(n)],
                        f"{n} queries len2=500 len1=3000 MAX")
    if name == "c4":      # len1=1e6, len2=2000, offsets split across GPUs
        s = 4 + seed_shift
        return Workload("c4", list(weights or [2, 1.5, 1.1, 1.3]), True, letters(s, 1_000_000),
                        [letters(s, 2000, stream=1)], "len1=1000000 len2=2000 MAX")
    if name == "c5":      # 65536 queries len2=64 vs len1=10000 MIN
        s = 5 + seed_shift
        n = 65536 if nq is None else nq
        m = query_matrix(s, n, 64)
        return Workload("c5", list(weights or [1, 3, 4, 2]), False, letters(s, 10000), [m[i].tobytes() for i in range(n)],
                        f"{n} queries len2=64 len1=10000 MIN")
    raise ValueError(f"unknown workload {name!r} (c2..c5; c1 is tests/golden/input_blocks.json block 0)")
