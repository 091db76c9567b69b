"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes front end for the two CPU checkers:

* ``port``  -- oracle/libpsa_oracle.so, the C restatement in oracle/psa_oracle.c
* ``ref``   -- oracle/_ref/libpsa_ref.so, the UNMODIFIED reference compiled from
               /root/reference by ``make -C oracle ref`` (capacity 10000/5000)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (parallel-sequence-alignment_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libpsa_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpsa_ref.so")
REF_O0_SO = os.path.join(HERE, "_ref", "libpsa_ref_O0.so")


def build(ref: bool = True) -> None:
    """Compile the C restatement (and the reference .so when /root/reference exists)."""
    targets = ["port"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


@dataclass
class Result:
    offset: int
    char_offset: int
    ch: str
    score: float
    counts: tuple = (0, 0, 0, 0)

    def mutant(self, seq2: str) -> str:
        if self.char_offset < 0:
            return seq2
        return seq2[: self.char_offset] + self.ch + seq2[self.char_offset + 1:]


class _CResult(C.Structure):
    _fields_ = [("offset", C.c_int), ("char_offset", C.c_int), ("ch", C.c_char),
                ("score", C.c_double), ("counts", C.c_longlong * 4)]

    def py(self) -> Result:
        return Result(self.offset, self.char_offset, self.ch.decode("latin1") if self.ch != b"\x00" else "",
                      self.score, tuple(self.counts))


def _w(weights):
    return (C.c_double * 4)(*[float(x) for x in weights])


def _b(s) -> bytes:
    return s if isinstance(s, (bytes, bytearray)) else s.encode("latin1")


class Port:
    """The C restatement (length-explicit, any size)."""

    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.psa_oracle_sign.restype = C.c_char
        L.psa_oracle_sign.argtypes = [C.c_char, C.c_char]
        L.psa_oracle_weight.restype = C.c_double
        L.psa_oracle_weight.argtypes = [C.c_char, C.POINTER(C.c_double)]
        L.psa_oracle_substitute.restype = C.c_char
        L.psa_oracle_substitute.argtypes = [C.c_char, C.c_char, C.POINTER(C.c_double), C.c_int]
        L.psa_oracle_is_swapable.argtypes = [C.c_int] * 4 + [C.c_double] * 2 + [C.c_int]
        L.psa_oracle_offset_naive.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_char_p, C.c_char_p, C.c_long,
                                              C.c_long, C.POINTER(_CResult)]
        L.psa_oracle_search.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_char_p, C.c_long, C.c_char_p, C.c_long,
                                        C.c_long, C.c_long, C.c_int, C.POINTER(_CResult)]
        L.psa_oracle_search_batch.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_char_p, C.c_long, C.c_char_p,
                                              C.POINTER(C.c_longlong), C.c_int, C.c_int, C.POINTER(_CResult)]
        L.psa_oracle_scores.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_char_p, C.c_long, C.c_char_p, C.c_long,
                                        C.c_long, C.c_long, C.POINTER(C.c_double)]

    def sign(self, a: str, b: str) -> str:
        r = self.lib.psa_oracle_sign(_b(a), _b(b))
        return r.decode("latin1") if r != b"\x00" else ""

    def substitute(self, c1: str, c2: str, weights, is_max: bool) -> str:
        r = self.lib.psa_oracle_substitute(_b(c1), _b(c2), _w(weights), int(is_max))
        return r.decode("latin1") if r != b"\x00" else ""

    def weight(self, sign: str, weights) -> float:
        return self.lib.psa_oracle_weight(_b(sign) if sign else b"\x00", _w(weights))

    def is_swapable(self, off1, coff1, off2, coff2, s1, s2, is_max) -> bool:
        return bool(self.lib.psa_oracle_is_swapable(off1, coff1, off2, coff2, s1, s2, int(is_max)))

    def offset_naive(self, weights, is_max, seq1, seq2, offset) -> Result:
        out = _CResult()
        s1, s2 = _b(seq1), _b(seq2)
        rc = self.lib.psa_oracle_offset_naive(_w(weights), int(is_max), s1, s2, len(s2), offset, C.byref(out))
        if rc:
            raise ValueError(f"psa_oracle_offset_naive rc={rc}")
        return out.py()

    def search(self, weights, is_max, seq1, seq2, first=0, last=None, nthreads=1) -> Result:
        s1, s2 = _b(seq1), _b(seq2)
        if last is None:
            last = len(s1) - len(s2) + 1
        out = _CResult()
        rc = self.lib.psa_oracle_search(_w(weights), int(is_max), s1, len(s1), s2, len(s2), first, last,
                                        nthreads, C.byref(out))
        if rc:
            raise ValueError(f"psa_oracle_search rc={rc}")
        return out.py()

    def search_batch(self, weights, is_max, seq1, queries, nthreads=None) -> list:
        s1 = _b(seq1)
        qs = [_b(q) for q in queries]
        offs = [0]
        for q in qs:
            offs.append(offs[-1] + len(q))
        cat = b"".join(qs)
        out = (_CResult * len(qs))()
        rc = self.lib.psa_oracle_search_batch(_w(weights), int(is_max), s1, len(s1), cat,
                                              (C.c_longlong * len(offs))(*offs), len(qs),
                                              nthreads or os.cpu_count() or 1, out)
        if rc:
            raise ValueError(f"psa_oracle_search_batch rc={rc}")
        return [o.py() for o in out]

    def scores(self, weights, is_max, seq1, seq2, first=0, last=None):
        s1, s2 = _b(seq1), _b(seq2)
        if last is None:
            last = len(s1) - len(s2) + 1
        buf = (C.c_double * max(last - first, 1))()
        rc = self.lib.psa_oracle_scores(_w(weights), int(is_max), s1, len(s1), s2, len(s2), first, last, buf)
        if rc:
            raise ValueError(f"psa_oracle_scores rc={rc}")
        return list(buf)[: last - first]


class Ref:
    """The unmodified reference (capacity-bound; see oracle/ref_harness.cpp)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        L = self.lib = C.CDLL(path)
        dp, ip, cp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_char)
        L.ref_search_seq.restype = C.c_double
        L.ref_search_seq.argtypes = [dp, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, ip, ip, cp]
        L.ref_search_omp.restype = C.c_double
        L.ref_search_omp.argtypes = [dp, C.c_int, C.c_char_p, C.c_char_p, C.c_int, ip, ip, cp]
        L.ref_divide_execute_tasks.restype = C.c_double
        L.ref_divide_execute_tasks.argtypes = [dp, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                               C.c_int, ip, ip, cp]
        if hasattr(L, "ref_np2"):
            L.ref_np2.restype = C.c_double
            L.ref_np2.argtypes = [dp, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, ip, ip, cp, dp]
        L.ref_offset_score.restype = C.c_double
        L.ref_offset_score.argtypes = [dp, C.c_int, C.c_char_p, C.c_char_p, C.c_int, ip, cp]
        L.ref_sign.restype = C.c_char
        L.ref_sign.argtypes = [C.c_char, C.c_char]
        L.ref_pair_sign.restype = C.c_char
        L.ref_pair_sign.argtypes = [C.c_char, C.c_char]
        L.ref_weight.restype = C.c_double
        L.ref_weight.argtypes = [C.c_char, dp]
        L.ref_substitute.restype = C.c_char
        L.ref_substitute.argtypes = [C.c_char, C.c_char, dp, C.c_int]
        L.ref_is_swapable.argtypes = [C.c_int] * 4 + [C.c_double] * 2 + [C.c_int]
        L.ref_read_input.argtypes = [C.c_char_p, dp, ip, C.c_char_p, C.c_char_p]
        L.ref_write_output.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double]
        L.ref_fill_hash()
        self.cap1 = L.ref_seq1_capacity()
        self.cap2 = L.ref_seq2_capacity()

    @staticmethod
    def _res(score, o, c, ch) -> Result:
        return Result(o.value, c.value, ch.value.decode("latin1") if ch.value != b"\x00" else "", score)

    def search(self, weights, is_max, seq1, seq2, first=0, last=None) -> Result:
        s1, s2 = _b(seq1), _b(seq2)
        if last is None:
            last = len(s1) - len(s2) + 1
        o, c, ch = C.c_int(), C.c_int(), C.c_char()
        sc = self.lib.ref_search_seq(_w(weights), int(is_max), s1, s2, first, last, C.byref(o), C.byref(c), C.byref(ch))
        return self._res(sc, o, c, ch)

    def search_omp(self, weights, is_max, seq1, seq2, nthreads) -> Result:
        o, c, ch = C.c_int(), C.c_int(), C.c_char()
        sc = self.lib.ref_search_omp(_w(weights), int(is_max), _b(seq1), _b(seq2), nthreads,
                                     C.byref(o), C.byref(c), C.byref(ch))
        return self._res(sc, o, c, ch)

    def divide_execute_tasks(self, weights, is_max, seq1, seq2, num_processes=1, pid=0, pct=0, nthreads=4) -> Result:
        o, c, ch = C.c_int(), C.c_int(), C.c_char()
        sc = self.lib.ref_divide_execute_tasks(_w(weights), int(is_max), _b(seq1), _b(seq2), num_processes, pid,
                                               pct, nthreads, C.byref(o), C.byref(c), C.byref(ch))
        return self._res(sc, o, c, ch)

    def np2(self, weights, is_max, seq1, seq2, pct=100, ndev=1, nthreads=4):
        """Emulated `mpiexec -np 2` run of the reference (two host threads as the two ranks, divide_execute_tasks each,
        MAXLOC/MINLOC merge); returns (Result, [score of rank 0, score of rank 1])."""
        o, c, ch = C.c_int(), C.c_int(), C.c_char()
        rs = (C.c_double * 2)()
        sc = self.lib.ref_np2(_w(weights), int(is_max), _b(seq1), _b(seq2), pct, ndev, nthreads,
                              C.byref(o), C.byref(c), C.byref(ch), rs)
        return self._res(sc, o, c, ch), list(rs)

    def offset_score(self, weights, is_max, seq1, seq2, offset) -> Result:
        c, ch = C.c_int(), C.c_char()
        sc = self.lib.ref_offset_score(_w(weights), int(is_max), _b(seq1), _b(seq2), offset, C.byref(c), C.byref(ch))
        return Result(offset, c.value, ch.value.decode("latin1") if ch.value != b"\x00" else "", sc)

    def sign(self, a, b) -> str:
        r = self.lib.ref_sign(_b(a), _b(b))
        return r.decode("latin1") if r != b"\x00" else ""

    def substitute(self, c1, c2, weights, is_max) -> str:
        r = self.lib.ref_substitute(_b(c1), _b(c2), _w(weights), int(is_max))
        return r.decode("latin1") if r != b"\x00" else ""

    def weight(self, sign, weights) -> float:
        return self.lib.ref_weight(_b(sign) if sign else b"\x00", _w(weights))

    def is_swapable(self, off1, coff1, off2, coff2, s1, s2, is_max) -> bool:
        return bool(self.lib.ref_is_swapable(off1, coff1, off2, coff2, s1, s2, int(is_max)))


def ref_available() -> bool:
    return os.path.exists(REF_SO)
