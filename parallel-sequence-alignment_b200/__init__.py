"""parallel-sequence-alignment_b200 -- B200-native mutant-offset search (host-side Python mirror).

A thin ctypes front end over the C ABI in include/psa_b200.h (libpsa_b200.so, built in-tree by
``make -C parallel-sequence-alignment_b200`` / ``__graft_entry__.build()``).  Names follow the
reference: ProgramData / Mutant (program_data.h, mutant.h), gpu_run_program (cuda_funcs.h:33),
read_seq_and_weights_from_file / write_results_to_file (cpu_funcs.c:353-378).

There is no CPU fallback and no torch dependency here: if the shared library is missing the import
fails, and every search raises PsaError(PSA_ERR_CUDA) on a host without a B200.

The directory name contains hyphens, so import it with
``importlib.import_module("parallel-sequence-alignment_b200")``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsa_b200.so")
CLI_PATH = os.path.join(_HERE, "psa_b200_cli")

PSA_OK, PSA_ERR_ARG, PSA_ERR_ALPHABET, PSA_ERR_WEIGHTS, PSA_ERR_CUDA = 0, -1, -2, -3, -4
PSA_ERR_NOMEM, PSA_ERR_STATE, PSA_ERR_IO = -5, -6, -7

SEQ1_MAX_LEN = 10000      # def.h:35 (without the NUL)
SEQ2_MAX_LEN = 5000       # def.h:36
MAXIMUM_STR, MINIMUM_STR = "maximum", "minimum"   # def.h:43-44

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no Python or CPU fallback for the search)")
_lib = C.CDLL(LIB_PATH)


class PsaError(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        self.status = status
        msg = _lib.psa_strerror(status).decode()
        super().__init__(f"psa status {status}: {msg}" + (f" ({detail})" if detail else ""))


class Mutant(C.Structure):
    """mutant.h:6-10"""
    _fields_ = [("offset", C.c_int), ("char_offset", C.c_int), ("ch", C.c_char)]


class ProgramData(C.Structure):
    """program_data.h:6-11"""
    _fields_ = [("is_max", C.c_int), ("weights", C.c_double * 4),
                ("seq1", C.c_char * (SEQ1_MAX_LEN + 1)), ("seq2", C.c_char * (SEQ2_MAX_LEN + 1))]


class _CResult(C.Structure):
    _fields_ = [("mutant", Mutant), ("rank", C.c_int32), ("score", C.c_double), ("counts", C.c_int64 * 4)]


class _CPairTable(C.Structure):
    _fields_ = [("sign", (C.c_char * 27) * 27), ("substitute", (C.c_char * 27) * 27),
                ("diff", (C.c_double * 27) * 27), ("rank", (C.c_uint8 * 27) * 27),
                ("nranks", C.c_int32), ("exact", C.c_int32), ("frac_bits", C.c_int32), ("key_slack", C.c_int64)]


class _CProblem(C.Structure):
    """psa_problem (include/psa_b200.h): one entry of a psa_search_many list."""
    _fields_ = [("weights", C.POINTER(C.c_double)), ("is_max", C.c_int32), ("nq", C.c_int32), ("seq1", C.c_void_p), ("len1", C.c_int64),
                ("seq2s", C.c_void_p), ("q_off", C.c_void_p), ("out", C.c_void_p), ("status", C.c_int32), ("reserved", C.c_int32)]


class _CShard(C.Structure):
    _fields_ = [("q_begin", C.c_int32), ("q_end", C.c_int32), ("first", C.c_int64), ("last", C.c_int64)]


@dataclass
class Shard:
    q_begin: int
    q_end: int
    first: int = -1
    last: int = -1


@dataclass
class Result:
    offset: int
    char_offset: int
    ch: str
    score: float
    counts: tuple
    rank: int = 0

    def mutant(self, seq2: str) -> str:
        """Seq2 with the single substitution applied (cpu_funcs.c:96-98)."""
        if self.char_offset < 0:
            return seq2
        return seq2[: self.char_offset] + self.ch + seq2[self.char_offset + 1:]


ALPHABET = [chr(ord("A") + i) for i in range(26)] + ["-"]


@dataclass
class PairTable:
    sign: List[List[str]]        # [seq2 symbol][seq1 symbol]
    substitute: List[List[str]]
    diff: List[List[float]]
    rank: List[List[int]]
    nranks: int
    exact: bool
    frac_bits: int
    key_slack: int


def _sig():
    dp = C.POINTER(C.c_double)
    _lib.psa_strerror.restype = C.c_char_p
    _lib.psa_strerror.argtypes = [C.c_int]
    _lib.psa_abi_version.restype = C.c_int
    _lib.psa_device_count.restype = C.c_int
    _lib.psa_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]
    _lib.psa_destroy.argtypes = [C.c_void_p]
    _lib.psa_last_error.restype = C.c_char_p
    _lib.psa_last_error.argtypes = [C.c_void_p]
    _lib.psa_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
    _lib.psa_get_stat.restype = C.c_longlong
    _lib.psa_get_stat.argtypes = [C.c_void_p, C.c_char_p]
    _lib.psa_build_pair_table.argtypes = [dp, C.c_int, C.c_longlong, C.POINTER(_CPairTable)]
    _lib.psa_plan_shards.argtypes = [C.c_int64, C.POINTER(C.c_int64), C.c_int32, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                     C.POINTER(_CShard)]
    _lib.psa_merge_results.argtypes = [C.c_int, C.POINTER(_CResult), C.c_int, C.POINTER(_CResult)]
    _lib.psa_plan_packing.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib.psa_plan_stream_pieces.argtypes = [C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    _lib.psa_plan_stripes.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int, C.c_int, C.POINTER(C.c_int)]
    batch = [C.c_void_p, dp, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]
    _lib.psa_search_batch.argtypes = batch + [C.POINTER(_CResult)]
    _lib.psa_batch_prepare.argtypes = batch
    _lib.psa_search_many.argtypes = [C.c_void_p, C.POINTER(_CProblem), C.c_int32, C.c_int]
    _lib.psa_batch_run.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    _lib.psa_batch_fetch.argtypes = [C.c_void_p, C.POINTER(_CResult)]
    _lib.psa_search_range.argtypes = [C.c_void_p, dp, C.c_int, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64,
                                      C.c_int64, C.c_int64, C.POINTER(_CResult)]
    _lib.psa_offset_scores.argtypes = [C.c_void_p, dp, C.c_int, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int64, C.c_int64,
                                       dp, C.POINTER(C.c_int32), C.c_char_p]
    _lib.psa_alloc_pinned.restype = C.c_void_p
    _lib.psa_alloc_pinned.argtypes = [C.c_size_t]
    _lib.psa_free_pinned.argtypes = [C.c_void_p]
    _lib.psa_gpu_run_program.restype = C.c_double
    _lib.psa_gpu_run_program.argtypes = [C.POINTER(ProgramData), C.POINTER(Mutant), C.c_int, C.c_int]
    _lib.psa_read_input_file.argtypes = [C.c_char_p, dp, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    _lib.psa_write_output_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double]
    _lib.psa_run_files.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(_CResult)]
    _lib.psa_run_files_all.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
    _lib.psa_read_query_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
    _lib.psa_free.argtypes = [C.c_void_p]
    _lib.psa_run_query_file.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32)]
    _lib.psa_search_batch_mutants.argtypes = batch + [C.POINTER(_CResult), C.c_void_p]
    _lib.psa_topk_offsets.argtypes = [C.c_void_p, dp, C.c_int, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int64, C.c_int64,
                                      C.c_int32, C.POINTER(C.c_int32), dp, C.POINTER(C.c_int32), C.c_char_p, C.POINTER(C.c_int32)]
    # the six host primitives of cuda_funcs.h:44-61 (C spellings)
    _lib.psa_get_hashtable_sign.restype = C.c_char
    _lib.psa_get_hashtable_sign.argtypes = [C.c_char, C.c_char]
    _lib.psa_get_pair_sign.restype = C.c_char
    _lib.psa_get_pair_sign.argtypes = [C.c_char, C.c_char]
    _lib.psa_get_weight.restype = C.c_double
    _lib.psa_get_weight.argtypes = [C.c_char, dp]
    _lib.psa_get_substitute.restype = C.c_char
    _lib.psa_get_substitute.argtypes = [C.c_char, C.c_char, dp, C.c_int]
    _lib.psa_is_swapable.argtypes = [C.POINTER(Mutant), C.POINTER(Mutant), C.c_double, C.c_double, C.c_int]
    _lib.psa_strlen.argtypes = [C.c_char_p]


_sig()
_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def c_weights(weights) -> C.Array:
    return (C.c_double * 4)(*[float(x) for x in weights])


def _w(weights) -> C.Array:
    return (C.c_double * 4)(*[float(x) for x in weights])


def _b(s) -> bytes:
    return bytes(s) if isinstance(s, (bytes, bytearray, memoryview)) else s.encode("latin1")


def _py(r: _CResult) -> Result:
    ch = r.mutant.ch
    return Result(r.mutant.offset, r.mutant.char_offset, ch.decode("latin1") if ch != b"\x00" else "",
                  r.score, tuple(r.counts), r.rank)


def abi_version() -> int:
    return _lib.psa_abi_version()


def device_count() -> int:
    """Number of visible sm_100 devices (0 on a CPU-only host)."""
    return _lib.psa_device_count()


def build_pair_table(weights, is_max: bool, max_len2: int = 1) -> PairTable:
    """Host-resolved 27x27 table (no GPU needed)."""
    t = _CPairTable()
    rc = _lib.psa_build_pair_table(_w(weights), int(bool(is_max)), int(max_len2), C.byref(t))
    if rc:
        raise PsaError(rc)
    dec = lambda c: c.decode("latin1") if c != b"\x00" else ""
    return PairTable(
        sign=[[dec(t.sign[a][b:b + 1]) for b in range(27)] for a in range(27)],
        substitute=[[dec(t.substitute[a][b:b + 1]) for b in range(27)] for a in range(27)],
        diff=[[t.diff[a][b] for b in range(27)] for a in range(27)],
        rank=[[t.rank[a][b] for b in range(27)] for a in range(27)],
        nranks=t.nranks, exact=bool(t.exact), frac_bits=t.frac_bits, key_slack=t.key_slack)


def plan_shards(len1: int, query_lens: Sequence[int], nshards: int, granule: int = 1024,
                first: int = -1, last: int = -1) -> List[Shard]:
    """The partition psa_search_batch applies over a context's GPUs (host only, no GPU needed)."""
    offs = [0]
    for n in query_lens:
        offs.append(offs[-1] + int(n))
    out = (_CShard * nshards)()
    rc = _lib.psa_plan_shards(len1, (C.c_int64 * len(offs))(*offs), len(query_lens), nshards, granule, first, last, out)
    if rc:
        raise PsaError(rc, "psa_plan_shards")
    return [Shard(o.q_begin, o.q_end, o.first, o.last) for o in out]


def plan_packing(len1: int, len2: int, nq: int, force: int = 0):
    """(queries per block, warps per block) of packed mode for nq equal-length queries, (0, 0) if it does not apply
    (host only, no GPU needed)."""
    q, w = C.c_int(0), C.c_int(0)
    rc = _lib.psa_plan_packing(len1, len2, nq, force, C.byref(q), C.byref(w))
    if rc:
        raise PsaError(rc, "psa_plan_packing")
    return q.value, w.value


def plan_stripes(len1: int, len2: int, nq: int, rank_planes: int = 0, sm_count: int = 148) -> dict:
    """Launch shape of stripe mode (host only, no GPU needed) for a window that carries `rank_planes` rank bit planes (0, 1 or 2);
    {"ok": 0, ...} when the mode does not apply."""
    shape = (C.c_int * 8)()
    rc = _lib.psa_plan_stripes(len1, len2, nq, int(rank_planes), sm_count, shape)
    if rc:
        raise PsaError(rc, "psa_plan_stripes")
    return dict(zip(("ok", "lanes", "queries_per_task", "passes", "team_warps", "teams", "blocks", "smem_bytes"), list(shape)))


def plan_stream_pieces(nbytes: int):
    """(pieces, piece_bytes) a one-shot stripe-mode call streams `nbytes` of queries in (psa_plan_stream_pieces; host only)."""
    n, b = C.c_int(0), C.c_int64(0)
    rc = _lib.psa_plan_stream_pieces(nbytes, C.byref(n), C.byref(b))
    if rc:
        raise PsaError(rc)
    return n.value, b.value


def merge_results(is_max: bool, parts: Sequence[Result]) -> Result:
    """Reference-order merge of per-shard answers of one query (shards in ascending offset order)."""
    arr = (_CResult * len(parts))()
    for a, p in zip(arr, parts):
        a.mutant.offset, a.mutant.char_offset = p.offset, p.char_offset
        a.mutant.ch = p.ch.encode("latin1") if p.ch else b"\x00"
        a.score, a.rank = p.score, getattr(p, "rank", 0)
        for k in range(4):
            a.counts[k] = p.counts[k] if k < len(p.counts) else 0
    out = _CResult()
    rc = _lib.psa_merge_results(int(bool(is_max)), arr, len(parts), C.byref(out))
    if rc:
        raise PsaError(rc, "psa_merge_results")
    return _py(out)


class PinnedBuffer:
    """Page-locked host memory (psa_alloc_pinned) exposed as a writable memoryview."""

    def __init__(self, nbytes: int):
        self.ptr = _lib.psa_alloc_pinned(nbytes)
        if not self.ptr:
            raise PsaError(PSA_ERR_CUDA, "cudaMallocHost failed")
        self.nbytes = nbytes
        self.view = memoryview((C.c_char * nbytes).from_address(self.ptr)).cast("B")

    def close(self):
        if self.ptr:
            self.view.release()
            _lib.psa_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """Queries packed the way the C ABI wants them: one concatenated buffer + nq+1 byte offsets."""

    def __init__(self, seq1, queries: Sequence, pinned: bool = False):
        qs = [_b(q) for q in queries]
        s1 = _b(seq1)
        self.nq = len(qs)
        self.len1 = len(s1)
        self.lens = [len(q) for q in qs]
        cat = b"".join(qs)
        offs = [0]
        for n in self.lens:
            offs.append(offs[-1] + n)
        self.q_off = (C.c_int64 * (self.nq + 1))(*offs)
        self._pins = []
        if pinned:
            p1 = PinnedBuffer(max(len(s1), 1)); p1.view[: len(s1)] = s1
            p2 = PinnedBuffer(max(len(cat), 1)); p2.view[: len(cat)] = cat
            p3 = PinnedBuffer(8 * (self.nq + 1)); C.memmove(p3.ptr, self.q_off, 8 * (self.nq + 1))
            self._pins = [p1, p2, p3]
            self.seq1_ptr, self.seq2s_ptr, self.q_off_ptr = p1.ptr, p2.ptr, p3.ptr
        else:
            self._s1 = C.create_string_buffer(s1, max(len(s1), 1))
            self._cat = C.create_string_buffer(cat, max(len(cat), 1))
            self.seq1_ptr = C.cast(self._s1, C.c_void_p).value
            self.seq2s_ptr = C.cast(self._cat, C.c_void_p).value
            self.q_off_ptr = C.cast(self.q_off, C.c_void_p).value
        self.h2d_bytes = len(s1) + len(cat) + 8 * (self.nq + 1) + 4 * (self.nq + 1)
        self.pair_evals = sum((self.len1 - n + 1) * n for n in self.lens)


class Context:
    """One process, 1..8 GPUs.  Replaces MPI ranks + per-call cudaMalloc of the reference."""

    def __init__(self, ngpus: int = 1, devices: Optional[Iterable[int]] = None):
        devs = list(devices) if devices is not None else None
        n = len(devs) if devs is not None else ngpus
        arr = (C.c_int * n)(*devs) if devs is not None else None
        h = C.c_void_p()
        rc = _lib.psa_create(C.byref(h), arr, n)
        if rc:
            raise PsaError(rc, "psa_create: this library runs on B200 (sm_100) GPUs only")
        self._h = h
        self.ngpus = n

    def close(self):
        if getattr(self, "_h", None):
            _lib.psa_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise PsaError(rc, _lib.psa_last_error(self._h).decode())

    def set_option(self, name: str, value: int):
        self._check(_lib.psa_set_option(self._h, name.encode(), int(value)))

    def stat(self, name: str) -> int:
        return _lib.psa_get_stat(self._h, name.encode())

    # -- searches ---------------------------------------------------------------------------
    def search_batch(self, weights, is_max: bool, seq1, queries=None, batch: Optional[Batch] = None) -> List[Result]:
        b = batch if batch is not None else Batch(seq1, queries)
        out = (_CResult * max(b.nq, 1))()
        self._check(_lib.psa_search_batch(self._h, _w(weights), int(bool(is_max)), b.seq1_ptr, b.len1,
                                          b.seq2s_ptr, b.q_off_ptr, b.nq, out))
        return [_py(out[i]) for i in range(b.nq)]

    def search_batch_raw(self, weights_c, is_max: bool, batch: Batch, out):
        """The bare C call (psa_search_batch) on preallocated buffers: `weights_c` a c_double[4], `out` a
        result array from new_result_array().  No Python objects are created per result."""
        self._check(_lib.psa_search_batch(self._h, weights_c, int(bool(is_max)), batch.seq1_ptr, batch.len1,
                                          batch.seq2s_ptr, batch.q_off_ptr, batch.nq, out))

    def make_problem_list(self, items):
        """psa_problem array for search_many_raw: items = [(weights_c, is_max, Batch, result array), ...] (kept alive by the caller)."""
        arr = (_CProblem * max(len(items), 1))()
        for k, (wc, is_max, b, out) in enumerate(items):
            arr[k].weights = C.cast(wc, C.POINTER(C.c_double))
            arr[k].is_max = int(bool(is_max)); arr[k].nq = b.nq
            arr[k].seq1 = b.seq1_ptr; arr[k].len1 = b.len1
            arr[k].seq2s = b.seq2s_ptr; arr[k].q_off = b.q_off_ptr
            arr[k].out = C.cast(out, C.c_void_p).value
        return arr

    def search_many_raw(self, problems, n: int, lanes: int = 0):
        """The bare C call (psa_search_many): n problems pipelined over the context's device slots, `lanes` lanes each."""
        self._check(_lib.psa_search_many(self._h, problems, n, lanes))

    def search_many(self, items, lanes: int = 0) -> List[List[Result]]:
        """items = [(weights, is_max, seq1, queries), ...] -> one result list per problem (psa_search_many)."""
        built = []
        for w, is_max, seq1, queries in items:
            b = Batch(seq1, queries)
            built.append((_w(w), is_max, b, (_CResult * max(b.nq, 1))()))
        arr = self.make_problem_list(built)
        self.search_many_raw(arr, len(built), lanes)
        return [[_py(out[i]) for i in range(b.nq)] for (_, _, b, out) in built]

    @staticmethod
    def new_result_array(nq: int, pinned: bool = False):
        """Result array for search_batch_raw; pinned=True puts it in page-locked memory, which the library then
        fills straight from the device-to-host copy (no staging, no host pass)."""
        n = max(nq, 1)
        if not pinned:
            return (_CResult * n)()
        buf = PinnedBuffer(C.sizeof(_CResult) * n)
        arr = (_CResult * n).from_address(buf.ptr)
        arr._pin = buf          # keep the allocation alive with the array
        return arr

    @staticmethod
    def result_from_array(out, i: int) -> Result:
        return _py(out[i])

    def search(self, weights, is_max: bool, seq1, seq2) -> Result:
        return self.search_batch(weights, is_max, seq1, [seq2])[0]

    def search_range(self, weights, is_max: bool, seq1, seq2, first: int, last: int) -> Result:
        s1, s2 = _b(seq1), _b(seq2)
        out = _CResult()
        self._check(_lib.psa_search_range(self._h, _w(weights), int(bool(is_max)), s1, len(s1), s2, len(s2),
                                          first, last, C.byref(out)))
        return _py(out)

    def offset_scores(self, weights, is_max: bool, seq1, seq2, first: int = 0, last: Optional[int] = None):
        """Score profile of one query: (scores, char_offsets, letters) for offsets [first,last)."""
        s1, s2 = _b(seq1), _b(seq2)
        if last is None:
            last = len(s1) - len(s2) + 1
        n = max(last - first, 1)
        scores = (C.c_double * n)()
        coffs = (C.c_int32 * n)()
        letters = C.create_string_buffer(n)
        self._check(_lib.psa_offset_scores(self._h, _w(weights), int(bool(is_max)), s1, len(s1), s2, len(s2), first, last,
                                           scores, coffs, letters))
        m = last - first
        return list(scores)[:m], list(coffs)[:m], letters.raw[:m].decode("latin1")

    # -- split phase (resident batch) ------------------------------------------------------------
    def prepare(self, weights, is_max: bool, batch: Batch):
        self._check(_lib.psa_batch_prepare(self._h, _w(weights), int(bool(is_max)), batch.seq1_ptr, batch.len1,
                                           batch.seq2s_ptr, batch.q_off_ptr, batch.nq))
        self._nq = batch.nq

    def run(self) -> float:
        """Kernels only; returns device milliseconds (CUDA events on the library's streams)."""
        ms = C.c_float()
        self._check(_lib.psa_batch_run(self._h, C.byref(ms)))
        return ms.value

    def fetch(self) -> List[Result]:
        out = (_CResult * max(self._nq, 1))()
        self._check(_lib.psa_batch_fetch(self._h, out))
        return [_py(out[i]) for i in range(self._nq)]

    def run_files(self, input_path: str, output_path: str) -> Result:
        out = _CResult()
        self._check(_lib.psa_run_files(self._h, input_path.encode(), output_path.encode(), C.byref(out)))
        return _py(out)


def read_query_file(path: str) -> List[bytes]:
    """FASTA or one-query-per-token list -> queries (host only, no GPU needed)."""
    p1, p2, n = C.c_void_p(), C.c_void_p(), C.c_int32()
    rc = _lib.psa_read_query_file(path.encode(), C.byref(p1), C.byref(p2), C.byref(n))
    if rc:
        raise PsaError(rc, path)
    try:
        offs = (C.c_int64 * (n.value + 1)).from_address(p2.value)
        cat = C.string_at(p1.value, offs[n.value])
        return [cat[offs[i]:offs[i + 1]] for i in range(n.value)]
    finally:
        _lib.psa_free(p1)
        _lib.psa_free(p2)


def _search_batch_mutants(self, weights, is_max: bool, seq1, queries):
    """(results, mutant strings): psa_search_batch + every query with its one substitution applied, written on the device."""
    b = Batch(seq1, queries)
    out = (_CResult * max(b.nq, 1))()
    buf = C.create_string_buffer(max(sum(b.lens), 1))
    self._check(_lib.psa_search_batch_mutants(self._h, _w(weights), int(bool(is_max)), b.seq1_ptr, b.len1, b.seq2s_ptr, b.q_off_ptr,
                                              b.nq, out, C.cast(buf, C.c_void_p)))
    raw, cuts = buf.raw, [0]
    for n in b.lens:
        cuts.append(cuts[-1] + n)
    return [_py(out[i]) for i in range(b.nq)], [raw[cuts[i]:cuts[i + 1]].decode("latin1") for i in range(b.nq)]


def _topk_offsets(self, weights, is_max: bool, seq1, seq2, k: int, first: int = 0, last: Optional[int] = None):
    """The k best offsets of one query in the reference order: list of (offset, score, char_offset, letter)."""
    s1, s2 = _b(seq1), _b(seq2)
    if last is None:
        last = len(s1) - len(s2) + 1
    offs, scores, coffs = (C.c_int32 * k)(), (C.c_double * k)(), (C.c_int32 * k)()
    letters = C.create_string_buffer(k)
    found = C.c_int32()
    self._check(_lib.psa_topk_offsets(self._h, _w(weights), int(bool(is_max)), s1, len(s1), s2, len(s2), first, last, k,
                                      offs, scores, coffs, letters, C.byref(found)))
    return [(offs[r], scores[r], coffs[r], letters.raw[r:r + 1].decode("latin1")) for r in range(found.value)]


def _run_query_file(self, input_path: str, queries_path: str, output_path: str) -> int:
    n = C.c_int32()
    self._check(_lib.psa_run_query_file(self._h, input_path.encode(), queries_path.encode(), output_path.encode(), C.byref(n)))
    return n.value


Context.search_batch_mutants = _search_batch_mutants
Context.topk_offsets = _topk_offsets
Context.run_query_file = _run_query_file


def _run_files_all(self, input_path: str, output_path: str) -> int:
    """Every problem block stacked in the input file -> one output stanza per block; returns the count."""
    n = C.c_int()
    self._check(_lib.psa_run_files_all(self._h, input_path.encode(), output_path.encode(), C.byref(n)))
    return n.value


Context.run_files_all = _run_files_all


def make_program_data(weights, is_max: bool, seq1, seq2) -> ProgramData:
    d = ProgramData()
    d.is_max = int(bool(is_max))
    for i in range(4):
        d.weights[i] = float(weights[i])
    d.seq1 = _b(seq1)
    d.seq2 = _b(seq2)
    return d


def gpu_run_program(data: ProgramData, first_offset: int, last_offset: int):
    """Drop-in for cuda_funcs.h:33.  Returns (score, Mutant).  Exits the process on CUDA failure,
    like the reference."""
    m = Mutant()
    score = _lib.psa_gpu_run_program(C.byref(data), C.byref(m), first_offset, last_offset)
    return score, m


def _ch(r: bytes) -> str:
    return r.decode("latin1") if r != b"\x00" else ""


# The host primitives cpu_funcs.c imports from cuda_funcs.cu (cuda_funcs.h:44-61), same names and argument meaning.
def get_hashtable_sign(c1: str, c2: str) -> str:
    return _ch(_lib.psa_get_hashtable_sign(_b(c1), _b(c2)))


def get_pair_sign(a: str, b: str) -> str:
    return _ch(_lib.psa_get_pair_sign(_b(a), _b(b)))


def get_weight(sign: str, weights) -> float:
    return _lib.psa_get_weight(_b(sign) if sign else b"\x00", _w(weights))


def get_substitute(c1: str, c2: str, weights, is_max: bool) -> str:
    return _ch(_lib.psa_get_substitute(_b(c1), _b(c2), _w(weights), int(bool(is_max))))


def is_swapable(m1: Mutant, m2: Mutant, score1: float, score2: float, is_max: bool) -> bool:
    return bool(_lib.psa_is_swapable(C.byref(m1), C.byref(m2), score1, score2, int(bool(is_max))))


def strlen_gpu(s) -> int:
    return _lib.psa_strlen(_b(s))


def read_seq_and_weights_from_file(path: str):
    """cpu_funcs.c:353-368 -> (weights, is_max, seq1, seq2)"""
    w = (C.c_double * 4)()
    mx = C.c_int()
    p1, p2 = C.c_void_p(), C.c_void_p()
    rc = _lib.psa_read_input_file(path.encode(), w, C.byref(mx), C.byref(p1), C.byref(p2))
    if rc:
        raise PsaError(rc, path)
    try:
        return list(w), bool(mx.value), C.string_at(p1).decode("latin1"), C.string_at(p2).decode("latin1")
    finally:
        _libc.free(p1)
        _libc.free(p2)


def write_results_to_file(path: str, mutant: str, offset: int, score: float):
    """cpu_funcs.c:373-378"""
    rc = _lib.psa_write_output_file(path.encode(), mutant.encode("latin1"), int(offset), float(score))
    if rc:
        raise PsaError(rc, path)
