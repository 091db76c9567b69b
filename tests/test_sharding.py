"""Host-side partitioning and merging (psa_plan_shards / psa_merge_results): the single-process
replacement of the reference's rank split (cpu_funcs.c:128-133) and MAXLOC/MINLOC reduce
(cpu_funcs.c:64-94).  Runs without a GPU: the per-shard compute is done by the oracle, the plan and
the merge by the product library -- including a real 2-process gloo run of the one-rank-per-GPU layout
that bench.py uses."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import same_answer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_query_shards_cover_and_balance(psa):
    lens = [500] * 1024
    for n in (1, 2, 4, 8):
        sh = psa.plan_shards(3000, lens, n)
        assert sh[0].q_begin == 0 and sh[-1].q_end == 1024
        assert all(a.q_end == b.q_begin for a, b in zip(sh, sh[1:]))
        sizes = [s.q_end - s.q_begin for s in sh]
        assert max(sizes) - min(sizes) <= 1 and all(s.first == -1 and s.last == -1 for s in sh)
    # ragged: balanced by pair evaluations, not by query count
    lens = [10] * 100 + [1500] * 4
    sh = psa.plan_shards(3000, lens, 2)
    work = lambda a, b: sum((3000 - n + 1) * n for n in lens[a:b])
    w0, w1 = work(sh[0].q_begin, sh[0].q_end), work(sh[1].q_begin, sh[1].q_end)
    assert abs(w0 - w1) <= max((3000 - n + 1) * n for n in lens)
    # more shards than queries: the surplus stays empty
    sh = psa.plan_shards(3000, [64, 64, 64], 8)
    assert sum(s.q_end - s.q_begin for s in sh) == 3 and all(s.q_end - s.q_begin <= 1 for s in sh)


def test_offset_range_shards(psa):
    for (len1, len2, n, first, last) in ((1_000_000, 2000, 8, -1, -1), (9711, 2131, 2, -1, -1), (5000, 100, 4, 37, 4321),
                                         (3000, 2000, 8, -1, -1), (200, 100, 4, -1, -1)):
        sh = [s for s in psa.plan_shards(len1, [len2], n, 1024, first, last) if s.q_end > s.q_begin]
        lo, hi = (0, len1 - len2 + 1) if last < 0 else (first, last)
        assert sh[0].first == lo and sh[-1].last == hi
        assert all(a.last == b.first for a, b in zip(sh, sh[1:]))
        assert all(s.first < s.last and (s.first == lo or s.first % 128 == 0) for s in sh)
    with pytest.raises(psa.PsaError):
        psa.plan_shards(100, [10, 10], 2, 1024, 0, 5)          # a range needs a single query
    with pytest.raises(psa.PsaError):
        psa.plan_shards(100, [200], 2)


def test_merge_follows_reference_order(psa, port):
    R = psa.Result
    for is_max in (True, False):
        better, worse = (2.0, 1.0) if is_max else (1.0, 2.0)
        a, b = R(10, 3, "A", worse, (1, 0, 0, 0)), R(900, 1, "B", better, (1, 0, 0, 0))
        assert psa.merge_results(is_max, [a, b]).offset == 900
        assert psa.merge_results(is_max, [R(10, 3, "A", better, (1, 0, 0, 0)), R(900, 1, "B", better, (1, 0, 0, 0))]).offset == 10
        none = R(-1, -1, "", float("-inf") if is_max else float("inf"), (0, 0, 0, 0))
        assert psa.merge_results(is_max, [none, a]).offset == 10
        assert psa.merge_results(is_max, [none, none]).offset == -1
        # agrees with the oracle's is_swapable on every pair
        for x, y in ((a, b), (b, a), (a, a)):
            take_second = port.is_swapable(x.offset, x.char_offset, y.offset, y.char_offset, x.score, y.score, is_max)
            got = psa.merge_results(is_max, [x, y])
            assert got.offset == (y.offset if take_second else x.offset)


@pytest.mark.parametrize("nshards", [1, 2, 3, 8])
def test_partition_invariance_with_oracle_compute(psa, port, synth, nshards):
    """Any split gives the unsplit answer: offset ranges of one query (config-4 style) and query blocks."""
    s1, s2 = synth.letters(21, 20000), synth.letters(22, 300)
    for w, is_max in (([1, 1, 1, 1], True), ([1, 3, 4, 2], False), ([2, 1.5, 1.1, 1.3], True)):
        whole = port.search(w, is_max, s1, s2)
        parts = [port.search(w, is_max, s1, s2, sh.first, sh.last)
                 for sh in psa.plan_shards(len(s1), [len(s2)], nshards, 1024) if sh.q_end > sh.q_begin]
        assert same_answer(psa.merge_results(is_max, parts), whole)
    qs = [synth.letters(30 + k, 40 + 17 * k) for k in range(13)]
    whole = port.search_batch([1, 3, 4, 2], False, s1[:4000], qs)
    got = []
    for sh in psa.plan_shards(4000, [len(q) for q in qs], nshards):
        got += port.search_batch([1, 3, 4, 2], False, s1[:4000], qs[sh.q_begin:sh.q_end])
    assert all(same_answer(g, e) for g, e in zip(got, whole)) and len(got) == len(whole)


WORKER = textwrap.dedent('''
    import importlib, os, sys
    sys.path.insert(0, os.environ["PSA_ROOT"])
    import torch, torch.distributed as dist
    import oracle
    psa = importlib.import_module("parallel-sequence-alignment_b200")
    synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    port = oracle.Port()
    s1, s2 = synth.letters(41, 30000), synth.letters(42, 257)
    w, is_max = [1, 3, 4, 2], False
    sh = psa.plan_shards(len(s1), [len(s2)], world, 1024)[rank]          # this rank's offset range (one rank per GPU)
    mine = port.search(w, is_max, s1, s2, sh.first, sh.last)             # oracle stands in for the GPU
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine.offset, mine.char_offset, mine.ch, mine.score, mine.counts))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)             # max-over-ranks timing reduce used by bench.py
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        merged = psa.merge_results(is_max, [psa.Result(*g) for g in gathered])
        whole = port.search(w, is_max, s1, s2)
        ok = (merged.offset, merged.char_offset, merged.ch, merged.score) == (whole.offset, whole.char_offset, whole.ch, whole.score)
        print("RESULT", ok, int(t.item()) == world, merged.offset)
    dist.destroy_process_group()
''')


def test_two_ranks_gloo(tmp_path):
    """world_size 2 over gloo on CPU: plan -> per-rank search -> gather -> reference-order merge."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    env = dict(os.environ, PSA_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no), OMP_NUM_THREADS="2")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port_no), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
    assert line and line[0].split()[1:3] == ["True", "True"], p.stdout[-2000:]


def test_packing_plan(psa):
    """Packed mode's launch shape (psa_plan_packing, host arithmetic): a block takes Q whole queries laid lane by lane,
    L = ceil(offsets / 32) lanes each, in ceil(Q L / 32) warps; chosen for the fewest idle lanes, only when that beats whole
    warps per query by 5 %."""
    # config 3: 2501 offsets = 79 lanes -> 3 warps plain (82 %), two queries in 5 warps packed (99 %)
    assert psa.plan_packing(3000, 500, 1024) == (2, 5)
    assert psa.plan_packing(3000, 500, 1) == (0, 0)                     # a single query has nobody to share with
    # 1024 offsets fill one warp exactly: nothing to gain
    assert psa.plan_packing(1523, 500, 100) == (0, 0)
    # 11 offsets = 1 lane: eight queries (the cap) share one warp
    assert psa.plan_packing(100, 90, 40) == (8, 1)
    # config 5 is 311 lanes per query: beyond one block, and batch mode's job anyway
    assert psa.plan_packing(10000, 64, 65536) == (0, 0)
    assert psa.plan_packing(3000, 1024, 50) == (0, 0)                   # len2 > 1023: more than one staged window
    # forcing: honoured when the block stays within 8 warps
    assert psa.plan_packing(3000, 500, 1024, force=3) == (3, 8)
    assert psa.plan_packing(3000, 500, 1024, force=4) == (0, 0)
    for len1, len2, nq in ((700, 300, 37), (2100, 1000, 11), (330, 64, 50), (4200, 200, 7), (1055, 32, 5)):
        q, w = psa.plan_packing(len1, len2, nq)
        lanes = (len1 - len2 + 1 + 31) // 32
        if q:
            assert 2 <= q <= min(8, nq) and w == (q * lanes + 31) // 32 <= 8
            assert q * lanes / (32 * w) >= 1.05 * lanes / (32 * ((lanes + 31) // 32))
    with pytest.raises(psa.PsaError):
        psa.plan_packing(10, 20, 5)


def test_stripe_plans(psa):
    """Stripe mode's launch shape (psa_plan_stripes, host arithmetic) for the shapes the design documents.
    Config 3 is ONE wave: a split plan -- 7 queries = 553 lanes = 18 passes per block, 16 whole passes + 2 x 2 parts on 20
    warps, 147 blocks.  Config 5 and anything with thousands of tasks: one-warp teams; there whole tasks are what a
    scheduler is dealt, so a GPU's share of config 5 on eight GPUs takes smaller tasks than the whole batch does."""
    p = psa.plan_stripes(3000, 500, 1024, 0)
    assert (p["ok"], p["lanes"], p["queries_per_task"], p["passes"], p["team_warps"], p["teams"], p["blocks"]) == (1, 79, 7, 18, 20, 1, 147)
    assert p["smem_bytes"] <= 224 * 1024
    p = psa.plan_stripes(10000, 64, 65536, 2)
    assert (p["ok"], p["lanes"], p["queries_per_task"], p["passes"], p["team_warps"], p["teams"], p["blocks"]) == (1, 311, 4, 39, 1, 20, 148)
    p = psa.plan_stripes(3000, 500, 16384, 0)
    assert (p["ok"], p["queries_per_task"], p["team_warps"], p["teams"]) == (1, 2, 1, 20)
    p = psa.plan_stripes(10000, 64, 8192, 2)
    assert p["ok"] == 1 and p["team_warps"] == 1 and p["queries_per_task"] < 4
    # a window beyond shared memory, or queries longer than the counters take: not stripe mode
    assert psa.plan_stripes(5000, 1000, 600, 2)["ok"] == 0
    assert psa.plan_stripes(200000, 1500, 64, 0)["ok"] == 0
    # any plan: teams x warps fit the block, lanes cover the offsets, a split plan has one team
    for len1, len2, nq in ((700, 300, 37), (2100, 1000, 11), (330, 64, 5000), (4200, 200, 700), (1055, 32, 50), (3000, 300, 600), (2600, 640, 900)):
        p = psa.plan_stripes(len1, len2, nq, 0)
        if p["ok"]:
            assert p["lanes"] * 32 >= len1 - len2 + 1 and p["team_warps"] * p["teams"] <= 20 and p["blocks"] <= 148
            assert p["passes"] == (p["queries_per_task"] * p["lanes"] + 31) // 32
            if p["team_warps"] > p["passes"]:
                assert p["teams"] == 1 and p["team_warps"] - p["passes"] <= 3


def test_stream_piece_plan(psa):
    """How a one-shot stripe-mode call cuts its query copy into pieces (psa_plan_stream_pieces, host arithmetic): nothing
    below 64 KB, pieces of at least 512 KB, at most eight, every piece a multiple of 128 bytes (the kernel waits for the piece
    that holds the END of the last 128-byte line it will touch), together exactly covering the bytes."""
    assert psa.plan_stream_pieces(0) == (0, 0) and psa.plan_stream_pieces(65535) == (0, 0)
    assert psa.plan_stream_pieces(512000) == (1, 512000)                  # config 3
    assert psa.plan_stream_pieces(65536 * 64) == (8, 524288)              # config 5
    assert psa.plan_stream_pieces(8192 * 64) == (1, 524288)               # a GPU's share of config 5 on eight
    for nbytes in (65536, 100001, 1 << 20, (1 << 20) + 1, 3_000_000, 4_480_000, 50_000_000, 2_000_000_001):
        n, b = psa.plan_stream_pieces(nbytes)
        assert 1 <= n <= 8 and b % 128 == 0 and (n - 1) * b < nbytes <= n * b
        assert n == 1 or b >= 512 * 1024 or n == 8
    with pytest.raises(psa.PsaError):
        psa.plan_stream_pieces(-1)
