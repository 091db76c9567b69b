"""Parity of the CUDA path against the oracle, through the C ABI (run with -m gpu on a B200).

Bit-exact bar: offset, char_offset, substitute letter and the score double must equal the
oracle's (== the reference's 1-thread CPU path, see tests/test_oracle.py)."""
import os
import random
import subprocess
import sys

import pytest

from conftest import PKG, ROOT, same_answer

pytestmark = pytest.mark.gpu
ALPHA = [chr(65 + i) for i in range(26)] + ["-"]
# (engine, rank planes, warps per block, batch mode, bit-sliced epilogue, derive top rank from class planes, stripe mode):
# the scalar engine; the default dispatch; the linear-plane scan kernels with every plane count (0 planes = every offset
# unresolved -> the in-kernel settle path carries the result), both tile shapes and both epilogues, with and without
# deriving the top-rank bit; and stripe mode forced wherever a batch qualifies (single queries included).
ENGINES = [(1, -1, 0, -1, 1, 1, 0), (0, -1, 0, -1, 1, 1, -1), (2, 0, 1, 0, 1, 1, 0), (2, 1, 2, 0, 0, 1, 0), (2, 4, 3, 0, 1, 1, 0),
           (2, 2, 4, 0, 0, 1, 0), (2, 0, 0, 1, 0, 1, 0), (2, 2, 0, 1, 1, 1, 0), (2, 4, 0, 1, 0, 1, 0), (2, 1, 0, 1, 1, 1, 0),
           (2, 1, 0, 0, 1, 0, 0), (2, 1, 0, 1, 1, 0, 0), (2, 1, 0, -1, 1, 1, 1), (2, 0, 0, -1, 1, 1, 1), (2, 1, 0, -1, 1, 0, 1)]
ENGINE_IDS = ["scalar", "auto", "scan-k0-w1-bs", "scan-k1-w2", "scan-k4-w3-bs", "scan-k2-w4", "batch-k0", "batch-k2-bs", "batch-k4",
              "batch-k1-bs", "scan-k1-planes", "batch-k1-planes", "stripe-k1", "stripe-k0", "stripe-k1-planes"]


def _set_engine(ctx, engine, planes=-1, warps=0, batch=-1, sliced=1, derive=1, stripe=None):
    """stripe: None = auto (-1) under the default dispatch (engine 0), off (0) when a test names the linear-plane kernels."""
    ctx.set_option("engine", engine)
    ctx.set_option("rank_planes", planes)
    ctx.set_option("scan_warps", warps)
    ctx.set_option("batch_mode", batch)
    ctx.set_option("sliced_keys", sliced)
    ctx.set_option("derive_rank", derive)
    ctx.set_option("stripe_mode", (-1 if engine == 0 else 0) if stripe is None else stripe)


@pytest.fixture(params=ENGINES, ids=ENGINE_IDS)
def engine(request, ctx):
    _set_engine(ctx, *request.param)
    yield request.param[0]
    _set_engine(ctx, 0)


def test_library_is_native_and_gpu_visible(psa):
    assert psa.device_count() >= 1


def test_input_blocks(ctx, engine, input_blocks):
    """The 10 problems of the reference's input.txt (block 0 is pinned by its output.txt)."""
    for b in input_blocks:
        r = ctx.search(b["weights"], b["goal"] == "maximum", b["seq1"], b["seq2"])
        assert same_answer(r, b["expect"]), (b["weights_text"], r)
        assert "%g" % r.score == b["expect"]["score_g"]
    b = input_blocks[0]
    r = ctx.search(b["weights"], False, b["seq1"], b["seq2"])
    assert r.mutant(b["seq2"]) + "\n%d %g" % (r.offset, r.score) == b["output_txt"]


def test_synthetic_golden(ctx, engine, synthetic_cases):
    for c in synthetic_cases:
        r = ctx.search(c["weights"], c["is_max"], c["seq1"], c["seq2"])
        assert same_answer(r, c["expect"]), (c["tag"], c["weights"], c["is_max"], r, c["expect"])


def test_counts_and_rank_reported(ctx, engine, port, input_blocks):
    b = input_blocks[0]
    r = ctx.search(b["weights"], False, b["seq1"], b["seq2"])
    o = port.search(b["weights"], False, b["seq1"], b["seq2"])
    assert r.counts == o.counts == (60, 312, 240, 1519) and sum(r.counts) == len(b["seq2"])


@pytest.mark.parametrize("w", [[1, 1, 1, 1], [1, 3, 4, 2], [2, 1.5, 1.1, 1.3], [1.5, 2.6, 0.1, 0.2], [0.8, 0.54, 2.6, 13.7]])
def test_config2_tie_breaking(ctx, engine, port, synth, w):
    """BASELINE config 2: 3000/2000, MIN (and MAX), tie-rich integer weights and FP weights."""
    wl = synth.workload("c2", weights=w)
    for is_max in (False, True):
        r = ctx.search(w, is_max, wl.seq1, wl.queries[0])
        assert same_answer(r, port.search(w, is_max, wl.seq1, wl.queries[0]))


def test_planted_exact_ties(ctx, engine, port, synth):
    """Seq2 planted at several offsets: equal best scores, the lowest offset must win."""
    core = synth.letters(77, 400)
    s1 = synth.letters(78, 1500) + core + synth.letters(79, 900) + core + synth.letters(80, 5000) + core
    for w in ([1, 3, 4, 2], [1.5, 2.6, 0.1, 0.2]):
        for is_max in (True, False):
            r = ctx.search(w, is_max, s1, core)
            assert same_answer(r, port.search(w, is_max, s1, core))
            if is_max:
                assert r.offset == 1500


def test_random_shapes(ctx, engine, port):
    rng = random.Random(99)
    wsets = [[1, 3, 4, 2], [1, 1, 1, 1], [2, 1.5, 1.1, 1.3], [0.1, 0.7, 0.3, 0.9], [7, 0, 2, 0.5], [0, 0, 0, 0]]
    for trial in range(120):
        w = rng.choice(wsets)
        is_max = trial % 2
        n1 = rng.choice([1, 2, 31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 4095, 4097, rng.randint(1, 9000)])
        n2 = rng.choice([1, n1, max(1, n1 - 1), rng.randint(1, n1), rng.randint(1, min(n1, 70))])
        alpha = ALPHA if trial % 5 == 0 else ALPHA[:26] if trial % 3 else "ACDG"
        s1 = "".join(rng.choice(alpha) for _ in range(n1))
        s2 = "".join(rng.choice(alpha) for _ in range(n2))
        r = ctx.search(w, is_max, s1, s2)
        assert same_answer(r, port.search(w, is_max, s1, s2)), (trial, w, is_max, n1, n2)


def test_ragged_batch(ctx, engine, port, synth):
    """Queries of very different lengths in one batch, incl. len2 == len1 and len2 == 1."""
    s1 = synth.letters(5, 5000)
    lens = [1, 2, 31, 32, 33, 64, 500, 777, 1024, 2047, 4999, 5000, 3, 64, 64, 1999]
    qs = [synth.letters(100 + k, n) for k, n in enumerate(lens)]
    for w, is_max in (([1, 3, 4, 2], False), ([1.5, 2.6, 0.1, 0.2], True)):
        got = ctx.search_batch(w, is_max, s1, qs)
        exp = port.search_batch(w, is_max, s1, qs)
        for k, (g, e) in enumerate(zip(got, exp)):
            assert same_answer(g, e), (k, lens[k], g, e)
            assert g.counts == e.counts
    assert ctx.search_batch([1, 1, 1, 1], True, s1, []) == []


def test_offset_ranges(ctx, engine, port, synth):
    """psa_search_range == the reference's [first,last) contract of gpu_run_program."""
    s1, s2 = synth.letters(8, 7000), synth.letters(9, 300)
    for w in ([1, 3, 4, 2], [2, 1.5, 1.1, 1.3]):
        for (f, l) in ((0, 6701), (0, 1), (6700, 6701), (17, 4113), (4096, 4097), (31, 33), (1000, 6000)):
            for is_max in (True, False):
                r = ctx.search_range(w, is_max, s1, s2, f, l)
                assert same_answer(r, port.search(w, is_max, s1, s2, f, l)), (w, f, l, is_max)


def test_gpu_run_program_dropin(psa, port, input_blocks):
    """The reference entry point (cuda_funcs.h:33) on the reference's own records."""
    for b in input_blocks[:4]:
        is_max = b["goal"] == "maximum"
        d = psa.make_program_data(b["weights"], is_max, b["seq1"], b["seq2"])
        total = len(b["seq1"]) - len(b["seq2"]) + 1
        score, m = psa.gpu_run_program(d, 0, total)
        e = b["expect"]
        assert (score, m.offset, m.char_offset, m.ch.decode()) == (e["score"], e["offset"], e["char_offset"], e["ch"])
        # two "ranks" splitting the offsets like cpu_funcs.c:128-133, merged with MAXLOC/MINLOC (ties -> rank 0)
        half = total // 2
        if half:
            s0, m0 = psa.gpu_run_program(d, 0, half)
            s1_, m1 = psa.gpu_run_program(d, half, total)
            win = (s1_, m1) if ((s1_ > s0) if is_max else (s1_ < s0)) else (s0, m0)
            assert (win[0], win[1].offset, win[1].char_offset) == (e["score"], e["offset"], e["char_offset"])


def test_rejects_bad_input(psa, ctx):
    with pytest.raises(psa.PsaError) as e:
        ctx.search([1, 1, 1, 1], True, "ABCDEFGH", "ABc")
    assert e.value.status == psa.PSA_ERR_ALPHABET
    with pytest.raises(psa.PsaError) as e:
        ctx.search([1, 1, 1, 1], True, "ABC*EFGH", "ABC")
    assert e.value.status == psa.PSA_ERR_ALPHABET
    with pytest.raises(psa.PsaError) as e:
        ctx.search([1, 1, 1, 1], True, "AB", "ABC")
    assert e.value.status == psa.PSA_ERR_ARG
    with pytest.raises(psa.PsaError) as e:
        ctx.search([1, float("nan"), 1, 1], True, "ABC", "AB")
    assert e.value.status == psa.PSA_ERR_WEIGHTS
    assert ctx.search([1, 1, 1, 1], True, "ABCDEFGH", "ABC").offset == 0      # context still usable


def test_run_files_and_cli(psa, ctx, tmp_path, input_blocks):
    import subprocess
    b = input_blocks[0]
    (tmp_path / "input.txt").write_text(" ".join(b["weights_text"]) + "\n" + b["seq1"] + "\n" + b["seq2"] + "\n" + b["goal"] + "\n")
    r = ctx.run_files(str(tmp_path / "input.txt"), str(tmp_path / "out_api.txt"))
    assert (tmp_path / "out_api.txt").read_text() == b["output_txt"] and r.offset == 4505
    p = subprocess.run([psa.CLI_PATH], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert (tmp_path / "output.txt").read_text() == b["output_txt"]
    assert "CUDA percentage set to 100" in p.stdout and "total time:" in p.stdout


def test_config3_sample_and_full_properties(ctx, port, synth):
    """BASELINE config 3 (1024 x 500 vs 3000, MAX): a 64-query sample bit-exact against the oracle,
    the full batch through size-independent properties."""
    _set_engine(ctx, 0)
    wl = synth.workload("c3")
    got = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:64])
    for g, e in zip(got, exp):
        assert same_answer(g, e)
    for q, g in zip(wl.queries, got):
        assert 0 <= g.offset <= 2500 and 0 <= g.char_offset < 500 and sum(g.counts) == 500
        n = port.offset_naive(wl.weights, wl.is_max, wl.seq1, q, g.offset)       # the reported score is the true score there
        assert (n.score, n.char_offset, n.ch, n.counts) == (g.score, g.char_offset, g.ch, g.counts)
    # batching must be invisible: same answers one query at a time and in reverse order
    again = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[::-1][:50])
    assert [(a.offset, a.score) for a in again] == [(g.offset, g.score) for g in got[::-1][:50]]


def test_config5_sample_and_properties(ctx, port, synth):
    """BASELINE config 5 (65536 x 64 vs 10000, MIN): 4096-query slice; 256 bit-exact vs the oracle."""
    _set_engine(ctx, 0)
    wl = synth.workload("c5", nq=4096)
    got = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:256])
    for g, e in zip(got, exp):
        assert same_answer(g, e)
    for q, g in list(zip(wl.queries, got))[::16]:
        n = port.offset_naive(wl.weights, wl.is_max, wl.seq1, q, g.offset)
        assert (n.score, n.char_offset, n.ch) == (g.score, g.char_offset, g.ch)


def test_config4_long_seq1(ctx, port, synth):
    """BASELINE config 4 (len1 = 1e6, len2 = 2000, non-dyadic weights): beyond the reference's static
    capacity, so the checker is the oracle port (multi-threaded)."""
    _set_engine(ctx, 0)
    wl = synth.workload("c4")
    r = ctx.search(wl.weights, wl.is_max, wl.seq1, wl.queries[0])
    assert same_answer(r, port.search(wl.weights, wl.is_max, wl.seq1, wl.queries[0], nthreads=8))
    r2 = ctx.search([1, 3, 4, 2], False, wl.seq1, wl.queries[0])
    assert same_answer(r2, port.search([1, 3, 4, 2], False, wl.seq1, wl.queries[0], nthreads=8))


def _multi_device_body(psa, port, synth, devices):
    """Offset ranges of one query and query blocks of a batch over the devices of ONE context must give the
    single-GPU / oracle answer: worker threads (one per extra device), per-device shards with their halo, and the
    host merge of the per-device candidates all run here."""
    s1 = synth.letters(61, 200_000)
    s2 = synth.letters(62, 1500)
    qs = [synth.letters(70 + k, 64 + 8 * k) for k in range(97)]
    eq = [synth.letters(170 + k, 200) for k in range(300)]                # equal lengths: packed / batch mode per device
    with psa.Context(devices=devices) as c:
        assert c.ngpus == len(devices)
        c.set_option("min_split_work", 0)                                 # these problems are small: split them all the same
        for w, is_max in (([1, 3, 4, 2], False), ([2, 1.5, 1.1, 1.3], True), ([1, 1, 1, 1], True)):
            r = c.search(w, is_max, s1, s2)
            assert same_answer(r, port.search(w, is_max, s1, s2, nthreads=8)), (w, is_max)
            got = c.search_batch(w, is_max, s1[:20000], qs)
            exp = port.search_batch(w, is_max, s1[:20000], qs)
            assert all(same_answer(g, e) for g, e in zip(got, exp)), (w, is_max)
            got = c.search_batch(w, is_max, s1[:3000], eq)
            exp = port.search_batch(w, is_max, s1[:3000], eq)
            assert all(same_answer(g, e) and g.counts == e.counts for g, e in zip(got, exp)), (w, is_max)
        # an explicit offset range is split too (the gpu_run_program contract on a multi-device context)
        r = c.search_range([1, 3, 4, 2], True, s1, s2, 777, 150_001)
        assert same_answer(r, port.search([1, 3, 4, 2], True, s1, s2, 777, 150_001, nthreads=8))
        # fewer queries than devices: the surplus devices stay idle
        got = c.search_batch([1, 3, 4, 2], False, s1[:5000], qs[:3])
        assert all(same_answer(g, e) for g, e in zip(got, port.search_batch([1, 3, 4, 2], False, s1[:5000], qs[:3])))
        # planted ties across the device boundary: the lowest offset must win wherever the split falls
        core = synth.letters(63, 700)
        tied = synth.letters(64, 50_000) + core + synth.letters(65, 80_000) + core + synth.letters(66, 30_000)
        r = c.search([1, 3, 4, 2], True, tied, core)
        assert r.offset == 50_000 and same_answer(r, port.search([1, 3, 4, 2], True, tied, core, nthreads=8))
        # split-phase form on several devices
        b = psa.Batch(s1[:3000], eq)
        c.prepare([1, 3, 4, 2], True, b)
        assert c.run() > 0
        exp = port.search_batch([1, 3, 4, 2], True, s1[:3000], eq)
        assert all(same_answer(g, e) for g, e in zip(c.fetch(), exp))
        # a bad symbol on a shard that is not the first device's
        with pytest.raises(psa.PsaError) as e:
            c.search_batch([1, 3, 4, 2], True, s1[:3000], eq[:-1] + [b"AB?D" * 50])
        assert e.value.status == psa.PSA_ERR_ALPHABET
        assert same_answer(c.search([1, 3, 4, 2], True, s1[:3000], eq[0]), exp[0])


@pytest.mark.parametrize("copies", [2, 3, 8])
def test_multi_device_context_on_one_gpu(psa, port, synth, copies):
    """The north_star split (one process, (query | offset range) shards over the context's devices, host merge) with the
    SAME ordinal repeated: every device slot has its own stream, buffers and worker thread, so the whole multi-device
    path -- psa_plan_shards, the per-shard copies, the concurrent enqueue and psa_merge_results -- runs on a 1-GPU box."""
    _multi_device_body(psa, port, synth, [0] * copies)


def test_multi_gpu_single_process(psa, port, synth):
    """The same on every visible GPU (the replacement of the reference's MPI ranks).  Needs >= 2 GPUs."""
    n = psa.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    _multi_device_body(psa, port, synth, list(range(n)))


def test_all_stacked_blocks(psa, ctx, tmp_path, input_blocks):
    """every block of an input.txt-style file (the reference parses only the first)."""
    _set_engine(ctx, 0)
    text = "\n\n".join(" ".join(b["weights_text"]) + "\n" + b["seq1"] + "\n" + b["seq2"] + "\n" + b["goal"] for b in input_blocks)
    (tmp_path / "input.txt").write_text(text + "\n\n")
    n = ctx.run_files_all(str(tmp_path / "input.txt"), str(tmp_path / "output.txt"))
    assert n == len(input_blocks)
    lines = (tmp_path / "output.txt").read_text().split("\n")
    assert len(lines) == 2 * n
    for k, b in enumerate(input_blocks):
        e = b["expect"]
        mut = b["seq2"][: e["char_offset"]] + e["ch"] + b["seq2"][e["char_offset"] + 1:]
        assert lines[2 * k] == mut and lines[2 * k + 1] == "%d %s" % (e["offset"], e["score_g"])
    import subprocess
    p = subprocess.run([psa.CLI_PATH, "--all-blocks"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and (tmp_path / "output.txt").read_text().split("\n") == lines


@pytest.mark.parametrize("slices,planes,fused", [(0, -1, 1), (2, -1, 1), (5, 0, 1), (16, 1, 1), (0, -1, 0), (4, 1, 0)])
def test_slice_mode_single_query(ctx, port, synth, input_blocks, slices, planes, fused):
    """One query cut along its alignment steps (k_scan slices + k_combine): same answers as the oracle for exact
    and re-scored weights, explicit ranges, unresolved ranks (0 planes) and ties -- with the finish step run by
    k_combine's last block (small grids, the default) or by its own kernel."""
    _set_engine(ctx, 2, planes)
    ctx.set_option("slices", slices)
    ctx.set_option("fused_finish", fused)
    ctx.set_option("single_launch", 0)                                    # this test is about the k_profile / k_scan / k_combine chain
    try:
        for b in (input_blocks[0], input_blocks[4], input_blocks[6], input_blocks[9]):
            r = ctx.search(b["weights"], b["goal"] == "maximum", b["seq1"], b["seq2"])
            assert same_answer(r, b["expect"]), (slices, planes, b["weights_text"], r)
        s1, s2 = synth.letters(91, 30000), synth.letters(92, 777)
        core = synth.letters(93, 900)
        tied = synth.letters(94, 3000) + core + synth.letters(95, 4000) + core + synth.letters(96, 100)
        for w in ([1, 3, 4, 2], [1.5, 2.6, 0.1, 0.2], [1, 1, 1, 1]):
            for is_max in (True, False):
                assert same_answer(ctx.search(w, is_max, s1, s2), port.search(w, is_max, s1, s2, nthreads=4))
                assert same_answer(ctx.search_range(w, is_max, s1, s2, 1000, 20011), port.search(w, is_max, s1, s2, 1000, 20011))
                assert same_answer(ctx.search(w, is_max, tied, core), port.search(w, is_max, tied, core))
        if slices >= 2:
            assert ctx.stat("slices") >= 2
        # the last search (9 000 offsets = 36 combine blocks): profile + scan + combine, + finish when not fused
        assert ctx.stat("kernel_launches") == (3 if fused else 4)
    finally:
        ctx.set_option("slices", 0)
        ctx.set_option("fused_finish", 1)
        ctx.set_option("single_launch", 1)
        _set_engine(ctx, 0)


@pytest.mark.parametrize("planes", [-1, 0])
def test_single_query_in_one_launch(ctx, port, synth, input_blocks, planes):
    """k_single: one query in exact order = ONE cooperative launch ((tile, slice) units with their own striped windows, grid
    barrier, combine, finish).  The reference's own problems, planted ties, explicit offset ranges (the gpu_run_program
    contract), unresolved ranks (0 planes), queries from 64 to 5000 symbols; weights that need the re-score keep the chain."""
    _set_engine(ctx, 2, planes)
    try:
        for b in input_blocks:
            r = ctx.search(b["weights"], b["goal"] == "maximum", b["seq1"], b["seq2"])
            assert same_answer(r, b["expect"]), (planes, b["weights_text"], r)
            if len(b["seq2"]) >= 64 and ctx.stat("exact"):
                assert (ctx.stat("single_launch"), ctx.stat("kernel_launches")) == (1, 1), b["weights_text"]
        s1 = synth.letters(191, 30000)
        core = synth.letters(193, 900)
        tied = synth.letters(194, 3000) + core + synth.letters(195, 4000) + core + synth.letters(196, 100)
        for n2 in (64, 65, 777, 2047, 5000):
            s2 = synth.letters(192, n2)
            for w in ([1, 3, 4, 2], [1, 1, 1, 1], [5, 1, 2, 3]):
                for is_max in (True, False):
                    assert same_answer(ctx.search(w, is_max, s1, s2), port.search(w, is_max, s1, s2, nthreads=4)), (n2, w, is_max)
                    assert ctx.stat("single_launch") == 1
                    assert same_answer(ctx.search_range(w, is_max, s1, s2, 1000, 20011), port.search(w, is_max, s1, s2, 1000, 20011)), (n2, w)
                    assert same_answer(ctx.search_range(w, is_max, s1, s2, 77, 78), port.search(w, is_max, s1, s2, 77, 78))
        for w in ([1, 3, 4, 2], [1, 1, 1, 1]):
            for is_max in (True, False):
                assert same_answer(ctx.search(w, is_max, tied, core), port.search(w, is_max, tied, core))
        ctx.search([1.5, 2.6, 0.1, 0.2], True, s1, synth.letters(192, 777))              # order needs the reference's double
        assert ctx.stat("single_launch") == 0
        with pytest.raises(Exception):
            ctx.search([1, 3, 4, 2], True, s1, "AB?D" * 50)
        assert same_answer(ctx.search([1, 3, 4, 2], True, tied, core), port.search([1, 3, 4, 2], True, tied, core))   # and it recovers
    finally:
        _set_engine(ctx, 0)


WEIGHT_ZOO = [[10, 20, 15, 5], [16, 16, 16, 1], [31, 1, 1, 1], [33, 2, 1, 1], [1.5, 0.25, 4, 2], [0.125, 0.5, 0.25, 8],
              [-1, 2, -3, 0.5], [-2, -2, -2, -2], [1e6, 3, 1e-3, 2], [2.0 ** 40, 1, 1, 1], [1e-9, 2e-9, 3e-9, 4e-9],
              [1000, 1, 1, 1], [0, 5, 0, 5], [3, 3, 3, 3]]


@pytest.mark.parametrize("w", WEIGHT_ZOO, ids=[str(w) for w in WEIGHT_ZOO])
def test_weight_zoo(ctx, engine, port, synth, w):
    """Weights on both sides of every internal switch: sliced / transposed epilogue (multiplier limit 31),
    dyadic fixed point, negative and zero weights, exact / re-scored ordering, huge and tiny magnitudes."""
    s1 = synth.letters(301, 2600)
    qs = [synth.letters(302, 40), synth.letters(303, 200), synth.letters(304, 1100), s1[700:760], s1[10:1500]]
    for is_max in (True, False):
        got = ctx.search_batch(w, is_max, s1, qs)
        exp = port.search_batch(w, is_max, s1, qs)
        for k, (g, e) in enumerate(zip(got, exp)):
            assert same_answer(g, e), (w, is_max, k, g, e)


def test_very_long_query_falls_back_to_scalar_engine(ctx, port, synth):
    """len2 > 32767 is beyond the 15 counter planes of the scan engine: the scalar engine takes over."""
    _set_engine(ctx, 0)
    s1, s2 = synth.letters(311, 45000), synth.letters(312, 40000)
    for w, is_max in (([1, 3, 4, 2], False), ([1.5, 2.6, 0.1, 0.2], True)):
        r = ctx.search(w, is_max, s1, s2)
        assert ctx.stat("engine") == 1
        assert same_answer(r, port.search(w, is_max, s1, s2, nthreads=8))


def test_offset_score_profile(ctx, port, synth):
    """psa_offset_scores: the reference's score of EVERY offset (find_best_mutant_offset), not just the best."""
    s1, s2 = synth.letters(401, 5000), synth.letters(402, 333)
    for w in ([1, 3, 4, 2], [1.5, 2.6, 0.1, 0.2]):
        for is_max in (True, False):
            scores, coffs, letters = ctx.offset_scores(w, is_max, s1, s2)
            assert scores == port.scores(w, is_max, s1, s2)
            for n in (0, 17, 2500, 4667):
                o = port.offset_naive(w, is_max, s1, s2, n)
                assert (scores[n], coffs[n], letters[n]) == (o.score, o.char_offset, o.ch)
            best = ctx.search(w, is_max, s1, s2)
            pick = max if is_max else min
            assert best.score == pick(scores) and best.offset == scores.index(best.score)
    sc, _, _ = ctx.offset_scores([1, 3, 4, 2], True, s1, s2, 100, 164)
    assert sc == port.scores([1, 3, 4, 2], True, s1, s2, 100, 164)
    _set_engine(ctx, 0)
    assert ctx.search([1, 3, 4, 2], True, s1, s2).offset >= 0          # the context is still good for searches


def test_many_tiny_ragged_queries(ctx, port, synth):
    """200 000 ragged queries of 1..48 letters against a short Seq1 (some as long as Seq1): per-query bookkeeping at scale."""
    import numpy as np
    _set_engine(ctx, 0)
    s1 = synth.letters(501, 48)
    rng = np.random.default_rng(5)
    lens = rng.integers(1, 49, size=200_000)
    pool = np.frombuffer(synth.letters(502, int(lens.sum())), dtype=np.uint8)
    cuts = np.concatenate(([0], np.cumsum(lens)))
    qs = [pool[cuts[k]:cuts[k + 1]].tobytes() for k in range(len(lens))]
    got = ctx.search_batch([1, 3, 4, 2], False, s1, qs)
    exp = port.search_batch([1, 3, 4, 2], False, s1, qs)
    bad = [k for k, (g, e) in enumerate(zip(got, exp)) if not same_answer(g, e)]
    assert not bad, (len(bad), bad[:5])


def test_fused_finish_matches_separate_finish(ctx, port, synth):
    """Long mode, one tile per query: the scan block finishes its own query (default) or leaves it to k_finish."""
    wl = synth.workload("c3", nq=96)
    res = {}
    for fused in (1, 0):
        _set_engine(ctx, 2, batch=0)
        ctx.set_option("fused_finish", fused)
        res[fused] = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
        assert ctx.stat("kernel_launches") == (2 if fused else 3)
    ctx.set_option("fused_finish", 1)
    _set_engine(ctx, 0)
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    for a, b, e in zip(res[1], res[0], exp):
        assert same_answer(a, e) and same_answer(b, e) and a.counts == b.counts == e.counts


PACK_SHAPES = [(3000, 500, 96), (700, 300, 37), (2100, 1000, 11), (330, 64, 50), (1500, 1023, 9), (4200, 200, 7), (100, 90, 40),
               (1055, 32, 5), (640, 129, 2), (2000, 1, 19)]


@pytest.mark.parametrize("pack,planes,sliced", [(1, -1, 1), (2, -1, 1), (3, 0, 0), (8, 4, 1), (5, 1, 0), (2, 2, 1)])
def test_packed_mode(ctx, port, pack, planes, sliced):
    """Packed mode (k_scan_packed): equal-length queries that each fit one window share blocks lane by lane.  Every
    forced packing factor, plane count and epilogue must give the oracle's answers, on random letters and on
    low-entropy letters (many exact ties and offsets the tracked rank planes do not resolve)."""
    rng = random.Random(1234 + pack)
    used = 0
    for len1, len2, nq in PACK_SHAPES:
        for w, is_max in (([1, 3, 4, 2], True), ([1, 3, 4, 2], False), ([2, 1.5, 1.1, 1.3], True), ([1, 1, 1, 1], True), ([5, 1, 2, 3], False)):
            alpha = rng.choice([ALPHA, ALPHA[:26], "ACDG", "AB", "A"])
            s1 = "".join(rng.choice(alpha) for _ in range(len1))
            qs = ["".join(rng.choice(alpha) for _ in range(len2)) for _ in range(nq)]
            if len2 <= len1 // 2:
                qs[nq // 2] = s1[len1 // 3: len1 // 3 + len2]            # one query that occurs verbatim
            _set_engine(ctx, 2, planes=planes, batch=0, sliced=sliced)
            ctx.set_option("pack_queries", pack)
            try:
                got = ctx.search_batch(w, is_max, s1, qs)
                used += ctx.stat("packed_queries") > 0
                if pack >= 2 and ((len1 - len2 + 1 + 31) // 32) * pack <= 256:
                    assert ctx.stat("packed_queries") == pack, (len1, len2, nq)
            finally:
                ctx.set_option("pack_queries", 1)
                _set_engine(ctx, 0)
            exp = port.search_batch(w, is_max, s1, qs)
            for k, (g, e) in enumerate(zip(got, exp)):
                assert same_answer(g, e), (len1, len2, nq, w, is_max, alpha, k, g, e)
                assert g.counts == e.counts
    assert used > 0


def test_packed_mode_is_chosen_for_config3_shape(ctx, port, synth):
    """2501 offsets = 79 lanes per query: two queries per 5-warp block instead of 3 warps each; turning it off
    changes nothing but the launch shape."""
    wl = synth.workload("c3", nq=65)                                      # odd count: the last block holds one query
    res = {}
    ctx.set_option("stripe_mode", 0)                                      # (stripe mode would take this batch first)
    for pack in (1, 0):
        ctx.set_option("pack_queries", pack)
        res[pack] = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
        assert (ctx.stat("packed_queries"), ctx.stat("packed_warps")) == ((2, 5) if pack else (0, 0))
    ctx.set_option("pack_queries", 1)
    ctx.set_option("stripe_mode", -1)
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    for a, b, e in zip(res[1], res[0], exp):
        assert same_answer(a, e) and same_answer(b, e) and a.counts == b.counts == e.counts


STRIPE_SHAPES = [(3000, 500, 96), (3000, 500, 1), (700, 300, 37), (2100, 1000, 11), (330, 64, 50), (1500, 1023, 9), (4200, 200, 7),
                 (10000, 64, 150), (1055, 32, 5), (640, 129, 2), (2000, 1, 19), (9000, 33, 23), (5000, 777, 3), (96, 64, 300), (20000, 100, 6)]


@pytest.mark.parametrize("planes,derive", [(-1, 1), (0, 1), (1, 0)])
def test_stripe_mode(ctx, port, planes, derive):
    """Stripe mode (k_stripe): one launch per batch -- every block builds the striped window of Seq1 in shared memory, warp
    teams scan Q queries per task against it and finish them.  Every shape (lanes per query from 1 to 622, tasks that
    straddle queries, a last task with fewer queries, steps padded to 32), plane count and both sources of the top-rank
    bit must give the oracle's answers, on random letters and on low-entropy letters (exact ties, unresolved offsets)."""
    rng = random.Random(4321 + planes)
    used = 0
    for len1, len2, nq in STRIPE_SHAPES:
        for w, is_max in (([1, 3, 4, 2], True), ([1, 3, 4, 2], False), ([1, 1, 1, 1], True), ([5, 1, 2, 3], False), ([10, 2, 3, 4], True)):
            alpha = rng.choice([ALPHA, ALPHA[:26], "ACDG", "AB", "A"])
            s1 = "".join(rng.choice(alpha) for _ in range(len1))
            qs = ["".join(rng.choice(alpha) for _ in range(len2)) for _ in range(nq)]
            if len2 <= len1 // 2:
                qs[nq // 2] = s1[len1 // 3: len1 // 3 + len2]            # one query that occurs verbatim
            _set_engine(ctx, 2, planes=planes, derive=derive, stripe=1)
            try:
                got = ctx.search_batch(w, is_max, s1, qs)
                on = ctx.stat("stripe_mode")
                used += on
                if on:
                    assert ctx.stat("kernel_launches") == 1
                    assert ctx.stat("stripe_lanes") == (len1 - len2 + 1 + 31) // 32
            finally:
                _set_engine(ctx, 0)
            exp = port.search_batch(w, is_max, s1, qs)
            for k, (g, e) in enumerate(zip(got, exp)):
                assert same_answer(g, e), (len1, len2, nq, w, is_max, alpha, k, g, e)
                assert g.counts == e.counts
    assert used >= 40, used


def test_stripe_mode_dispatch(ctx, port, synth):
    """The default dispatch takes equal-length batches of queries of 128+ symbols with exact small keys in stripe mode
    (config 3: teams of 5 warps on 2 queries of 79 lanes each); short queries (config 5: one warp per query, 311 lanes, 10
    passes -- when forced), batches whose weights need the re-score path and ragged batches keep the linear-plane kernels."""
    _set_engine(ctx, 0)
    wl = synth.workload("c3", nq=301)
    got = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    assert (ctx.stat("stripe_mode"), ctx.stat("stripe_lanes"), ctx.stat("kernel_launches")) == (1, 79, 1)
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    assert all(same_answer(g, e) and g.counts == e.counts for g, e in zip(got, exp))
    wl = synth.workload("c5", nq=6000)                                    # short queries, many of them: one warp per task
    got = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    assert (ctx.stat("stripe_mode"), ctx.stat("stripe_lanes"), ctx.stat("stripe_team_warps"), ctx.stat("kernel_launches")) == (1, 311, 1, 1)
    ctx.set_option("stripe_mode", 0)
    linear = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    assert (ctx.stat("stripe_mode"), ctx.stat("batch_mode")) == (0, 1)
    ctx.set_option("stripe_mode", -1)
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:500])
    assert all(same_answer(g, e) and g.counts == e.counts for g, e in zip(got, exp))
    assert [(a.offset, a.char_offset, a.ch, a.score, a.counts) for a in linear] == [(a.offset, a.char_offset, a.ch, a.score, a.counts) for a in got]
    few = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:40])  # short queries, few of them: teams of several warps -> linear kernels
    assert ctx.stat("stripe_mode") == 0 and all(same_answer(g, e) for g, e in zip(few, exp))
    wl = synth.workload("c3", nq=40)
    ctx.search_batch([1.5, 2.6, 0.1, 0.2], True, wl.seq1, wl.queries)              # order needs the reference's double
    assert ctx.stat("stripe_mode") == 0
    ctx.search_batch(wl.weights, True, wl.seq1, wl.queries[:-1] + [wl.queries[-1][:499]])   # ragged
    assert ctx.stat("stripe_mode") == 0


def test_zero_copy_results_match_copied_results(psa, ctx, port, synth):
    """Small result sets are written by the kernels straight into page-locked host memory (the caller's array when it
    is page-locked, the context's staging buffer otherwise); the copy-engine path must give the same records."""
    import ctypes as C
    wl = synth.workload("c3", nq=200)
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    batch = psa.Batch(wl.seq1, wl.queries, pinned=True)
    w = (C.c_double * 4)(*wl.weights)
    for zc in (1, 0):
        ctx.set_option("zero_copy_results", zc)
        for pinned in (True, False):
            out = ctx.new_result_array(batch.nq, pinned=pinned)
            ctx.search_batch_raw(w, wl.is_max, batch, out)
            for k, e in enumerate(exp):
                assert same_answer(ctx.result_from_array(out, k), e), (zc, pinned, k)
        # single query (staged for the merge) and a bad symbol (flag path) on both settings
        r = ctx.search(wl.weights, wl.is_max, wl.seq1, wl.queries[3])
        assert same_answer(r, exp[3])
        with pytest.raises(psa.PsaError):
            ctx.search_batch(wl.weights, wl.is_max, wl.seq1, [wl.queries[0], b"AB?D"])
        assert same_answer(ctx.search(wl.weights, wl.is_max, wl.seq1, wl.queries[5]), exp[5])      # and it recovers
    ctx.set_option("zero_copy_results", 1)


def test_streamed_queries_and_large_zero_copy_results(psa, ctx, port, synth):
    """One-shot stripe-mode calls copy their queries on a second stream in pieces while k_stripe is already running
    (stream_wait on per-piece flags) and write every record straight into page-locked host memory, whatever the size of the
    result set.  Same records with streaming on and off, pinned and pageable buffers, one piece and several, a second
    call on the same context (fresh flag tag), and a later split-phase run of the same shape (resident: no flags)."""
    import ctypes as C
    for name, nq, pieces in (("c3", 300, 1), ("c5", 20000, 2), ("c5", 70000, 8)):
        wl = synth.workload(name, nq=nq)
        w = (C.c_double * 4)(*wl.weights)
        exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:150])
        ref_records = None
        for stream in (1, 0, 1):
            ctx.set_option("stream_queries", stream)
            for pinned in (True, False):
                batch = psa.Batch(wl.seq1, wl.queries, pinned=pinned)
                out = ctx.new_result_array(batch.nq, pinned=pinned)
                ctx.search_batch_raw(w, wl.is_max, batch, out)
                assert ctx.stat("stripe_mode") == 1 and ctx.stat("kernel_launches") == 1
                assert ctx.stat("streamed_chunks") == (pieces if stream else 0), (name, nq, stream, ctx.stat("streamed_chunks"))
                recs = [ctx.result_from_array(out, k) for k in range(batch.nq)]
                for k, e in enumerate(exp):
                    assert same_answer(recs[k], e) and recs[k].counts == e.counts, (name, stream, pinned, k)
                flat = [(r.offset, r.char_offset, r.ch, r.score, r.counts) for r in recs]
                if ref_records is None:
                    ref_records = flat
                assert flat == ref_records, (name, stream, pinned)
        ctx.set_option("stream_queries", 1)
        # split phase after a streamed one-shot call of the same shape: the batch is resident, nothing waits for a flag
        batch = psa.Batch(wl.seq1, wl.queries, pinned=True)
        ctx.prepare(wl.weights, wl.is_max, batch)
        ctx.run()
        got = ctx.fetch()
        assert [(r.offset, r.char_offset, r.ch, r.score, r.counts) for r in got] == ref_records
    # a bad symbol in a streamed batch is still reported, and the context recovers
    wl = synth.workload("c3", nq=300)
    bad = list(wl.queries)
    bad[250] = bad[250][:100] + b"?" + bad[250][101:]
    with pytest.raises(psa.PsaError):
        ctx.search_batch(wl.weights, wl.is_max, wl.seq1, bad)
    assert ctx.stat("streamed_chunks") == 1
    got = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:32])
    assert all(same_answer(g, e) for g, e in zip(got, exp))


def test_streamed_queries_with_copied_flags(psa, synth):
    """The same with the stream memory operation unavailable: the flags are 4-byte copies from page-locked memory
    (PSA_NO_STREAM_MEMOPS is read once per process, hence the child process)."""
    code = (
        "import sys, importlib; sys.path.insert(0, %r)\n"
        "psa = importlib.import_module(%r); synth = importlib.import_module(%r + '.synth')\n"
        "import oracle; port = oracle.Port()\n"
        "wl = synth.workload('c5', nq=20000)\n"
        "with psa.Context(1) as ctx:\n"
        "    for rep in range(2):\n"
        "        got = ctx.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)\n"
        "        assert ctx.stat('streamed_chunks') == 2, ctx.stat('streamed_chunks')\n"
        "    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:100])\n"
        "    assert all((g.offset, g.char_offset, g.ch, g.score, g.counts) == (e.offset, e.char_offset, e.ch, e.score, e.counts) for g, e in zip(got, exp))\n"
        "print('ok')\n" % (ROOT, PKG, PKG))
    env = dict(os.environ, PSA_NO_STREAM_MEMOPS="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_gated_timed_runs(psa, port, synth):
    """psa_batch_run behind the host-released gate (option gate_timed_runs): same answers, the first run of a prepared batch
    is never gated (a kernel's first launch must not sit behind a closed gate), an option change in between makes the next
    run a first run again; stripe mode, the single-query launch and the k_profile / k_scan / k_finish chain."""
    with psa.Context(1) as c:
        c.set_option("gate_timed_runs", 1)
        for name, nq in (("c3", 64), ("c2", None), ("c5", 3000)):
            wl = synth.workload(name, nq=nq)
            b = psa.Batch(wl.seq1, wl.queries, pinned=True)
            c.prepare(wl.weights, wl.is_max, b)
            for k in range(4):
                assert c.run() > 0
                if k == 1:
                    c.set_option("kernel_events", 1)
            c.set_option("kernel_events", 0)
            got = c.fetch()
            exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:40])
            assert all(same_answer(g, e) for g, e in zip(got, exp)), name
        # re-scored order (the chain with programmatic dependent launches) and a ragged batch
        wl = synth.workload("c3", nq=40)
        for weights, queries in (([1.5, 2.6, 0.1, 0.2], wl.queries), (wl.weights, wl.queries[:-1] + [wl.queries[-1][:321]])):
            b = psa.Batch(wl.seq1, queries, pinned=True)
            c.prepare(weights, wl.is_max, b)
            for _ in range(3):
                assert c.run() > 0
            exp = port.search_batch(weights, wl.is_max, wl.seq1, queries)
            assert all(same_answer(g, e) for g, e in zip(c.fetch(), exp))


def test_search_many_pipelined(psa, port, synth):
    """psa_search_many: a list of independent problems (different Seq1, weights, goals, batch shapes -- stripe mode, the
    single-query launch, ragged batches, a re-scored order) pipelined over lanes; every problem's records equal what the
    oracle gives for it alone.  One to four lanes on one device slot, two lanes on three slots, a failing problem in the list."""
    rng = random.Random(77)
    items = []
    for k in range(14):
        kind = k % 5
        s1 = synth.letters(200 + k, rng.choice([700, 3000, 5000]))
        if kind == 0:
            qs = [bytes(synth.letters(300 + 31 * k + j, 200)) for j in range(rng.randint(2, 90))]      # equal lengths: stripe mode
        elif kind == 1:
            qs = [bytes(synth.letters(400 + k, rng.randint(64, 600)))]                                  # one query: k_single
        elif kind == 2:
            qs = [bytes(synth.letters(500 + 17 * k + j, rng.randint(1, 300))) for j in range(rng.randint(2, 40))]   # ragged
        elif kind == 3:
            qs = [bytes(synth.letters(600 + 13 * k + j, 64)) for j in range(400)]                        # many short queries
        else:
            qs = [bytes(synth.letters(700 + k + j, 150)) for j in range(5)]
        w = [[1, 3, 4, 2], [1, 1, 1, 1], [1.5, 2.6, 0.1, 0.2], [5, 1, 2, 3]][k % 4]
        items.append((w, bool(k & 1), bytes(s1), qs))
    exp = [port.search_batch(w, mx, s1, qs) for (w, mx, s1, qs) in items]

    def check(got):
        for g_list, e_list in zip(got, exp):
            assert len(g_list) == len(e_list)
            assert all(same_answer(g, e) for g, e in zip(g_list, e_list))

    with psa.Context(1) as c:
        for lanes in (0, 1, 3, 8):
            check(c.search_many(items, lanes=lanes))
        assert c.search_many([], lanes=2) == []
        bad = list(items)
        bad[5] = (bad[5][0], bad[5][1], bad[5][2], [b"AB?D"])
        with pytest.raises(psa.PsaError):
            c.search_many(bad, lanes=2)
        check(c.search_many(items, lanes=2))                               # the lanes recover
        with pytest.raises(psa.PsaError):
            c.search_many(items, lanes=9)
    with psa.Context(devices=[0, 0, 0]) as c:
        check(c.search_many(items, lanes=2))
        check(c.search_many(items[:2], lanes=2))                           # fewer problems than lanes


def test_stripe_split_plans_random(psa, ctx, port):
    """Equal-length batches that are ONE wave of long queries take the split plans of stripe mode (one task per block, the
    last passes of the task cut in two along their steps and merged through shared memory).  Shapes aimed at that regime
    (about 148 x 17 warp passes of work) and random ones around it -- small alphabets plant ties by the thousand -- against
    the oracle (every query of the small batches, a sample of the large ones) and, record for record, against the
    linear-plane kernels (stripe mode off); the team size the plan reports is checked against passes + split."""
    rng = random.Random(4242)
    seen_split = 0
    for trial in range(14):
        aimed = trial < 8
        len1 = rng.randint(700, 2800)
        len2 = rng.randint(256, min(len1 - 1, 700))
        noff = len1 - len2 + 1
        lanes = (noff + 31) // 32
        if aimed:
            q_task = max(1, (rng.choice([15, 16, 17, 18, 19]) * 32) // lanes)
            nq = max(2, 148 * q_task - rng.randint(0, 2 * q_task))
        else:
            nq = max(8, min(rng.randint(20, 420), int(1.1e8 // (noff * len2))))
        alpha = rng.choice([ALPHA, ALPHA[:26], "ACDG", "AB", "NDEQKHRST", "A-"])
        w = rng.choice([[1, 3, 4, 2], [1, 1, 1, 1], [5, 1, 2, 3], [10, 2, 3, 4], [2, 2, 1, 3]])
        is_max = bool(rng.getrandbits(1))
        s1 = "".join(rng.choice(alpha) for _ in range(len1))
        pool = ["".join(rng.choice(alpha) for _ in range(len2)) for _ in range(min(nq, 64))]
        qs = [pool[k % len(pool)] if k >= len(pool) and rng.random() < 0.5 else "".join(rng.choice(alpha) for _ in range(len2)) if k >= len(pool) else pool[k]
              for k in range(nq)]
        at = rng.randrange(noff)
        qs[nq // 2] = s1[at: at + len2]                                   # a query cut out of Seq1: a planted best offset
        _set_engine(ctx, 0)
        got = ctx.search_batch(w, is_max, s1, qs)
        mode, split, q_task, t_warps = ctx.stat("stripe_mode"), ctx.stat("stripe_split"), ctx.stat("stripe_queries_per_task"), ctx.stat("stripe_team_warps")
        sample = list(range(nq)) if not aimed else sorted(set(rng.sample(range(nq), 40) + [0, nq // 2, nq - 1]))
        exp = port.search_batch(w, is_max, s1, [qs[k] for k in sample])
        assert all(same_answer(got[k], e) and got[k].counts == e.counts for k, e in zip(sample, exp)), (trial, len1, len2, nq, w, is_max, mode, split, q_task, t_warps)
        if aimed or trial % 3 == 0:
            ctx.set_option("stripe_mode", 0)
            lin = ctx.search_batch(w, is_max, s1, qs)
            ctx.set_option("stripe_mode", -1)
            assert [(a.offset, a.char_offset, a.ch, a.score, a.counts) for a in lin] == [(a.offset, a.char_offset, a.ch, a.score, a.counts) for a in got], (trial, mode, split)
        if mode:
            seen_split += split > 0
            assert split == 0 or t_warps == (q_task * lanes + 31) // 32 + split, (trial, split, q_task, t_warps)
    assert seen_split >= 2, seen_split                                  # (pass counts that are multiples of four need no split)


def test_small_calls_stay_on_one_device(psa, port, synth):
    """A call is spread over at most (its pair evaluations / min_split_work) device slots: by default a small batch runs on
    one slot of a multi-slot context (idle slots are not even woken), a larger one on as many as its work pays for, and
    min_split_work = 0 splits everything; same answers every time."""
    wl = synth.workload("c3", nq=512)                                     # 6.4e8 pair evaluations
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:48])
    with psa.Context(devices=[0, 0, 0]) as c:
        for knob, used in ((None, 1), (300_000_000, 2), (100_000_000, 3), (0, 3)):
            if knob is not None:
                c.set_option("min_split_work", knob)
            got = c.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries)
            assert c.stat("devices_used") == used, (knob, c.stat("devices_used"))
            assert all(same_answer(g, e) for g, e in zip(got, exp))
        long1 = synth.letters(9, 60_000)                                  # a single query: offset ranges over the three slots
        e1 = port.search(wl.weights, wl.is_max, long1, wl.queries[0], nthreads=8)
        r = c.search(wl.weights, wl.is_max, long1, wl.queries[0])
        assert c.stat("devices_used") == 3 and same_answer(r, e1)
        c.set_option("min_split_work", 2_500_000_000)
        r = c.search(wl.weights, wl.is_max, long1, wl.queries[0])
        assert c.stat("devices_used") == 1 and same_answer(r, e1)


def test_random_batches(ctx, port):
    """Random batches through the default dispatch (long / packed / batch mode, fused or separate finish, exact or
    re-scored order, zero-copy or copied results): equal-length and ragged, tiny and multi-tile, five alphabets."""
    rng = random.Random(20261018)
    wsets = [[1, 3, 4, 2], [1, 1, 1, 1], [2, 1.5, 1.1, 1.3], [5, 1, 2, 3], [0.1, 0.7, 0.3, 0.9], [10, 2, 3, 4], [1.5, 2.6, 0.1, 0.2]]
    modes = set()
    for trial in range(60):
        w = rng.choice(wsets)
        is_max = bool(rng.getrandbits(1))
        len1 = rng.choice([rng.randint(1, 400), rng.randint(400, 3000), rng.randint(3000, 9000)])
        nq = rng.choice([1, 2, 3, rng.randint(4, 40), rng.randint(40, 300)])
        alpha = rng.choice([ALPHA, ALPHA[:26], "ACDG", "AB", "A-"])
        s1 = "".join(rng.choice(alpha) for _ in range(len1))
        if rng.getrandbits(1):                                            # equal lengths
            n2 = rng.choice([1, len1, rng.randint(1, len1), rng.randint(1, min(len1, 200))])
            lens = [n2] * nq
        else:
            lens = [rng.choice([1, len1, rng.randint(1, len1), rng.randint(1, min(len1, 64))]) for _ in range(nq)]
        if sum((len1 - n + 1) * n for n in lens) > 60_000_000:            # keep the oracle's share of the test short
            lens = [min(n, 300) for n in lens]
        qs = ["".join(rng.choice(alpha) for _ in range(n)) for n in lens]
        if lens[0] < len1:
            qs[0] = s1[(len1 - lens[0]) // 2: (len1 - lens[0]) // 2 + lens[0]]     # an exact occurrence
        got = ctx.search_batch(w, is_max, s1, qs)
        modes.add((ctx.stat("batch_mode"), ctx.stat("packed_queries") > 0, ctx.stat("slices") > 1, ctx.stat("exact")))
        exp = port.search_batch(w, is_max, s1, qs)
        for k, (g, e) in enumerate(zip(got, exp)):
            assert same_answer(g, e), (trial, w, is_max, len1, lens[k], nq, alpha, k, g, e)
            assert g.counts == e.counts
    assert len(modes) >= 4, modes


def test_reference_program_links_against_the_library(tmp_path, input_blocks):
    """The drop-in proof: the reference's OWN executable -- its unmodified main.c, cpu_funcs.c (file I/O,
    divide_execute_tasks, the call to gpu_run_program at cpu_funcs.c:180) and mpi_funcs.c, MPI stubbed to one rank --
    linked against libpsa_b200.so (oracle/Makefile target `ref` builds oracle/_ref/mpiCudaOpenMP_dropin where
    /root/reference exists).  With every offset on the GPU its output.txt must be byte-identical to the reference's
    own CPU answer."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "mpiCudaOpenMP_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mpiCudaOpenMP_dropin not built (needs /root/reference)")
    for k, b in enumerate(input_blocks):
        d = tmp_path / f"blk{k}"
        d.mkdir()
        (d / "input.txt").write_text(" ".join(b["weights_text"]) + "\n" + b["seq1"] + "\n" + b["seq2"] + "\n" + b["goal"] + "\n")
        e = b["expect"]
        mut = b["seq2"][: e["char_offset"]] + e["ch"] + b["seq2"][e["char_offset"] + 1:]
        for argv in (["100"], []) if k in (0, 4, 6) else (["100"],):      # no argument = the reference's auto policy (GPU for big problems)
            p = subprocess.run([exe] + argv, cwd=d, capture_output=True, text=True, timeout=120)
            assert p.returncode == 0, (k, argv, p.stderr[-500:])
            assert "CUDA percentage set to 100" in p.stdout, (k, argv, p.stdout)
            assert (d / "output.txt").read_text() == "%s\n%d %s" % (mut, e["offset"], e["score_g"]), (k, argv)


def test_reference_program_links_with_no_reference_cuda_file(tmp_path, input_blocks):
    """The full drop-in: main.c + cpu_funcs.c + mpi_funcs.c of the reference compiled against OUR
    include/cuda_funcs.h and linked against libpsa_b200.so only -- no reference cuda_funcs.o in the link; gpu_run_program
    AND the six host primitives of cuda_funcs.h:44-61 resolve to the library.  output.txt must be byte-identical to the
    reference's own answer for all offsets on the GPU (100), all on the reference's OpenMP loop (0) and its sequential
    loop (-100) -- the CPU loops then run the reference's find_best_mutant_cpu over our primitives."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "mpiCudaOpenMP_dropin2")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mpiCudaOpenMP_dropin2 not built (needs /root/reference)")
    for k, b in enumerate(input_blocks):
        d = tmp_path / f"blk{k}"
        d.mkdir()
        (d / "input.txt").write_text(" ".join(b["weights_text"]) + "\n" + b["seq1"] + "\n" + b["seq2"] + "\n" + b["goal"] + "\n")
        e = b["expect"]
        mut = b["seq2"][: e["char_offset"]] + e["ch"] + b["seq2"][e["char_offset"] + 1:]
        for argv in (["100"], ["0"], ["-100"]) if k in (1, 2, 3, 5, 7, 8, 9) else (["100"],):
            p = subprocess.run([exe] + argv, cwd=d, capture_output=True, text=True, timeout=300)
            assert p.returncode == 0, (k, argv, p.stderr[-500:])
            assert (d / "output.txt").read_text() == "%s\n%d %s" % (mut, e["offset"], e["score_g"]), (k, argv)


def test_topk_offsets(ctx, port, synth):
    """psa_topk_offsets: the k best offsets in the reference order (best score first, ties by ascending offset), selected on the
    device; rank 0 is the search answer."""
    s1, s2 = synth.letters(601, 6000), synth.letters(602, 210)
    core = synth.letters(603, 150)
    tied = synth.letters(604, 700) + core + synth.letters(605, 900) + core + synth.letters(606, 400) + core
    for seq1, seq2 in ((s1, s2), (tied, core)):
        for w in ([1, 3, 4, 2], [1, 1, 1, 1], [1.5, 2.6, 0.1, 0.2]):
            for is_max in (True, False):
                sc = port.scores(w, is_max, seq1, seq2)
                order = sorted(range(len(sc)), key=lambda n: (-sc[n] if is_max else sc[n], n))
                for k, (f, l) in ((1, (0, len(sc))), (17, (0, len(sc))), (40, (100, 131)), (5, (len(sc) - 3, len(sc)))):
                    got = ctx.topk_offsets(w, is_max, seq1, seq2, k, f, l)
                    exp = [n for n in order if f <= n < l][:k]
                    assert [g[0] for g in got] == exp, (w, is_max, k, f, l)
                    assert [g[1] for g in got] == [sc[n] for n in exp]
                    o = port.offset_naive(w, is_max, seq1, seq2, got[0][0])
                    assert (got[0][2], got[0][3]) == (o.char_offset, o.ch)
                best = ctx.search(w, is_max, seq1, seq2)
                top = ctx.topk_offsets(w, is_max, seq1, seq2, 3)
                assert (top[0][0], top[0][1], top[0][2], top[0][3]) == (best.offset, best.score, best.char_offset, best.ch)
    _set_engine(ctx, 0)


def test_mutant_strings_on_the_device(psa, ctx, port, synth):
    """psa_search_batch_mutants: every query with its one substitution applied, written by a kernel (equal-length and ragged
    batches, one query, several device slots)."""
    s1 = synth.letters(611, 4000)
    eq = [synth.letters(620 + k, 300) for k in range(70)]
    ragged = [synth.letters(700 + k, 5 + 37 * k) for k in range(40)]
    for qs in (eq, ragged, [eq[0]]):
        for w, is_max in (([1, 3, 4, 2], True), ([2, 1.5, 1.1, 1.3], False)):
            res, muts = ctx.search_batch_mutants(w, is_max, s1, qs)
            exp = port.search_batch(w, is_max, s1, qs)
            for q, r, m, e in zip(qs, res, muts, exp):
                assert same_answer(r, e)
                assert m == e.mutant(q.decode()) and sum(a != b for a, b in zip(m, q.decode())) <= 1
    with psa.Context(devices=[0, 0, 0]) as c:
        c.set_option("min_split_work", 0)
        res, muts = c.search_batch_mutants([1, 3, 4, 2], False, s1, ragged)
        exp = port.search_batch([1, 3, 4, 2], False, s1, ragged)
        assert [m for m in muts] == [e.mutant(q.decode()) for q, e in zip(ragged, exp)]


def test_query_file_run_and_cli(psa, ctx, port, tmp_path, input_blocks):
    """weights, Seq1 and goal from an input.txt, the queries from a FASTA file; one reference-format stanza per query."""
    import subprocess
    b = input_blocks[5]
    (tmp_path / "input.txt").write_text(" ".join(b["weights_text"]) + "\n" + b["seq1"] + "\n" + b["seq2"] + "\n" + b["goal"] + "\n")
    n1 = len(b["seq1"])
    qs = [b["seq1"][3:3 + n] for n in (5, 9, 17)] + [b["seq2"], "ACDEFGHIKLMNPQRSTVWY"[: min(20, n1)]]
    (tmp_path / "q.fa").write_text("".join(f">q{k}\n{q[:7]}\n{q[7:]}\n" for k, q in enumerate(qs)))
    assert ctx.run_query_file(str(tmp_path / "input.txt"), str(tmp_path / "q.fa"), str(tmp_path / "out.txt")) == len(qs)
    exp = port.search_batch(b["weights"], b["goal"] == "maximum", b["seq1"], qs)
    want = "\n".join("%s\n%d %g" % (e.mutant(q), e.offset, e.score) for q, e in zip(qs, exp))
    assert (tmp_path / "out.txt").read_text() == want
    p = subprocess.run([psa.CLI_PATH, "--queries", "q.fa"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and (tmp_path / "output.txt").read_text() == want, p.stderr
