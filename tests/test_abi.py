"""The C ABI: the library loads, exports every symbol include/psa_b200.h declares (plus the
reference's C++-mangled entry point), has the reference's record layouts, reads and writes the
reference's file formats, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "psa_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psa_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(psa):
    names = declared_functions()
    assert len(names) >= 20
    lib = C.CDLL(psa.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/psa_b200.h but not exported"


def test_reference_mangled_entry_point(psa):
    """cpu_funcs.c (built as C++) imports _Z15gpu_run_programP5_dataP7_mutantii (cuda_funcs.h:33)."""
    out = subprocess.run(["nm", "-D", "--defined-only", psa.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert re.search(r" T _Z15gpu_run_programP5_dataP7_mutantii$", out, flags=re.M)
    assert re.search(r" T psa_gpu_run_program$", out, flags=re.M)


MANGLED = ["_Z15gpu_run_programP5_dataP7_mutantii", "_Z14get_substituteccPdi", "_Z18get_hashtable_signcc", "_Z10get_weightcPd",
           "_Z13get_pair_signcc", "_Z11is_swapableP7_mutantS0_ddi", "_Z10strlen_gpuPc"]


def test_every_symbol_cpu_funcs_imports_is_exported(psa):
    """cpu_funcs.o of the reference (built as C++) imports gpu_run_program + six host primitives from cuda_funcs.cu
    (cuda_funcs.h:33, 44-61).  The library exports all seven under the same mangled names, so cuda_funcs.cu can leave
    the reference's link altogether; include/cuda_funcs.h declares them."""
    out = subprocess.run(["nm", "-D", "--defined-only", psa.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for name in MANGLED:
        assert re.search(r" T %s$" % re.escape(name), out, flags=re.M), name
    hdr = open(os.path.join(ROOT, "include", "cuda_funcs.h")).read()
    for fn in ("gpu_run_program", "get_substitute", "get_hashtable_sign", "get_weight", "get_pair_sign", "is_swapable", "strlen_gpu"):
        assert re.search(r"\b%s\s*\(" % fn, hdr), fn
    assert "__constant__" not in re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)      # no definitions in the header any more
    # what the reference's own object file asks for (when the reference is present): nothing else out of cuda_funcs.cu
    obj = os.path.join(ROOT, "oracle", "_ref", "d2_c_funcs.o")
    if os.path.exists(obj):
        und = subprocess.run(["nm", "-u", obj], capture_output=True, text=True, check=True).stdout
        wanted = set(re.findall(r"\bU (_Z\S+)", und))
        assert wanted <= set(MANGLED), wanted - set(MANGLED)


def test_host_primitives_match_the_reference(psa, ref):
    """The six exported primitives against the reference's own (compiled from cuda_funcs.cu) on every symbol pair."""
    import itertools
    alpha = [chr(65 + i) for i in range(26)] + ["-"]
    for a, b in itertools.product(alpha, alpha):
        assert psa.get_hashtable_sign(a, b) == ref.sign(a, b), (a, b)
    for a, b in itertools.product(alpha + ["a", "*", "[", "@", " "], repeat=2):
        assert psa.get_pair_sign(a, b) == chr(ref.lib.ref_pair_sign(a.encode(), b.encode())[0]), (a, b)
    for a in ("a", "[", "@", "*", "1"):                                   # outside A..Z and not '-': no sign (cuda_funcs.cu:428-429)
        assert psa.get_hashtable_sign(a, "A") == "" == ref.sign(a, "A") and psa.get_hashtable_sign("Q", a) == "" == ref.sign("Q", a)
    wsets = [[1, 3, 4, 2], [2, 1.5, 1.1, 1.3], [1.5, 2.6, 0.1, 0.2], [0.8, 0.54, 2.6, 13.7], [1, 1, 1, 1], [0, 0, 0, 0], [-1, 2, -3, 0.5]]
    for w in wsets:
        for sign in ("*", ":", ".", "_", "", "x"):
            assert psa.get_weight(sign, w) == ref.weight(sign, w)
        for is_max in (0, 1):
            for a, b in itertools.product(alpha, alpha):
                assert psa.get_substitute(a, b, w, is_max) == ref.substitute(a, b, w, is_max), (w, is_max, a, b)
    M = psa.Mutant
    for (o1, c1, o2, c2) in ((5, 1, 5, 1), (5, 1, 4, 9), (4, 9, 5, 1), (5, 2, 5, 1), (5, 1, 5, 2), (-1, -1, 0, 0)):
        for s1, s2 in ((1.0, 2.0), (2.0, 1.0), (1.5, 1.5), (float("-inf"), 0.0), (float("inf"), float("inf"))):
            for is_max in (0, 1):
                assert psa.is_swapable(M(o1, c1, b"A"), M(o2, c2, b"B"), s1, s2, is_max) == ref.is_swapable(o1, c1, o2, c2, s1, s2, is_max)
    assert psa.strlen_gpu("") == 0 and psa.strlen_gpu("HELLO") == 5 and psa.strlen_gpu("A" * 9999) == 9999


def test_reference_cpu_loops_over_our_primitives(tmp_path, input_blocks):
    """No GPU needed: the reference program linked with NO cuda_funcs.o (oracle/_ref/mpiCudaOpenMP_dropin2, see
    oracle/Makefile) run with argv[1] = 0 (its OpenMP loop) and -100 (its sequential loop): find_best_mutant_cpu then
    calls get_hashtable_sign / get_substitute / get_weight / is_swapable / strlen_gpu out of libpsa_b200.so for every
    pair.  output.txt must be byte-identical to the reference's own answers."""
    exe = os.path.join(ROOT, "oracle", "_ref", "mpiCudaOpenMP_dropin2")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mpiCudaOpenMP_dropin2 not built (needs /root/reference)")
    for k in (1, 2, 3, 5, 7, 8):
        b = input_blocks[k]
        d = tmp_path / f"blk{k}"
        d.mkdir()
        (d / "input.txt").write_text(" ".join(b["weights_text"]) + "\n" + b["seq1"] + "\n" + b["seq2"] + "\n" + b["goal"] + "\n")
        e = b["expect"]
        mut = b["seq2"][: e["char_offset"]] + e["ch"] + b["seq2"][e["char_offset"] + 1:]
        for argv in ("0", "-100"):
            p = subprocess.run([exe, argv], cwd=d, capture_output=True, text=True, timeout=300)
            assert p.returncode == 0, (k, argv, p.stderr[-300:])
            assert (d / "output.txt").read_text() == "%s\n%d %s" % (mut, e["offset"], e["score_g"]), (k, argv)


def test_library_contains_sm100a_code_only(psa):
    out = subprocess.run(["cuobjdump", "-lelf", psa.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_record_layouts(psa):
    assert C.sizeof(psa.ProgramData) == 15048           # program_data.h:6-11
    assert psa.ProgramData.weights.offset == 8 and psa.ProgramData.seq1.offset == 40
    assert psa.ProgramData.seq2.offset == 10041
    assert C.sizeof(psa.Mutant) == 12 and psa.Mutant.ch.offset == 8     # mutant.h:6-10
    assert psa.abi_version() == 1


def test_problem_list_layout_and_argument_checks(psa):
    """psa_problem (psa_search_many) as the ctypes mirror lays it out == the static_assert in psa_engine.cu; a null
    context is refused before anything touches a GPU."""
    P = psa._CProblem
    assert C.sizeof(P) == 64 and P.seq1.offset == 16 and P.out.offset == 48 and P.status.offset == 56
    assert psa._lib.psa_search_many(None, None, 0, 0) == psa.PSA_ERR_ARG


def test_layouts_agree_with_reference_build(psa, ref):
    assert ref.lib.ref_sizeof_program_data() == C.sizeof(psa.ProgramData)
    assert ref.lib.ref_sizeof_mutant() == C.sizeof(psa.Mutant)


def test_input_output_formats(psa, tmp_path, input_blocks):
    b = input_blocks[1]
    p = tmp_path / "input.txt"
    p.write_text("  ".join(b["weights_text"]) + "\n" + b["seq1"] + "\n" + b["seq2"] + "\n" + b["goal"] + "\n\n1 1 1 1\nAAA\nA\nminimum\n")
    w, is_max, s1, s2 = psa.read_seq_and_weights_from_file(str(p))
    assert (w, is_max, s1, s2) == (b["weights"], True, b["seq1"], b["seq2"])
    p.write_text("1 2 3 4 ABC AB maximal")                 # anything but "maximum" is a minimum
    assert psa.read_seq_and_weights_from_file(str(p)) == ([1, 2, 3, 4], False, "ABC", "AB")
    for broken in ("1 2 3", "1 2 3 4 ABC", "x 2 3 4 A B maximum"):
        p.write_text(broken)
        with pytest.raises(psa.PsaError) as e:
            psa.read_seq_and_weights_from_file(str(p))
        assert e.value.status == psa.PSA_ERR_IO
    with pytest.raises(psa.PsaError):
        psa.read_seq_and_weights_from_file(str(tmp_path / "missing.txt"))
    o = tmp_path / "output.txt"
    psa.write_results_to_file(str(o), "HELLO", 4505, -4879.0)
    assert o.read_bytes() == b"HELLO\n4505 -4879"           # "%s\n%d %g", no trailing newline
    psa.write_results_to_file(str(o), "X", 21, 41.7)
    assert o.read_bytes() == b"X\n21 41.7"
    psa.write_results_to_file(str(o), "X", 9, -191.79999999999998)
    assert o.read_bytes() == b"X\n9 -191.8"


def test_io_matches_reference_io(psa, ref, tmp_path, input_blocks):
    b = input_blocks[3]
    p = tmp_path / "input.txt"
    p.write_text(" ".join(b["weights_text"]) + "\n" + b["seq1"] + "\n" + b["seq2"] + "\n" + b["goal"])
    w = (C.c_double * 4)()
    mx = C.c_int()
    s1, s2 = C.create_string_buffer(10001), C.create_string_buffer(5001)
    assert ref.lib.ref_read_input(str(p).encode(), w, C.byref(mx), s1, s2) == 0
    assert psa.read_seq_and_weights_from_file(str(p)) == (list(w), bool(mx.value), s1.value.decode(), s2.value.decode())
    for score in (-4879.0, 30.9, 1e21, -0.000123456789, 5.0):
        ref.lib.ref_write_output(str(tmp_path / "r.txt").encode(), b"MUTANT", 17, score)
        psa.write_results_to_file(str(tmp_path / "o.txt"), "MUTANT", 17, score)
        assert (tmp_path / "r.txt").read_bytes() == (tmp_path / "o.txt").read_bytes()


def test_no_cpu_fallback(psa):
    """On a host without a B200 the product must fail loudly, not compute on the CPU."""
    if psa.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(psa.PsaError) as e:
        psa.Context(1)
    assert e.value.status == psa.PSA_ERR_CUDA


def test_product_does_not_link_the_oracle(psa):
    out = subprocess.run(["ldd", psa.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "psa_ref" not in out
    syms = subprocess.run(["nm", "-D", psa.LIB_PATH], capture_output=True, text=True).stdout
    assert "psa_oracle" not in syms and "ref_search" not in syms
    src = os.path.join(ROOT, "parallel-sequence-alignment_b200")
    for dirpath, _, files in os.walk(src):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "oracle/" not in text and "psa_oracle" not in text, f


def test_reference_arm_does_not_load_the_product_library():
    """bench.py --impl reference must run the reference only: it loads synth.py by file path, so the product package (whose
    import maps libpsa_b200.so) never enters the process."""
    code = ("import sys, bench; s = bench.load_synth(); wl = bench.make_workload(s, 'c3', 0, nq=4); "
            "assert len(wl.queries) == 4 and len(wl.seq1) == 3000; "
            "assert not any('parallel-sequence-alignment_b200' in m for m in sys.modules), [m for m in sys.modules if 'parallel' in m]; "
            "maps = open('/proc/self/maps').read(); assert 'libpsa_b200' not in maps; print('ok')")
    p = subprocess.run([os.sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-800:]


def test_query_file_reader(psa, tmp_path):
    """psa_read_query_file (host only): FASTA records and plain token lists."""
    f = tmp_path / "q.fa"
    f.write_text(">sp|P1|first query\nHELLO\nWORLD\n\n>second\nACDEFGHIK\n;comment\n>empty\n>third  \n  KK-LL \n")
    assert psa.read_query_file(str(f)) == [b"HELLOWORLD", b"ACDEFGHIK", b"KK-LL"]
    f.write_text("HELLO WORLD\nACDEFGHIK\n\n\tKK-LL")
    assert psa.read_query_file(str(f)) == [b"HELLO", b"WORLD", b"ACDEFGHIK", b"KK-LL"]
    f.write_text("")
    assert psa.read_query_file(str(f)) == []
    with pytest.raises(psa.PsaError) as e:
        psa.read_query_file(str(tmp_path / "missing.fa"))
    assert e.value.status == psa.PSA_ERR_IO
