"""Test plumbing.

* ``-m "not gpu"``: oracle vs golden vectors / the compiled reference, host logic (table resolver,
  I/O, partitioning), and that the C-ABI library loads and exports what include/psa_b200.h declares.
* ``-m gpu``: parity of the CUDA path against the oracle, through the C ABI.

The oracle (oracle/) is imported here and only here (plus smoke() and bench.py's cpu baseline).
"""
import importlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "parallel-sequence-alignment_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # make sure the product library exists (nvcc cross-compiles without a GPU); never rebuild if present
    lib = os.path.join(ROOT, PKG, "libpsa_b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, PKG)], check=True)
    port = os.path.join(ROOT, "oracle", "libpsa_oracle.so")
    if not os.path.exists(port):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)


@pytest.fixture(scope="session")
def psa():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module(PKG + ".synth")


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.Port()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libpsa_ref.so not built (needs /root/reference)")
    return oracle.Ref()


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def input_blocks():
    return load_golden("input_blocks.json")


@pytest.fixture(scope="session")
def synthetic_cases():
    return load_golden("synthetic.json")


@pytest.fixture(scope="session")
def ctx(psa):
    """One single-GPU context for the whole gpu session (buffers persist across calls by design)."""
    c = psa.Context(1)
    yield c
    c.close()


def same_answer(got, exp):
    """got: Result-like (offset,char_offset,ch,score); exp: golden 'expect' dict or Result."""
    if isinstance(exp, dict):
        e = (exp["offset"], exp["char_offset"], exp["ch"], exp["score"])
    else:
        e = (exp.offset, exp.char_offset, exp.ch, exp.score)
    return (got.offset, got.char_offset, got.ch, got.score) == e
