"""The C++ host mirror (include/smgx.hpp) — the reference is compiled code (Rust), so the interface a maintainer programs against is
mirrored in C++ too.  tests/cpp/test_cache_aware.cpp ports the reference's own cache_aware.rs unit tests and replays seeded streams
against the CPU oracle; this module builds it and runs the host-only subset on CPU and the whole program on a B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
LIB = os.path.join(ROOT, "smg_b200", "libsmgx.so")


def _binary():
    if not os.path.exists(LIB):
        pytest.fail("smg_b200/libsmgx.so is missing: run __graft_entry__.build()")
    subprocess.check_call(["make", "-s", "-C", CPP, "test_cache_aware"])
    return os.path.join(CPP, "test_cache_aware")


def _run(args):
    r = subprocess.run([_binary(), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_cpp_mirror_host_subset():
    out = _run(["--host"])
    assert "0 failures" in out and "3 tests" in out


@pytest.mark.gpu
def test_cpp_mirror_full():
    out = _run([])
    assert "0 failures" in out and "20 tests" in out, out
