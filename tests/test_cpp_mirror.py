"""The C++ host mirror (include/smgx.hpp) — the reference is compiled code (Rust), so the interface a maintainer programs against is
mirrored in C++ too.  tests/cpp/test_cache_aware.cpp ports the reference's own cache_aware.rs unit tests and replays seeded streams
against the CPU oracle; this module builds it and runs the host-only subset on CPU and the whole program on a B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
LIB = os.path.join(ROOT, "smg_b200", "libsmgx.so")


def _binary(name="test_cache_aware"):
    if not os.path.exists(LIB):
        pytest.fail("smg_b200/libsmgx.so is missing: run __graft_entry__.build()")
    subprocess.check_call(["make", "-s", "-C", CPP, name])
    return os.path.join(CPP, name)


def _run(args, name="test_cache_aware", timeout=300):
    r = subprocess.run([_binary(name), *args], capture_output=True, text=True, timeout=timeout)   # a hang fails the test instead of the box
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_cpp_mirror_host_subset():
    out = _run(["--host"])
    assert "0 failures" in out and "4 tests" in out


@pytest.mark.gpu
def test_cpp_mirror_full():
    out = _run([])
    assert "0 failures" in out and "21 tests" in out, out


def test_cpp_batcher_builds():
    assert os.path.exists(_binary("test_batcher"))


def test_cpp_batcher_logic_against_mock_device():
    """smgx::Batcher's concurrency (slot reservation, ring growth, back-pressure, ticket redemption, exhaustion report) with the five
    C-ABI calls it makes replaced by a mock 4-lane device: runs on CPU, and a regression shows up here as a timeout, not on a GPU box."""
    out = _run([], name="test_batcher_logic", timeout=240)
    assert out.strip().endswith("ok") and " 0 wrong" in out


@pytest.mark.gpu
@pytest.mark.parametrize("threads,per_thread,window,wait_us,mapped", [(8, 1500, 64, 100, 1), (32, 400, 1, 50, 1), (8, 1500, 64, 100, 0), (32, 400, 1, 50, 0),
                                                                      (3, 2000, 700, 200, 0)])
def test_cpp_batcher_picks_equal_oracle(threads, per_thread, window, wait_us, mapped):
    """Many caller threads through smgx::Batcher (include/smgx_batcher.hpp), both transports — mapped = the zero-copy latency path
    (smgx_submit_tokens_mapped, group commit, callers spin on the completion word), staged = smgx_submit_tokens / smgx_wait: every
    per-request pick equals the oracle's; blocking route() (window 1), a task pool with 64 outstanding requests and, on the staged
    transport, windows wide enough to fill 4096-request batches."""
    import json
    out = _run([str(threads), str(per_thread), str(window), str(wait_us), str(mapped)], name="test_batcher", timeout=120)
    res = json.loads(out.strip().splitlines()[-1])
    assert res["mismatches_vs_oracle"] == 0 and res["requests"] == threads * per_thread
    assert res["batches"] >= 1 and res["decisions_per_s"] > 0
    assert ("mapped" in res["transport"]) == bool(mapped)
