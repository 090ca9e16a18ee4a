"""Pins the prefix_hash / hash-ring oracle (oracle/prefix_hash.h): the reference's own unit tests
(model_gateway/src/worker/hash_ring.rs:152-198, model_gateway/src/policies/prefix_hash.rs:236-414, names kept) and golden
vectors made with the independent Python blake3 / xxhash packages."""
import numpy as np
import pytest

from oracle import orc
from tests import prefix_hash_golden as G

W3 = ["http://w1:8000", "http://w2:8000", "http://w3:8000"]


def _pick(pol, urls, ring, tokens, loads=None, healthy=None):
    return pol.select_worker(urls, loads or [0] * len(urls), healthy or [1] * len(urls), ring, tokens)


# ---- hash_ring.rs tests ----
def test_empty_ring_returns_none():
    ring = orc.HashRing([])
    assert ring.is_empty() and len(ring) == 0 and ring.worker_count() == 0
    assert ring.find_healthy_url("any-key", lambda u: True) is None


def test_len_scales_with_virtual_nodes():
    ring = orc.HashRing(["http://a", "http://b", "http://c"])
    assert not ring.is_empty() and len(ring) == 3 * 150 and ring.worker_count() == 3


def test_find_healthy_url_is_deterministic():
    ring = orc.HashRing(["http://a", "http://b", "http://c"])
    first = ring.find_healthy_url("routing-key", lambda u: True)
    assert first is not None
    for _ in range(10):
        assert ring.find_healthy_url("routing-key", lambda u: True) == first


def test_find_healthy_url_skips_unhealthy():
    ring = orc.HashRing(["http://a", "http://b", "http://c"])
    assert ring.find_healthy_url("routing-key", lambda u: u != "http://a") in ("http://b", "http://c")


def test_find_healthy_url_returns_none_when_all_unhealthy():
    assert orc.HashRing(["http://a", "http://b"]).find_healthy_url("k", lambda u: False) is None


# ---- prefix_hash.rs tests ----
def test_prefix_hash_consistent_routing():
    pol, ring = orc.PrefixHashPolicy(), orc.HashRing(W3)
    first, _ = _pick(pol, W3, ring, list(range(1, 11)))
    assert first is not None
    for _ in range(10):
        assert _pick(pol, W3, ring, list(range(1, 11)))[0] == first


def test_different_prefixes_distribute():
    pol, ring = orc.PrefixHashPolicy(), orc.HashRing(W3)
    seen = {_pick(pol, W3, ring, [i, i + 1, i + 2, i + 3])[0] for i in range(100)}
    assert len(seen) > 1


def test_shared_prefix_routes_same():
    pol, ring = orc.PrefixHashPolicy(prefix_token_count=5), orc.HashRing(W3)
    assert _pick(pol, W3, ring, [1, 2, 3, 4, 5, 100, 200, 300])[0] == _pick(pol, W3, ring, [1, 2, 3, 4, 5, 999, 888, 777])[0]


def test_no_tokens_returns_none():
    pol, ring = orc.PrefixHashPolicy(), orc.HashRing(W3[:1])
    assert _pick(pol, W3[:1], ring, []) == (None, "no_tokens")
    assert _pick(pol, W3[:1], ring, None) == (None, "no_tokens")


def test_no_healthy_workers():
    pol, ring = orc.PrefixHashPolicy(), orc.HashRing(W3[:1])
    assert _pick(pol, W3[:1], ring, [1, 2, 3], healthy=[0]) == (None, "no_healthy_workers")
    assert pol.select_worker([], [], [], ring, [1, 2, 3]) == (None, "no_healthy_workers")


def test_load_ok_calculation():
    pol = orc.PrefixHashPolicy(load_factor=1.25)
    assert pol.load_ok(30, 100, 4) and not pol.load_ok(35, 100, 4)
    assert pol.load_ok(0, 0, 4) and pol.load_ok(100, 0, 0)


def test_policy_name():
    assert orc.PrefixHashPolicy().name() == "prefix_hash"


def test_overloaded_initial_walks_to_least_loaded_ok():   # prefix_hash.rs:171-187
    pol, ring = orc.PrefixHashPolicy(), orc.HashRing(W3)
    first, br = _pick(pol, W3, ring, [7, 7, 7])
    assert br == "ring_hit"
    loads = [1, 1, 1]
    loads[first] = 50
    idx, br = _pick(pol, W3, ring, [7, 7, 7], loads=loads)
    assert br == "load_balance_walk" and idx == min(i for i in range(3) if i != first)   # FIRST minimum among the acceptable ones
    # nobody passes the check (factor 0) → the initial worker anyway
    idx, br = _pick(orc.PrefixHashPolicy(load_factor=0.0), W3, ring, [7, 7, 7], loads=[5, 5, 5])
    assert br == "load_balance_walk" and idx == first
    # no ring → least loaded healthy
    assert _pick(pol, W3, None, [7, 7, 7], loads=[4, 2, 2]) == (1, "fallback_least_load")


# ---- goldens (independent blake3 / xxhash) ----
def test_golden_ring_positions_and_rings():
    g = G.load()
    for e in g["positions"]:
        assert orc.HashRing.hash_position(e["key"]) == int(e["pos"])
    for r in g["rings"]:
        ring = orc.HashRing(r["urls"])
        pos, url = ring.entries()
        assert len(ring) == r["len"]
        assert [[str(int(p)), int(u)] for p, u in zip(pos[:12], url[:12])] == r["head"]
        assert [[str(int(p)), int(u)] for p, u in zip(pos[-4:], url[-4:])] == r["tail"]
        assert int(np.bitwise_xor.reduce(pos)) == int(r["xor_of_positions"])
        assert sum((i + 1) * (int(u) + 1) for i, u in enumerate(url)) % (1 << 61) == r["url_checksum"]
        assert np.all(pos[1:] >= pos[:-1])


def test_golden_prefix_hashes():
    for e in G.load()["prefix_hashes"]:
        t = e["tokens"] if "tokens" in e else G.stream(e["seed"], e["n"])
        assert orc.PrefixHashPolicy(prefix_token_count=e["k"]).compute_prefix_hash(t) == int(e["hash"]), e


def test_golden_decisions():
    n = 0
    for c in G.load()["decisions"]:
        pol = orc.PrefixHashPolicy(c["prefix_token_count"], c["load_factor"])
        ring = None if c["ring_urls"] is None else orc.HashRing(c["ring_urls"])
        reqs = [G.expand(p) for p in c["requests"]]
        flat = np.array([t for r in reqs for t in r], np.uint32)
        off = np.zeros(len(reqs) + 1, np.uint64)
        np.cumsum([len(r) for r in reqs], out=off[1:])
        idx, br, _ = pol.select_batch(c["urls"], c["loads"], c["healthy"], ring, flat, off)
        for i, (want_idx, want_br) in enumerate(c["picks"]):
            assert (int(idx[i]), orc.PREFIX_BRANCHES[int(br[i])]) == (want_idx, want_br), (c["urls"][:2], i)
            n += 1
    assert n == 960
