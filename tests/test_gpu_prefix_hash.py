"""prefix_hash policy + consistent hash ring on the GPU (smg_b200/csrc/prefix_hash.cu) through the C ABI: the reference's own unit
tests (hash_ring.rs:152-198, prefix_hash.rs:236-414, names kept), the blake3/xxhash golden vectors, and seeded parity with the
oracle — bit-exact picks, branches, prefix hashes and ring entries."""
import numpy as np
import pytest

from oracle import orc
from tests import prefix_hash_golden as G

pytestmark = pytest.mark.gpu

W3 = ["http://w1:8000", "http://w2:8000", "http://w3:8000"]


def _policy(**cfg):
    from smg_b200 import PrefixHashConfig, PrefixHashPolicy
    return PrefixHashPolicy(PrefixHashConfig(**cfg))


def _workers(urls, loads=None, healthy=None):
    from smg_b200 import BasicWorker
    ws = [BasicWorker(u) for u in urls]
    for i, w in enumerate(ws):
        if loads is not None:
            w.set_load(int(loads[i]))
        if healthy is not None:
            w.set_healthy(bool(healthy[i]))
    return ws


def _info(tokens, ring):
    from smg_b200 import SelectWorkerInfo
    return SelectWorkerInfo(tokens=tokens, hash_ring=ring)


# ---- hash_ring.rs tests ----
def test_empty_ring_returns_none():
    from smg_b200 import HashRing
    ring = HashRing([])
    assert ring.is_empty() and len(ring) == 0 and ring.worker_count() == 0
    assert ring.find_healthy_url("any-key", lambda u: True) is None


def test_len_scales_with_virtual_nodes():
    from smg_b200 import HashRing
    ring = HashRing(["http://a", "http://b", "http://c"])
    assert not ring.is_empty() and len(ring) == 450 and ring.worker_count() == 3


def test_find_healthy_url_deterministic_skips_unhealthy_none_when_all_unhealthy():
    from smg_b200 import HashRing
    ring = HashRing(["http://a", "http://b", "http://c"])
    first = ring.find_healthy_url("routing-key", lambda u: True)
    assert first is not None and all(ring.find_healthy_url("routing-key", lambda u: True) == first for _ in range(10))
    assert ring.find_healthy_url("routing-key", lambda u: u != "http://a") in ("http://b", "http://c")
    assert HashRing(["http://a", "http://b"]).find_healthy_url("k", lambda u: False) is None


# ---- prefix_hash.rs tests ----
def test_prefix_hash_consistent_routing():
    pol = _policy()
    ws, ring = _workers(W3), pol.hash_ring(W3)
    first, _ = pol.select_worker_impl(ws, _info(list(range(1, 11)), ring))
    assert first is not None
    for _ in range(10):
        assert pol.select_worker(ws, _info(list(range(1, 11)), ring)) == first


def test_different_prefixes_distribute():
    pol = _policy()
    ws, ring = _workers(W3), pol.hash_ring(W3)
    idx, _ = pol.select_worker_batch(ws, [[i, i + 1, i + 2, i + 3] for i in range(100)], ring=ring)
    assert len(set(idx.tolist())) > 1


def test_shared_prefix_routes_same():
    pol = _policy(prefix_token_count=5)
    ws, ring = _workers(W3), pol.hash_ring(W3)
    assert pol.select_worker(ws, _info([1, 2, 3, 4, 5, 100, 200, 300], ring)) == pol.select_worker(ws, _info([1, 2, 3, 4, 5, 999, 888, 777], ring))


def test_no_tokens_returns_none():
    pol = _policy()
    ws, ring = _workers(W3[:1]), pol.hash_ring(W3[:1])
    assert pol.select_worker_impl(ws, _info([], ring)) == (None, "no_tokens")
    assert pol.select_worker_impl(ws, _info(None, ring)) == (None, "no_tokens")


def test_no_healthy_workers():
    pol = _policy()
    ws, ring = _workers(W3[:1], healthy=[0]), pol.hash_ring(W3[:1])
    assert pol.select_worker_impl(ws, _info([1, 2, 3], ring)) == (None, "no_healthy_workers")
    assert pol.select_worker_impl([], _info([1, 2, 3], ring)) == (None, "no_healthy_workers")


def test_policy_name_and_factory():
    from smg_b200 import PolicyFactory
    assert _policy().name() == "prefix_hash"
    assert PolicyFactory.create_by_name("PrefixHash").name() == "prefix_hash"


def test_overloaded_initial_walks_and_no_ring_fallback():   # prefix_hash.rs:171-199
    pol = _policy()
    ring = pol.hash_ring(W3)
    first, br = pol.select_worker_impl(_workers(W3), _info([7, 7, 7], ring))
    assert br == "ring_hit"
    loads = [1, 1, 1]
    loads[first] = 50
    idx, br = pol.select_worker_impl(_workers(W3, loads), _info([7, 7, 7], ring))
    assert br == "load_balance_walk" and idx == min(i for i in range(3) if i != first)
    zero = _policy(load_factor=0.0)
    assert zero.select_worker_impl(_workers(W3, [5, 5, 5]), _info([7, 7, 7], zero.hash_ring(W3))) == (first, "load_balance_walk")
    assert pol.select_worker_impl(_workers(W3, [4, 2, 2]), _info([7, 7, 7], None)) == (1, "fallback_least_load")


# ---- goldens ----
def test_golden_rings_and_positions():
    from smg_b200 import HashRing
    g = G.load()
    probe = HashRing(["http://a"])
    for r in g["rings"]:
        ring = HashRing(r["urls"])
        pos, url = ring.entries()
        assert len(ring) == r["len"]
        assert [[str(int(p)), int(u)] for p, u in zip(pos[:12], url[:12])] == r["head"]
        assert [[str(int(p)), int(u)] for p, u in zip(pos[-4:], url[-4:])] == r["tail"]
        assert int(np.bitwise_xor.reduce(pos)) == int(r["xor_of_positions"])
        assert sum((i + 1) * (int(u) + 1) for i, u in enumerate(url)) % (1 << 61) == r["url_checksum"]
        opos, ourl = orc.HashRing(r["urls"]).entries()
        assert np.array_equal(pos, opos) and np.array_equal(url, ourl)
    # key positions, observed through find_healthy_url: the chosen URL must be the oracle's for every golden key
    oring = orc.HashRing(["http://a", "http://b", "http://c"])
    ring = HashRing(["http://a", "http://b", "http://c"])
    keys = [e["key"] for e in g["positions"]] + [f"key-{i}" for i in range(500)]
    for ok in (lambda u: True, lambda u: u != "http://b", lambda u: u == "http://c"):
        assert ring.find_healthy_urls(keys, ok) == [oring.find_healthy_url(k, ok) for k in keys]
    del probe


def test_golden_prefix_hashes():
    g = G.load()
    by_k = {}
    for e in g["prefix_hashes"]:
        by_k.setdefault(e["k"], []).append(e)
    for k, es in by_k.items():
        pol = _policy(prefix_token_count=k)
        got = pol.compute_prefix_hashes([e["tokens"] if "tokens" in e else G.stream(e["seed"], e["n"]) for e in es])
        assert [int(h) for h in got] == [int(e["hash"]) for e in es], k


def test_golden_decisions():
    n = 0
    for c in G.load()["decisions"]:
        pol = _policy(prefix_token_count=c["prefix_token_count"], load_factor=c["load_factor"])
        ring = None if c["ring_urls"] is None else pol.hash_ring(c["ring_urls"])
        ws = _workers(c["urls"], c["loads"], c["healthy"])
        idx, br = pol.select_worker_batch(ws, [G.expand(p) for p in c["requests"]], ring=ring)
        for i, (want_idx, want_br) in enumerate(c["picks"]):
            assert (int(idx[i]), br[i]) == (want_idx, want_br), (c["urls"][:2], i)
            n += 1
    assert n == 960


# ---- seeded parity with the oracle at size ----
@pytest.mark.parametrize("seed,n_workers,k,factor", [(1, 64, 256, 1.25), (2, 7, 256, 1.0), (3, 512, 64, 1.25), (4, 33, 300, 2.0), (5, 16, 1000, 1.25),
                                                      (6, 3, 0, 1.25)])
def test_random_stream_parity_with_oracle(seed, n_workers, k, factor):
    rng = np.random.default_rng(seed)
    urls = [f"http://worker-{i}.svc:{8000 + i}" for i in range(n_workers)]
    pol = _policy(prefix_token_count=k, load_factor=factor)
    opol = orc.PrefixHashPolicy(k, factor)
    ring_urls = urls if seed != 4 else urls[:20] + ["http://ghost:1", "http://ghost:2"]
    ring, oring = pol.hash_ring(ring_urls), orc.HashRing(ring_urls)
    lens = [0, 1, 2, 3, 4, 5, 17, 32, 33, 60, 61, 62, 63, 64, 65, 127, 128, 255, 256, 257, 300, 511, 512, 513, 700, 1023, 1024, 1025, 2047, 2048, 2049, 2500]
    for rnd in range(4):
        loads = rng.integers(0, [1, 6, 40, 400][rnd], size=n_workers)
        healthy = (rng.random(n_workers) > [0.0, 0.1, 0.5, 0.97][rnd]).astype(np.uint8)
        reqs = []
        for _ in range(700):
            n = int(rng.choice(lens)) if rng.random() < 0.6 else int(rng.integers(0, 600))
            t = rng.integers(0, 1 << 32 if rnd == 1 else 128000, size=n, dtype=np.uint64).astype(np.uint32)
            if reqs and rng.random() < 0.2:
                t = np.concatenate([reqs[-1][: max(k, 1)], t]).astype(np.uint32)
            reqs.append(t)
        ws = _workers(urls, loads, healthy)
        idx, br = pol.select_worker_batch(ws, reqs, ring=ring)
        flat = np.concatenate(reqs) if sum(len(r) for r in reqs) else np.zeros(0, np.uint32)
        off = np.zeros(len(reqs) + 1, np.uint64)
        np.cumsum([len(r) for r in reqs], out=off[1:])
        oidx, obr, _ = opol.select_batch(urls, loads, healthy, oring, flat, off)
        assert np.array_equal(idx, oidx), (rnd, np.nonzero(idx != oidx)[0][:5])
        assert br == [orc.PREFIX_BRANCHES[int(b)] for b in obr]
        got = pol.compute_prefix_hashes(reqs)
        assert [int(h) for h in got] == [opol.compute_prefix_hash(r) for r in reqs]


def test_duplicate_urls_in_slice_resolve_to_last_healthy():   # healthy_url_map is collected over healthy workers; later duplicates overwrite
    urls = ["http://a", "http://b", "http://a", "http://c", "http://a"]
    pol, opol = _policy(), orc.PrefixHashPolicy()
    ring, oring = pol.hash_ring(["http://a", "http://b", "http://c"]), orc.HashRing(["http://a", "http://b", "http://c"])
    reqs = [[i, i * 7, 3] for i in range(300)]
    flat = np.array([t for r in reqs for t in r], np.uint32)
    off = np.arange(0, 3 * len(reqs) + 1, 3, dtype=np.uint64)
    for healthy in ([1, 1, 1, 1, 1], [1, 1, 1, 1, 0], [1, 1, 0, 1, 0], [0, 1, 0, 1, 0], [0, 0, 1, 0, 0]):
        idx, br = pol.select_worker_batch(_workers(urls, [0] * 5, healthy), reqs, ring=ring)
        oidx, obr, _ = opol.select_batch(urls, [0] * 5, healthy, oring, flat, off)
        assert np.array_equal(idx, oidx) and br == [orc.PREFIX_BRANCHES[int(b)] for b in obr], healthy


def test_unaligned_request_starts_and_device_batches():
    """Odd token offsets take the 4-byte load path; the device-resident multi-batch entry must agree with the host-buffer call."""
    import ctypes as C
    from smg_b200 import _lib
    rng = np.random.default_rng(9)
    urls = [f"http://w{i}:8000" for i in range(64)]
    pol, opol = _policy(), orc.PrefixHashPolicy()
    ring, oring = pol.hash_ring(urls), orc.HashRing(urls)
    loads = rng.integers(0, 30, size=64)
    ws = _workers(urls, loads)
    h, model = pol._h, pol._push_fleet(ws, ring)
    batches = []
    for b in range(3):
        reqs = [rng.integers(0, 128000, size=int(rng.choice([61, 63, 255, 257, 301, 512])), dtype=np.uint64).astype(np.uint32) for _ in range(200)]
        tokens = np.concatenate(reqs)
        off = np.zeros(len(reqs) + 1, np.uint32)
        np.cumsum([len(r) for r in reqs], out=off[1:])
        batches.append((tokens, off))
    err = _lib.new_err()
    L = h.L
    d_tok, d_off, d_out = [], [], []
    for tokens, off in batches:
        for arr, lst in ((tokens, d_tok), (off, d_off)):
            ptr = L.smgx_device_alloc(h.p, arr.nbytes, C.byref(err))
            h.call("smgx_memcpy_h2d", ptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes)
            lst.append(ptr)
        d_out.append(L.smgx_device_alloc(h.p, (off.size - 1) * 4, C.byref(err)))
    nb = len(batches)
    vp = C.c_void_p * nb
    ns = (C.c_uint32 * nb)(*[off.size - 1 for _, off in batches])
    h.call("smgx_prefix_hash_select_many_tokens_device", model, nb, vp(*d_tok), vp(*d_off), ns, vp(*d_out))
    h.call("smgx_synchronize")
    for k, (tokens, off) in enumerate(batches):
        got = np.zeros(off.size - 1, np.int32)
        h.call("smgx_memcpy_d2h", got.ctypes.data_as(C.c_void_p), d_out[k], got.nbytes)
        oidx, _, _ = opol.select_batch(urls, loads, [1] * 64, oring, tokens, off.astype(np.uint64))
        assert np.array_equal(got, oidx), k
    for ptr in d_tok + d_off + d_out:
        L.smgx_device_free(h.p, ptr)
