"""BASELINE config 3 as a parity case: 256 workers, 64 system prompts of 16·U[8,96] tokens with Zipf(1.1) popularity, request = system
prompt + 16·U[1,16] fresh tokens (agent / RAG traffic: multi-tenant shared-prefix nodes are the norm).  Token-tree mode in both batch
modes and event-driven mode (4 bitset words → warp kernel), each against the oracle.  Seed 43."""
import numpy as np
import pytest

from oracle import orc
from smg_b200 import synth

pytestmark = pytest.mark.gpu
CFG = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=16)
W, S = 256, 64


def _traffic(rng):
    prompts = [rng.integers(0, 50000, size=16 * int(rng.integers(8, 97)), dtype=np.uint32) for _ in range(S)]
    zw = 1.0 / np.arange(1, S + 1) ** 1.1
    zw /= zw.sum()

    def request():
        p = prompts[int(rng.choice(S, p=zw))]
        return np.concatenate([p, rng.integers(0, 50000, size=16 * int(rng.integers(1, 17)), dtype=np.uint32)])
    return prompts, request


@pytest.mark.parametrize("mode", ["sequential", "snapshot"])
def test_config3_token_tree_mode(mode):
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy
    rng = np.random.default_rng(43)
    urls = synth.worker_urls(W)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG), tree_batch_mode=mode)
    ws = [BasicWorker(u) for u in urls]
    pol.init_workers(ws)
    orc.reset_globals()
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **CFG)
    op.set_workers(urls)
    _, request = _traffic(rng)
    for batch_no in range(6):
        loads = synth.poisson_loads(W, 8, 43 + batch_no)
        for w, l in zip(ws, loads):
            w.set_load(int(l))
        op.set_state(loads, [1] * W, [1] * W)
        reqs = [request() for _ in range(400)]
        idx, info = pol.select_worker_batch(ws, reqs)
        if mode == "snapshot":
            flat = np.concatenate(reqs).astype(np.uint32)
            off = np.zeros(len(reqs) + 1, np.uint64)
            np.cumsum([len(r) for r in reqs], out=off[1:])
            want, br, ma, _ = op.select_batch_tokens(flat, off, snapshot=True)
            assert np.array_equal(idx, want) and [i.branch for i in info] == list(br) and [i.matched for i in info] == list(ma)
        else:
            for i, r in enumerate(reqs):
                d = op.select_worker(tokens=r)
                assert d.idx == idx[i] and d.branch == orc.BRANCHES[info[i].branch] and d.matched == info[i].matched, (batch_no, i)
        if batch_no >= 2:
            assert sum(1 for i in info if i.branch == 4) > 100      # shared system prompts really route by cache match
    assert pol.token_tree().entries() == op.token_tree().entries()
    pol.evict_cache(2000); op.evict_cache(2000)
    assert pol.token_tree().entries() == op.token_tree().entries()


def test_config3_event_mode():
    from smg_b200 import BasicWorker, CacheAwareConfig, CacheAwarePolicy
    rng = np.random.default_rng(43)
    urls = synth.worker_urls(W)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **CFG))
    ws = [BasicWorker(u) for u in urls]
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(16)
    ix = mon.create_indexer("unknown", 64)
    pol.set_kv_event_monitor(mon)
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **CFG)
    op.set_workers(urls)
    oix = orc.PositionalIndexer(64)
    op.attach_indexer("unknown", oix)
    op.set_kv_event_monitor(True)
    for u in urls:
        assert ix.intern_worker(u) == oix.intern_worker(u)
    prompts, request = _traffic(rng)
    seq = 1
    for p in prompts:                       # every system prompt is cached by a handful of workers, to different depths
        hs = orc.compute_request_content_hashes(p, 16)
        for w in rng.choice(W, size=int(rng.integers(2, 12)), replace=False):
            depth = int(rng.integers(len(hs) // 2, len(hs) + 1))
            blocks = [(seq + i, hs[i]) for i in range(depth)]
            seq += depth
            ix.apply_stored(int(w), blocks)
            oix.apply_stored(int(w), blocks)
    for batch_no in range(3):
        loads = synth.poisson_loads(W, 8, 143 + batch_no)
        healthy = (rng.random(W) > 0.05).astype(np.uint8)
        for w, l, h in zip(ws, loads, healthy):
            w.set_load(int(l)); w.set_healthy(bool(h))
        op.set_state(loads, healthy, [1] * W)
        reqs = [request() for _ in range(1024)]
        idx, info = pol.select_worker_batch(ws, reqs)
        flat = np.concatenate(reqs).astype(np.uint32)
        off = np.zeros(len(reqs) + 1, np.uint64)
        np.cumsum([len(r) for r in reqs], out=off[1:])
        want, br, ma, _ = op.select_batch_tokens(flat, off)
        assert np.array_equal(idx, want) and [i.branch for i in info] == list(br)
        assert [i.matched * 16 for i in info] == list(ma)
        assert (np.asarray(br) == 2).sum() > 900
