"""KV-event ingest: KvEventMonitor::apply_event through smgx_kv_events_apply — the reference's unit tests
(model_gateway/src/worker/kv_event_monitor.rs:629-760: convert_kv_block incl. the negative i64 block hash, stored with / without parent,
fresh-chain fallback, removed, cleared) and a random event stream compared with the oracle's apply_event + find_matches."""
import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu


def _mk(jump=64):
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0))
    mon = pol.kv_event_monitor(None)
    ix = mon.create_indexer("unknown", jump)
    pol.set_kv_event_monitor(mon)
    return pol, mon, ix


def _blk(h, toks, bs=4):
    return {"block_hash": h, "token_ids": toks, "block_size": bs}


def test_convert_kv_block_and_negative_hash():  # :629-668
    pol, mon, ix = _mk()
    w = ix.intern_worker("http://w1:8000")
    mon.apply_events("unknown", w, [{"stored": {"blocks": [_blk(42, [1, 2, 3, 4])], "parent_block_hash": None}}])
    assert ix.find_matches([orc.compute_content_hash([1, 2, 3, 4])])[0] == {w: 1}
    # block_hash −1 is SequenceHash(u64::MAX): a child chained to it by its u64 value must find its parent
    mon.apply_events("unknown", w, [{"stored": {"blocks": [_blk(-1, [10, 20], 2)], "parent_block_hash": 42}}])
    assert mon.apply_events("unknown", w, [{"stored": {"blocks": [_blk(7, [30, 40], 2)], "parent_block_hash": -1}}]) == 0
    assert ix.current_size() == 3
    # empty token_ids hash like compute_content_hash(&[])
    mon.apply_events("unknown", w, [{"stored": {"blocks": [_blk(100, [], 0)], "parent_block_hash": None}}])
    assert ix.find_matches([orc.compute_content_hash([])])[0] == {w: 1}


def test_apply_stored_parent_chain_and_fallback():  # :674-752
    pol, mon, ix = _mk()
    w = ix.intern_worker("http://w1:8000")
    mon.apply_events("unknown", w, [{"stored": {"blocks": [_blk(1, [10, 20, 30, 40]), _blk(2, [50, 60, 70, 80])], "parent_block_hash": None}}])
    assert ix.current_size() == 2
    pol2, mon2, ix2 = _mk()
    w2 = ix2.intern_worker("http://w1:8000")
    assert mon2.apply_events("unknown", w2, [{"stored": {"blocks": [_blk(1, [10, 20, 30, 40])], "parent_block_hash": None}}]) == 0
    assert mon2.apply_events("unknown", w2, [{"stored": {"blocks": [_blk(2, [50, 60, 70, 80])], "parent_block_hash": 1}}]) == 0
    assert ix2.current_size() == 2
    pol3, mon3, ix3 = _mk()
    w3 = ix3.intern_worker("http://new-worker:8000")
    assert mon3.apply_events("unknown", w3, [{"stored": {"blocks": [_blk(1, [10, 20, 30, 40])], "parent_block_hash": 999}}]) == 1   # fresh chain
    assert ix3.current_size() == 1


def test_removed_cleared_and_block_size_learning():
    pol, mon, ix = _mk()
    w = ix.intern_worker("http://w1:8000")
    mon.apply_events("unknown", w, [{"stored": {"blocks": [_blk(1, [1, 2, 3, 4]), _blk(2, [5, 6, 7, 8])], "parent_block_hash": None}},
                                    {"removed": {"block_hashes": [2]}}])
    assert ix.current_size() == 1
    mon.apply_events("unknown", w, [{"cleared": {}}])
    assert ix.current_size() == 0
    # the first stored block taught block_size = 4 (learn_block_size :270-296): a 4-token request now scores through the event path
    from smg_b200 import BasicWorker
    ws = [BasicWorker("http://w1:8000")]
    pol.init_workers(ws)
    mon.apply_events("unknown", w, [{"stored": {"blocks": [_blk(9, [1, 2, 3, 4])], "parent_block_hash": None}}])
    idx, info = pol.select_worker_batch(ws, [[1, 2, 3, 4]])
    assert idx[0] == 0 and info[0].branch == 2 and info[0].matched == 1


@pytest.mark.parametrize("seed", [1, 2])
def test_random_event_stream_matches_oracle(seed):
    rng = np.random.default_rng(seed)
    pol, mon, ix = _mk(jump=8)
    oix = orc.PositionalIndexer(8)
    urls = [f"http://w{i}:8000" for i in range(6)]
    wids = [ix.intern_worker(u) for u in urls]
    assert wids == [oix.intern_worker(u) for u in urls]
    seqs = [rng.integers(0, 300, size=(int(rng.integers(2, 20)), 4)) for _ in range(25)]   # 25 sequences of 4-token blocks, shared prefixes below
    for s in seqs[5:]:                                # shared prefixes: copy the first k blocks of one of the first five sequences
        base = seqs[int(rng.integers(0, 5))]
        k = int(rng.integers(0, min(len(s), len(base)) + 1))
        s[:k] = base[:k]
    stored = {w: [] for w in wids}      # block hashes a worker currently holds, in chain order per store
    next_hash = [1]
    fall_g = fall_o = 0
    for step in range(60):
        w = int(rng.choice(wids))
        events = []
        for _ in range(int(rng.integers(1, 5))):
            r = rng.random()
            if r < 0.65:
                s = seqs[int(rng.integers(0, len(seqs)))]
                depth = int(rng.integers(1, len(s) + 1))
                hs = list(range(next_hash[0], next_hash[0] + depth))
                next_hash[0] += depth
                cont = rng.random() < 0.3 and stored[w]
                parent = int(rng.choice(stored[w])) if cont else (None if rng.random() < 0.8 else 10**9 + step)   # sometimes an unknown parent
                blocks = [_blk(h, [int(t) for t in s[i]]) for i, h in enumerate(hs)]
                events.append({"stored": {"blocks": blocks, "parent_block_hash": parent}})
                stored[w] += hs
            elif r < 0.9 and stored[w]:
                k = int(rng.integers(1, min(4, len(stored[w])) + 1))
                victims = [int(x) for x in rng.choice(stored[w], size=k, replace=False)]
                events.append({"removed": {"block_hashes": victims}})
                stored[w] = [h for h in stored[w] if h not in victims]
            elif r >= 0.97:
                events.append({"cleared": {}})
                stored[w] = []
        fall_g += mon.apply_events("unknown", w, events)
        fall_o += orc.apply_kv_events(oix, w, events)
        assert ix.current_size() == oix.current_size() and ix.entry_count() == oix.entry_count(), step
    assert fall_g == fall_o and fall_g > 0
    for s in seqs:
        hs = orc.compute_request_content_hashes(s.reshape(-1), 4)
        assert ix.find_matches(hs) == oix.find_matches(hs)
