"""Policy-level known answers ported from the reference (model_gateway/src/policies/cache_aware.rs:999-2015,
policies/mod.rs:192-262).  `mk(config)` returns a policy object with the CacheAwarePolicy surface; workers are
smg_b200.BasicWorker.  TEXT_ / TOKEN_ / EVENT_ prefixes name the cache-aware mode a scenario exercises so the GPU
suite can run the modes that exist in the CUDA build."""
from oracle import orc
from smg_b200.policy import BasicWorker, CacheAwareConfig, SelectWorkerInfo


def W(*urls):
    return [BasicWorker(u) for u in urls]


def cfg(**kw):
    kw.setdefault("eviction_interval_secs", 0)
    return CacheAwareConfig(**kw)


def TEXT_balanced_load(mk):  # cache_aware.rs:999
    pol = mk(cfg())
    ws = W("http://w1:8000", "http://w2:8000")
    pol.init_workers(ws)
    i1 = pol.select_worker(ws, SelectWorkerInfo(request_text="hello world"))
    i2 = pol.select_worker(ws, SelectWorkerInfo(request_text="hello world"))
    i3 = pol.select_worker(ws, SelectWorkerInfo(request_text="hello"))
    assert i1 is not None and i1 == i2 == i3


def TEXT_imbalanced_load(mk):  # :1063
    pol = mk(cfg(cache_threshold=0.5, balance_abs_threshold=5, balance_rel_threshold=2.0, max_tree_size=10000, block_size=16))
    ws = W("http://w1:8000", "http://w2:8000")
    for _ in range(20):
        ws[0].increment_load()
    pol.init_workers(ws)
    for _ in range(5):
        assert pol.select_worker(ws, SelectWorkerInfo(request_text="test")) == 1


def TEXT_worker_removal(mk):  # :1103
    pol = mk(cfg())
    ws = W("http://w1:8000", "http://w2:8000")
    pol.init_workers(ws)
    pol.select_worker(ws, SelectWorkerInfo(request_text="test1"))
    pol.select_worker(ws, SelectWorkerInfo(request_text="test2"))
    pol.remove_worker_by_url("http://w1:8000")
    ws[0].set_healthy(False)
    assert pol.select_worker(ws, SelectWorkerInfo(request_text="test1")) == 1


def TEXT_single_worker(mk):  # :1364
    pol = mk(cfg())
    ws = W("http://w1:8000")
    pol.init_workers(ws)
    assert pol.select_worker(ws, SelectWorkerInfo(request_text="test request")) == 0


def TEXT_threshold_and_min_load(mk):
    """match_rate > cache_threshold routes to the tenant, otherwise first-min-load (cache_aware.rs:921-938); rate in f32."""
    pol = mk(cfg(cache_threshold=0.5))
    ws = W("http://a", "http://b", "http://c")
    ws[0].set_load(3); ws[1].set_load(1); ws[2].set_load(1)
    pol.init_workers(ws)
    assert pol.select_worker(ws, SelectWorkerInfo(request_text="abcdefghij")) == 1      # empty tree → rate 0 → first min load (b)
    ws[1].set_load(9)
    assert pol.select_worker(ws, SelectWorkerInfo(request_text="abcdefghijXX")) == 1    # 10/12 > 0.5 → tenant b despite its load
    assert pol.select_worker(ws, SelectWorkerInfo(request_text="abcdeZZZZZZ")) == 2     # 5/11 ≤ 0.5 → min load → c (first min among a=3,b=9,c=1)
    assert pol.select_worker(ws, SelectWorkerInfo(request_text="abcdeZZZZZ")) == 2      # now "abcdeZZZZZ" is cached on c: 10/10


def TEXT_stale_tenant_falls_back_to_first_healthy(mk):  # :940-964
    pol = mk(cfg(cache_threshold=0.5))
    ws = W("http://a", "http://b", "http://c")
    ws[0].set_load(5); ws[1].set_load(5)
    pol.init_workers(ws)
    assert pol.select_worker(ws, SelectWorkerInfo(request_text="shared prefix text")) == 2
    ws[2].set_healthy(False)                       # the cached tenant is now unhealthy → healthy_indices.first()
    assert pol.select_worker(ws, SelectWorkerInfo(request_text="shared prefix text")) == 0
    assert pol.select_worker(ws[:2], SelectWorkerInfo(request_text="shared prefix text")) == 0   # tenant URL not in the slice


def TEXT_no_healthy_returns_none(mk):  # :653-655
    pol = mk(cfg())
    ws = W("http://a", "http://b")
    pol.init_workers(ws)
    for w in ws:
        w.set_healthy(False)
    assert pol.select_worker(ws, SelectWorkerInfo(request_text="x")) is None


def TOKEN_no_monitor_uses_token_tree(mk):  # :1801
    pol = mk(cfg(block_size=4))
    ws = W("http://w1:8000", "http://w2:8000")
    pol.init_workers(ws)
    idx = pol.select_worker(ws, SelectWorkerInfo(tokens=[1, 2, 3, 4]))
    assert idx in (0, 1)


def TOKEN_empty_indexer_falls_through_to_token_tree(mk):  # :1960
    pol = mk(cfg(block_size=4))
    ws = W("http://w1:8000", "http://w2:8000")
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(4)
    mon.create_indexer("unknown", 4)
    pol.set_kv_event_monitor(mon)
    toks = list(range(1, 33))
    i1 = pol.select_worker(ws, SelectWorkerInfo(tokens=toks))
    i2 = pol.select_worker(ws, SelectWorkerInfo(tokens=toks))
    assert i1 in (0, 1) and i1 == i2


def TOKEN_tree_affinity_and_threshold(mk):
    """select_worker_with_tokens (:834-904): rate = matched / UNALIGNED input length in f32."""
    pol = mk(cfg(cache_threshold=0.5))
    ws = W("http://a", "http://b")
    ws[0].set_load(2)
    pol.init_workers(ws)
    seq = list(range(100, 164))                                  # 4 pages
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=seq)) == 1        # empty tree → min load b
    ws[1].set_load(7)
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=seq + [1] * 15)) == 1   # 64/79 > 0.5 → b
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=seq[:32] + [9] * 40)) == 0   # 32/72 ≤ 0.5 → min load a
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=seq[:32] + [9] * 40)) == 0   # now cached on a (64/72)
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=[5, 6, 7])) == 0              # < one page: rate 0 → min load


def TOKEN_imbalanced_updates_tree(mk):
    """Imbalanced picks still match + insert into the token tree (cache_aware.rs:380-402)."""
    pol = mk(cfg(cache_threshold=0.5, balance_abs_threshold=5, balance_rel_threshold=2.0))
    ws = W("http://a", "http://b")
    for _ in range(20):
        ws[0].increment_load()
    pol.init_workers(ws)
    seq = list(range(1, 33))
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=seq)) == 1        # imbalanced → min load b, inserted for b
    ws[0].set_load(0); ws[1].set_load(3)                                    # balanced again, a is the min-load worker
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=seq)) == 1        # tree remembers b (32/32 > 0.5)


def _store(ix, url, chunks):
    wid = ix.intern_worker(url)
    ix.apply_stored(wid, [(i + 1, orc.compute_content_hash(c)) for i, c in enumerate(chunks)])
    return wid


def _event(mk, urls, block_size=4, jump=4, **kw):
    pol = mk(cfg(block_size=block_size, **kw))
    ws = W(*urls)
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(block_size)
    ix = mon.create_indexer("unknown", jump)
    pol.set_kv_event_monitor(mon)
    return pol, ws, mon, ix


def EVENT_overlap_selects_cached_worker(mk):  # :1686
    pol, ws, mon, ix = _event(mk, ["http://w1:8000", "http://w2:8000"])
    _store(ix, "http://w1:8000", [[1, 2, 3, 4], [5, 6, 7, 8]])
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=[1, 2, 3, 4, 5, 6, 7, 8])) == 0


def EVENT_no_overlap_uses_min_load(mk):  # :1725
    pol, ws, mon, ix = _event(mk, ["http://w1:8000", "http://w2:8000"])
    for _ in range(3):
        ws[0].increment_load()
    _store(ix, "http://w1:8000", [[1, 2, 3, 4]])
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=[100, 200, 300, 400])) == 1


def EVENT_short_request_uses_min_load(mk):  # :1764
    pol, ws, mon, ix = _event(mk, ["http://w1:8000", "http://w2:8000"])
    for _ in range(3):
        ws[0].increment_load()
    _store(ix, "http://w1:8000", [[1, 2, 3, 4]])
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=[1, 2, 3])) == 1


def EVENT_uses_monitor_block_size(mk):  # :1856
    pol, ws, mon, ix = _event(mk, ["http://w1:8000", "http://w2:8000"])
    ix.apply_stored(ix.intern_worker("http://w1:8000"), [(1, orc.compute_content_hash([1, 2, 3, 4, 5, 6, 7, 8]))])
    mon.set_block_size("unknown", 8)
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=[1, 2, 3, 4, 5, 6, 7, 8])) == 0


def EVENT_imbalanced_skips_event_driven(mk):  # :1916
    pol, ws, mon, ix = _event(mk, ["http://w1:8000", "http://w2:8000"], balance_abs_threshold=5, balance_rel_threshold=2.0)
    for _ in range(20):
        ws[0].increment_load()
    _store(ix, "http://w1:8000", [[1, 2, 3, 4]])
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=[1, 2, 3, 4])) == 1


def EVENT_tiebreaks(mk):  # :1500, :1547 + LAST max on full ties
    pol, ws, mon, ix = _event(mk, ["http://a", "http://b", "http://c"])
    blk = [(1, orc.compute_content_hash([1, 2, 3, 4]))]
    for u in ("http://a", "http://b", "http://c"):
        ix.apply_stored(ix.intern_worker(u), blk)
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=[1, 2, 3, 4])) == 2          # full tie → last max
    ws[2].set_load(4)
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=[1, 2, 3, 4])) == 1          # lower load wins
    ix.apply_stored(ix.worker_id("http://b"), [(2, orc.compute_content_hash([5, 6, 7, 8]))], parent=1)
    assert pol.select_worker(ws, SelectWorkerInfo(tokens=[1, 2, 3, 4])) == 0          # equal load: smaller tree wins


ALL = {k: v for k, v in sorted(globals().items()) if k.split("_")[0] in ("TEXT", "TOKEN", "EVENT") and callable(v)}
