"""GPU parity for the pick itself: CacheAwarePolicy::select_worker in event-driven mode, through the C ABI.
Ports the reference's known-answer tests (cache_aware.rs:1433-1957, policies/mod.rs:192-262) and adds randomized
bit-exact parity against the oracle at fleet sizes 64 / 256 / 300, ragged lengths, every block size the reference's
benches use, unhealthy / open-circuit workers and the f32 imbalance gate."""
import numpy as np
import pytest

from oracle import orc
from smg_b200 import synth

pytestmark = pytest.mark.gpu


def _policy(**cfg):
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy
    c = CacheAwareConfig(eviction_interval_secs=0, **cfg)
    return CacheAwarePolicy(c)


def _workers(urls):
    from smg_b200 import BasicWorker
    return [BasicWorker(u) for u in urls]


def _store(ix, url, chunks, jump=None):
    wid = ix.intern_worker(url)
    ix.apply_stored(wid, [(i + 1, orc.compute_content_hash(c)) for i, c in enumerate(chunks)])
    return wid


def _setup(urls, block_size=4, jump=4):
    pol = _policy(block_size=block_size)
    ws = _workers(urls)
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(block_size)
    ix = mon.create_indexer("unknown", jump)
    pol.set_kv_event_monitor(mon)
    return pol, ws, mon, ix


def _sel(pol, ws, tokens):
    from smg_b200 import SelectWorkerInfo
    return pol.select_worker(ws, SelectWorkerInfo(tokens=tokens))


def test_name_and_flags():
    pol = _policy()
    assert pol.name() == "cache_aware" and pol.needs_request_text()


def test_score_overlap_selects_best_match():  # cache_aware.rs:1433
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"])
    _store(ix, "http://w1:8000", [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [13, 14, 15, 16]])
    assert _sel(pol, ws, list(range(1, 17))) == 0


def test_score_overlap_load_tiebreak():  # :1500
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"])
    for _ in range(10):
        ws[0].increment_load()
    for u in ("http://w1:8000", "http://w2:8000"):
        ix.apply_stored(ix.intern_worker(u), [(1, orc.compute_content_hash([1, 2, 3, 4]))])
    assert _sel(pol, ws, [1, 2, 3, 4]) == 1


def test_score_overlap_tree_size_tiebreak():  # :1547
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"])
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    blk = [(1, orc.compute_content_hash([1, 2, 3, 4]))]
    ix.apply_stored(w1, blk)
    ix.apply_stored(w2, blk)
    ix.apply_stored(w2, [(2, orc.compute_content_hash([5, 6, 7, 8]))], parent=1)
    assert _sel(pol, ws, [1, 2, 3, 4]) == 0


def test_full_tie_takes_last_max():
    """max_by_key returns the LAST maximum: equal score, load and tree size → the highest slice index."""
    pol, ws, mon, ix = _setup(["http://a", "http://b", "http://c"])
    blk = [(1, orc.compute_content_hash([1, 2, 3, 4]))]
    for u in ("http://a", "http://b", "http://c"):
        ix.apply_stored(ix.intern_worker(u), blk)
    assert _sel(pol, ws, [1, 2, 3, 4]) == 2


def test_score_overlap_partial_match():  # :1616
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"])
    chunks = [[4 * i + 1, 4 * i + 2, 4 * i + 3, 4 * i + 4] for i in range(4)]
    _store(ix, "http://w1:8000", chunks)
    _store(ix, "http://w2:8000", chunks[:2])
    assert _sel(pol, ws, list(range(1, 17))) == 0


def test_event_driven_overlap_selects_cached_worker():  # :1686
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"])
    _store(ix, "http://w1:8000", [[1, 2, 3, 4], [5, 6, 7, 8]])
    assert _sel(pol, ws, [1, 2, 3, 4, 5, 6, 7, 8]) == 0


def test_event_driven_no_overlap_uses_min_load():  # :1725
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"])
    for _ in range(3):
        ws[0].increment_load()
    _store(ix, "http://w1:8000", [[1, 2, 3, 4]])
    assert _sel(pol, ws, [100, 200, 300, 400]) == 1


def test_event_driven_short_request_uses_min_load():  # :1764
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"])
    for _ in range(3):
        ws[0].increment_load()
    _store(ix, "http://w1:8000", [[1, 2, 3, 4]])
    assert _sel(pol, ws, [1, 2, 3]) == 1


def test_event_driven_uses_monitor_block_size():  # :1856
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"], block_size=4)
    ix.apply_stored(ix.intern_worker("http://w1:8000"), [(1, orc.compute_content_hash([1, 2, 3, 4, 5, 6, 7, 8]))])
    mon.set_block_size("unknown", 8)
    assert _sel(pol, ws, [1, 2, 3, 4, 5, 6, 7, 8]) == 0


def test_imbalanced_skips_event_driven():  # :1916 (abs 5 / rel 2.0, loads 20 vs 0 → idx 1)
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy
    pol = CacheAwarePolicy(CacheAwareConfig(balance_abs_threshold=5, balance_rel_threshold=2.0, eviction_interval_secs=0, block_size=4))
    ws = _workers(["http://w1:8000", "http://w2:8000"])
    for _ in range(20):
        ws[0].increment_load()
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(4)
    ix = mon.create_indexer("unknown", 4)
    _store(ix, "http://w1:8000", [[1, 2, 3, 4]])
    pol.set_kv_event_monitor(mon)
    idx, info = pol.select_worker_batch(ws, [[1, 2, 3, 4]])
    assert idx[0] == 1 and info[0].branch == 1


def test_unhealthy_and_open_circuit_workers_are_skipped():  # policies/mod.rs:192-262
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000", "http://w3:8000"])
    chunks = [[1, 2, 3, 4], [5, 6, 7, 8]]
    for u in ("http://w1:8000", "http://w2:8000", "http://w3:8000"):
        _store(ix, u, chunks)
    ws[2].set_healthy(False)          # would win the last-max tie-break
    assert _sel(pol, ws, [1, 2, 3, 4, 5, 6, 7, 8]) == 1
    ws[1].set_circuit_ok(False)
    assert _sel(pol, ws, [1, 2, 3, 4, 5, 6, 7, 8]) == 0
    ws[0].set_healthy(False)
    assert _sel(pol, ws, [1, 2, 3, 4, 5, 6, 7, 8]) is None   # no healthy worker → None (cache_aware.rs:653-655)


def test_worker_not_in_slice_is_ignored():
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"])
    _store(ix, "http://ghost:8000", [[1, 2, 3, 4]])       # cached only on a worker the router did not pass in
    for _ in range(2):
        ws[0].increment_load()
    idx, info = pol.select_worker_batch(ws, [[1, 2, 3, 4]])
    assert idx[0] == 1 and info[0].branch == 3            # no eligible overlap → min-load fallback


def _oracle_policy(urls, cfg, jump, bs):
    op = orc.CacheAwarePolicy(eviction_interval_secs=0, **cfg)
    op.set_workers(urls)
    oix = orc.PositionalIndexer(jump)
    op.attach_indexer("unknown", oix)
    op.set_block_size("unknown", bs)
    op.set_kv_event_monitor(True)
    return op, oix


@pytest.mark.parametrize("seed,n_workers,T,bs,jump,B", [
    (1, 64, 512, 16, 64, 512),     # config-2 shape, scaled down
    (2, 64, 512, 16, 8, 256),      # CI bench jump size
    (3, 256, 1024, 16, 64, 256),   # 4 bitset words
    (4, 300, 256, 16, 32, 256),    # 8 words, non power of two fleet
    (5, 64, 512, 64, 64, 256),     # block size 64 (XXH3 long-input path)
    (6, 10, 96, 4, 4, 300),        # tiny blocks / tiny jump
    (7, 100, 2048, 32, 32, 128),   # > 32 blocks per request
    (8, 64, 8192, 16, 64, 48),     # BASELINE sweep upper bound
])
def test_random_select_parity(seed, n_workers, T, bs, jump, B):
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy
    rng = np.random.default_rng(seed)
    cfg = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=bs)   # CLI defaults (main.rs:156-165)
    urls = synth.worker_urls(n_workers)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **cfg))
    ws = _workers(urls)
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(bs)
    ix = mon.create_indexer("unknown", jump)
    pol.set_kv_event_monitor(mon)
    op, oix = _oracle_policy(urls, cfg, jump, bs)

    n_seq = 200
    seqs = synth.gen_sequences(n_seq, T, seed)
    # several workers per sequence with different depths → multi-worker sets, drains inside and across jumps
    order = rng.permutation(n_workers)
    for u in (urls[i] for i in order):           # intern in a shuffled order: worker id ≠ slice index
        assert ix.intern_worker(u) == oix.intern_worker(u)
    P = T // bs
    seq_id = 1
    for s in range(n_seq):
        hashes = orc.compute_request_content_hashes(seqs[s], bs)
        for w in rng.choice(n_workers, size=rng.integers(1, 5), replace=False):
            depth = int(rng.integers(1, P + 1))
            blocks = [(seq_id + i, hashes[i]) for i in range(depth)]
            seq_id += depth
            wid = ix.worker_id(urls[w])
            ix.apply_stored(wid, blocks)
            oix.apply_stored(wid, blocks)
            if rng.random() < 0.15 and depth > 2:     # mid-sequence removal → non-prefix-closed state
                victim = [blocks[int(rng.integers(0, depth))][0]]
                ix.apply_removed(wid, victim)
                oix.apply_removed(wid, victim)
    loads = synth.poisson_loads(n_workers, 8, seed)
    healthy = (rng.random(n_workers) > 0.1).astype(np.uint8)
    circuit = (rng.random(n_workers) > 0.05).astype(np.uint8)
    for i, w in enumerate(ws):
        w.set_load(int(loads[i])); w.set_healthy(bool(healthy[i])); w.set_circuit_ok(bool(circuit[i]))
    op.set_state(loads, healthy, circuit)

    q = synth.gen_queries(seqs, B, seed, block=bs)
    # ragged: truncate each request to a random length (incl. < one block and non-multiples of the block size)
    lens = rng.integers(0, T + 1, size=B)
    lens[: B // 4] = T
    reqs = [q[i, : lens[i]] for i in range(B)]
    idx, info = pol.select_worker_batch(ws, reqs)
    flat = np.concatenate(reqs).astype(np.uint32) if sum(lens) else np.zeros(0, np.uint32)
    offs = np.zeros(B + 1, np.uint64); np.cumsum(lens, out=offs[1:])
    oidx, obr, oma, _ = op.select_batch_tokens(flat, offs)
    assert np.array_equal(idx, oidx)
    assert [i.branch for i in info] == list(obr)
    assert [i.matched * bs for i in info] == list(oma)
    assert (np.asarray(obr) == 2).sum() > B // 10     # the test really exercises overlap picks …
    assert (np.asarray(obr) == 3).sum() > 0           # … and min-load fallbacks


def test_imbalance_gate_f32_boundary():
    """(max as f32) > (min as f32 * rel) is evaluated in f32 exactly like the reference (cache_aware.rs:669-670)."""
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy
    urls = ["http://a", "http://b"]
    for mn, mx, rel, abs_thr in [(10, 15, 1.5, 4), (10, 16, 1.5, 4), (16777217, 25165826, 1.5, 4), (3, 5, 1.6666666, 1),
                                 (3, 5, 1.6666667, 1), (0, 65, 1.5, 64), (0, 64, 1.5, 64), (7, 8, 1.1, 0), (1000, 1101, 1.1, 100)]:
        cfg = dict(cache_threshold=0.3, balance_abs_threshold=abs_thr, balance_rel_threshold=rel, block_size=4)
        pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **cfg))
        ws = _workers(urls)
        ws[0].set_load(mx); ws[1].set_load(mn)
        pol.init_workers(ws)
        mon = pol.kv_event_monitor(4)
        ix = mon.create_indexer("unknown", 4)
        _store(ix, "http://a", [[1, 2, 3, 4]])
        pol.set_kv_event_monitor(mon)
        op, oix = _oracle_policy(urls, cfg, 4, 4)
        oix.apply_stored(oix.intern_worker("http://a"), [(1, orc.compute_content_hash([1, 2, 3, 4]))])
        op.set_state([mx, mn], [1, 1], [1, 1])
        idx, info = pol.select_worker_batch(ws, [[1, 2, 3, 4]])
        d = op.select_worker(tokens=[1, 2, 3, 4])
        assert idx[0] == d.idx and orc.BRANCHES[info[0].branch] == d.branch, (mn, mx, rel, abs_thr)


def test_index_updates_between_batches_are_visible():
    """Writes after a query batch reach the device mirror (dirty-slot scatter) before the next batch."""
    pol, ws, mon, ix = _setup(["http://w1:8000", "http://w2:8000"], block_size=4, jump=4)
    ws[0].set_load(5)
    w2 = _store(ix, "http://w2:8000", [[9, 9, 9, 9]])
    assert _sel(pol, ws, [1, 2, 3, 4]) == 1                 # no overlap → min load (w2)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, [(1, orc.compute_content_hash([1, 2, 3, 4]))])
    assert _sel(pol, ws, [1, 2, 3, 4]) == 0                 # now cached on w1
    ix.apply_removed(w1, [1])
    assert _sel(pol, ws, [1, 2, 3, 4]) == 1
    ix.apply_cleared(w2)
    idx, info = pol.select_worker_batch(ws, [[1, 2, 3, 4]])   # empty indexer → falls through to the token tree (:723-729)
    assert idx[0] == 1 and orc.BRANCHES[info[0].branch] == "tree_min_load"


def test_multi_batch_device_path_matches_oracle():
    """smgx_select_many_tokens_device (several batches per launch, blockIdx.y = batch) — same picks as the oracle."""
    import ctypes as C
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy, _lib
    W, T, bs, B, NB = 64, 512, 16, 300, 37          # 37 batches → two launches (32 + 5), ragged last sizes
    cfg = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=bs)
    urls = synth.worker_urls(W)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **cfg), max_tokens_per_request=T)
    ws = _workers(urls)
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(bs)
    ix = mon.create_indexer("unknown", 64)
    pol.set_kv_event_monitor(mon)
    op, oix = _oracle_policy(urls, cfg, 64, bs)
    seqs = synth.gen_sequences(300, T, 9)
    for u in urls:
        assert ix.intern_worker(u) == oix.intern_worker(u)
    for s_ in range(len(seqs)):
        hs = orc.compute_request_content_hashes(seqs[s_], bs)
        blocks = [(s_ * 100 + i + 1, h) for i, h in enumerate(hs)]
        ix.apply_stored(s_ % W, blocks)
        oix.apply_stored(s_ % W, blocks)
    loads = synth.poisson_loads(W, 8, 9)
    for w, l in zip(ws, loads):
        w.set_load(int(l))
    op.set_state(loads, [1] * W, [1] * W)
    model = pol._push_fleet(ws)
    h, L = pol._h, _lib.load()
    err = _lib.new_err()
    d_tok, d_off, d_out, ns, host = [], [], [], [], []
    for j in range(NB):
        n = B - j                                     # different n per batch
        tokens, offsets = synth.ragged(synth.gen_queries(seqs, n, 100 + j))
        host.append((tokens, offsets))
        dt = L.smgx_device_alloc(h.p, tokens.nbytes, C.byref(err)); h.call("smgx_memcpy_h2d", dt, tokens.ctypes.data_as(C.c_void_p), tokens.nbytes)
        do = L.smgx_device_alloc(h.p, offsets.nbytes, C.byref(err)); h.call("smgx_memcpy_h2d", do, offsets.ctypes.data_as(C.c_void_p), offsets.nbytes)
        d_tok.append(dt); d_off.append(do); d_out.append(L.smgx_device_alloc(h.p, n * 4, C.byref(err))); ns.append(n)
    TOK = (C.c_void_p * NB)(*d_tok); OFF = (C.c_void_p * NB)(*d_off); OUT = (C.c_void_p * NB)(*d_out); NS = (C.c_uint32 * NB)(*ns)
    h.call("smgx_select_many_tokens_device", model, NB, TOK, OFF, NS, T, OUT)
    h.call("smgx_synchronize")
    for j in range(NB):
        got = np.zeros(ns[j], np.int32)
        h.call("smgx_memcpy_d2h", got.ctypes.data_as(C.c_void_p), d_out[j], ns[j] * 4)
        want, _, _, _ = op.select_batch_tokens(host[j][0], host[j][1].astype(np.uint64))
        assert np.array_equal(got, want), j


def test_concurrent_callers_share_the_pipeline():
    """select_worker is called from many tokio tasks at once: 8 threads hammer the synchronous batch call on one policy —
    more threads than stream lanes — and every batch must equal the single-threaded answer."""
    import threading
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy
    cfg = dict(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5, block_size=16)
    urls = synth.worker_urls(16)
    pol = CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0, **cfg))
    ws = _workers(urls)
    pol.init_workers(ws)
    mon = pol.kv_event_monitor(16)
    ix = mon.create_indexer("unknown", 8)
    pol.set_kv_event_monitor(mon)
    seqs = synth.gen_sequences(64, 256, 3)
    for u in urls:
        ix.intern_worker(u)
    sid = 1
    for s in range(64):
        hs = orc.compute_request_content_hashes(seqs[s], 16)
        ix.apply_stored(s % 16, [(sid + i, hs[i]) for i in range(len(hs))])
        sid += len(hs)
    batches = []
    for k in range(8):
        q = synth.gen_queries(seqs, 200, 100 + k, block=16)
        tokens, offsets = synth.ragged(q)
        batches.append((tokens, offsets))
    model = pol._push_fleet(ws)
    import ctypes as C

    def run(tokens, offsets):
        out = np.full(200, -9, np.int32)
        pol._h.call("smgx_select_batch_tokens", model, tokens.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), 200,
                    out.ctypes.data_as(C.c_void_p), None)
        return out

    want = [run(t, o) for t, o in batches]
    errors = []

    def worker(k):
        try:
            for _ in range(30):
                if not np.array_equal(run(*batches[k]), want[k]):
                    errors.append(k)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
