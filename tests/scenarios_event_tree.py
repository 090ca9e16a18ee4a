"""Known-answer scenarios for the event-driven KV index, ported 1:1 from the reference's unit tests
(crates/kv_index/src/event_tree.rs:801-2020).  Each scenario takes `mk(jump_size)` returning an object with the
PositionalIndexer surface (intern_worker, worker_id, apply_stored, apply_removed, apply_cleared, remove_worker,
current_size, find_matches) and `H`, a namespace with compute_content_hash / compute_next_seq_hash /
compute_request_content_hashes — so the same list runs against the CPU oracle (tests/test_oracle_event_tree.py)
and against the CUDA path through the C-ABI (tests/test_gpu_event_tree.py).
"""
import pytest


def make_blocks(H, content_hashes):
    """event_tree.rs:777-794 — seq_hash = rolling hash of the content hashes."""
    blocks, prev = [], 0
    for i, ch in enumerate(content_hashes):
        seq = ch if i == 0 else H.compute_next_seq_hash(prev, ch)
        prev = seq
        blocks.append((seq, ch))
    return blocks


def store_via_continuations(H, ix, worker, content, chunk):
    """event_tree.rs:1678-1698"""
    wid = ix.intern_worker(worker)
    blocks = make_blocks(H, content)
    off, parent = 0, None
    while off < len(blocks):
        part = blocks[off:off + chunk]
        ix.apply_stored(wid, part, parent)
        parent = part[-1][0]
        off += chunk
    return wid


def s_new_indexer_is_empty(mk, H):  # :801
    ix = mk(32)
    sc, _ = ix.find_matches([1, 2, 3])
    assert sc == {} and ix.current_size() == 0


def s_store_and_find_single_worker(mk, H):  # :810
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    sc, ts = ix.find_matches([10, 20, 30])
    assert sc.get(w1) == 3 and ts.get(w1) == 3


def s_store_partial_prefix_match(mk, H):  # :823
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    sc, _ = ix.find_matches([10, 20, 30, 40, 50])
    assert sc.get(w1) == 3


def s_store_no_match(mk, H):  # :836
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    sc, _ = ix.find_matches([99, 88, 77])
    assert sc == {}


def s_two_workers_different_depths(mk, H):  # :848
    ix = mk(64)
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    ix.apply_stored(w2, make_blocks(H, [10, 20]))
    sc, _ = ix.find_matches([10, 20, 30, 40])
    assert sc.get(w1) == 3 and sc.get(w2) == 2


def s_remove_blocks(mk, H):  # :869
    ix = mk(64)
    blocks = make_blocks(H, [10, 20, 30])
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, blocks)
    ix.apply_removed(w1, [blocks[2][0]])
    sc, ts = ix.find_matches([10, 20, 30])
    assert sc.get(w1) == 2 and ts.get(w1) == 2


def s_clear_worker(mk, H):  # :885
    ix = mk(64)
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    ix.apply_stored(w2, make_blocks(H, [10, 20]))
    ix.apply_cleared(w1)
    sc, _ = ix.find_matches([10, 20, 30])
    assert w1 not in sc and sc.get(w2) == 2


def s_tree_sizes(mk, H):  # :908
    ix = mk(64)
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    ix.apply_stored(w2, make_blocks(H, [10, 20]))
    _, ts = ix.find_matches([10])
    assert ts.get(w1) == 3 and ts.get(w2) == 2


def s_store_with_parent_hash(mk, H):  # :929
    ix = mk(64)
    b1 = make_blocks(H, [10, 20])
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, b1)
    ix.apply_stored(w1, [(300, 30), (400, 40)], parent=b1[1][0])
    sc, ts = ix.find_matches([10, 20, 30, 40])
    assert sc.get(w1) == 4 and ts.get(w1) == 4


def s_store_with_parent_error_worker_not_tracked(mk, H):  # :959
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    with pytest.raises(Exception, match="WorkerNotTracked"):
        ix.apply_stored(w1, make_blocks(H, [10, 20]), parent=999)


def s_store_with_parent_error_parent_not_found(mk, H):  # :969
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20]))
    with pytest.raises(Exception, match="ParentBlockNotFound"):
        ix.apply_stored(w1, make_blocks(H, [30]), parent=999_999)


def s_remove_missing_block_is_noop(mk, H):  # :982
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    ix.apply_removed(w1, [999])
    assert ix.current_size() == 3


def s_remove_unknown_worker_is_noop(mk, H):  # :994
    ix = mk(64)
    w1 = ix.intern_worker("http://unknown:8000")
    ix.apply_removed(w1, [1])


def s_remove_worker(mk, H):  # :1002
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    ix.remove_worker(w1)
    sc, _ = ix.find_matches([10, 20, 30])
    assert sc == {} and ix.current_size() == 0


def s_multiple_workers_same_position(mk, H):  # :1016
    ix = mk(64)
    ws = [ix.intern_worker(f"http://w{i}:8000") for i in (1, 2, 3)]
    for w in ws:
        ix.apply_stored(w, make_blocks(H, [10]))
    sc, _ = ix.find_matches([10])
    assert all(sc.get(w) == 1 for w in ws)


def s_empty_blocks_is_noop(mk, H):  # :1041
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, [])
    assert ix.current_size() == 0


def s_single_block_sequence(mk, H):  # :1050
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [42]))
    sc, _ = ix.find_matches([42])
    assert sc.get(w1) == 1


def s_jump_search_long_prefix(mk, H):  # :1080
    ix = mk(4)
    vals = list(range(1, 21))
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, vals))
    sc, _ = ix.find_matches(vals)
    assert sc.get(w1) == 20


def s_jump_search_worker_drains_mid_jump(mk, H):  # :1093
    ix = mk(4)
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    ix.apply_stored(w1, make_blocks(H, list(range(1, 11))))
    ix.apply_stored(w2, make_blocks(H, list(range(1, 7))))
    sc, _ = ix.find_matches(list(range(1, 11)))
    assert sc.get(w1) == 10 and sc.get(w2) == 6


def s_jump_search_multiple_drains(mk, H):  # :1116
    ix = mk(3)
    w = [ix.intern_worker(f"http://w{i}:8000") for i in (1, 2, 3)]
    for wid, n in zip(w, (12, 7, 4)):
        ix.apply_stored(wid, make_blocks(H, list(range(1, n + 1))))
    sc, _ = ix.find_matches(list(range(1, 13)))
    assert [sc.get(x) for x in w] == [12, 7, 4]


def s_seq_entry_single_to_multi_upgrade(mk, H):  # :1175
    ix = mk(64)
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    ix.apply_stored(w1, [(100, 10)])
    ix.apply_stored(w2, [(200, 10)])
    sc, _ = ix.find_matches([10])
    assert sc.get(w1) == 1 and sc.get(w2) == 1


def s_seq_entry_distinct_prefix_same_content(mk, H):  # :1209
    ix = mk(64)
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 99]))
    ix.apply_stored(w2, make_blocks(H, [20, 99]))
    sc, _ = ix.find_matches([10, 99])
    assert sc.get(w1) == 2 and w2 not in sc
    sc, _ = ix.find_matches([20, 99])
    assert sc.get(w2) == 2 and w1 not in sc


def s_early_exit_returns_score_one(mk, H):  # :1245
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    sc, ts = ix.find_matches([10, 20, 30], early_exit=True)
    assert sc.get(w1) == 1 and ts.get(w1) == 3


def s_early_exit_no_match(mk, H):  # :1260
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    sc, _ = ix.find_matches([99, 88], early_exit=True)
    assert sc == {}


def s_worker_id(mk, H):  # :1276, :1282
    ix = mk(32)
    assert ix.worker_id("http://unknown:8000") is None
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10]))
    assert ix.worker_id("http://w1:8000") == w1
    assert ix.intern_worker("http://w1:8000") == w1


def s_tree_sizes_after_store_and_remove(mk, H):  # :1297
    ix = mk(64)
    blocks = make_blocks(H, [10, 20, 30, 40, 50])
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, blocks)
    assert ix.current_size() == 5
    ix.apply_removed(w1, [blocks[3][0], blocks[4][0]])
    assert ix.current_size() == 3
    _, ts = ix.find_matches([10, 20, 30])
    assert ts.get(w1) == 3


def s_duplicate_store_does_not_inflate_tree_size(mk, H):  # :1315
    ix = mk(64)
    blocks = make_blocks(H, [10, 20, 30])
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, blocks)
    ix.apply_stored(w1, blocks)
    sc, ts = ix.find_matches([10, 20, 30])
    assert ts.get(w1) == 3 and sc.get(w1) == 3


def s_remove_worker_nonexistent_is_noop(mk, H):  # :1340
    ix = mk(32)
    w = ix.intern_worker("http://ghost:8000")
    ix.remove_worker(w)
    assert ix.current_size() == 0


def s_cleanup_no_leak(mk, H):  # :1396
    ix = mk(32)
    blocks = make_blocks(H, [10, 20, 30])
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    ix.apply_stored(w1, blocks)
    ix.apply_stored(w2, blocks)
    assert ix.entry_count() > 0
    ix.remove_worker(w1)
    assert ix.entry_count() > 0
    ix.remove_worker(w2)
    assert ix.entry_count() == 0


def s_seq_hash_rolling_correctness(mk, H):  # :1429
    content = [10, 20, 30, 40, 50]
    blocks = make_blocks(H, content)
    prev = None
    for i, c in enumerate(content):
        prev = c if i == 0 else H.compute_next_seq_hash(prev, c)
        assert prev == blocks[i][0]


def s_query_prefix_of_stored(mk, H):  # :1446
    ix = mk(32)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30, 40, 50]))
    sc, ts = ix.find_matches([10, 20])
    assert sc.get(w1) == 2 and ts.get(w1) == 5


def s_disjoint_workers_no_shared_prefix(mk, H):  # :1459
    ix = mk(32)
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    ix.apply_stored(w2, make_blocks(H, [99, 88, 77]))
    sc, _ = ix.find_matches([10, 20, 30])
    assert sc.get(w1) == 3 and w2 not in sc
    sc, _ = ix.find_matches([99, 88, 77])
    assert sc.get(w2) == 3 and w1 not in sc


def s_current_size_across_operations(mk, H):  # :1490
    ix = mk(32)
    assert ix.current_size() == 0
    blocks = make_blocks(H, [10, 20, 30])
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    ix.apply_stored(w1, blocks)
    assert ix.current_size() == 3
    ix.apply_stored(w2, blocks)
    assert ix.current_size() == 6
    ix.apply_removed(w1, [blocks[2][0]])
    assert ix.current_size() == 5
    ix.apply_cleared(w2)
    assert ix.current_size() == 2
    ix.remove_worker(w1)
    assert ix.current_size() == 0


def s_end_to_end_store_and_query(mk, H):  # :1575
    ix = mk(32)
    tokens = list(range(1, 17))
    chs = [H.compute_content_hash(tokens[i:i + 4]) for i in range(0, 16, 4)]
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, [(0xBEEF0000 + i, ch) for i, ch in enumerate(chs)])
    sc, _ = ix.find_matches(H.compute_request_content_hashes(tokens, 4))
    assert sc.get(w1) == 4


def s_end_to_end_partial_overlap(mk, H):  # :1604
    ix = mk(32)
    cached = list(range(1, 9))
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, [(i + 1, H.compute_content_hash(cached[4 * i:4 * i + 4])) for i in range(2)])
    sc, ts = ix.find_matches(H.compute_request_content_hashes(list(range(1, 17)), 4))
    assert sc.get(w1) == 2 and ts.get(w1) == 2


def s_end_to_end_different_backends_same_content(mk, H):  # :1629
    ix = mk(4)
    tokens = list(range(1, 9))
    chs = [H.compute_content_hash(tokens[i:i + 4]) for i in range(0, 8, 4)]
    sg, vl = ix.intern_worker("http://sglang:8000"), ix.intern_worker("http://vllm:8000")
    ix.apply_stored(sg, [(0xAAAA0000 + i, ch) for i, ch in enumerate(chs)])
    ix.apply_stored(vl, [(0xBBBB0000 + i, ch) for i, ch in enumerate(chs)])
    sc, _ = ix.find_matches(H.compute_request_content_hashes(tokens, 4))
    assert sc.get(sg) == 2 and sc.get(vl) == 2


def s_divergence_at_jump_boundaries(mk, H):  # :1701
    ix = mk(32)
    full = list(range(1, 129))
    fid = ix.intern_worker("http://full:8000")
    ix.apply_stored(fid, make_blocks(H, full))
    for depth in (31, 32, 33, 63, 64, 65):
        wid = ix.intern_worker(f"http://depth{depth}:8000")
        ix.apply_stored(wid, make_blocks(H, full[:depth]))
    sc, _ = ix.find_matches(full)
    assert sc.get(fid) == 128
    for depth in (31, 32, 33, 63, 64, 65):
        assert sc.get(ix.worker_id(f"http://depth{depth}:8000")) == depth


def s_exact_jump_size_sequences(mk, H):  # :1741
    ix = mk(32)
    for n in (32, 64, 96):
        content = list(range(1, n + 1))
        wid = ix.intern_worker(f"http://len{n}:8000")
        ix.apply_stored(wid, make_blocks(H, content))
        sc, _ = ix.find_matches(content)
        assert sc.get(wid) == n


def s_off_by_one_jump_boundaries(mk, H):  # :1762
    ix = mk(32)
    full = list(range(1, 129))
    for n in (31, 33, 63, 65, 95, 97):
        wid = ix.intern_worker(f"http://len{n}:8000")
        ix.apply_stored(wid, make_blocks(H, full[:n]))
        sc, _ = ix.find_matches(full[:n])
        assert sc.get(wid) == n


def s_staggered_workers_across_jump_boundaries(mk, H):  # :1784
    ix = mk(32)
    full = list(range(1, 101))
    depths = (10, 20, 35, 64, 100)
    for d in depths:
        ix.apply_stored(ix.intern_worker(f"http://w{d}:8000"), make_blocks(H, full[:d]))
    sc, _ = ix.find_matches(full)
    for d in depths:
        assert sc.get(ix.worker_id(f"http://w{d}:8000")) == d


def s_shared_prefix_diverge_at_jump_boundary(mk, H):  # :1810
    ix = mk(32)
    shared = list(range(1, 41))
    c1 = shared + list(range(1001, 1061))
    c2 = shared + list(range(2001, 2021))
    w1, w2, w3 = (ix.intern_worker(f"http://w{i}:8000") for i in (1, 2, 3))
    ix.apply_stored(w1, make_blocks(H, c1))
    ix.apply_stored(w2, make_blocks(H, c2))
    ix.apply_stored(w3, make_blocks(H, shared))
    sc, _ = ix.find_matches(c1)
    assert sc.get(w1) == 100 and sc.get(w2) == 40 and sc.get(w3) == 40


def s_very_long_sequence(mk, H):  # :1846
    ix = mk(64)
    content = list(range(1, 1001))
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, content))
    assert ix.find_matches(content)[0].get(w1) == 1000
    assert ix.find_matches(content[:500])[0].get(w1) == 500
    assert ix.find_matches(content[:499] + [999999])[0].get(w1) == 499


def s_deep_continuation_chain(mk, H):  # :1871
    ix = mk(64)
    content = list(range(1, 201))
    w1 = store_via_continuations(H, ix, "http://w1:8000", content, 10)
    assert ix.current_size() == 200
    assert ix.find_matches(content)[0].get(w1) == 200
    assert ix.find_matches(content[:150])[0].get(w1) == 150


def s_continuation_chain_with_multiple_workers(mk, H):  # :1888
    ix = mk(32)
    content = list(range(1, 101))
    w1 = store_via_continuations(H, ix, "http://w1:8000", content, 10)
    w2 = store_via_continuations(H, ix, "http://w2:8000", content[:50], 10)
    sc, _ = ix.find_matches(content)
    assert sc.get(w1) == 100 and sc.get(w2) == 50


def s_multiple_disjoint_sequences_per_worker(mk, H):  # :1905
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 20, 30]))
    ix.apply_stored(w1, make_blocks(H, [100, 200, 300, 400]))
    assert ix.find_matches([100, 200, 300, 400])[0].get(w1) == 4
    assert ix.find_matches([10, 20, 30])[0].get(w1) == 3


def s_long_sequence_partial_removal(mk, H):  # :1928
    ix = mk(32)
    content = list(range(1, 101))
    blocks = make_blocks(H, content)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, blocks)
    ix.apply_removed(w1, [b[0] for b in blocks[80:]])
    assert ix.current_size() == 80
    assert ix.find_matches(content)[0].get(w1) == 80
    assert ix.find_matches(content[:80])[0].get(w1) == 80


def s_remove_parent_does_not_cascade(mk, H):  # :1949
    ix = mk(1)
    blocks = make_blocks(H, [10, 20, 30, 40, 50])
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, blocks)
    ix.apply_removed(w1, [blocks[1][0]])
    assert ix.current_size() == 4
    assert ix.find_matches([10, 20, 30, 40, 50])[0].get(w1) == 1


def s_long_sequence_clear_and_rebuild(mk, H):  # :1965
    ix = mk(32)
    w1 = ix.intern_worker("http://w1:8000")
    original = list(range(1, 101))
    ix.apply_stored(w1, make_blocks(H, original))
    ix.apply_cleared(w1)
    assert ix.current_size() == 0
    replacement = list(range(1001, 1101))
    ix.apply_stored(w1, make_blocks(H, replacement))
    assert w1 not in ix.find_matches(original)[0]
    assert ix.find_matches(replacement)[0].get(w1) == 100


def s_interleaved_long_sequences(mk, H):  # :1991
    ix = mk(32)
    content = list(range(1, 101))
    depths = (25, 50, 75, 100)
    for d in depths:
        ix.apply_stored(ix.intern_worker(f"http://w{d}:8000"), make_blocks(H, content[:d]))
    sc, ts = ix.find_matches(content)
    for d in depths:
        wid = ix.worker_id(f"http://w{d}:8000")
        assert sc.get(wid) == d and ts.get(wid) == d


# ---- beyond the reference's tests: non-prefix-closed states that exercise the count-only jump test (:720),
# ---- the retain guard (:611, :641) and the Multi stickiness (:196-212).  Expected values derived by hand from the code.

def s_jump_overcounts_past_gap(mk, H):
    """A mid-sequence removal lands between jump points: count at the jump destination still equals |active|,
    so the reference skips the gap and scores the full length (documented caveat, event_tree.rs:456-460)."""
    ix = mk(4)
    content = list(range(1, 10))  # len 9: jump points 4, 8
    blocks = make_blocks(H, content)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, blocks)
    ix.apply_removed(w1, [blocks[2][0]])  # gap at position 2
    assert ix.find_matches(content)[0].get(w1) == 9


def s_retain_guard_keeps_absent_worker(mk, H):
    """workers.len() >= active.len() skips the retain even when an active worker is absent at that position."""
    ix = mk(1)  # jump 1 → every position is a jump destination and a 1-wide scan on mismatch
    a, b, c = (ix.intern_worker(f"http://{x}:8000") for x in "abc")
    ix.apply_stored(a, make_blocks(H, [10, 20, 30]))
    ix.apply_stored(b, make_blocks(H, [10]))
    cb = make_blocks(H, [10, 20])
    ix.apply_stored(c, cb)
    ix.apply_removed(c, [cb[0][0]])   # no cascade: c keeps (1,20) with the same prefix hash → still a Single {a,c}
    # query [10,20,30]: active {a,b}; pos1 set {a,c}: count 2 == |active| → skipped; pos2 {a}: count 1 != 2 →
    # scan pos 2: |{a}| < 2 → b drained at 2.  b never owned position 1, yet scores 2; c is never scored.
    sc, _ = ix.find_matches([10, 20, 30])
    assert sc.get(a) == 3 and sc.get(b) == 2 and c not in sc


def s_multi_entry_never_downgrades(mk, H):
    """After Single→Multi, removing one prefix leaves a one-key Multi that still requires the rolling prefix hash."""
    ix = mk(64)
    w1, w2 = ix.intern_worker("http://w1:8000"), ix.intern_worker("http://w2:8000")
    b1 = make_blocks(H, [10, 99])
    b2 = make_blocks(H, [20, 99])
    ix.apply_stored(w1, b1)
    ix.apply_stored(w2, b2)               # (1,99) is Multi now
    ix.apply_removed(w2, [b2[1][0]])      # Multi with the single key prefix(10,99)
    # a Single entry would match [20,99] at position 1 without checking the prefix; the sticky Multi must not.
    sc, _ = ix.find_matches([20, 99])
    assert sc.get(w2) == 1 and w1 not in sc
    sc, _ = ix.find_matches([10, 99])
    assert sc.get(w1) == 2


def s_single_entry_ignores_prefix_hash(mk, H):
    """A Single entry matches on (position, content) alone (:229-235): a different history still scores."""
    ix = mk(64)
    w1 = ix.intern_worker("http://w1:8000")
    ix.apply_stored(w1, make_blocks(H, [10, 99]))
    w2 = ix.intern_worker("http://w2:8000")
    ix.apply_stored(w2, make_blocks(H, [20]))
    # query [20, 99]: position 0 → {w2}; position 1 (jump dest) is Single {w1}: count 1 == |active| → skip → w2 scores 2
    sc, _ = ix.find_matches([20, 99])
    assert sc.get(w2) == 2 and w1 not in sc


SCENARIOS = [v for k, v in sorted(globals().items()) if k.startswith("s_") and callable(v)]
