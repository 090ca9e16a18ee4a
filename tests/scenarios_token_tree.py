"""Known-answer scenarios for the page-aligned multi-tenant token radix tree, ported from the reference's unit tests
(crates/kv_index/src/token_tree.rs:1192-2697; thread-only tests omitted).  `mk(policy)` returns an object with the
TokenTree surface: insert_tokens, match_prefix_with_counts (→ .tenant, .matched_token_count, .input_token_count),
evict_tenant, evict_tenant_by_size, tenant_token_size, clear, entries.  Runs against the oracle (CPU) and the
CUDA-backed tree (GPU)."""

PAGE = 16
LRU, LFU, FIFO, MRU, FILO, PRIORITY = range(6)


def make_tokens(base, pages):  # token_tree.rs:1188-1190
    return [base + i for i in range(pages * PAGE)]


def s_basic_insert_match(mk):  # :1193
    t = mk(LRU)
    tokens = make_tokens(1, 2)
    t.insert_tokens(tokens, "tenant1")
    r = t.match_prefix_with_counts(tokens)
    assert r.matched_token_count == 32 and r.tenant == "tenant1"
    r = t.match_prefix_with_counts(make_tokens(1, 1))
    assert r.matched_token_count == PAGE and r.tenant == "tenant1"
    r = t.match_prefix_with_counts(tokens + [100, 101, 102, 103, 104])
    assert r.matched_token_count == 32 and r.input_token_count == 37


def s_short_sequences_skipped(mk):  # :1219
    t = mk(LRU)
    t.insert_tokens([1, 2, 3, 4, 5], "tenant1")
    assert t.tenant_token_size("tenant1") == 0
    r = t.match_prefix_with_counts([1, 2, 3, 4, 5])
    assert r.matched_token_count == 0 and r.input_token_count == 5


def s_multiple_tenants(mk):  # :1236
    t = mk(LRU)
    tokens = make_tokens(1, 1)
    t.insert_tokens(tokens, "tenant1")
    t.insert_tokens(tokens, "tenant2")
    r = t.match_prefix_with_counts(tokens)
    assert r.matched_token_count == PAGE and r.tenant in ("tenant1", "tenant2")


def s_prefix_split(mk):  # :1250, :2002, :2074
    t = mk(LRU)
    long_t, short_t = make_tokens(1, 3), make_tokens(1, 1)
    t.insert_tokens(long_t, "tenant1")
    t.insert_tokens(short_t, "tenant2")
    assert t.match_prefix_with_counts(short_t).matched_token_count == PAGE
    r = t.match_prefix_with_counts(long_t)
    assert r.matched_token_count == 3 * PAGE and r.tenant == "tenant1"


def s_empty_input(mk):  # :1272
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 1), "tenant1")
    r = t.match_prefix_with_counts([])
    assert r.matched_token_count == 0 and r.input_token_count == 0


def s_no_match(mk):  # :1284
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 1), "tenant1")
    r = t.match_prefix_with_counts(make_tokens(1000, 1))
    assert r.matched_token_count == 0 and r.tenant == "empty"


def s_eviction_reduces_counts(mk):  # :1297
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 2), "tenant1")
    t.insert_tokens(make_tokens(1, 3), "tenant1")
    before = t.tenant_token_size("tenant1")
    assert before > 0
    t.evict_tenant("tenant1", 0)
    assert t.tenant_token_size("tenant1") < before


def s_prefix_match_with_counts(mk):  # :1351
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 2), "tenant1")
    r = t.match_prefix_with_counts(make_tokens(1, 2))
    assert (r.matched_token_count, r.input_token_count) == (2 * PAGE, 2 * PAGE)
    r = t.match_prefix_with_counts(make_tokens(1, 1))
    assert (r.matched_token_count, r.input_token_count) == (PAGE, PAGE)
    r = t.match_prefix_with_counts(make_tokens(1, 3))
    assert (r.matched_token_count, r.input_token_count) == (2 * PAGE, 3 * PAGE)


def s_disjoint_paths(mk):  # :1376
    t = mk(LRU)
    for base, ten in ((1, "tenant1"), (1000, "tenant2"), (2000, "tenant3")):
        t.insert_tokens(make_tokens(base, 1), ten)
    for base, ten in ((1, "tenant1"), (1000, "tenant2"), (2000, "tenant3")):
        r = t.match_prefix_with_counts(make_tokens(base, 1))
        assert r.matched_token_count == PAGE and r.tenant == ten


def s_branching_paths(mk):  # :1401
    t = mk(LRU)
    seqs = [make_tokens(1, 1) + make_tokens(b, 1) for b in (100, 200, 300)]
    for s, ten in zip(seqs, ("tenant1", "tenant2", "tenant3")):
        t.insert_tokens(s, ten)
    for s, ten in zip(seqs[:2], ("tenant1", "tenant2")):
        r = t.match_prefix_with_counts(s)
        assert r.matched_token_count == 2 * PAGE and r.tenant == ten
    assert t.match_prefix_with_counts(make_tokens(1, 1)).matched_token_count == PAGE


def s_clear(mk):  # :1453
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 1), "tenant1")
    t.insert_tokens(make_tokens(1000, 1), "tenant2")
    assert t.tenant_token_size("tenant1") > 0
    t.clear()
    assert t.tenant_token_size("tenant1") == 0 and t.tenant_token_size("tenant2") == 0
    assert t.match_prefix_with_counts(make_tokens(1, 1)).matched_token_count == 0


def s_tenant_token_count(mk):  # :1471
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 2), "tenant1")
    t.insert_tokens(make_tokens(1, 3), "tenant1")
    t.insert_tokens(make_tokens(1000, 1), "tenant2")
    assert t.tenant_token_size("tenant1") >= PAGE and t.tenant_token_size("tenant2") >= PAGE


def s_cold_start(mk):  # :1493
    t = mk(LRU)
    r = t.match_prefix_with_counts([1, 2, 3, 4, 5])
    assert (r.matched_token_count, r.input_token_count) == (0, 5)
    r = t.match_prefix_with_counts(make_tokens(1, 1))
    assert (r.matched_token_count, r.input_token_count) == (0, PAGE)


def s_exact_match_seq(mk):  # :1508
    t = mk(LRU)
    for i in range(100):
        t.insert_tokens(make_tokens(i * 1000, 2), f"tenant{i}")
    for i in range(100):
        r = t.match_prefix_with_counts(make_tokens(i * 1000, 2))
        assert r.matched_token_count == 2 * PAGE and r.tenant == f"tenant{i}"


def s_existing_is_prefix_of_new(mk):  # :2024
    t = mk(LRU)
    short_t, long_t = make_tokens(1, 1), make_tokens(1, 3)
    t.insert_tokens(short_t, "tenant1")
    t.insert_tokens(long_t, "tenant2")
    assert t.match_prefix_with_counts(short_t).matched_token_count == PAGE
    r = t.match_prefix_with_counts(long_t)
    assert r.matched_token_count == 3 * PAGE and r.tenant == "tenant2"


def s_counts_accuracy(mk):  # :2046
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 4), "tenant1")
    for pages, want in ((4, 4), (2, 2), (6, 4)):
        r = t.match_prefix_with_counts(make_tokens(1, pages))
        assert (r.matched_token_count, r.input_token_count) == (want * PAGE, pages * PAGE)


def s_multiple_splits_same_path(mk):  # :2094
    t = mk(LRU)
    for pages, ten in ((4, "tenant1"), (3, "tenant2"), (2, "tenant3"), (1, "tenant4")):
        t.insert_tokens(make_tokens(1, pages), ten)
    for pages in (1, 2, 3, 4):
        assert t.match_prefix_with_counts(make_tokens(1, pages)).matched_token_count == pages * PAGE


def s_very_long_sequences(mk):  # :2330
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 64), "tenant1")
    r = t.match_prefix_with_counts(make_tokens(1, 64))
    assert r.matched_token_count == 64 * PAGE and r.tenant == "tenant1"
    assert t.match_prefix_with_counts(make_tokens(1, 32)).matched_token_count == 32 * PAGE


def s_many_tenants_same_path(mk):  # :2348
    t = mk(LRU)
    tokens = make_tokens(1, 2)
    for i in range(100):
        t.insert_tokens(tokens, f"tenant{i}")
    assert t.match_prefix_with_counts(tokens).matched_token_count == 2 * PAGE
    assert t.tenant_token_size("tenant7") > 0


def s_token_id_edge_values(mk):  # :2367
    t = mk(LRU)
    zeros = list(range(PAGE)); zeros[0] = 0
    maxp = list(range(PAGE)); maxp[0] = 0xFFFFFFFF
    mixed = list(range(PAGE)); mixed[0] = 0; mixed[1] = 0xFFFFFFFF
    t.insert_tokens(zeros, "tenant1")
    t.insert_tokens(maxp, "tenant2")
    t.insert_tokens(mixed, "tenant3")
    r = t.match_prefix_with_counts(zeros)
    assert r.matched_token_count == PAGE and r.tenant in ("tenant1", "tenant3")
    r = t.match_prefix_with_counts(maxp)
    assert r.matched_token_count == PAGE and r.tenant == "tenant2"


def s_hit_ratio(mk):  # :2397
    t = mk(LRU)
    one = make_tokens(1, 1)
    t.insert_tokens(one, "tenant1")
    r = t.match_prefix_with_counts(one)
    assert r.matched_token_count / r.input_token_count == 1.0
    r = t.match_prefix_with_counts(make_tokens(1, 2))
    assert r.matched_token_count / r.input_token_count == 0.5
    assert t.match_prefix_with_counts(make_tokens(1000, 1)).matched_token_count == 0


def _three(t):
    t1, t2, t3 = make_tokens(1, 1), make_tokens(100, 1), make_tokens(200, 1)
    for s in (t1, t2, t3):
        t.insert_tokens(s, "tenant1")
    return t1, t2, t3


def s_eviction_policy_lru(mk):  # :2435
    t = mk(LRU)
    t1, t2, t3 = _three(t)
    t.match_prefix_with_counts(t2)
    t.evict_tenant("tenant1", 2 * PAGE)
    assert t.match_prefix_with_counts(t1).matched_token_count == 0
    assert t.match_prefix_with_counts(t2).matched_token_count == PAGE
    assert t.match_prefix_with_counts(t3).matched_token_count == PAGE


def s_eviction_policy_mru(mk):  # :2474
    t = mk(MRU)
    t1, t2, t3 = _three(t)
    t.evict_tenant("tenant1", 2 * PAGE)
    assert t.match_prefix_with_counts(t3).matched_token_count == 0
    assert t.match_prefix_with_counts(t1).matched_token_count == PAGE


def s_eviction_policy_fifo(mk):  # :2504
    t = mk(FIFO)
    t1, t2, t3 = _three(t)
    for _ in range(10):
        t.match_prefix_with_counts(t1)
    t.evict_tenant("tenant1", 2 * PAGE)
    assert t.match_prefix_with_counts(t1).matched_token_count == 0


def s_eviction_policy_filo(mk):  # :2533
    t = mk(FILO)
    t1, t2, t3 = _three(t)
    t.evict_tenant("tenant1", 2 * PAGE)
    assert t.match_prefix_with_counts(t3).matched_token_count == 0
    assert t.match_prefix_with_counts(t1).matched_token_count == PAGE


def s_eviction_policy_lfu(mk):  # :2563
    t = mk(LFU)
    t1, t2, t3 = _three(t)
    for _ in range(20):
        t.match_prefix_with_counts(t1)
    for _ in range(5):
        t.match_prefix_with_counts(t3)
    t.evict_tenant("tenant1", 2 * PAGE)
    assert t.match_prefix_with_counts(t2).matched_token_count == 0
    assert t.match_prefix_with_counts(t1).matched_token_count == PAGE


def s_iter_entries(mk):  # :2640-2697
    t = mk(LRU)
    assert t.entries() == []
    tokens = list(range(16))
    t.insert_tokens(tokens, "worker-1")
    leaf = [e for e in t.entries() if e[0] == tokens]
    assert leaf and [x[0] for x in leaf[0][1]] == ["worker-1"]
    t2 = mk(LRU)
    prefix = list(range(16))
    a, b = prefix + list(range(16)), prefix + list(range(100, 116))
    t2.insert_tokens(a, "wa")
    t2.insert_tokens(b, "wb")
    paths = [e[0] for e in t2.entries()]
    assert a in paths and b in paths and paths.index(a) < paths.index(b)


# ---- beyond the reference's tests: exact bookkeeping the pick depends on (derived from the code) ----

def s_repeated_insert_inflates_tenant_count(mk):
    """`advance` is added on every full-edge traversal even if the tenant already owns the node (:467-474, :594-597)."""
    t = mk(LRU)
    seq = make_tokens(1, 2)
    t.insert_tokens(seq, "w")
    assert t.tenant_token_size("w") == 32
    t.insert_tokens(seq, "w")
    assert t.tenant_token_size("w") == 64


def s_split_clones_tenants_and_counts_only_new_owner(mk):
    """Input-is-prefix split (:475-523): the intermediate clones the child's tenants; common prefix counted only for a new owner."""
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 3), "a")          # one leaf of 48 tokens
    t.insert_tokens(make_tokens(1, 1), "b")          # split at page 1; b is new on the prefix → +16
    assert t.tenant_token_size("a") == 48 and t.tenant_token_size("b") == 16
    r = t.match_prefix_with_counts(make_tokens(1, 1))
    assert r.matched_token_count == 16 and r.tenant in ("a", "b")           # clone kept `a` on the intermediate
    if hasattr(r, "valid"):
        assert sorted(r.valid) == ["a", "b"]
    t.insert_tokens(make_tokens(1, 1), "a")          # full-edge traversal of the 16-token intermediate → +16 (quirk)
    assert t.tenant_token_size("a") == 64


def s_diverging_insert_creates_branch(mk):
    """Diverge split (:524-585): intermediate + suffix + new branch; new branch tokens + common prefix for a new owner."""
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 3), "a")
    other = make_tokens(1, 1) + make_tokens(500, 2)
    t.insert_tokens(other, "b")
    assert t.tenant_token_size("b") == 48            # 32 new branch + 16 common prefix
    assert t.match_prefix_with_counts(other).tenant == "b"
    assert t.match_prefix_with_counts(make_tokens(1, 3)).tenant == "a"
    assert t.match_prefix_with_counts(make_tokens(1, 2) + make_tokens(900, 1)).matched_token_count == 32   # partial edge match


def s_eviction_walks_up_and_keeps_inflated_count(mk):
    """Leaf-first eviction with parent promotion (:832-850, :943-985) against the inflated per-tenant count (:467-474)."""
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 1), "a")          # N1{a}
    t.insert_tokens(make_tokens(1, 2), "b")          # N1{a,b} (+16)  N2{b} (+16)               → b = 32
    t.insert_tokens(make_tokens(1, 3), "b")          # N1 (+16) N2 (+16) N3{b} (+16)              → b = 80
    assert t.tenant_token_size("b") == 80
    t.evict_tenant("b", 32)                          # must shed 48: N3, then promoted N2, then promoted N1
    assert t.tenant_token_size("b") == 32            # the counter still says 32 although b owns nothing any more
    r = t.match_prefix_with_counts(make_tokens(1, 3))
    assert r.matched_token_count == 16 and r.tenant == "a"


def s_partial_page_tail_ignored(mk):
    t = mk(LRU)
    t.insert_tokens(make_tokens(1, 2) + [7, 8, 9], "a")      # 35 tokens → 32 inserted
    assert t.tenant_token_size("a") == 32
    r = t.match_prefix_with_counts(make_tokens(1, 2) + [7, 8, 9, 10])
    assert (r.matched_token_count, r.input_token_count) == (32, 36)


def s_evict_by_size_only_over_limit_tenants(mk):  # :1011-1024
    t = mk(LRU)
    for i in range(5):
        t.insert_tokens(make_tokens(1000 * i, 2), "big")
    t.insert_tokens(make_tokens(90000, 1), "small")
    t.evict_tenant_by_size(64)
    assert t.tenant_token_size("big") <= 64 and t.tenant_token_size("small") == 16
    # LRU: the two most recently inserted paths of `big` survive
    assert t.match_prefix_with_counts(make_tokens(4000, 2)).matched_token_count == 32
    assert t.match_prefix_with_counts(make_tokens(0, 2)).matched_token_count == 0


SCENARIOS = [v for k, v in sorted(globals().items()) if k.startswith("s_") and callable(v)]
