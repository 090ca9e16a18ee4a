"""Mesh wire format scenarios shared by the oracle (CPU) and product (GPU) test modules: the reference's snapshot / merge unit tests
(crates/kv_index/src/string_tree.rs:2804-2935 and snapshot.rs:66-109, names kept) plus byte-level checks of the bincode layout.
`mk()` returns an empty tree with insert_text / match_prefix_with_counts / snapshot_bytes / merge_snapshot_bytes / entries /
get_tenant_char_count / get_used_size_per_tenant / node_count; `from_bytes(data)` builds a tree from a snapshot."""
from oracle.orc import decode_snapshot, encode_snapshot


def _match(t, text):
    r = t.match_prefix_with_counts(text)
    return r.matched_char_count, r.tenant


def test_snapshot_empty_tree(mk, from_bytes):
    snap = decode_snapshot(mk().snapshot_bytes())
    assert snap == [("", [], 0)]            # node_count() == 1: just the root


def test_snapshot_round_trip_single_entry(mk, from_bytes):
    t = mk()
    t.insert_text("Hello world", "worker-1")
    r = from_bytes(t.snapshot_bytes())
    assert _match(r, "Hello world") == (11, "worker-1")


def test_snapshot_round_trip_shared_prefixes(mk, from_bytes):
    t = mk()
    t.insert_text("Hello world", "worker-1")
    t.insert_text("Hello there", "worker-2")
    t.insert_text("Goodbye", "worker-3")
    r = from_bytes(t.snapshot_bytes())
    assert _match(r, "Hello world") == (11, "worker-1")
    assert _match(r, "Hello there") == (11, "worker-2")
    assert _match(r, "Goodbye") == (7, "worker-3")


def test_snapshot_size_vs_flat_ops(mk, from_bytes):
    t = mk()
    prefix = "A" * 10000
    for i in range(100):
        t.insert_text(f"{prefix}_{i}", f"worker-{i}")
    data = t.snapshot_bytes()
    assert len(data) < sum(len(f"{prefix}_{i}") for i in range(100)) // 2


def test_snapshot_bincode_round_trip(mk, from_bytes):
    t = mk()
    t.insert_text("Hello world", "worker-1")
    t.insert_text("Hello there", "worker-2")
    data = t.snapshot_bytes()
    assert encode_snapshot(decode_snapshot(data)) == data
    assert _match(from_bytes(data), "Hello world")[0] == 11


def test_snapshot_wire_layout(mk, from_bytes):
    """Pre-order, children in char order, u64-LE lengths, u32-LE child_count (bincode 1.3 defaults); epochs as stored."""
    t = mk()
    t.insert_text("Hello world", "w1")      # epoch 0
    t.insert_text("Hello there", "w2")      # split: "Hello " keeps w1's epoch, w2 joins with 0; leaf "there" gets epoch 1
    t.insert_text("héllo", "w1")            # multi-byte edge; epoch 2
    nodes = decode_snapshot(t.snapshot_bytes())
    assert [(e, cc) for e, _, cc in nodes] == [("", 2, ), ("Hello ", 2), ("there", 0), ("world", 0), ("héllo", 0)]
    assert sorted(nodes[0][1]) == [("w1", 0), ("w2", 0)]
    assert sorted(nodes[1][1]) == [("w1", 0), ("w2", 0)]
    assert nodes[2][1] == [("w2", 1)] and nodes[3][1] == [("w1", 0)] and nodes[4][1] == [("w1", 2)]
    raw = t.snapshot_bytes()
    assert raw[:8] == (5).to_bytes(8, "little") and raw[8:16] == (0).to_bytes(8, "little")   # 5 nodes; root edge ""
    assert raw[-4:] == (0).to_bytes(4, "little")


def test_from_snapshot_accounting_and_structure(mk, from_bytes):
    t = mk()
    for text, w in [("alpha", "w1"), ("alpha-beta", "w2"), ("zulu", "w3"), ("alp", "w3"), ("日本語のテキスト", "w2"), ("日本", "w1")]:
        t.insert_text(text, w)
    r = from_bytes(t.snapshot_bytes())
    assert r.entries() == t.entries()                                  # test_iter_entries_round_trips_via_snapshot
    assert r.get_used_size_per_tenant() == t.get_used_size_per_tenant()
    assert r.get_tenant_char_count() == r.get_used_size_per_tenant()   # restore_node counts every listed tenant once per node
    assert r.node_count() == t.node_count()
    assert r.snapshot_bytes() == t.snapshot_bytes()


def test_from_snapshot_skips_empty_child_edges_and_truncation(mk, from_bytes):
    """restore_node: a child with an empty edge is skipped with its whole subtree (:1280-1296); a short node list just stops."""
    nodes = [("", [("w1", 5)], 3), ("", [("bad", 1)], 1), ("lost", [("bad", 2)], 0), ("ab", [("w1", 7)], 1), ("c", [("w2", 9)], 0)]
    r = from_bytes(encode_snapshot(nodes))                             # root claims 3 children, only 2 subtrees follow
    assert r.entries() == [("", [("w1", 5)]), ("ab", [("w1", 7)]), ("abc", [("w2", 9)])]
    assert r.get_tenant_char_count() == {"w1": 2, "w2": 1}


def test_merge_disjoint_trees(mk, from_bytes):
    t1, t2 = mk(), mk()
    t1.insert_text("Hello", "worker-1")
    t2.insert_text("Goodbye", "worker-2")
    t1.merge_snapshot_bytes(t2.snapshot_bytes())
    assert _match(t1, "Hello")[0] == 5
    assert _match(t1, "Goodbye") == (7, "worker-2")


def test_merge_overlapping_trees(mk, from_bytes):
    t1, t2 = mk(), mk()
    t1.insert_text("Hello world", "worker-1")
    t2.insert_text("Hello there", "worker-2")
    t1.merge_snapshot_bytes(t2.snapshot_bytes())
    assert _match(t1, "Hello world")[0] == 11
    assert _match(t1, "Hello there") == (11, "worker-2")


def test_merge_three_cases_epochs_and_counts(mk, from_bytes):
    local, remote = mk(), mk()
    local.insert_text("shared-prefix/one", "a")        # epochs 0..
    local.insert_text("shared", "b")
    local.insert_text("exact", "a")
    remote.insert_text("shared-prefix/two", "c")        # case 3 below "shared" → "-prefix/" split
    remote.insert_text("shared-prefix", "a")
    remote.insert_text("exact", "c")                    # case 1
    remote.insert_text("exactly", "a")                  # case 2 (local "exact" is a prefix)
    remote.insert_text("novel", "d")                    # no local child
    local.merge_snapshot_bytes(remote.snapshot_bytes())
    ents = dict(local.entries())
    # case 2 descends into the deeper local child WITHOUT comparing its edge with the remote remainder (:1437-1441): the remote
    # "-prefix" node is merged into local "-prefix/one", so remote "/two" lands under it.  Restated as is.
    for path in ["shared", "shared-prefix/one", "shared-prefix/one/two", "exact", "exactly", "novel"]:
        assert path in ents, path
    assert "shared-prefix/two" not in ents and {t for t, _ in ents["shared-prefix/one"]} == {"a", "c"}
    assert {t for t, _ in ents["exact"]} == {"a", "c"}
    assert {t for t, _ in ents["novel"]} == {"d"}
    assert local.get_tenant_char_count() == local.get_used_size_per_tenant()
    assert _match(local, "shared-prefix/two")[0] == 14 and _match(local, "exactly")[0] == 7
    # merging the same snapshot again changes nothing (remote epochs are not newer)
    before = (local.entries(), local.get_tenant_char_count())
    local.merge_snapshot_bytes(remote.snapshot_bytes())
    assert (local.entries(), local.get_tenant_char_count()) == before


def test_merge_remote_prefix_of_local_drops_remote_children(mk, from_bytes):
    """Case 3 with an empty remote remainder (:1517): the split node takes the remote tenants, the remote child's children are dropped."""
    local = mk()
    local.insert_text("abcdef", "l")
    remote = from_bytes(encode_snapshot([("", [("r", 0)], 1), ("abc", [("r", 4)], 1), ("xyz", [("r", 5)], 0)]))
    local.merge_snapshot_bytes(remote.snapshot_bytes())
    ents = dict(local.entries())
    assert set(ents) == {"", "abc", "abcdef"} and dict(ents["abc"]) == {"l": 0, "r": 4}
    assert local.get_tenant_char_count() == {"l": 6, "r": 3}


def test_malformed_bytes_are_rejected(mk, from_bytes):
    import pytest
    t = mk()
    t.insert_text("Hello", "w")
    data = t.snapshot_bytes()
    for bad in (data[:-1], b"\x01", (3).to_bytes(8, "little") + data[8:], data[:8] + b"\x02\x00\x00\x00\x00\x00\x00\x00\xff\xfe" + data[8:]):
        with pytest.raises(Exception):
            from_bytes(bad)


def test_trailing_bytes_are_accepted(mk, from_bytes):
    """TreeSnapshot::from_bytes is bincode::deserialize (snapshot.rs:49-51; bincode 1.3 = DefaultOptions + fixint + allow_trailing_bytes):
    a padded / extended payload deserialises to the same snapshot, for from_snapshot and for merge_snapshot."""
    t = mk()
    t.insert_text("Hello world", "w1")
    t.insert_text("Hello there", "w2")
    data = t.snapshot_bytes()
    for pad in (b"\x00", b"\xff" * 7, b"trailing garbage that is not a node"):
        assert decode_snapshot(data + pad, allow_trailing=True) == decode_snapshot(data)
        r = from_bytes(data + pad)
        assert r.snapshot_bytes() == data
        m = mk()
        m.insert_text("Help", "w3")
        m2 = mk()
        m2.insert_text("Help", "w3")
        m.merge_snapshot_bytes(data + pad)
        m2.merge_snapshot_bytes(data)
        shape = lambda tr: [(e, sorted(n for n, _ in tens)) for e, tens in tr.entries()]   # epochs come from a process-wide clock
        assert shape(m) == shape(m2) and m.get_tenant_char_count() == m2.get_tenant_char_count()


ALL = {k: v for k, v in globals().items() if k.startswith("test_")}
