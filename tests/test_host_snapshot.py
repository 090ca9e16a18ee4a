"""Mesh wire format on the product's host-authoritative string tree (no GPU: device_id = -1): snapshot bytes, from_snapshot and
merge_snapshot must equal the oracle byte for byte — ported reference tests that need no match, and seeded random trees."""
import numpy as np
import pytest

from oracle import orc
from tests import scenarios_snapshot as SC

NO_MATCH = ["test_snapshot_empty_tree", "test_snapshot_size_vs_flat_ops", "test_snapshot_wire_layout", "test_from_snapshot_accounting_and_structure",
            "test_from_snapshot_skips_empty_child_edges_and_truncation", "test_merge_remote_prefix_of_local_drops_remote_children",
            "test_malformed_bytes_are_rejected", "test_trailing_bytes_are_accepted"]


def _mk():
    from smg_b200 import Tree
    return Tree.standalone(device_id=-1)


def _from_bytes(data):
    from smg_b200 import Tree
    return Tree.from_snapshot_bytes(data, device_id=-1)


@pytest.mark.parametrize("name", NO_MATCH)
def test_host_snapshot_scenarios(name):
    SC.ALL[name](_mk, _from_bytes)


WORDS = ["route ", "the ", "request ", "to ", "a ", "worker ", "préfixe ", "κλειδί ", "缓存 ", "/v1/chat", "/v1/", "ab", "abc", "a", "🙂"]


def _random_texts(rng, n):
    out = []
    for _ in range(n):
        k = int(rng.integers(1, 7))
        out.append("".join(WORDS[int(i)] for i in rng.integers(0, len(WORDS), size=k)))
    return out


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_random_trees_snapshot_load_merge_equal_oracle(seed):
    rng = np.random.default_rng(seed)
    orc.reset_globals()
    pa, pb, oa, ob = _mk(), _mk(), orc.Tree(), orc.Tree()
    # the product draws epochs per policy handle, the oracle from one process-wide counter: interleave inserts the same way on both
    # by building A completely, then B (B's epochs start where A stopped in the oracle; in the product each handle starts at 0) —
    # so only A's snapshot is compared directly, and B is shipped to both as the ORACLE's bytes.
    for text in _random_texts(rng, 60):
        w = f"http://w{int(rng.integers(0, 5))}:8000"
        pa.insert_text(text, w); oa.insert_text(text, w)
    assert pa.snapshot_bytes() == oa.snapshot_bytes()
    for text in _random_texts(rng, 60):
        ob.insert_text(text, f"http://w{int(rng.integers(2, 8))}:8000")
    remote = ob.snapshot_bytes()
    # from_snapshot
    pr, orr = _from_bytes(remote), orc.Tree.from_snapshot_bytes(remote)
    assert pr.snapshot_bytes() == orr.snapshot_bytes() == remote
    assert pr.get_tenant_char_count() == orr.get_tenant_char_count()
    assert pr.entries() == orr.entries() and pr.node_count() == orr.node_count()
    # merge_snapshot, twice (the second is a no-op), then keep inserting on the merged tree
    for _ in range(2):
        pa.merge_snapshot_bytes(remote); oa.merge_snapshot_bytes(remote)
        assert pa.snapshot_bytes() == oa.snapshot_bytes()
        assert pa.get_tenant_char_count() == oa.get_tenant_char_count()
        assert pa.get_used_size_per_tenant() == oa.get_used_size_per_tenant()
        assert pa.node_count() == oa.node_count()
    orc.reset_globals()
    oa2 = orc.Tree.from_snapshot_bytes(oa.snapshot_bytes())
    pa2 = _from_bytes(pa.snapshot_bytes())
    for text in _random_texts(rng, 40):
        pa2.insert_text(text, "http://w1:8000"); oa2.insert_text(text, "http://w1:8000")
    assert pa2.snapshot_bytes() == oa2.snapshot_bytes()
    pa2.evict_tenant_by_size(40); oa2.evict_tenant_by_size(40)
    assert pa2.snapshot_bytes() == oa2.snapshot_bytes()
