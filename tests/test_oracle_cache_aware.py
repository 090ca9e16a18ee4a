"""Oracle policy pinned against the reference's policy-level unit tests (cache_aware.rs:999-2015, mod.rs:192-262)."""
import pytest

from oracle import orc
from tests import scenarios_cache_aware as S
from tests.policy_adapters import OraclePolicy


def _mk(config):
    orc.reset_globals()
    return OraclePolicy(config)


@pytest.mark.parametrize("name", list(S.ALL), ids=list(S.ALL))
def test_policy_scenarios(name):
    S.ALL[name](_mk)


def test_init_workers_does_not_touch_token_tree_root():
    """insert_tokens(&[], url) returns before registering the tenant at the root (token_tree.rs:403-407), while
    insert_text("", url) does register it (string_tree.rs:399-408, :552-556)."""
    orc.reset_globals()
    p = orc.CacheAwarePolicy(eviction_interval_secs=0)
    p.set_workers(["http://a", "http://b"])
    assert p.token_tree().match_prefix_with_counts(list(range(16))).tenant == "empty"
    r = p.string_tree().match_prefix_with_counts("zzz")
    assert r.matched_char_count == 0 and r.tenant in ("http://a", "http://b") and sorted(r.valid) == ["http://a", "http://b"]


def test_snapshot_batch_equals_sequential_without_conflicts_and_differs_with():
    """begin/end_snapshot_batch (oracle/cache_aware.h): every request walks the pre-batch tree, then side effects replay in
    request order.  Without intra-batch conflicts it is the one-by-one execution, timestamps included."""
    import numpy as np
    from oracle import orc

    def run(snapshot, conflict):
        orc.reset_globals()
        p = orc.CacheAwarePolicy(cache_threshold=0.3, balance_abs_threshold=64, balance_rel_threshold=1.5)
        p.set_workers([f"http://w{i}:8000" for i in range(4)])
        p.set_state([3, 1, 2, 0], [1] * 4, [1] * 4)
        rng = np.random.default_rng(1)
        reqs = [rng.integers(0, 1000, size=64, dtype=np.uint32) for _ in range(8)]
        if conflict:
            reqs[5] = reqs[2].copy()
        flat, off = np.concatenate(reqs), np.arange(9, dtype=np.uint64) * 64
        outs = [p.select_batch_tokens(flat, off, snapshot=snapshot)[:3] for _ in range(2)]
        texts = ["alpha one", "beta two", "gamma three"] + (["alpha one"] if conflict else [])
        outs.append(p.select_batch_text(texts, snapshot=snapshot)[:4])
        return outs, p.token_tree().entries(), p.string_tree().entries()

    a, b = run(False, False), run(True, False)
    assert all(np.array_equal(x, y) for u, v in zip(a[0], b[0]) for x, y in zip(u, v)) and a[1:] == b[1:]
    a, b = run(False, True), run(True, True)
    assert a[0][0][2][5] == 64 and b[0][0][2][5] == 0          # the duplicate sees the first insert only one-by-one
    assert a[0][2][2][3] == len("alpha one") and b[0][2][2][3] == 0
