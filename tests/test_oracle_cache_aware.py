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
