"""Oracle token tree pinned against the reference's own unit tests (token_tree.rs:1192-2697, ported in
scenarios_token_tree.py).  CPU only."""
import pytest

from oracle import orc
from tests import scenarios_token_tree as S


def _mk(policy):
    orc.reset_globals()
    return orc.TokenTree(policy)


@pytest.mark.parametrize("scenario", S.SCENARIOS, ids=lambda f: f.__name__)
def test_token_tree_scenarios(scenario):
    scenario(_mk)


def test_last_tenant_refresh_follows_global_timestamp_phase():
    """touch_tenant refreshes last_tenant only when the freshly drawn global timestamp has ts & 0xF == 0
    (token_tree.rs:308-313); Node::new draws one timestamp too (:219).  Deterministic in a fresh process."""
    orc.reset_globals()
    t = orc.TokenTree()
    seq = S.make_tokens(1, 1)
    t.insert_tokens(seq, "zeta")     # ts 0 = creation, ts 1 = touch → no refresh; any-tenant falls back to the map
    t.insert_tokens(seq, "alpha")    # ts 2
    r = t.match_prefix_with_counts(seq)   # slow path: lexicographically smallest, full valid set; touch draws ts 3
    assert r.tenant == "alpha" and sorted(r.valid) == ["alpha", "zeta"]
    for _ in range(12):              # ts 4..15
        t.match_prefix_with_counts(seq)
    t.insert_tokens(seq, "zeta")     # ts 16 → 16 & 0xF == 0 → last_tenant = zeta, now deterministic
    r = t.match_prefix_with_counts(seq)
    assert r.tenant == "zeta" and r.valid == ["zeta"]
    assert orc.lib().orc_token_ts() == 18
