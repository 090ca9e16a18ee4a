"""Oracle string tree pinned against the reference's own unit tests (string_tree.rs:1640-3016, ported in
scenarios_string_tree.py).  CPU only."""
import pytest

from oracle import orc
from tests import scenarios_string_tree as S


def _mk():
    orc.reset_globals()
    return orc.Tree()


@pytest.mark.parametrize("scenario", S.SCENARIOS, ids=lambda f: f.__name__)
def test_string_tree_scenarios(scenario):
    scenario(_mk)


def test_match_epoch_refresh_is_one_in_eight():
    """match_prefix_with_counts draws one epoch per call and refreshes the tenant timestamp iff epoch & 7 == 0
    (string_tree.rs:633-637)."""
    orc.reset_globals()
    t = orc.Tree()
    t.insert_text("hello", "a")           # epoch 0 → leaf a@0
    for i in range(7):
        t.match_prefix_with_counts("hello")   # epochs 1..7: no refresh
    assert dict(t.entries())["hello"] == [("a", 0)]
    t.match_prefix_with_counts("hello")       # epoch 8 → refresh
    assert dict(t.entries())["hello"] == [("a", 8)]
    assert orc.lib().orc_string_epoch() == 9
