"""Pins the oracle's mesh wire format (oracle/string_tree.h: snapshot / load_snapshot / merge_snapshot + bincode) with the
reference's own unit tests and byte-level layout checks (tests/scenarios_snapshot.py)."""
import pytest

from oracle import orc
from tests import scenarios_snapshot as SC


def _mk():
    return orc.Tree()


@pytest.mark.parametrize("name", sorted(SC.ALL))
def test_oracle_snapshot(name):
    orc.reset_globals()
    SC.ALL[name](_mk, orc.Tree.from_snapshot_bytes)
