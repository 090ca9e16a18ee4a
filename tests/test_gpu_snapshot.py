"""Mesh wire format end to end on the GPU-backed string tree: after from_snapshot / merge_snapshot the DEVICE mirror must answer
match_prefix_with_counts exactly like the oracle tree that went through the same bytes (reference tests + seeded random trees)."""
import numpy as np
import pytest

from oracle import orc
from tests import scenarios_snapshot as SC
from tests.test_host_snapshot import _random_texts

pytestmark = pytest.mark.gpu


def _mk():
    from smg_b200 import Tree
    return Tree.standalone(device_id=0)


def _from_bytes(data):
    from smg_b200 import Tree
    return Tree.from_snapshot_bytes(data, device_id=0)


@pytest.mark.parametrize("name", sorted(SC.ALL))
def test_gpu_snapshot_scenarios(name):
    SC.ALL[name](_mk, _from_bytes)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_matches_after_load_and_merge_equal_oracle(seed):
    rng = np.random.default_rng(seed)
    orc.reset_globals()
    pa, oa, ob = _mk(), orc.Tree(), orc.Tree()
    for text in _random_texts(rng, 80):
        w = f"http://w{int(rng.integers(0, 5))}:8000"
        pa.insert_text(text, w); oa.insert_text(text, w)
    assert pa.match_prefix_with_counts("route the request").matched_char_count == oa.match_prefix_with_counts("route the request").matched_char_count
    for text in _random_texts(rng, 80):
        ob.insert_text(text, f"http://w{int(rng.integers(2, 8))}:8000")
    remote = ob.snapshot_bytes()
    pr, orr = _from_bytes(remote), orc.Tree.from_snapshot_bytes(remote)
    pa.merge_snapshot_bytes(remote); oa.merge_snapshot_bytes(remote)
    assert pa.snapshot_bytes() == oa.snapshot_bytes()
    # the oracle's epoch counter is process-wide, the product's per handle: compare what a match returns (matched chars, tenant),
    # on trees whose last_tenant caches and tenant sets are equal; epochs drawn by the matches themselves are not compared
    for q in _random_texts(rng, 200) + ["", "zzz", "route", "缓", "/v1/chat/completions"]:
        for pt, ot in ((pr, orr), (pa, oa)):
            g, w = pt.match_prefix_with_counts(q), ot.match_prefix_with_counts(q)
            assert (g.matched_char_count, g.input_char_count) == (w.matched_char_count, w.input_char_count), q
            assert g.tenant == w.tenant or g.tenant in w.valid, q
