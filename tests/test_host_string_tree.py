"""Host side of the string-tree mode without a GPU (device_id = -1): the library's host-authoritative tree must evolve
exactly like the oracle under inserts, evictions and prefix_match_tenant — node structure, tenants, epochs and the
maintained per-tenant char counts.  (The walk + pick run on the GPU and are covered by tests/test_gpu_string_tree.py.)"""
import random

import pytest

from oracle import orc
from smg_b200.policy import CacheAwareConfig, Tree, _Handle

WORDS = ["hello", "help", "helicopter", "你好", "你好嗎", "你心情好嗎", "héllo", "hé", "👋🌍", "👋", "app", "apple", "application", "", "a", "ab", "abc"]


def _mk():
    orc.reset_globals()
    return Tree(_Handle(CacheAwareConfig(eviction_interval_secs=0), -1)), orc.Tree()


@pytest.mark.parametrize("seed", range(6))
def test_host_tree_tracks_oracle_under_inserts_and_eviction(seed):
    t, o = _mk()
    r = random.Random(seed)
    tenants = [f"http://w{i}:8000" for i in range(5)]
    for step in range(300):
        text = " ".join(r.choice(WORDS) for _ in range(r.randrange(0, 5)))
        ten = r.choice(tenants)
        t.insert_text(text, ten)
        o.insert_text(text, ten)
        if step % 37 == 36:
            q, qt = r.choice(WORDS) + " " + r.choice(WORDS), r.choice(tenants + ["nobody"])
            assert t.prefix_match_tenant(q, qt) == o.prefix_match_tenant(q, qt)
        if step % 97 == 96:
            lim = r.randrange(0, 120)
            t.evict_tenant_by_size(lim)
            o.evict_tenant_by_size(lim)
            assert t.entries() == o.entries()
    assert t.entries() == o.entries()
    assert t.get_tenant_char_count() == o.get_tenant_char_count()
    assert t.get_used_size_per_tenant() == o.get_used_size_per_tenant()
    assert t.node_count() == o.node_count()


def test_match_without_device_fails_loudly():
    from smg_b200 import SmgxError
    t, _ = _mk()
    t.insert_text("hello", "a")
    with pytest.raises(SmgxError):
        t.match_prefix_with_counts("hello")     # no CPU fallback for the walk


def test_background_eviction_thread_bounds_the_trees():
    """eviction_interval_secs > 0 starts the reference's periodic eviction (cache_aware.rs:126-199) inside the library."""
    import time
    h = _Handle(CacheAwareConfig(eviction_interval_secs=1, max_tree_size=40), -1)
    t = Tree(h)
    for i in range(60):
        t.insert_text(f"entry number {i:03d} with some tail", "http://w0:8000")
    assert t.get_used_size_per_tenant()["http://w0:8000"] > 40
    deadline = time.time() + 6
    while time.time() < deadline and t.get_used_size_per_tenant().get("http://w0:8000", 0) > 40:
        time.sleep(0.2)
    assert t.get_used_size_per_tenant().get("http://w0:8000", 0) <= 40
    h.close()          # joins the thread
