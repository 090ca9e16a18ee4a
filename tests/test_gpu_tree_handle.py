"""TreeHandle in the product (smgx_tree_apply_known_remote_insert / smgx_tree_apply_repair_page; cache_aware.rs:454-645): the reference's
three unit tests (cache_aware.rs:1159-1361, the same ones that pin the oracle in tests/test_oracle_tree_handle.py) through the C ABI, plus a
randomized repair exchange compared with the oracle entry by entry."""
import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu


def _policy():
    from smg_b200 import CacheAwareConfig, CacheAwarePolicy
    return CacheAwarePolicy(CacheAwareConfig(eviction_interval_secs=0))


def test_apply_known_remote_insert_round_trip():
    policy = _policy()
    text, tokens = "remote_text", [1, 2, 3, 4]
    assert policy.apply_repair_page("model1", "string", [("string", text, [("http://w1", 1)])]) == 1
    assert policy.apply_repair_page("model1", "token", [("token", tokens, [("http://w1", 1)])]) == 1
    text_hash, token_hash = orc.hash_node_path(text), orc.hash_token_path(tokens)
    assert policy.apply_known_remote_insert("model1", "string", text_hash, "http://w2")
    assert policy.apply_known_remote_insert("model1", "token", token_hash, "http://w2")
    assert not policy.apply_known_remote_insert("model1", "token", text_hash, "http://w2")
    assert not policy.apply_known_remote_insert("model1", "string", 0xDEADBEEF, "http://w2")
    assert not policy.apply_known_remote_insert("unknown_model", "string", text_hash, "http://w2")
    ents = dict(policy.string_tree("model1").entries())
    assert {t for t, _ in ents[text]} == {"http://w1", "http://w2"}


def test_apply_repair_page_seeds_hash_index():
    policy = _policy()
    text, tokens = "repaired text", [1] * 16
    assert policy.apply_repair_page("model1", "string", [("string", text, [("http://w1", 1)])]) == 1
    assert policy.apply_known_remote_insert("model1", "string", orc.hash_node_path(text), "http://w2")
    assert policy.apply_repair_page("model1", "token", [("token", tokens, [("http://w1", 1)])]) == 1
    assert policy.apply_known_remote_insert("model1", "token", orc.hash_token_path(tokens), "http://w2")
    assert policy.hash_index_get(orc.hash_node_path(text), "text", "model1") == text and policy.hash_index_size("text", "model1") == 1
    assert policy.hash_index_get(orc.hash_token_path(tokens), "tokens", "model1") == tokens and policy.hash_index_size("tokens", "model1") == 1
    tt = policy.token_tree("model1")
    assert tt.tenant_token_size("http://w1") == 16 and tt.tenant_token_size("http://w2") == 16


def test_apply_known_remote_insert_from_request_hot_path():
    from smg_b200 import BasicWorker, SelectWorkerInfo
    policy = _policy()
    ws = [BasicWorker("http://w1:8000"), BasicWorker("http://w2:8000")]
    policy.init_workers(ws)
    text = "the quick brown fox jumps over the lazy dog"
    assert policy.select_worker(ws, SelectWorkerInfo(request_text=text)) is not None
    tokens = list(range(32))
    assert policy.select_worker(ws, SelectWorkerInfo(tokens=tokens)) is not None
    assert policy.apply_known_remote_insert("", "string", orc.hash_node_path(text), "http://w3:8000")
    assert policy.apply_known_remote_insert("", "token", orc.hash_token_path(tokens), "http://w3:8000")
    policy.select_worker(ws, SelectWorkerInfo(request_text=text))
    policy.select_worker(ws, SelectWorkerInfo(tokens=tokens))
    assert policy.apply_known_remote_insert("unknown", "string", orc.hash_node_path(text), "http://w3:8000")
    assert policy.apply_known_remote_insert("unknown", "token", orc.hash_token_path(tokens), "http://w3:8000")
    assert "http://w3:8000" in {t for t, _ in dict(policy.string_tree().entries())[text]}
    assert policy.token_tree().tenant_token_size("http://w3:8000") == 32


def test_repair_page_skips_entries_of_the_other_kind_and_is_idempotent():
    policy = _policy()
    page = [("string", "abc", [("http://w1", 5), ("http://w2", 6)]), ("token", [9] * 16, [("http://w1", 1)]), ("string", "abd", [("http://w1", 7)])]
    assert policy.apply_repair_page("m", "string", page) == 2
    before = ([p for p, _ in policy.string_tree("m").entries()], policy.string_tree("m").get_tenant_char_count())
    assert policy.apply_repair_page("m", "string", page) == 2
    after = ([p for p, _ in policy.string_tree("m").entries()], policy.string_tree("m").get_tenant_char_count())
    assert before == after
    assert policy.apply_repair_page("m", "token", page) == 1
    assert policy.token_tree("m").tenant_token_size("http://w1") == 16


@pytest.mark.parametrize("seed", [1, 2])
def test_random_repair_exchange_matches_oracle(seed):
    """A peer's iter_entries stream applied page by page on both sides, then remote inserts by hash: same trees, same hash_index."""
    rng = np.random.default_rng(seed)
    orc.reset_globals()
    policy, op = _policy(), orc.CacheAwarePolicy(eviction_interval_secs=0)
    tenants = [f"http://w{i}:8000" for i in range(5)]
    words = ["alpha", "beta", "gamma", "delta", "épsilon", "ζ", "chat", "system: you are", " helpful", "\n"]
    for page_no in range(6):
        page = []
        for _ in range(int(rng.integers(3, 12))):
            ts = [(tenants[int(t)], int(rng.integers(1, 100))) for t in rng.choice(5, size=int(rng.integers(1, 4)), replace=False)]
            if rng.random() < 0.5:
                page.append(("string", "".join(words[int(w)] for w in rng.integers(0, len(words), size=int(rng.integers(0, 6)))), ts))
            else:
                base = list(rng.integers(0, 50, size=16)) * int(rng.integers(1, 4))
                page.append(("token", [int(x) for x in base[: int(rng.integers(0, len(base) + 1))]] + [int(x) for x in rng.integers(0, 9, size=int(rng.integers(0, 20)))], ts))
        for kind in ("string", "token"):
            assert policy.apply_repair_page("m", kind, page) == op.apply_repair_page("m", kind, page)
        for kind, path, _ in page:
            h = orc.hash_node_path(path) if kind == "string" else orc.hash_token_path(path)
            w = tenants[int(rng.integers(0, 5))]
            assert policy.apply_known_remote_insert("m", kind, h, w) == op.apply_known_remote_insert("m", kind, h, w)
        assert not policy.apply_known_remote_insert("m", "string", int(rng.integers(1, 2**62)), tenants[0])
    assert [(p, sorted(t for t, _ in ts)) for p, ts in policy.string_tree("m").entries()] == [(p, sorted(t for t, _ in ts)) for p, ts in op.string_tree("m").entries()]
    assert [(list(p), sorted(t for t, _ in ts)) for p, ts in policy.token_tree("m").entries()] == [(list(p), sorted(t for t, _ in ts)) for p, ts in op.token_tree("m").entries()]
    for t in tenants:
        assert policy.token_tree("m").tenant_token_size(t) == op.token_tree("m").tenant_token_size(t)
    assert policy.string_tree("m").get_tenant_char_count() == op.string_tree("m").get_tenant_char_count()
    assert policy.hash_index_size("text", "m") == len(op.hash_index("text", "m")) and policy.hash_index_size("tokens", "m") == len(op.hash_index("tokens", "m"))
    for h, val in op.hash_index("tokens", "m").items():
        assert policy.hash_index_get(h, "tokens", "m") == list(val)
    for h, val in op.hash_index("text", "m").items():
        assert policy.hash_index_get(h, "text", "m") == val
