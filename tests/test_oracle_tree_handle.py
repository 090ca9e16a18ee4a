"""Pins the oracle's TreeHandle (oracle/cache_aware.h: apply_known_remote_insert / apply_repair_page, cache_aware.rs:443-645) with the
reference's own unit tests (cache_aware.rs:1159-1361, names kept)."""
from oracle import orc


def _policy():
    orc.reset_globals()
    return orc.CacheAwarePolicy(eviction_interval_secs=0)


def test_apply_known_remote_insert_round_trip():
    policy = _policy()
    text, tokens = "remote_text", [1, 2, 3, 4]
    assert policy.apply_repair_page("model1", "string", [("string", text, [("http://w1", 1)])]) == 1
    assert policy.apply_repair_page("model1", "token", [("token", tokens, [("http://w1", 1)])]) == 1
    text_hash, token_hash = orc.hash_node_path(text), orc.hash_token_path(tokens)
    assert policy.apply_known_remote_insert("model1", "string", text_hash, "http://w2")
    assert policy.apply_known_remote_insert("model1", "token", token_hash, "http://w2")
    assert not policy.apply_known_remote_insert("model1", "token", text_hash, "http://w2")        # a string hash asked of the token tree
    assert not policy.apply_known_remote_insert("model1", "string", 0xDEADBEEF, "http://w2")      # unknown hash
    assert not policy.apply_known_remote_insert("unknown_model", "string", text_hash, "http://w2")
    # the remote tenant is now on the stored path
    ents = dict(policy.string_tree("model1").entries())
    assert {t for t, _ in ents[text]} == {"http://w1", "http://w2"}


def test_apply_repair_page_seeds_hash_index():
    policy = _policy()
    text, tokens = "repaired text", [1] * 16
    assert policy.apply_repair_page("model1", "string", [("string", text, [("http://w1", 1)])]) == 1
    assert policy.apply_known_remote_insert("model1", "string", orc.hash_node_path(text), "http://w2")
    assert policy.apply_repair_page("model1", "token", [("token", tokens, [("http://w1", 1)])]) == 1
    assert policy.apply_known_remote_insert("model1", "token", orc.hash_token_path(tokens), "http://w2")
    assert policy.hash_index("text", "model1") == {orc.hash_node_path(text): text}
    assert policy.hash_index("tokens", "model1") == {orc.hash_token_path(tokens): tokens}
    tt = policy.token_tree("model1")
    assert tt.tenant_token_size("http://w1") == 16 and tt.tenant_token_size("http://w2") == 16


def test_apply_known_remote_insert_from_request_hot_path():
    """The request path itself seeds hash_index (cache_aware.rs:881-886, :950-956): hash(full request) → matched prefix."""
    policy = _policy()
    policy.set_workers(["http://w1:8000", "http://w2:8000"])
    text = "the quick brown fox jumps over the lazy dog"
    assert policy.select_worker(request_text=text).idx is not None
    tokens = list(range(32))
    assert policy.select_worker(tokens=tokens).idx is not None
    # both hashes are known now, whatever prefix they map to (the first request of a kind matches nothing: the value may be empty)
    assert policy.apply_known_remote_insert("", "string", orc.hash_node_path(text), "http://w3:8000")
    assert policy.apply_known_remote_insert("", "token", orc.hash_token_path(tokens), "http://w3:8000")
    # a second, identical request maps the hash to the full path; the remote insert then lands on it
    policy.select_worker(request_text=text)
    policy.select_worker(tokens=tokens)
    assert policy.apply_known_remote_insert("unknown", "string", orc.hash_node_path(text), "http://w3:8000")
    assert policy.apply_known_remote_insert("unknown", "token", orc.hash_token_path(tokens), "http://w3:8000")
    assert "http://w3:8000" in {t for t, _ in dict(policy.string_tree().entries())[text]}
    assert policy.token_tree().tenant_token_size("http://w3:8000") == 32


def test_repair_page_skips_entries_of_the_other_kind_and_is_idempotent():
    policy = _policy()
    page = [("string", "abc", [("http://w1", 5), ("http://w2", 6)]), ("token", [9] * 16, [("http://w1", 1)]), ("string", "abd", [("http://w1", 7)])]
    assert policy.apply_repair_page("m", "string", page) == 2          # the token entry is logged and skipped (:606-613)
    before = (policy.string_tree("m").entries(), policy.string_tree("m").get_tenant_char_count())
    assert policy.apply_repair_page("m", "string", page) == 2
    after = (policy.string_tree("m").entries(), policy.string_tree("m").get_tenant_char_count())
    assert [p for p, _ in before[0]] == [p for p, _ in after[0]] and before[1] == after[1]   # same structure and sizes; epochs move
    assert policy.token_tree("m") is None
    assert policy.apply_repair_page("m", "token", page) == 1
    assert policy.token_tree("m").tenant_token_size("http://w1") == 16
