"""Known-answer scenarios for the char-level multi-tenant radix tree (HTTP text routing), ported from the reference's
unit tests (crates/kv_index/src/string_tree.rs:1640-2780; thread-only and random tests omitted).  `mk()` returns an
object with the kv_index::Tree surface."""


def legacy(t, text):  # prefix_match_legacy (:653-657)
    r = t.match_prefix_with_counts(text)
    return text[: r.matched_char_count], r.tenant


def s_tenant_char_count(mk):  # :1640 — maintained counts == sizes recomputed from the tree, through inserts and eviction
    t = mk()
    phases = [[("apple", "tenant1"), ("apricot", "tenant1"), ("banana", "tenant1"), ("amplify", "tenant2"), ("application", "tenant2")],
              [("apartment", "tenant1"), ("appetite", "tenant2"), ("ball", "tenant1"), ("box", "tenant2")],
              [("zebra", "tenant1"), ("zebra", "tenant2"), ("zero", "tenant1"), ("zero", "tenant2")]]
    for ph in phases:
        for text, ten in ph:
            t.insert_text(text, ten)
        assert t.get_tenant_char_count() == t.get_used_size_per_tenant()
    t.evict_tenant_by_size(10)
    assert t.get_tenant_char_count() == t.get_used_size_per_tenant()
    assert all(v <= 10 for v in t.get_used_size_per_tenant().values())


def s_cold_start(mk):  # :1704
    assert legacy(mk(), "hello") == ("", "empty")


def s_exact_match_seq(mk):  # :1714
    t = mk()
    for text, ten in (("hello", "tenant1"), ("apple", "tenant2"), ("banana", "tenant3")):
        t.insert_text(text, ten)
    for text, ten in (("hello", "tenant1"), ("apple", "tenant2"), ("banana", "tenant3")):
        assert legacy(t, text) == (text, ten)


def s_utf8_split_seq(mk):  # :1950
    t = mk()
    pairs = [("你好嗎", "tenant1"), ("你好喔", "tenant2"), ("你心情好嗎", "tenant3")]
    for text, ten in pairs:
        t.insert_text(text, ten)
    for text, ten in pairs:
        assert legacy(t, text) == (text, ten)


def s_simple_eviction(mk):  # :2027
    t = mk()
    t.insert_text("hello", "tenant1")
    t.insert_text("hello", "tenant2")
    t.insert_text("world", "tenant2")
    assert t.get_used_size_per_tenant() == {"tenant1": 5, "tenant2": 10}
    t.evict_tenant_by_size(5)
    assert t.get_used_size_per_tenant() == {"tenant1": 5, "tenant2": 5}
    assert legacy(t, "world") == ("world", "tenant2")


def s_get_used_size_per_tenant(mk):  # :2194
    t = mk()
    t.insert_text("hello", "tenant1")
    t.insert_text("world", "tenant1")
    assert t.get_used_size_per_tenant()["tenant1"] == 10
    t.insert_text("hello", "tenant2")
    t.insert_text("help", "tenant2")
    s = t.get_used_size_per_tenant()
    assert s["tenant1"] == 10 and s["tenant2"] == 6
    t.insert_text("你好", "tenant3")
    assert t.get_used_size_per_tenant()["tenant3"] == 2


def s_prefix_match_tenant(mk):  # :2227
    t = mk()
    for text, ten in (("hello", "tenant1"), ("hello", "tenant2"), ("hello world", "tenant2"), ("help", "tenant1"), ("helicopter", "tenant2")):
        t.insert_text(text, ten)
    want = {("hello", "tenant1"): "hello", ("help", "tenant1"): "help", ("hel", "tenant1"): "hel", ("hello world", "tenant1"): "hello",
            ("helicopter", "tenant1"): "hel", ("hello", "tenant2"): "hello", ("hello world", "tenant2"): "hello world",
            ("helicopter", "tenant2"): "helicopter", ("hel", "tenant2"): "hel", ("help", "tenant2"): "hel",
            ("hello", "tenant3"): "", ("help", "tenant3"): ""}
    for (text, ten), exp in want.items():
        assert t.prefix_match_tenant(text, ten) == exp, (text, ten)


def s_empty_string_input(mk):  # :2266
    t = mk()
    t.insert_text("", "tenant1")
    assert legacy(t, "") == ("", "tenant1")
    t.insert_text("hello", "tenant2")
    assert legacy(t, "") == ("", "tenant1")        # the root's cached last_tenant sticks (:598-603)


def s_single_character_operations(mk):  # :2285
    t = mk()
    for text, ten in (("a", "tenant1"), ("b", "tenant2"), ("c", "tenant1")):
        t.insert_text(text, ten)
    assert legacy(t, "a") == ("a", "tenant1")
    assert legacy(t, "b") == ("b", "tenant2")
    assert legacy(t, "abc") == ("a", "tenant1")


def s_prefix_is_subset_of_existing(mk):  # :2308
    t = mk()
    t.insert_text("application", "tenant1")
    t.insert_text("app", "tenant2")
    m, ten = legacy(t, "app")
    assert m == "app" and ten in ("tenant1", "tenant2")
    assert legacy(t, "application") == ("application", "tenant1")
    assert legacy(t, "apple")[0] == "appl"


def s_existing_is_prefix_of_new(mk):  # :2334
    t = mk()
    t.insert_text("app", "tenant1")
    t.insert_text("application", "tenant2")
    m, ten = legacy(t, "app")
    assert m == "app" and ten in ("tenant1", "tenant2")
    assert legacy(t, "application") == ("application", "tenant2")
    assert legacy(t, "applesauce")[0] == "appl"


def s_counts_accuracy(mk):  # :2360
    t = mk()
    t.insert_text("hello world", "tenant1")
    for text, m, n in (("hello world", 11, 11), ("hello", 5, 5), ("hello world and more", 11, 20), ("goodbye", 0, 7)):
        r = t.match_prefix_with_counts(text)
        assert (r.matched_char_count, r.input_char_count) == (m, n)
    assert t.match_prefix_with_counts("hello world").tenant == "tenant1"


def s_counts_utf8(mk):  # :2388
    t = mk()
    t.insert_text("你好世界呀", "tenant1")
    r = t.match_prefix_with_counts("你好世界呀")
    assert (r.matched_char_count, r.input_char_count) == (5, 5)
    r = t.match_prefix_with_counts("你好")
    assert (r.matched_char_count, r.input_char_count) == (2, 2)
    t.insert_text("hello你好", "tenant2")
    r = t.match_prefix_with_counts("hello你好世界")
    assert (r.matched_char_count, r.input_char_count) == (7, 9)


def s_split_at_first_and_last_character(mk):  # :2412, :2434
    t = mk()
    t.insert_text("abc", "tenant1")
    t.insert_text("aXX", "tenant2")
    assert legacy(t, "abc") == ("abc", "tenant1") and legacy(t, "aXX") == ("aXX", "tenant2") and legacy(t, "a")[0] == "a"
    t = mk()
    t.insert_text("abcd", "tenant1")
    t.insert_text("abcX", "tenant2")
    assert legacy(t, "abcd") == ("abcd", "tenant1") and legacy(t, "abcX") == ("abcX", "tenant2") and legacy(t, "abc")[0] == "abc"


def s_multiple_splits_same_path(mk):  # :2456
    t = mk()
    for text, ten in (("abcdefgh", "tenant1"), ("abcdef", "tenant2"), ("abcd", "tenant3"), ("ab", "tenant4")):
        t.insert_text(text, ten)
    for text in ("abcdefgh", "abcdef", "abcd", "ab", "a"):
        assert legacy(t, text)[0] == text


def s_ascii_utf8_consistency(mk):  # :2517
    t = mk()
    t.insert_text("hello", "tenant1")
    t.insert_text("你好", "tenant2")
    t.insert_text("hello你好", "tenant3")
    for text, n in (("hello", 5), ("你好", 2), ("hello你好", 7)):
        r = t.match_prefix_with_counts(text)
        assert (r.matched_char_count, r.input_char_count) == (n, n)


def s_emoji_handling(mk):  # :2549
    t = mk()
    t.insert_text("hello 👋", "tenant1")
    t.insert_text("hello 👋🌍", "tenant2")
    assert legacy(t, "hello 👋") == ("hello 👋", "tenant1")
    assert legacy(t, "hello 👋🌍") == ("hello 👋🌍", "tenant2")
    r = t.match_prefix_with_counts("hello 👋")
    assert (r.matched_char_count, r.input_char_count) == (7, 7)


def s_eviction_edge_cases(mk):  # :2573, :2584, :2601
    t = mk()
    t.evict_tenant_by_size(100)
    assert t.get_used_size_per_tenant() == {}
    t.insert_text("hello", "tenant1")
    t.insert_text("world", "tenant1")
    t.evict_tenant_by_size(0)
    assert all(v == 0 for v in t.get_used_size_per_tenant().values())
    t = mk()
    for i in range(100):
        t.insert_text(f"entry{i:03d}", "tenant1")
    assert t.get_used_size_per_tenant()["tenant1"] > 50
    t.evict_tenant_by_size(50)
    assert t.get_used_size_per_tenant().get("tenant1", 0) <= 50


def s_last_tenant_cache(mk):  # :2626 + leaf creation sets last_tenant (:438), split copies it (:478)
    t = mk()
    t.insert_text("hello", "tenant1")
    assert legacy(t, "hello")[1] == "tenant1"
    t.insert_text("hello", "tenant2")
    assert legacy(t, "hello") == ("hello", "tenant1")      # cached creator still valid → deterministic
    t.insert_text("help", "tenant0")                        # split "hel" copies last_tenant = tenant1
    assert legacy(t, "hel") == ("hel", "tenant1")
    assert legacy(t, "help") == ("help", "tenant0")


def s_very_long_and_special(mk):  # :2741, :2762, :2780
    t = mk()
    long_text = "a" * 10000
    t.insert_text(long_text, "tenant1")
    assert t.match_prefix_with_counts(long_text).matched_char_count == 10000
    assert t.match_prefix_with_counts(long_text[:5000]).matched_char_count == 5000
    for i in range(100):
        t.insert_text("shared/path", f"tenant{i}")
    assert legacy(t, "shared/path")[0] == "shared/path"
    for text in ("hello\nworld", "tab\there", "quote\"inside", "back\\slash", "nul\x00byte"):
        t.insert_text(text, "sp")
        assert legacy(t, text)[0] == text


def s_iter_entries(mk):  # :2934-3016
    t = mk()
    assert t.entries() == []
    t.insert_text("hello", "worker-1")
    es = dict((p, [x[0] for x in ts]) for p, ts in t.entries())
    assert es.get("hello") == ["worker-1"]
    t.insert_text("help", "worker-2")
    paths = [p for p, _ in t.entries()]
    assert "hel" in paths and "hello" in paths and "help" in paths
    assert paths.index("hel") < paths.index("hello") < paths.index("help")     # pre-order, children in char order
    t2 = mk()
    t2.insert_text("你好世界", "w")
    assert [p for p, _ in t2.entries() if p] == ["你好世界"]


SCENARIOS = [v for k, v in sorted(globals().items()) if k.startswith("s_") and callable(v)]
