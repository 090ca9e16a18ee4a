// C++ host-mirror tests: smgx::CacheAwarePolicy (include/smgx.hpp, over libsmgx.so) against the reference's own unit tests for the
// path (model_gateway/src/policies/cache_aware.rs, test names kept) and against the CPU oracle (oracle/*.h) on seeded streams.
// Test infrastructure: the oracle is the checker here, never the thing under test.
//   test_cache_aware            every test (needs a B200)
//   test_cache_aware --host     the subset that runs without a GPU (device_id = -1: writers work, every select must fail loudly)
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../../include/smgx.hpp"
#include "../../oracle/cache_aware.h"
#include "../../oracle/prefix_hash.h"
#include "../../oracle/string_tree.h"

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                                                \
    do {                                                                                           \
        ++g_checks;                                                                                \
        if (!(cond)) { ++g_fail; std::fprintf(stderr, "FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
#define CHECK_EQ(a, b)                                                                                                   \
    do {                                                                                                                 \
        ++g_checks;                                                                                                      \
        const long long _a = (long long)(a), _b = (long long)(b);                                                        \
        if (_a != _b) { ++g_fail; std::fprintf(stderr, "FAIL %s:%d  %s == %s  (%lld vs %lld)\n", __FILE__, __LINE__, #a, #b, _a, _b); } \
    } while (0)

using smgx::BasicWorker;
using smgx::CacheAwareConfig;
using smgx::CacheAwarePolicy;
using smgx::SelectWorkerInfo;
using smgx::Workers;

static int g_device = 0;

static CacheAwareConfig test_config() {   // cache_aware.rs:1422-1430
    CacheAwareConfig c;
    c.cache_threshold = 0.5f; c.balance_abs_threshold = 32; c.balance_rel_threshold = 1.1f; c.eviction_interval_secs = 0; c.max_tree_size = 10000; c.block_size = 4;
    return c;
}
static Workers two_workers() {
    return Workers{std::make_shared<BasicWorker>("http://w1:8000"), std::make_shared<BasicWorker>("http://w2:8000")};
}
static BasicWorker& bw(const Workers& ws, size_t i) { return static_cast<BasicWorker&>(*ws[i]); }
static SelectWorkerInfo text(const std::string& s) { SelectWorkerInfo i; i.request_text = s; return i; }
static SelectWorkerInfo toks(std::vector<uint32_t> t) { SelectWorkerInfo i; i.tokens = std::move(t); return i; }

// setup_indexer_with_blocks (cache_aware.rs:1662-1683): one worker's chain of blocks with seq hashes 1..n
static void store_blocks(smgx::PositionalIndexer& ix, const std::string& url, const std::vector<std::vector<uint32_t>>& blocks, uint32_t bs) {
    const uint32_t wid = ix.intern_worker(url);
    std::vector<uint64_t> seq;
    std::vector<uint32_t> flat;
    for (size_t i = 0; i < blocks.size(); ++i) { seq.push_back(i + 1); flat.insert(flat.end(), blocks[i].begin(), blocks[i].end()); }
    ix.apply_stored_tokens(wid, seq, flat, bs);
}

// ---------------------------------------------------------------------------------------------------------------------------------
static void test_cache_aware_with_balanced_load() {   // cache_aware.rs:998-1060
    CacheAwareConfig c; c.eviction_interval_secs = 0;
    CacheAwarePolicy policy(c, g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    auto i1 = policy.select_worker(ws, text("hello world"));
    auto i2 = policy.select_worker(ws, text("hello world"));
    auto i3 = policy.select_worker(ws, text("hello"));
    CHECK(i1 && i2 && i3);
    CHECK_EQ(*i1, *i2);
    CHECK_EQ(*i1, *i3);
}

static void test_cache_aware_with_imbalanced_load() {   // cache_aware.rs:1062-1100
    CacheAwareConfig c; c.cache_threshold = 0.5f; c.balance_abs_threshold = 5; c.balance_rel_threshold = 2.0f; c.eviction_interval_secs = 0; c.max_tree_size = 10000;
    CacheAwarePolicy policy(c, g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    bw(ws, 0).set_load(20);   // worker1.increment_load() × 20
    for (int k = 0; k < 3; ++k) {
        auto idx = policy.select_worker(ws, text("test"));
        CHECK(idx.has_value());
        CHECK_EQ(*idx, 1);
    }
}

static void test_cache_aware_worker_removal() {   // cache_aware.rs:1102-1157 (the slice shrinks; the tree keeps the tenant)
    CacheAwareConfig c; c.eviction_interval_secs = 0;
    CacheAwarePolicy policy(c, g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    policy.select_worker(ws, text("test1"));
    policy.select_worker(ws, text("test2"));
    policy.remove_worker_by_url("http://w1:8000");
    Workers remaining{ws[1]};
    auto idx = policy.select_worker(remaining, text("test1"));
    CHECK(idx.has_value());
    CHECK_EQ(*idx, 0);
}

static void test_no_healthy_returns_none() {   // cache_aware.rs:651-655
    CacheAwarePolicy policy(test_config(), g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    bw(ws, 0).set_healthy(false);
    bw(ws, 1).set_circuit_ok(false);
    CHECK(!policy.select_worker(ws, text("x")).has_value());
    CHECK(!policy.select_worker(ws, toks({1, 2, 3, 4})).has_value());
    CHECK(!policy.select_worker(Workers{}, text("x")).has_value());
}

static void test_event_driven_overlap_selects_cached_worker() {   // cache_aware.rs:1685-1722
    CacheAwarePolicy policy(test_config(), g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    auto monitor = policy.kv_event_monitor();
    auto ix = monitor->create_indexer("unknown", 64);
    monitor->set_block_size("unknown", 4);
    store_blocks(*ix, "http://w1:8000", {{1, 2, 3, 4}, {5, 6, 7, 8}}, 4);
    policy.set_kv_event_monitor(monitor);
    auto idx = policy.select_worker(ws, toks({1, 2, 3, 4, 5, 6, 7, 8}));
    CHECK(idx.has_value());
    CHECK_EQ(*idx, 0);
    CHECK_EQ(bw(ws, 0).processed(), 1);
}

static void test_event_driven_no_overlap_uses_min_load() {   // cache_aware.rs:1724-1761
    CacheAwarePolicy policy(test_config(), g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    auto monitor = policy.kv_event_monitor();
    auto ix = monitor->create_indexer("unknown", 64);
    monitor->set_block_size("unknown", 4);
    store_blocks(*ix, "http://w1:8000", {{1, 2, 3, 4}}, 4);
    policy.set_kv_event_monitor(monitor);
    bw(ws, 0).set_load(5);
    auto d = policy.select_worker_batch(ws, {{99, 98, 97, 96}});
    CHECK_EQ(d[0].idx, 1);
    CHECK_EQ(d[0].info.branch, SMGX_BR_EVENT_MIN_LOAD);
}

static void test_event_driven_short_request_uses_min_load() {   // cache_aware.rs:1763-1798
    CacheAwarePolicy policy(test_config(), g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    auto monitor = policy.kv_event_monitor();
    auto ix = monitor->create_indexer("unknown", 64);
    monitor->set_block_size("unknown", 4);
    store_blocks(*ix, "http://w1:8000", {{1, 2, 3, 4}}, 4);
    policy.set_kv_event_monitor(monitor);
    bw(ws, 0).set_load(3);
    auto d = policy.select_worker_batch(ws, {{1, 2}});   // < one block
    CHECK_EQ(d[0].idx, 1);
    CHECK_EQ(d[0].info.branch, SMGX_BR_EVENT_MIN_LOAD);
}

static void test_no_monitor_uses_token_tree() {   // cache_aware.rs:1800-1833
    CacheAwarePolicy policy(test_config(), g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    std::vector<uint32_t> t(32);
    for (uint32_t i = 0; i < 32; ++i) t[i] = i + 1;
    auto i1 = policy.select_worker(ws, toks(t));
    auto i2 = policy.select_worker(ws, toks(t));
    CHECK(i1 && i2);
    CHECK_EQ(*i1, *i2);
    auto d = policy.select_worker_batch(ws, {t});
    CHECK_EQ(d[0].info.branch, SMGX_BR_TREE_MATCH);
    CHECK_EQ(d[0].info.matched, 32);
}

static void test_imbalanced_skips_event_driven() {   // cache_aware.rs:1915-1957
    CacheAwareConfig c = test_config(); c.balance_abs_threshold = 5; c.balance_rel_threshold = 1.5f;
    CacheAwarePolicy policy(c, g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    auto monitor = policy.kv_event_monitor();
    auto ix = monitor->create_indexer("unknown", 64);
    monitor->set_block_size("unknown", 4);
    store_blocks(*ix, "http://w1:8000", {{1, 2, 3, 4}, {5, 6, 7, 8}}, 4);
    policy.set_kv_event_monitor(monitor);
    bw(ws, 0).set_load(20);
    auto d = policy.select_worker_batch(ws, {{1, 2, 3, 4, 5, 6, 7, 8}});
    CHECK_EQ(d[0].idx, 1);   // w1 holds the blocks but is overloaded
    CHECK_EQ(d[0].info.branch, SMGX_BR_IMBALANCED_MIN_LOAD);
}

static void test_empty_indexer_falls_through_to_token_tree() {   // cache_aware.rs:1959-1990
    CacheAwarePolicy policy(test_config(), g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    auto monitor = policy.kv_event_monitor();
    monitor->create_indexer("unknown", 64);
    policy.set_kv_event_monitor(monitor);
    std::vector<uint32_t> t(32);
    for (uint32_t i = 0; i < 32; ++i) t[i] = i + 1;
    auto d = policy.select_worker_batch(ws, {t});
    CHECK(d[0].idx >= 0);
    CHECK(d[0].info.branch == SMGX_BR_TREE_MIN_LOAD || d[0].info.branch == SMGX_BR_TREE_MATCH);
}

static void test_kv_events_fresh_chain_fallback() {   // kv_event_monitor.rs:525-597
    CacheAwarePolicy policy(test_config(), g_device);
    Workers ws = two_workers();
    policy.init_workers(ws);
    auto monitor = policy.kv_event_monitor();
    auto ix = monitor->create_indexer("unknown", 64);
    const uint32_t w2 = ix->intern_worker("http://w2:8000");
    smgx::KvCacheEvent ev;
    ev.blocks.push_back({11, {1, 2, 3, 4}, 4});   // block_size 4 is learned from the first stored block (:270-296)
    ev.blocks.push_back({12, {5, 6, 7, 8}, 4});
    ev.parent_block_hash = 4242;                  // unknown parent → stored as a fresh chain (:559-571)
    CHECK_EQ(monitor->apply_events("unknown", w2, {ev}), 1);
    CHECK_EQ(ix->current_size(), 2);
    smgx::KvCacheEvent more;
    more.blocks.push_back({13, {9, 10, 11, 12}, 4});
    more.parent_block_hash = 12;
    CHECK_EQ(monitor->apply_events("unknown", w2, {more}), 0);
    policy.set_kv_event_monitor(monitor);
    auto d = policy.select_worker_batch(ws, {{1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13}});
    CHECK_EQ(d[0].idx, 1);
    CHECK_EQ(d[0].info.branch, SMGX_BR_EVENT_OVERLAP);
    CHECK_EQ(d[0].info.matched, 3);
    auto ov = ix->find_matches({orc::compute_content_hash(std::vector<uint32_t>{1, 2, 3, 4}.data(), 4)});
    CHECK_EQ(ov.scores.size(), 1);
    CHECK_EQ(ov.scores[w2], 1);
    CHECK_EQ(ov.tree_sizes[w2], 3);
}

// ---- seeded streams against the oracle --------------------------------------------------------------------------------------------
static std::vector<orc::Worker> mirror(const Workers& ws) {
    std::vector<orc::Worker> o;
    for (auto& w : ws) { orc::Worker x; x.url = w->url(); x.model_id = w->model_id(); x.load = w->load(); x.healthy = w->is_healthy(); x.circuit_ok = w->circuit_breaker_can_execute(); o.push_back(x); }
    return o;
}

static void test_event_stream_matches_oracle() {
    const uint32_t W = 24, BS = 16, CHAINS = 40;
    std::mt19937_64 rng(20260921);
    CacheAwareConfig c = test_config(); c.block_size = BS;
    CacheAwarePolicy policy(c, g_device);
    orc::CacheAwareConfig oc; oc.cache_threshold = c.cache_threshold; oc.balance_abs_threshold = c.balance_abs_threshold; oc.balance_rel_threshold = c.balance_rel_threshold;
    oc.eviction_interval_secs = 0; oc.max_tree_size = c.max_tree_size; oc.block_size = BS;
    orc::CacheAwarePolicy opol(oc);
    Workers ws;
    for (uint32_t i = 0; i < W; ++i) ws.push_back(std::make_shared<BasicWorker>("http://w" + std::to_string(i) + ":8000", "m"));
    policy.init_workers(ws);
    opol.init_workers(mirror(ws));
    auto monitor = policy.kv_event_monitor();
    auto ix = monitor->create_indexer("m", 8);
    monitor->set_block_size("m", BS);
    orc::PositionalIndexer oix(8);
    std::vector<orc::WorkerBlockMap> owb(W);
    opol.set_monitor(true); opol.attach_indexer("m", &oix); opol.set_block_size("m", BS);

    // CHAINS prompts of 4..40 blocks; every worker caches a random prefix of a few of them
    std::vector<std::vector<uint32_t>> chains(CHAINS);
    for (auto& ch : chains) { ch.resize((4 + rng() % 37) * BS); for (auto& t : ch) t = (uint32_t)(rng() % 50000); }
    uint64_t next_seq = 1;
    for (uint32_t w = 0; w < W; ++w) {
        const uint32_t wid = ix->intern_worker(ws[w]->url());
        const uint32_t owid = oix.intern_worker(ws[w]->url());
        CHECK_EQ(wid, owid);
        for (int k = 0; k < 5; ++k) {
            const auto& ch = chains[rng() % CHAINS];
            const size_t nb = 1 + rng() % (ch.size() / BS);
            std::vector<uint64_t> seq(nb), con(nb);
            for (size_t b = 0; b < nb; ++b) { seq[b] = next_seq++; con[b] = orc::compute_content_hash(ch.data() + b * BS, BS); }
            ix->apply_stored_tokens(wid, seq, std::vector<uint32_t>(ch.begin(), ch.begin() + nb * BS), BS);
            CHECK_EQ(oix.apply_stored(owid, seq.data(), con.data(), nb, false, 0, owb[owid]), orc::APPLY_OK);
        }
    }
    policy.set_kv_event_monitor(monitor);
    CHECK_EQ(ix->current_size(), oix.current_size());

    for (int round = 0; round < 6; ++round) {
        for (uint32_t w = 0; w < W; ++w) { bw(ws, w).set_load(rng() % (round < 3 ? 20 : 60)); bw(ws, w).set_healthy(rng() % 7 != 0); bw(ws, w).set_circuit_ok(rng() % 11 != 0); }
        std::vector<std::vector<uint32_t>> reqs;
        for (int r = 0; r < 200; ++r) {
            const auto& ch = chains[rng() % CHAINS];
            size_t keep = rng() % (ch.size() + 1);
            std::vector<uint32_t> q(ch.begin(), ch.begin() + keep);
            const size_t tail = rng() % 48;
            for (size_t i = 0; i < tail; ++i) q.push_back((uint32_t)(rng() % 50000));
            reqs.push_back(std::move(q));
        }
        auto got = policy.select_worker_batch(ws, reqs);
        auto ows = mirror(ws);
        for (size_t r = 0; r < reqs.size(); ++r) {
            orc::Decision d = opol.select_worker(ows, nullptr, reqs[r].data(), reqs[r].size(), true);
            CHECK_EQ(got[r].idx, d.idx);
            CHECK_EQ(got[r].info.branch, d.branch);
            if (d.branch == orc::BR_EVENT_OVERLAP) CHECK_EQ(got[r].info.matched, d.score);
        }
    }
}

static void test_tree_streams_match_oracle() {
    const uint32_t W = 8;
    std::mt19937_64 rng(77);
    CacheAwareConfig c = test_config(); c.block_size = 16;
    CacheAwarePolicy policy(c, g_device);
    orc::CacheAwareConfig oc; oc.eviction_interval_secs = 0; oc.block_size = 16;
    orc::tree_globals() = orc::TreeGlobals();   // the reference's logical clocks are process-global statics; a fresh policy starts them at 0
    orc::CacheAwarePolicy opol(oc);
    Workers ws;
    for (uint32_t i = 0; i < W; ++i) ws.push_back(std::make_shared<BasicWorker>("http://w" + std::to_string(i) + ":8000"));
    policy.init_workers(ws);
    opol.init_workers(mirror(ws));
    std::vector<std::vector<uint32_t>> prompts(12);
    for (auto& p : prompts) { p.resize(16 * (2 + rng() % 10)); for (auto& t : p) t = (uint32_t)(rng() % 1000); }
    const char* words[] = {"route ", "the ", "request ", "to ", "a ", "worker ", "with ", "cached ", "préfixe ", "κλειδί ", "缓存 "};
    for (int round = 0; round < 5; ++round) {
        for (uint32_t w = 0; w < W; ++w) bw(ws, w).set_load(rng() % 12);
        std::vector<std::vector<uint32_t>> reqs;
        std::vector<std::string> texts;
        for (int r = 0; r < 64; ++r) {
            const auto& p = prompts[rng() % prompts.size()];
            std::vector<uint32_t> q(p.begin(), p.begin() + rng() % (p.size() + 1));
            for (size_t i = rng() % 20; i > 0; --i) q.push_back((uint32_t)(rng() % 1000));
            reqs.push_back(std::move(q));
            std::string s;
            for (size_t i = rng() % 12; i > 0; --i) s += words[rng() % (i > 6 ? 4 : 11)];
            texts.push_back(std::move(s));
        }
        auto ows = mirror(ws);
        auto got = policy.select_worker_batch(ws, reqs);
        for (size_t r = 0; r < reqs.size(); ++r) {
            orc::Decision d = opol.select_worker(ows, nullptr, reqs[r].data(), reqs[r].size(), true);
            CHECK_EQ(got[r].idx, d.idx);
            CHECK_EQ(got[r].info.branch, d.branch);
            CHECK_EQ(got[r].info.matched, d.matched);
        }
        auto gtxt = policy.select_worker_batch_request_text(ws, texts);
        for (size_t r = 0; r < texts.size(); ++r) {
            orc::Decision d = opol.select_worker(ows, &texts[r], nullptr, 0, false);
            CHECK_EQ(gtxt[r].idx, d.idx);
            CHECK_EQ(gtxt[r].info.branch, d.branch);
            CHECK_EQ(gtxt[r].info.matched, d.matched);
        }
    }
}

// ---- prefix_hash (policies/prefix_hash.rs tests :236-414) -------------------------------------------------------------------------
static Workers three_workers() {
    return Workers{std::make_shared<BasicWorker>("http://w1:8000"), std::make_shared<BasicWorker>("http://w2:8000"), std::make_shared<BasicWorker>("http://w3:8000")};
}
static const std::vector<std::string> kW3 = {"http://w1:8000", "http://w2:8000", "http://w3:8000"};

static void test_prefix_hash_consistent_routing() {   // prefix_hash.rs:262-284
    smgx::PrefixHashPolicy policy({}, g_device);
    Workers ws = three_workers();
    SelectWorkerInfo info = toks({1, 2, 3, 4, 5, 6, 7, 8, 9, 10});
    info.hash_ring = policy.hash_ring(kW3);
    auto first = policy.select_worker(ws, info);
    CHECK(first.has_value());
    for (int k = 0; k < 10; ++k) CHECK(policy.select_worker(ws, info) == first);
    CHECK(std::string(policy.name()) == "prefix_hash");
    CHECK(!policy.needs_request_text());
}

static void test_prefix_hash_no_tokens_and_no_healthy() {   // prefix_hash.rs:341-386
    smgx::PrefixHashPolicy policy({}, g_device);
    Workers ws{std::make_shared<BasicWorker>("http://w1:8000")};
    auto ring = policy.hash_ring({"http://w1:8000"});
    auto d = policy.select_worker_batch(ws, {std::vector<uint32_t>{}, std::nullopt}, ring);
    CHECK_EQ(d[0].idx, -1); CHECK_EQ(d[0].info.branch, SMGX_PH_NO_TOKENS);
    CHECK_EQ(d[1].idx, -1); CHECK_EQ(d[1].info.branch, SMGX_PH_NO_TOKENS);
    bw(ws, 0).set_healthy(false);
    d = policy.select_worker_batch(ws, {std::vector<uint32_t>{1, 2, 3}}, ring);
    CHECK_EQ(d[0].idx, -1); CHECK_EQ(d[0].info.branch, SMGX_PH_NO_HEALTHY_WORKERS);
}

static void test_hash_ring_surface() {   // hash_ring.rs:152-198
    smgx::PrefixHashPolicy policy({}, g_device);
    auto empty = policy.hash_ring({}, "empty");
    CHECK(empty->is_empty());
    CHECK(!empty->find_healthy_url("any-key", [](const std::string&) { return true; }).has_value());
    auto ring = policy.hash_ring({"http://a", "http://b", "http://c"}, "abc");
    CHECK_EQ(ring->len(), 450);
    CHECK_EQ(ring->worker_count(), 3);
    orc::HashRing oring({"http://a", "http://b", "http://c"});
    for (int k = 0; k < 50; ++k) {
        const std::string key = "routing-key-" + std::to_string(k);
        auto got = ring->find_healthy_url(key, [](const std::string& u) { return u != "http://a"; });
        int64_t want = oring.find_healthy(key, [](const std::string& u) { return u != "http://a"; });
        CHECK(got.has_value() && want >= 0 && *got == oring.url((size_t)want));
    }
    CHECK(!ring->find_healthy_url("k", [](const std::string&) { return false; }).has_value());
}

static void test_prefix_hash_stream_matches_oracle() {
    const uint32_t W = 40;
    std::mt19937_64 rng(4242);
    smgx::PrefixHashConfig cfg; cfg.prefix_token_count = 128; cfg.load_factor = 1.1;
    smgx::PrefixHashPolicy policy(cfg, g_device);
    orc::PrefixHashConfig ocfg; ocfg.prefix_token_count = 128; ocfg.load_factor = 1.1;
    orc::PrefixHashPolicy opol(ocfg);
    Workers ws;
    std::vector<std::string> urls;
    for (uint32_t i = 0; i < W; ++i) { urls.push_back("http://w" + std::to_string(i) + ":8000"); ws.push_back(std::make_shared<BasicWorker>(urls.back(), "m")); }
    auto ring = policy.hash_ring(urls, "m");
    orc::HashRing oring(urls);
    for (int round = 0; round < 4; ++round) {
        std::vector<orc::PrefixWorker> ows(W);
        for (uint32_t w = 0; w < W; ++w) {
            bw(ws, w).set_load(rng() % (round == 0 ? 1 : 25)); bw(ws, w).set_healthy(rng() % 5 != 0);
            ows[w].url = urls[w]; ows[w].load = ws[w]->load(); ows[w].healthy = ws[w]->is_healthy();
        }
        std::vector<std::optional<std::vector<uint32_t>>> reqs;
        for (int r = 0; r < 300; ++r) {
            std::vector<uint32_t> t(rng() % 400);
            for (auto& x : t) x = (uint32_t)(rng() % 128000);
            if (r % 50 == 7) reqs.push_back(std::nullopt); else reqs.push_back(std::move(t));
        }
        auto got = policy.select_worker_batch(ws, reqs, round == 3 ? nullptr : ring);
        for (size_t r = 0; r < reqs.size(); ++r) {
            orc::PrefixBranch br;
            const int64_t want = opol.select_worker(ows, reqs[r] ? reqs[r]->data() : nullptr, reqs[r] ? reqs[r]->size() : 0, round == 3 ? nullptr : &oring, &br);
            CHECK_EQ(got[r].idx, want);
            CHECK_EQ(got[r].info.branch, br);
        }
    }
}

// ---- host-only subset -----------------------------------------------------------------------------------------------------------------
static void test_policy_surface() {   // cache_aware.rs:704-710, mod.rs:106-117
    CacheAwareConfig d;
    CHECK(d.cache_threshold == 0.5f && d.balance_abs_threshold == 32 && d.balance_rel_threshold == 1.1f);
    CHECK(d.eviction_interval_secs == 30 && d.max_tree_size == 10000 && d.block_size == 16);
    CacheAwarePolicy policy(test_config(), -1);
    CHECK(std::string(policy.name()) == "cache_aware");
    CHECK(policy.needs_request_text());
    CHECK(smgx::normalize_model_key("") == "unknown");
    CHECK(smgx::normalize_model_key("llama") == "llama");
    policy.on_request_complete("http://w1:8000", true);
}

static void test_indexer_writers_and_apply_errors() {   // event_tree.rs tests: apply_stored / removed / cleared and ApplyError
    CacheAwarePolicy policy(test_config(), -1);
    auto monitor = policy.kv_event_monitor();
    bool threw = false;
    try { monitor->create_indexer("m", 0); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);   // "jump_size must be greater than 0" (event_tree.rs:280)
    auto ix = monitor->create_indexer("m", 4);
    CHECK(!ix->worker_id("http://w1:8000").has_value());
    const uint32_t w1 = ix->intern_worker("http://w1:8000");
    CHECK_EQ(ix->intern_worker("http://w1:8000"), w1);
    CHECK_EQ(*ix->worker_id("http://w1:8000"), w1);
    ix->apply_stored(w1, {{1, 100}, {2, 200}, {3, 300}});
    CHECK_EQ(ix->current_size(), 3);
    ix->apply_stored(w1, {{4, 400}}, 3);   // continuation under seq hash 3
    CHECK_EQ(ix->current_size(), 4);
    int code = 0;
    try { ix->apply_stored(w1, {{9, 900}}, 777); } catch (const smgx::ApplyError& e) { code = e.code; }
    CHECK_EQ(code, SMGX_PARENT_BLOCK_NOT_FOUND);
    const uint32_t w2 = ix->intern_worker("http://w2:8000");
    code = 0;
    try { ix->apply_stored(w2, {{9, 900}}, 1); } catch (const smgx::ApplyError& e) { code = e.code; }
    CHECK_EQ(code, SMGX_WORKER_NOT_TRACKED);
    ix->apply_removed(w1, {4, 12345});
    CHECK_EQ(ix->current_size(), 3);
    ix->apply_cleared(w1);
    CHECK_EQ(ix->current_size(), 0);
    // Removed / Cleared events need no hashing and run on the host mirror; Stored events hash token ids on the GPU
    ix->apply_stored(w2, {{11, 1100}, {12, 1200}});
    smgx::KvCacheEvent rm; rm.kind = smgx::KvCacheEvent::Removed; rm.block_hashes = {12};
    smgx::KvCacheEvent cl; cl.kind = smgx::KvCacheEvent::Cleared;
    CHECK_EQ(monitor->apply_events("m", w2, {rm}), 0);
    CHECK_EQ(ix->current_size(), 1);
    monitor->apply_events("m", w2, {cl});
    CHECK_EQ(ix->current_size(), 0);
    smgx::KvCacheEvent st;
    st.blocks.push_back({11, {1, 2, 3, 4}, 4});
    code = 0;
    try { monitor->apply_events("m", w2, {st}); } catch (const smgx::Error& e) { code = e.code; }
    CHECK_EQ(code, SMGX_DEVICE_ERROR);
}

static void test_string_tree_snapshot_wire_format() {   // snapshot.rs:66-109, string_tree.rs:2804-2935 — host-authoritative tree, no GPU needed
    orc::tree_globals() = orc::TreeGlobals();
    CacheAwarePolicy pa(test_config(), -1), pb(test_config(), -1);
    auto ta = pa.string_tree(), tb = pb.string_tree();
    orc::StringTree oa, ob;
    const char* texts[] = {"Hello world", "Hello there", "Goodbye", "héllo wörld", "Hell", "日本語のテキスト", "日本"};
    int k = 0;
    for (const char* t : texts) { const std::string w = "worker-" + std::to_string(k++ % 3); ta->insert_text(t, w); oa.insert_text(t, w); }
    const std::string snap = ta->snapshot_bytes();
    CHECK(snap == orc::StringTree::snapshot_to_bytes(oa.snapshot()));
    orc::tree_globals() = orc::TreeGlobals();
    for (const char* t : {"Hello wonder", "Good", "zulu", "日本酒"}) { tb->insert_text(t, "worker-9"); ob.insert_text(t, "worker-9"); }
    const std::string remote = tb->snapshot_bytes();
    CHECK(remote == orc::StringTree::snapshot_to_bytes(ob.snapshot()));
    ta->merge_snapshot(remote);                                  // Tree::merge_snapshot
    orc::StringTree::TreeSnapshot rs;
    CHECK(orc::StringTree::snapshot_from_bytes(remote, rs));
    oa.merge_snapshot(rs);
    CHECK(ta->snapshot_bytes() == orc::StringTree::snapshot_to_bytes(oa.snapshot()));
    CHECK_EQ(ta->node_count(), oa.node_count());
    CacheAwarePolicy pc(test_config(), -1);
    auto tc = pc.string_tree();
    tc->load_snapshot(ta->snapshot_bytes());                     // Tree::from_snapshot
    CHECK(tc->snapshot_bytes() == ta->snapshot_bytes());
    int code = 0;
    try { tc->load_snapshot(remote.substr(0, remote.size() - 3)); } catch (const smgx::Error& e) { code = e.code; }
    CHECK_EQ(code, SMGX_INVALID_ARGUMENT);
    CHECK(tc->snapshot_bytes() == ta->snapshot_bytes());         // a malformed snapshot leaves the tree untouched
}

static void test_select_without_device_fails_loudly() {   // there is no CPU fallback behind the interface
    CacheAwarePolicy policy(test_config(), -1);
    Workers ws = two_workers();
    policy.init_workers(ws);
    int code = 0;
    try { policy.select_worker(ws, text("hello")); } catch (const smgx::Error& e) { code = e.code; }
    CHECK_EQ(code, SMGX_DEVICE_ERROR);
    code = 0;
    try { policy.select_worker(ws, toks({1, 2, 3, 4})); } catch (const smgx::Error& e) { code = e.code; }
    CHECK_EQ(code, SMGX_DEVICE_ERROR);
}

int main(int argc, char** argv) {
    const bool host_only = argc > 1 && std::string(argv[1]) == "--host";
    struct T { const char* name; void (*fn)(); bool gpu; };
    const T tests[] = {
        {"policy_surface", test_policy_surface, false},
        {"indexer_writers_and_apply_errors", test_indexer_writers_and_apply_errors, false},
        {"select_without_device_fails_loudly", test_select_without_device_fails_loudly, false},
        {"string_tree_snapshot_wire_format", test_string_tree_snapshot_wire_format, false},
        {"cache_aware_with_balanced_load", test_cache_aware_with_balanced_load, true},
        {"cache_aware_with_imbalanced_load", test_cache_aware_with_imbalanced_load, true},
        {"cache_aware_worker_removal", test_cache_aware_worker_removal, true},
        {"no_healthy_returns_none", test_no_healthy_returns_none, true},
        {"event_driven_overlap_selects_cached_worker", test_event_driven_overlap_selects_cached_worker, true},
        {"event_driven_no_overlap_uses_min_load", test_event_driven_no_overlap_uses_min_load, true},
        {"event_driven_short_request_uses_min_load", test_event_driven_short_request_uses_min_load, true},
        {"no_monitor_uses_token_tree", test_no_monitor_uses_token_tree, true},
        {"imbalanced_skips_event_driven", test_imbalanced_skips_event_driven, true},
        {"empty_indexer_falls_through_to_token_tree", test_empty_indexer_falls_through_to_token_tree, true},
        {"kv_events_fresh_chain_fallback", test_kv_events_fresh_chain_fallback, true},
        {"event_stream_matches_oracle", test_event_stream_matches_oracle, true},
        {"tree_streams_match_oracle", test_tree_streams_match_oracle, true},
        {"prefix_hash_consistent_routing", test_prefix_hash_consistent_routing, true},
        {"prefix_hash_no_tokens_and_no_healthy", test_prefix_hash_no_tokens_and_no_healthy, true},
        {"hash_ring_surface", test_hash_ring_surface, true},
        {"prefix_hash_stream_matches_oracle", test_prefix_hash_stream_matches_oracle, true},
    };
    int ran = 0;
    for (const T& t : tests) {
        if (host_only && t.gpu) continue;
        const int before = g_fail;
        try { t.fn(); } catch (const std::exception& e) { ++g_fail; std::fprintf(stderr, "FAIL %s threw: %s\n", t.name, e.what()); }
        std::printf("%s %s\n", g_fail == before ? "ok  " : "FAIL", t.name);
        ++ran;
    }
    std::printf("%d tests, %d checks, %d failures\n", ran, g_checks, g_fail);
    return g_fail ? 1 : 0;
}
