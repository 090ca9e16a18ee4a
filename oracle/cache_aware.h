/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into / imported by the product path.
 *
 * CPU restatement of the reference's cache-aware worker pick, following
 *   model_gateway/src/policies/cache_aware.rs  (reference @ 1c5701cf)
 *     :219-283 init_workers / add_worker / add_worker_by_url   :285-308 remove_worker (no-op)
 *     :311-352 evict_cache                                      :356-440 select_worker_min_load
 *     :648-690 select_worker (healthy filter, min/max load over ALL workers, f32 imbalance gate, mode select)
 *     :723-729 has_event_indexer   :736-769 select_worker_event_driven   :776-831 score_overlap
 *     :834-904 select_worker_with_tokens   :907-974 select_worker_with_text
 *   model_gateway/src/policies/mod.rs :94-117 CacheAwareConfig (+Default) :137-144 get_healthy_worker_indices
 *                                     :151-157 normalize_model_key ("" → "unknown")
 * Rust std semantics honoured: min_by_key → FIRST minimum, max_by_key → LAST maximum, `usize as f32`
 * round-to-nearest-even, saturating_sub.
 * Pinned by the reference's unit tests ported in tests/test_oracle_cache_aware.py (cache_aware.rs:999-2015,
 * mod.rs:192-262).
 *
 * The mesh hash_index side effect (cache_aware.rs:397-401, 420-424, 881-886, 950-956) — blake3 path hash of the FULL request
 * → copy of the matched prefix — is restated too (hash_index_tokens / hash_index_text); it never feeds back into the pick.
 * The "no tree for model → rand" branch (:896-903, :965-973) returns healthy[0] and
 * reports the whole healthy set as valid.
 *
 * "Snapshot batches" (begin_snapshot_batch / end_snapshot_batch): select_worker takes &self and is called from many
 * tokio tasks at once; it is not atomic — a task's read-only walk, its match side effects and its insert can interleave
 * with other tasks' (the reference promises eventual consistency only, token_tree.rs:1035-1037).  One admissible
 * interleaving of a batch of concurrent calls is: every call performs its read-only walk + decision against the same
 * tree state, then each call, in request order, applies its match side effects and its insert.  Between
 * begin/end the policy executes exactly that interleaving (side effects are queued and replayed by end_snapshot_batch).
 * With no intra-batch conflicts it is identical, timestamps included, to calling select_worker one by one.
 */
#pragma once
#include <functional>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "blake3_ref.h"
#include "positional_indexer.h"
#include "string_tree.h"
#include "token_tree.h"

namespace orc {

struct CacheAwareConfig {  // mod.rs:94-117
    float cache_threshold = 0.5f;
    uint64_t balance_abs_threshold = 32;
    float balance_rel_threshold = 1.1f;
    uint64_t eviction_interval_secs = 30;
    uint64_t max_tree_size = 10000;
    uint64_t block_size = 16;
};

struct Worker {  // the scalars the path reads from `trait Worker` (worker/worker.rs:114,151-153,187,208,217)
    std::string url;
    std::string model_id;
    uint64_t load = 0;
    bool healthy = true;          // status == Ready
    bool circuit_ok = true;       // circuit_breaker_can_execute()
    uint64_t processed = 0;       // increment_processed()
};

enum Branch {
    BR_NO_HEALTHY = 0,
    BR_IMBALANCED_MIN_LOAD = 1,
    BR_EVENT_OVERLAP = 2,
    BR_EVENT_MIN_LOAD = 3,
    BR_TREE_MATCH = 4,
    BR_TREE_MIN_LOAD = 5,
    BR_TREE_FALLBACK_FIRST_HEALTHY = 6,
    BR_NO_TREE_RANDOM = 7,
};

struct Decision {
    int64_t idx = -1;  // index into the passed worker slice; -1 == None
    Branch branch = BR_NO_HEALTHY;
    size_t matched = 0, input = 0;
    uint32_t score = 0;            // event mode: overlap score of the winner
    std::vector<int64_t> valid;    // every idx the reference could legally return for this state
};

static inline const char* normalize_model_key(const std::string& m) { return m.empty() ? "unknown" : m.c_str(); }

class CacheAwarePolicy {
public:
    explicit CacheAwarePolicy(const CacheAwareConfig& c) : cfg_(c) {}

    // cache_aware.rs:219-249
    void init_workers(const std::vector<Worker>& ws) {
        for (auto& w : ws) add_worker_by_url(w.url, normalize_model_key(w.model_id));
    }
    void add_worker_by_url(const std::string& url, const std::string& model) {  // :269-283
        string_tree(model, true)->insert_text("", url);
        token_tree(model, true)->insert_tokens(nullptr, 0, url);
    }
    void evict_cache(size_t max_size) {  // :311-331
        for (auto& kv : string_trees_) kv.second->evict_tenant_by_size(max_size);
        for (auto& kv : token_trees_) kv.second->evict_tenant_by_size(max_size);
        for (auto& kv : hash_index_text_) if (kv.second.size() > max_size) kv.second.clear();      // :335-351
        for (auto& kv : hash_index_tokens_) if (kv.second.size() > max_size) kv.second.clear();
    }

    // KvEventMonitor surface the policy reads (worker/kv_event_monitor.rs get_indexer / block_size)
    void set_monitor(bool present) { monitor_ = present; }
    void attach_indexer(const std::string& model, PositionalIndexer* ix) { indexers_[model] = ix; }
    void set_block_size(const std::string& model, size_t bs) { learned_bs_[model] = bs; }

    StringTree* string_tree(const std::string& model, bool create = false) {
        auto it = string_trees_.find(model);
        if (it != string_trees_.end()) return it->second.get();
        if (!create) return nullptr;
        return (string_trees_[model] = std::make_unique<StringTree>()).get();
    }
    TokenTree* token_tree(const std::string& model, bool create = false) {
        auto it = token_trees_.find(model);
        if (it != token_trees_.end()) return it->second.get();
        if (!create) return nullptr;
        return (token_trees_[model] = std::make_unique<TokenTree>()).get();
    }

    // cache_aware.rs:648-690.  `text`/`tokens` may be null (Option::None).
    Decision select_worker(std::vector<Worker>& ws, const std::string* text, const uint32_t* tokens, size_t n_tokens,
                           bool has_tokens) {
        Decision d;
        std::vector<size_t> healthy;
        for (size_t i = 0; i < ws.size(); ++i) if (ws[i].healthy && ws[i].circuit_ok) healthy.push_back(i);
        if (healthy.empty()) return d;
        std::string model = normalize_model_key(ws[healthy[0]].model_id);

        uint64_t mn = UINT64_MAX, mx = 0;
        for (auto& w : ws) { mn = std::min(mn, w.load); mx = std::max(mx, w.load); }
        if (mn == UINT64_MAX) mn = 0;
        uint64_t diff = mx >= mn ? mx - mn : 0;
        volatile float fmax = (float)mx;
        volatile float fprod = (float)mn * cfg_.balance_rel_threshold;
        bool imbalanced = diff > cfg_.balance_abs_threshold && fmax > fprod;

        if (imbalanced) return min_load_path(ws, text, tokens, n_tokens, has_tokens, healthy, model);
        if (has_tokens) {
            if (has_event_indexer(model)) return event_driven(ws, tokens, n_tokens, healthy, model);
            return with_tokens(ws, tokens, n_tokens, healthy, model);
        }
        static const std::string empty;
        return with_text(ws, text ? *text : empty, healthy, model);
    }

    void begin_snapshot_batch() { deferring_ = true; deferred_.clear(); }
    void end_snapshot_batch() {
        deferring_ = false;
        for (auto& f : deferred_) f();
        deferred_.clear();
    }

    const std::map<uint64_t, std::vector<uint32_t>>& hash_index_tokens(const std::string& model) { return hash_index_tokens_[model]; }
    const std::map<uint64_t, std::string>& hash_index_text(const std::string& model) { return hash_index_text_[model]; }

    // ---- TreeHandle (cache_aware.rs:443-645): what the mesh adapter calls to apply remote tenant inserts and repair pages ----
    enum TreeKind { TREE_STRING = 0, TREE_TOKEN = 1 };   // :445-448
    // :496-553: true iff node_hash is in hash_index[model] of that kind AND the tree exists; then the stored prefix is inserted for worker_url
    bool apply_known_remote_insert(const std::string& model_in, TreeKind kind, uint64_t node_hash, const std::string& worker_url) {
        const std::string model = normalize_model_key(model_in);
        if (kind == TREE_STRING) {
            auto mi = hash_index_text_.find(model);
            if (mi == hash_index_text_.end() && hash_index_tokens_.find(model) == hash_index_tokens_.end()) return false;   // hash_index.get(model_id)?
            if (mi == hash_index_text_.end()) return false;
            auto it = mi->second.find(node_hash);
            if (it == mi->second.end()) return false;
            StringTree* t = string_tree(model);
            if (!t) return false;   // populate-site invariant violated (:512-526)
            t->insert_text(it->second, worker_url);
            return true;
        }
        auto mi = hash_index_tokens_.find(model);
        if (mi == hash_index_tokens_.end()) return false;
        auto it = mi->second.find(node_hash);
        if (it == mi->second.end()) return false;
        TokenTree* t = token_tree(model);
        if (!t) return false;
        t->insert_tokens(it->second.data(), it->second.size(), worker_url);
        return true;
    }
    // :575-645, one entry of a page whose variant matches the page kind: every listed tenant is inserted (epochs are ignored) and the
    // path is recorded under its blake3 path hash; the tree is created on first use
    void apply_repair_entry_text(const std::string& model_in, const std::string& path, const std::vector<std::string>& tenants) {
        const std::string model = normalize_model_key(model_in);
        StringTree* t = string_tree(model, true);
        for (auto& tenant : tenants) t->insert_text(path, tenant);
        hash_index_text_[model][hash_node_path(path)] = path;
    }
    void apply_repair_entry_tokens(const std::string& model_in, const uint32_t* tokens, size_t n, const std::vector<std::string>& tenants) {
        const std::string model = normalize_model_key(model_in);
        TokenTree* t = token_tree(model, true);
        for (auto& tenant : tenants) t->insert_tokens(tokens, n, tenant);
        hash_index_tokens_[model][hash_token_path(tokens, n)] = std::vector<uint32_t>(tokens, tokens + n);
    }
    // open_repair_stream (:555-573) is iter_entries of the model's tree: string_tree(model)->entries() / token_tree(model)->entries()

    bool has_event_indexer(const std::string& model) const {  // :723-729
        if (!monitor_) return false;
        auto it = indexers_.find(model);
        return it != indexers_.end() && it->second->current_size() > 0;
    }

    // cache_aware.rs:776-831
    static int64_t score_overlap(std::vector<Worker>& ws, const uint32_t* tokens, size_t n, const std::vector<size_t>& healthy,
                                 const PositionalIndexer& ix, size_t bs, uint32_t* score_out = nullptr) {
        auto hashes = compute_request_content_hashes(tokens, n, bs);
        if (hashes.empty()) return -1;
        OverlapScores ov = ix.find_matches(hashes, false);
        if (ov.scores.empty()) return -1;
        bool have = false;
        size_t best = 0; uint32_t bscore = 0; uint64_t bload = 0; size_t bts = 0;
        for (size_t idx : healthy) {
            auto wid = ix.worker_id(ws[idx].url);
            uint32_t score = 0; size_t ts = 0;
            if (wid) {
                auto s = ov.scores.find(*wid); if (s != ov.scores.end()) score = s->second;
                auto t = ov.tree_sizes.find(*wid); if (t != ov.tree_sizes.end()) ts = t->second;
            }
            if (score == 0) continue;
            uint64_t load = ws[idx].load;
            // key = (score, Reverse(load), Reverse(tree_size)); max_by_key keeps the LAST maximum → `>=`
            bool ge = !have || score > bscore || (score == bscore && (load < bload || (load == bload && ts <= bts)));
            if (ge) { have = true; best = idx; bscore = score; bload = load; bts = ts; }
        }
        if (!have) return -1;
        ws[best].processed++;
        if (score_out) *score_out = bscore;
        return (int64_t)best;
    }

private:
    static size_t first_min_load(const std::vector<Worker>& ws, const std::vector<size_t>& healthy) {
        size_t best = healthy[0];
        for (size_t i : healthy) if (ws[i].load < ws[best].load) best = i;  // strict < keeps the FIRST minimum
        return best;
    }

    // cache_aware.rs:356-440
    Decision min_load_path(std::vector<Worker>& ws, const std::string* text, const uint32_t* tokens, size_t n, bool has_tokens,
                           const std::vector<size_t>& healthy, const std::string& model) {
        Decision d;
        size_t idx = first_min_load(ws, healthy);
        if (has_tokens) {
            if (TokenTree* t = token_tree(model)) {
                TokenMatch m = match_tokens(t, tokens, n);
                d.matched = m.matched; d.input = m.input;
                std::string url = ws[idx].url;
                run_or_defer([t, tokens, n, url] { t->insert_tokens(tokens, n, url); });
                hash_index_tokens_[model][hash_token_path(tokens, n)] = std::vector<uint32_t>(tokens, tokens + m.matched);   // :397-401
            }
        } else if (text) {
            if (StringTree* t = string_tree(model)) {
                StringMatch m = match_text(t, *text);
                d.matched = m.matched; d.input = m.input;
                std::string url = ws[idx].url, tx = *text;
                run_or_defer([t, tx, url] { t->insert_text(tx, url); });
                hash_index_text_[model][hash_node_path(tx)] = utf8_encode(utf8_decode(tx).substr(0, m.matched));               // :420-424
            }
        }
        ws[idx].processed++;
        d.idx = (int64_t)idx; d.branch = BR_IMBALANCED_MIN_LOAD; d.valid = {d.idx};
        return d;
    }

    // cache_aware.rs:736-769
    Decision event_driven(std::vector<Worker>& ws, const uint32_t* tokens, size_t n, const std::vector<size_t>& healthy,
                          const std::string& model) {
        Decision d;
        PositionalIndexer* ix = indexers_[model];
        size_t bs = learned_bs_.count(model) ? learned_bs_[model] : (size_t)cfg_.block_size;
        d.input = n;
        int64_t idx = score_overlap(ws, tokens, n, healthy, *ix, bs, &d.score);
        if (idx >= 0) { d.idx = idx; d.branch = BR_EVENT_OVERLAP; d.matched = (size_t)d.score * bs; d.valid = {idx}; return d; }
        size_t m = first_min_load(ws, healthy);
        ws[m].processed++;
        d.idx = (int64_t)m; d.branch = BR_EVENT_MIN_LOAD; d.valid = {d.idx};
        return d;
    }

    template <class Tree, class Match, class InsertFn>
    Decision tree_decide(std::vector<Worker>& ws, const Match& m, const std::vector<size_t>& healthy, InsertFn insert) {
        Decision d;
        d.matched = m.matched; d.input = m.input;
        volatile float rate = m.input == 0 ? 0.0f : (float)m.matched / (float)m.input;
        int64_t sel = -1;
        if (rate > cfg_.cache_threshold) {
            auto resolve = [&](const std::string& tenant) -> int64_t {
                for (size_t i = 0; i < ws.size(); ++i)
                    if (ws[i].url == tenant) return ws[i].healthy ? (int64_t)i : -1;  // .filter(is_healthy) only
                return -1;
            };
            sel = resolve(m.tenant);
            d.branch = BR_TREE_MATCH;
            for (auto& t : m.valid) {
                int64_t v = resolve(t);
                if (v < 0) v = (int64_t)healthy[0];
                if (std::find(d.valid.begin(), d.valid.end(), v) == d.valid.end()) d.valid.push_back(v);
            }
        } else {
            sel = (int64_t)first_min_load(ws, healthy);
            d.branch = BR_TREE_MIN_LOAD;
            d.valid = {sel};
        }
        if (sel >= 0) {
            insert(ws[(size_t)sel].url);
            ws[(size_t)sel].processed++;
            d.idx = sel;
            return d;
        }
        d.idx = (int64_t)healthy[0];  // :892-894 — no insert, no processed increment
        d.branch = BR_TREE_FALLBACK_FIRST_HEALTHY;
        if (d.valid.empty()) d.valid = {d.idx};
        return d;
    }

    // cache_aware.rs:834-904
    Decision with_tokens(std::vector<Worker>& ws, const uint32_t* tokens, size_t n, const std::vector<size_t>& healthy,
                         const std::string& model) {
        TokenTree* t = token_tree(model);
        if (!t) return random_healthy(healthy);
        TokenMatch m = match_tokens(t, tokens, n);
        return tree_decide<TokenTree>(ws, m, healthy, [&](const std::string& url) {
            run_or_defer([t, tokens, n, url] { t->insert_tokens(tokens, n, url); });   // `tokens` outlives the batch call
            hash_index_tokens_[model][hash_token_path(tokens, n)] = std::vector<uint32_t>(tokens, tokens + m.matched);          // :881-886
        });
    }
    // cache_aware.rs:907-974
    Decision with_text(std::vector<Worker>& ws, const std::string& text, const std::vector<size_t>& healthy,
                       const std::string& model) {
        StringTree* t = string_tree(model);
        if (!t) return random_healthy(healthy);
        StringMatch m = match_text(t, text);
        return tree_decide<StringTree>(ws, m, healthy, [&](const std::string& url) {
            std::string tx = text;
            run_or_defer([t, tx, url] { t->insert_text(tx, url); });
            hash_index_text_[model][hash_node_path(tx)] = utf8_encode(utf8_decode(tx).substr(0, m.matched));                    // :950-956
        });
    }
    // match + its side effects: immediate, or (snapshot batch) read-only now and side effects queued
    TokenMatch match_tokens(TokenTree* t, const uint32_t* tokens, size_t n) {
        if (!deferring_) return t->match_prefix_with_counts(tokens, n);
        auto touches = std::make_shared<std::vector<TokenTree::PendingTouch>>();
        TokenMatch m = t->match_prefix_with_counts(tokens, n, touches.get());
        deferred_.push_back([t, touches] { t->apply_touches(*touches); });
        return m;
    }
    StringMatch match_text(StringTree* t, const std::string& text) {
        if (!deferring_) return t->match_prefix_with_counts(text);
        auto eff = std::make_shared<StringTree::PendingEffect>();
        StringMatch m = t->match_prefix_with_counts(text, eff.get());
        deferred_.push_back([t, eff] { t->apply_match_effects(*eff); });
        return m;
    }
    void run_or_defer(std::function<void()> f) { if (deferring_) deferred_.push_back(std::move(f)); else f(); }

    static Decision random_healthy(const std::vector<size_t>& healthy) {
        Decision d;
        d.idx = (int64_t)healthy[0]; d.branch = BR_NO_TREE_RANDOM;
        for (size_t h : healthy) d.valid.push_back((int64_t)h);
        return d;
    }

    CacheAwareConfig cfg_;
    std::map<std::string, std::map<uint64_t, std::vector<uint32_t>>> hash_index_tokens_;   // hash_index[model].token_tree (cache_aware.rs:95-101)
    std::map<std::string, std::map<uint64_t, std::string>> hash_index_text_;              // hash_index[model].string_tree
    bool deferring_ = false;
    std::vector<std::function<void()>> deferred_;
    std::map<std::string, std::unique_ptr<StringTree>> string_trees_;
    std::map<std::string, std::unique_ptr<TokenTree>> token_trees_;
    bool monitor_ = false;
    std::map<std::string, PositionalIndexer*> indexers_;
    std::map<std::string, size_t> learned_bs_;
};

}  // namespace orc
