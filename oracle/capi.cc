/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into / imported by the product path.
 *
 * Flat C surface over the oracle classes so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs can drive it through ctypes.  See the class headers for reference file:line.
 */
#include "power_of_two.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "cache_aware.h"
#include "tuned_event.h"
#include "prefix_hash.h"

using namespace orc;

namespace {
struct IndexerBox {
    PositionalIndexer ix;
    std::map<uint32_t, WorkerBlockMap> wbs;  // "caller-owned" reverse maps (event_tree.rs:246), one per worker
    explicit IndexerBox(size_t j) : ix(j) {}
};
struct PolicyBox {
    CacheAwarePolicy pol;
    std::vector<Worker> workers;
    explicit PolicyBox(const CacheAwareConfig& c) : pol(c) {}
};
size_t copy_str(const std::string& s, char* out, size_t cap) {
    if (out && cap) { size_t n = std::min(cap - 1, s.size()); memcpy(out, s.data(), n); out[n] = 0; }
    return s.size();
}
std::string join(const std::vector<std::string>& v) {
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) { if (i) s.push_back('\n'); s += v[i]; }
    return s;
}
}  // namespace

extern "C" {

void orc_reset_globals() { tree_globals() = TreeGlobals(); }
uint64_t orc_token_ts() { return tree_globals().token_ts; }
uint64_t orc_string_epoch() { return tree_globals().string_epoch; }

// ---- hashing ----
uint64_t orc_xxh3_64(const void* d, size_t n, uint64_t seed) { return xxh3_64(d, n, seed); }
uint64_t orc_content_hash(const uint32_t* t, size_t n) { return compute_content_hash(t, n); }
uint64_t orc_next_seq_hash(uint64_t prev, uint64_t cur) { return compute_next_seq_hash(prev, cur); }
size_t orc_request_content_hashes(const uint32_t* t, size_t n, size_t bs, uint64_t* out, size_t cap) {
    auto v = compute_request_content_hashes(t, n, bs);
    for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    return v.size();
}

// ---- PositionalIndexer ----
void* orc_indexer_new(size_t jump) { return new IndexerBox(jump); }
void orc_indexer_free(void* h) { delete (IndexerBox*)h; }
uint32_t orc_indexer_intern_worker(void* h, const char* url) { return ((IndexerBox*)h)->ix.intern_worker(url); }
int64_t orc_indexer_worker_id(void* h, const char* url) {
    auto r = ((IndexerBox*)h)->ix.worker_id(url);
    return r ? (int64_t)*r : -1;
}
int orc_indexer_apply_stored(void* h, uint32_t wid, const uint64_t* seq, const uint64_t* content, size_t n, int has_parent,
                             uint64_t parent) {
    auto* b = (IndexerBox*)h;
    return b->ix.apply_stored(wid, seq, content, n, has_parent != 0, parent, b->wbs[wid]);
}
void orc_indexer_apply_removed(void* h, uint32_t wid, const uint64_t* seq, size_t n) {
    auto* b = (IndexerBox*)h;
    b->ix.apply_removed(wid, seq, n, b->wbs[wid]);
}
void orc_indexer_apply_cleared(void* h, uint32_t wid) {
    auto* b = (IndexerBox*)h;
    b->ix.apply_cleared(wid, b->wbs[wid]);
}
void orc_indexer_remove_worker(void* h, uint32_t wid) {
    auto* b = (IndexerBox*)h;
    b->ix.remove_worker(wid, b->wbs[wid]);
    b->wbs.erase(wid);
}
size_t orc_indexer_current_size(void* h) { return ((IndexerBox*)h)->ix.current_size(); }
size_t orc_indexer_entry_count(void* h) { return ((IndexerBox*)h)->ix.entry_count(); }
size_t orc_indexer_tree_size(void* h, uint32_t wid) { return ((IndexerBox*)h)->ix.tree_size(wid); }
size_t orc_indexer_worker_blocks(void* h, uint32_t wid) { return ((IndexerBox*)h)->wbs[wid].size(); }
size_t orc_indexer_find_matches(void* h, const uint64_t* hashes, size_t n, int early_exit, uint32_t* ids, uint32_t* scores,
                                uint64_t* tree_sizes, size_t cap) {
    std::vector<uint64_t> v(hashes, hashes + n);
    OverlapScores ov = ((IndexerBox*)h)->ix.find_matches(v, early_exit != 0);
    size_t i = 0;
    for (auto& kv : ov.scores) {
        if (i < cap) { ids[i] = kv.first; scores[i] = kv.second; tree_sizes[i] = ov.tree_sizes[kv.first]; }
        ++i;
    }
    return i;
}

// ---- TokenTree ----
void* orc_ttree_new(int policy) { return new TokenTree((EvictionPolicy)policy); }
void orc_ttree_free(void* h) { delete (TokenTree*)h; }
void orc_ttree_insert(void* h, const uint32_t* t, size_t n, const char* tenant) { ((TokenTree*)h)->insert_tokens(t, n, tenant); }
// out_counts: [matched, input, nodes_visited, edge_tokens_compared]; valid set '\n'-joined into valid_out
size_t orc_ttree_match(void* h, const uint32_t* t, size_t n, char* tenant_out, size_t cap, uint64_t* out_counts,
                       char* valid_out, size_t valid_cap) {
    TokenMatch m = ((TokenTree*)h)->match_prefix_with_counts(t, n);
    out_counts[0] = m.matched; out_counts[1] = m.input; out_counts[2] = m.nodes_visited; out_counts[3] = m.edge_tokens_compared;
    copy_str(join(m.valid), valid_out, valid_cap);
    return copy_str(m.tenant, tenant_out, cap);
}
void orc_ttree_evict_tenant(void* h, const char* tenant, size_t max_tokens) { ((TokenTree*)h)->evict_tenant(tenant, max_tokens); }
void orc_ttree_evict_by_size(void* h, size_t max_size) { ((TokenTree*)h)->evict_tenant_by_size(max_size); }
size_t orc_ttree_tenant_size(void* h, const char* tenant) { return ((TokenTree*)h)->tenant_token_size(tenant); }
size_t orc_ttree_node_count(void* h) { return ((TokenTree*)h)->node_count(); }
void orc_ttree_clear(void* h) { ((TokenTree*)h)->clear(); }
void orc_ttree_set_priority(void* h, const uint32_t* t, size_t n, int32_t p) { ((TokenTree*)h)->set_priority_on_path(t, n, p); }
// serialises iter_entries as text: one line per entry "tok,tok,...|tenant=ts;tenant=ts"
size_t orc_ttree_entries(void* h, char* out, size_t cap) {
    std::vector<std::pair<std::vector<uint32_t>, std::vector<std::pair<std::string, uint64_t>>>> es;
    ((TokenTree*)h)->entries(es);
    std::string s;
    for (auto& e : es) {
        for (size_t i = 0; i < e.first.size(); ++i) { if (i) s.push_back(','); s += std::to_string(e.first[i]); }
        s.push_back('|');
        for (size_t i = 0; i < e.second.size(); ++i) { if (i) s.push_back(';'); s += e.second[i].first + "=" + std::to_string(e.second[i].second); }
        s.push_back('\n');
    }
    return copy_str(s, out, cap);
}

// ---- StringTree ----
void* orc_stree_new() { return new StringTree(); }
void orc_stree_free(void* h) { delete (StringTree*)h; }
void orc_stree_insert(void* h, const char* text, size_t n, const char* tenant) { ((StringTree*)h)->insert_text(std::string(text, n), tenant); }
size_t orc_stree_match(void* h, const char* text, size_t n, char* tenant_out, size_t cap, uint64_t* out_counts, char* valid_out,
                       size_t valid_cap) {
    StringMatch m = ((StringTree*)h)->match_prefix_with_counts(std::string(text, n));
    out_counts[0] = m.matched; out_counts[1] = m.input; out_counts[2] = m.nodes_visited;
    copy_str(join(m.valid), valid_out, valid_cap);
    return copy_str(m.tenant, tenant_out, cap);
}
size_t orc_stree_prefix_match_tenant(void* h, const char* text, size_t n, const char* tenant, char* out, size_t cap) {
    return copy_str(((StringTree*)h)->prefix_match_tenant(std::string(text, n), tenant), out, cap);
}
int orc_stree_force_cached_tenant(void* h, const char* text, size_t n, const char* tenant) {
    return ((StringTree*)h)->force_cached_tenant(std::string(text, n), tenant) ? 1 : 0;
}
void orc_stree_evict_by_size(void* h, size_t max_size) { ((StringTree*)h)->evict_tenant_by_size(max_size); }
void orc_stree_evict_by_tenant(void* h, const char* tenant, size_t max_chars) { ((StringTree*)h)->evict_by_tenant(tenant, max_chars); }
void orc_stree_remove_tenant_all(void* h, const char* tenant) { ((StringTree*)h)->remove_tenant_all(tenant); }
size_t orc_stree_tenant_size(void* h, const char* tenant) { return ((StringTree*)h)->tenant_char_size(tenant); }
size_t orc_stree_node_count(void* h) { return ((StringTree*)h)->node_count(); }
size_t orc_stree_used_sizes(void* h, char* out, size_t cap) {
    std::string s;
    for (auto& kv : ((StringTree*)h)->used_size_per_tenant()) s += kv.first + "=" + std::to_string(kv.second) + "\n";
    return copy_str(s, out, cap);
}
size_t orc_stree_char_counts(void* h, char* out, size_t cap) {
    std::string s;
    for (auto& kv : ((StringTree*)h)->tenant_char_counts()) s += kv.first + "=" + std::to_string(kv.second) + "\n";
    return copy_str(s, out, cap);
}
size_t orc_stree_entries(void* h, char* out, size_t cap) {
    std::vector<std::pair<std::string, std::vector<std::pair<std::string, uint64_t>>>> es;
    ((StringTree*)h)->entries(es);
    std::string s;
    for (auto& e : es) {
        s += e.first; s.push_back('\x1f');
        for (size_t i = 0; i < e.second.size(); ++i) { if (i) s.push_back(';'); s += e.second[i].first + "=" + std::to_string(e.second[i].second); }
        s.push_back('\x1e');
    }
    return copy_str(s, out, cap);
}

// mesh wire format (snapshot.rs): bincode bytes of Tree::snapshot(); load = from_snapshot, merge = merge_snapshot. 0 = decode error
size_t orc_stree_snapshot_bytes(void* h, char* out, size_t cap) {
    std::string b = StringTree::snapshot_to_bytes(((StringTree*)h)->snapshot());
    if (out && cap >= b.size()) memcpy(out, b.data(), b.size());
    return b.size();
}
int orc_stree_load_snapshot_bytes(void* h, const char* bytes, size_t n) {
    StringTree::TreeSnapshot snap;
    if (!StringTree::snapshot_from_bytes(std::string(bytes, n), snap)) return 0;
    ((StringTree*)h)->load_snapshot(snap);
    return 1;
}
int orc_stree_merge_snapshot_bytes(void* h, const char* bytes, size_t n) {
    StringTree::TreeSnapshot snap;
    if (!StringTree::snapshot_from_bytes(std::string(bytes, n), snap)) return 0;
    ((StringTree*)h)->merge_snapshot(snap);
    return 1;
}

// ---- CacheAwarePolicy ----
void* orc_policy_new(float cache_threshold, uint64_t abs_thr, float rel_thr, uint64_t evict_secs, uint64_t max_tree, uint64_t block_size) {
    CacheAwareConfig c;
    c.cache_threshold = cache_threshold; c.balance_abs_threshold = abs_thr; c.balance_rel_threshold = rel_thr;
    c.eviction_interval_secs = evict_secs; c.max_tree_size = max_tree; c.block_size = block_size;
    return new PolicyBox(c);
}
void orc_policy_free(void* h) { delete (PolicyBox*)h; }
// Defines the worker slice later passed to select (urls[i], models[i]); calls init_workers when `init` != 0.
void orc_policy_set_workers(void* h, const char* const* urls, const char* const* models, size_t n, int init) {
    auto* b = (PolicyBox*)h;
    b->workers.clear();
    for (size_t i = 0; i < n; ++i) { Worker w; w.url = urls[i]; w.model_id = models ? models[i] : ""; b->workers.push_back(w); }
    if (init) b->pol.init_workers(b->workers);
}
void orc_policy_set_state(void* h, const uint64_t* loads, const uint8_t* healthy, const uint8_t* circuit_ok, size_t n) {
    auto* b = (PolicyBox*)h;
    for (size_t i = 0; i < n && i < b->workers.size(); ++i) {
        if (loads) b->workers[i].load = loads[i];
        if (healthy) b->workers[i].healthy = healthy[i] != 0;
        if (circuit_ok) b->workers[i].circuit_ok = circuit_ok[i] != 0;
    }
}
uint64_t orc_policy_processed(void* h, size_t i) { return ((PolicyBox*)h)->workers[i].processed; }
void orc_policy_set_monitor(void* h, int present) { ((PolicyBox*)h)->pol.set_monitor(present != 0); }
void orc_policy_attach_indexer(void* h, const char* model, void* indexer) {
    ((PolicyBox*)h)->pol.attach_indexer(model, &((IndexerBox*)indexer)->ix);
}
void orc_policy_set_block_size(void* h, const char* model, size_t bs) { ((PolicyBox*)h)->pol.set_block_size(model, bs); }
int orc_policy_has_event_indexer(void* h, const char* model) { return ((PolicyBox*)h)->pol.has_event_indexer(model) ? 1 : 0; }
void orc_policy_evict_cache(void* h, size_t max_size) { ((PolicyBox*)h)->pol.evict_cache(max_size); }
void* orc_policy_token_tree(void* h, const char* model) { return ((PolicyBox*)h)->pol.token_tree(model); }
void* orc_policy_string_tree(void* h, const char* model) { return ((PolicyBox*)h)->pol.string_tree(model); }

// out: [idx, branch, matched, input, score, n_valid]; valid idx list into valid_out (cap entries)
void orc_policy_select(void* h, const char* text, size_t text_len, int has_text, const uint32_t* tokens, size_t n_tokens,
                       int has_tokens, int64_t* out, int64_t* valid_out, size_t valid_cap) {
    auto* b = (PolicyBox*)h;
    std::string t;
    if (has_text) t.assign(text, text_len);
    Decision d = b->pol.select_worker(b->workers, has_text ? &t : nullptr, tokens, n_tokens, has_tokens != 0);
    out[0] = d.idx; out[1] = d.branch; out[2] = (int64_t)d.matched; out[3] = (int64_t)d.input; out[4] = d.score;
    out[5] = (int64_t)d.valid.size();
    for (size_t i = 0; i < d.valid.size() && i < valid_cap; ++i) valid_out[i] = d.valid[i];
}

// Batch of token requests against ONE fleet snapshot (the product's batch contract): ragged tokens, offsets[n+1].
// Event-driven / imbalanced / tree branches all go through select_worker in request order.
// out_idx[i], out_branch[i], out_matched[i] per request.  Returns elapsed seconds (for the cpu_baseline leg).
double orc_policy_select_batch_tokens(void* h, const uint32_t* tokens, const uint64_t* offsets, size_t n, int32_t* out_idx,
                                      uint8_t* out_branch, uint32_t* out_matched) {
    auto* b = (PolicyBox*)h;
    auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < n; ++i) {
        Decision d = b->pol.select_worker(b->workers, nullptr, tokens + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), true);
        out_idx[i] = (int32_t)d.idx;
        if (out_branch) out_branch[i] = (uint8_t)d.branch;
        if (out_matched) out_matched[i] = (uint32_t)d.matched;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Same, as one "snapshot batch" (cache_aware.h): every request walks/decides against the pre-batch trees, then the
// match side effects and inserts are applied in request order.
double orc_policy_select_batch_tokens_snapshot(void* h, const uint32_t* tokens, const uint64_t* offsets, size_t n, int32_t* out_idx,
                                               uint8_t* out_branch, uint32_t* out_matched) {
    auto* b = (PolicyBox*)h;
    auto t0 = std::chrono::steady_clock::now();
    b->pol.begin_snapshot_batch();
    for (size_t i = 0; i < n; ++i) {
        Decision d = b->pol.select_worker(b->workers, nullptr, tokens + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), true);
        out_idx[i] = (int32_t)d.idx;
        if (out_branch) out_branch[i] = (uint8_t)d.branch;
        if (out_matched) out_matched[i] = (uint32_t)d.matched;
    }
    b->pol.end_snapshot_batch();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// The request STREAM as the reference router sees it: after every pick the router takes a WorkerLoadGuard on the chosen worker
// (routers/http/router.rs:319-321, gated on policy.name() == "cache_aware"; worker/worker.rs:1067-1070 increment_load), so request i+1
// reads load()+1 there.  Requests are routed one after another, each pick bumping its worker's load; no request completes inside the
// batch.  The box's own fleet vector is left untouched (the bumps live in a copy); out_loads (nullable) receives the loads after the batch.
double orc_policy_select_batch_tokens_feedback(void* h, const uint32_t* tokens, const uint64_t* offsets, size_t n, int32_t* out_idx,
                                               uint8_t* out_branch, uint32_t* out_matched, uint64_t* out_loads) {
    auto* b = (PolicyBox*)h;
    std::vector<Worker> ws = b->workers;
    auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < n; ++i) {
        Decision d = b->pol.select_worker(ws, nullptr, tokens + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), true);
        out_idx[i] = (int32_t)d.idx;
        if (out_branch) out_branch[i] = (uint8_t)d.branch;
        if (out_matched) out_matched[i] = (uint32_t)d.matched;
        if (d.idx >= 0) ++ws[(size_t)d.idx].load;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (out_loads) for (size_t i = 0; i < ws.size(); ++i) out_loads[i] = ws[i].load;
    return secs;
}
// Batch of text requests (HTTP routing, cache_aware.rs:907-974): ragged UTF-8, offsets[n+1]; snapshot != 0 → snapshot batch.
double orc_policy_select_batch_text(void* h, const char* text, const uint64_t* offsets, size_t n, int snapshot, int32_t* out_idx,
                                    uint8_t* out_branch, uint32_t* out_matched, uint32_t* out_input) {
    auto* b = (PolicyBox*)h;
    auto t0 = std::chrono::steady_clock::now();
    if (snapshot) b->pol.begin_snapshot_batch();
    for (size_t i = 0; i < n; ++i) {
        std::string t(text + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
        Decision d = b->pol.select_worker(b->workers, &t, nullptr, 0, false);
        out_idx[i] = (int32_t)d.idx;
        if (out_branch) out_branch[i] = (uint8_t)d.branch;
        if (out_matched) out_matched[i] = (uint32_t)d.matched;
        if (out_input) out_input[i] = (uint32_t)d.input;
    }
    if (snapshot) b->pol.end_snapshot_batch();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- blake3 path hashes + hash_index ----
void orc_blake3(const uint8_t* data, size_t n, uint8_t* digest32) { b3::hash(data, n, digest32); }
uint64_t orc_hash_path_bytes(const uint8_t* data, size_t n) { return hash_path_bytes(data, n); }
uint64_t orc_hash_token_path(const uint32_t* toks, size_t n) { return hash_token_path(toks, n); }
// "hash=len:elem,elem,...\n" (tokens) or "hash=<utf8 prefix>\x1e" (text) for every entry, ascending by hash
size_t orc_policy_hash_index(void* h, const char* model, int text_kind, char* out, size_t cap) {
    auto* b = (PolicyBox*)h;
    std::string s;
    if (text_kind) for (auto& kv : b->pol.hash_index_text(model)) { s += std::to_string(kv.first) + "=" + kv.second; s.push_back('\x1e'); }
    else for (auto& kv : b->pol.hash_index_tokens(model)) {
        s += std::to_string(kv.first) + "=";
        for (size_t i = 0; i < kv.second.size(); ++i) { if (i) s.push_back(','); s += std::to_string(kv.second[i]); }
        s.push_back('\x1e');
    }
    return copy_str(s, out, cap);
}

// Read-only event-mode scoring with `threads` persistent host threads (the reference's concurrent-read design:
// select_worker takes &self, the index is only read, requests are spread over a tokio worker pool).  `steps` batches
// are routed back to back: batch s = tokens/offsets of (s % n_batches); threads are spawned ONCE, every thread owns a
// fixed shard of each batch; with step_barrier != 0 a barrier separates consecutive steps (a "step" completes before the next starts,
// as on the GPU side), with 0 the threads run free — the throughput a pool of independent tasks reaches.
// Each thread gets a private copy of the fleet vector (the processed counter is the only thing select_worker writes
// in event mode).  Falls back to one thread when the model has no populated indexer (tree modes mutate the tree).
// Returns elapsed seconds.
double orc_policy_select_steps_mt(void* h, const uint32_t* const* tokens, const uint64_t* const* offsets, size_t n_batches,
                                  size_t n, size_t steps, int32_t* out_idx, int threads, int step_barrier) {
    auto* b = (PolicyBox*)h;
    if (threads < 1) threads = 1;
    if (!b->workers.empty()) {
        std::string model = normalize_model_key(b->workers[0].model_id);
        if (!b->pol.has_event_indexer(model)) threads = 1;
    }
    std::atomic<size_t> arrived{0};
    std::atomic<size_t> generation{0};
    auto barrier = [&](size_t& local_gen) {
        size_t g = local_gen++;
        if (arrived.fetch_add(1) + 1 == (size_t)threads) { arrived.store(0); generation.store(g + 1); }
        else { while (generation.load(std::memory_order_acquire) <= g) std::this_thread::yield(); }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t) {
        ts.emplace_back([&, t]() {
            std::vector<Worker> ws = b->workers;
            size_t lo = n * (size_t)t / (size_t)threads, hi = n * (size_t)(t + 1) / (size_t)threads;
            size_t gen = 0;
            for (size_t s = 0; s < steps; ++s) {
                const uint32_t* tk = tokens[s % n_batches];
                const uint64_t* off = offsets[s % n_batches];
                for (size_t i = lo; i < hi; ++i) {
                    Decision d = b->pol.select_worker(ws, nullptr, tk + off[i], (size_t)(off[i + 1] - off[i]), true);
                    out_idx[i] = (int32_t)d.idx;
                }
                if (step_barrier) barrier(gen);   // 0: free-running threads, as independent tokio tasks are — the CPU's best throughput
            }
        });
    }
    for (auto& th : ts) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// "port-tuned" (oracle/tuned_event.h): same decisions from a flat bitset index, threads created before the clock starts.
// cfg thresholds are passed again because the policy keeps them private.  Returns elapsed seconds of the routed steps only.
double orc_tuned_select_steps_mt(void* h, void* indexer, float rel_thr, uint64_t abs_thr, size_t block_size, const uint32_t* const* tokens,
                                 const uint64_t* const* offsets, size_t n_batches, size_t n, size_t steps, int32_t* out_idx, int threads) {
    auto* b = (PolicyBox*)h;
    CacheAwareConfig c;
    c.balance_rel_threshold = rel_thr; c.balance_abs_threshold = abs_thr; c.block_size = block_size;
    TunedEventRouter r(((IndexerBox*)indexer)->ix, b->workers, c, block_size);
    return r.run(tokens, offsets, n_batches, n, steps, out_idx, threads);
}

// TreeHandle (cache_aware.rs:443-645)
int orc_policy_apply_known_remote_insert(void* h, const char* model, int kind, uint64_t node_hash, const char* worker_url) {
    return ((PolicyBox*)h)->pol.apply_known_remote_insert(model, kind ? CacheAwarePolicy::TREE_TOKEN : CacheAwarePolicy::TREE_STRING, node_hash, worker_url) ? 1 : 0;
}
// one repair entry; tenants = '\n'-separated URLs
void orc_policy_apply_repair_entry(void* h, const char* model, int kind, const void* data, size_t n, const char* tenants) {
    std::vector<std::string> ts;
    std::string cur;
    for (const char* c = tenants; ; ++c) { if (*c == '\n' || *c == 0) { if (!cur.empty()) ts.push_back(cur); cur.clear(); if (*c == 0) break; } else cur.push_back(*c); }
    if (kind) ((PolicyBox*)h)->pol.apply_repair_entry_tokens(model, (const uint32_t*)data, n, ts);
    else ((PolicyBox*)h)->pol.apply_repair_entry_text(model, std::string((const char*)data, n), ts);
}

// ---- prefix_hash policy + hash ring (prefix_hash.h) ----
void* orc_ring_new(const char* const* urls, size_t n) {
    std::vector<std::string> u;
    for (size_t i = 0; i < n; ++i) u.emplace_back(urls[i]);
    return new HashRing(u);
}
void orc_ring_free(void* r) { delete (HashRing*)r; }
size_t orc_ring_len(void* r) { return ((HashRing*)r)->len(); }
size_t orc_ring_worker_count(void* r) { return ((HashRing*)r)->worker_count(); }
uint64_t orc_ring_hash_position(const char* s, size_t n) { return HashRing::hash_position(std::string(s, n)); }
size_t orc_ring_entries(void* r, uint64_t* pos, uint32_t* url, size_t cap) {
    const auto& e = ((HashRing*)r)->entries();
    for (size_t i = 0; i < e.size() && i < cap; ++i) { pos[i] = e[i].pos; url[i] = e[i].url; }
    return e.size();
}
// healthy[u] per constructor URL; returns the URL index or -1
int64_t orc_ring_find_healthy(void* r, const char* key, size_t n, const uint8_t* healthy) {
    auto* ring = (HashRing*)r;
    std::unordered_map<std::string, bool> h;
    for (size_t u = 0; u < ring->worker_count(); ++u) h[ring->url(u)] = healthy[u] != 0;
    return ring->find_healthy(std::string(key, n), [&](const std::string& url) { return h[url]; });
}
uint64_t orc_prefix_hash(const uint32_t* tokens, size_t n, size_t prefix_token_count) {
    PrefixHashConfig c; c.prefix_token_count = prefix_token_count;
    return PrefixHashPolicy(c).compute_prefix_hash(tokens, n);
}
int orc_prefix_load_ok(double load_factor, uint64_t worker_load, uint64_t total, size_t n) {
    PrefixHashConfig c; c.load_factor = load_factor;
    return PrefixHashPolicy(c).load_ok(worker_load, total, n) ? 1 : 0;
}
// One batch against one fleet snapshot.  ring may be null (info.hash_ring = None); has_tokens 0 = info.tokens None for every request.
// Returns elapsed seconds.
double orc_prefix_select_batch(size_t prefix_token_count, double load_factor, const char* const* urls, const uint64_t* loads, const uint8_t* healthy,
                               size_t n_workers, void* ring, const uint32_t* tokens, const uint64_t* offsets, size_t n, int has_tokens,
                               int32_t* out_idx, uint8_t* out_branch) {
    PrefixHashConfig c; c.prefix_token_count = prefix_token_count; c.load_factor = load_factor;
    PrefixHashPolicy pol(c);
    std::vector<PrefixWorker> ws(n_workers);
    for (size_t i = 0; i < n_workers; ++i) { ws[i].url = urls[i]; ws[i].load = loads[i]; ws[i].healthy = healthy[i] != 0; }
    auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < n; ++i) {
        PrefixBranch br;
        out_idx[i] = (int32_t)pol.select_worker(ws, has_tokens ? tokens + offsets[i] : nullptr, (size_t)(offsets[i + 1] - offsets[i]), (HashRing*)ring, &br);
        if (out_branch) out_branch[i] = (uint8_t)br;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// power_of_two: one batch against one fleet snapshot (loads are NOT bumped between requests: a snapshot, like the product's batch call); request i
// uses draws 2i and 2i + 1 of the stream `seed`.  cached_urls / cached_usage = the policy's cached_loads (update_loads), effective_token_usage per URL.
void orc_p2c_select_batch(const char* const* urls, const uint64_t* loads, const uint8_t* healthy, const uint8_t* circuit, size_t n_workers,
                          const char* const* cached_urls, const double* cached_usage, size_t n_cached, uint64_t seed, size_t n, int32_t* out_idx,
                          int32_t* out_pairs, uint8_t* out_metric, uint64_t* out_processed) {
    PowerOfTwoPolicy pol;
    std::vector<std::pair<std::string, double>> cl;
    for (size_t i = 0; i < n_cached; ++i) cl.emplace_back(cached_urls[i], cached_usage[i]);
    pol.update_loads(cl);
    std::vector<Worker> ws(n_workers);
    for (size_t i = 0; i < n_workers; ++i) { ws[i].url = urls[i]; ws[i].load = loads[i]; ws[i].healthy = healthy[i] != 0; ws[i].circuit_ok = circuit ? circuit[i] != 0 : true; }
    for (size_t i = 0; i < n; ++i) {
        P2cStream rng(seed);
        rng.k = 2 * (uint64_t)i;
        const P2cDecision d = pol.select_worker(ws, rng);
        out_idx[i] = (int32_t)d.idx;
        if (out_pairs) { out_pairs[2 * i] = (int32_t)d.cand1; out_pairs[2 * i + 1] = (int32_t)d.cand2; }
        if (out_metric) out_metric[i] = (uint8_t)d.metric;
    }
    if (out_processed) for (size_t i = 0; i < n_workers; ++i) out_processed[i] = ws[i].processed;
}
double orc_p2c_effective_token_usage(const double* token_usage, size_t n) { return PowerOfTwoPolicy::effective_token_usage(std::vector<double>(token_usage, token_usage + n)); }

}  // extern "C"
