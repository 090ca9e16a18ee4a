/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into / imported by the product path.
 *
 * CPU restatement of the reference's event-driven KV index, following
 *   crates/kv_index/src/event_tree.rs  (reference @ 1c5701cf)
 *     :35      XXH3_SEED = 1337
 *     :48      MAX_WORKERS = 2048 (cap lifted here for BASELINE config 4; see kMaxWorkers)
 *     :122-151 compute_content_hash / compute_request_content_hashes
 *     :162-236 SeqEntry {Single, Multi} incl. the Single→Multi upgrade that never downgrades
 *     :305-366 apply_stored   :380-404 apply_removed   :410-420 apply_cleared   :426-435 remove_worker
 *     :438-444 current_size   :477-482 compute_next_seq_hash   :509-525 intern_worker
 *     :534-574 get_workers_lazy / count_workers_at   :582-657 linear_scan_drain
 *     :659-753 jump_search_matches
 * Pinned by the reference's own unit tests ported in tests/test_oracle_event_tree.py
 * (event_tree.rs:848-2020) and the XXH3 vectors in tests/golden/.
 *
 * The reference's FxHashSet/FxHashMap iteration order never reaches the result
 * (per-worker scores are order-independent), so plain ordered containers are used.
 */
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>
#include <optional>

#include "xxh3_ref.h"

namespace orc {

static const uint64_t kXxh3Seed = 1337;   // event_tree.rs:35
static const size_t kMaxWorkers = 65536;  // reference asserts < 2048 (event_tree.rs:48,520-523)

// event_tree.rs:122-129 — streaming XXH3 over LE bytes of each u32 == one-shot over the LE buffer
static inline uint64_t compute_content_hash(const uint32_t* toks, size_t n) {
    std::vector<uint8_t> buf(n * 4);
    for (size_t i = 0; i < n; ++i) {
        uint32_t t = toks[i];
        buf[4 * i] = (uint8_t)t; buf[4 * i + 1] = (uint8_t)(t >> 8);
        buf[4 * i + 2] = (uint8_t)(t >> 16); buf[4 * i + 3] = (uint8_t)(t >> 24);
    }
    return xxh3_64(buf.data(), buf.size(), kXxh3Seed);
}

// event_tree.rs:141-151 — full blocks only, block_size==0 → empty
static inline std::vector<uint64_t> compute_request_content_hashes(const uint32_t* toks, size_t n, size_t bs) {
    std::vector<uint64_t> out;
    if (bs == 0) return out;
    for (size_t off = 0; off + bs <= n; off += bs) out.push_back(compute_content_hash(toks + off, bs));
    return out;
}

// event_tree.rs:477-482
static inline uint64_t compute_next_seq_hash(uint64_t prev, uint64_t cur) {
    uint8_t b[16];
    for (int i = 0; i < 8; ++i) { b[i] = (uint8_t)(prev >> (8 * i)); b[8 + i] = (uint8_t)(cur >> (8 * i)); }
    return xxh3_64(b, 16, kXxh3Seed);
}

struct SeqEntry {  // event_tree.rs:162-236
    bool multi = false;
    uint64_t single_hash = 0;
    std::set<uint32_t> single_workers;
    std::map<uint64_t, std::set<uint32_t>> map;

    void insert(uint64_t h, uint32_t w) {
        if (!multi) {
            if (single_hash == h) { single_workers.insert(w); return; }
            map[single_hash] = std::move(single_workers);
            single_workers.clear();
            map[h].insert(w);
            multi = true;
            return;
        }
        map[h].insert(w);
    }
    // returns true when the entry is now empty and must be dropped
    bool remove(uint64_t h, uint32_t w) {
        if (!multi) {
            if (single_hash == h) { single_workers.erase(w); return single_workers.empty(); }
            return false;
        }
        auto it = map.find(h);
        if (it != map.end()) {
            it->second.erase(w);
            if (it->second.empty()) map.erase(it);
        }
        return map.empty();
    }
    const std::set<uint32_t>* get(uint64_t h) const {
        if (!multi) return single_hash == h ? &single_workers : nullptr;
        auto it = map.find(h);
        return it == map.end() ? nullptr : &it->second;
    }
    const std::set<uint32_t>* workers_if_single() const { return multi ? nullptr : &single_workers; }
};

struct BlockRec { size_t pos; uint64_t content; uint64_t prefix; };
using WorkerBlockMap = std::unordered_map<uint64_t, BlockRec>;  // event_tree.rs:246

enum ApplyStatus { APPLY_OK = 0, APPLY_WORKER_NOT_TRACKED = 1, APPLY_PARENT_NOT_FOUND = 2 };

struct OverlapScores {
    std::map<uint32_t, uint32_t> scores;
    std::map<uint32_t, size_t> tree_sizes;
};

class PositionalIndexer {
public:
    explicit PositionalIndexer(size_t jump) : jump_size_(jump ? jump : 1), tree_sizes_(kMaxWorkers, 0) {}

    std::optional<uint32_t> worker_id(const std::string& url) const {
        auto it = worker_to_id_.find(url);
        if (it == worker_to_id_.end()) return std::nullopt;
        return it->second;
    }
    uint32_t intern_worker(const std::string& url) {  // :509-525
        auto it = worker_to_id_.find(url);
        if (it != worker_to_id_.end()) return it->second;
        uint32_t id = next_worker_id_++;
        worker_to_id_[url] = id;
        return id;
    }

    // :305-366
    ApplyStatus apply_stored(uint32_t wid, const uint64_t* seq_hashes, const uint64_t* content_hashes, size_t n,
                             bool has_parent, uint64_t parent, WorkerBlockMap& wb) {
        if (n == 0) return APPLY_OK;
        size_t start = 0;
        bool have_prev = false;
        uint64_t prev = 0;
        if (has_parent) {
            if (wb.empty()) return APPLY_WORKER_NOT_TRACKED;
            auto it = wb.find(parent);
            if (it == wb.end()) return APPLY_PARENT_NOT_FOUND;
            start = it->second.pos + 1;
            prev = it->second.prefix;
            have_prev = true;
        }
        size_t fresh = 0;
        for (size_t i = 0; i < n; ++i) {
            size_t pos = start + i;
            uint64_t c = content_hashes[i];
            uint64_t pfx = have_prev ? compute_next_seq_hash(prev, c) : c;
            auto key = std::make_pair(pos, c);
            auto it = index_.find(key);
            if (it == index_.end()) {
                SeqEntry e; e.single_hash = pfx; e.single_workers.insert(wid);
                index_.emplace(key, std::move(e));
            } else {
                it->second.insert(pfx, wid);
            }
            bool existed = wb.count(seq_hashes[i]) != 0;
            wb[seq_hashes[i]] = BlockRec{pos, c, pfx};
            if (!existed) ++fresh;
            prev = pfx; have_prev = true;
        }
        if (fresh) tree_sizes_[wid] += fresh;
        return APPLY_OK;
    }

    // :380-404
    void apply_removed(uint32_t wid, const uint64_t* seq_hashes, size_t n, WorkerBlockMap& wb) {
        size_t removed = 0;
        for (size_t i = 0; i < n; ++i) {
            auto it = wb.find(seq_hashes[i]);
            if (it == wb.end()) continue;
            BlockRec r = it->second;
            wb.erase(it);
            drop(r, wid);
            ++removed;
        }
        if (removed) tree_sizes_[wid] -= removed;  // fetch_sub wraps like usize
    }
    // :410-420
    void apply_cleared(uint32_t wid, WorkerBlockMap& wb) {
        for (auto& kv : wb) drop(kv.second, wid);
        wb.clear();
        tree_sizes_[wid] = 0;
    }
    // :426-435
    void remove_worker(uint32_t wid, WorkerBlockMap& wb) { apply_cleared(wid, wb); }

    size_t current_size() const {  // :438-444
        size_t s = 0;
        for (uint32_t i = 0; i < next_worker_id_; ++i) s += tree_sizes_[i];
        return s;
    }
    size_t entry_count() const { return index_.size(); }
    size_t tree_size(uint32_t wid) const { return tree_sizes_[wid]; }
    size_t jump_size() const { return jump_size_; }
    uint32_t worker_count() const { return next_worker_id_; }
    // read-only visit of every (position, content hash) → SeqEntry (used to build the tuned bench variant, oracle/tuned_event.h)
    template <class F> void for_each_entry(F&& f) const { for (auto& kv : index_) f(kv.first.first, kv.first.second, kv.second); }

    // :659-753 (find_matches :461)
    OverlapScores find_matches(const std::vector<uint64_t>& seq, bool early_exit) const {
        OverlapScores out;
        if (seq.empty()) return out;
        std::vector<uint64_t> pfx;  // lazily computed rolling prefix hashes (:486-501)
        std::vector<uint32_t> active;
        {
            const SeqEntry* e = find(0, seq[0]);
            if (!e) return out;
            const std::set<uint32_t>* ws = e->workers_if_single();
            if (!ws) { ensure(pfx, 0, seq); ws = e->get(pfx[0]); }
            if (!ws) return out;
            active.assign(ws->begin(), ws->end());
        }
        if (active.empty()) return out;
        size_t len = seq.size();
        std::map<uint32_t, uint32_t> sc;
        if (early_exit) {
            for (uint32_t w : active) sc[w] = 1;
            finish(out, sc);
            return out;
        }
        size_t cur = 0;
        while (cur < len - 1 && !active.empty()) {
            size_t next = std::min(cur + jump_size_, len - 1);
            size_t count = count_at(next, seq, pfx);
            if (count == active.size()) {
                cur = next;
            } else {
                linear_scan_drain(seq, pfx, active, sc, cur + 1, next + 1);
                cur = next;
            }
        }
        for (uint32_t w : active) sc[w] = (uint32_t)len;
        finish(out, sc);
        return out;
    }

private:
    struct KeyHash {
        size_t operator()(const std::pair<size_t, uint64_t>& k) const {
            return (size_t)(k.second ^ (k.first * 0x9E3779B97F4A7C15ULL));
        }
    };
    const SeqEntry* find(size_t pos, uint64_t c) const {
        auto it = index_.find(std::make_pair(pos, c));
        return it == index_.end() ? nullptr : &it->second;
    }
    void drop(const BlockRec& r, uint32_t wid) {
        auto it = index_.find(std::make_pair(r.pos, r.content));
        if (it != index_.end() && it->second.remove(r.prefix, wid)) index_.erase(it);
    }
    static void ensure(std::vector<uint64_t>& pfx, size_t target, const std::vector<uint64_t>& seq) {
        while (pfx.size() <= target) {
            size_t p = pfx.size();
            pfx.push_back(p == 0 ? seq[0] : compute_next_seq_hash(pfx[p - 1], seq[p]));
        }
    }
    size_t count_at(size_t pos, const std::vector<uint64_t>& seq, std::vector<uint64_t>& pfx) const {  // :555-574
        const SeqEntry* e = find(pos, seq[pos]);
        if (!e) return 0;
        if (auto ws = e->workers_if_single()) return ws->size();
        ensure(pfx, pos, seq);
        auto ws = e->get(pfx[pos]);
        return ws ? ws->size() : 0;
    }
    void linear_scan_drain(const std::vector<uint64_t>& seq, std::vector<uint64_t>& pfx, std::vector<uint32_t>& active,
                           std::map<uint32_t, uint32_t>& sc, size_t lo, size_t hi) const {  // :582-657
        for (size_t pos = lo; pos < hi; ++pos) {
            if (active.empty()) break;
            const SeqEntry* e = find(pos, seq[pos]);
            if (!e) {
                for (uint32_t w : active) sc[w] = (uint32_t)pos;
                active.clear();
                break;
            }
            const std::set<uint32_t>* ws = e->workers_if_single();
            if (!ws) {
                ensure(pfx, pos, seq);
                ws = e->get(pfx[pos]);
                if (!ws) {
                    for (uint32_t w : active) sc[w] = (uint32_t)pos;
                    active.clear();
                    break;
                }
            }
            if (ws->size() < active.size()) {  // retain guard (:611, :641)
                size_t i = 0;
                while (i < active.size()) {
                    if (ws->count(active[i])) { ++i; }
                    else { sc[active[i]] = (uint32_t)pos; active[i] = active.back(); active.pop_back(); }
                }
            }
        }
    }
    void finish(OverlapScores& out, std::map<uint32_t, uint32_t>& sc) const {
        out.scores = sc;
        for (auto& kv : sc) out.tree_sizes[kv.first] = tree_sizes_[kv.first];
    }

    size_t jump_size_;
    std::unordered_map<std::pair<size_t, uint64_t>, SeqEntry, KeyHash> index_;
    std::vector<size_t> tree_sizes_;
    std::unordered_map<std::string, uint32_t> worker_to_id_;
    uint32_t next_worker_id_ = 0;
};

}  // namespace orc
