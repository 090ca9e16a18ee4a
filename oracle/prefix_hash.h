/* ORACLE — test infrastructure only (never linked into, imported by or executed from the product path).
 *
 * CPU restatement of the reference's `prefix_hash` policy and the consistent hash ring it reads:
 *   model_gateway/src/worker/hash_ring.rs
 *     :17     VIRTUAL_NODES_PER_WORKER = 150
 *     :45-70  HashRing::new: for every URL, 150 entries at blake3("{url}#{vnode}")[..8] (LE), sorted by position
 *     :78-86  hash_position
 *     :102-134 find_healthy_url: partition_point(pos < key_pos), walk clockwise (wrapping), first URL for which is_healthy holds
 *     :137-150 is_empty / len / worker_count
 *   model_gateway/src/policies/prefix_hash.rs
 *     :38-58   PrefixHashConfig (+Default: 256 tokens, load factor 1.25)
 *     :106-113 compute_prefix_hash: xxh3_64 (seed 0) of the LE bytes of the first min(n, prefix_token_count) tokens
 *     :116-127 load_ok: f64, (total + 1) / n * load_factor, `<=`
 *     :130-201 find_worker_with_load_balance   :203-222 select_worker_impl
 * Third-party algorithms: blake3 = "1.5" (blake3_ref.h), xxhash-rust xxh3 (xxh3_ref.h), both restated from their published
 * specifications and pinned to official vectors.
 * Pinned by the reference's own unit tests (hash_ring.rs:152-198, prefix_hash.rs:236-414) ported in tests/test_oracle_prefix_hash.py
 * and by ring positions / prefix hashes produced by the independent Python `blake3` and `xxhash` modules (tests/golden/prefix_hash_vectors.json).
 *
 * Determinism notes: `sort_unstable_by_key` leaves the order of equal ring positions unspecified (a 64-bit collision between two
 * of ≤ 150·n keys); this restatement keeps insertion order.  `healthy_url_map` is a HashMap keyed by URL: when two slice entries
 * share a URL the LAST one wins (collect() overwrites), which is restated.
 */
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

#include "blake3_ref.h"
#include "xxh3_ref.h"

namespace orc {

constexpr size_t kVirtualNodesPerWorker = 150;   // hash_ring.rs:17

class HashRing {
public:
    HashRing() = default;
    explicit HashRing(const std::vector<std::string>& urls) {   // hash_ring.rs:45-70
        for (size_t u = 0; u < urls.size(); ++u) {
            urls_.push_back(urls[u]);
            for (size_t v = 0; v < kVirtualNodesPerWorker; ++v) entries_.push_back({hash_position(urls[u] + "#" + std::to_string(v)), (uint32_t)u});
        }
        std::stable_sort(entries_.begin(), entries_.end(), [](const Entry& a, const Entry& b) { return a.pos < b.pos; });
    }
    static uint64_t hash_position(const std::string& s) {   // :78-86
        uint8_t d[32];
        b3::hash((const uint8_t*)s.data(), s.size(), d);
        uint64_t h = 0;
        for (int i = 7; i >= 0; --i) h = (h << 8) | d[i];
        return h;
    }
    // :102-134; returns the index into the constructor's URL list, -1 = None
    template <class F> int64_t find_healthy(const std::string& key, F is_healthy) const {
        if (entries_.empty()) return -1;
        const uint64_t kp = hash_position(key);
        size_t start = std::partition_point(entries_.begin(), entries_.end(), [&](const Entry& e) { return e.pos < kp; }) - entries_.begin();
        std::vector<uint8_t> checked(urls_.size(), 0);
        for (size_t i = 0; i < entries_.size(); ++i) {
            const Entry& e = entries_[(start + i) % entries_.size()];
            if (checked[e.url]) continue;
            checked[e.url] = 1;
            if (is_healthy(urls_[e.url])) return e.url;
        }
        return -1;
    }
    bool is_empty() const { return entries_.empty(); }
    size_t len() const { return entries_.size(); }
    size_t worker_count() const { return entries_.size() / kVirtualNodesPerWorker; }
    const std::string& url(size_t i) const { return urls_[i]; }
    struct Entry { uint64_t pos; uint32_t url; };
    const std::vector<Entry>& entries() const { return entries_; }

private:
    std::vector<std::string> urls_;
    std::vector<Entry> entries_;
};

struct PrefixHashConfig {   // prefix_hash.rs:38-58
    size_t prefix_token_count = 256;
    double load_factor = 1.25;
};

enum PrefixBranch { PH_NO_HEALTHY_WORKERS = 0, PH_NO_TOKENS = 1, PH_RING_HIT = 2, PH_LOAD_BALANCE_WALK = 3, PH_FALLBACK_LEAST_LOAD = 4 };   // :61-83

struct PrefixWorker { std::string url; uint64_t load = 0; bool healthy = true; };

class PrefixHashPolicy {
public:
    explicit PrefixHashPolicy(const PrefixHashConfig& c = PrefixHashConfig()) : cfg_(c) {}
    static const char* name() { return "prefix_hash"; }   // :231-233

    uint64_t compute_prefix_hash(const uint32_t* tokens, size_t n) const {   // :106-113
        const size_t k = std::min(n, cfg_.prefix_token_count);
        return xxh3_64(tokens, k * 4, 0);
    }
    bool load_ok(uint64_t worker_load, uint64_t total_load, size_t num_workers) const {   // :116-127
        if (total_load == 0 || num_workers == 0) return true;
        const double avg = (double)(total_load + 1) / (double)num_workers;
        const double threshold = avg * cfg_.load_factor;
        return (double)worker_load <= threshold;
    }
    // :203-222 + :130-201.  tokens == nullptr ⇔ info.tokens is None; ring == nullptr ⇔ info.hash_ring is None.  Returns idx or -1.
    int64_t select_worker(const std::vector<PrefixWorker>& ws, const uint32_t* tokens, size_t n, const HashRing* ring, PrefixBranch* branch) const {
        PrefixBranch dummy;
        if (!branch) branch = &dummy;
        if (ws.empty()) { *branch = PH_NO_HEALTHY_WORKERS; return -1; }
        if (!tokens || n == 0) { *branch = PH_NO_TOKENS; return -1; }
        const uint64_t ph = compute_prefix_hash(tokens, n);
        std::vector<size_t> healthy;
        for (size_t i = 0; i < ws.size(); ++i) if (ws[i].healthy) healthy.push_back(i);
        if (healthy.empty()) { *branch = PH_NO_HEALTHY_WORKERS; return -1; }
        uint64_t total = 0;
        for (size_t i : healthy) total += ws[i].load;
        const size_t nw = healthy.size();
        if (ring) {
            char key[17];
            std::snprintf(key, sizeof key, "%016llx", (unsigned long long)ph);   // format!("{prefix_hash:016x}")
            std::unordered_map<std::string, size_t> url_map;
            for (size_t i : healthy) url_map[ws[i].url] = i;   // collect(): a later duplicate URL overwrites
            int64_t u = ring->find_healthy(key, [&](const std::string& url) { return url_map.count(url) != 0; });
            if (u >= 0) {
                const size_t idx = url_map[ring->url((size_t)u)];
                if (load_ok(ws[idx].load, total, nw)) { *branch = PH_RING_HIT; return (int64_t)idx; }
                bool have = false;
                size_t best = 0;
                for (size_t i : healthy)
                    if (load_ok(ws[i].load, total, nw) && (!have || ws[i].load < ws[best].load)) { have = true; best = i; }   // min_by_key → FIRST minimum
                *branch = PH_LOAD_BALANCE_WALK;
                return (int64_t)(have ? best : idx);
            }
        }
        size_t best = healthy[0];
        for (size_t i : healthy) if (ws[i].load < ws[best].load) best = i;
        *branch = PH_FALLBACK_LEAST_LOAD;
        return (int64_t)best;
    }

private:
    PrefixHashConfig cfg_;
};

}  // namespace orc
