/*
 * ORACLE — TEST / BENCH INFRASTRUCTURE ONLY. Never linked into / imported by the product path.
 *
 * PowerOfTwoPolicy (model_gateway/src/policies/power_of_two.rs:18-135) restated on the CPU.  The reference draws its two candidates
 * from rand::rng() (:48), a thread-local generator no caller can seed, so its own tests are statistical (:169-206, :210-246) or built so
 * that the outcome does not depend on the draw (:249-260, :267-320, :327-420); they are ported in tests/test_oracle_power_of_two.py.
 * To make a CUDA implementation checkable pick by pick, this restatement draws from an explicit counter-based stream instead:
 * draw k of a stream = splitmix64-finalise(seed + GOLDEN·(k + 1)); a draw x selects index ⌊x·n / 2^64⌋ of n.  Everything else follows the
 * reference line by line.  (Parity of the draw procedure with `rand` itself is out of reach by construction: "parity unpinned" for the
 * stream, pinned for the decision rule by the reference's deterministic tests.)
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "cache_aware.h"

namespace orc {

struct P2cStream {
    uint64_t seed, k = 0;
    explicit P2cStream(uint64_t s) : seed(s) {}
    static uint64_t finalise(uint64_t x) {
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
        x ^= x >> 27; x *= 0x94D049BB133111EBULL;
        return x ^ (x >> 31);
    }
    uint64_t next() { ++k; return finalise(seed + 0x9E3779B97F4A7C15ULL * k); }
    size_t random_range(size_t n) { return (size_t)(((unsigned __int128)next() * n) >> 64); }   // rng.random_range(0..n)
};

struct P2cDecision { int64_t idx = -1; int64_t cand1 = -1, cand2 = -1; int metric = 2; };   // metric: 0 request_count, 1 token_usage, 2 none

class PowerOfTwoPolicy {
public:
    // update_loads (:129-135): cached.extend(loads); value = WorkerLoadResponse::effective_token_usage() (protocols worker.rs:1039-1044)
    void update_loads(const std::vector<std::pair<std::string, double>>& loads) { for (auto& kv : loads) cached_[kv.first] = kv.second; }
    static double effective_token_usage(const std::vector<double>& token_usage_per_dp_rank) {
        if (token_usage_per_dp_rank.empty()) return 0.0;
        double s = 0; for (double v : token_usage_per_dp_rank) s += v;
        return s / (double)token_usage_per_dp_rank.size();
    }
    // select_worker (:36-120); `ws[idx].processed` is bumped on the pick (:110)
    P2cDecision select_worker(std::vector<Worker>& ws, P2cStream& rng) const {
        P2cDecision d;
        std::vector<size_t> healthy;                                    // get_healthy_worker_indices (policies/mod.rs:137-144)
        for (size_t i = 0; i < ws.size(); ++i) if (ws[i].healthy && ws[i].circuit_ok) healthy.push_back(i);
        if (healthy.empty()) return d;                                  // :40-42
        if (healthy.size() == 1) { d.idx = (int64_t)healthy[0]; return d; }   // :44-46 (no increment_processed on this path)
        const size_t idx1 = rng.random_range(healthy.size());           // :49
        const size_t idx2 = (idx1 + 1 + rng.random_range(healthy.size() - 1)) % healthy.size();   // :51-52
        const size_t w1 = healthy[idx1], w2 = healthy[idx2];
        auto r1 = cached_.find(ws[w1].url), r2 = cached_.find(ws[w2].url);
        double l1, l2;
        if (r1 != cached_.end() && r2 != cached_.end()) { l1 = r1->second; l2 = r2->second; d.metric = 1; }   // :70-78
        else { l1 = (double)ws[w1].load; l2 = (double)ws[w2].load; d.metric = 0; }                             // :79-88
        d.cand1 = (int64_t)w1; d.cand2 = (int64_t)w2;
        d.idx = (int64_t)(l1 <= l2 ? w1 : w2);                          // :91-95
        ++ws[(size_t)d.idx].processed;                                   // :110
        return d;
    }
    const char* name() const { return "power_of_two"; }

private:
    std::unordered_map<std::string, double> cached_;
};

}  // namespace orc
