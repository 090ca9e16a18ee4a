/*
 * ORACLE — TEST / BENCH INFRASTRUCTURE ONLY. Never linked into / imported by the product path.
 *
 * "port-tuned": the event-driven cache_aware decision (cache_aware.rs:648-690, :736-831 over event_tree.rs:461-753) restated a second
 * time with the data-structure choices a performance-minded CPU implementation would make, so that bench.py's reference arm
 * (--impl reference) is timed against something closer to what the reference's Rust path costs than the plain restatement
 * (oracle/positional_indexer.h: std::set / std::map worker sets, one heap vector per hashed block):
 *   - the index is one flat open-addressed table keyed by (position, content hash) with worker sets as fixed-width bitsets,
 *   - content hashes are computed straight from the token words into a stack buffer (no per-block allocation),
 *   - the fleet snapshot (healthy filter, min/max load, f32 imbalance gate, first-min-load) is reduced once per batch,
 *   - worker threads are created before the timed region and released by a barrier.
 * Read-only: built from a populated orc::PositionalIndexer (for_each_entry); its results are checked against the plain oracle in
 * tests/test_oracle_tuned.py before any timing is trusted.  The algorithm (count-only jump test, retain guard, Single entries
 * ignoring the prefix hash, sticky Multi, last-max tie-breaks) is unchanged.
 */
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "cache_aware.h"
#include "positional_indexer.h"

namespace orc {

class TunedEventRouter {
public:
    // ws: the worker slice (urls, loads, health) the decisions are made against; ix: populated indexer; cfg: thresholds + block size
    TunedEventRouter(const PositionalIndexer& ix, const std::vector<Worker>& ws, const CacheAwareConfig& cfg, size_t block_size)
        : jump_(ix.jump_size()), bs_(block_size) {
        const uint32_t n_ids = ix.worker_count();
        words_ = std::max<uint32_t>(1, (n_ids + 63) / 64);
        size_t cap = 16;
        while (cap < ix.entry_count() * 2 + 16) cap <<= 1;
        slots_.assign(cap, Slot{0, 0, 0, 0, 0});
        mask_ = cap - 1;
        ix.for_each_entry([&](size_t pos, uint64_t content, const SeqEntry& e) {
            size_t h = slot_of(pos, content);
            while (slots_[h].state) h = (h + 1) & mask_;
            Slot& s = slots_[h];
            s.content = content; s.pos = (uint32_t)pos;
            if (!e.multi) { s.state = 1; s.prefix = e.single_hash; s.payload = new_row(e.single_workers); }
            else {
                s.state = 2; s.payload = multi_.size(); s.prefix = e.map.size();
                for (auto& kv : e.map) multi_.push_back(MultiRec{kv.first, new_row(kv.second)});
            }
        });
        tree_sizes_.resize(n_ids);
        for (uint32_t i = 0; i < n_ids; ++i) tree_sizes_[i] = ix.tree_size(i);
        // fleet snapshot (select_worker prologue, cache_aware.rs:651-670) + per-id best slice entry (duplicate URLs: lowest load, then LAST index)
        elig_.assign(words_, 0);
        slice_of_id_.assign(n_ids, -1);
        load_of_id_.assign(n_ids, 0);
        uint64_t mn = UINT64_MAX, mx = 0;
        uint64_t hl = UINT64_MAX;
        for (size_t i = 0; i < ws.size(); ++i) {
            mn = std::min(mn, ws[i].load); mx = std::max(mx, ws[i].load);
            if (!(ws[i].healthy && ws[i].circuit_ok)) continue;
            ++n_healthy_;
            if (ws[i].load < hl) { hl = ws[i].load; min_load_idx_ = (int64_t)i; }
            auto id = ix.worker_id(ws[i].url);
            if (!id) continue;
            elig_[*id >> 6] |= 1ULL << (*id & 63);
            if (slice_of_id_[*id] < 0 || ws[i].load <= load_of_id_[*id]) { slice_of_id_[*id] = (int64_t)i; load_of_id_[*id] = ws[i].load; }
        }
        if (mn == UINT64_MAX) mn = 0;
        volatile float fmax = (float)mx;
        volatile float fprod = (float)mn * cfg.balance_rel_threshold;
        imbalanced_ = (mx - mn) > cfg.balance_abs_threshold && fmax > fprod;
    }

    // one decision; scratch = caller-owned buffers (no allocation on the hot path)
    struct Scratch { std::vector<uint64_t> ch, pfx, active, ws, last; };
    int64_t select(const uint32_t* tokens, size_t n_tokens, Scratch& sc) const {
        if (n_healthy_ == 0) return -1;
        if (imbalanced_) return min_load_idx_;
        const size_t nb = bs_ ? n_tokens / bs_ : 0;
        if (nb == 0 || tree_sizes_.empty()) return min_load_idx_;
        sc.ch.resize(nb);
        for (size_t b = 0; b < nb; ++b) sc.ch[b] = xxh3_64((const uint8_t*)(tokens + b * bs_), bs_ * 4, kXxh3Seed);   // little-endian host: LE bytes of the u32s
        sc.active.assign(words_, 0); sc.last.assign(words_, 0); sc.ws.assign(words_, 0);
        size_t pfx_n = 0;
        sc.pfx.resize(nb);
        uint32_t last_score = 0;
        bool have_last = false;
        auto prefix_at = [&](size_t p) {
            while (pfx_n <= p) { sc.pfx[pfx_n] = pfx_n == 0 ? sc.ch[0] : compute_next_seq_hash(sc.pfx[pfx_n - 1], sc.ch[pfx_n]); ++pfx_n; }
            return sc.pfx[p];
        };
        auto entry_set = [&](size_t pos, uint64_t*& out) -> bool {   // get_workers_lazy (:534-553)
            const Slot* s = find(pos, sc.ch[pos]);
            if (!s) return false;
            if (s->state == 1) { out = const_cast<uint64_t*>(&rows_[s->payload * words_]); return true; }
            const uint64_t want = prefix_at(pos);
            for (uint64_t k = 0; k < s->prefix; ++k)
                if (multi_[s->payload + k].prefix == want) { out = const_cast<uint64_t*>(&rows_[multi_[s->payload + k].row * words_]); return true; }
            return false;
        };
        auto popc = [&](const uint64_t* w) { size_t c = 0; for (uint32_t i = 0; i < words_; ++i) c += (size_t)__builtin_popcountll(w[i]); return c; };
        auto drained = [&](size_t pos, const uint64_t* gone) {   // workers leaving at `pos` get score pos; only the latest eligible set matters
            bool any = false;
            for (uint32_t i = 0; i < words_; ++i) any = any || (gone[i] & elig_[i]);
            if (any) { for (uint32_t i = 0; i < words_; ++i) sc.last[i] = gone[i] & elig_[i]; last_score = (uint32_t)pos; have_last = true; }
        };
        uint64_t* set = nullptr;
        if (!entry_set(0, set)) return min_load_idx_;
        memcpy(sc.active.data(), set, words_ * 8);
        size_t n_active = popc(sc.active.data());
        if (n_active == 0) return min_load_idx_;
        const size_t len = nb;
        size_t cur = 0;
        while (cur < len - 1 && n_active) {
            const size_t next = std::min(cur + jump_, len - 1);
            size_t count = 0;
            if (entry_set(next, set)) count = popc(set);
            if (count != n_active) {   // linear_scan_drain(cur+1 ..= next) (:582-657)
                for (size_t pos = cur + 1; pos <= next && n_active; ++pos) {
                    uint64_t* w = nullptr;
                    if (!entry_set(pos, w)) { drained(pos, sc.active.data()); std::fill(sc.active.begin(), sc.active.end(), 0); n_active = 0; break; }
                    if (popc(w) < n_active) {   // retain guard
                        for (uint32_t i = 0; i < words_; ++i) sc.ws[i] = sc.active[i] & ~w[i];
                        drained(pos, sc.ws.data());
                        for (uint32_t i = 0; i < words_; ++i) sc.active[i] &= w[i];
                        n_active = popc(sc.active.data());
                    }
                }
            }
            cur = next;
        }
        // score_overlap (cache_aware.rs:776-831): survivors score len, else the latest eligible drained set
        bool any = false;
        for (uint32_t i = 0; i < words_; ++i) { sc.ws[i] = sc.active[i] & elig_[i]; any = any || sc.ws[i]; }
        if (!any) { if (!have_last) return min_load_idx_; sc.ws = sc.last; (void)last_score; }
        int64_t best = -1;
        uint64_t bl = 0, bt = 0;
        for (uint32_t i = 0; i < words_; ++i) {
            uint64_t w = sc.ws[i];
            while (w) {
                const uint32_t id = i * 64 + (uint32_t)__builtin_ctzll(w);
                w &= w - 1;
                const int64_t sl = slice_of_id_[id];
                const uint64_t l = load_of_id_[id], t = tree_sizes_[id];
                if (best < 0 || l < bl || (l == bl && (t < bt || (t == bt && sl > best)))) { best = sl; bl = l; bt = t; }
            }
        }
        return best;
    }

    // `steps` batches back to back over `threads` persistent threads; threads exist before the clock starts.  → seconds
    double run(const uint32_t* const* tokens, const uint64_t* const* offsets, size_t n_batches, size_t n, size_t steps, int32_t* out_idx, int threads) const {
        if (threads < 1) threads = 1;
        std::atomic<int> ready{0};
        std::atomic<bool> go{false};
        std::vector<std::thread> ts;
        for (int t = 0; t < threads; ++t) {
            ts.emplace_back([&, t]() {
                Scratch sc;
                const size_t lo = n * (size_t)t / (size_t)threads, hi = n * (size_t)(t + 1) / (size_t)threads;
                ready.fetch_add(1);
                while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
                for (size_t s = 0; s < steps; ++s) {
                    const uint32_t* tk = tokens[s % n_batches];
                    const uint64_t* off = offsets[s % n_batches];
                    for (size_t i = lo; i < hi; ++i) out_idx[i] = (int32_t)select(tk + off[i], (size_t)(off[i + 1] - off[i]), sc);
                }
            });
        }
        while (ready.load() < threads) std::this_thread::yield();
        auto t0 = std::chrono::steady_clock::now();
        go.store(true, std::memory_order_release);
        for (auto& th : ts) th.join();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

private:
    struct Slot { uint64_t content; uint32_t pos; uint32_t state; uint64_t prefix; uint64_t payload; };   // state 0 empty, 1 Single, 2 Multi (prefix = count)
    struct MultiRec { uint64_t prefix; uint64_t row; };
    size_t slot_of(size_t pos, uint64_t content) const { return (size_t)(((content ^ ((uint64_t)pos * 0x9E3779B97F4A7C15ULL)) * 0xD6E8FEB86659FD93ULL) >> 32) & mask_; }
    const Slot* find(size_t pos, uint64_t content) const {
        size_t h = slot_of(pos, content);
        for (;;) {
            const Slot& s = slots_[h];
            if (!s.state) return nullptr;
            if (s.content == content && s.pos == (uint32_t)pos) return &s;
            h = (h + 1) & mask_;
        }
    }
    uint64_t new_row(const std::set<uint32_t>& ws) {
        const uint64_t r = rows_.size() / words_;
        rows_.resize(rows_.size() + words_, 0);
        for (uint32_t w : ws) rows_[r * words_ + (w >> 6)] |= 1ULL << (w & 63);
        return r;
    }

    size_t jump_, bs_;
    uint32_t words_ = 1;
    size_t mask_ = 0;
    std::vector<Slot> slots_;
    std::vector<uint64_t> rows_;
    std::vector<MultiRec> multi_;
    std::vector<uint64_t> tree_sizes_, elig_, load_of_id_;
    std::vector<int64_t> slice_of_id_;
    size_t n_healthy_ = 0;
    int64_t min_load_idx_ = -1;
    bool imbalanced_ = false;
};

}  // namespace orc
