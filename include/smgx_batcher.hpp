// smgx_batcher.hpp — the per-request front end of the drop-in: LoadBalancingPolicy::select_worker is called once per request from many
// router tasks at a time (routers/http/router.rs:175, routers/grpc/common/stages/worker_selection.rs:157); the GPU path wants batches.
// Batcher coalesces concurrent route() / enqueue() calls into batches of up to max_batch requests (or whatever arrived within max_wait of the
// first one), keeps up to smgx_pipeline_depth() batches in flight through smgx_submit_tokens / smgx_wait, and hands every caller its
// own pick.  Every request of a batch sees one fleet snapshot — the state the reference's select_worker would read at that instant.
//
// Header-only C++17 over the C ABI (include/smgx.h); the same structure is what INTEGRATION.md's Rust `batcher.route()` does with
// tokio::sync::oneshot instead of condition variables.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "smgx.h"

namespace smgx {

class Batcher {
public:
    struct Options {
        uint32_t max_batch = 4096;                       // requests per batch (≤ the policy's max_batch)
        uint32_t tokens_per_batch = 4096 * 1024;         // pinned staging per batch, in tokens
        std::chrono::microseconds max_wait{100};         // a batch leaves at most this long after its first request arrived
        uint32_t ring = 6;                               // batches allocated up front: open + in flight + being read by their callers
        uint32_t max_ring = 64;                          // the ring grows on demand up to this many batches (a caller may hold tickets of many)
    };
    struct Stats { uint64_t requests = 0, batches = 0, full_batches = 0; };

    Batcher(smgx_policy* policy, std::string model_key, const Options& o) : p_(policy), model_(std::move(model_key)), opt_(o) {
        if (opt_.ring < 3) opt_.ring = 3;
        if (opt_.max_ring < opt_.ring) opt_.max_ring = opt_.ring;
        ring_.reserve(opt_.max_ring);
        slots_.reset(new std::atomic<Batch*>[opt_.max_ring]);
        for (uint32_t i = 0; i < opt_.max_ring; ++i) slots_[i].store(nullptr, std::memory_order_relaxed);
        for (uint32_t i = 0; i < opt_.ring; ++i) add_batch();
        make_open(0);
        dispatcher_ = std::thread([this] { dispatch_loop(); });
        completer_ = std::thread([this] { complete_loop(); });
    }
    ~Batcher() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_work_.notify_all(); cv_inflight_.notify_all(); cv_space_.notify_all();
        if (dispatcher_.joinable()) dispatcher_.join();
        if (completer_.joinable()) completer_.join();
        for (auto& b : ring_) { smgx_free_pinned(b->tokens); smgx_free_pinned(b->offsets); smgx_free_pinned(b->out); smgx_free_pinned(b->info); }
    }
    Batcher(const Batcher&) = delete;
    Batcher& operator=(const Batcher&) = delete;

    // A request in flight.  Every ticket must be redeemed with get() exactly once: a batch is recycled when its last ticket is in.
    struct Ticket { uint32_t batch = 0, slot = 0; uint64_t gen = 0; };

    // Hands one request to the open batch and returns at once (the oneshot-channel half of an async select_worker).  The common case
    // is one compare-and-swap on the open batch's reservation word plus the copy of the tokens; no lock is taken.  A caller that
    // holds unredeemed tickets must never wait for a batch to come free — it may be waiting for itself — so when every batch of a ring
    // already grown to max_ring is in flight or unread this throws; redeem tickets first or raise max_ring.
    Ticket enqueue(const uint32_t* tokens, uint32_t n) { return enqueue_impl(tokens, n, false); }

    // The pick for a ticket: index into the model's worker slice, or -1 for None.  Blocks until the ticket's batch has come back.
    int32_t get(const Ticket& t, smgx_decision_info* info = nullptr) {
        Batch* b = batch_at(t.batch);
        int status;
        std::string error;
        int32_t idx;
        {
            std::unique_lock<std::mutex> lk(b->mu);
            b->cv_done.wait(lk, [&] { return b->done_gen == t.gen; });
            status = b->status;
            if (status != SMGX_SUCCESS) error = b->error;
            idx = b->out[t.slot];
            if (info) *info = b->info[t.slot];
        }
        if (b->readers.fetch_sub(1, std::memory_order_acq_rel) == 1) {   // the last ticket in recycles the batch
            std::lock_guard<std::mutex> g(mu_);
            b->filled.store(0, std::memory_order_relaxed); b->close_now = false; b->state = FREE; ++b->gen;
            if (ring_[open_]->state != OPEN) make_open(t.batch);   // no open batch (ring at max_ring): callers are waiting for this one
            cv_space_.notify_all();
        }
        if (status != SMGX_SUCCESS) throw std::runtime_error("smgx::Batcher: " + error);
        return idx;
    }
    // select_worker for one request, blocking (≈ max_wait + one GPU round trip).  Safe from any number of threads.
    int32_t route(const uint32_t* tokens, uint32_t n, smgx_decision_info* info = nullptr) { return get(enqueue_impl(tokens, n, true), info); }
    Stats stats() const { std::lock_guard<std::mutex> g(mu_); return stats_; }

private:
    // reservation word of a batch: bit 63 = closed, bits 40..62 = requests reserved, bits 0..39 = tokens reserved
    static constexpr uint64_t kClosed = 1ull << 63, kSlotOne = 1ull << 40, kTokMask = kSlotOne - 1;
    static uint32_t slots_of(uint64_t r) { return (uint32_t)((r & ~kClosed) >> 40); }
    static uint64_t tokens_of(uint64_t r) { return r & kTokMask; }

    enum State { FREE, OPEN, SUBMITTED };
    struct Batch {
        uint32_t* tokens = nullptr; uint32_t* offsets = nullptr; int32_t* out = nullptr; smgx_decision_info* info = nullptr;
        std::atomic<uint64_t> rsv{kClosed};      // closed until the batch is opened
        std::atomic<uint32_t> filled{0}, readers{0};
        std::atomic<uint64_t> gen{1};
        State state = FREE;                      // mu_
        bool close_now = false;                  // mu_
        uint32_t n = 0;                          // requests, fixed when the batch is closed
        uint64_t ticket = 0;
        std::chrono::steady_clock::time_point first;   // mu_
        uint64_t first_set = 0;                  // mu_: generation whose first request has arrived
        std::mutex mu;                           // completion: done_gen / status / error
        std::condition_variable cv_done;
        uint64_t done_gen = 0;
        int status = SMGX_SUCCESS;
        std::string error;
    };
    Batch* batch_at(uint32_t i) const { return slots_[i].load(std::memory_order_acquire); }
    std::unique_ptr<Batch> new_batch() {
        auto b = std::make_unique<Batch>();
        b->tokens = (uint32_t*)smgx_alloc_pinned((size_t)opt_.tokens_per_batch * 4);
        b->offsets = (uint32_t*)smgx_alloc_pinned(((size_t)opt_.max_batch + 1) * 4);
        b->out = (int32_t*)smgx_alloc_pinned((size_t)opt_.max_batch * 4);
        b->info = (smgx_decision_info*)smgx_alloc_pinned((size_t)opt_.max_batch * sizeof(smgx_decision_info));
        if (!b->tokens || !b->offsets || !b->out || !b->info) throw std::runtime_error("smgx::Batcher: pinned allocation failed");
        b->offsets[0] = 0;
        return b;
    }
    void add_batch() {   // mu_ held (or construction)
        ring_.push_back(new_batch());
        slots_[ring_.size() - 1].store(ring_.back().get(), std::memory_order_release);
    }
    void make_open(uint32_t i) {   // mu_ held
        Batch* b = ring_[i].get();
        open_ = i; b->state = OPEN;
        b->rsv.store(0, std::memory_order_release);   // publishes gen / filled of the new generation to the lock-free reservers
        open_idx_.store(i, std::memory_order_release);
    }
    // mu_ held: make some FREE batch the open one; grow the ring when every batch is in flight or still being read (a caller holding
    // tickets of all of them must never be the one waiting for them).  With the ring at max_ring, callers wait on cv_space_.
    void open_next() {
        const uint32_t nb = (uint32_t)ring_.size();
        bool opened = false;
        for (uint32_t k = 1; k <= nb && !opened; ++k) {
            const uint32_t c = (open_ + k) % nb;
            if (ring_[c]->state == FREE) { make_open(c); opened = true; }
        }
        if (!opened && nb < opt_.max_ring) { add_batch(); make_open(nb); }
        cv_space_.notify_all();   // also when nothing could be opened: callers that asked for the close re-evaluate (wait for a batch, or throw)
    }
    // one CAS: claim a request slot and its token range in batch b; false = closed, full, or no room for n tokens
    bool reserve(Batch* b, uint32_t n, uint32_t* slot, uint64_t* at) {
        uint64_t r = b->rsv.load(std::memory_order_acquire);
        for (;;) {
            if ((r & kClosed) || slots_of(r) >= opt_.max_batch || tokens_of(r) + n > opt_.tokens_per_batch) return false;
            if (b->rsv.compare_exchange_weak(r, r + kSlotOne + n, std::memory_order_acq_rel, std::memory_order_acquire)) break;
        }
        *slot = slots_of(r); *at = tokens_of(r);
        return true;
    }
    Ticket enqueue_impl(const uint32_t* tokens, uint32_t n, bool may_wait) {
        if (n > opt_.tokens_per_batch) throw std::invalid_argument("smgx::Batcher: request longer than tokens_per_batch");
        Ticket t;
        uint64_t at = 0;
        t.batch = open_idx_.load(std::memory_order_acquire);
        Batch* b = batch_at(t.batch);
        if (!reserve(b, n, &t.slot, &at)) {   // slow path: the open batch is full, closed, or has just been rotated
            std::unique_lock<std::mutex> lk(mu_);
            for (;;) {
                if (stop_) throw std::runtime_error("smgx::Batcher: stopped");
                t.batch = open_;
                b = ring_[open_].get();
                if (b->state == OPEN) {
                    if (reserve(b, n, &t.slot, &at)) break;
                    b->close_now = true; cv_work_.notify_one();   // no room for this request: ship what is there
                } else if (!may_wait)   // no open batch: the ring is at max_ring and every batch is in flight or unread
                    throw std::runtime_error("smgx::Batcher: every batch of the ring is in flight or unread; redeem tickets or raise max_ring");
                cv_space_.wait(lk);
            }
        }
        t.gen = b->gen.load(std::memory_order_acquire);
        b->offsets[t.slot + 1] = (uint32_t)(at + n);
        if (n) std::memcpy(b->tokens + at, tokens, (size_t)n * 4);
        b->filled.fetch_add(1, std::memory_order_release);   // the dispatcher submits once every reserved slot is filled
        if (t.slot == 0) { std::lock_guard<std::mutex> g(mu_); b->first = std::chrono::steady_clock::now(); b->first_set = b->gen.load(std::memory_order_relaxed); cv_work_.notify_one(); }
        else if (t.slot + 1 == opt_.max_batch) { std::lock_guard<std::mutex> g(mu_); cv_work_.notify_one(); }
        return t;
    }

    void dispatch_loop() {
        std::unique_lock<std::mutex> lk(mu_);
        while (!stop_) {
            Batch& b = *ring_[open_];
            const uint64_t g = b.gen.load(std::memory_order_relaxed);
            if (b.state != OPEN || b.first_set != g) { cv_work_.wait(lk); continue; }   // nothing has arrived in this generation yet
            const auto deadline = b.first + opt_.max_wait;
            cv_work_.wait_until(lk, deadline, [&] { return stop_ || b.close_now || slots_of(b.rsv.load(std::memory_order_relaxed)) >= opt_.max_batch; });
            if (stop_) break;
            // close: no reservation succeeds after this; later requests go to the next batch of the ring
            const uint64_t r = b.rsv.fetch_or(kClosed, std::memory_order_acq_rel);
            const uint32_t n = slots_of(r), mine = open_;
            b.n = n;
            b.state = SUBMITTED;
            b.readers.store(n, std::memory_order_release);
            ++stats_.batches; stats_.requests += n;
            if (n == opt_.max_batch) ++stats_.full_batches;
            open_next();
            lk.unlock();
            while (b.filled.load(std::memory_order_acquire) != n) std::this_thread::yield();   // callers still copying their tokens in
            char* err = nullptr;
            smgx_status st;
            for (;;) {
                st = smgx_submit_tokens(p_, model_.c_str(), b.tokens, b.offsets, n, b.out, b.info, &b.ticket, &err);
                if (st != SMGX_BUSY) break;
                if (err) { smgx_free_string(err); err = nullptr; }
                std::this_thread::yield();   // every lane busy: the completer is about to free one
            }
            if (st != SMGX_SUCCESS) { finish(b, st, err ? err : "submit failed"); if (err) smgx_free_string(err); lk.lock(); }
            else { lk.lock(); inflight_.push_back(mine); cv_inflight_.notify_one(); }
        }
    }
    void complete_loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_inflight_.wait(lk, [&] { return stop_ || !inflight_.empty(); });
            if (inflight_.empty()) { if (stop_) return; continue; }
            Batch& b = *ring_[inflight_.front()];
            inflight_.pop_front();
            lk.unlock();
            char* err = nullptr;
            const smgx_status st = smgx_wait(p_, b.ticket, &err);
            finish(b, st, err ? err : "");
            if (err) smgx_free_string(err);
            lk.lock();
        }
    }
    void finish(Batch& b, int status, const std::string& error) {   // mu_ NOT held
        { std::lock_guard<std::mutex> g(b.mu); b.status = status; b.error = error; b.done_gen = b.gen.load(std::memory_order_relaxed); }
        b.cv_done.notify_all();
    }

    smgx_policy* p_;
    std::string model_;
    Options opt_;
    mutable std::mutex mu_;
    std::condition_variable cv_work_, cv_space_, cv_inflight_;
    std::vector<std::unique_ptr<Batch>> ring_;                 // mu_; grows up to max_ring
    std::unique_ptr<std::atomic<Batch*>[]> slots_;             // ring_[i].get(), readable without mu_
    std::atomic<uint32_t> open_idx_{0};
    std::deque<uint32_t> inflight_;
    uint32_t open_ = 0;
    bool stop_ = false;
    Stats stats_;
    std::thread dispatcher_, completer_;
};

}  // namespace smgx
