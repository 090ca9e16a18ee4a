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
        ring_.reserve(opt_.max_ring);   // never reallocates: tickets index into it
        for (uint32_t i = 0; i < opt_.ring; ++i) ring_.push_back(new_batch());
        ring_[0]->state = OPEN;
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

    // Hands one request to the open batch and returns at once (the oneshot-channel half of an async select_worker).  A caller that
    // holds unredeemed tickets must never wait for a batch to come free — it may be waiting for itself — so when every batch of a ring
    // already grown to max_ring is in flight or unread this throws; redeem tickets first or raise max_ring.
    Ticket enqueue(const uint32_t* tokens, uint32_t n) { return enqueue_impl(tokens, n, false); }
private:
    Ticket enqueue_impl(const uint32_t* tokens, uint32_t n, bool may_wait) {
        if (n > opt_.tokens_per_batch) throw std::invalid_argument("smgx::Batcher: request longer than tokens_per_batch");
        std::unique_lock<std::mutex> lk(mu_);
        Batch* b;
        for (;;) {   // reserve a slot in the open batch
            if (stop_) throw std::runtime_error("smgx::Batcher: stopped");
            b = ring_[open_].get();
            if (b->state == OPEN && b->n < opt_.max_batch && (uint64_t)b->used + n <= opt_.tokens_per_batch) break;
            if (b->state == OPEN && b->n > 0) { b->close_now = true; cv_work_.notify_one(); cv_space_.wait(lk); continue; }   // no room: ship what is there
            // no open batch: the ring is at max_ring and every batch is in flight or unread
            if (!may_wait) throw std::runtime_error("smgx::Batcher: every batch of the ring is in flight or unread; redeem tickets or raise max_ring");
            cv_space_.wait(lk);
        }
        Ticket t;
        t.batch = open_; t.slot = b->n++; t.gen = b->gen;
        const uint32_t at = b->used;
        b->used += n;
        b->offsets[t.slot + 1] = b->used;
        if (t.slot == 0) { b->first = std::chrono::steady_clock::now(); cv_work_.notify_one(); }
        else if (b->n == opt_.max_batch) cv_work_.notify_one();
        lk.unlock();
        if (n) std::memcpy(b->tokens + at, tokens, (size_t)n * 4);   // the copy runs outside the lock; the dispatcher waits for `filled`
        b->filled.fetch_add(1, std::memory_order_release);
        return t;
    }

public:
    // The pick for a ticket: index into the model's worker slice, or -1 for None.  Blocks until the ticket's batch has come back.
    int32_t get(const Ticket& t, smgx_decision_info* info = nullptr) {
        std::unique_lock<std::mutex> lk(mu_);
        Batch* b = ring_[t.batch].get();
        b->cv_done.wait(lk, [&] { return b->done_gen == t.gen; });
        const int status = b->status;
        const std::string error = status == SMGX_SUCCESS ? std::string() : b->error;
        const int32_t idx = b->out[t.slot];
        if (info) *info = b->info[t.slot];
        if (--b->readers == 0) {   // the last ticket in re-opens the batch
            b->n = 0; b->used = 0; b->filled.store(0, std::memory_order_relaxed); b->close_now = false; b->state = FREE; ++b->gen;
            if (ring_[open_]->state != OPEN) { open_ = t.batch; b->state = OPEN; }   // no open batch (ring at max_ring): callers are waiting for this one
            cv_space_.notify_all();
        }
        lk.unlock();
        if (status != SMGX_SUCCESS) throw std::runtime_error("smgx::Batcher: " + error);
        return idx;
    }
    // select_worker for one request, blocking (≈ max_wait + one GPU round trip).  Safe from any number of threads.
    int32_t route(const uint32_t* tokens, uint32_t n, smgx_decision_info* info = nullptr) { return get(enqueue_impl(tokens, n, true), info); }
    Stats stats() const { std::lock_guard<std::mutex> g(mu_); return stats_; }
    std::string debug_state() const {   // one line per batch: state / requests / filled / readers / generation
        std::lock_guard<std::mutex> g(mu_);
        std::string s = "open=" + std::to_string(open_) + " inflight=" + std::to_string(inflight_.size());
        for (auto& b : ring_)
            s += " [" + std::to_string((int)b->state) + " n=" + std::to_string(b->n) + " f=" + std::to_string(b->filled.load()) + " r=" + std::to_string(b->readers) +
                 " g=" + std::to_string(b->gen) + "/" + std::to_string(b->done_gen) + "]";
        return s;
    }

private:
    enum State { FREE, OPEN, SUBMITTED };
    struct Batch {
        uint32_t* tokens = nullptr; uint32_t* offsets = nullptr; int32_t* out = nullptr; smgx_decision_info* info = nullptr;
        uint32_t n = 0, used = 0, readers = 0;
        std::atomic<uint32_t> filled{0};
        State state = FREE;
        bool close_now = false;
        uint64_t gen = 1, done_gen = 0, ticket = 0;
        int status = SMGX_SUCCESS;
        std::string error;
        std::chrono::steady_clock::time_point first;
        std::condition_variable cv_done;
    };
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
    // mu_ held: make some FREE batch the open one; grow the ring when every batch is in flight or still being read (a caller holding
    // tickets of all of them must never be the one waiting for them).  With the ring at max_ring, callers wait on cv_space_.
    void open_next() {
        const uint32_t nb = (uint32_t)ring_.size();
        for (uint32_t k = 1; k <= nb; ++k) {
            const uint32_t c = (open_ + k) % nb;
            if (ring_[c]->state == FREE) { open_ = c; ring_[c]->state = OPEN; cv_space_.notify_all(); return; }
        }
        if (nb < opt_.max_ring) {
            ring_.push_back(new_batch());
            open_ = nb; ring_[nb]->state = OPEN;
        }
        cv_space_.notify_all();   // also when nothing could be opened: callers that asked for the close re-evaluate (wait for a batch, or throw)
    }

    void dispatch_loop() {
        std::unique_lock<std::mutex> lk(mu_);
        while (!stop_) {
            Batch& b = *ring_[open_];
            if (b.state != OPEN || b.n == 0) { cv_work_.wait(lk); continue; }
            const auto deadline = b.first + opt_.max_wait;
            cv_work_.wait_until(lk, deadline, [&] { return stop_ || b.n == opt_.max_batch || b.close_now; });
            if (stop_) break;
            // close: later requests go to the next batch of the ring (callers wait on cv_space_ while it is still being read)
            b.state = SUBMITTED;
            b.readers = b.n;
            ++stats_.batches; stats_.requests += b.n;
            if (b.n == opt_.max_batch) ++stats_.full_batches;
            const uint32_t mine = open_, n = b.n;
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
            lk.lock();
            if (st != SMGX_SUCCESS) { finish(b, st, err ? err : "submit failed"); if (err) smgx_free_string(err); }
            else { inflight_.push_back(mine); cv_inflight_.notify_one(); }
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
            lk.lock();
            finish(b, st, err ? err : "");
            if (err) smgx_free_string(err);
        }
    }
    void finish(Batch& b, int status, const std::string& error) {   // mu_ held
        b.status = status; b.error = error; b.done_gen = b.gen;
        b.cv_done.notify_all();
    }

    smgx_policy* p_;
    std::string model_;
    Options opt_;
    mutable std::mutex mu_;
    std::condition_variable cv_work_, cv_space_, cv_inflight_;
    std::vector<std::unique_ptr<Batch>> ring_;
    std::deque<uint32_t> inflight_;
    uint32_t open_ = 0;
    bool stop_ = false;
    Stats stats_;
    std::thread dispatcher_, completer_;
};

}  // namespace smgx
