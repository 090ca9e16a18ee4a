// smgx_batcher.hpp — the per-request front end of the drop-in: LoadBalancingPolicy::select_worker is called once per request from many
// router tasks at a time (routers/http/router.rs:175, routers/grpc/common/stages/worker_selection.rs:157); the GPU path wants batches.
// Batcher coalesces concurrent route() / enqueue() calls into batches of up to max_batch requests (or whatever arrived within max_wait of the
// first one), keeps up to smgx_pipeline_depth() batches in flight through smgx_submit_tokens / smgx_wait, and hands every caller its
// own pick.  Every request of a batch sees one fleet snapshot — the state the reference's select_worker would read at that instant.
//
// Two transports, chosen by Options::mapped:
//   mapped = true (default; event-driven mode) — the LATENCY path.  The dispatcher is a busy-polling group-commit loop: the moment the
//     open batch holds a request and fewer than max_inflight batches are on the GPU it closes the batch and hands it to
//     smgx_submit_tokens_mapped — the kernel reads the batch's pinned staging in place (zero copy), writes the picks back into pinned
//     memory and raises a completion word the callers themselves spin on.  No staging copy, no cudaStreamSynchronize, no completer
//     thread, no mutex or condition variable on the request path.  While max_inflight batches are out, arrivals pile into the open
//     batch — batches grow with load instead of with a fixed window.  A model that is currently routed through a tree (SMGX_NOT_FOUND
//     from the mapped call) falls back to the staged transport for that batch.
//   mapped = false — the staged transport of round 1: smgx_submit_tokens / smgx_wait (H2D copy, ≤ pipeline_depth tickets in flight, a
//     completer thread publishing per-batch results through a condition variable), batches close max_wait after their first request.
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
        bool mapped = true;                              // zero-copy latency transport (see above); false = staged copies + tickets
        uint32_t max_inflight = 3;                       // mapped: batches on the GPU at once; arrivals beyond that join the open batch
        std::chrono::microseconds linger{0};             // mapped: keep a non-full batch open at least this long after its first request
        std::chrono::nanoseconds quiet{1500};            // mapped: … and until no request has joined it for this long (arrival pause)
        uint32_t max_request_tokens = 0;                 // mapped: bound on the longest request (0 = the policy's max_tokens_per_request)
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
        if (opt_.mapped) dispatcher_ = std::thread([this] { dispatch_loop_mapped(); });
        else {
            dispatcher_ = std::thread([this] { dispatch_loop(); });
            completer_ = std::thread([this] { complete_loop(); });
        }
    }
    ~Batcher() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; stop_flag_.store(true, std::memory_order_release); }
        cv_work_.notify_all(); cv_space_.notify_all();
        if (dispatcher_.joinable()) dispatcher_.join();   // first: nothing is submitted after this …
        { std::lock_guard<std::mutex> g(mu_); dispatcher_done_ = true; }
        cv_inflight_.notify_all();
        if (completer_.joinable()) completer_.join();      // … then the completer drains what is still in flight before the buffers go
        for (auto& b : ring_) { smgx_free_pinned(b->tokens); smgx_free_pinned(b->offsets); smgx_free_pinned(b->out); smgx_free_pinned(b->info); smgx_free_pinned((void*)b->done_flag); }
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
        if (opt_.mapped) {   // the GPU (or the fallback path) stores the batch's generation into the pinned completion word
            uint32_t spins = 0;
            while (__atomic_load_n(b->done_flag, __ATOMIC_ACQUIRE) != t.gen) {
                if (++spins < 4096) cpu_relax();
                else std::this_thread::yield();   // oversubscribed box: give the core away instead of burning the time slice
            }
            std::lock_guard<std::mutex> lk(b->mu);   // uncontended; orders the reads below after a fallback's status write
            status = b->status;
            if (status != SMGX_SUCCESS) error = b->error;
            idx = b->out[t.slot];
            if (info) *info = b->info[t.slot];
        } else {
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
        uint64_t* done_flag = nullptr;           // pinned: generation of the last completed use (written by the GPU in mapped mode)
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
        b->done_flag = (uint64_t*)smgx_alloc_pinned(64);
        if (!b->tokens || !b->offsets || !b->out || !b->info || !b->done_flag) throw std::runtime_error("smgx::Batcher: pinned allocation failed");
        b->offsets[0] = 0;
        *b->done_flag = 0;
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
        if (opt_.mapped) return t;   // the dispatcher polls the reservation word: no lock, no wake-up on the request path
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
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    // mapped transport: busy-polling group commit (see the header comment).  Only this thread closes batches.
    void dispatch_loop_mapped() {
        std::deque<std::pair<Batch*, uint64_t>> flying;   // (batch, generation) handed to the GPU, oldest first
        auto retire = [&] { while (!flying.empty() && __atomic_load_n(flying.front().first->done_flag, __ATOMIC_ACQUIRE) == flying.front().second) flying.pop_front(); };
        std::chrono::steady_clock::time_point first_seen{}, last_change{};
        uint32_t last_have = 0;
        uint64_t seen_gen = 0;
        uint32_t seen_batch = ~0u;
        uint32_t idle = 0;
        while (!stop_flag_.load(std::memory_order_acquire)) {
            retire();
            const uint32_t oi = open_idx_.load(std::memory_order_acquire);
            Batch* ob = batch_at(oi);
            const uint64_t r = ob->rsv.load(std::memory_order_acquire);
            const uint32_t have = (r & kClosed) ? 0 : slots_of(r);
            bool ship = false;
            if (have) {
                const bool full = have >= opt_.max_batch || tokens_of(r) + (opt_.max_request_tokens ? opt_.max_request_tokens : 1) > opt_.tokens_per_batch;
                if (!full) {
                    // ship once arrivals PAUSE: callers that enqueue back to back (a task pool with many requests outstanding) keep the batch open
                    // and it grows towards max_batch; a burst of blocking callers is complete within a microsecond or two and leaves at once
                    const uint64_t g = ob->gen.load(std::memory_order_relaxed);
                    const auto now = std::chrono::steady_clock::now();
                    if (seen_batch != oi || seen_gen != g) { seen_batch = oi; seen_gen = g; first_seen = now; last_have = 0; }
                    if (have != last_have) { last_have = have; last_change = now; }
                    ship = now - last_change >= opt_.quiet && now - first_seen >= opt_.linger;
                } else ship = true;
                if (ship && flying.size() >= opt_.max_inflight && !full) ship = false;   // GPU busy: let the batch grow (group commit)
                if (ship && flying.size() >= 2 * (size_t)opt_.max_inflight) ship = false; // even full batches queue only so deep
            }
            bool close_req = false;
            if (!ship) {
                // a caller may be asking for a rotation (its request did not fit) or for space; serve that under the lock, rarely
                ++idle;
                if ((idle & 63) != 0) { cpu_relax(); continue; }
                if (idle > (1u << 18) && !have && flying.empty()) std::this_thread::sleep_for(std::chrono::microseconds(20));   // long idle: stop burning the core
                std::lock_guard<std::mutex> g(mu_);
                close_req = ring_[open_]->state == OPEN && ring_[open_]->close_now && slots_of(ring_[open_]->rsv.load(std::memory_order_relaxed)) > 0;
                if (!close_req) continue;
            }
            idle = 0;
            std::unique_lock<std::mutex> lk(mu_);
            if (stop_) break;
            Batch& b = *ring_[open_];
            if (b.state != OPEN) continue;
            if (!close_req && ring_.size() >= opt_.max_ring) {
                // ring pressure: callers hold tickets of almost every batch (deep windows produce many small batches).  Let this one fill up
                // for a while instead of burning the last free batches on a handful of requests each.
                uint32_t free_batches = 0;
                for (auto& x : ring_) free_batches += x->state == FREE;
                if (free_batches <= 2 && slots_of(b.rsv.load(std::memory_order_relaxed)) < opt_.max_batch / 4 &&
                    std::chrono::steady_clock::now() - first_seen < std::chrono::microseconds(200)) continue;
            }
            const uint64_t rr = b.rsv.fetch_or(kClosed, std::memory_order_acq_rel);
            const uint32_t n = slots_of(rr);
            if (n == 0) { b.rsv.store(rr & ~kClosed, std::memory_order_release); continue; }   // raced with a recycle: nothing to ship
            b.n = n;
            b.state = SUBMITTED;
            b.readers.store(n, std::memory_order_release);
            ++stats_.batches; stats_.requests += n;
            if (n == opt_.max_batch) ++stats_.full_batches;
            const uint64_t gen = b.gen.load(std::memory_order_relaxed);
            open_next();
            lk.unlock();
            while (b.filled.load(std::memory_order_acquire) != n) cpu_relax();   // callers still copying their tokens in
            { std::lock_guard<std::mutex> g(b.mu); b.status = SMGX_SUCCESS; b.error.clear(); }
            char* err = nullptr;
            smgx_status st = smgx_submit_tokens_mapped(p_, model_.c_str(), b.tokens, b.offsets, n, opt_.max_request_tokens, b.out, b.info, b.done_flag, gen, &err);
            if (st == SMGX_NOT_FOUND) {   // the model is routed through a tree right now: staged transport for this batch, synchronously
                if (err) { smgx_free_string(err); err = nullptr; }
                uint64_t ticket = 0;
                for (;;) {
                    st = smgx_submit_tokens(p_, model_.c_str(), b.tokens, b.offsets, n, b.out, b.info, &ticket, &err);
                    if (st != SMGX_BUSY) break;
                    if (err) { smgx_free_string(err); err = nullptr; }
                    std::this_thread::yield();
                }
                if (st == SMGX_SUCCESS) st = smgx_wait(p_, ticket, &err);
                { std::lock_guard<std::mutex> g(b.mu); b.status = st; b.error = err ? err : ""; }
                __atomic_store_n(b.done_flag, gen, __ATOMIC_RELEASE);
            } else if (st != SMGX_SUCCESS) {
                { std::lock_guard<std::mutex> g(b.mu); b.status = st; b.error = err ? err : "mapped submit failed"; }
                __atomic_store_n(b.done_flag, gen, __ATOMIC_RELEASE);
            } else flying.emplace_back(&b, gen);
            if (err) smgx_free_string(err);
        }
        // drain: the callers of batches still on the GPU are spinning on their flags; the GPU raises them
        for (uint32_t spins = 0; !flying.empty() && spins < 200000000u; ++spins) { retire(); cpu_relax(); }
    }
    void complete_loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_inflight_.wait(lk, [&] { return dispatcher_done_ || !inflight_.empty(); });
            if (inflight_.empty()) { if (dispatcher_done_) return; continue; }   // the dispatcher has been joined: nothing more can arrive
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
    bool stop_ = false, dispatcher_done_ = false;
    std::atomic<bool> stop_flag_{false};
    Stats stats_;
    std::thread dispatcher_, completer_;
};

}  // namespace smgx
