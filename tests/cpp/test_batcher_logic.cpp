// Concurrency logic of smgx::Batcher (include/smgx_batcher.hpp) WITHOUT a GPU: the five C-ABI calls it makes are replaced by a mock
// "device" with 4 lanes and a 150 µs round trip, so slot reservation, ring growth, back-pressure (SMGX_BUSY, full ring) and result
// delivery can be exercised — and cannot hang a GPU box.  Not linked against libsmgx.so.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/smgx_batcher.hpp"

namespace {
std::mutex g_mu;
std::map<uint64_t, std::chrono::steady_clock::time_point> g_inflight;
uint64_t g_next = 1;
std::atomic<uint64_t> g_batches{0}, g_busy{0};
int32_t mock_pick(const uint32_t* t, uint32_t n) { return n ? (int32_t)((t[0] * 2654435761u + n) % 64u) : -1; }
}  // namespace

extern "C" {
void* smgx_alloc_pinned(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
void smgx_free_pinned(void* p) { std::free(p); }
void smgx_free_string(char* s) { std::free(s); }
smgx_status smgx_submit_tokens(smgx_policy*, const char*, const uint32_t* tokens, const uint32_t* offsets, uint32_t n, int32_t* out, smgx_decision_info* info,
                               uint64_t* ticket, char**) {
    std::lock_guard<std::mutex> g(g_mu);
    if (g_inflight.size() >= 4) { ++g_busy; return SMGX_BUSY; }
    for (uint32_t i = 0; i < n; ++i) {
        out[i] = mock_pick(tokens + offsets[i], offsets[i + 1] - offsets[i]);
        if (info) { info[i].matched = 0; info[i].input = offsets[i + 1] - offsets[i]; info[i].branch = 0; info[i].nodes = 0; }
    }
    *ticket = g_next++;
    g_inflight[*ticket] = std::chrono::steady_clock::now() + std::chrono::microseconds(150);
    ++g_batches;
    return SMGX_SUCCESS;
}
// mapped transport: picks computed at once, the "GPU" raises the completion word 30 µs later from its own thread
struct MockFlag { uint64_t* flag; uint64_t value; std::chrono::steady_clock::time_point due; };
std::mutex g_fmu;
std::vector<MockFlag> g_flags;
std::atomic<bool> g_gpu_stop{false};
std::atomic<uint64_t> g_mapped{0};
void mock_gpu() {
    while (!g_gpu_stop.load()) {
        {
            std::lock_guard<std::mutex> g(g_fmu);
            const auto now = std::chrono::steady_clock::now();
            for (size_t i = 0; i < g_flags.size();) {
                if (g_flags[i].due <= now) { __atomic_store_n(g_flags[i].flag, g_flags[i].value, __ATOMIC_RELEASE); g_flags[i] = g_flags.back(); g_flags.pop_back(); }
                else ++i;
            }
        }
        std::this_thread::sleep_for(std::chrono::microseconds(5));
    }
}
smgx_status smgx_submit_tokens_mapped(smgx_policy*, const char*, const uint32_t* tokens, const uint32_t* offsets, uint32_t n, uint32_t, int32_t* out,
                                      smgx_decision_info* info, uint64_t* done_flag, uint64_t done_value, char**) {
    for (uint32_t i = 0; i < n; ++i) {
        out[i] = mock_pick(tokens + offsets[i], offsets[i + 1] - offsets[i]);
        if (info) { info[i].matched = 0; info[i].input = offsets[i + 1] - offsets[i]; info[i].branch = 0; info[i].nodes = 0; }
    }
    ++g_mapped; ++g_batches;
    std::lock_guard<std::mutex> g(g_fmu);
    g_flags.push_back({done_flag, done_value, std::chrono::steady_clock::now() + std::chrono::microseconds(30)});
    return SMGX_SUCCESS;
}
smgx_status smgx_wait(smgx_policy*, uint64_t ticket, char**) {
    std::chrono::steady_clock::time_point due;
    { std::lock_guard<std::mutex> g(g_mu); due = g_inflight.at(ticket); }
    std::this_thread::sleep_until(due);
    std::lock_guard<std::mutex> g(g_mu);
    g_inflight.erase(ticket);
    return SMGX_SUCCESS;
}
}

static int run(int T, int R, int WIN, uint32_t max_batch, uint32_t ring, uint32_t max_ring, int wait_us, bool mapped) {
    smgx::Batcher::Options o;
    o.mapped = mapped;
    o.max_batch = max_batch; o.tokens_per_batch = max_batch * 40; o.ring = ring; o.max_ring = max_ring; o.max_wait = std::chrono::microseconds(wait_us);
    std::atomic<size_t> bad{0};
    smgx::Batcher::Stats st;
    {
        smgx::Batcher b(nullptr, "m", o);
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                std::vector<smgx::Batcher::Ticket> tk(WIN);
                std::vector<std::vector<uint32_t>> req(WIN);
                uint32_t x = 12345u + (uint32_t)t * 977u;
                for (int r0 = 0; r0 < R; r0 += WIN) {
                    const int cnt = std::min(WIN, R - r0);
                    for (int k = 0; k < cnt; ++k) {
                        x = x * 1664525u + 1013904223u;
                        req[k].assign((x >> 8) % 33, 0);   // 0..32 tokens, ragged, some empty
                        for (auto& v : req[k]) { x = x * 1664525u + 1013904223u; v = x; }
                        if (WIN > 1) tk[k] = b.enqueue(req[k].data(), (uint32_t)req[k].size());
                    }
                    if (WIN == 1) {   // the blocking per-request call
                        smgx_decision_info info;
                        if (b.route(req[0].data(), (uint32_t)req[0].size(), &info) != mock_pick(req[0].data(), (uint32_t)req[0].size()) || info.input != req[0].size()) ++bad;
                        continue;
                    }
                    for (int k = 0; k < cnt; ++k) {
                        smgx_decision_info info;
                        const int32_t got = b.get(tk[k], &info);
                        if (got != mock_pick(req[k].data(), (uint32_t)req[k].size()) || info.input != req[k].size()) ++bad;
                    }
                }
            });
        for (auto& x : th) x.join();
        st = b.stats();
    }
    std::printf("%s T=%d R=%d window=%d max_batch=%u ring=%u/%u: %llu requests in %llu batches (%llu full), %llu BUSY retries, %zu wrong\n", mapped ? "mapped" : "staged", T, R, WIN, max_batch, ring, max_ring,
                (unsigned long long)st.requests, (unsigned long long)st.batches, (unsigned long long)st.full_batches, (unsigned long long)g_busy.load(), bad.load());
    return (bad.load() == 0 && st.requests == (uint64_t)T * R) ? 0 : 1;
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    int rc = 0;
    std::thread gpu(mock_gpu);
    for (int m = 0; m < 2; ++m) {
        const bool mapped = m == 1;
        rc |= run(8, 3000, 1, 4096, 6, 64, 50, mapped);       // blocking route(): batches of ≤ 8
        rc |= run(8, 4000, 64, 4096, 6, 64, 100, mapped);     // task pool, 512 outstanding
        rc |= run(3, 6000, 700, 4096, 3, 64, 200, mapped);    // a window that spans more batches than the initial ring: the ring must grow
        rc |= run(16, 4000, 1, 4, 3, 3, 100, mapped);         // blocking callers, tiny batches, ring pinned at 3: callers wait for space
        {   // a caller holding tickets is told when the ring is exhausted instead of waiting for itself
            smgx::Batcher::Options o;
            o.mapped = mapped;
            o.max_batch = 2; o.tokens_per_batch = 64; o.ring = 3; o.max_ring = 3; o.max_wait = std::chrono::microseconds(10);
            smgx::Batcher b(nullptr, "m", o);
            std::vector<smgx::Batcher::Ticket> tk;
            const uint32_t one = 7;
            bool threw = false;
            try { for (int k = 0; k < 100; ++k) tk.push_back(b.enqueue(&one, 1)); } catch (const std::runtime_error&) { threw = true; }
            for (auto& t : tk) if (b.get(t) != mock_pick(&one, 1)) rc |= 1;
            if (!threw || tk.size() > 6) { std::printf("ring exhaustion not reported (%zu tickets)\n", tk.size()); rc |= 1; }
            if (b.route(&one, 1) != mock_pick(&one, 1)) rc |= 1;   // usable again once the tickets are in
        }
        rc |= run(4, 20000, 2000, 256, 3, 512, 20, mapped);   // full batches back to back, windows spanning dozens of batches
    }
    if (g_mapped.load() == 0) { std::printf("mapped transport never used\n"); rc |= 1; }
    g_gpu_stop.store(true);
    gpu.join();
    std::printf(rc ? "FAILED\n" : "ok\n");
    return rc;
}
