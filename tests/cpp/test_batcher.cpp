// Per-request front end (include/smgx_batcher.hpp) on a B200: T caller threads route R requests each through smgx::Batcher against the
// event-driven index; every pick must equal the oracle's pick for that request (event-mode picks do not depend on arrival order), and
// the run reports throughput and per-request latency.  Test infrastructure: the oracle is the checker.
//   test_batcher [threads] [requests_per_thread] [window] [max_wait_us] [mapped 1|0] [max_inflight] [linger_us] [quiet_ns]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "../../include/smgx.hpp"
#include "../../include/smgx_batcher.hpp"
#include "../../oracle/cache_aware.h"

int main(int argc, char** argv) {
    const int T = argc > 1 ? std::atoi(argv[1]) : 16, R = argc > 2 ? std::atoi(argv[2]) : 2000, WIN = argc > 3 ? std::atoi(argv[3]) : 32;
    const int wait_us = argc > 4 ? std::atoi(argv[4]) : 100;
    const bool mapped = argc > 5 ? std::atoi(argv[5]) != 0 : true;
    const int max_inflight = argc > 6 ? std::atoi(argv[6]) : 3, linger_us = argc > 7 ? std::atoi(argv[7]) : 0, quiet_ns = argc > 8 ? std::atoi(argv[8]) : 1500;
    const uint32_t W = 64, BS = 16, TOK = 512, SEQS = 20000;
    std::mt19937_64 rng(7);
    smgx::CacheAwareConfig cfg; cfg.eviction_interval_secs = 0; cfg.cache_threshold = 0.3f; cfg.balance_abs_threshold = 64; cfg.balance_rel_threshold = 1.5f; cfg.block_size = BS;
    smgx::CacheAwarePolicy policy(cfg, 0, 4096, TOK);
    orc::CacheAwareConfig oc; oc.cache_threshold = 0.3f; oc.balance_abs_threshold = 64; oc.balance_rel_threshold = 1.5f; oc.eviction_interval_secs = 0; oc.block_size = BS;
    orc::CacheAwarePolicy opol(oc);
    smgx::Workers ws;
    std::vector<orc::Worker> ows(W);
    for (uint32_t i = 0; i < W; ++i) {
        auto w = std::make_shared<smgx::BasicWorker>("http://w" + std::to_string(i) + ":8000", "m");
        w->set_load(rng() % 16);
        ows[i].url = w->url(); ows[i].model_id = "m"; ows[i].load = w->load();
        ws.push_back(w);
    }
    policy.init_workers(ws);
    opol.init_workers(ows);
    auto monitor = policy.kv_event_monitor();
    auto ix = monitor->create_indexer("m", 64);
    monitor->set_block_size("m", BS);
    orc::PositionalIndexer oix(64);
    std::vector<orc::WorkerBlockMap> owb(W);
    opol.set_monitor(true); opol.attach_indexer("m", &oix); opol.set_block_size("m", BS);
    std::vector<std::vector<uint32_t>> seqs(SEQS, std::vector<uint32_t>(TOK));
    uint64_t next_seq = 1;
    for (uint32_t s = 0; s < SEQS; ++s) {
        for (auto& t : seqs[s]) t = (uint32_t)(rng() % 50000);
        const uint32_t w = s % W;
        const uint32_t wid = ix->intern_worker(ws[w]->url());
        oix.intern_worker(ws[w]->url());
        std::vector<uint64_t> seq(TOK / BS), con(TOK / BS);
        for (uint32_t b = 0; b < TOK / BS; ++b) { seq[b] = next_seq++; con[b] = orc::compute_content_hash(seqs[s].data() + b * BS, BS); }
        std::vector<smgx::StoredBlock> blocks(TOK / BS);
        for (uint32_t b = 0; b < TOK / BS; ++b) blocks[b] = {seq[b], con[b]};
        ix->apply_stored(wid, blocks);
        oix.apply_stored(wid, seq.data(), con.data(), seq.size(), false, 0, owb[wid]);
    }
    policy.set_kv_event_monitor(monitor);
    // the fleet snapshot the batches read (the Worker scalars are pushed once; the gateway refreshes them as loads change)
    policy.select_worker_batch(ws, {seqs[0]});

    // requests: 80 % a cached sequence, 10 % a cached prefix + novel tail, 10 % novel
    const size_t N = (size_t)T * R;
    std::vector<std::vector<uint32_t>> reqs(N);
    for (auto& q : reqs) {
        const uint64_t u = rng() % 10;
        q = seqs[rng() % SEQS];
        if (u == 8) { const uint32_t keep = BS * (1 + rng() % 31); for (uint32_t i = keep; i < TOK; ++i) q[i] = (uint32_t)(rng() % 50000); }
        else if (u == 9) for (auto& t : q) t = (uint32_t)(rng() % 50000);
    }
    std::vector<int32_t> got(N, -2);
    std::vector<float> lat_us(N, 0.f);
    smgx::Batcher::Options bo;
    bo.max_batch = 4096; bo.tokens_per_batch = 4096 * TOK; bo.max_wait = std::chrono::microseconds(wait_us);
    bo.ring = 24;   // sized for the run: no pinned allocation on the request path
    bo.mapped = mapped; bo.max_inflight = (uint32_t)max_inflight; bo.linger = std::chrono::microseconds(linger_us); bo.max_request_tokens = TOK; bo.quiet = std::chrono::nanoseconds(quiet_ns);
    double secs;
    smgx::Batcher::Stats st;
    {
        smgx::Batcher batcher(policy.handle()->p, "m", bo);
        {   // warm-up, untimed: sizes the device staging of the lanes.  Redeemed in windows: a caller must not hold more tickets than the ring
            // has batches, and how many batches a window becomes depends on how fast the dispatcher turns them around (a 16 384-ticket window
            // became more than max_ring = 64 batches once the kernels got faster)
            const size_t win = mapped ? 512 : 2048;
            for (size_t k0 = 0; k0 < 4 * 4096; k0 += win) {
                std::vector<smgx::Batcher::Ticket> warm;
                for (size_t k = k0; k < k0 + win; ++k) warm.push_back(batcher.enqueue(reqs[k % N].data(), (uint32_t)reqs[k % N].size()));
                for (auto& t : warm) batcher.get(t);
            }
        }
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                std::vector<smgx::Batcher::Ticket> tk(WIN);
                std::vector<std::chrono::steady_clock::time_point> at(WIN);
                for (int r0 = 0; r0 < R; r0 += WIN) {   // a router task pool: WIN requests outstanding per thread
                    const int cnt = std::min(WIN, R - r0);
                    if (WIN == 1) {   // the blocking per-request call: route() may wait for a batch to come free, enqueue() (a caller that may hold tickets) must not
                        const size_t id = (size_t)t * R + r0;
                        at[0] = std::chrono::steady_clock::now();
                        got[id] = batcher.route(reqs[id].data(), (uint32_t)reqs[id].size());
                        lat_us[id] = std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - at[0]).count();
                        continue;
                    }
                    for (int k = 0; k < cnt; ++k) { const size_t id = (size_t)t * R + r0 + k; at[k] = std::chrono::steady_clock::now(); tk[k] = batcher.enqueue(reqs[id].data(), (uint32_t)reqs[id].size()); }
                    for (int k = 0; k < cnt; ++k) {
                        const size_t id = (size_t)t * R + r0 + k;
                        got[id] = batcher.get(tk[k]);
                        lat_us[id] = std::chrono::duration<float, std::micro>(std::chrono::steady_clock::now() - at[k]).count();
                    }
                }
            });
        for (auto& x : th) x.join();
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        st = batcher.stats();
    }
    size_t bad = 0;
    for (size_t i = 0; i < N; ++i) {
        const orc::Decision d = opol.select_worker(ows, nullptr, reqs[i].data(), reqs[i].size(), true);
        if (got[i] != (int32_t)d.idx) { if (++bad <= 5) std::fprintf(stderr, "MISMATCH request %zu: got %d, oracle %lld\n", i, got[i], (long long)d.idx); }
    }
    std::sort(lat_us.begin(), lat_us.end());
    std::printf("{\"mode\": \"per-request front end (smgx::Batcher), event-driven cache_aware, 64 workers, 512-token requests\", \"transport\": \"%s\", "
                "\"threads\": %d, \"outstanding_per_thread\": %d, \"requests\": %zu, \"max_wait_us\": %d, \"max_inflight\": %d, \"linger_us\": %d, \"quiet_ns\": %d, \"decisions_per_s\": %.1f, \"batches\": %llu, "
                "\"mean_batch\": %.1f, \"p50_latency_us\": %.1f, \"p99_latency_us\": %.1f, \"p999_latency_us\": %.1f, \"mismatches_vs_oracle\": %zu}\n",
                mapped ? "mapped zero-copy (smgx_submit_tokens_mapped, callers spin on the completion word)" : "staged (smgx_submit_tokens / smgx_wait)",
                T, WIN, N, wait_us, max_inflight, linger_us, quiet_ns, N / secs, (unsigned long long)st.batches, st.batches ? (double)st.requests / st.batches : 0.0, lat_us[N / 2],
                lat_us[(size_t)(N * 0.99)], lat_us[(size_t)(N * 0.999)], bad);
    return bad ? 1 : 0;
}
