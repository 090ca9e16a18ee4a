// sm_100a kernels for the event-driven cache-aware pick.
//
//   K2b  content hashes (XXH3-64, seed 1337) of every full block of a request
//          ← compute_request_content_hashes      crates/kv_index/src/event_tree.rs:141-151
//        jump search over the GPU-resident positional index, exact reference semantics incl. the count-only
//        jump test, the retain guard and Single/Multi entries
//          ← PositionalIndexer::jump_search_matches / linear_scan_drain / count_workers_at  event_tree.rs:555-753
//   K3   per-worker (overlap, load, tree_size) → argmax worker with the reference's tie-breaks, min-load fallback
//          ← CacheAwarePolicy::score_overlap / select_worker_event_driven   cache_aware.rs:736-831
//        fleet_prepare: healthy filter, min/max load, f32 imbalance gate, first-min-load
//          ← select_worker prologue   cache_aware.rs:651-670, policies/mod.rs:137-144
//
// Mapping: one warp per request.  Lanes hash blocks lane, lane+32, … (64 B each, 4×LDG.128) into a per-warp
// shared-memory array; index probes are 32-byte slot loads (one sector); worker sets are bitsets so set
// algebra is popc/and/andnot — replicated in every lane when the fleet has ≤ 64 interned workers (W1), otherwise
// one u64 word per lane (≤ 2048 workers) reduced with warp collectives.  No tensor-core work exists on this path.
#include <cuda_runtime.h>

#include "kernels.h"
#include "xxh3.cuh"

namespace smgx {
namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src), hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
    return mk64(lo, hi);
}
__device__ __forceinline__ uint64_t shfl64_xor(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)v, m), hi = __shfl_xor_sync(FULL, (uint32_t)(v >> 32), m);
    return mk64(lo, hi);
}

__device__ __forceinline__ Slot load_slot(const Slot* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Slot s;
    s.content = mk64(a.x, a.y); s.pos = a.z; s.state = a.w;
    s.prefix = mk64(b.x, b.y); s.payload = mk64(b.z, b.w);
    return s;
}
__device__ __forceinline__ MultiNode load_node(const MultiNode* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    MultiNode n;
    n.prefix = mk64(a.x, a.y); n.payload = mk64(a.z, a.w); n.next = b.x; n.pad0 = 0; n.pad1 = 0;
    return n;
}

// DashMap::get(&(position, content_hash)) — linear probing over 32 B slots.  Works per-thread (divergent) or
// warp-uniform (same address in every lane → one broadcast transaction).
__device__ __forceinline__ bool probe(const EventIndexView& v, uint32_t pos, uint64_t content, Slot& out) {
    uint32_t h = slot_hash(pos, content) & v.mask;
#pragma unroll 1
    for (;;) {
        Slot s = load_slot(v.slots + h);
        if (s.state == SLOT_EMPTY) return false;
        if (s.state != SLOT_TOMB && s.content == content && s.pos == pos) { out = s; return true; }
        h = (h + 1) & v.mask;
    }
}

// ---- worker sets --------------------------------------------------------------------------------------------
template <bool W1> __device__ __forceinline__ uint64_t load_set(const EventIndexView& v, uint64_t payload, int lane) {
    if (W1) return payload;
    return (uint32_t)lane < v.words ? __ldg(v.rows + (size_t)(uint32_t)payload * v.words + lane) : 0ULL;
}
template <bool W1> __device__ __forceinline__ uint32_t set_popc(uint64_t s) {
    if (W1) return (uint32_t)__popcll(s);
    return __reduce_add_sync(FULL, (uint32_t)__popcll(s));
}
template <bool W1> __device__ __forceinline__ bool set_any(uint64_t s) {
    if (W1) return s != 0;
    return __any_sync(FULL, s != 0);
}
// popcount of a whole set by ONE lane (scan path: every lane sizes the entry it probed itself)
template <bool W1> __device__ __forceinline__ uint32_t lane_popc_set(const EventIndexView& v, uint64_t payload) {
    if (W1) return (uint32_t)__popcll(payload);
    const uint64_t* r = v.rows + (size_t)(uint32_t)payload * v.words;
    uint32_t c = 0;
    for (uint32_t i = 0; i < v.words; ++i) c += (uint32_t)__popcll(__ldg(r + i));
    return c;
}

// rolling prefix hash, computed lazily and only when a Multi entry needs it (event_tree.rs:486-501)
struct PrefixCache { int pos; uint64_t val; };
__device__ __forceinline__ uint64_t prefix_at(const uint64_t* ch, int p, PrefixCache& pc) {
    if (pc.pos < 0 || p < pc.pos) { pc.pos = 0; pc.val = ch[0]; }
    while (pc.pos < p) { ++pc.pos; pc.val = xxh3_pair(pc.val, ch[pc.pos], kSeed); }
    return pc.val;
}

// workers_if_single() / SeqEntry::get(prefix) (event_tree.rs:216-235); warp-uniform slot
template <bool W1>
__device__ __forceinline__ bool entry_set(const EventIndexView& v, uint32_t state, uint64_t payload, const uint64_t* ch, int p, int lane,
                                          PrefixCache& pc, uint64_t& set) {
    if (state == SLOT_SINGLE) { set = load_set<W1>(v, payload, lane); return true; }
    uint64_t want = prefix_at(ch, p, pc);
    for (uint32_t i = (uint32_t)payload; i != kNil;) {
        MultiNode nd = load_node(v.multi + i);
        if (nd.prefix == want) { set = load_set<W1>(v, nd.payload, lane); return true; }
        i = nd.next;
    }
    return false;
}

// ---- event sinks: what happens when workers leave the active set at `pos` (score = pos) -------------------
template <bool W1> struct SelectSink {  // keeps only what the argmax can still need
    uint64_t elig, last;
    uint32_t last_score;
    __device__ __forceinline__ void on_event(uint32_t pos, uint64_t set) {
        uint64_t e = set & elig;
        if (set_any<W1>(e)) { last = e; last_score = pos; }
    }
};
template <bool W1> struct DumpSink {  // OverlapScores.scores materialised (find_matches API)
    uint32_t* scores;
    int lane;
    __device__ __forceinline__ void on_event(uint32_t pos, uint64_t set) {
        if (W1) {
            if ((set >> lane) & 1) scores[lane] = pos;
            if ((set >> (lane + 32)) & 1) scores[lane + 32] = pos;
        } else {
            uint64_t w = set;
            while (w) { int b = __ffsll((long long)w) - 1; w &= w - 1; scores[lane * 64 + b] = pos; }
        }
    }
};

// linear_scan_drain (event_tree.rs:582-657) over positions lo..=hi.  All positions of a 32-wide chunk are probed
// in parallel (one lane each); positions that provably leave `active` untouched (Single entry whose set is at
// least as large as the active set — the reference's retain guard) are skipped by ballot, the rest are resolved
// in order.
template <bool W1, class Sink>
__device__ __forceinline__ void scan_drain(const EventIndexView& v, const uint64_t* ch, int lo, int hi, int lane, uint64_t& active,
                                           Sink& sink, PrefixCache& pc) {
    for (int base = lo; base <= hi; base += 32) {
        if (!set_any<W1>(active)) break;
        int p = base + lane;
        bool in_range = p <= hi;
        Slot my{0, 0, SLOT_EMPTY, 0, 0};
        bool found = false;
        uint32_t cnt = 0;
        if (in_range) {
            found = probe(v, (uint32_t)p, ch[p], my);
            if (found && my.state == SLOT_SINGLE) cnt = lane_popc_set<W1>(v, my.payload);
        }
        unsigned remaining = __ballot_sync(FULL, in_range);
        while (remaining) {
            if (!set_any<W1>(active)) break;
            uint32_t nact = set_popc<W1>(active);
            bool noop = found && my.state == SLOT_SINGLE && cnt >= nact;
            unsigned b = __ballot_sync(FULL, in_range && !noop) & remaining;
            if (!b) break;
            int k = __ffs((int)b) - 1;
            int kp = base + k;
            bool kfound = __shfl_sync(FULL, (int)found, k) != 0;
            if (!kfound) { sink.on_event((uint32_t)kp, active); active = 0; break; }
            uint32_t kstate = __shfl_sync(FULL, my.state, k);
            uint64_t kpayload = shfl64(my.payload, k);
            uint64_t ws;
            if (!entry_set<W1>(v, kstate, kpayload, ch, kp, lane, pc, ws)) { sink.on_event((uint32_t)kp, active); active = 0; break; }
            if (set_popc<W1>(ws) < nact) {
                uint64_t drained = active & ~ws;
                sink.on_event((uint32_t)kp, drained);
                active &= ws;
            }
            remaining &= ~((2u << k) - 1u);
        }
    }
}

// jump_search_matches (event_tree.rs:659-753).  `ch[0..len)` content hashes (shared memory), len ≥ 1.
// On return the survivors (score = len) are returned; drained workers were reported through the sink.
//
// The jump destinations 0, J, 2J, …, len-1 do not depend on the data (current_pos always advances to next_pos,
// :721/:733), so they are probed SPECULATIVELY in parallel — one lane per destination, 32 at a time — and the
// count-only jump test (:720) is then replayed in order from registers.  Only a failed test costs a second
// memory round trip (scan_drain).  For len ≤ J+1 (BASELINE config 2: 32 blocks, jump 64) that is exactly the
// reference's two probes, issued concurrently.
template <bool W1, class Sink>
__device__ __forceinline__ uint64_t jump_search(const EventIndexView& v, const uint64_t* ch, int len, int lane, Sink& sink, bool early_exit) {
    PrefixCache pc{-1, 0};
    const uint64_t J = v.jump;
    const int last = len - 1;
    const int m = last == 0 ? 0 : (int)(((uint64_t)last + J - 1) / J);  // destinations after position 0
    uint64_t active = 0;
    int cur = 0;
    for (int base = 0; base <= m; base += 32) {
        const int di = base + lane;
        const bool in_range = di <= m;
        int dpos = 0;
        if (in_range) { uint64_t q = (uint64_t)di * J; dpos = q < (uint64_t)last ? (int)q : last; }
        Slot my{0, 0, SLOT_EMPTY, 0, 0};
        bool found = false;
        uint32_t cnt = 0;
        if (in_range) {
            found = probe(v, (uint32_t)dpos, ch[dpos], my);
            if (!W1 && found && my.state == SLOT_SINGLE) cnt = lane_popc_set<W1>(v, my.payload);
        }
        const int nloc = (m - base + 1) < 32 ? (m - base + 1) : 32;
        for (int k = 0; k < nloc; ++k) {
            const int kd = base + k;
            int kpos;
            { uint64_t q = (uint64_t)kd * J; kpos = q < (uint64_t)last ? (int)q : last; }
            const bool kfound = __shfl_sync(FULL, (int)found, k) != 0;
            const uint32_t kstate = __shfl_sync(FULL, my.state, k);
            const uint64_t kpayload = shfl64(my.payload, k);
            if (kd == 0) {
                if (!kfound) return 0;
                if (!entry_set<W1>(v, kstate, kpayload, ch, 0, lane, pc, active)) return 0;
                if (!set_any<W1>(active)) return 0;
                if (early_exit) { sink.on_event(1, active); return 0; }
                continue;
            }
            uint32_t count = 0;  // count_workers_at(kpos) (:555-574)
            if (kfound) {
                if (kstate == SLOT_SINGLE) count = W1 ? (uint32_t)__popcll(kpayload) : __shfl_sync(FULL, cnt, k);
                else { uint64_t ws; if (entry_set<W1>(v, kstate, kpayload, ch, kpos, lane, pc, ws)) count = set_popc<W1>(ws); }
            }
            if (count != set_popc<W1>(active)) scan_drain<W1>(v, ch, cur + 1, kpos, lane, active, sink, pc);
            cur = kpos;
            if (!set_any<W1>(active)) return 0;
        }
    }
    return active;
}

// ---- K2b part 1: content hashes of a request's blocks into shared memory -----------------------------------
__device__ __forceinline__ void hash_blocks(const uint32_t* __restrict__ tok, uint32_t nb, uint32_t bs, uint64_t* ch, int lane) {
    for (uint32_t b = lane; b < nb; b += 32) {
        const uint32_t* p = tok + (size_t)b * bs;
        uint64_t h;
        if (bs == 16) {
            uint32_t w[16];
            if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
                for (int i = 0; i < 4; ++i) { uint4 t = __ldg(q + i); w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w; }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = __ldg(p + i);
            }
            h = xxh3_16words(w, kSeed);
        } else {
            h = xxh3_words(p, bs, kSeed);
        }
        ch[b] = h;
    }
}

// ---- K3: max_by_key((score, Reverse(load), Reverse(tree_size))) with LAST max = highest slice index ---------
// Per-worker scalars of the fleet snapshot.  W1 (≤ 64 interned workers): lane l keeps workers l and l+32 in
// registers, so the final pick costs shuffles only; otherwise they are read from the (L1/L2-resident) id-space arrays.
struct FleetRegs { int32_t sl0, sl1; uint64_t ld0, ld1, ts0, ts1; };

template <bool W1>
__device__ __forceinline__ int32_t arg_best(const EventIndexView& v, const FleetView& f, const FleetRegs& fr, uint64_t winset, int lane) {
    bool have = false;
    uint64_t bl = 0, bt = 0;
    int32_t bs = -1;
    auto consider = [&](int32_t sl, uint64_t ld, uint64_t ts) {
        bool better = !have || ld < bl || (ld == bl && (ts < bt || (ts == bt && sl > bs)));
        if (better) { have = true; bl = ld; bt = ts; bs = sl; }
    };
    if (W1) {
        if (__popcll(winset) == 1) {
            int id = __ffsll((long long)winset) - 1;
            int32_t a = __shfl_sync(FULL, fr.sl0, id & 31), b = __shfl_sync(FULL, fr.sl1, id & 31);
            return id < 32 ? a : b;
        }
        if ((winset >> lane) & 1) consider(fr.sl0, fr.ld0, fr.ts0);
        if ((winset >> (lane + 32)) & 1) consider(fr.sl1, fr.ld1, fr.ts1);
    } else {
        uint64_t w = winset;
        while (w) {
            int b = __ffsll((long long)w) - 1;
            w &= w - 1;
            uint32_t id = (uint32_t)(lane * 64 + b);
            consider(f.slice_of_id[id], f.load_of_id[id], v.tree_sizes[id]);
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        bool oh = __shfl_xor_sync(FULL, (int)have, d) != 0;
        uint64_t ol = shfl64_xor(bl, d), ot = shfl64_xor(bt, d);
        int32_t os = __shfl_xor_sync(FULL, bs, d);
        bool better = oh && (!have || ol < bl || (ol == bl && (ot < bt || (ot == bt && os > bs))));
        if (better) { have = true; bl = ol; bt = ot; bs = os; }
    }
    return bs;
}

// One request, one warp: the whole pick.
template <bool W1>
__device__ __forceinline__ void select_one(const EventIndexView& v, const FleetView& f, const FleetDerived& fd, const FleetRegs& fr, uint64_t elig,
                                           const uint32_t* __restrict__ tokens, const uint32_t* __restrict__ offsets, uint32_t r,
                                           uint32_t block_size, uint32_t max_blocks, uint64_t* ch, int lane, int32_t* out_idx,
                                           smgx_decision_info* out_info, uint32_t* err_flag) {
    const uint32_t off = offsets[r], ntok = offsets[r + 1] - off;
    int32_t out = -1;
    uint32_t branch = SMGX_BR_NO_HEALTHY, matched = 0;
    if (fd.n_healthy == 0) {
        // None
    } else if (fd.imbalanced) {
        out = fd.min_load_idx;
        branch = SMGX_BR_IMBALANCED_MIN_LOAD;
    } else {
        const uint32_t nb = block_size ? ntok / block_size : 0;
        if (nb > max_blocks) {
            if (lane == 0) atomicExch(err_flag, 1u);
            branch = 255;
        } else {
            uint64_t winset = 0;
            uint32_t score = 0;
            if (nb > 0 && v.n_workers > 0) {
                hash_blocks(tokens + off, nb, block_size, ch, lane);
                __syncwarp();
                SelectSink<W1> sink{elig, 0, 0};
                uint64_t surv = jump_search<W1>(v, ch, (int)nb, lane, sink, false) & elig;
                if (set_any<W1>(surv)) { winset = surv; score = nb; }
                else { winset = sink.last; score = sink.last_score; }
                __syncwarp();
            }
            if (set_any<W1>(winset)) {
                out = arg_best<W1>(v, f, fr, winset, lane);
                branch = SMGX_BR_EVENT_OVERLAP;
                matched = score;
            } else {
                out = fd.min_load_idx;
                branch = SMGX_BR_EVENT_MIN_LOAD;
            }
        }
    }
    if (lane == 0) {
        out_idx[r] = out;
        if (out_info) {
            smgx_decision_info di;
            di.matched = matched; di.input = ntok; di.branch = (uint8_t)branch;
            di.reserved[0] = di.reserved[1] = di.reserved[2] = 0;
            out_info[r] = di;
        }
    }
}

template <bool W1>
__device__ __forceinline__ FleetRegs load_fleet_regs(const EventIndexView& v, const FleetView& f, int lane) {
    FleetRegs fr{-1, -1, 0, 0, 0, 0};
    if (W1) {
        if ((uint32_t)lane < v.n_workers) { fr.sl0 = f.slice_of_id[lane]; fr.ld0 = f.load_of_id[lane]; fr.ts0 = v.tree_sizes[lane]; }
        if ((uint32_t)lane + 32 < v.n_workers) { fr.sl1 = f.slice_of_id[lane + 32]; fr.ld1 = f.load_of_id[lane + 32]; fr.ts1 = v.tree_sizes[lane + 32]; }
    }
    return fr;
}

template <bool W1>
__global__ void __launch_bounds__(256) event_select_kernel(EventIndexView v, FleetView f, SelectArgs a) {
    extern __shared__ uint64_t smem_ch[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    uint64_t* ch = smem_ch + (size_t)wic * a.max_blocks;
    const FleetDerived fd = *f.derived;
    const uint64_t elig = W1 ? f.elig[0] : ((uint32_t)lane < v.words ? f.elig[lane] : 0ULL);
    const FleetRegs fr = load_fleet_regs<W1>(v, f, lane);
    for (uint32_t r = blockIdx.x * wpc + wic; r < a.n; r += gridDim.x * wpc)
        select_one<W1>(v, f, fd, fr, elig, a.tokens, a.offsets, r, a.block_size, a.max_blocks, ch, lane, a.out_idx, a.out_info, a.err_flag);
}

// blockIdx.y = batch
template <bool W1>
__global__ void __launch_bounds__(256) event_select_multi_kernel(EventIndexView v, FleetView f, const __grid_constant__ MultiArgs a) {
    extern __shared__ uint64_t smem_ch[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    uint64_t* ch = smem_ch + (size_t)wic * a.max_blocks;
    const FleetDerived fd = *f.derived;
    const uint64_t elig = W1 ? f.elig[0] : ((uint32_t)lane < v.words ? f.elig[lane] : 0ULL);
    const FleetRegs fr = load_fleet_regs<W1>(v, f, lane);
    const BatchDesc& b = a.b[blockIdx.y];
    for (uint32_t r = blockIdx.x * wpc + wic; r < b.n; r += gridDim.x * wpc)
        select_one<W1>(v, f, fd, fr, elig, b.tokens, b.offsets, r, a.block_size, a.max_blocks, ch, lane, b.out_idx, b.out_info, a.err_flag);
}

template <bool W1>
__global__ void find_matches_kernel(EventIndexView v, const uint64_t* __restrict__ hashes, uint32_t n, int early_exit, uint32_t* scores) {
    extern __shared__ uint64_t smem_ch[];
    const int lane = threadIdx.x & 31;
    for (uint32_t i = lane; i < n; i += 32) smem_ch[i] = hashes[i];
    __syncwarp();
    if (n == 0 || v.n_workers == 0) return;
    DumpSink<W1> sink{scores, lane};
    uint64_t surv = jump_search<W1>(v, smem_ch, (int)n, lane, sink, early_exit != 0);
    sink.on_event(n, surv);
}

__global__ void content_hashes_kernel(const uint32_t* __restrict__ tokens, uint32_t n_tokens, uint32_t bs, uint64_t* out) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nb = bs ? n_tokens / bs : 0;
    if (b < nb) out[b] = xxh3_words(tokens + (size_t)b * bs, bs, kSeed);
}

__global__ void fill_kernel(uint32_t* d, uint32_t value, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) d[i] = value;
}

// select_worker prologue on the device: one CTA, the fleet is at most a few thousand workers.
__global__ void __launch_bounds__(256) fleet_prepare_kernel(FleetRaw raw, FleetDerived* out, int32_t* slice_of_id, uint64_t* load_of_id,
                                                            unsigned long long* elig) {
    __shared__ uint64_t s_mn[256], s_mx[256], s_hl[256];
    __shared__ int32_t s_hidx[256], s_fh[256];
    __shared__ uint32_t s_nh[256];
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < raw.n_ids; i += 256) { slice_of_id[i] = -1; load_of_id[i] = 0; }
    for (uint32_t i = tid; i < raw.words; i += 256) elig[i] = 0;
    __syncthreads();
    uint64_t mn = ~0ULL, mx = 0, hl = ~0ULL;
    int32_t hidx = 0x7fffffff, fh = 0x7fffffff;
    uint32_t nh = 0;
    for (uint32_t i = tid; i < raw.n_slice; i += 256) {
        uint64_t load = raw.loads[i];
        mn = load < mn ? load : mn;      // min/max over ALL workers, healthy or not (cache_aware.rs:662-666)
        mx = load > mx ? load : mx;
        if ((raw.flags[i] & 3) == 3) {   // is_healthy() && circuit_breaker_can_execute()  (mod.rs:137-144)
            ++nh;
            fh = (int32_t)i < fh ? (int32_t)i : fh;
            if (load < hl || (load == hl && (int32_t)i < hidx)) { hl = load; hidx = (int32_t)i; }
            int32_t id = raw.id_of_slice[i];
            if (id >= 0 && (uint32_t)id < raw.n_ids) {
                atomicOr(&elig[id >> 6], 1ULL << (id & 63));
                slice_of_id[id] = (int32_t)i;
                load_of_id[id] = load;
            }
        }
    }
    s_mn[tid] = mn; s_mx[tid] = mx; s_hl[tid] = hl; s_hidx[tid] = hidx; s_fh[tid] = fh; s_nh[tid] = nh;
    __syncthreads();
    for (int d = 128; d; d >>= 1) {
        if (tid < d) {
            s_mn[tid] = s_mn[tid + d] < s_mn[tid] ? s_mn[tid + d] : s_mn[tid];
            s_mx[tid] = s_mx[tid + d] > s_mx[tid] ? s_mx[tid + d] : s_mx[tid];
            if (s_hl[tid + d] < s_hl[tid] || (s_hl[tid + d] == s_hl[tid] && s_hidx[tid + d] < s_hidx[tid])) {
                s_hl[tid] = s_hl[tid + d]; s_hidx[tid] = s_hidx[tid + d];
            }
            s_fh[tid] = s_fh[tid + d] < s_fh[tid] ? s_fh[tid + d] : s_fh[tid];
            s_nh[tid] += s_nh[tid + d];
        }
        __syncthreads();
    }
    if (tid == 0) {
        FleetDerived d;
        uint64_t mn0 = s_mn[0] == ~0ULL ? 0 : s_mn[0], mx0 = s_mx[0];
        d.min_load = mn0; d.max_load = mx0;
        d.n_healthy = s_nh[0];
        d.min_load_idx = s_nh[0] ? s_hidx[0] : -1;
        d.first_healthy = s_nh[0] ? s_fh[0] : -1;
        // usize→f32 casts round to nearest-even; the product is an f32 multiply (no FMA contraction possible here)
        float fmax = __ull2float_rn(mx0), fmin = __ull2float_rn(mn0);
        d.imbalanced = ((mx0 - mn0) > raw.abs_threshold && fmax > __fmul_rn(fmin, raw.rel_threshold)) ? 1u : 0u;
        *out = d;
    }
}

}  // namespace

void launch_fleet_prepare(const FleetRaw& raw, FleetDerived* d_derived, int32_t* d_slice_of_id, uint64_t* d_load_of_id, uint64_t* d_elig,
                          cudaStream_t stream) {
    fleet_prepare_kernel<<<1, 256, 0, stream>>>(raw, d_derived, d_slice_of_id, d_load_of_id, reinterpret_cast<unsigned long long*>(d_elig));
    SMGX_CUDA(cudaGetLastError());
}

void launch_event_select(const EventIndexView& ix, const FleetView& fleet, const SelectArgs& a, int sm_count, cudaStream_t stream) {
    if (a.n == 0) return;
    // warps per CTA bounded by the shared-memory scratch (8 B per block per warp)
    size_t per_warp = (size_t)std::max<uint32_t>(a.max_blocks, 1) * 8;
    int wpc = 8;
    while (wpc > 1 && per_warp * wpc > 96 * 1024) wpc >>= 1;
    size_t smem = per_warp * wpc;
    if (smem > 200 * 1024) throw Error(SMGX_INVALID_ARGUMENT, "request too long for the per-warp scratch (max_tokens_per_request)");
    unsigned ctas_needed = (a.n + wpc - 1) / wpc;
    // resident CTAs per SM: 2048 threads / (wpc*32), also bounded by shared memory (227 KB)
    unsigned per_sm = std::min<unsigned>(2048 / (wpc * 32), (unsigned)std::max<size_t>(1, (220 * 1024) / std::max<size_t>(smem, 1)));
    unsigned cap = (unsigned)sm_count * std::max(1u, per_sm);
    unsigned grid = std::min(ctas_needed, cap);   // grid-stride beyond one full wave
    auto k = ix.words == 1 ? event_select_kernel<true> : event_select_kernel<false>;
    if (smem > 48 * 1024) SMGX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, wpc * 32, smem, stream>>>(ix, fleet, a);
    SMGX_CUDA(cudaGetLastError());
}

void launch_event_select_multi(const EventIndexView& ix, const FleetView& fleet, const MultiArgs& a, int sm_count, cudaStream_t stream) {
    if (a.count == 0) return;
    size_t per_warp = (size_t)std::max<uint32_t>(a.max_blocks, 1) * 8;
    int wpc = 8;
    while (wpc > 1 && per_warp * wpc > 96 * 1024) wpc >>= 1;
    size_t smem = per_warp * wpc;
    if (smem > 200 * 1024) throw Error(SMGX_INVALID_ARGUMENT, "request too long for the per-warp scratch (max_tokens_per_request)");
    uint32_t max_n = 0;
    for (uint32_t j = 0; j < a.count; ++j) max_n = std::max(max_n, a.b[j].n);
    unsigned ctas_x = std::max(1u, (max_n + wpc - 1) / wpc);
    (void)sm_count;
    auto k = ix.words == 1 ? event_select_multi_kernel<true> : event_select_multi_kernel<false>;
    if (smem > 48 * 1024) SMGX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<dim3(ctas_x, a.count), wpc * 32, smem, stream>>>(ix, fleet, a);
    SMGX_CUDA(cudaGetLastError());
}

void launch_find_matches(const EventIndexView& ix, const uint64_t* d_hashes, uint32_t n, bool early_exit, uint32_t* d_scores,
                         cudaStream_t stream) {
    size_t smem = (size_t)std::max<uint32_t>(n, 1) * 8;
    if (smem > 200 * 1024) throw Error(SMGX_INVALID_ARGUMENT, "too many content hashes for one find_matches call");
    auto k = ix.words == 1 ? find_matches_kernel<true> : find_matches_kernel<false>;
    if (smem > 48 * 1024) SMGX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<1, 32, smem, stream>>>(ix, d_hashes, n, early_exit ? 1 : 0, d_scores);
    SMGX_CUDA(cudaGetLastError());
}

void launch_content_hashes(const uint32_t* d_tokens, uint32_t n_tokens, uint32_t block_size, uint64_t* d_out, cudaStream_t stream) {
    uint32_t nb = block_size ? n_tokens / block_size : 0;
    if (!nb) return;
    content_hashes_kernel<<<(nb + 127) / 128, 128, 0, stream>>>(d_tokens, n_tokens, block_size, d_out);
    SMGX_CUDA(cudaGetLastError());
}

void launch_fill(uint32_t* d, uint32_t value, size_t n_words, cudaStream_t stream) {
    if (!n_words) return;
    unsigned grid = (unsigned)std::min<size_t>((n_words + 255) / 256, 148 * 8);
    fill_kernel<<<grid, 256, 0, stream>>>(d, value, n_words);
    SMGX_CUDA(cudaGetLastError());
}

}  // namespace smgx
