// K1 — GPU tokenizer: vocabulary tables + pre-tokenise / BPE / compaction kernels.  See tokenizer.h and bpe.cuh.
#include "tokenizer.h"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_map>

#include "unicode_tables.h"

namespace smgx {

namespace {

constexpr uint32_t kInvalidTok = 0xFFFFFFFFu;

// ---- K1a: pre-tokenisation, parallel over bytes ---------------------------------------------------------------
// Grid: blockIdx.y = request, blockIdx.x = 256-byte chunk of that request, one thread per byte.
// Special tokens first (tiktoken `encode`: the regex only ever sees the text between two specials):
//   special_candidates_kernel   does a special string start at this byte?  (longest at a position)
//   special_resolve_kernel      one warp per request: accept candidates left to right, dropping those inside an accepted span
//   flags: 0 ordinary byte, 1 piece start, 2 first byte of a special, 3 inside a special, 4 unresolved candidate
__global__ void __launch_bounds__(256) special_candidates_kernel(BpeView v, const uint8_t* __restrict__ text, const uint32_t* __restrict__ offsets,
                                                                 uint8_t* __restrict__ flags, uint32_t* __restrict__ tmp_ids,
                                                                 uint64_t* __restrict__ tmp_rk) {
    const uint32_t r = blockIdx.y, beg = offsets[r], rend = offsets[r + 1];
    const uint32_t i = beg + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rend) return;
    const uint8_t c = text[i];
    if (!((v.special_first[c >> 5] >> (c & 31)) & 1)) return;
    uint32_t len, id;
    if (special_at(v, text, i, rend, len, id)) { flags[i] = 4; tmp_ids[i] = id; tmp_rk[i] = len; }
}
__global__ void __launch_bounds__(256) special_resolve_kernel(const uint32_t* __restrict__ offsets, uint32_t n, uint8_t* __restrict__ flags,
                                                              uint32_t* __restrict__ tmp_ids, const uint64_t* __restrict__ tmp_rk,
                                                              uint32_t* __restrict__ totals) {
    const int lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t beg = offsets[r], end = offsets[r + 1];
    uint32_t skip_until = beg, count = 0;
    for (uint32_t base = beg; base < end; base += 32) {
        const uint32_t i = base + lane;
        unsigned m = __ballot_sync(0xffffffffu, i < end && flags[i] == 4);
        while (m) {
            const int k = __ffs((int)m) - 1;
            m &= m - 1;
            const uint32_t p = base + k;
            if (p < skip_until) { if (lane == 0) flags[p] = 0; continue; }   // starts inside an accepted special
            const uint32_t len = (uint32_t)tmp_rk[p];
            for (uint32_t j = lane; j < len; j += 32) {
                flags[p + j] = j == 0 ? 2 : 3;
                if (j) tmp_ids[p + j] = 0xFFFFFFFFu;
            }
            skip_until = p + len;
            ++count;
        }
        __syncwarp();
    }
    if (lane == 0) totals[r] = count;
}

// The regex scan itself.  Positions where a piece MUST start whatever came before — the end of a letter run (every
// alternative that consumes a letter stops at the end of the run or hands over to one that does), both ends of a digit
// run (digits are consumed by \p{N}{1,3} only), the start of the request and the byte after a special — cut the text
// into short independent segments; one thread per such "anchor" replays the hand-compiled regex until the next anchor
// and appends every piece it finds to a dense piece list (CTA-staged, one global atomic per CTA).
struct Piece { uint32_t start, len, req; };

__device__ __forceinline__ uint32_t prev_cp(const uint8_t* s, uint32_t i, uint32_t lo) {   // code point ending right before byte i
    uint32_t j = i - 1;
    while (j > lo && (s[j] & 0xC0) == 0x80) --j;
    uint32_t l;
    return utf8_decode(s, j, i, l);
}
__device__ __forceinline__ bool is_anchor(const uint8_t* s, uint32_t i, uint32_t beg, const uint8_t* flags, const UnicodeView& u) {
    if (i == beg) return true;
    if (flags[i - 1] >= 2) return true;                      // byte after a special token
    uint32_t l;
    const uint8_t c = char_class(utf8_decode(s, i, i + 4, l), u);   // valid UTF-8: a lead byte's continuation bytes exist
    const uint8_t pc = char_class(prev_cp(s, i, beg), u);
    return (pc == CH_LETTER && c != CH_LETTER) || (c == CH_NUMBER) != (pc == CH_NUMBER);
}
// Window form of next_piece_cl100k: the CTA's 256-byte window is classified ONCE into shared memory (phase 0 of the kernel:
// s_cls[k] = class | utf-8 length << 4 for a lead byte, kCont for a continuation byte; s_stop[k] = first byte of a special) and
// the regex then runs on classes instead of decoding and binary-searching every code point again.  Same alternatives, same order,
// same analytic resolution of the look-ahead as next_piece_cl100k; whenever a decision needs a byte beyond the window (and the
// window does not end the request) it returns kBail and the caller falls back to the global-memory matcher for that piece.
constexpr uint8_t kCont = 0x0F;
constexpr uint32_t kBail = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t next_piece_window(const uint8_t* __restrict__ s_txt, const uint8_t* __restrict__ s_cls, const uint8_t* __restrict__ s_stop,
                                                      uint32_t k, uint32_t W, bool ends_text) {
#define WIN_NEED(j) do { if ((j) >= W && !ends_text) return kBail; } while (0)
#define WIN_END(j) ((j) >= W || s_stop[(j)])
    const uint8_t c0 = s_cls[k] & 0x0F;
    const uint32_t len0 = s_cls[k] >> 4;
    const uint8_t b0 = s_txt[k];
    // 1. (?i:'s|'t|'re|'ve|'m|'ll|'d)  (+ U+017F for 's')
    if (b0 == '\'') {
        WIN_NEED(k + 1);
        if (!WIN_END(k + 1)) {
            const uint8_t b1 = s_txt[k + 1];
            if (b1 < 0x80) {
                const uint32_t c1 = b1 | 0x20;
                if (c1 == 's' || c1 == 't' || c1 == 'm' || c1 == 'd') return k + 2;
                WIN_NEED(k + 2);
                if (!WIN_END(k + 2) && s_txt[k + 2] < 0x80) {
                    const uint32_t c2 = s_txt[k + 2] | 0x20;
                    if ((c1 == 'r' && c2 == 'e') || (c1 == 'v' && c2 == 'e') || (c1 == 'l' && c2 == 'l')) return k + 3;
                }
            } else {
                WIN_NEED(k + 2);
                if (k + 2 < W && b1 == 0xC5 && s_txt[k + 2] == 0xBF) return k + 3;   // 'ſ
            }
        }
    }
    const bool nl0 = b0 == '\n' || b0 == '\r';
    // 2. [^\r\n\p{L}\p{N}]?\p{L}+
    {
        uint32_t j = k;
        bool ok = c0 == CH_LETTER;
        if (!ok && c0 != CH_NUMBER && !nl0) {
            const uint32_t k1 = k + len0;
            WIN_NEED(k1);
            if (!WIN_END(k1) && (s_cls[k1] & 0x0F) == CH_LETTER) { j = k1; ok = true; }
        }
        if (ok) {
            for (;;) {
                WIN_NEED(j);
                if (WIN_END(j) || (s_cls[j] & 0x0F) != CH_LETTER) break;
                j += s_cls[j] >> 4;
            }
            return j;
        }
    }
    // 3. \p{N}{1,3}
    if (c0 == CH_NUMBER) {
        uint32_t j = k + len0;
        for (int c = 1; c < 3; ++c) {
            WIN_NEED(j);
            if (WIN_END(j) || (s_cls[j] & 0x0F) != CH_NUMBER) break;
            j += s_cls[j] >> 4;
        }
        return j;
    }
    // 4.  ?[^\s\p{L}\p{N}]+[\r\n]*
    {
        uint32_t j = k;
        bool ok = c0 == CH_OTHER;
        if (!ok && b0 == ' ') {
            WIN_NEED(k + 1);
            if (!WIN_END(k + 1) && (s_cls[k + 1] & 0x0F) == CH_OTHER) { j = k + 1; ok = true; }
        }
        if (ok) {
            for (;;) {
                WIN_NEED(j);
                if (WIN_END(j) || (s_cls[j] & 0x0F) != CH_OTHER) break;
                j += s_cls[j] >> 4;
            }
            for (;;) {
                WIN_NEED(j);
                if (WIN_END(j) || !(s_txt[j] == '\n' || s_txt[j] == '\r')) break;
                ++j;
            }
            return j;
        }
    }
    // 5-7. whitespace run starting at k
    uint32_t j = k, last_nl_end = 0, last_start = k, count = 0;
    for (;;) {
        WIN_NEED(j);
        if (WIN_END(j) || (s_cls[j] & 0x0F) != CH_SPACE) break;
        last_start = j;
        const bool nl = s_txt[j] == '\n' || s_txt[j] == '\r';
        j += s_cls[j] >> 4;
        ++count;
        if (nl) last_nl_end = j;
    }
    if (last_nl_end) return last_nl_end;   // 5. \s*[\r\n]+
    if (WIN_END(j)) return j;              // 6. \s+(?!\S): the run reaches the end of the (slice of) text
    if (count >= 2) return last_start;     // 6. give back one char so that whitespace follows
    return j;                              // 7. \s+
#undef WIN_NEED
#undef WIN_END
}

__global__ void __launch_bounds__(256) pretokenize_kernel(BpeView v, const uint8_t* __restrict__ text, const uint32_t* __restrict__ offsets,
                                                          uint8_t* __restrict__ flags, Piece* __restrict__ pieces, uint32_t* __restrict__ n_pieces) {
    __shared__ Piece s_list[320];
    __shared__ uint32_t s_anchor[256];
    __shared__ uint32_t s_wcnt[8];
    __shared__ uint32_t s_count, s_base;
    __shared__ uint8_t s_txt[256], s_cls[256], s_stop[256], s_flag[256];
    if (threadIdx.x == 0) s_count = 0;
    const uint32_t r = blockIdx.y, beg = offsets[r], rend = offsets[r + 1];
    const uint32_t cstart = beg + blockIdx.x * blockDim.x;
    const uint32_t i = cstart + threadIdx.x;
    const uint32_t W = cstart < rend ? min(256u, rend - cstart) : 0;
    const bool ends_text = cstart + W == rend;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // phase 0 (every thread, one byte each): classify the window once
    if (threadIdx.x < W) {
        const uint8_t b = text[i], fl = flags[i];
        s_txt[threadIdx.x] = b;
        s_flag[threadIdx.x] = fl;
        s_stop[threadIdx.x] = fl == 2;
        if ((b & 0xC0) == 0x80) s_cls[threadIdx.x] = kCont;
        else { uint32_t l; const uint32_t cp = utf8_decode(text, i, rend, l); s_cls[threadIdx.x] = (uint8_t)(char_class(cp, v.uni) | (l << 4)); }
    }
    __syncthreads();
    // anchor test on the window (is_anchor on classes); the char before the window is classified from global memory
    auto anchor_at = [&](uint32_t k) -> bool {   // k < W, lead byte, not inside a special
        const uint32_t gi = cstart + k;
        if (gi == beg) return true;
        if ((k ? s_flag[k - 1] : flags[gi - 1]) >= 2) return true;
        const uint8_t c = s_cls[k] & 0x0F;
        uint8_t pc;
        int q = (int)k - 1;
        while (q >= 0 && s_cls[q] == kCont) --q;
        if (q >= 0) pc = s_cls[q] & 0x0F;
        else pc = char_class(prev_cp(text, gi, beg), v.uni);
        return (pc == CH_LETTER && c != CH_LETTER) || (c == CH_NUMBER) != (pc == CH_NUMBER);
    };
    // phase 1: which bytes are anchors?  They are then packed densely, so that phase 2 — the serial regex replay — runs with full
    // warps instead of one live lane in six.
    const bool anchor = threadIdx.x < W && s_cls[threadIdx.x] != kCont && s_flag[threadIdx.x] < 2 && anchor_at(threadIdx.x);
    const unsigned bal = __ballot_sync(0xffffffffu, anchor);
    if (lane == 0) s_wcnt[wid] = (uint32_t)__popc(bal);
    __syncthreads();
    uint32_t before = 0, n_anchor = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t c = s_wcnt[w]; if (w < wid) before += c; n_anchor += c; }
    if (anchor) s_anchor[before + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = threadIdx.x;
    __syncthreads();
    // phase 2: thread k replays the regex from the k-th anchor up to the next one; a special's first byte (flags == 2) ends the
    // slice.  The warp iterates in lockstep (one piece per live lane and round).
    if ((uint32_t)(wid * 32) < n_anchor) {
        bool live = threadIdx.x < n_anchor;
        uint32_t p = live ? cstart + s_anchor[threadIdx.x] : 0;
        while (__any_sync(0xffffffffu, live)) {
            if (live) {
                flags[p] = 1;
                uint32_t q = kBail;
                if (p - cstart < W) { const uint32_t e = next_piece_window(s_txt, s_cls, s_stop, p - cstart, W, ends_text); if (e != kBail) q = cstart + e; }
                if (q == kBail) q = next_piece_cl100k(text, p, rend, v.uni, flags);   // the piece looks beyond the window
                const uint32_t slot = atomicAdd(&s_count, 1u);
                if (slot < 320) s_list[slot] = Piece{p, q - p, r};
                else { const uint32_t g = atomicAdd(n_pieces, 1u); pieces[g] = Piece{p, q - p, r}; }   // overflow: straight to the global list
                bool stop;
                if (q >= rend) stop = true;
                else if (q - cstart < W) stop = s_flag[q - cstart] == 2 || anchor_at(q - cstart);
                else stop = flags[q] == 2 || is_anchor(text, q, beg, flags, v.uni);
                if (stop) live = false; else p = q;
            }
        }
    }
    __syncthreads();
    const uint32_t cnt = s_count < 320 ? s_count : 320;
    if (threadIdx.x == 0 && cnt) s_base = atomicAdd(n_pieces, cnt);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) pieces[s_base + k] = s_list[k];
}

// K1b: BPE over the dense piece list, three kernels so that every one of them runs on full warps:
//   bpe_lookup_kernel   thread per piece: whole-piece vocabulary hit → one token (most pieces of natural text); misses are appended to
//                       two lists — pieces of ≤ kSmemSyms bytes and longer ones (warp-aggregated atomics)
//   bpe_merge_kernel    thread per short miss: tiktoken's byte-pair merge with symbols and pair ranks in shared memory (column per
//                       thread); the warp runs the merge loop in lockstep, one merge per live lane and round
//   bpe_long_kernel     warp per long miss (indentation runs, rules of dashes, URLs …): up to 32 bytes one symbol per lane, the
//                       leftmost-lowest-rank pair found with one redux.sync per merge, the tail shifted down by shuffles; up to 256
//                       bytes a linked list in shared memory with a strided minimum scan; beyond that (very rare) lane 0 merges in
//                       place in the global scratch
// Unused slots of a piece are set INVALID for the compaction.  counters: [0] pieces, [1] short misses, [2] long misses.
constexpr uint32_t kSmemSyms = 15;
constexpr uint32_t kLongSyms = 256;   // longest piece the warp-per-piece kernel keeps in shared memory
__global__ void __launch_bounds__(256) bpe_lookup_kernel(BpeView v, const uint8_t* __restrict__ text, const Piece* __restrict__ pieces,
                                                         uint32_t* __restrict__ counters, uint32_t* __restrict__ tmp_ids, uint32_t* __restrict__ totals,
                                                         uint32_t* __restrict__ miss_short, uint32_t* __restrict__ miss_long) {
    const uint32_t np = counters[0];
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t t0 = blockIdx.x * blockDim.x; t0 < np; t0 += stride) {
        const uint32_t t = t0 + threadIdx.x;
        int kind = 0;   // 1 short miss, 2 long miss
        if (t < np) {
            const Piece pc = pieces[t];
            uint32_t id;
            if (v.whole_piece && piece_lookup(v, text + pc.start, pc.len, id)) {
                tmp_ids[pc.start] = id;
                for (uint32_t j = pc.start + 1; j < pc.start + pc.len; ++j) tmp_ids[j] = kInvalidTok;
                atomicAdd(&totals[pc.req], 1u);
            } else kind = pc.len <= kSmemSyms ? 1 : 2;
        }
#pragma unroll
        for (int k = 1; k <= 2; ++k) {
            const unsigned bal = __ballot_sync(0xffffffffu, kind == k);
            if (!bal) continue;
            uint32_t base = 0;
            if (lane == __ffs((int)bal) - 1) base = atomicAdd(&counters[k], (uint32_t)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, __ffs((int)bal) - 1);
            if (kind == k) (k == 1 ? miss_short : miss_long)[base + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = t;
        }
    }
}

__global__ void __launch_bounds__(256) bpe_merge_kernel(BpeView v, const uint8_t* __restrict__ text, const Piece* __restrict__ pieces,
                                                        const uint32_t* __restrict__ counters, const uint32_t* __restrict__ miss_short,
                                                        uint32_t* __restrict__ tmp_ids, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_id[kSmemSyms][256];     // symbol ids
    __shared__ uint32_t s_rank[kSmemSyms][256];   // rank of the pair (i, i+1); kRankMax = no merge
    __shared__ uint32_t s_mid[kSmemSyms][256];    // id that pair merges into
    const uint32_t n = counters[1], tid = threadIdx.x;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t t0 = blockIdx.x * blockDim.x; t0 < n; t0 += stride) {   // t0 is CTA-uniform: whole warps stay in the loop together
        const bool valid = t0 + tid < n;
        Piece pc{0, 0, 0};
        if (valid) pc = pieces[miss_short[t0 + tid]];
        const uint32_t i = pc.start, plen = pc.len;
        uint32_t m = plen;
        if (valid) {
            for (uint32_t j = 0; j < plen; ++j) s_id[j][tid] = v.byte_token[text[i + j]];
            for (uint32_t j = 0; j + 1 < plen; ++j) {
                const uint64_t pr = pair_lookup(v, s_id[j][tid], s_id[j + 1][tid]);
                s_rank[j][tid] = (uint32_t)(pr >> 32); s_mid[j][tid] = (uint32_t)pr;
            }
        }
        bool merging = valid && plen >= 2;
        while (__any_sync(0xffffffffu, merging)) {
            if (merging) {
                uint32_t best = kRankMax, bi = 0;
                for (uint32_t j = 0; j + 1 < m; ++j) { const uint32_t rk = s_rank[j][tid]; if (rk < best) { best = rk; bi = j; } }   // leftmost minimum
                if (best == kRankMax) merging = false;
                else {
                    s_id[bi][tid] = s_mid[bi][tid];
                    for (uint32_t j = bi + 1; j + 1 < m; ++j) { s_id[j][tid] = s_id[j + 1][tid]; s_rank[j][tid] = s_rank[j + 1][tid]; s_mid[j][tid] = s_mid[j + 1][tid]; }
                    --m;
                    if (bi + 1 < m) { const uint64_t pr = pair_lookup(v, s_id[bi][tid], s_id[bi + 1][tid]); s_rank[bi][tid] = (uint32_t)(pr >> 32); s_mid[bi][tid] = (uint32_t)pr; }
                    else s_rank[bi][tid] = kRankMax;
                    if (bi > 0) { const uint64_t pr = pair_lookup(v, s_id[bi - 1][tid], s_id[bi][tid]); s_rank[bi - 1][tid] = (uint32_t)(pr >> 32); s_mid[bi - 1][tid] = (uint32_t)pr; }
                    if (m < 2) merging = false;
                }
            }
        }
        if (valid) {
            for (uint32_t j = 0; j < m; ++j) tmp_ids[i + j] = s_id[j][tid];
            for (uint32_t j = i + m; j < i + plen; ++j) tmp_ids[j] = kInvalidTok;
            atomicAdd(&totals[pc.req], m);
        }
    }
}

__global__ void __launch_bounds__(256) bpe_long_kernel(BpeView v, const uint8_t* __restrict__ text, const Piece* __restrict__ pieces,
                                                       const uint32_t* __restrict__ counters, const uint32_t* __restrict__ miss_long,
                                                       uint32_t* __restrict__ tmp_ids, uint64_t* __restrict__ tmp_rk, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_id[8][kLongSyms], s_rank[8][kLongSyms], s_mid[8][kLongSyms];
    __shared__ uint16_t s_nxt[8][kLongSyms], s_prv[8][kLongSyms];
    const uint32_t n = counters[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t nw = gridDim.x * (blockDim.x >> 5);
    for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += nw) {
        const Piece pc = pieces[miss_long[k]];
        const uint32_t i = pc.start, plen = pc.len;
        uint32_t m;
        if (plen <= 32) {
            m = plen;
            uint32_t id = (uint32_t)lane < plen ? v.byte_token[text[i + lane]] : 0;
            uint32_t nid = __shfl_down_sync(0xffffffffu, id, 1);
            uint32_t rank = kRankMax, mid = 0;
            if ((uint32_t)lane + 1 < m) { const uint64_t pr = pair_lookup(v, id, nid); rank = (uint32_t)(pr >> 32); mid = (uint32_t)pr; }
            for (;;) {
                // ranks are vocabulary positions (< 2^27), so rank * 32 + lane orders by rank, then by position
                const uint32_t key = ((uint32_t)lane + 1 < m && rank != kRankMax) ? ((rank << 5) | (uint32_t)lane) : 0xFFFFFFFFu;
                const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
                if (kmin == 0xFFFFFFFFu) break;
                const int bi = (int)(kmin & 31u);
                if (lane == bi) id = mid;
                const uint32_t id_dn = __shfl_down_sync(0xffffffffu, id, 1), rank_dn = __shfl_down_sync(0xffffffffu, rank, 1), mid_dn = __shfl_down_sync(0xffffffffu, mid, 1);
                if (lane > bi) { id = id_dn; rank = rank_dn; mid = mid_dn; }
                --m;
                nid = __shfl_down_sync(0xffffffffu, id, 1);
                if (lane == bi || lane == bi - 1) {
                    rank = kRankMax;
                    if ((uint32_t)lane + 1 < m) { const uint64_t pr = pair_lookup(v, id, nid); rank = (uint32_t)(pr >> 32); mid = (uint32_t)pr; }
                }
            }
            if ((uint32_t)lane < plen) tmp_ids[i + lane] = (uint32_t)lane < m ? id : kInvalidTok;
        } else if (plen <= kLongSyms) {
            // 33 … kLongSyms bytes: symbols stay in place in shared memory, linked by next/prev indices (list order = index order, so
            // "leftmost" = smallest index); a merge unlinks one symbol and re-ranks two pairs; the minimum is a strided scan + 2 redux
            uint32_t* w_id = s_id[wid]; uint32_t* w_rank = s_rank[wid]; uint32_t* w_mid = s_mid[wid];
            uint16_t* w_nxt = s_nxt[wid]; uint16_t* w_prv = s_prv[wid];
            for (uint32_t j = lane; j < plen; j += 32) { w_id[j] = v.byte_token[text[i + j]]; w_nxt[j] = (uint16_t)(j + 1); w_prv[j] = (uint16_t)(j ? j - 1 : 0xFFFF); }
            __syncwarp();
            for (uint32_t j = lane; j < plen; j += 32) {
                uint32_t rk = kRankMax, md = 0;
                if (j + 1 < plen) { const uint64_t pr = pair_lookup(v, w_id[j], w_id[j + 1]); rk = (uint32_t)(pr >> 32); md = (uint32_t)pr; }
                w_rank[j] = rk; w_mid[j] = md;
            }
            __syncwarp();
            m = plen;
            for (;;) {
                uint32_t best = kRankMax, bpos = 0xFFFFFFFFu;
                for (uint32_t j = lane; j < plen; j += 32) { const uint32_t rk = w_rank[j]; if (rk < best) { best = rk; bpos = j; } }   // ascending j: leftmost per lane
                const uint32_t rmin = __reduce_min_sync(0xffffffffu, best);
                if (rmin == kRankMax) break;
                const uint32_t b = __reduce_min_sync(0xffffffffu, best == rmin ? bpos : 0xFFFFFFFFu);
                const uint32_t c = w_nxt[b], d = w_nxt[c], a = w_prv[b];
                __syncwarp();
                if (lane == 0) {
                    w_id[b] = w_mid[b]; w_nxt[b] = (uint16_t)d; w_rank[c] = kRankMax; w_id[c] = kInvalidTok;
                    if (d < plen) w_prv[d] = (uint16_t)b;
                }
                __syncwarp();
                if (lane == 0) {
                    uint32_t rk = kRankMax, md = 0;
                    if (d < plen) { const uint64_t pr = pair_lookup(v, w_id[b], w_id[d]); rk = (uint32_t)(pr >> 32); md = (uint32_t)pr; }
                    w_rank[b] = rk; w_mid[b] = md;
                } else if (lane == 1 && a != 0xFFFF) {
                    const uint64_t pr = pair_lookup(v, w_id[a], w_id[b]);
                    w_rank[a] = (uint32_t)(pr >> 32); w_mid[a] = (uint32_t)pr;
                }
                --m;
                __syncwarp();
            }
            uint32_t written = 0;   // compact the surviving symbols to the front of the piece's slots
            for (uint32_t base = 0; base < plen; base += 32) {
                const uint32_t j = base + lane;
                const bool alive = j < plen && w_id[j] != kInvalidTok;
                const unsigned bal = __ballot_sync(0xffffffffu, alive);
                if (alive) tmp_ids[i + written + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = w_id[j];
                written += (uint32_t)__popc(bal);
            }
            for (uint32_t j = written + lane; j < plen; j += 32) tmp_ids[i + j] = kInvalidTok;
            __syncwarp();
        } else {
            m = 0;
            if (lane == 0) {
                m = byte_pair_merge(v, text + i, plen, tmp_ids + i, tmp_rk + i);
                for (uint32_t j = i + m; j < i + plen; ++j) tmp_ids[j] = kInvalidTok;
            }
            m = __shfl_sync(0xffffffffu, m, 0);
        }
        if (lane == 0) atomicAdd(&totals[pc.req], m);
    }
}

// exclusive scan of per-request token counts → token offsets (n + 1); one CTA, n ≤ a few 10^5
__global__ void __launch_bounds__(1024) scan_counts_kernel(const uint32_t* __restrict__ totals, uint32_t n, uint32_t* __restrict__ tok_offsets) {
    __shared__ uint32_t s_part[1024];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < n ? totals[i] : 0;
        s_part[threadIdx.x] = x;
        __syncthreads();
        for (uint32_t d = 1; d < 1024; d <<= 1) {
            uint32_t t = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) tok_offsets[i] = s_carry + s_part[threadIdx.x] - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry += s_part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) tok_offsets[n] = s_carry;
}

// K1c: one warp per request: stream-compact the valid token slots of its byte range into the ragged output
__global__ void __launch_bounds__(256) compact_tokens_kernel(const uint32_t* __restrict__ offsets, uint32_t n, const uint32_t* __restrict__ tmp_ids,
                                                             const uint32_t* __restrict__ tok_offsets, uint32_t* __restrict__ out_tokens) {
    const int lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t beg = offsets[r], end = offsets[r + 1];
    uint32_t out = tok_offsets[r];
    for (uint32_t base = beg; base < end; base += 32) {
        uint32_t i = base + lane;
        uint32_t v = i < end ? tmp_ids[i] : kInvalidTok;
        unsigned m = __ballot_sync(0xffffffffu, v != kInvalidTok);
        if (v != kInvalidTok) out_tokens[out + __popc(m & ((1u << lane) - 1u))] = v;
        out += __popc(m);
    }
}

uint32_t pow2_at_least(uint64_t x) { uint32_t p = 1024; while (p < x) p <<= 1; return p; }

std::string b64decode(const std::string& in) {
    static int8_t T[256];
    static bool init = false;
    if (!init) {
        memset(T, -1, sizeof(T));
        const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; ++i) T[(uint8_t)a[i]] = (int8_t)i;
        init = true;
    }
    std::string out;
    uint32_t val = 0;
    int bits = -8;
    for (unsigned char c : in) {
        if (c == '=') break;
        if (T[c] < 0) throw Error(SMGX_TOKENIZATION_ERROR, "invalid base64 in tiktoken file");
        val = (val << 6) | (uint32_t)T[c];
        bits += 6;
        if (bits >= 0) { out.push_back((char)((val >> bits) & 0xFF)); bits -= 8; }
    }
    return out;
}

}  // namespace

Tokenizer* Tokenizer::from_tiktoken_file(const std::string& path, const std::vector<std::pair<std::string, uint32_t>>& specials, bool device) {
    // load_tiktoken_bpe (tiktoken.rs:346-367): lines of `base64(token) rank`
    std::ifstream f(path);
    if (!f) throw Error(SMGX_TOKENIZATION_ERROR, "cannot open tiktoken file '" + path + "'");
    std::vector<std::string> toks;
    std::vector<uint32_t> ranks;
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        std::istringstream ls(line);
        std::string b64, rk;
        if (!(ls >> b64)) throw Error(SMGX_TOKENIZATION_ERROR, "missing token in tiktoken file");
        if (!(ls >> rk)) throw Error(SMGX_TOKENIZATION_ERROR, "missing rank in tiktoken file");
        toks.push_back(b64decode(b64));
        ranks.push_back((uint32_t)std::stoul(rk));
    }
    return new Tokenizer(toks, ranks, specials, device);
}

Tokenizer::Tokenizer(const std::vector<std::string>& tokens, const std::vector<uint32_t>& ranks,
                     const std::vector<std::pair<std::string, uint32_t>>& specials, bool device)
    : device_(device) {
    build(tokens, ranks, nullptr, true, specials);
}
Tokenizer::Tokenizer(const std::vector<std::string>& tokens, const std::vector<uint32_t>& ids, const std::vector<std::pair<uint32_t, uint32_t>>& merges,
                     bool ignore_merges, const std::vector<std::pair<std::string, uint32_t>>& specials, bool device)
    : device_(device) {
    build(tokens, ids, &merges, ignore_merges, specials);
}

void Tokenizer::build(const std::vector<std::string>& tokens, const std::vector<uint32_t>& ranks, const std::vector<std::pair<uint32_t, uint32_t>>* merges,
                      bool whole_piece, const std::vector<std::pair<std::string, uint32_t>>& specials) {
    SMGX_REQUIRE(tokens.size() == ranks.size() && !tokens.empty(), "empty vocabulary");
    std::unordered_map<std::string, uint32_t> enc;
    enc.reserve(tokens.size() * 2);
    uint32_t max_len = 1, max_id = 0;
    for (size_t i = 0; i < tokens.size(); ++i) {
        SMGX_REQUIRE(!tokens[i].empty(), "empty token in vocabulary");
        enc[tokens[i]] = ranks[i];
        max_len = std::max<uint32_t>(max_len, (uint32_t)tokens[i].size());
        max_id = std::max(max_id, ranks[i]);
    }
    for (auto& sp : specials) max_id = std::max(max_id, sp.second);
    vocab_size_ = max_id + 1;
    SMGX_REQUIRE(vocab_size_ < (1u << 27), "vocabulary too large (merge ranks are packed into 27 bits)");

    std::vector<uint32_t> byte_token(256);
    for (int b = 0; b < 256; ++b) {
        auto it = enc.find(std::string(1, (char)b));
        if (it == enc.end()) throw Error(SMGX_TOKENIZATION_ERROR, "vocabulary lacks the single-byte token " + std::to_string(b));
        byte_token[b] = it->second;
    }
    // blob of token bytes + piece table keyed by the bytes
    std::vector<uint8_t> blob;
    std::vector<PieceSlot> pieces(pow2_at_least(tokens.size() * 2), PieceSlot{0, 0, 0, 0, 0});
    uint32_t pmask = (uint32_t)pieces.size() - 1;
    for (auto& kv : enc) {
        uint32_t off = (uint32_t)blob.size();
        blob.insert(blob.end(), kv.first.begin(), kv.first.end());
        uint64_t hsh = bytes_hash((const uint8_t*)kv.first.data(), (uint32_t)kv.first.size());
        uint32_t h = (uint32_t)(hsh >> 32) & pmask;
        while (pieces[h].hash) h = (h + 1) & pmask;
        pieces[h] = PieceSlot{hsh, kv.second, (uint32_t)kv.first.size(), off, 0};
    }
    // pair table (left id, right id) → (priority, merged id)
    struct PairRec { uint64_t key; uint32_t rank, id; };
    std::vector<PairRec> pair_list;
    if (!merges) {
        // tiktoken: every split of every token into two vocab entries (rank of the concatenation = the merge priority)
        for (auto& kv : enc) {
            const std::string& t = kv.first;
            for (size_t k = 1; k < t.size(); ++k) {
                auto l = enc.find(t.substr(0, k));
                if (l == enc.end()) continue;
                auto r = enc.find(t.substr(k));
                if (r == enc.end()) continue;
                pair_list.push_back({((uint64_t)l->second << 32) | r->second, kv.second, kv.second});
            }
        }
    } else {
        // HuggingFace BPE: only the listed pairs merge, in list order; the merged token is the concatenation (models/bpe/model.rs)
        std::unordered_map<uint32_t, const std::string*> by_id;
        for (auto& kv : enc) by_id[kv.second] = &kv.first;
        for (size_t k = 0; k < merges->size(); ++k) {
            auto l = by_id.find((*merges)[k].first), r = by_id.find((*merges)[k].second);
            if (l == by_id.end() || r == by_id.end()) throw Error(SMGX_TOKENIZATION_ERROR, "merge refers to an unknown token id");
            auto t = enc.find(*l->second + *r->second);
            if (t == enc.end()) throw Error(SMGX_TOKENIZATION_ERROR, "merge result is not in the vocabulary");
            pair_list.push_back({((uint64_t)l->first << 32) | r->first, (uint32_t)k, t->second});
        }
    }
    n_pairs_ = (uint32_t)pair_list.size();
    std::vector<PairSlot> pairs(pow2_at_least(pair_list.size() * 2 + 2), PairSlot{kPairEmpty, kRankMax, 0});
    uint32_t qmask = (uint32_t)pairs.size() - 1;
    for (auto& pr : pair_list) {
        uint32_t h = (uint32_t)(mix64(pr.key) >> 32) & qmask;
        while (pairs[h].key != kPairEmpty && pairs[h].key != pr.key) h = (h + 1) & qmask;
        if (pairs[h].key == pr.key && pairs[h].rank <= pr.rank) continue;   // a repeated pair keeps its first (best) priority
        pairs[h] = PairSlot{pr.key, pr.rank, pr.id};
    }
    dview_.whole_piece = whole_piece ? 1u : 0u;
    // specials
    std::vector<SpecialTok> sp;
    uint32_t first_bits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (auto& s : specials) {
        SMGX_REQUIRE(!s.first.empty(), "empty special token");
        uint32_t off = (uint32_t)blob.size();
        blob.insert(blob.end(), s.first.begin(), s.first.end());
        sp.push_back(SpecialTok{off, (uint32_t)s.first.size(), s.second, 0});
        uint8_t c = (uint8_t)s.first[0];
        first_bits[c >> 5] |= 1u << (c & 31);
    }
    if (!device_) return;
    auto up = [](DevBuf& d, const void* src, size_t bytes) {
        d.reserve(std::max<size_t>(bytes, 16));
        if (bytes) SMGX_CUDA(cudaMemcpy(d.ptr, src, bytes, cudaMemcpyHostToDevice));
    };
    up(d_byte_token_, byte_token.data(), 256 * 4);
    up(d_pairs_, pairs.data(), pairs.size() * sizeof(PairSlot));
    up(d_pieces_, pieces.data(), pieces.size() * sizeof(PieceSlot));
    up(d_blob_, blob.data(), blob.size());
    up(d_specials_, sp.data(), sp.size() * sizeof(SpecialTok));
    up(d_uni_lo_, kUniLo, sizeof(kUniLo));
    up(d_uni_hi_, kUniHi, sizeof(kUniHi));
    up(d_uni_cls_, kUniCls, sizeof(kUniCls));
    dview_.byte_token = d_byte_token_.as<uint32_t>();
    dview_.pairs = d_pairs_.as<PairSlot>();
    dview_.pair_mask = qmask;
    dview_.pieces = d_pieces_.as<PieceSlot>();
    dview_.piece_mask = pmask;
    dview_.blob = d_blob_.as<uint8_t>();
    dview_.max_token_len = max_len;
    dview_.specials = d_specials_.as<SpecialTok>();
    dview_.n_special = (uint32_t)sp.size();
    for (int i = 0; i < 8; ++i) dview_.special_first[i] = first_bits[i];
    dview_.uni = UnicodeView{d_uni_lo_.as<uint32_t>(), d_uni_hi_.as<uint32_t>(), d_uni_cls_.as<uint8_t>(), kUniRanges};
}

Tokenizer::~Tokenizer() {
    d_byte_token_.release(); d_pairs_.release(); d_pieces_.release(); d_blob_.release(); d_specials_.release();
    d_uni_lo_.release(); d_uni_hi_.release(); d_uni_cls_.release();
}

void Tokenizer::encode_batch(const uint8_t* d_text, const uint32_t* d_offsets, uint32_t n, uint32_t first_byte, uint32_t total_bytes, uint32_t max_len,
                             uint32_t* d_tokens, uint32_t* d_tok_offsets, Scratch& sc, cudaStream_t stream, uint64_t* launches) const {
    if (!device_) throw Error(SMGX_DEVICE_ERROR, "tokenizer was loaded without a device (host-mirror policy): no GPU path, no CPU fallback");
    if (n == 0) return;
    sc.flags.reserve(std::max<uint32_t>(total_bytes, 1));
    sc.tmp_ids.reserve((size_t)std::max<uint32_t>(total_bytes, 1) * 4);
    sc.tmp_rk.reserve((size_t)std::max<uint32_t>(total_bytes, 1) * 8);
    sc.totals.reserve((size_t)n * 4);
    sc.pieces.reserve((size_t)std::max<uint32_t>(total_bytes - first_byte, 1) * sizeof(Piece));
    sc.n_pieces.reserve(16);
    sc.miss.reserve(((size_t)std::max<uint32_t>(total_bytes - first_byte, 1) + 1) * 8);
    if (total_bytes) {
        SMGX_CUDA(cudaMemsetAsync(sc.flags.ptr, 0, total_bytes, stream));
    }
    {
        const dim3 grid2((max_len + 255) / 256, n);
        if (dview_.n_special && max_len) {
            special_candidates_kernel<<<grid2, 256, 0, stream>>>(dview_, d_text, d_offsets, sc.flags.as<uint8_t>(), sc.tmp_ids.as<uint32_t>(),
                                                               sc.tmp_rk.as<uint64_t>());
            SMGX_CUDA(cudaGetLastError());
            ++*launches;
            special_resolve_kernel<<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, stream>>>(d_offsets, n, sc.flags.as<uint8_t>(), sc.tmp_ids.as<uint32_t>(),
                                                                                               sc.tmp_rk.as<uint64_t>(), sc.totals.as<uint32_t>());
            SMGX_CUDA(cudaGetLastError());
            ++*launches;
        } else {
            SMGX_CUDA(cudaMemsetAsync(sc.totals.ptr, 0, (size_t)n * 4, stream));
        }
        SMGX_CUDA(cudaMemsetAsync(sc.n_pieces.ptr, 0, 16, stream));
        if (max_len) {
            pretokenize_kernel<<<grid2, 256, 0, stream>>>(dview_, d_text, d_offsets, sc.flags.as<uint8_t>(), sc.pieces.as<Piece>(), sc.n_pieces.as<uint32_t>());
            SMGX_CUDA(cudaGetLastError());
            ++*launches;
            // the piece count lives on the device, so the grids are sized for the worst case the text allows and stride over the lists
            const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)(total_bytes - first_byte) / 2 + 255) / 256 + 1, 148ull * 16);
            uint32_t* cnt = sc.n_pieces.as<uint32_t>();
            uint32_t* miss_short = sc.miss.as<uint32_t>();
            uint32_t* miss_long = miss_short + (total_bytes - first_byte) + 1;
            bpe_lookup_kernel<<<grid, 256, 0, stream>>>(dview_, d_text, sc.pieces.as<Piece>(), cnt, sc.tmp_ids.as<uint32_t>(), sc.totals.as<uint32_t>(), miss_short, miss_long);
            bpe_merge_kernel<<<std::min(grid, 148u * 4), 256, 0, stream>>>(dview_, d_text, sc.pieces.as<Piece>(), cnt, miss_short, sc.tmp_ids.as<uint32_t>(), sc.totals.as<uint32_t>());
            bpe_long_kernel<<<std::min(grid, 148u * 8), 256, 0, stream>>>(dview_, d_text, sc.pieces.as<Piece>(), cnt, miss_long, sc.tmp_ids.as<uint32_t>(),
                                                                    sc.tmp_rk.as<uint64_t>(), sc.totals.as<uint32_t>());
            SMGX_CUDA(cudaGetLastError());
            *launches += 3;
        }
    }
    scan_counts_kernel<<<1, 1024, 0, stream>>>(sc.totals.as<uint32_t>(), n, d_tok_offsets);
    SMGX_CUDA(cudaGetLastError());
    ++*launches;
    compact_tokens_kernel<<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, stream>>>(d_offsets, n, sc.tmp_ids.as<uint32_t>(), d_tok_offsets, d_tokens);
    SMGX_CUDA(cudaGetLastError());
    ++*launches;
}

}  // namespace smgx
