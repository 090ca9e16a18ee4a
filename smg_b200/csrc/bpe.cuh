// K1 — tokenize: the cl100k pre-tokenizer and tiktoken-style byte-pair merge as host+device functions.
//
// Replaces, token for token, what the reference gets from the un-vendored crate tiktoken-rs through
//   TiktokenTokenizer::encode → CoreBPE::encode_with_special_tokens   crates/tokenizer/src/tiktoken.rs:444-462
// with the pattern CL100K_BASE_PATTERN (tiktoken.rs:28):
//   (?i:'s|'t|'re|'ve|'m|'ll|'d) | [^\r\n\p{L}\p{N}]?\p{L}+ | \p{N}{1,3} | ?[^\s\p{L}\p{N}]+[\r\n]* | \s*[\r\n]+ | \s+(?!\S) | \s+
// The regex is compiled by hand into next_piece_cl100k(): leftmost-first alternation, greedy quantifiers and the one
// backtracking look-ahead are resolved analytically (see the comments at each alternative).
#pragma once
#include <cstdint>

#include "xxh3.cuh"  // SMGX_HD

namespace smgx {

enum : uint8_t { CH_OTHER = 0, CH_LETTER = 1, CH_NUMBER = 2, CH_SPACE = 3 /* White_Space */ };

struct UnicodeView {
    const uint32_t* lo;   // sorted range starts
    const uint32_t* hi;
    const uint8_t* cls;
    uint32_t n;
};

// UTF-8 decode at s[i] (input is valid UTF-8, as `&str` guarantees in the reference); len = bytes consumed.
SMGX_HD uint32_t utf8_decode(const uint8_t* s, uint32_t i, uint32_t n, uint32_t& len) {
    uint32_t c = s[i];
    if (c < 0x80) { len = 1; return c; }
    if ((c >> 5) == 0x6 && i + 1 < n) { len = 2; return ((c & 0x1F) << 6) | (s[i + 1] & 0x3F); }
    if ((c >> 4) == 0xE && i + 2 < n) { len = 3; return ((c & 0x0F) << 12) | ((s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F); }
    if ((c >> 3) == 0x1E && i + 3 < n) { len = 4; return ((c & 0x07) << 18) | ((s[i + 1] & 0x3F) << 12) | ((s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F); }
    len = 1;
    return 0xFFFD;
}

SMGX_HD uint8_t char_class(uint32_t cp, const UnicodeView& u) {
    if (cp < 0x80) {
        if ((cp >= 'a' && cp <= 'z') || (cp >= 'A' && cp <= 'Z')) return CH_LETTER;
        if (cp >= '0' && cp <= '9') return CH_NUMBER;
        if (cp == ' ' || (cp >= 0x09 && cp <= 0x0D)) return CH_SPACE;
        return CH_OTHER;
    }
    uint32_t a = 0, b = u.n;   // first range with hi >= cp
    while (a < b) {
        uint32_t m = (a + b) >> 1;
        if (u.hi[m] < cp) a = m + 1; else b = m;
    }
    if (a < u.n && u.lo[a] <= cp) return u.cls[a];
    return CH_OTHER;
}
SMGX_HD bool is_newline(uint32_t cp) { return cp == '\n' || cp == '\r'; }

// End (byte index) of the regex match that starts at byte i of s[0..n).  i < n and i is not a slice end.
// tiktoken runs the regex only on the ordinary text between two special tokens; `stop` (nullable) marks the first byte of
// every special with the value 2, and such a byte acts as the end of the string (for the run scans and for `(?!\S)`).
SMGX_HD uint32_t next_piece_cl100k(const uint8_t* s, uint32_t i, uint32_t n, const UnicodeView& u, const uint8_t* stop = nullptr) {
#define SMGX_AT_END(j) ((j) >= n || (stop != nullptr && stop[(j)] == 2))
    uint32_t len0;
    const uint32_t cp0 = utf8_decode(s, i, n, len0);
    const uint8_t cls0 = char_class(cp0, u);

    // 1. (?i:'s|'t|'re|'ve|'m|'ll|'d) — Unicode-aware case folding adds U+017F (ſ) for 's'
    if (cp0 == '\'' && !SMGX_AT_END(i + 1)) {
        const uint32_t c1 = s[i + 1] | 0x20;
        if (s[i + 1] < 0x80) {
            if (c1 == 's' || c1 == 't' || c1 == 'm' || c1 == 'd') return i + 2;
            if (!SMGX_AT_END(i + 2) && s[i + 2] < 0x80) {
                const uint32_t c2 = s[i + 2] | 0x20;
                if ((c1 == 'r' && c2 == 'e') || (c1 == 'v' && c2 == 'e') || (c1 == 'l' && c2 == 'l')) return i + 3;
            }
        } else if (i + 2 < n && s[i + 1] == 0xC5 && s[i + 2] == 0xBF) {
            return i + 3;  // 'ſ
        }
    }
    // 2. [^\r\n\p{L}\p{N}]?\p{L}+   — the optional char is taken only when a letter follows it (otherwise the
    //    alternative can only match with the optional empty, which needs a letter at i)
    {
        uint32_t j = i;
        bool ok = cls0 == CH_LETTER;
        if (!ok && cls0 != CH_NUMBER && !is_newline(cp0)) {
            uint32_t k = i + len0;
            if (!SMGX_AT_END(k)) {
                uint32_t l1;
                uint32_t cp1 = utf8_decode(s, k, n, l1);
                if (char_class(cp1, u) == CH_LETTER) { j = k; ok = true; }
            }
        }
        if (ok) {
            while (!SMGX_AT_END(j)) {
                uint32_t l;
                uint32_t cp = utf8_decode(s, j, n, l);
                if (char_class(cp, u) != CH_LETTER) break;
                j += l;
            }
            return j;
        }
    }
    // 3. \p{N}{1,3}
    if (cls0 == CH_NUMBER) {
        uint32_t j = i + len0;
        for (int c = 1; c < 3 && !SMGX_AT_END(j); ++c) {
            uint32_t l;
            uint32_t cp = utf8_decode(s, j, n, l);
            if (char_class(cp, u) != CH_NUMBER) break;
            j += l;
        }
        return j;
    }
    // 4.  ?[^\s\p{L}\p{N}]+[\r\n]*
    {
        uint32_t j = i;
        bool ok = cls0 == CH_OTHER;
        if (!ok && cp0 == ' ' && !SMGX_AT_END(i + 1)) {
            uint32_t l1;
            uint32_t cp1 = utf8_decode(s, i + 1, n, l1);
            if (char_class(cp1, u) == CH_OTHER) { j = i + 1; ok = true; }
        }
        if (ok) {
            while (!SMGX_AT_END(j)) {
                uint32_t l;
                uint32_t cp = utf8_decode(s, j, n, l);
                if (char_class(cp, u) != CH_OTHER) break;
                j += l;
            }
            while (!SMGX_AT_END(j) && (s[j] == '\n' || s[j] == '\r')) ++j;
            return j;
        }
    }
    // 5-7. whitespace run starting at i (cls0 == CH_SPACE)
    uint32_t j = i, last_nl_end = 0, last_start = i, count = 0;
    while (!SMGX_AT_END(j)) {
        uint32_t l;
        uint32_t cp = utf8_decode(s, j, n, l);
        if (char_class(cp, u) != CH_SPACE) break;
        last_start = j;
        j += l;
        ++count;
        if (is_newline(cp)) last_nl_end = j;
    }
    if (last_nl_end) return last_nl_end;   // 5. \s*[\r\n]+  : greedy \s* backtracks to the last newline of the run
    if (SMGX_AT_END(j)) return j;          // 6. \s+(?!\S)   : run reaches the end of the (slice of) text
    if (count >= 2) return last_start;     // 6.             : give back one char so that whitespace follows
    return j;                              // 7. \s+
#undef SMGX_AT_END
}

// ---- vocabulary tables ---------------------------------------------------------------------------------------
struct PairSlot { uint64_t key; uint32_t rank; uint32_t id; };          // key = left_id << 32 | right_id ; empty = ~0
struct PieceSlot { uint64_t hash; uint32_t id; uint32_t len; uint32_t off; uint32_t pad; };  // hash 0 = empty
struct SpecialTok { uint32_t off, len, id, pad; };

struct BpeView {
    const uint32_t* byte_token;   // [256] id of each single-byte token
    const PairSlot* pairs;
    uint32_t pair_mask;
    const PieceSlot* pieces;
    uint32_t piece_mask;
    const uint8_t* blob;          // token bytes (piece verification) + special strings
    uint32_t max_token_len;
    const SpecialTok* specials;
    uint32_t n_special;
    uint32_t special_first[8];    // bitmap of first bytes of special strings
    uint32_t whole_piece;         // 1: a piece that is a vocabulary entry becomes that token without merging (tiktoken; HF ignore_merges)
    UnicodeView uni;
};

constexpr uint64_t kPairEmpty = ~0ULL;
constexpr uint32_t kRankMax = 0xFFFFFFFFu;

SMGX_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDULL; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ULL; x ^= x >> 33;
    return x;
}
SMGX_HD uint64_t bytes_hash(const uint8_t* p, uint32_t n) {
    uint64_t h = 0xCBF29CE484222325ULL;
    for (uint32_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001B3ULL; }
    h = mix64(h ^ n);
    return h ? h : 1;
}
// rank + merged id of the adjacent pair (l, r); rank kRankMax when bytes(l)+bytes(r) is not a vocab entry
SMGX_HD uint64_t pair_lookup(const BpeView& v, uint32_t l, uint32_t r) {
    const uint64_t key = ((uint64_t)l << 32) | r;
    uint32_t h = (uint32_t)(mix64(key) >> 32) & v.pair_mask;
    for (;;) {
        const PairSlot s = v.pairs[h];
        if (s.key == key) return ((uint64_t)s.rank << 32) | s.id;
        if (s.key == kPairEmpty) return ((uint64_t)kRankMax << 32);
        h = (h + 1) & v.pair_mask;
    }
}
// `self.encoder.get(piece)` — whole piece is one vocab entry?
SMGX_HD bool piece_lookup(const BpeView& v, const uint8_t* p, uint32_t n, uint32_t& id) {
    if (n > v.max_token_len) return false;
    const uint64_t hsh = bytes_hash(p, n);
    uint32_t h = (uint32_t)(hsh >> 32) & v.piece_mask;
    for (;;) {
        const PieceSlot s = v.pieces[h];
        if (s.hash == 0) return false;
        if (s.hash == hsh && s.len == n) {
            bool eq = true;
            for (uint32_t i = 0; i < n; ++i) if (v.blob[s.off + i] != p[i]) { eq = false; break; }
            if (eq) { id = s.id; return true; }
        }
        h = (h + 1) & v.piece_mask;
    }
}

// tiktoken `_byte_pair_merge` over token ids: ids[0..n) start as the single-byte tokens of the piece, rk[i] packs
// (rank, merged id) of the pair (ids[i], ids[i+1]).  Returns the number of tokens left in ids[].
SMGX_HD uint32_t byte_pair_merge(const BpeView& v, const uint8_t* p, uint32_t n, uint32_t* ids, uint64_t* rk) {
    for (uint32_t i = 0; i < n; ++i) ids[i] = v.byte_token[p[i]];
    if (n < 2) return n;
    for (uint32_t i = 0; i + 1 < n; ++i) rk[i] = pair_lookup(v, ids[i], ids[i + 1]);
    uint32_t m = n;
    for (;;) {
        uint32_t best = kRankMax, bi = 0;
        for (uint32_t i = 0; i + 1 < m; ++i) {       // leftmost minimum (strict <)
            uint32_t r = (uint32_t)(rk[i] >> 32);
            if (r < best) { best = r; bi = i; }
        }
        if (best == kRankMax) break;
        ids[bi] = (uint32_t)rk[bi];
        for (uint32_t i = bi + 1; i + 1 < m; ++i) { ids[i] = ids[i + 1]; rk[i] = rk[i + 1]; }
        --m;
        rk[bi] = (bi + 1 < m) ? pair_lookup(v, ids[bi], ids[bi + 1]) : ((uint64_t)kRankMax << 32);
        if (bi > 0) rk[bi - 1] = pair_lookup(v, ids[bi - 1], ids[bi]);
    }
    return m;
}

// Longest special token starting at s[i], if any (special strings are matched before the regex, tiktoken `encode`).
SMGX_HD bool special_at(const BpeView& v, const uint8_t* s, uint32_t i, uint32_t n, uint32_t& len, uint32_t& id) {
    const uint8_t c = s[i];
    if (!((v.special_first[c >> 5] >> (c & 31)) & 1)) return false;
    bool found = false;
    for (uint32_t k = 0; k < v.n_special; ++k) {
        const SpecialTok t = v.specials[k];
        if (t.len > n - i || (found && t.len <= len)) continue;
        bool eq = true;
        for (uint32_t b = 0; b < t.len; ++b) if (v.blob[t.off + b] != s[i + b]) { eq = false; break; }
        if (eq) { found = true; len = t.len; id = t.id; }
    }
    return found;
}

}  // namespace smgx
