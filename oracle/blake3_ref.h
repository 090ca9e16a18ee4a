/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into / imported by the product path.
 *
 * CPU restatement of BLAKE3 (unkeyed hash mode) as used by the reference's mesh path hashes:
 *   crates/mesh/src/hash.rs :22-30 hash_node_path (blake3 of the UTF-8 bytes, low 8 bytes LE, 0 → 1)
 *                           :40-52 hash_token_path (blake3 of the little-endian u32 bytes of every id, same truncation)
 * The algorithm lives in the third-party crate blake3 = "1.5" (crates/mesh/Cargo.toml; no Cargo.lock in the tree), absent from
 * /root/reference.  Restated from the published BLAKE3 specification (§2: compression function, chunk chaining values, binary
 * tree of parent nodes, ROOT flag).  Pinned by tests/golden/blake3_vectors.json, generated with the Python `blake3` 1.0.8
 * bindings of the official implementation (tests/golden/gen_blake3_golden.py) — including the official empty-input digest.
 */
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {
namespace b3 {

static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const int PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline void g(uint32_t* s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx; s[d] = rotr(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = rotr(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 7);
}
// out[16]: first 8 words are the new chaining value
static inline void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter, uint32_t block_len, uint32_t flags, uint32_t out[16]) {
    uint32_t s[16], m[16], t[16];
    for (int i = 0; i < 8; ++i) s[i] = cv[i];
    for (int i = 0; i < 4; ++i) s[8 + i] = IV[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = block_len; s[15] = flags;
    memcpy(m, block, 64);
    for (int r = 0; r < 7; ++r) {
        g(s, 0, 4, 8, 12, m[0], m[1]);   g(s, 1, 5, 9, 13, m[2], m[3]);
        g(s, 2, 6, 10, 14, m[4], m[5]);  g(s, 3, 7, 11, 15, m[6], m[7]);
        g(s, 0, 5, 10, 15, m[8], m[9]);  g(s, 1, 6, 11, 12, m[10], m[11]);
        g(s, 2, 7, 8, 13, m[12], m[13]); g(s, 3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; ++i) t[i] = m[PERM[i]];
        memcpy(m, t, 64);
    }
    for (int i = 0; i < 8; ++i) { out[i] = s[i] ^ s[i + 8]; out[i + 8] = s[i + 8] ^ cv[i]; }
}

static inline void load_block(const uint8_t* p, size_t len, uint32_t w[16]) {
    uint8_t buf[64] = {0};
    memcpy(buf, p, len);
    for (int i = 0; i < 16; ++i) w[i] = (uint32_t)buf[4 * i] | ((uint32_t)buf[4 * i + 1] << 8) | ((uint32_t)buf[4 * i + 2] << 16) | ((uint32_t)buf[4 * i + 3] << 24);
}

// 32-byte digest of data[0..n)
static inline void hash(const uint8_t* data, size_t n, uint8_t digest[32]) {
    std::vector<std::vector<uint32_t>> stack;   // chaining values of completed subtrees
    const size_t n_chunks = n == 0 ? 1 : (n + 1023) / 1024;
    uint32_t pend_cv[8], pend_block[16], pend_len = 0, pend_flags = 0;
    uint64_t pend_counter = 0;
    for (size_t c = 0; c < n_chunks; ++c) {
        const uint8_t* cp = data + c * 1024;
        const size_t clen = n == 0 ? 0 : (c + 1 == n_chunks ? n - c * 1024 : 1024);
        const size_t n_blocks = clen == 0 ? 1 : (clen + 63) / 64;
        uint32_t cv[8];
        memcpy(cv, IV, 32);
        for (size_t b = 0; b < n_blocks; ++b) {
            const size_t blen = clen == 0 ? 0 : (b + 1 == n_blocks ? clen - b * 64 : 64);
            uint32_t w[16], out[16];
            load_block(cp + b * 64, blen, w);
            uint32_t flags = (b == 0 ? CHUNK_START : 0) | (b + 1 == n_blocks ? CHUNK_END : 0);
            if (b + 1 == n_blocks && c + 1 == n_chunks) {   // the very last block is compressed at finalisation (it may be the root)
                memcpy(pend_cv, cv, 32); memcpy(pend_block, w, 64); pend_len = (uint32_t)blen; pend_flags = flags; pend_counter = c;
                break;
            }
            compress(cv, w, c, (uint32_t)blen, flags, out);
            memcpy(cv, out, 32);
            if (b + 1 == n_blocks) {   // chunk finished and more input follows: add_chunk_chaining_value
                std::vector<uint32_t> nv(cv, cv + 8);
                uint64_t total = c + 1;
                while ((total & 1) == 0) {
                    uint32_t blk[16], o[16];
                    memcpy(blk, stack.back().data(), 32); memcpy(blk + 8, nv.data(), 32);
                    stack.pop_back();
                    compress(IV, blk, 0, 64, PARENT, o);
                    nv.assign(o, o + 8);
                    total >>= 1;
                }
                stack.push_back(nv);
            }
        }
    }
    uint32_t out[16];
    if (stack.empty()) compress(pend_cv, pend_block, pend_counter, pend_len, pend_flags | ROOT, out);
    else {
        compress(pend_cv, pend_block, pend_counter, pend_len, pend_flags, out);
        while (true) {
            uint32_t blk[16];
            memcpy(blk, stack.back().data(), 32); memcpy(blk + 8, out, 32);
            stack.pop_back();
            compress(IV, blk, 0, 64, PARENT | (stack.empty() ? ROOT : 0), out);
            if (stack.empty()) break;
        }
    }
    for (int i = 0; i < 8; ++i) { digest[4 * i] = (uint8_t)out[i]; digest[4 * i + 1] = (uint8_t)(out[i] >> 8); digest[4 * i + 2] = (uint8_t)(out[i] >> 16); digest[4 * i + 3] = (uint8_t)(out[i] >> 24); }
}

}  // namespace b3

// crates/mesh/src/hash.rs:22-30 / :40-52
static inline uint64_t hash_path_bytes(const uint8_t* data, size_t n) {
    uint8_t d[32];
    b3::hash(data, n, d);
    uint64_t h = 0;
    for (int i = 7; i >= 0; --i) h = (h << 8) | d[i];
    return h == 0 ? 1 : h;
}
static inline uint64_t hash_node_path(const std::string& s) { return hash_path_bytes((const uint8_t*)s.data(), s.size()); }
static inline uint64_t hash_token_path(const uint32_t* toks, size_t n) {
    std::vector<uint8_t> b(n * 4);
    for (size_t i = 0; i < n; ++i) { b[4 * i] = (uint8_t)toks[i]; b[4 * i + 1] = (uint8_t)(toks[i] >> 8); b[4 * i + 2] = (uint8_t)(toks[i] >> 16); b[4 * i + 3] = (uint8_t)(toks[i] >> 24); }
    return hash_path_bytes(b.data(), b.size());
}

}  // namespace orc
