// XXH3-64 (seeded) for the KV-event index, host + device, specialised for the only inputs the path ever
// hashes: little-endian u32 token words (length a multiple of 4 bytes) and 16-byte (prev ‖ cur) pairs.
//
// Replaces, bit-exactly, what the reference computes through the `xxhash-rust` crate at
//   crates/kv_index/src/event_tree.rs:122-129  compute_content_hash   (Xxh3::with_seed(1337), streaming)
//   crates/kv_index/src/event_tree.rs:477-482  compute_next_seq_hash  (xxh3_64_with_seed(16 B, 1337))
// Written from the public XXH3 specification; every length class reachable with 4-byte words is covered
// (0, 4, 8, 12-16, 20-128, 132-240, >240 bytes).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SMGX_HD __host__ __device__ __forceinline__
#else
#define SMGX_HD inline
#endif

namespace smgx {

constexpr uint64_t kSeed = 1337;  // event_tree.rs:35

// The 192-byte default secret as 48 little-endian u32 words (secret reads are at 4-byte-aligned or odd offsets;
// odd offsets (mid-size tail, long-input merge) go through sec64_unaligned()).
#define SMGX_XXH3_SECRET_WORDS                                                                                   \
    0x396cfeb8u, 0xbe4ba423u, 0x2c81017cu, 0x1cad21f7u, 0xe96dd4deu, 0xdb979083u, 0xa4a44072u, 0x1f67b3b7u,     \
    0x4ee679cbu, 0x78e5c0ccu, 0x7dd05a82u, 0x2172ffccu, 0x744608b8u, 0x8e2443f7u, 0xe69035e0u, 0x4c263a81u,     \
    0xbb52283cu, 0xcb00c391u, 0x8b65d088u, 0xa32e531bu, 0x97486471u, 0x4ef90da2u, 0x46ef1938u, 0xd8acdea9u,     \
    0x3f76faa8u, 0x3f349ce3u, 0xc7bbdcf9u, 0x1d4f0bc7u, 0x4be0518au, 0x3159b4cdu, 0xc97e9fc8u, 0x647378d9u,     \
    0x83acc5eau, 0xc3ebd334u, 0xffa081c5u, 0xeb6313fau, 0x51dd0d17u, 0x49daf0b7u, 0x265516d3u, 0x9e68d429u,     \
    0x58be162bu, 0xfca1477du, 0xd1b8f88fu, 0xce31d07au, 0x8f3acb45u, 0x28041695u, 0xcafbd7afu, 0x7e404bbbu

// compile-time copy: constant offsets fold to immediates on both host and device
constexpr uint32_t kSecretC[48] = {SMGX_XXH3_SECRET_WORDS};
static const uint32_t kSecretHost[48] = {SMGX_XXH3_SECRET_WORDS};
#if defined(__CUDACC__)
// run-time offsets (generic lengths) read the secret from constant memory on the device
static __device__ __constant__ uint32_t kSecretDev[48] = {SMGX_XXH3_SECRET_WORDS};
#endif
SMGX_HD uint32_t secw(int i) {
#if defined(__CUDA_ARCH__)
    return kSecretDev[i];
#else
    return kSecretHost[i];
#endif
}
template <int OFF> SMGX_HD constexpr uint64_t sec64c() {
    return (uint64_t)kSecretC[OFF >> 2] | ((uint64_t)kSecretC[(OFF >> 2) + 1] << 32);
}

constexpr uint64_t P32_1 = 0x9E3779B1ULL, P32_2 = 0x85EBCA77ULL, P32_3 = 0xC2B2AE3DULL;
constexpr uint64_t P64_1 = 0x9E3779B185EBCA87ULL, P64_2 = 0xC2B2AE3D27D4EB4FULL, P64_3 = 0x165667B19E3779F9ULL,
                   P64_4 = 0x85EBCA77C2B2AE63ULL, P64_5 = 0x27D4EB2F165667C5ULL;
constexpr uint64_t PMX1 = 0x165667919E3779F9ULL, PMX2 = 0x9FB21C651E98DF25ULL;

SMGX_HD uint64_t mk64(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }
// secret bytes [off, off+8) for off % 4 == 0
SMGX_HD uint64_t sec64(int off) { return mk64(secw(off >> 2), secw((off >> 2) + 1)); }
// secret bytes [off, off+8) for any off
SMGX_HD uint64_t sec64_unaligned(int off) {
    int w = off >> 2, sh = (off & 3) * 8;
    uint64_t lo = mk64(secw(w), secw(w + 1));
    if (sh == 0) return lo;
    uint64_t hi = secw(w + 2);
    return (lo >> sh) | (hi << (64 - sh));
}
SMGX_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
SMGX_HD uint64_t bswap64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return mk64(__byte_perm(hi, 0, 0x0123), __byte_perm(lo, 0, 0x0123));
#else
    return __builtin_bswap64(x);
#endif
}
SMGX_HD uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}
SMGX_HD uint64_t mul128_fold64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return (a * b) ^ __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    return (uint64_t)p ^ (uint64_t)(p >> 64);
#endif
}
SMGX_HD uint64_t avalanche(uint64_t h) {
    h ^= h >> 37;
    h *= PMX1;
    return h ^ (h >> 32);
}
SMGX_HD uint64_t xxh64_avalanche(uint64_t h) {
    h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32;
    return h;
}
SMGX_HD uint64_t rrmxmx(uint64_t h, uint64_t len) {
    h ^= rotl64(h, 49) ^ rotl64(h, 24);
    h *= PMX2;
    h ^= (h >> 35) + len;
    h *= PMX2;
    return h ^ (h >> 28);
}
// one 16-byte lane mix: input words (w0..w3), secret bytes [soff, soff+16) with soff % 4 == 0
SMGX_HD uint64_t mix16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int soff, uint64_t seed) {
    return mul128_fold64(mk64(w0, w1) ^ (sec64(soff) + seed), mk64(w2, w3) ^ (sec64(soff + 8) - seed));
}
SMGX_HD uint64_t mix16_unaligned(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int soff, uint64_t seed) {
    return mul128_fold64(mk64(w0, w1) ^ (sec64_unaligned(soff) + seed), mk64(w2, w3) ^ (sec64_unaligned(soff + 8) - seed));
}

// ---- the two fixed-size hot cases -------------------------------------------------------------------------

// 64-byte block (block_size 16): spec path 17..128 bytes with len = 64.
template <int SOFF> SMGX_HD uint64_t mix16c(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint64_t seed) {
    constexpr uint64_t s0 = sec64c<SOFF>(), s1 = sec64c<SOFF + 8>();
    return mul128_fold64(mk64(w0, w1) ^ (s0 + seed), mk64(w2, w3) ^ (s1 - seed));
}
SMGX_HD uint64_t xxh3_16words(const uint32_t (&w)[16], uint64_t seed) {
    uint64_t acc = 64ULL * P64_1;
    // len > 32: mix(in+16, sec+32) + mix(in+len-32, sec+48)
    acc += mix16c<32>(w[4], w[5], w[6], w[7], seed);
    acc += mix16c<48>(w[8], w[9], w[10], w[11], seed);
    // always: mix(in, sec) + mix(in+len-16, sec+16)
    acc += mix16c<0>(w[0], w[1], w[2], w[3], seed);
    acc += mix16c<16>(w[12], w[13], w[14], w[15], seed);
    return avalanche(acc);
}

// 16-byte (prev ‖ cur) pair: spec path 9..16 bytes with len = 16.
SMGX_HD uint64_t xxh3_pair(uint64_t prev, uint64_t cur, uint64_t seed) {
    constexpr uint64_t k1 = sec64c<24>() ^ sec64c<32>(), k2 = sec64c<40>() ^ sec64c<48>();
    uint64_t bf1 = k1 + seed, bf2 = k2 - seed;
    uint64_t lo = prev ^ bf1, hi = cur ^ bf2;
    uint64_t acc = 16ULL + bswap64(lo) + hi + mul128_fold64(lo, hi);
    return avalanche(acc);
}

// ---- generic: n little-endian u32 words at `w` (any n ≥ 0) ------------------------------------------------
SMGX_HD uint64_t rd64w(const uint32_t* w, int word) { return mk64(w[word], w[word + 1]); }

SMGX_HD void acc512(uint64_t (&acc)[8], const uint32_t* in /*16 words*/, const uint64_t* cs /*8 qwords*/) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t dv = rd64w(in, 2 * i);
        uint64_t dk = dv ^ cs[i];
        acc[i ^ 1] += dv;
        acc[i] += (dk & 0xFFFFFFFFULL) * (dk >> 32);
    }
}

// n ≤ 60 words (≤ 240 bytes): the short and mid-size forms
SMGX_HD uint64_t xxh3_words_upto60(const uint32_t* w, uint32_t n, uint64_t seed) {
    const uint64_t len = (uint64_t)n * 4;
    if (n == 0) return xxh64_avalanche(seed ^ (sec64(56) ^ sec64(64)));
    if (n <= 2) {  // 4..8 bytes
        uint64_t s = seed ^ ((uint64_t)bswap32((uint32_t)seed) << 32);
        uint32_t i1 = w[0], i2 = w[n - 1];
        uint64_t bitflip = (sec64(8) ^ sec64(16)) - s;
        uint64_t in64 = (uint64_t)i2 + ((uint64_t)i1 << 32);
        return rrmxmx(in64 ^ bitflip, len);
    }
    if (n <= 4) {  // 12, 16 bytes
        uint64_t bf1 = (sec64(24) ^ sec64(32)) + seed;
        uint64_t bf2 = (sec64(40) ^ sec64(48)) - seed;
        uint64_t lo = rd64w(w, 0) ^ bf1, hi = rd64w(w, n - 2) ^ bf2;
        uint64_t acc = len + bswap64(lo) + hi + mul128_fold64(lo, hi);
        return avalanche(acc);
    }
    if (n <= 32) {  // 20..128 bytes
        uint64_t acc = len * P64_1;
        if (n > 8) {
            if (n > 16) {
                if (n > 24) {
                    acc += mix16(w[12], w[13], w[14], w[15], 96, seed);
                    acc += mix16(w[n - 16], w[n - 15], w[n - 14], w[n - 13], 112, seed);
                }
                acc += mix16(w[8], w[9], w[10], w[11], 64, seed);
                acc += mix16(w[n - 12], w[n - 11], w[n - 10], w[n - 9], 80, seed);
            }
            acc += mix16(w[4], w[5], w[6], w[7], 32, seed);
            acc += mix16(w[n - 8], w[n - 7], w[n - 6], w[n - 5], 48, seed);
        }
        acc += mix16(w[0], w[1], w[2], w[3], 0, seed);
        acc += mix16(w[n - 4], w[n - 3], w[n - 2], w[n - 1], 16, seed);
        return avalanche(acc);
    }
    {  // 132..240 bytes
        uint64_t acc = len * P64_1;
        uint32_t rounds = n / 4;
        for (uint32_t i = 0; i < 8; ++i) acc += mix16(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3], 16 * i, seed);
        acc = avalanche(acc);
        for (uint32_t i = 8; i < rounds; ++i)
            acc += mix16_unaligned(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3], 16 * (i - 8) + 3, seed);
        acc += mix16_unaligned(w[n - 4], w[n - 3], w[n - 2], w[n - 1], 136 - 17, seed);
        return avalanche(acc);
    }
}

SMGX_HD uint64_t xxh3_words(const uint32_t* w, uint32_t n, uint64_t seed) {
    if (n <= 60) return xxh3_words_upto60(w, n, seed);
    const uint64_t len = (uint64_t)n * 4;
    // > 240 bytes: stripes of 64 B against the seed-derived secret
    uint64_t cs[24];
#pragma unroll
    for (int i = 0; i < 12; ++i) { cs[2 * i] = sec64(16 * i) + seed; cs[2 * i + 1] = sec64(16 * i + 8) - seed; }
    uint64_t acc[8] = {P32_3, P64_1, P64_2, P64_3, P64_4, P32_2, P64_5, P32_1};
    const uint64_t block_len = 1024;  // 16 stripes
    uint64_t nb_blocks = (len - 1) / block_len;
    for (uint64_t b = 0; b < nb_blocks; ++b) {
        for (int s = 0; s < 16; ++s) acc512(acc, w + (b * block_len + s * 64) / 4, cs + s);
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // scramble with secret bytes [128, 192)
            uint64_t a = acc[i];
            a ^= a >> 47;
            a ^= cs[16 + i];
            a *= P32_1;
            acc[i] = a;
        }
    }
    uint64_t nb_stripes = ((len - 1) - block_len * nb_blocks) / 64;
    for (uint64_t s = 0; s < nb_stripes; ++s) acc512(acc, w + (nb_blocks * block_len + s * 64) / 4, cs + s);
    {   // last stripe against secret bytes [121, 185): unaligned view of the derived secret
        uint64_t last[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int off = 192 - 64 - 7 + 8 * i;  // byte offset into cs
            int q = off >> 3, sh = (off & 7) * 8;
            last[i] = (cs[q] >> sh) | (cs[q + 1] << (64 - sh));
        }
        acc512(acc, w + n - 16, last);
    }
    uint64_t r = len * P64_1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int o0 = 11 + 16 * i, o1 = o0 + 8;
        uint64_t k0 = (cs[o0 >> 3] >> ((o0 & 7) * 8)) | (cs[(o0 >> 3) + 1] << (64 - (o0 & 7) * 8));
        uint64_t k1 = (cs[o1 >> 3] >> ((o1 & 7) * 8)) | (cs[(o1 >> 3) + 1] << (64 - (o1 & 7) * 8));
        r += mul128_fold64(acc[2 * i] ^ k0, acc[2 * i + 1] ^ k1);
    }
    return avalanche(r);
}

}  // namespace smgx
