// BLAKE3 compression function for device code (shared by blake3.cu and prefix_hash.cu).  Written from the published specification.
#pragma once
#include <cstdint>

namespace smgx {
namespace b3 {

static __device__ __constant__ uint32_t kIV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

#define B3_G(a, b, c, d, mx, my)                        \
    do {                                                \
        a = a + b + (mx); d = rotr(d ^ a, 16);          \
        c = c + d;        b = rotr(b ^ c, 12);          \
        a = a + b + (my); d = rotr(d ^ a, 8);           \
        c = c + d;        b = rotr(b ^ c, 7);           \
    } while (0)

// cv[8] ← first 8 output words of compress(cv, m, counter, block_len, flags)
__device__ __forceinline__ void compress(uint32_t cv[8], const uint32_t m[16], uint64_t counter, uint32_t block_len, uint32_t flags) {
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = kIV[0], s9 = kIV[1], s10 = kIV[2], s11 = kIV[3];
    uint32_t s12 = (uint32_t)counter, s13 = (uint32_t)(counter >> 32), s14 = block_len, s15 = flags;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        // fully unrolled: the schedule indices are compile-time constants, m[] stays in registers
        constexpr uint8_t S[7][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
            {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8},
            {3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1},
            {10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6},
            {12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4},
            {9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7},
            {11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13}};
        B3_G(s0, s4, s8, s12, m[S[r][0]], m[S[r][1]]);
        B3_G(s1, s5, s9, s13, m[S[r][2]], m[S[r][3]]);
        B3_G(s2, s6, s10, s14, m[S[r][4]], m[S[r][5]]);
        B3_G(s3, s7, s11, s15, m[S[r][6]], m[S[r][7]]);
        B3_G(s0, s5, s10, s15, m[S[r][8]], m[S[r][9]]);
        B3_G(s1, s6, s11, s12, m[S[r][10]], m[S[r][11]]);
        B3_G(s2, s7, s8, s13, m[S[r][12]], m[S[r][13]]);
        B3_G(s3, s4, s9, s14, m[S[r][14]], m[S[r][15]]);
    }
    cv[0] = s0 ^ s8; cv[1] = s1 ^ s9; cv[2] = s2 ^ s10; cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12; cv[5] = s5 ^ s13; cv[6] = s6 ^ s14; cv[7] = s7 ^ s15;
}

}  // namespace b3
}  // namespace smgx
