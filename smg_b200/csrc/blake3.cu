// a16 — mesh path hashes on the device: smg_mesh::hash_token_path / hash_node_path (crates/mesh/src/hash.rs:22-52).
// BLAKE3 (unkeyed) of each request's bytes — the little-endian u32 ids of a token path, or the UTF-8 bytes of a text path —
// truncated to the low 8 bytes (LE) with 0 remapped to 1.  The algorithm is the third-party crate blake3 = "1.5"
// (crates/mesh/Cargo.toml); implemented here from the published specification.
//   kernel 1: one thread per 1 KiB chunk → chunk chaining value (16 dependent compressions; a single-chunk input is finished
//             here with the ROOT flag)
//   kernel 2: one thread per request → binary tree of parent nodes over the chunk values (left subtree = largest power of two)
// Requests are already in HBM (the select path uploaded them), 2 KiB each: the work is ALU-bound, ≈33 compressions per request.
#include "blake3.h"

namespace smgx {
namespace {

__constant__ uint32_t kIV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
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

// block of `len` ≤ 64 bytes at p, zero padded, as 16 little-endian words
__device__ __forceinline__ void load_block(const uint8_t* __restrict__ p, uint32_t len, uint32_t m[16]) {
    if (len == 64 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + q);
            m[4 * q] = v.x; m[4 * q + 1] = v.y; m[4 * q + 2] = v.z; m[4 * q + 3] = v.w;
        }
        return;
    }
    if (len == 64 && (reinterpret_cast<uintptr_t>(p) & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 16; ++q) m[q] = __ldg(reinterpret_cast<const uint32_t*>(p) + q);
        return;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) if ((uint32_t)(4 * q + b) < len) w |= (uint32_t)p[4 * q + b] << (8 * b);
        m[q] = w;
    }
}

__device__ __forceinline__ uint64_t finish_hash(const uint32_t cv[8]) {
    const uint64_t h = ((uint64_t)cv[1] << 32) | cv[0];   // low 8 digest bytes, little-endian
    return h == 0 ? 1 : h;                                // GLOBAL_EVICTION_HASH carve-out (hash.rs:26-30)
}

// chunk_start[r] = index of request r's first chunk in the CV scratch (prefix sum of max(1, ceil(bytes / 1024)))
__global__ void __launch_bounds__(128) blake3_chunks_kernel(const uint8_t* __restrict__ data, const uint32_t* __restrict__ offsets, uint32_t elem_bytes,
                                                            const uint32_t* __restrict__ chunk_start, uint32_t n, uint32_t total_chunks,
                                                            uint32_t* __restrict__ cvs, uint64_t* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_chunks) return;
    uint32_t lo = 0, hi = n;   // last r with chunk_start[r] <= t
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (chunk_start[mid] <= t) lo = mid; else hi = mid; }
    const uint32_t r = lo, c = t - chunk_start[r], n_chunks = chunk_start[r + 1] - chunk_start[r];
    const uint64_t beg = (uint64_t)offsets[r] * elem_bytes, bytes = (uint64_t)(offsets[r + 1] - offsets[r]) * elem_bytes;
    const uint8_t* cp = data + beg + (uint64_t)c * 1024;
    const uint32_t clen = (uint32_t)min((uint64_t)1024, bytes - (uint64_t)c * 1024);
    const uint32_t n_blocks = clen == 0 ? 1 : (clen + 63) / 64;
    uint32_t cv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cv[i] = kIV[i];
    for (uint32_t b = 0; b < n_blocks; ++b) {
        const uint32_t blen = clen == 0 ? 0 : (b + 1 == n_blocks ? clen - b * 64 : 64);
        uint32_t m[16];
        load_block(cp + b * 64, blen, m);
        uint32_t flags = (b == 0 ? CHUNK_START : 0u) | (b + 1 == n_blocks ? CHUNK_END : 0u);
        if (b + 1 == n_blocks && n_chunks == 1) flags |= ROOT;
        compress(cv, m, c, blen, flags);
    }
    if (n_chunks == 1) { out[r] = finish_hash(cv); return; }
#pragma unroll
    for (int i = 0; i < 8; ++i) cvs[(size_t)t * 8 + i] = cv[i];
}

__global__ void __launch_bounds__(128) blake3_parents_kernel(const uint32_t* __restrict__ chunk_start, uint32_t n, const uint32_t* __restrict__ cvs,
                                                             uint64_t* __restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t first = chunk_start[r], k = chunk_start[r + 1] - first;
    if (k <= 1) return;
    uint32_t stack[32][8];   // chaining values of completed subtrees (k < 2^32 chunks)
    int sp = 0;
    uint32_t cur[8], m[16];
    for (uint32_t c = 0; c + 1 < k; ++c) {   // add_chunk_chaining_value for every chunk but the last
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = cvs[(size_t)(first + c) * 8 + i];
        uint32_t total = c + 1;
        while ((total & 1) == 0) {
            --sp;
#pragma unroll
            for (int i = 0; i < 8; ++i) { m[i] = stack[sp][i]; m[8 + i] = cur[i]; }
#pragma unroll
            for (int i = 0; i < 8; ++i) cur[i] = kIV[i];
            compress(cur, m, 0, 64, PARENT);
            total >>= 1;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) stack[sp][i] = cur[i];
        ++sp;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = cvs[(size_t)(first + k - 1) * 8 + i];
    while (sp > 0) {   // fold the right edge of the tree; the last parent is the root
        --sp;
#pragma unroll
        for (int i = 0; i < 8; ++i) { m[i] = stack[sp][i]; m[8 + i] = cur[i]; }
#pragma unroll
        for (int i = 0; i < 8; ++i) cur[i] = kIV[i];
        compress(cur, m, 0, 64, PARENT | (sp == 0 ? ROOT : 0u));
    }
    out[r] = finish_hash(cur);
}
}  // namespace

void launch_blake3_paths(const uint8_t* d_data, const uint32_t* d_offsets, uint32_t elem_bytes, const uint32_t* d_chunk_start, uint32_t n,
                         uint32_t total_chunks, uint32_t* d_cv_scratch, uint64_t* d_out, cudaStream_t stream, uint64_t* launches) {
    if (n == 0) return;
    blake3_chunks_kernel<<<(total_chunks + 127) / 128, 128, 0, stream>>>(d_data, d_offsets, elem_bytes, d_chunk_start, n, total_chunks, d_cv_scratch, d_out);
    ++*launches;
    if (total_chunks > n) {   // at least one request spans several chunks
        blake3_parents_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_chunk_start, n, d_cv_scratch, d_out);
        ++*launches;
    }
    SMGX_CUDA(cudaGetLastError());
}

}  // namespace smgx
