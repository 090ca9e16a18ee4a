// a16 — mesh path hashes on the device: smg_mesh::hash_token_path / hash_node_path (crates/mesh/src/hash.rs:22-52).
// BLAKE3 (unkeyed) of each request's bytes — the little-endian u32 ids of a token path, or the UTF-8 bytes of a text path —
// truncated to the low 8 bytes (LE) with 0 remapped to 1.  The algorithm is the third-party crate blake3 = "1.5"
// (crates/mesh/Cargo.toml); implemented here from the published specification.
//   kernel 1: one thread per 1 KiB chunk → chunk chaining value (16 dependent compressions; a single-chunk input is finished
//             here with the ROOT flag)
//   kernel 2: one thread per request → binary tree of parent nodes over the chunk values (left subtree = largest power of two)
// Requests are already in HBM (the select path uploaded them), 2 KiB each: the work is ALU-bound, ≈33 compressions per request.
#include "blake3.h"
#include "blake3.cuh"

namespace smgx {
namespace {
using namespace b3;

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

__device__ __forceinline__ uint64_t finish_hash(const uint32_t cv[8], uint32_t remap_zero) {
    const uint64_t h = ((uint64_t)cv[1] << 32) | cv[0];   // low 8 digest bytes, little-endian
    return (h == 0 && remap_zero) ? 1 : h;                // GLOBAL_EVICTION_HASH carve-out (hash.rs:26-30); ring positions keep 0 (hash_ring.rs:78-86)
}

// chunk_start[r] = index of request r's first chunk in the CV scratch (prefix sum of max(1, ceil(bytes / 1024)))
__global__ void __launch_bounds__(128) blake3_chunks_kernel(const uint8_t* __restrict__ data, const uint32_t* __restrict__ offsets, uint32_t elem_bytes,
                                                            const uint32_t* __restrict__ chunk_start, uint32_t n, uint32_t total_chunks,
                                                            uint32_t* __restrict__ cvs, uint64_t* __restrict__ out, uint32_t remap_zero) {
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
    if (n_chunks == 1) { out[r] = finish_hash(cv, remap_zero); return; }
#pragma unroll
    for (int i = 0; i < 8; ++i) cvs[(size_t)t * 8 + i] = cv[i];
}

__global__ void __launch_bounds__(128) blake3_parents_kernel(const uint32_t* __restrict__ chunk_start, uint32_t n, const uint32_t* __restrict__ cvs,
                                                             uint64_t* __restrict__ out, uint32_t remap_zero) {
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
    out[r] = finish_hash(cur, remap_zero);
}
}  // namespace

void launch_blake3_paths(const uint8_t* d_data, const uint32_t* d_offsets, uint32_t elem_bytes, const uint32_t* d_chunk_start, uint32_t n,
                         uint32_t total_chunks, uint32_t* d_cv_scratch, uint64_t* d_out, cudaStream_t stream, uint64_t* launches, bool remap_zero) {
    if (n == 0) return;
    blake3_chunks_kernel<<<(total_chunks + 127) / 128, 128, 0, stream>>>(d_data, d_offsets, elem_bytes, d_chunk_start, n, total_chunks, d_cv_scratch, d_out, remap_zero ? 1u : 0u);
    ++*launches;
    if (total_chunks > n) {   // at least one request spans several chunks
        blake3_parents_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_chunk_start, n, d_cv_scratch, d_out, remap_zero ? 1u : 0u);
        ++*launches;
    }
    SMGX_CUDA(cudaGetLastError());
}

}  // namespace smgx
