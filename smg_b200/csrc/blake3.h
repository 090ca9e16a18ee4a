// a16 — batched BLAKE3 path hashes (crates/mesh/src/hash.rs:22-52) on the device.  See blake3.cu.
#pragma once
#include <cstdint>

#include "common.h"

namespace smgx {

// Request r = data[offsets[r] * elem_bytes .. offsets[r+1] * elem_bytes) (elem_bytes 4: token ids, 1: UTF-8 text).
// remap_zero: the mesh path hashes reserve 0 (hash.rs:26-30); ring positions (hash_ring.rs:78-86) do not.
// d_chunk_start (n + 1): prefix sum of max(1, ceil(bytes_r / 1024)); d_cv_scratch: total_chunks × 8 u32; d_out: n u64.
void launch_blake3_paths(const uint8_t* d_data, const uint32_t* d_offsets, uint32_t elem_bytes, const uint32_t* d_chunk_start, uint32_t n,
                         uint32_t total_chunks, uint32_t* d_cv_scratch, uint64_t* d_out, cudaStream_t stream, uint64_t* launches, bool remap_zero = true);

}  // namespace smgx
