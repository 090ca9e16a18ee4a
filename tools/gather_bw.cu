// Micro-benchmark: achievable HBM read bandwidth for RANDOM reads of G contiguous bytes out of a large buffer, as a function
// of G.  It bounds the radix-tree walk (labels of 512 B – 1 KB at random places in a ≈1 GB arena, 16 B slots, 16 B headers):
// the streaming-copy peak in MEASURED_PEAKS.json is not reachable by a pointer-chasing gather.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_bw tools/gather_bw.cu ; ./tools/gather_bw [buffer_MiB]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// one warp per read of G bytes (G >= 512: 16 B per lane per step), or G/16 lanes per read for smaller G
__global__ void gather_kernel(const uint4* __restrict__ buf, const uint32_t* __restrict__ idx, uint32_t n_reads, uint32_t quads_per_read,
                              uint32_t lanes_per_read, uint4* __restrict__ sink) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = t / lanes_per_read, rank = t % lanes_per_read;
    const uint32_t n_groups = gridDim.x * blockDim.x / lanes_per_read;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (uint32_t r = group; r < n_reads; r += n_groups) {
        const uint4* p = buf + (size_t)idx[r] * quads_per_read;
        for (uint32_t q = rank; q < quads_per_read; q += lanes_per_read) {
            const uint4 v = __ldg(p + q);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[t] = acc;
}

int main(int argc, char** argv) {
    const size_t mib = argc > 1 ? (size_t)atoll(argv[1]) : 1024;
    const size_t bytes = mib << 20;
    uint4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
    uint4* sink; CK(cudaMalloc(&sink, 1 << 24));
    const uint32_t n_reads_max = 1u << 22;
    uint32_t* d_idx; CK(cudaMalloc(&d_idx, n_reads_max * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("{\"buffer_mib\": %zu, \"results\": [", mib);
    bool firstp = true;
    for (uint32_t G : {32u, 64u, 128u, 256u, 512u, 1024u, 2048u, 4096u, 16384u}) {
        const size_t slots = bytes / G;
        const uint32_t n_reads = (uint32_t)std::min<size_t>(n_reads_max, (size_t)(256u << 20) / G * 4);   // ≥ 1 GiB of traffic, capped
        std::vector<uint32_t> idx(n_reads);
        uint64_t x = 88172645463325252ULL;
        for (auto& v : idx) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (uint32_t)(x % slots); }
        CK(cudaMemcpy(d_idx, idx.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice));
        const uint32_t quads = G / 16 ? G / 16 : 1, lanes = quads < 32 ? quads : 32;
        const int grid = 148 * 8;
        for (int w = 0; w < 2; ++w) gather_kernel<<<grid, 256>>>(buf, d_idx, n_reads, quads, lanes, sink);
        CK(cudaEventRecord(e0));
        const int reps = 5;
        for (int w = 0; w < reps; ++w) gather_kernel<<<grid, 256>>>(buf, d_idx, n_reads, quads, lanes, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double gbs = (double)n_reads * (G < 16 ? 16 : G) * reps / (ms * 1e-3) / 1e9;
        printf("%s{\"read_bytes\": %u, \"gbs\": %.1f}", firstp ? "" : ", ", G, gbs);
        firstp = false;
    }
    printf("]}\n");
    return 0;
}
