// power_of_two (policies/power_of_two.rs): "randomly selects two workers and routes to the one with lower load".
//
// The reference draws from a thread-local ChaCha generator nobody can seed (rand::rng(), :48), so there is no stream to be bit-equal
// to; what is reproduced is the procedure — idx1 uniform in 0..h, idx2 = (idx1 + 1 + uniform(0..h-1)) % h (:49-52), the metric rule
// (token usage only when BOTH candidates have a cached load response, otherwise request counts for both, :66-88) and the tie rule
// (load1 <= load2 keeps the first candidate, :91-95) — over a counter-based stream: draw k of a call is mix(seed + C·(k + 1)) with the
// splitmix64 finaliser, request i of the batch uses draws 2i and 2i + 1, and a draw x becomes an index as ⌊x·h / 2^64⌋.  The oracle
// (oracle/power_of_two.h) states the same stream independently; the GPU tests compare picks bit for bit.
#include "power_of_two.h"

#include "common.h"

namespace smgx {
namespace {

__device__ __forceinline__ uint64_t p2c_mix(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__device__ __forceinline__ uint64_t p2c_draw(uint64_t seed, uint64_t k) { return p2c_mix(seed + 0x9E3779B97F4A7C15ULL * (k + 1)); }

__global__ void __launch_bounds__(256) power_of_two_kernel(const P2cArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    int32_t out = -1, c1 = -1, c2 = -1;
    uint8_t metric = 2;
    const uint32_t h = a.n_healthy;
    if (h == 1) out = a.healthy[0];                                                    // :44-46
    else if (h >= 2) {
        const uint32_t i1 = (uint32_t)__umul64hi(p2c_draw(a.seed, 2ull * i), (uint64_t)h);
        const uint32_t i2 = (i1 + 1 + (uint32_t)__umul64hi(p2c_draw(a.seed, 2ull * i + 1), (uint64_t)(h - 1))) % h;   // never i1
        c1 = a.healthy[i1]; c2 = a.healthy[i2];
        const double u1 = a.usage[c1], u2 = a.usage[c2];
        const bool both = u1 == u1 && u2 == u2;                                         // neither is NaN: both workers have a cached load response
        const double l1 = both ? u1 : (double)a.loads[c1], l2 = both ? u2 : (double)a.loads[c2];
        metric = both ? 1 : 0;
        out = l1 <= l2 ? c1 : c2;                                                        // :91-95
    }
    a.out_idx[i] = out;
    if (a.out_pair) { a.out_pair[2 * i] = c1; a.out_pair[2 * i + 1] = c2; }
    if (a.out_metric) a.out_metric[i] = metric;
}

}  // namespace

void launch_power_of_two(const P2cArgs& a, cudaStream_t stream) {
    if (!a.n) return;
    power_of_two_kernel<<<(a.n + 255) / 256, 256, 0, stream>>>(a);
    SMGX_CUDA(cudaGetLastError());
}

}  // namespace smgx
