// Shared helpers for the sm_100a kernels of libavmnist_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/avmnist_b200.h"

namespace b200 {

void set_error(const char* fmt, ...);

// Launch-status helper: returns 0 or the cudaError_t (positive) after recording the message.
static inline int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define B200_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            b200::set_error(__VA_ARGS__);  \
            return (code);                 \
        }                                  \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- fixed-order reductions of one float per warp / per block (run-to-run deterministic: no float atomics) ----
// Sum of the warps' values (v: the warp's value, valid in lane 0); result valid in thread 0.  All threads of the block must call.
__device__ __forceinline__ float block_sum_ordered(float v) {
    __shared__ float s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    float a = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < nw; ++w) a += s_warp[w];
    __syncthreads();
    return a;
}
// Grid-wide sum of one float per block (valid in thread 0) in BLOCK ORDER: every block stores its partial and draws a ticket; the
// block that draws the last one adds the partials by index and writes out[0].  work[0] = ticket (unsigned, 0 at launch, left 0),
// work[1 + b] = partial of block b: work needs 1 + gridDim.x floats.
__device__ __forceinline__ void grid_sum_ordered(float block_partial, float* work, float* out) {
    if (threadIdx.x == 0) {
        volatile float* parts = work + 1;
        parts[blockIdx.x] = block_partial;
        __threadfence();
        const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(work), 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            float a = 0.f;
            for (unsigned b = 0; b < gridDim.x; ++b) a += parts[b];
            out[0] = a;
            *reinterpret_cast<volatile unsigned*>(work) = 0u;
        }
    }
}

// ---- Philox4x32-10 (counter-based RNG; same construction as curand/torch, used with our own key/counter map) ----
struct Philox {
    uint32_t key[2];
    __host__ __device__ Philox(uint64_t seed) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
    }
    __device__ __forceinline__ uint4 operator()(uint64_t counter, uint64_t stream) const {
        uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32);
        uint32_t c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
// uniform in [0,1) with 24 random bits (torch's CPU uniform_ for float uses the same 24-bit mantissa construction)
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// uniform in [0,1) with 53 bits from two words
__device__ __forceinline__ double u01d(uint32_t a, uint32_t b) {
    return (double)((((uint64_t)a << 32) | b) >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace b200
