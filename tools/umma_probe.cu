// Micro-benchmark: cycles per tcgen05.mma (bf16, SS operands, SWIZZLE_NONE K-major, dummy zero data) for
//   cta_group::1  M=128  N in {16, 64, 256}   with 1 or 2 resident CTAs per SM
//   cta_group::2  M=256  N in {16, 64, 256}   (CTA pairs)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

constexpr int ITERS = 4096;

template <int CG, int N>
__global__ void __launch_bounds__(128) probe(long long* out, int chains, int issuers) {
    constexpr int COLS = N <= 16 ? 128 : (N <= 64 ? 256 : 256);     // TMEM columns: up to 4 resident CTAs per SM must fit
    extern __shared__ __align__(1024) uint8_t smem[];       // A: 128 rows x 16 k (4 KB) ; B: up to 256 rows x 16 k (8 KB)
    __shared__ __align__(8) uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t& bar = bars[warp];
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    long long t0 = 0, t1 = 0;
    if (lane == 0 && warp < issuers && rank == 0) {
        const uint32_t a = smem_u32(smem), b = a + 4096;
        const uint64_t ad = smem_desc(a, 2048, 128);                         // K chunks 2048 B apart, 8-row groups 128 B apart
        const uint64_t bd = smem_desc(b, 4096, 128);
        const uint32_t idesc = idesc_bf16(CG == 2 ? 256 : 128, N);
        t0 = clock64();
        for (int i = 0; i < ITERS; ++i) {
            const uint32_t d = tmem + (uint32_t)((warp * chains + i % chains) * N) % (uint32_t)COLS;   // `chains` independent accumulators
            if (CG == 1)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        }
        if (CG == 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    }
    if (lane == 0 && warp < issuers) {
        uint32_t spins = 0;
        while (!mbar_try_wait(smem_u32(&bar), 0)) {
            if (++spins > (1u << 24)) __trap();
        }
        t1 = clock64();
        if (rank == 0) out[blockIdx.x * 4 + warp] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(COLS) : "memory");
    }
}

template <int CG, int N>
void run(const char* name, int grid, int chains, size_t smem, int issuers = 1) {
    long long* d;
    cudaMalloc(&d, sizeof(long long) * grid * 4);
    cudaMemset(d, 0, sizeof(long long) * grid * 4);
    cudaFuncSetAttribute(probe<CG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CG, N>, d, chains, issuers);
        if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: run failed: %s\n", name, cudaGetErrorString(e)); return; }
    }
    std::vector<long long> h(grid * 4);
    cudaMemcpy(h.data(), d, sizeof(long long) * grid * 4, cudaMemcpyDeviceToHost);
    std::vector<long long> v;
    for (auto x : h) if (x > 0) v.push_back(x);
    std::sort(v.begin(), v.end());
    if (v.empty()) { printf("%s: no samples\n", name); return; }
    printf("%-58s cycles/MMA median %.1f  (min %.1f max %.1f, %zu issuers)\n", name, (double)v[v.size() / 2] / ITERS, (double)v.front() / ITERS,
           (double)v.back() / ITERS, v.size());
    cudaFree(d);
}

int main() {
    const size_t small = 16384, big = 120 * 1024;      // big: forces 1 CTA per SM
    run<1, 16>("cta_group::1 M=128 N=16  1 CTA/SM, 1 accumulator", 148, 1, big);
    run<1, 16>("cta_group::1 M=128 N=16  1 CTA/SM, 8 accumulators", 148, 8, big);
    run<1, 16>("cta_group::1 M=128 N=16  2 CTA/SM", 296, 1, small);
    run<1, 16>("cta_group::1 M=128 N=16  4 CTA/SM", 592, 1, small);
    run<1, 64>("cta_group::1 M=128 N=64  1 CTA/SM", 148, 1, big);
    run<1, 64>("cta_group::1 M=128 N=64  2 CTA/SM", 296, 1, small);
    run<1, 256>("cta_group::1 M=128 N=256 1 CTA/SM", 148, 1, big);
    run<1, 16>("cta_group::1 M=128 N=16  1 CTA/SM, 2 issuing warps", 148, 1, big, 2);
    run<1, 16>("cta_group::1 M=128 N=16  1 CTA/SM, 4 issuing warps", 148, 1, big, 4);
    run<1, 64>("cta_group::1 M=128 N=64  1 CTA/SM, 4 issuing warps", 148, 1, big, 4);
    run<1, 64>("cta_group::1 M=128 N=64  2 CTA/SM, 2 issuing warps", 296, 1, small, 2);
    run<2, 16>("cta_group::2 M=256 N=16  (74 pairs, 1 CTA/SM)", 148, 1, big);
    run<2, 32>("cta_group::2 M=256 N=32  (74 pairs, 1 CTA/SM)", 148, 1, big);
    run<2, 64>("cta_group::2 M=256 N=64  (74 pairs, 1 CTA/SM)", 148, 1, big);
    run<2, 128>("cta_group::2 M=256 N=128 (74 pairs, 1 CTA/SM)", 148, 1, big);
    run<2, 256>("cta_group::2 M=256 N=256 (74 pairs, 1 CTA/SM)", 148, 1, big);
    run<2, 16>("cta_group::2 M=256 N=16  (148 pairs, 2 CTA/SM)", 296, 1, small);
    run<2, 64>("cta_group::2 M=256 N=64  (148 pairs, 2 CTA/SM)", 296, 1, small);
    return 0;
}
