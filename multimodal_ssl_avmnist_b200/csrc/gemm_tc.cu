// Linear layers on the tensor cores: tcgen05.mma kind::tf32 (fp32 tensors in HBM are consumed as they are: 10-bit mantissa
// operands, fp32 TMEM accumulators), TMA 128-byte-swizzled operand tiles, 128 x 128 output tile per CTA, split-K over
// gridDim.z:
//     C[m,n] = sum_k A[m*lda + k] * B[n*ldb + k]          (both operands K-major)
// kind::tf32 silently produces zeros for MN-major (transposed) shared-memory operands on sm_100a (measured), so the data
// gradient dx = dy W and the weight gradient dW = dy^T x get their transposed operand(s) from transpose_f32_kernel first
// (W^T is tiny; dy^T and x^T cost one extra read + write of tensors that are small next to the activations).
// Reference call sites: nn.Linear forward/backward of models/dino.py:223-226, 459-468, 1244-1248.
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200 {
namespace {

using namespace umma;

constexpr int GBM = 128, GBN = 128, GBK = 32;            // tile (GBK fp32 = one 128-byte swizzle row)
// Two pipeline depths: 6 stages (one CTA per SM) for long reductions; 3 stages (two CTAs per SM) for K-per-split <= 1024, where
// a tile is a handful of k-blocks and the second resident CTA hides the first one's pipeline fill and store epilogue.
constexpr int G_A_BYTES = GBM * GBK * 4, G_B_BYTES = GBN * GBK * 4, G_STAGE = G_A_BYTES + G_B_BYTES;
constexpr int g_smem(int stages) { return stages * G_STAGE + 1024 + 256; }

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return smem_desc(saddr, lbo_bytes, sbo_bytes) | (2ull << 61);      // layout type 2 = SWIZZLE_128B
}
__host__ __device__ constexpr uint32_t idesc_tf32(int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// epi: 0 store, 1 +bias, 2 +bias,ReLU, 3 +bias,ReLU,dropout(mask), 4 split-K partial (C holds [splits][M][ldc]),
//      5 InfoNCE: e = exp(x*fparam - fparam) stored as bf16 into C (ldc in bf16 elements) + per-(column tile, row) sums of e in aux
// BF16 = true: bf16 operands (kind::f16, UMMA K = 16, 64-element k-blocks); false: fp32 operands consumed as tf32.
template <bool BF16, int G_STAGES>
__global__ void __launch_bounds__(192, G_STAGES <= 3 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C, int64_t ldc, int M, int N,
               int K, int k_per_split, int epi, const float* __restrict__ bias, const uint8_t* __restrict__ mask,
               float keep_scale, float* __restrict__ aux, float fparam) {
    constexpr int GBKE = BF16 ? 64 : 32;                                 // k-block in elements (always 128 bytes)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                      // swizzle atoms need 1024-byte alignment
    uint8_t* smem = smem_raw + (base - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE);       // full[S], empty[S], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + G_STAGES * G_STAGE + 200);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (G_STAGES + s); };
    const uint32_t done_bar = bar0 + 8u * (2 * G_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
    const int nkb = (kend - kbeg + GBKE - 1) / GBKE;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<GBN>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % G_STAGES, use = kb / G_STAGES;
                mbar_wait(empty_bar(s), (use & 1) ^ 1);
                mbar_expect_tx(full_bar(s), G_STAGE);
                const uint32_t sa = base + s * G_STAGE, sb = sa + G_A_BYTES;
                const int k = kbeg + kb * GBKE;
                tma_load_2d(sa, &tmA, full_bar(s), k, m0);             // box (32 k, 128 rows), 128-byte swizzle
                tma_load_2d(sb, &tmB, full_bar(s), k, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = BF16 ? idesc_bf16(GBN) : idesc_tf32(GBN, false, false);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % G_STAGES, use = kb / G_STAGES;
                mbar_wait(full_bar(s), use & 1);
                tc_fence_after_sync();
                const uint32_t sa = base + s * G_STAGE, sb = sa + G_A_BYTES;
#pragma unroll
                for (int ks = 0; ks < GBK / 8; ++ks) {                 // UMMA K = 8 for tf32
                    // +32 bytes inside the swizzled 128-byte row per K step; 8-row groups are 1024 bytes apart
                    const uint64_t ad = desc_sw128(sa + ks * 32, 16, 1024);
                    const uint64_t bd = desc_sw128(sb + ks * 32, 16, 1024);
                    mma_tf32(tmem_base, ad, bd, idesc, (kb | ks) != 0);
                }
                mma_commit(empty_bar(s));
            }
            mma_commit(done_bar);
        }
    } else {
        const int quad = warp & 3;
        const int gm = m0 + quad * 32 + lane;
        if (nkb > 0) {
            mbar_wait(done_bar, 0);
            tc_fence_after_sync();
        }
        float* crow = C + (epi == 4 ? (int64_t)blockIdx.z * M * ldc : 0) + (int64_t)gm * ldc;
        float esum = 0.f;
        const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        const bool wide_ok = ((ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(C) & 31) == 0) && epi != 5;
#pragma unroll 1
        for (int cc = 0; cc < GBN / 16; ++cc) {
            uint32_t v[16];
            if (nkb > 0) {
                tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + cc * 16, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0u;
            }
            const int gn0 = n0 + cc * 16;
            if (gm >= M || gn0 >= N) continue;
            if (epi == 5) {
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float e0 = (gn0 + 2 * j < N) ? expf(__uint_as_float(v[2 * j]) * fparam - fparam) : 0.f;
                    const float e1 = (gn0 + 2 * j + 1 < N) ? expf(__uint_as_float(v[2 * j + 1]) * fparam - fparam) : 0.f;
                    esum += e0 + e1;
                    pk[j] = pack_bf16(e0, e1);
                }
                __nv_bfloat16* erow = reinterpret_cast<__nv_bfloat16*>(C) + (int64_t)gm * ldc + gn0;
                if (gn0 + 16 <= N && (ldc & 7) == 0) {
                    reinterpret_cast<uint4*>(erow)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    reinterpret_cast<uint4*>(erow)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (gn0 + j < N) reinterpret_cast<uint16_t*>(erow)[j] = (uint16_t)((j & 1) ? (pk[j >> 1] >> 16) : (pk[j >> 1] & 0xFFFFu));
                }
                continue;
            }
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int gn = gn0 + j;
                float x = __uint_as_float(v[j]);
                if (epi >= 1 && epi <= 3 && bias != nullptr && gn < N) x += __ldg(bias + gn);
                if (epi == 2 || epi == 3) x = fmaxf(x, 0.f);
                if (epi == 3 && gn < N) x = mask[(int64_t)gm * N + gn] ? x * keep_scale : 0.f;
                f[j] = x;
            }
            if (wide_ok && gn0 + 16 <= N) {             // a thread owns 64 contiguous bytes of its row: two 256-bit stores
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    st_global_256(crow + gn0 + 8 * j, make_uint4(__float_as_uint(f[8 * j]), __float_as_uint(f[8 * j + 1]), __float_as_uint(f[8 * j + 2]), __float_as_uint(f[8 * j + 3])),
                                  make_uint4(__float_as_uint(f[8 * j + 4]), __float_as_uint(f[8 * j + 5]), __float_as_uint(f[8 * j + 6]), __float_as_uint(f[8 * j + 7])));
            } else if (vec_ok && gn0 + 16 <= N) {
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(crow + gn0 + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (gn0 + j < N) crow[gn0 + j] = f[j];
            }
        }
        if (epi == 5 && gm < M) aux[(int64_t)blockIdx.x * M + gm] = esum;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<GBN>(tmem_base);
    }
}

int encode_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems, uint32_t box_inner,
                       uint32_t box_outer, bool swizzle128 = true, bool bf16 = false) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled entry point not available (%d)", (int)e);
            return -30;
        }
        fn = (EncodeFn)p;
    }
    cuuint64_t d[2] = {inner, outer};
    cuuint64_t s[1] = {ld_elems * (bf16 ? 2 : 4)};
    cuuint32_t b[2] = {box_inner, box_outer};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), d, s, b, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (fp32 2d) failed (%d)", (int)r);
        return -31;
    }
    return 0;
}

}  // namespace

// TMA needs 16-byte aligned bases and row pitches
bool gemm_tc_usable(const void* A, int64_t lda, const void* Bm, int64_t ldb) {
    return ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bm)) & 15) == 0 && (lda & 3) == 0 && (ldb & 3) == 0;
}

// C[M,N] = A x B^T, both operands K-major.  splits > 1: C must hold [splits][M][ldc] partials (epi is forced to 4).
// bf16 = 1: A and B are bf16 (lda / ldb in bf16 elements, multiples of 8); aux / fparam: see epi 5.
int launch_gemm_tc_ex(const void* A, int64_t lda, const void* Bm, int64_t ldb, void* C, int64_t ldc, int M, int N, int K, int splits, int epi,
                      const float* bias, const uint8_t* mask, float keep_scale, float* aux, float fparam, int bf16, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem(6));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem(6));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem(3));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem(3));
        if (e != cudaSuccess) {
            set_error("gemm_tc: cannot set %d bytes of shared memory: %s", g_smem(6), cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    const int gbke = bf16 ? 64 : 32;
    CUtensorMap ta, tb;
    int rc = encode_tmap_f32_2d(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, gbke, GBM, true, bf16 != 0);
    if (rc) return rc;
    rc = encode_tmap_f32_2d(&tb, Bm, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, gbke, GBN, true, bf16 != 0);
    if (rc) return rc;
    if (splits < 1) splits = 1;
    int kps = ((K + splits - 1) / splits + gbke - 1) / gbke * gbke;
    splits = (K + kps - 1) / kps;
    dim3 grid((N + GBN - 1) / GBN, (M + GBM - 1) / GBM, splits);
    const bool shallow = kps <= 1024 && (long)grid.x * grid.y * grid.z > sm_count();      // enough tiles for two CTAs per SM
    const int e = splits > 1 ? 4 : epi;
    float* Cf = reinterpret_cast<float*>(C);
    if (bf16 && shallow)
        gemm_tc_kernel<true, 3><<<grid, 192, g_smem(3), st>>>(ta, tb, Cf, ldc, M, N, K, kps, e, bias, mask, keep_scale, aux, fparam);
    else if (bf16)
        gemm_tc_kernel<true, 6><<<grid, 192, g_smem(6), st>>>(ta, tb, Cf, ldc, M, N, K, kps, e, bias, mask, keep_scale, aux, fparam);
    else if (shallow)
        gemm_tc_kernel<false, 3><<<grid, 192, g_smem(3), st>>>(ta, tb, Cf, ldc, M, N, K, kps, e, bias, mask, keep_scale, aux, fparam);
    else
        gemm_tc_kernel<false, 6><<<grid, 192, g_smem(6), st>>>(ta, tb, Cf, ldc, M, N, K, kps, e, bias, mask, keep_scale, aux, fparam);
    return launch_status("gemm_tc_kernel");
}

int launch_gemm_tc(const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc, int M, int N, int K, int splits, int epi,
                   const float* bias, const uint8_t* mask, float keep_scale, cudaStream_t st) {
    return launch_gemm_tc_ex(A, lda, Bm, ldb, C, ldc, M, N, K, splits, epi, bias, mask, keep_scale, nullptr, 0.f, 0, st);
}

// dst[c*ldd + r] = src[r*lds + c], 32 x 32 tiles through shared memory (coalesced both ways)
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd,
                                                            int R, int Cc) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + tx;
        tile[ty + 8 * i][tx] = (r < R && c < Cc) ? __ldg(src + (int64_t)r * lds + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        if (c < Cc && r < R) dst[(int64_t)c * ldd + r] = tile[tx][ty + 8 * i];
    }
}

int launch_transpose_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int R, int Cc, cudaStream_t st) {
    transpose_f32_kernel<<<dim3((Cc + 31) / 32, (R + 31) / 32), 256, 0, st>>>(src, lds, dst, ldd, R, Cc);
    return launch_status("transpose_f32_kernel");
}

int gemm_tc_splits(int M, int N, int K) {         // split the reduction so that about one wave of CTAs is in flight
    const int tiles = ((N + GBN - 1) / GBN) * ((M + GBM - 1) / GBM);
    int splits = (sm_count() + tiles - 1) / tiles;
    const int max_splits = (K + 4 * GBK - 1) / (4 * GBK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int kps = ((K + splits - 1) / splits + GBK - 1) / GBK * GBK;
    return (K + kps - 1) / kps;
}

}  // namespace b200
