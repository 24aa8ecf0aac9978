// Linear layers (FP32 SIMT GEMM tiles in this revision) and their fused element-wise middles:
// bias / ReLU / dropout epilogues, BatchNorm1d + GELU(erf) + dropout forward and backward, column statistics.
#include "common.cuh"

namespace b200 {

// tensor-core GEMM (gemm_tc.cu)
bool gemm_tc_usable(const void* A, int64_t lda, const void* Bm, int64_t ldb);
int launch_gemm_tc(const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc, int M, int N, int K, int splits, int epi,
                   const float* bias, const uint8_t* mask, float keep_scale, cudaStream_t st);
int launch_transpose_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int R, int Cc, cudaStream_t st);
int gemm_tc_splits(int M, int N, int K);

constexpr int BM = 64, BN = 64, BK = 16, GT = 256, LDS_PAD = 4;

// C[m,n] (+)= sum_k A(m,k) * B(n,k)
//   A_K: A element (m,k) at A[m*lda + k] (k contiguous) else at A[k*lda + m] (m contiguous); same for B.
// EPI: 0 store, 1 bias, 2 bias+relu, 3 bias+relu+dropout(mask), 4 atomicAdd (split-K)
template <bool A_K, bool B_K, int EPI>
__global__ void __launch_bounds__(GT) gemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bm, int64_t ldb,
                                                  float* __restrict__ C, int64_t ldc, int M, int N, int K, int k_per_split,
                                                  const float* __restrict__ bias, const uint8_t* __restrict__ mask, float keep_scale) {
    __shared__ __align__(16) float As[BK][BM + LDS_PAD];
    __shared__ __align__(16) float Bs[BK][BN + LDS_PAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- stage A tile ----
        if (A_K) {
            const int r = tid >> 2, kq = (tid & 3) * 4;
            const int gm = m0 + r, gk = k0 + kq;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (gm < M) {
                const float* p = A + (int64_t)gm * lda + gk;
                if (gk + 3 < kend && ((((uintptr_t)p) & 15) == 0)) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(p));
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (gk + i < kend) v[i] = __ldg(p + i);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) As[kq + i][r] = v[i];
        } else {
            const int kr = tid >> 4, mq = (tid & 15) * 4;
            const int gk = k0 + kr, gm = m0 + mq;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (gk < kend) {
                const float* p = A + (int64_t)gk * lda + gm;
                if (gm + 3 < M && ((((uintptr_t)p) & 15) == 0)) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(p));
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (gm + i < M) v[i] = __ldg(p + i);
                }
            }
            *reinterpret_cast<float4*>(&As[kr][mq]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        // ---- stage B tile ----
        if (B_K) {
            const int r = tid >> 2, kq = (tid & 3) * 4;
            const int gn = n0 + r, gk = k0 + kq;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (gn < N) {
                const float* p = Bm + (int64_t)gn * ldb + gk;
                if (gk + 3 < kend && ((((uintptr_t)p) & 15) == 0)) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(p));
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (gk + i < kend) v[i] = __ldg(p + i);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[kq + i][r] = v[i];
        } else {
            const int kr = tid >> 4, nq = (tid & 15) * 4;
            const int gk = k0 + kr, gn = n0 + nq;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (gk < kend) {
                const float* p = Bm + (int64_t)gk * ldb + gn;
                if (gn + 3 < N && ((((uintptr_t)p) & 15) == 0)) {
                    float4 t = __ldg(reinterpret_cast<const float4*>(p));
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (gn + i < N) v[i] = __ldg(p + i);
                }
            }
            *reinterpret_cast<float4*>(&Bs[kr][nq]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            float* cp = C + (int64_t)gm * ldc + gn;
            if (EPI == 4) {
                cp[(int64_t)blockIdx.z * M * ldc] = v;       // split-K partial: C holds [splits][M][ldc], reduced in fixed order
            } else {
                if (EPI >= 1 && bias) v += __ldg(bias + gn);
                if (EPI >= 2) v = fmaxf(v, 0.f);
                if (EPI == 3) v = mask[(int64_t)gm * N + gn] ? v * keep_scale : 0.f;
                *cp = v;
            }
        }
    }
}

// out[i] = (accumulate ? out[i] : 0) + sum_p part[p][i], p in fixed order (deterministic split-K reduction)
__global__ void __launch_bounds__(256) reduce_splits_kernel(const float* __restrict__ part, int n_parts, int64_t n, float* __restrict__ out,
                                                            int accumulate) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float a = accumulate ? out[i] : 0.f;
        for (int p = 0; p < n_parts; ++p) a += part[(int64_t)p * n + i];
        out[i] = a;
    }
}

// column sums of dy[M,N] -> db[N]   (one warp per 32 columns, grid-stride over row blocks, deterministic tree)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, int64_t ld, int M, int N, float* __restrict__ out,
                                                     int accumulate) {
    // block = 32 columns x 8 row-lanes
    __shared__ float red[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    float a = 0.f;
    if (col < N)
        for (int m = rl; m < M; m += 8) a += __ldg(dy + (int64_t)m * ld + col);
    red[rl][threadIdx.x & 31] = a;
    __syncthreads();
    if (rl == 0 && col < N) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x & 31];
        out[col] = accumulate ? out[col] + s : s;
    }
}

// partial column sums: part[blockIdx.y][col] = sum over this block's row slice (reduced in fixed order by reduce_splits_kernel)
__global__ void __launch_bounds__(256) colsum_part_kernel(const float* __restrict__ dy, int64_t ld, int M, int N, float* __restrict__ part) {
    __shared__ float red[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    const int rows = (M + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * rows, r1 = min(M, r0 + rows);
    float a = 0.f;
    if (col < N)
        for (int m = r0 + rl; m < r1; m += 8) a += __ldg(dy + (int64_t)m * ld + col);
    red[rl][threadIdx.x & 31] = a;
    __syncthreads();
    if (rl == 0 && col < N) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x & 31];
        part[(int64_t)blockIdx.y * N + col] = s;
    }
}

// column statistics in double: stats[c] = {sum, sum of squares}; grid = (col blocks, row splits), atomics on doubles
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ h, double* __restrict__ stats, int M, int C) {
    __shared__ double r1[8][33], r2[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    double a = 0.0, b = 0.0;
    if (col < C)
        for (int m = blockIdx.y * 8 + rl; m < M; m += gridDim.y * 8) {
            const float v = __ldg(h + (int64_t)m * C + col);
            a += v;
            b += (double)v * v;
        }
    r1[rl][threadIdx.x & 31] = a;
    r2[rl][threadIdx.x & 31] = b;
    __syncthreads();
    if (rl == 0 && col < C) {
        double s = 0.0, q = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) { s += r1[r][threadIdx.x & 31]; q += r2[r][threadIdx.x & 31]; }
        atomicAdd(&stats[col * 2], s);
        atomicAdd(&stats[col * 2 + 1], q);
    }
}

__device__ __forceinline__ float gelu_erf(float y) { return 0.5f * y * (1.0f + erff(y * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float y) {
    const float cdf = 0.5f * (1.0f + erff(y * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * y * y);
    return cdf + y * pdf;
}

__global__ void __launch_bounds__(256) bn1d_gelu_drop_fwd_kernel(const float* __restrict__ h, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, const uint8_t* __restrict__ mask,
                                                                 float keep_scale, float* __restrict__ g, int64_t n, int C) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float v = gelu_erf(fmaf(__ldg(scale + c), __ldg(h + i), __ldg(shift + c)));
        if (mask) v = mask[i] ? v * keep_scale : 0.f;
        g[i] = v;
    }
}

// backward: dy = dg * keep * gelu'(y);  pass 1: column sums of dy and dy*xhat;  pass 2: dh = scale*(dy - s1/M - xhat*s2/M)
template <bool APPLY>
__global__ void __launch_bounds__(256) bn1d_gelu_drop_bwd_kernel(const float* __restrict__ h, const float* __restrict__ dg,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 const uint8_t* __restrict__ mask, float keep_scale,
                                                                 double* __restrict__ sums, float* __restrict__ dh, int M, int C) {
    __shared__ double r1[8][33], r2[8][33];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    double a = 0.0, b = 0.0;
    if (col < C) {
        const float sc = __ldg(scale + col), sh = __ldg(shift + col), mu = __ldg(mean + col), is = __ldg(invstd + col);
        float k1 = 0.f, k2 = 0.f;
        if (APPLY) {
            k1 = (float)(sums[col * 2] / (double)M);
            k2 = (float)(sums[col * 2 + 1] / (double)M);
        }
        for (int m = blockIdx.y * 8 + rl; m < M; m += gridDim.y * 8) {
            const int64_t i = (int64_t)m * C + col;
            const float hv = __ldg(h + i);
            const float y = fmaf(sc, hv, sh);
            float dy = __ldg(dg + i) * gelu_erf_grad(y);
            if (mask) dy = mask[i] ? dy * keep_scale : 0.f;
            const float xh = (hv - mu) * is;
            if (!APPLY) {
                a += dy;
                b += (double)dy * xh;
            } else {
                dh[i] = sc * (dy - k1 - xh * k2);
            }
        }
    }
    if (!APPLY) {
        r1[rl][threadIdx.x & 31] = a;
        r2[rl][threadIdx.x & 31] = b;
        __syncthreads();
        if (rl == 0 && col < C) {
            double s = 0.0, q = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) { s += r1[r][threadIdx.x & 31]; q += r2[r][threadIdx.x & 31]; }
            atomicAdd(&sums[col * 2], s);
            atomicAdd(&sums[col * 2 + 1], q);
        }
    }
}

__global__ void __launch_bounds__(256) act_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, float keep_scale, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dy[i] = __ldg(y + i) > 0.f ? dy[i] * keep_scale : 0.f;
}

static int ew_grid(int64_t n) {
    int64_t g = (n + 1023) / 1024;
    int64_t cap = (int64_t)sm_count() * 8;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}
static int row_splits(int M, int col_blocks) {
    int want = (sm_count() * 4 + col_blocks - 1) / col_blocks;
    int maxs = (M + 63) / 64;
    if (want > maxs) want = maxs;
    return want < 1 ? 1 : want;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_linear_fwd(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t ldy, int M, int N, int K,
                    int act, const uint8_t* mask, float drop_p, void* stream) {
    B200_REQUIRE(x && w && y && M > 0 && N > 0 && K > 0 && ldx >= K && ldy >= N, B200_E_ARG, "linear_fwd: bad arguments");
    B200_REQUIRE(act >= 0 && act <= 2 && (act != 2 || (mask && drop_p >= 0.f && drop_p < 1.f)), B200_E_ARG, "linear_fwd: bad activation");
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
    cudaStream_t st = as_stream(stream);
    const float ks = 1.0f / (1.0f - drop_p);
    if (act == 0) gemm_kernel<true, true, 1><<<grid, GT, 0, st>>>(x, ldx, w, K, y, ldy, M, N, K, K, bias, nullptr, 1.f);
    else if (act == 1) gemm_kernel<true, true, 2><<<grid, GT, 0, st>>>(x, ldx, w, K, y, ldy, M, N, K, K, bias, nullptr, 1.f);
    else gemm_kernel<true, true, 3><<<grid, GT, 0, st>>>(x, ldx, w, K, y, ldy, M, N, K, K, bias, mask, ks);
    return launch_status("linear_fwd");
}

int b200_linear_bwd_data(const float* dy, int64_t lddy, const float* w, float* dx, int64_t lddx, int M, int N, int K, void* stream) {
    B200_REQUIRE(dy && w && dx && M > 0 && N > 0 && K > 0 && lddy >= N && lddx >= K, B200_E_ARG, "linear_bwd_data: bad arguments");
    // dx[m,k] = sum_n dy[m,n] w[n,k]:  C[M,K]; A = dy (reduction index n contiguous), B(k, n) = w[n*K + k] (k contiguous)
    dim3 grid((K + BN - 1) / BN, (M + BM - 1) / BM, 1);
    gemm_kernel<true, false, 0><<<grid, GT, 0, as_stream(stream)>>>(dy, lddy, w, K, dx, lddx, M, K, N, N, nullptr, nullptr, 1.f);
    return launch_status("linear_bwd_data");
}

static int wgrad_splits(int M, int N, int K) {
    const int tiles = ((K + BN - 1) / BN) * ((N + BM - 1) / BM);
    int splits = (sm_count() * 2 + tiles - 1) / tiles;
    const int max_splits = (M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int kps = ((M + splits - 1) / splits + BK - 1) / BK * BK;
    return (M + kps - 1) / kps;
}

int64_t b200_linear_bwd_weight_work_floats(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const int splits = wgrad_splits(M, N, K);
    return (int64_t)splits * N * K;      // also covers splits == 1 with accumulate
}

int b200_linear_bwd_weight(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dw, float* db, float* work, int M, int N,
                           int K, int accumulate, void* stream) {
    B200_REQUIRE(dy && x && dw && M > 0 && N > 0 && K > 0 && lddy >= N && ldx >= K, B200_E_ARG, "linear_bwd_weight: bad arguments");
    cudaStream_t st = as_stream(stream);
    // dw[n,k] = sum_m dy[m,n] x[m,k]:  C[N,K]; A(n, m) = dy[m*lddy + n] (n contiguous), B(k, m) = x[m*ldx + k] (k contiguous)
    dim3 grid((K + BN - 1) / BN, (N + BM - 1) / BM, 1);
    const int splits = wgrad_splits(M, N, K);
    const int kps = ((M + splits - 1) / splits + BK - 1) / BK * BK;
    if (splits == 1 && !accumulate) {
        gemm_kernel<false, false, 0><<<grid, GT, 0, st>>>(dy, lddy, x, ldx, dw, K, N, K, M, M, nullptr, nullptr, 1.f);
    } else {
        B200_REQUIRE(work, B200_E_ARG, "linear_bwd_weight: split-K needs the work buffer (b200_linear_bwd_weight_work_floats)");
        grid.z = splits;
        gemm_kernel<false, false, 4><<<grid, GT, 0, st>>>(dy, lddy, x, ldx, work, K, N, K, M, kps, nullptr, nullptr, 1.f);
        const int64_t n = (int64_t)N * K;
        reduce_splits_kernel<<<(int)((n + 1023) / 1024 < sm_count() * 8 ? (n + 1023) / 1024 : sm_count() * 8), 256, 0, st>>>(work, splits, n, dw,
                                                                                                                          accumulate);
    }
    int rc = launch_status("linear_bwd_weight");
    if (rc || !db) return rc;
    colsum_kernel<<<(N + 31) / 32, 256, 0, st>>>(dy, lddy, M, N, db, accumulate);
    return launch_status("linear_bwd_bias");
}

// ---- tensor-core (tcgen05, tf32) variants: same arguments and results up to tf32 operand rounding; operands that TMA cannot
//      address (row pitch not a multiple of 16 bytes, e.g. the 10-way classifier outputs) run on the SIMT kernels ----
static int colsum_splits(int M) {
    int rs = (M + 511) / 512;
    return rs < 1 ? 1 : (rs > 64 ? 64 : rs);
}

int b200_linear_fwd_tc(const float* x, int64_t ldx, const float* w, const float* bias, float* y, int64_t ldy, int M, int N, int K,
                       int act, const uint8_t* mask, float drop_p, void* stream) {
    B200_REQUIRE(x && w && y && M > 0 && N > 0 && K > 0 && ldx >= K && ldy >= N, B200_E_ARG, "linear_fwd_tc: bad arguments");
    B200_REQUIRE(act >= 0 && act <= 2 && (act != 2 || (mask && drop_p >= 0.f && drop_p < 1.f)), B200_E_ARG, "linear_fwd_tc: bad activation");
    if (!gemm_tc_usable(x, ldx, w, K)) return b200_linear_fwd(x, ldx, w, bias, y, ldy, M, N, K, act, mask, drop_p, stream);
    return launch_gemm_tc(x, ldx, w, K, y, ldy, M, N, K, 1, act + 1, bias, mask, 1.0f / (1.0f - drop_p), as_stream(stream));
}

static int64_t pad4(int64_t n) { return (n + 3) / 4 * 4; }

int64_t b200_linear_bwd_data_tc_work_floats(int M, int N, int K) {
    (void)M;
    return (N > 0 && K > 0) ? (int64_t)K * pad4(N) : 0;
}

int b200_linear_bwd_data_tc(const float* dy, int64_t lddy, const float* w, float* dx, int64_t lddx, float* work, int M, int N, int K,
                            void* stream) {
    B200_REQUIRE(dy && w && dx && work && M > 0 && N > 0 && K > 0 && lddy >= N && lddx >= K, B200_E_ARG, "linear_bwd_data_tc: bad arguments");
    if (!gemm_tc_usable(dy, lddy, work, 4)) return b200_linear_bwd_data(dy, lddy, w, dx, lddx, M, N, K, stream);
    cudaStream_t st = as_stream(stream);
    // dx[m,k] = sum_n dy[m,n] w[n,k] = sum_n dy[m,n] wT[k,n]:  B = w^T [K][N] (K-major in the reduction index n)
    const int64_t ldt = pad4(N);
    int rc = launch_transpose_f32(w, K, work, ldt, N, K, st);
    if (rc) return rc;
    return launch_gemm_tc(dy, lddy, work, ldt, dx, lddx, M, K, N, 1, 0, nullptr, nullptr, 1.f, st);
}

int64_t b200_linear_bwd_weight_tc_work_floats(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const int64_t simt = b200_linear_bwd_weight_work_floats(M, N, K);
    const int64_t tc = (int64_t)gemm_tc_splits(N, K, M) * N * K + (int64_t)colsum_splits(M) * N + (int64_t)(N + K) * pad4(M);
    return simt > tc ? simt : tc;
}

int b200_linear_bwd_weight_tc(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dw, float* db, float* work, int M, int N,
                              int K, int accumulate, void* stream) {
    B200_REQUIRE(dy && x && dw && work && M > 0 && N > 0 && K > 0 && lddy >= N && ldx >= K, B200_E_ARG, "linear_bwd_weight_tc: bad arguments");
    if ((reinterpret_cast<uintptr_t>(work) & 15) != 0) return b200_linear_bwd_weight(dy, lddy, x, ldx, dw, db, work, M, N, K, accumulate, stream);
    cudaStream_t st = as_stream(stream);
    // dw[n,k] = sum_m dy[m,n] x[m,k] = sum_m dyT[n,m] xT[k,m]: both transposed copies are K-major in the reduction index m
    const int splits = gemm_tc_splits(N, K, M);
    const int64_t nk = (int64_t)N * K, ldt = pad4(M);
    float* part = work + (int64_t)splits * nk;                 // colsum partials
    const int rs = colsum_splits(M);
    float* dyT = part + (int64_t)rs * N;
    dyT += (4 - ((dyT - work) & 3)) & 3;                      // keep 16-byte alignment
    float* xT = dyT + (int64_t)N * ldt;
    int rc = launch_transpose_f32(dy, lddy, dyT, ldt, M, N, st);
    if (rc) return rc;
    rc = launch_transpose_f32(x, ldx, xT, ldt, M, K, st);
    if (rc) return rc;
    if (splits == 1 && !accumulate) {
        rc = launch_gemm_tc(dyT, ldt, xT, ldt, dw, K, N, K, M, 1, 0, nullptr, nullptr, 1.f, st);
    } else {
        rc = launch_gemm_tc(dyT, ldt, xT, ldt, work, K, N, K, M, splits, 4, nullptr, nullptr, 1.f, st);
        if (rc) return rc;
        const int kps = ((M + splits - 1) / splits + 31) / 32 * 32;
        const int eff = (M + kps - 1) / kps;
        reduce_splits_kernel<<<(int)((nk + 1023) / 1024 < sm_count() * 8 ? (nk + 1023) / 1024 : sm_count() * 8), 256, 0, st>>>(work, eff, nk, dw,
                                                                                                                            accumulate);
        rc = launch_status("linear_bwd_weight_tc reduce");
    }
    if (rc || !db) return rc;
    colsum_part_kernel<<<dim3((N + 31) / 32, rs), 256, 0, st>>>(dy, lddy, M, N, part);
    reduce_splits_kernel<<<(N + 255) / 256, 256, 0, st>>>(part, rs, N, db, accumulate);
    return launch_status("linear_bwd_bias_tc");
}

int b200_act_bwd(float* dy, const float* y, const uint8_t* mask, float drop_p, int64_t n, void* stream) {
    (void)mask;   // y already carries the dropout zeros: y > 0 <=> unit active and kept
    B200_REQUIRE(dy && y && n > 0 && drop_p >= 0.f && drop_p < 1.f, B200_E_ARG, "act_bwd: bad arguments");
    act_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(dy, y, 1.0f / (1.0f - drop_p), n);
    return launch_status("act_bwd");
}

int b200_colstats(const float* h, double* stats, int M, int C, void* stream) {
    B200_REQUIRE(h && stats && M > 0 && C > 0, B200_E_ARG, "colstats: bad arguments");
    const int cb = (C + 31) / 32;
    colstats_kernel<<<dim3(cb, row_splits(M, cb)), 256, 0, as_stream(stream)>>>(h, stats, M, C);
    return launch_status("colstats");
}

int b200_bn1d_gelu_drop_fwd(const float* h, const float* scale, const float* shift, const uint8_t* mask, float drop_p, float* g,
                            int M, int C, void* stream) {
    B200_REQUIRE(h && scale && shift && g && M > 0 && C > 0 && drop_p >= 0.f && drop_p < 1.f, B200_E_ARG, "bn1d_gelu_drop_fwd: bad arguments");
    const int64_t n = (int64_t)M * C;
    bn1d_gelu_drop_fwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(h, scale, shift, drop_p > 0.f ? mask : nullptr,
                                                                         1.0f / (1.0f - drop_p), g, n, C);
    return launch_status("bn1d_gelu_drop_fwd");
}

int b200_bn1d_gelu_drop_bwd_reduce(const float* h, const float* dg, const float* scale, const float* shift, const float* mean,
                                   const float* invstd, const uint8_t* mask, float drop_p, double* sums, int M, int C, void* stream) {
    B200_REQUIRE(h && dg && scale && shift && mean && invstd && sums && M > 0 && C > 0, B200_E_ARG, "bn1d_gelu_drop_bwd_reduce: bad arguments");
    const int cb = (C + 31) / 32;
    bn1d_gelu_drop_bwd_kernel<false><<<dim3(cb, row_splits(M, cb)), 256, 0, as_stream(stream)>>>(
        h, dg, scale, shift, mean, invstd, drop_p > 0.f ? mask : nullptr, 1.0f / (1.0f - drop_p), sums, nullptr, M, C);
    return launch_status("bn1d_gelu_drop_bwd_reduce");
}

int b200_bn1d_gelu_drop_bwd_apply(const float* h, const float* dg, const float* scale, const float* shift, const float* mean,
                                  const float* invstd, const uint8_t* mask, float drop_p, const double* sums, float* dh, int M, int C,
                                  void* stream) {
    B200_REQUIRE(h && dg && scale && shift && mean && invstd && sums && dh && M > 0 && C > 0, B200_E_ARG, "bn1d_gelu_drop_bwd_apply: bad arguments");
    const int cb = (C + 31) / 32;
    bn1d_gelu_drop_bwd_kernel<true><<<dim3(cb, row_splits(M, cb)), 256, 0, as_stream(stream)>>>(
        h, dg, scale, shift, mean, invstd, drop_p > 0.f ? mask : nullptr, 1.0f / (1.0f - drop_p), const_cast<double*>(sums), dh, M, C);
    return launch_status("bn1d_gelu_drop_bwd_apply");
}

}  // extern "C"
