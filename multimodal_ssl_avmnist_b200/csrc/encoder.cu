// Small-conv encoder kernels (FP32 direct convolution on the CUDA cores, shared-memory resident tiles):
//   conv forward (+bias, +per-view BatchNorm statistics), data gradient (same kernel, flipped weights),
//   weight gradient (output-stationary, deterministic two-pass reduction), BatchNorm finalisation and the fused
//   BN-apply + ReLU + 2x2 max-pool forward / backward, global average pool.
// Layout: NCHW fp32.  A launch covers all view-calls of a step (N = n_views * n_per_view); BatchNorm statistics
// stay segmented per view-call.
#include "common.cuh"

namespace b200 {

// =========================================================================================================
// Direct convolution, one CTA = (sample, row tile), all output channels, input channels in chunks of CC.
//   thread = (cout group of TCO channels) x (output row) x (strip of SX output pixels)
//   inner loop per (cin, ky): SX+K-1 input values from smem, K*TCO weights (broadcast float4), SX*K*TCO FMAs.
// FLIP: weights are read as w[co_in][ci_out] spatially flipped => the same kernel computes the data gradient.
// =========================================================================================================
template <int CIN, int COUT, int H, int W, int K, int PAD, int TY, int SX, int TCO, int CC, bool FLIP, bool STATS>
struct ConvCfg {
    static constexpr int HO = H + 2 * PAD - K + 1;
    static constexpr int WO = W + 2 * PAD - K + 1;
    static constexpr int NSTRIP = (WO + SX - 1) / SX;
    static constexpr int NTILE = (HO + TY - 1) / TY;
    static constexpr int NGROUP = COUT / TCO;
    static constexpr int NPOS = TY * NSTRIP;
    static constexpr int THREADS = ((NPOS * NGROUP + 31) / 32) * 32;
    static constexpr int TIH = TY + K - 1;                                  // input tile rows
    static constexpr int TIW = ((NSTRIP * SX + K - 1 + 3) / 4) * 4;         // input tile row pitch (floats)
    static constexpr int NCHUNK = CIN / CC;
    static constexpr int IN_FLOATS = CC * TIH * TIW;
    static constexpr int W_FLOATS = CC * K * K * COUT;
    static constexpr size_t SMEM = (size_t)(IN_FLOATS + W_FLOATS) * sizeof(float) + (STATS ? COUT * 2 * sizeof(double) : 0);
    static_assert(CIN % CC == 0 && COUT % TCO == 0 && TCO % 4 == 0, "bad conv tiling");
};

template <int CIN, int COUT, int H, int W, int K, int PAD, int TY, int SX, int TCO, int CC, bool FLIP, bool STATS>
__global__ void __launch_bounds__(ConvCfg<CIN, COUT, H, W, K, PAD, TY, SX, TCO, CC, FLIP, STATS>::THREADS)
conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ z,
            double* __restrict__ stats, int N, int n_per_view) {
    using C = ConvCfg<CIN, COUT, H, W, K, PAD, TY, SX, TCO, CC, FLIP, STATS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_in = reinterpret_cast<float*>(smem_raw);
    float* s_w = s_in + C::IN_FLOATS;
    double* s_stat = reinterpret_cast<double*>(s_w + C::W_FLOATS);

    const int tid = threadIdx.x;
    const bool active = tid < C::NPOS * C::NGROUP;
    const int pos = tid % C::NPOS, grp = tid / C::NPOS;
    const int ly = pos / C::NSTRIP, lx = (pos % C::NSTRIP) * SX;
    const int co0 = grp * TCO;

    float ssum[TCO], ssq[TCO];
#pragma unroll
    for (int t = 0; t < TCO; ++t) ssum[t] = ssq[t] = 0.f;
    int cur_view = -1;
    if (STATS) {
        for (int i = tid; i < COUT * 2; i += C::THREADS) s_stat[i] = 0.0;
    }
    bool w_loaded = false;

    const int n_items = N * C::NTILE;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int n = item / C::NTILE, tile = item - n * C::NTILE;
        const int y0 = tile * TY;
        if (STATS) {
            const int view = n / n_per_view;
            if (view != cur_view) {
                if (cur_view >= 0) {   // flush the statistics of the previous view-call
                    if (active) {
#pragma unroll
                        for (int t = 0; t < TCO; ++t) {
                            atomicAdd(&s_stat[(co0 + t) * 2], (double)ssum[t]);
                            atomicAdd(&s_stat[(co0 + t) * 2 + 1], (double)ssq[t]);
                            ssum[t] = ssq[t] = 0.f;
                        }
                    }
                    __syncthreads();
                    for (int i = tid; i < COUT * 2; i += C::THREADS) {
                        atomicAdd(&stats[(size_t)cur_view * COUT * 2 + i], s_stat[i]);
                        s_stat[i] = 0.0;
                    }
                    __syncthreads();
                }
                cur_view = view;
            }
        }
        float acc[SX][TCO];
#pragma unroll
        for (int p = 0; p < SX; ++p)
#pragma unroll
            for (int t = 0; t < TCO; ++t) acc[p][t] = 0.f;

        for (int ch = 0; ch < C::NCHUNK; ++ch) {
            __syncthreads();   // previous chunk / item fully consumed
            // ---- input tile: rows y0-PAD .. y0-PAD+TIH-1, cols -PAD .. ; zero outside the image ----
            const float* xin = x + ((size_t)n * CIN + ch * CC) * H * W;
            for (int e = tid; e < C::IN_FLOATS; e += C::THREADS) {
                const int c = e / (C::TIH * C::TIW), r = (e / C::TIW) % C::TIH, col = e % C::TIW;
                const int gy = y0 - PAD + r, gx = col - PAD;
                float v = 0.f;
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(xin + ((size_t)c * H + gy) * W + gx);
                s_in[e] = v;
            }
            // ---- weights of this chunk: s_w[c][ky][kx][co] ----
            if (C::NCHUNK > 1 || !w_loaded) {
                for (int e = tid; e < C::W_FLOATS; e += C::THREADS) {
                    const int co = e % COUT, kk = (e / COUT) % (K * K), c = e / (COUT * K * K);
                    float v;
                    if (!FLIP) v = __ldg(w + ((size_t)co * CIN + ch * CC + c) * K * K + kk);
                    else v = __ldg(w + ((size_t)(ch * CC + c) * COUT + co) * K * K + (K * K - 1 - kk));
                    s_w[e] = v;
                }
                w_loaded = true;
            }
            __syncthreads();
            if (active) {
#pragma unroll 1
                for (int c = 0; c < CC; ++c) {
#pragma unroll
                    for (int ky = 0; ky < K; ++ky) {
                        float in[SX + K - 1];
                        const float* row = s_in + (c * C::TIH + ly + ky) * C::TIW + lx;
                        if (SX % 4 == 0 && (SX + K - 1) % 4 == 0) {
#pragma unroll
                            for (int q = 0; q < (SX + K - 1) / 4; ++q) {
                                float4 v = reinterpret_cast<const float4*>(row)[q];
                                in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < SX + K - 1; ++q) in[q] = row[q];
                        }
                        const float* wp = s_w + ((c * K + ky) * K) * COUT + co0;
#pragma unroll
                        for (int kx = 0; kx < K; ++kx) {
                            float wv[TCO];
#pragma unroll
                            for (int q = 0; q < TCO / 4; ++q) {
                                float4 v = reinterpret_cast<const float4*>(wp + kx * COUT)[q];
                                wv[4 * q] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w;
                            }
#pragma unroll
                            for (int p = 0; p < SX; ++p)
#pragma unroll
                                for (int t = 0; t < TCO; ++t) acc[p][t] = fmaf(in[p + kx], wv[t], acc[p][t]);
                        }
                    }
                }
            }
        }
        // ---- epilogue: bias, store, statistics ----
        const int oy = y0 + ly;
        if (active && oy < C::HO) {
#pragma unroll
            for (int t = 0; t < TCO; ++t) {
                const float bv = bias ? __ldg(bias + co0 + t) : 0.f;
                float* zp = z + (((size_t)n * COUT + co0 + t) * C::HO + oy) * C::WO + lx;
#pragma unroll
                for (int p = 0; p < SX; ++p) {
                    if (lx + p < C::WO) {
                        const float v = acc[p][t] + bv;
                        zp[p] = v;
                        if (STATS) {
                            ssum[t] += v;
                            ssq[t] = fmaf(v, v, ssq[t]);
                        }
                    }
                }
            }
        }
    }
    if (STATS && cur_view >= 0) {
        if (active) {
#pragma unroll
            for (int t = 0; t < TCO; ++t) {
                atomicAdd(&s_stat[(co0 + t) * 2], (double)ssum[t]);
                atomicAdd(&s_stat[(co0 + t) * 2 + 1], (double)ssq[t]);
            }
        }
        __syncthreads();
        for (int i = tid; i < COUT * 2; i += C::THREADS) atomicAdd(&stats[(size_t)cur_view * COUT * 2 + i], s_stat[i]);
    }
}

template <int CIN, int COUT, int H, int W, int K, int PAD, int TY, int SX, int TCO, int CC, bool FLIP, bool STATS>
static int launch_conv(const float* x, const float* w, const float* bias, float* z, double* stats, int N, int n_per_view,
                       cudaStream_t st) {
    using C = ConvCfg<CIN, COUT, H, W, K, PAD, TY, SX, TCO, CC, FLIP, STATS>;
    auto kern = conv_kernel<CIN, COUT, H, W, K, PAD, TY, SX, TCO, CC, FLIP, STATS>;
    static bool attr = false;
    static int ctas_per_sm = 1;
    if (!attr) {
        if (C::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
            B200_REQUIRE(e == cudaSuccess, B200_E_SMEM, "conv: cannot reserve %zu B shared memory", C::SMEM);
        }
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, C::SMEM) == cudaSuccess && occ > 0) ctas_per_sm = occ;
        attr = true;
    }
    // persistent grid: a multiple of the SM count; items are (sample, row-tile) pairs in sample-major order so that a
    // CTA's consecutive items stay inside one view-call for as long as possible (few statistic flushes)
    const long items = (long)N * C::NTILE;
    long grid = (long)sm_count() * ctas_per_sm;
    if (grid > items) grid = items;
    kern<<<(int)grid, C::THREADS, C::SMEM, st>>>(x, w, bias, z, stats, N, n_per_view);
    return launch_status("conv");
}

// =========================================================================================================
// Weight gradient, CIN > 1: output-stationary.  thread = (cout group TCO, ci within the CTA's slice, ky) and keeps
// K*TCO accumulators while sweeping every pixel of the CTA's items; CTAs write partial sums that a second kernel
// adds in a fixed order (deterministic).  grid = (persistent CTAs, CIN/CIB slices).
// =========================================================================================================
template <int CIN, int COUT, int H, int W, int K, int PAD, int TY, int TCO, int CIB>
struct WgCfg {
    static constexpr int HO = H + 2 * PAD - K + 1;
    static constexpr int WO = W + 2 * PAD - K + 1;
    static constexpr int NTILE = (HO + TY - 1) / TY;
    static constexpr int NGROUP = COUT / TCO;
    static constexpr int THREADS = ((NGROUP * CIB * K + 31) / 32) * 32;
    static constexpr int TIH = TY + K - 1;
    static constexpr int TIW = WO + K - 1 + ((WO + K - 1) % 2 == 0 ? 1 : 0);    // odd pitch: ky rows land in different banks
    static constexpr int X_FLOATS = CIB * TIH * TIW + 1;
    static constexpr int DZ_FLOATS = COUT * TY * WO;
    static constexpr size_t SMEM = (size_t)(X_FLOATS + DZ_FLOATS) * sizeof(float);
    static constexpr int NSLICE = CIN / CIB;
};

template <int CIN, int COUT, int H, int W, int K, int PAD, int TY, int TCO, int CIB>
__global__ void __launch_bounds__(WgCfg<CIN, COUT, H, W, K, PAD, TY, TCO, CIB>::THREADS)
conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ part, float* __restrict__ part_db, int N) {
    using C = WgCfg<CIN, COUT, H, W, K, PAD, TY, TCO, CIB>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_x = reinterpret_cast<float*>(smem_raw);
    float* s_dz = s_x + C::X_FLOATS;
    const int tid = threadIdx.x;
    const bool active = tid < C::NGROUP * CIB * K;
    // lanes vary fastest over (ci, ky) so that a warp shares few cout groups (dz reads are broadcasts)
    const int ky = tid % K, cil = (tid / K) % CIB, grp = tid / (K * CIB);
    const int co0 = grp * TCO;
    const int slice = blockIdx.y;
    float acc[K][TCO];
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
        for (int t = 0; t < TCO; ++t) acc[a][t] = 0.f;
    float dbacc[TCO];
#pragma unroll
    for (int t = 0; t < TCO; ++t) dbacc[t] = 0.f;
    const bool do_db = (slice == 0) && (ky == 0) && (cil == 0);

    const int n_items = N * C::NTILE;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int n = item / C::NTILE, tile = item - n * C::NTILE;
        const int y0 = tile * TY;
        __syncthreads();
        const float* xin = x + ((size_t)n * CIN + slice * CIB) * H * W;
        for (int e = tid; e < CIB * C::TIH * C::TIW; e += C::THREADS) {
            const int c = e / (C::TIH * C::TIW), r = (e / C::TIW) % C::TIH, col = e % C::TIW;
            const int gy = y0 - PAD + r, gx = col - PAD;
            float v = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(xin + ((size_t)c * H + gy) * W + gx);
            s_x[e] = v;
        }
        const float* dzin = dz + (size_t)n * COUT * C::HO * C::WO;
        for (int e = tid; e < C::DZ_FLOATS; e += C::THREADS) {
            const int co = e / (TY * C::WO), r = (e / C::WO) % TY, col = e % C::WO;
            const int gy = y0 + r;
            s_dz[e] = (gy < C::HO) ? __ldg(dzin + ((size_t)co * C::HO + gy) * C::WO + col) : 0.f;
        }
        __syncthreads();
        if (active) {
#pragma unroll 1
            for (int r = 0; r < TY; ++r) {
                const float* xr = s_x + (cil * C::TIH + r + ky) * C::TIW;
                float win[K];
#pragma unroll
                for (int a = 0; a < K - 1; ++a) win[a + 1] = xr[a];
#pragma unroll 2
                for (int col = 0; col < C::WO; ++col) {
#pragma unroll
                    for (int a = 0; a < K - 1; ++a) win[a] = win[a + 1];
                    win[K - 1] = xr[col + K - 1];
                    float dv[TCO];
#pragma unroll
                    for (int t = 0; t < TCO; ++t) dv[t] = s_dz[((co0 + t) * TY + r) * C::WO + col];
#pragma unroll
                    for (int a = 0; a < K; ++a)
#pragma unroll
                        for (int t = 0; t < TCO; ++t) acc[a][t] = fmaf(win[a], dv[t], acc[a][t]);
                    if (do_db) {
#pragma unroll
                        for (int t = 0; t < TCO; ++t) dbacc[t] += dv[t];
                    }
                }
            }
        }
    }
    if (active) {
        // partial layout: part[blockIdx.x][co][ci][ky][kx]
        float* pp = part + (size_t)blockIdx.x * COUT * CIN * K * K;
#pragma unroll
        for (int t = 0; t < TCO; ++t)
#pragma unroll
            for (int a = 0; a < K; ++a)
                pp[(((size_t)(co0 + t) * CIN + slice * CIB + cil) * K + ky) * K + a] = acc[a][t];
        if (do_db) {
#pragma unroll
            for (int t = 0; t < TCO; ++t) part_db[(size_t)blockIdx.x * COUT + co0 + t] = dbacc[t];
        }
    }
}

// Weight gradient, CIN == 1: thread = (cout, strip of SX pixels); accumulators are the K*K taps; block reduction.
template <int COUT, int H, int W, int K, int PAD, int TY, int SX>
struct Wg1Cfg {
    static constexpr int HO = H + 2 * PAD - K + 1;
    static constexpr int WO = W + 2 * PAD - K + 1;
    static constexpr int NSTRIP = (WO + SX - 1) / SX;
    static constexpr int NTILE = (HO + TY - 1) / TY;
    static constexpr int NPOS = TY * NSTRIP;
    static constexpr int THREADS = ((NPOS * COUT + 31) / 32) * 32;
    static constexpr int TIH = TY + K - 1;
    static constexpr int TIW = NSTRIP * SX + K - 1;
    static constexpr int X_FLOATS = TIH * TIW;
    static constexpr int RED_FLOATS = COUT * (K * K + 1);
    static constexpr size_t SMEM = (size_t)(X_FLOATS + RED_FLOATS + THREADS) * sizeof(float);      // + one staging float per thread
};

template <int COUT, int H, int W, int K, int PAD, int TY, int SX>
__global__ void __launch_bounds__(Wg1Cfg<COUT, H, W, K, PAD, TY, SX>::THREADS)
conv_wgrad_c1_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ part, float* __restrict__ part_db, int N) {
    using C = Wg1Cfg<COUT, H, W, K, PAD, TY, SX>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_x = reinterpret_cast<float*>(smem_raw);
    float* s_red = s_x + C::X_FLOATS;
    const int tid = threadIdx.x;
    const bool active = tid < C::NPOS * COUT;
    const int pos = tid % C::NPOS, co = tid / C::NPOS;
    const int ly = pos / C::NSTRIP, lx = (pos % C::NSTRIP) * SX;
    float acc[K * K];
#pragma unroll
    for (int a = 0; a < K * K; ++a) acc[a] = 0.f;
    float dbacc = 0.f;
    for (int i = tid; i < C::RED_FLOATS; i += C::THREADS) s_red[i] = 0.f;

    const int n_items = N * C::NTILE;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int n = item / C::NTILE, tile = item - n * C::NTILE;
        const int y0 = tile * TY;
        __syncthreads();
        const float* xin = x + (size_t)n * H * W;
        for (int e = tid; e < C::X_FLOATS; e += C::THREADS) {
            const int r = e / C::TIW, col = e % C::TIW;
            const int gy = y0 - PAD + r, gx = col - PAD;
            s_x[e] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(xin + (size_t)gy * W + gx) : 0.f;
        }
        __syncthreads();
        const int oy = y0 + ly;
        if (active && oy < C::HO) {
            float dv[SX];
            const float* dp = dz + (((size_t)n * COUT + co) * C::HO + oy) * C::WO + lx;
#pragma unroll
            for (int p = 0; p < SX; ++p) {
                dv[p] = (lx + p < C::WO) ? __ldg(dp + p) : 0.f;
                dbacc += dv[p];
            }
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                float in[SX + K - 1];
#pragma unroll
                for (int q = 0; q < SX + K - 1; ++q) in[q] = s_x[(ly + ky) * C::TIW + lx + q];
#pragma unroll
                for (int kx = 0; kx < K; ++kx)
#pragma unroll
                    for (int p = 0; p < SX; ++p) acc[ky * K + kx] = fmaf(in[p + kx], dv[p], acc[ky * K + kx]);
            }
        }
    }
    __syncthreads();
    // block reduction in a FIXED order (run-to-run deterministic, like the reference's deterministic=True): tap by tap every
    // thread stages its accumulator, then one thread per output channel adds the NPOS positions in index order
    float* s_stage = s_red + C::RED_FLOATS;
#pragma unroll
    for (int a = 0; a <= K * K; ++a) {
        s_stage[tid] = active ? (a < K * K ? acc[a < K * K ? a : 0] : dbacc) : 0.f;
        __syncthreads();
        if (tid < COUT) {
            float sum = 0.f;
            for (int p = 0; p < C::NPOS; ++p) sum += s_stage[tid * C::NPOS + p];
            s_red[tid * (K * K + 1) + a] = sum;
        }
        __syncthreads();
    }
    for (int i = tid; i < COUT * K * K; i += C::THREADS) {
        const int c = i / (K * K), a = i % (K * K);
        part[(size_t)blockIdx.x * COUT * K * K + i] = s_red[c * (K * K + 1) + a];
    }
    for (int i = tid; i < COUT; i += C::THREADS) part_db[(size_t)blockIdx.x * COUT + i] = s_red[i * (K * K + 1) + K * K];
}

__global__ void __launch_bounds__(256) reduce_parts_kernel(const float* __restrict__ part, int n_parts, int n, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int p = 0; p < n_parts; ++p) a += part[(size_t)p * n + i];
        out[i] = a;
    }
}

static int wgrad_grid(long items, int slices, int ctas_per_sm) {
    long g = (long)sm_count() * ctas_per_sm / slices;
    if (g < 1) g = 1;
    if (g > items) g = items;
    return (int)g;
}

template <int CIN, int COUT, int H, int W, int K, int PAD, int TY, int TCO, int CIB>
static int launch_wgrad(const float* x, const float* dz, float* dw, float* db, float* work, int N, cudaStream_t st, bool query, int64_t* need) {
    using C = WgCfg<CIN, COUT, H, W, K, PAD, TY, TCO, CIB>;
    auto kern = conv_wgrad_kernel<CIN, COUT, H, W, K, PAD, TY, TCO, CIB>;
    const long items = (long)N * C::NTILE;
    const int gx = wgrad_grid(items, C::NSLICE, 2);
    const int64_t wn = (int64_t)COUT * CIN * K * K;
    if (query) {
        *need = (int64_t)wgrad_grid((long)1 << 40, C::NSLICE, 2) * (wn + COUT);
        return 0;
    }
    static bool attr = false;
    if (!attr) {
        if (C::SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
            B200_REQUIRE(e == cudaSuccess, B200_E_SMEM, "conv wgrad: cannot reserve %zu B shared memory", C::SMEM);
        }
        attr = true;
    }
    float* part = work;
    float* part_db = work + (size_t)gx * wn;
    kern<<<dim3(gx, C::NSLICE), C::THREADS, C::SMEM, st>>>(x, dz, part, part_db, N);
    int rc = launch_status("conv_wgrad");
    if (rc) return rc;
    reduce_parts_kernel<<<(int)((wn + 255) / 256), 256, 0, st>>>(part, gx, (int)wn, dw);
    if (db) reduce_parts_kernel<<<1, 256, 0, st>>>(part_db, gx, COUT, db);
    return launch_status("conv_wgrad_reduce");
}

template <int COUT, int H, int W, int K, int PAD, int TY, int SX>
static int launch_wgrad_c1(const float* x, const float* dz, float* dw, float* db, float* work, int N, cudaStream_t st, bool query, int64_t* need) {
    using C = Wg1Cfg<COUT, H, W, K, PAD, TY, SX>;
    auto kern = conv_wgrad_c1_kernel<COUT, H, W, K, PAD, TY, SX>;
    const long items = (long)N * C::NTILE;
    const int gx = wgrad_grid(items, 1, 2);
    const int64_t wn = (int64_t)COUT * K * K;
    if (query) {
        *need = (int64_t)wgrad_grid((long)1 << 40, 1, 2) * (wn + COUT);
        return 0;
    }
    float* part = work;
    float* part_db = work + (size_t)gx * wn;
    kern<<<gx, C::THREADS, C::SMEM, st>>>(x, dz, part, part_db, N);
    int rc = launch_status("conv_wgrad_c1");
    if (rc) return rc;
    reduce_parts_kernel<<<(int)((wn + 255) / 256), 256, 0, st>>>(part, gx, (int)wn, dw);
    if (db) reduce_parts_kernel<<<1, 256, 0, st>>>(part_db, gx, COUT, db);
    return launch_status("conv_wgrad_c1_reduce");
}

// =========================================================================================================
// BatchNorm finalisation: one thread per channel walks the view-calls in order (running statistics are updated
// sequentially, exactly like 6 consecutive nn.BatchNorm calls).
// =========================================================================================================
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ rmean, float* __restrict__ rvar, int64_t* __restrict__ nbt,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, int n_views, int C, double count, float momentum, float eps, int train) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float g = gamma[c], b = beta[c];
    if (!train) {
        const float is = rsqrtf(rvar[c] + eps);
        for (int v = 0; v < n_views; ++v) {
            scale[v * C + c] = g * is;
            shift[v * C + c] = b - rmean[c] * g * is;
            if (mean_out) mean_out[v * C + c] = rmean[c];
            if (invstd_out) invstd_out[v * C + c] = is;
        }
        return;
    }
    float rm = rmean[c], rv = rvar[c];
    for (int v = 0; v < n_views; ++v) {
        const double s = stats[((size_t)v * C + c) * 2], ss = stats[((size_t)v * C + c) * 2 + 1];
        const double m = s / count;
        double var = ss / count - m * m;
        var = var < 0.0 ? 0.0 : var;
        const float is = (float)(1.0 / sqrt(var + (double)eps));
        const float mf = (float)m;
        scale[v * C + c] = g * is;
        shift[v * C + c] = b - mf * g * is;
        mean_out[v * C + c] = mf;
        invstd_out[v * C + c] = is;
        const float unbiased = (float)(var * (count / (count > 1.0 ? count - 1.0 : 1.0)));
        rm = (1.f - momentum) * rm + momentum * mf;
        rv = (1.f - momentum) * rv + momentum * unbiased;
    }
    rmean[c] = rm;
    rvar[c] = rv;
    if (c == 0 && nbt) nbt[0] += n_views;
}

// =========================================================================================================
// BN-apply + ReLU + 2x2 max-pool.  One warp per (sample, channel) plane (grid-stride), float2 loads of the two
// source rows.  Pure HBM streaming: reads H*W*4 B, writes H*W B per plane.
// =========================================================================================================
__device__ __forceinline__ void load_win(const float* __restrict__ zp, int W, int py, int px, float (&zv)[4]);
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_kernel(const float* __restrict__ z, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, float* __restrict__ out, int N,
                                                               int n_per_view, int C, int H, int W) {
    const int lane = threadIdx.x & 31;
    const int HP = H >> 1, WP = W >> 1;
    const long planes = (long)N * C;
    for (long pl = (long)blockIdx.x * 8 + (threadIdx.x >> 5); pl < planes; pl += (long)gridDim.x * 8) {
        const int n = (int)(pl / C), c = (int)(pl - (long)n * C);
        const int v = n / n_per_view;
        const float a = __ldg(scale + v * C + c), b = __ldg(shift + v * C + c);
        const float* zp = z + pl * H * W;
        float* op = out + pl * HP * WP;
        for (int e = lane; e < HP * WP; e += 32) {
            const int py = e / WP, px = e - py * WP;
            float zv[4];
            load_win(zp, W, py, px, zv);
            const float m = fmaxf(fmaxf(fmaf(a, zv[0], b), fmaf(a, zv[1], b)), fmaxf(fmaf(a, zv[2], b), fmaf(a, zv[3], b)));
            op[e] = fmaxf(m, 0.f);
        }
    }
}

// 2x2 window load: float2 when the row pitch is even (8-byte aligned), scalar otherwise (7x7 planes)
__device__ __forceinline__ void load_win(const float* __restrict__ zp, int W, int py, int px, float (&zv)[4]) {
    const float* r0 = zp + (2 * py) * W + 2 * px;
    if ((W & 1) == 0) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(r0));
        const float2 b = __ldg(reinterpret_cast<const float2*>(r0 + W));
        zv[0] = a.x; zv[1] = a.y; zv[2] = b.x; zv[3] = b.y;
    } else {
        zv[0] = __ldg(r0); zv[1] = __ldg(r0 + 1); zv[2] = __ldg(r0 + W); zv[3] = __ldg(r0 + W + 1);
    }
}

// argmax in PyTorch scan order (first maximum wins); returns the position 0..3 and the maximum
__device__ __forceinline__ int pool_argmax(float y0, float y1, float y2, float y3, float& m) {
    int k = 0;
    m = y0;
    if (y1 > m) { m = y1; k = 1; }
    if (y2 > m) { m = y2; k = 2; }
    if (y3 > m) { m = y3; k = 3; }
    return k;
}

template <bool APPLY>
__global__ void __launch_bounds__(256) bn_relu_pool_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dout,
                                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                                               const float* __restrict__ mean, const float* __restrict__ invstd,
                                                               double* __restrict__ sums, float* __restrict__ dz, int N, int n_per_view,
                                                               int C, int H, int W, int n_views) {
    extern __shared__ double s_sums[];     // [n_views][C][2]  (reduce pass only)
    const int lane = threadIdx.x & 31;
    const int HP = H >> 1, WP = W >> 1;
    const long planes = (long)N * C;
    const float inv_cnt = 1.0f / ((float)n_per_view * (float)H * (float)W);
    if (!APPLY) {
        for (int i = threadIdx.x; i < n_views * C * 2; i += blockDim.x) s_sums[i] = 0.0;
        __syncthreads();
    }
    for (long pl = (long)blockIdx.x * 8 + (threadIdx.x >> 5); pl < planes; pl += (long)gridDim.x * 8) {
        const int n = (int)(pl / C), c = (int)(pl - (long)n * C);
        const int v = n / n_per_view;
        const float a = __ldg(scale + v * C + c), b = __ldg(shift + v * C + c);
        const float mu = __ldg(mean + v * C + c), is = __ldg(invstd + v * C + c);
        const float* zp = z + pl * H * W;
        const float* dp = dout + pl * HP * WP;
        float k1 = 0.f, k2 = 0.f;
        if (APPLY) {
            k1 = (float)(sums[((size_t)v * C + c) * 2] ) * inv_cnt;
            k2 = (float)(sums[((size_t)v * C + c) * 2 + 1]) * inv_cnt;
        }
        float s1 = 0.f, s2 = 0.f;
        for (int e = lane; e < HP * WP; e += 32) {
            const int py = e / WP, px = e - py * WP;
            float zv[4];
            load_win(zp, W, py, px, zv);
            float m;
            const int k = pool_argmax(fmaf(a, zv[0], b), fmaf(a, zv[1], b), fmaf(a, zv[2], b), fmaf(a, zv[3], b), m);
            const float g = (m > 0.f) ? __ldg(dp + e) : 0.f;
            if (!APPLY) {
                s1 += g;
                s2 += g * ((zv[k] - mu) * is);
            } else {
                float o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float xh = (zv[q] - mu) * is;
                    const float dy = (q == k) ? g : 0.f;
                    o[q] = a * (dy - k1 - xh * k2);
                }
                float* zo = dz + pl * H * W + (2 * py) * W + 2 * px;
                if ((W & 1) == 0) {
                    *reinterpret_cast<float2*>(zo) = make_float2(o[0], o[1]);
                    *reinterpret_cast<float2*>(zo + W) = make_float2(o[2], o[3]);
                } else {
                    zo[0] = o[0]; zo[1] = o[1]; zo[W] = o[2]; zo[W + 1] = o[3];
                }
            }
        }
        if (APPLY && ((H | W) & 1)) {
            // odd planes (7x7 -> floor-pooled 3x3): the last row / column is in no pooling window but still
            // receives the dense BatchNorm terms
            float* zo = dz + pl * H * W;
            for (int e = lane; e < H * W; e += 32) {
                const int y = e / W, x = e - y * W;
                if (y >= 2 * HP || x >= 2 * WP) zo[e] = a * (0.f - k1 - ((__ldg(zp + e) - mu) * is) * k2);
            }
        }
        if (!APPLY) {
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) {
                atomicAdd(&s_sums[((size_t)v * C + c) * 2], (double)s1);
                atomicAdd(&s_sums[((size_t)v * C + c) * 2 + 1], (double)s2);
            }
        }
    }
    if (!APPLY) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_views * C * 2; i += blockDim.x)
            if (s_sums[i] != 0.0) atomicAdd(&sums[i], s_sums[i]);
    }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta, int n_views,
                                      int C, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double g = 0.0, b = 0.0;
    for (int v = 0; v < n_views; ++v) {
        b += sums[((size_t)v * C + c) * 2];
        g += sums[((size_t)v * C + c) * 2 + 1];
    }
    if (accumulate) {
        dgamma[c] += (float)g;
        dbeta[c] += (float)b;
    } else {
        dgamma[c] = (float)g;
        dbeta[c] = (float)b;
    }
}

__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, long planes, int HW) {
    const float inv = 1.0f / (float)HW;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < planes; p += (long)gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int i = 0; i < HW; ++i) a += __ldg(x + p * HW + i);
        out[p] = a * inv;
    }
}
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, long planes, int HW) {
    const float inv = 1.0f / (float)HW;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < planes * HW; e += (long)gridDim.x * blockDim.x)
        dx[e] = __ldg(dout + e / HW) * inv;
}

static int stream_grid(long warps_needed) {
    long g = (warps_needed + 7) / 8;
    long cap = (long)sm_count() * 8;
    if (g < 1) g = 1;
    return (int)(g < cap ? g : cap);
}

// ---- shape dispatch ---------------------------------------------------------------------------------------
// key = (Cin, Cout, H, W, K, pad)
#define CONV_SHAPES(X)                                                                 \
    /*      CIN COUT  H    W   K PAD  TY SX TCO CC */                                 \
    X(1, 32, 28, 28, 5, 2, 7, 4, 8, 1)        /* image conv1 (LeNet)            */     \
    X(32, 64, 14, 14, 5, 0, 10, 5, 8, 8)      /* image conv2                    */     \
    X(1, 8, 112, 112, 5, 2, 8, 4, 8, 1)       /* audio conv1                    */     \
    X(8, 16, 56, 56, 5, 2, 8, 4, 8, 8)        /* audio conv2                    */     \
    X(16, 32, 28, 28, 5, 2, 7, 4, 8, 16)      /* audio conv3                    */     \
    X(32, 64, 14, 14, 5, 2, 14, 7, 8, 8)      /* audio conv4                    */     \
    X(1, 32, 28, 28, 3, 1, 7, 4, 8, 1)        /* image_simple conv1             */     \
    X(32, 64, 14, 14, 3, 1, 14, 7, 8, 16)     /* image_simple conv2             */     \
    X(64, 128, 7, 7, 3, 1, 7, 7, 8, 16)       /* image_simple conv3             */     \
    X(1, 32, 112, 112, 3, 1, 4, 4, 8, 1)      /* multi_simple audio conv1 (models/dino.py:43-72) */ \
    X(32, 64, 56, 56, 3, 1, 8, 4, 8, 8)       /* multi_simple audio conv2       */     \
    X(64, 128, 28, 28, 3, 1, 7, 4, 8, 8)      /* multi_simple audio conv3       */     \
    X(128, 256, 14, 14, 3, 1, 14, 7, 8, 8)    /* multi_simple audio conv4       */

// data-gradient instances: (CIN' = Cout, COUT' = Cin, H' = HO, W' = WO, K, PAD' = K-1-pad)
#define DGRAD_SHAPES(X)                                                                \
    /* fwd: Cin Cout H  W  K pad |  TY SX TCO CC */                                    \
    X(32, 64, 14, 14, 5, 0, 14, 7, 8, 8)                                               \
    X(8, 16, 56, 56, 5, 2, 14, 4, 8, 16)                                               \
    X(16, 32, 28, 28, 5, 2, 14, 4, 8, 16)                                              \
    X(32, 64, 14, 14, 5, 2, 14, 7, 8, 16)                                              \
    X(32, 64, 14, 14, 3, 1, 14, 7, 8, 16)                                              \
    X(64, 128, 7, 7, 3, 1, 7, 7, 8, 16)                                                \
    X(32, 64, 56, 56, 3, 1, 14, 4, 8, 16)                                              \
    X(64, 128, 28, 28, 3, 1, 14, 4, 8, 16)                                             \
    X(128, 256, 14, 14, 3, 1, 14, 7, 8, 16)

#define WGRAD_SHAPES(X)                                                                \
    /*      CIN COUT  H   W  K PAD  TY TCO CIB */                                      \
    X(32, 64, 14, 14, 5, 0, 10, 8, 8)                                                  \
    X(8, 16, 56, 56, 5, 2, 8, 4, 8)                                                    \
    X(16, 32, 28, 28, 5, 2, 14, 8, 16)                                                 \
    X(32, 64, 14, 14, 5, 2, 14, 8, 8)                                                  \
    X(32, 64, 14, 14, 3, 1, 14, 8, 8)                                                  \
    X(64, 128, 7, 7, 3, 1, 7, 8, 8)                                                    \
    X(32, 64, 56, 56, 3, 1, 4, 8, 8)                                                   \
    X(64, 128, 28, 28, 3, 1, 4, 8, 8)                                                  \
    X(128, 256, 14, 14, 3, 1, 7, 8, 8)

#define WGRAD1_SHAPES(X)                                                               \
    /*     COUT  H    W   K PAD  TY SX */                                              \
    X(32, 28, 28, 5, 2, 4, 4)                                                          \
    X(8, 112, 112, 5, 2, 4, 4)                                                         \
    X(32, 28, 28, 3, 1, 4, 4)                                                          \
    X(32, 112, 112, 3, 1, 1, 4)

}  // namespace b200

using namespace b200;

extern "C" {

int b200_conv_supported(int Cin, int Cout, int H, int W, int K, int pad) {
#define X(CI, CO, HH, WW, KK, PP, TY, SX, TCO, CC) \
    if (Cin == CI && Cout == CO && H == HH && W == WW && K == KK && pad == PP) return 1;
    CONV_SHAPES(X)
#undef X
    return 0;
}

int b200_conv_fwd(const float* x, const float* w, const float* bias, float* z, double* stats, int N, int n_per_view, int Cin,
                  int Cout, int H, int W, int K, int pad, void* stream) {
    B200_REQUIRE(x && w && z && N > 0 && n_per_view > 0 && N % n_per_view == 0, B200_E_ARG, "conv_fwd: bad arguments");
    cudaStream_t st = as_stream(stream);
#define X(CI, CO, HH, WW, KK, PP, TY, SX, TCO, CC)                                                                  \
    if (Cin == CI && Cout == CO && H == HH && W == WW && K == KK && pad == PP) {                                     \
        if (stats) return launch_conv<CI, CO, HH, WW, KK, PP, TY, SX, TCO, CC, false, true>(x, w, bias, z, stats, N, n_per_view, st); \
        return launch_conv<CI, CO, HH, WW, KK, PP, TY, SX, TCO, CC, false, false>(x, w, bias, z, nullptr, N, n_per_view, st);          \
    }
    CONV_SHAPES(X)
#undef X
    B200_REQUIRE(false, B200_E_SHAPE, "conv_fwd: shape (%d,%d,%d,%d,%d,%d) not compiled", Cin, Cout, H, W, K, pad);
}

int b200_conv_bwd_data(const float* dz, const float* w, float* dx, int N, int Cin, int Cout, int H, int W, int K, int pad,
                       void* stream) {
    B200_REQUIRE(dz && w && dx && N > 0, B200_E_ARG, "conv_bwd_data: bad arguments");
    cudaStream_t st = as_stream(stream);
#define X(CI, CO, HH, WW, KK, PP, TY, SX, TCO, CC)                                                                   \
    if (Cin == CI && Cout == CO && H == HH && W == WW && K == KK && pad == PP)                                        \
        return launch_conv<CO, CI, HH + 2 * PP - KK + 1, WW + 2 * PP - KK + 1, KK, KK - 1 - PP, TY, SX, TCO, CC, true, false>( \
            dz, w, nullptr, dx, nullptr, N, N, st);
    DGRAD_SHAPES(X)
#undef X
    B200_REQUIRE(false, B200_E_SHAPE, "conv_bwd_data: shape (%d,%d,%d,%d,%d,%d) not compiled", Cin, Cout, H, W, K, pad);
}

static int wgrad_dispatch(const float* x, const float* dz, float* dw, float* db, float* work, int N, int Cin, int Cout, int H,
                          int W, int K, int pad, cudaStream_t st, bool query, int64_t* need) {
#define X(CI, CO, HH, WW, KK, PP, TY, TCO, CIB) \
    if (Cin == CI && Cout == CO && H == HH && W == WW && K == KK && pad == PP) \
        return launch_wgrad<CI, CO, HH, WW, KK, PP, TY, TCO, CIB>(x, dz, dw, db, work, N, st, query, need);
    WGRAD_SHAPES(X)
#undef X
#define X(CO, HH, WW, KK, PP, TY, SX) \
    if (Cin == 1 && Cout == CO && H == HH && W == WW && K == KK && pad == PP) \
        return launch_wgrad_c1<CO, HH, WW, KK, PP, TY, SX>(x, dz, dw, db, work, N, st, query, need);
    WGRAD1_SHAPES(X)
#undef X
    B200_REQUIRE(false, B200_E_SHAPE, "conv_bwd_weight: shape (%d,%d,%d,%d,%d,%d) not compiled", Cin, Cout, H, W, K, pad);
}

int64_t b200_conv_bwd_weight_work_floats(int N, int Cin, int Cout, int H, int W, int K, int pad) {
    int64_t need = 0;
    int rc = wgrad_dispatch(nullptr, nullptr, nullptr, nullptr, nullptr, N, Cin, Cout, H, W, K, pad, nullptr, true, &need);
    return rc ? (int64_t)rc : need;
}

int b200_conv_bwd_weight(const float* x, const float* dz, float* dw, float* db, float* work, int N, int Cin, int Cout, int H,
                         int W, int K, int pad, void* stream) {
    B200_REQUIRE(x && dz && dw && work && N > 0, B200_E_ARG, "conv_bwd_weight: bad arguments");
    int64_t need = 0;
    return wgrad_dispatch(x, dz, dw, db, work, N, Cin, Cout, H, W, K, pad, as_stream(stream), false, &need);
}

int b200_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd, int n_views, int C,
                     int64_t count, float momentum, float eps, int train, void* stream) {
    B200_REQUIRE(gamma && beta && running_mean && running_var && scale && shift && n_views > 0 && C > 0, B200_E_ARG,
                 "bn_finalize: bad arguments");
    B200_REQUIRE(!train || (stats && mean && invstd && count > 0), B200_E_ARG, "bn_finalize: train mode needs stats/mean/invstd");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(stats, gamma, beta, running_mean, running_var,
                                                                       num_batches_tracked, scale, shift, mean, invstd, n_views, C,
                                                                       (double)count, momentum, eps, train);
    return launch_status("bn_finalize");
}

int b200_bn_relu_pool_fwd(const float* z, const float* scale, const float* shift, float* out, int N, int n_per_view, int C, int H,
                          int W, void* stream) {
    B200_REQUIRE(z && scale && shift && out && N > 0 && n_per_view > 0 && C > 0, B200_E_ARG, "bn_relu_pool_fwd: bad arguments");
    bn_relu_pool_fwd_kernel<<<stream_grid((long)N * C), 256, 0, as_stream(stream)>>>(z, scale, shift, out, N, n_per_view, C, H, W);
    return launch_status("bn_relu_pool_fwd");
}

int b200_bn_relu_pool_bwd_reduce(const float* z, const float* dout, const float* scale, const float* shift, const float* mean,
                                 const float* invstd, double* sums, int N, int n_per_view, int C, int H, int W, void* stream) {
    B200_REQUIRE(z && dout && scale && shift && mean && invstd && sums && N > 0 && n_per_view > 0 && N % n_per_view == 0, B200_E_ARG,
                 "bn_relu_pool_bwd_reduce: bad arguments");
    const int n_views = N / n_per_view;
    const size_t smem = (size_t)n_views * C * 2 * sizeof(double);
    B200_REQUIRE(smem <= 48 * 1024, B200_E_SHAPE, "bn_relu_pool_bwd_reduce: n_views*C too large");
    bn_relu_pool_bwd_kernel<false><<<stream_grid((long)N * C), 256, smem, as_stream(stream)>>>(z, dout, scale, shift, mean, invstd, sums,
                                                                                             nullptr, N, n_per_view, C, H, W, n_views);
    return launch_status("bn_relu_pool_bwd_reduce");
}

int b200_bn_relu_pool_bwd_apply(const float* z, const float* dout, const float* scale, const float* shift, const float* mean,
                                const float* invstd, const double* sums, float* dz, int N, int n_per_view, int C, int H, int W,
                                void* stream) {
    B200_REQUIRE(z && dout && scale && shift && mean && invstd && sums && dz && N > 0 && n_per_view > 0, B200_E_ARG,
                 "bn_relu_pool_bwd_apply: bad arguments");
    bn_relu_pool_bwd_kernel<true><<<stream_grid((long)N * C), 256, 0, as_stream(stream)>>>(
        z, dout, scale, shift, mean, invstd, const_cast<double*>(sums), dz, N, n_per_view, C, H, W, N / n_per_view);
    return launch_status("bn_relu_pool_bwd_apply");
}

int b200_bn_param_grads(const double* sums, float* dgamma, float* dbeta, int n_views, int C, int accumulate, void* stream) {
    B200_REQUIRE(sums && dgamma && dbeta && n_views > 0 && C > 0, B200_E_ARG, "bn_param_grads: bad arguments");
    bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(sums, dgamma, dbeta, n_views, C, accumulate);
    return launch_status("bn_param_grads");
}

int b200_avgpool_fwd(const float* x, float* out, int N, int C, int HW, void* stream) {
    B200_REQUIRE(x && out && N > 0 && C > 0 && HW > 0, B200_E_ARG, "avgpool_fwd: bad arguments");
    const long planes = (long)N * C;
    avgpool_fwd_kernel<<<stream_grid((planes + 31) / 32), 256, 0, as_stream(stream)>>>(x, out, planes, HW);
    return launch_status("avgpool_fwd");
}

int b200_avgpool_bwd(const float* dout, float* dx, int N, int C, int HW, void* stream) {
    B200_REQUIRE(dout && dx && N > 0 && C > 0 && HW > 0, B200_E_ARG, "avgpool_bwd: bad arguments");
    const long planes = (long)N * C;
    avgpool_bwd_kernel<<<stream_grid((planes * HW + 31) / 32), 256, 0, as_stream(stream)>>>(dout, dx, planes, HW);
    return launch_status("avgpool_bwd");
}

}  // extern "C"
