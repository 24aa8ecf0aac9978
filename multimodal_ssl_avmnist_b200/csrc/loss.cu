// Fused loss kernels: DINO centre-softmax-cross-entropy (fwd + bwd + teacher column sums), centre EMA,
// MSE alignment, 10-way CE, InfoNCE (tiled, the [B,B] matrix never reaches HBM), cosine consistency.
// All reductions inside a row are warp shuffles; cross-CTA reductions go through per-CTA partials that are
// summed in a fixed order (deterministic).
#include <cuda_bf16.h>

#include "common.cuh"

namespace b200 {

constexpr float NORM_EPS = 1e-12f;   // F.normalize eps
constexpr int LOSS_WARPS = 8;

// ---------------------------------------------------------------------------------------------------------
// DINO loss.  One warp per sample (grid-stride); lane owns columns lane + 32*j.
// Algorithmic HBM bytes per sample: (Vs+Vt)*D*4 read + Vs*D*4 written  (7168 B at Vs=6, Vt=2, D=128).
// ---------------------------------------------------------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(LOSS_WARPS * 32)
dino_loss_kernel(const float* __restrict__ s, const float* __restrict__ t, const float* __restrict__ center,
                 const float* __restrict__ t_colmean, int Vs, int Vt, int B, float tau_s, float tau_t, float gscale,
                 int variant, float* __restrict__ grad_s, float* __restrict__ part_loss, float* __restrict__ part_colsum) {
    constexpr int D = NPL * 32;
    __shared__ float red[LOSS_WARPS][D];
    __shared__ float red_loss[LOSS_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float c = 1.0f / ((float)Vs * (float)Vt * (float)B);
    float colsum[NPL], cen[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        colsum[j] = 0.f;
        cen[j] = __ldg(center + lane + 32 * j);
    }
    float loss_acc = 0.f;
    for (int b = blockIdx.x * LOSS_WARPS + warp; b < B; b += gridDim.x * LOSS_WARPS) {
        float pbar[NPL];
#pragma unroll
        for (int j = 0; j < NPL; ++j) pbar[j] = 0.f;
        for (int u = 0; u < Vt; ++u) {
            const float* tp = t + ((size_t)u * B + b) * D;
            float x[NPL];
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                float v = __ldg(tp + lane + 32 * j);
                colsum[j] += v;                    // centre EMA uses the UNcentred projections (dino.py:717)
                v -= cen[j];
                x[j] = v;
                ss += v * v;
            }
            float denom = fmaxf(sqrtf(warp_sum(ss)), NORM_EPS);
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                float v = x[j] / denom;
                if (variant == 1) v -= __ldg(t_colmean + (size_t)u * D + lane + 32 * j);
                v = v / tau_t;
                x[j] = v;
                mx = fmaxf(mx, v);
            }
            mx = warp_max(mx);
            float se = 0.f;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                x[j] = expf(x[j] - mx);
                se += x[j];
            }
            se = warp_sum(se);
#pragma unroll
            for (int j = 0; j < NPL; ++j) pbar[j] += x[j] / se;
        }
        for (int v = 0; v < Vs; ++v) {
            const float* sp = s + ((size_t)v * B + b) * D;
            float sh[NPL], z[NPL];
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                sh[j] = __ldg(sp + lane + 32 * j);
                ss += sh[j] * sh[j];
            }
            float denom = fmaxf(sqrtf(warp_sum(ss)), NORM_EPS);
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                sh[j] = sh[j] / denom;
                z[j] = sh[j] / tau_s;
                mx = fmaxf(mx, z[j]);
            }
            mx = warp_max(mx);
            float se = 0.f;
#pragma unroll
            for (int j = 0; j < NPL; ++j) se += expf(z[j] - mx);
            float lse = mx + logf(warp_sum(se));
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                float q = z[j] - lse;
                loss_acc += pbar[j] * q;
                float dz = ((float)Vt * expf(q) - pbar[j]) * c * gscale;
                z[j] = dz / tau_s;               // d loss / d s_hat
                dot += sh[j] * z[j];
            }
            dot = warp_sum(dot);
            float* gp = grad_s + ((size_t)v * B + b) * D;
#pragma unroll
            for (int j = 0; j < NPL; ++j) gp[lane + 32 * j] = (z[j] - sh[j] * dot) / denom;
        }
    }
    loss_acc = warp_sum(loss_acc);
#pragma unroll
    for (int j = 0; j < NPL; ++j) red[warp][lane + 32 * j] = colsum[j];
    if (lane == 0) red_loss[warp] = loss_acc;
    __syncthreads();
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < LOSS_WARPS; ++w) a += red[w][k];
        part_colsum[(size_t)blockIdx.x * D + k] = a;
    }
    if (threadIdx.x == 0) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < LOSS_WARPS; ++w) a += red_loss[w];
        part_loss[blockIdx.x] = -c * a;
    }
}

// column mean over b of normalize(t - center): one CTA per (view, 32-column slab); warps stride over rows.
__global__ void __launch_bounds__(256) teacher_norm_colmean_kernel(const float* __restrict__ t, const float* __restrict__ center,
                                                                   int B, int D, float* __restrict__ out) {
    // grid.x = Vt.  Each warp handles rows b = warp, warp+8, ...; lane owns columns lane+32j (loop over j).
    extern __shared__ float sm[];        // [8][D]
    const int u = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* acc = sm + warp * D;
    for (int k = lane; k < D; k += 32) acc[k] = 0.f;
    for (int b = warp; b < B; b += 8) {
        const float* tp = t + ((size_t)u * B + b) * D;
        float ss = 0.f;
        for (int k = lane; k < D; k += 32) {
            float v = __ldg(tp + k) - __ldg(center + k);
            ss += v * v;
        }
        float denom = fmaxf(sqrtf(warp_sum(ss)), NORM_EPS);
        for (int k = lane; k < D; k += 32) acc[k] += (__ldg(tp + k) - __ldg(center + k)) / denom;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += sm[w * D + k];
        out[(size_t)u * D + k] = a / (float)B;
    }
}

__global__ void center_update_kernel(float* __restrict__ center, const float* __restrict__ part_colsum,
                                     const float* __restrict__ part_loss, int n_parts, int D, float inv_rows, float mc,
                                     float omc, float* __restrict__ loss_out, float* __restrict__ colsum_out) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < D; k += gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int p = 0; p < n_parts; ++p) a += part_colsum[(size_t)p * D + k];
        if (colsum_out) colsum_out[k] = a;
        else center[k] = __fadd_rn(__fmul_rn(center[k], mc), __fmul_rn(__fmul_rn(a, inv_rows), omc));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && loss_out && part_loss) {
        float a = 0.f;
        for (int p = 0; p < n_parts; ++p) a += part_loss[p];
        loss_out[0] = a;
    }
}

// ---------------------------------------------------------------------------------------------------------
// MSE alignment: mean((normalize(a) - normalize(b))^2); one warp per row.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mse_align_kernel(const float* __restrict__ a, const float* __restrict__ b, int B, int D,
                                                        float gscale, float* __restrict__ ga, float* __restrict__ gb,
                                                        float* __restrict__ loss_out, float* __restrict__ work) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_n = 1.0f / ((float)B * (float)D);
    float lacc = 0.f;
    for (int r = blockIdx.x * 8 + warp; r < B; r += gridDim.x * 8) {
        const float* ap = a + (size_t)r * D;
        const float* bp = b + (size_t)r * D;
        float sa = 0.f, sb = 0.f;
        for (int k = lane; k < D; k += 32) {
            float x = __ldg(ap + k), y = __ldg(bp + k);
            sa += x * x;
            sb += y * y;
        }
        float da = fmaxf(sqrtf(warp_sum(sa)), NORM_EPS), db = fmaxf(sqrtf(warp_sum(sb)), NORM_EPS);
        float dota = 0.f, dotb = 0.f;
        for (int k = lane; k < D; k += 32) {
            float x = __ldg(ap + k) / da, y = __ldg(bp + k) / db;
            float d = x - y;
            lacc += d * d;
            float g = 2.f * d * inv_n * gscale;     // d loss / d a_hat ;  d loss / d b_hat = -g
            dota += x * g;
            dotb += y * (-g);
        }
        dota = warp_sum(dota);
        dotb = warp_sum(dotb);
        for (int k = lane; k < D; k += 32) {
            float x = __ldg(ap + k) / da, y = __ldg(bp + k) / db;
            float g = 2.f * (x - y) * inv_n * gscale;
            ga[(size_t)r * D + k] = (g - x * dota) / da;
            gb[(size_t)r * D + k] = (-g - y * dotb) / db;
        }
    }
    grid_sum_ordered(block_sum_ordered(warp_sum(lacc) * inv_n), work, loss_out);
}

// 10-way (C <= 32) cross entropy, mean over rows; one thread per row.
__global__ void __launch_bounds__(128) ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int B, int C,
                                                 float gscale, float* __restrict__ g, float* __restrict__ loss_out,
                                                 float* __restrict__ work) {
    float lacc = 0.f;
    const float invB = 1.0f / (float)B;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x) {
        const float* lp = logits + (size_t)r * C;
        float mx = -INFINITY;
        for (int k = 0; k < C; ++k) mx = fmaxf(mx, __ldg(lp + k));
        float se = 0.f;
        for (int k = 0; k < C; ++k) se += expf(__ldg(lp + k) - mx);
        float lse = mx + logf(se);
        int y = (int)labels[r];
        lacc += lse - __ldg(lp + y);
        for (int k = 0; k < C; ++k) g[(size_t)r * C + k] = (expf(__ldg(lp + k) - lse) - (k == y ? 1.f : 0.f)) * invB * gscale;
    }
    grid_sum_ordered(block_sum_ordered(warp_sum(lacc) * invB), work, loss_out);
}

// ---------------------------------------------------------------------------------------------------------
// InfoNCE.  sim_ij = <a_hat_i, b_hat_j>/temp is bounded by 1/temp, so exp(sim - 1/temp) never overflows and plain
// sums replace the online max.  rows_kernel<false>: rowsum_i = sum_j exp(sim_ij - 1/temp) for a 64-row block against
// all columns; called twice (a,b) and (b,a) -> row and column sums.  rows_kernel<true>: recomputes the tile,
// forms G_ij = (e_ij/rowsum_i + e_ij/colsum_j) * 0.5/B - delta_ij/B and accumulates dA_hat = G B_hat / temp for the
// 64-row block (no atomics; called twice with swapped roles).  FP32 CUDA-core tiles in this revision.
// ---------------------------------------------------------------------------------------------------------
constexpr int NT = 64;   // tile edge
template <bool GRAD>
__global__ void __launch_bounds__(256) infonce_rows_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int B, int D,
                                                           float inv_temp, const float* __restrict__ rowsum_in,
                                                           const float* __restrict__ colsum_in, float coef,
                                                           float* __restrict__ out, int partner_half) {
    // partner_half > 0: NT-Xent on ONE set of 2*half rows (A == Bm): self-similarities are masked out of the sums and the
    // positive of row i is row (i + half) mod B instead of the diagonal
    // smem: At[D][NT] (k-major), Bt[D][NT], Bs[NT][D] (row-major copy for the G*B product), G[NT][NT+1]
    extern __shared__ float sm[];
    float* At = sm;
    float* Bt = At + (size_t)D * NT;
    float* Bs = Bt + (size_t)D * NT;
    float* G = Bs + (GRAD ? (size_t)NT * D : 0);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.x * NT;
    for (int e = tid; e < NT * D; e += 256) {
        int r = e / D, k = e - r * D;
        At[k * NT + r] = (i0 + r < B) ? __ldg(A + (size_t)(i0 + r) * D + k) : 0.f;
    }
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    // gradient accumulators: thread owns rows 4*(tid/16).. +3? -> use mapping r4 = ty (4 rows), columns tx + 16*c
    float acc[4][16];
    if (GRAD) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[a][c] = 0.f;
    }
    float rsi[4] = {1.f, 1.f, 1.f, 1.f};
    if (GRAD) {
#pragma unroll
        for (int a = 0; a < 4; ++a) rsi[a] = (i0 + ty * 4 + a < B) ? __ldg(rowsum_in + i0 + ty * 4 + a) : 1.f;
    }
    for (int j0 = 0; j0 < B; j0 += NT) {
        __syncthreads();
        for (int e = tid; e < NT * D; e += 256) {
            int r = e / D, k = e - r * D;
            float v = (j0 + r < B) ? __ldg(Bm + (size_t)(j0 + r) * D + k) : 0.f;
            Bt[k * NT + r] = v;
            if (GRAD) Bs[r * D + k] = v;
        }
        __syncthreads();
        float sacc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) sacc[a][c] = 0.f;
        for (int k = 0; k < D; ++k) {
            float4 av = *reinterpret_cast<const float4*>(At + k * NT + ty * 4);
            float4 bv = *reinterpret_cast<const float4*>(Bt + k * NT + tx * 4);
            float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) sacc[a][c] = fmaf(aa[a], bb[c], sacc[a][c]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int gi = i0 + ty * 4 + a, gj = j0 + tx * 4 + c;
                float e = (gi < B && gj < B && !(partner_half > 0 && gi == gj)) ? expf(sacc[a][c] * inv_temp - inv_temp) : 0.f;
                if (!GRAD) {
                    rs[a] += e;
                } else {
                    float csj = (gj < B) ? __ldg(colsum_in + gj) : 1.f;
                    const int pos = partner_half > 0 ? (gi + partner_half) % B : gi;
                    float g = (e / rsi[a] + e / csj) * (0.5f * coef) - ((gj == pos && gi < B) ? coef : 0.f);
                    G[(ty * 4 + a) * (NT + 1) + tx * 4 + c] = g;
                }
            }
        }
        if (GRAD) {
            __syncthreads();
            // acc[a][c] += sum_j G[row a][j] * Bs[j][col]  with cols = tx + 16*c (c < D/16)
            const int nc = D / 16;
            for (int j = 0; j < NT; ++j) {
                float gv[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) gv[a] = G[(ty * 4 + a) * (NT + 1) + j];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (c < nc) {
                        float bv = Bs[j * D + tx + 16 * c];
#pragma unroll
                        for (int a = 0; a < 4; ++a) acc[a][c] = fmaf(gv[a], bv, acc[a][c]);
                    }
                }
            }
        }
    }
    if (!GRAD) {
        // reduce the 16 tx-lanes that share a row (they are contiguous lanes of a half-warp)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float v = rs[a];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (tx == 0 && i0 + ty * 4 + a < B) out[i0 + ty * 4 + a] = v;
        }
    } else {
        const int nc = D / 16;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            int gi = i0 + ty * 4 + a;
            if (gi < B) {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c < nc) out[(size_t)gi * D + tx + 16 * c] = acc[a][c] * inv_temp;
            }
        }
    }
}

// normalise rows: xh = x / max(||x||, eps); also the diagonal similarity for the loss is formed later
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ x, int B, int D, float* __restrict__ xh,
                                                             float* __restrict__ denom_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + warp; r < B; r += gridDim.x * 8) {
        float ss = 0.f;
        for (int k = lane; k < D; k += 32) {
            float v = __ldg(x + (size_t)r * D + k);
            ss += v * v;
        }
        float den = fmaxf(sqrtf(warp_sum(ss)), NORM_EPS);
        for (int k = lane; k < D; k += 32) xh[(size_t)r * D + k] = __ldg(x + (size_t)r * D + k) / den;
        if (lane == 0) denom_out[r] = den;
    }
}

// loss = sum_i [log rowsum_i + log colsum_i + 2/temp - 2 sim_ii] / (2B);  one warp per row for the diagonal dot
__global__ void __launch_bounds__(256) infonce_loss_kernel(const float* __restrict__ ah, const float* __restrict__ bh,
                                                           const float* __restrict__ rowsum, const float* __restrict__ colsum, int B,
                                                           int D, float inv_temp, float* __restrict__ loss_out, float* __restrict__ work) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int r = blockIdx.x * 8 + warp; r < B; r += gridDim.x * 8) {
        float d = 0.f;
        for (int k = lane; k < D; k += 32) d += __ldg(ah + (size_t)r * D + k) * __ldg(bh + (size_t)r * D + k);
        d = warp_sum(d);
        if (lane == 0) acc += logf(rowsum[r]) + logf(colsum[r]) + 2.f * inv_temp - 2.f * d * inv_temp;
    }
    grid_sum_ordered(block_sum_ordered(acc / (2.f * (float)B)), work, loss_out);
}

// NT-Xent: loss = sum_i [log rowsum_i + 1/temp - sim(i, partner(i))] / N  (rowsum excludes the self-similarity)
__global__ void __launch_bounds__(256) ntxent_loss_kernel(const float* __restrict__ rh, const float* __restrict__ rowsum, int N, int D,
                                                          float inv_temp, float* __restrict__ loss_out, float* __restrict__ work) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = N / 2;
    float acc = 0.f;
    for (int r = blockIdx.x * 8 + warp; r < N; r += gridDim.x * 8) {
        const int p = (r + half) % N;
        float d = 0.f;
        for (int k = lane; k < D; k += 32) d += __ldg(rh + (size_t)r * D + k) * __ldg(rh + (size_t)p * D + k);
        d = warp_sum(d);
        if (lane == 0) acc += logf(rowsum[r]) + inv_temp - d * inv_temp;
    }
    grid_sum_ordered(block_sum_ordered(acc / (float)N), work, loss_out);
}

// backward through the row normalisation: dx = (dxh - xh <xh, dxh>) / denom
__global__ void __launch_bounds__(256) normalize_bwd_kernel(const float* __restrict__ xh, const float* __restrict__ dxh,
                                                            const float* __restrict__ denom, int B, int D, float gscale,
                                                            float* __restrict__ dx) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + warp; r < B; r += gridDim.x * 8) {
        float dot = 0.f;
        for (int k = lane; k < D; k += 32) dot += __ldg(xh + (size_t)r * D + k) * __ldg(dxh + (size_t)r * D + k);
        dot = warp_sum(dot);
        float den = denom[r];
        for (int k = lane; k < D; k += 32)
            dx[(size_t)r * D + k] = (__ldg(dxh + (size_t)r * D + k) - __ldg(xh + (size_t)r * D + k) * dot) / den * gscale;
    }
}

// ---------------------------------------------------------------------------------------------------------
// InfoNCE on the tensor cores (csrc/gemm_tc.cu): S = Ah Bh^T as a tcgen05 tf32 GEMM whose epilogue turns every tile into
// e = exp(S/temp - 1/temp), stores it as bf16 and emits per-(column tile, row) partial sums (-> rowsum, fixed order);
// the same call with swapped operands gives E^T and the column sums.  The gradients are then plain bf16 tensor-core
// GEMMs   T = E [Bh | Bh/colsum]   and   T' = E^T [Ah | Ah/rowsum]   followed by an element-wise combine:
//     dAh_i = (0.5/B (T1_i / rowsum_i + T2_i) - Bh_i / B) / temp.
// E is B x B bf16 (134 MB at B = 8192): written once, read once per side; 3 x 2 B^2 D FLOP on the tensor cores replace the
// 6 x 2 B^2 D FLOP of recomputing SIMT tiles.
// ---------------------------------------------------------------------------------------------------------
int launch_gemm_tc_ex(const void* A, int64_t lda, const void* Bm, int64_t ldb, void* C, int64_t ldc, int M, int N, int K, int splits, int epi,
                      const float* bias, const uint8_t* mask, float keep_scale, float* aux, float fparam, int bf16, cudaStream_t st);

// tf32 keeps 10 mantissa bits: at 1/temp = 14 that is a 4e-3 error on the logits.  Split x = hi + lo (hi = x with the low 13
// mantissa bits cleared: exact in tf32; lo = x - hi: exact in fp32) and lay the pieces out along K so that ONE tf32 GEMM
// computes  a_hi b_hi + a_hi b_lo + a_lo b_hi  (the 3xTF32 scheme, error ~2^-21):  X1 = [hi | hi | lo],  X2 = [hi | lo | hi].
__global__ void __launch_bounds__(256) infonce_split3_kernel(const float* __restrict__ xh, int B, int D, float* __restrict__ x1,
                                                             float* __restrict__ x2) {
    const size_t n = (size_t)B * D;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const size_t i = e / D;
        const int d = (int)(e - i * D);
        const float x = xh[e];
        const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u), lo = x - hi;
        float* r1 = x1 + i * 3 * D;
        float* r2 = x2 + i * 3 * D;
        r1[d] = hi; r1[D + d] = hi; r1[2 * D + d] = lo;
        r2[d] = hi; r2[D + d] = lo; r2[2 * D + d] = hi;
    }
}

__global__ void __launch_bounds__(256) sum_parts_kernel(const float* __restrict__ part, int n_parts, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f;
    for (int p = 0; p < n_parts; ++p) a += part[(size_t)p * n + i];
    out[i] = a;
}

// YT bf16 [2D][ld]: YT[d][j] = xh[j][d], YT[D+d][j] = xh[j][d] / s[j]   (K-major B operand of the gradient GEMM)
__global__ void __launch_bounds__(256) infonce_build_yt_kernel(const float* __restrict__ xh, const float* __restrict__ s, int B, int D, int ld,
                                                               __nv_bfloat16* __restrict__ yt) {
    __shared__ float tile[32][33];
    const int j0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = j0 + ty + 8 * i, d = d0 + tx;
        tile[ty + 8 * i][tx] = (j < B && d < D) ? __ldg(xh + (size_t)j * D + d) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = d0 + ty + 8 * i, j = j0 + tx;
        if (d < D && j < B) {
            const float v = tile[tx][ty + 8 * i];
            yt[(size_t)d * ld + j] = __float2bfloat16_rn(v);
            yt[(size_t)(D + d) * ld + j] = __float2bfloat16_rn(v / __ldg(s + j));
        }
    }
}

// dxh[i][d] = inv_temp * (0.5*coef*(T[i][d] / own[i] + T[i][D+d]) - coef * other_h[i][d])
__global__ void __launch_bounds__(256) infonce_combine_kernel(const float* __restrict__ T, const float* __restrict__ own,
                                                              const float* __restrict__ other_h, int B, int D, float inv_temp, float coef,
                                                              float* __restrict__ dxh) {
    const size_t n = (size_t)B * D;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const size_t i = e / D;
        const int d = (int)(e - i * D);
        const float t1 = T[i * 2 * D + d], t2 = T[i * 2 * D + D + d];
        dxh[e] = inv_temp * (0.5f * coef * (t1 / __ldg(own + i) + t2) - coef * __ldg(other_h + e));
    }
}

// ---------------------------------------------------------------------------------------------------------
// Cosine consistency (unimodal): mean over pairs i<j of mean_b (1 - <e_i, e_j>)^2, e = normalize(emb).
// One CTA (128 threads) per sample; the V normalised vectors sit in shared memory.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) cosine_consistency_kernel(const float* __restrict__ emb, int V, int B, int D, float gscale,
                                                                 float* __restrict__ grad, float* __restrict__ loss_out,
                                                                 float* __restrict__ work) {
    extern __shared__ float sm[];          // e[V][D], ge[V][D], den[V], simm[V*V]
    float* e = sm;
    float* ge = e + (size_t)V * D;
    float* den = ge + (size_t)V * D;
    float* simm = den + V;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int npairs = V * (V - 1) / 2;
    const float coef = 1.0f / ((float)npairs * (float)B);
    for (int v = warp; v < V; v += nw) {
        const float* p = emb + ((size_t)v * B + b) * D;
        float ss = 0.f;
        for (int k = lane; k < D; k += 32) ss += __ldg(p + k) * __ldg(p + k);
        float dn = fmaxf(sqrtf(warp_sum(ss)), NORM_EPS);
        for (int k = lane; k < D; k += 32) {
            e[v * D + k] = __ldg(p + k) / dn;
            ge[v * D + k] = 0.f;
        }
        if (lane == 0) den[v] = dn;
    }
    __syncthreads();
    for (int pr = warp; pr < V * V; pr += nw) {
        int i = pr / V, j = pr - i * V;
        float d = 0.f;
        for (int k = lane; k < D; k += 32) d += e[i * D + k] * e[j * D + k];
        d = warp_sum(d);
        if (lane == 0) simm[pr] = d;
    }
    __syncthreads();
    {
        float l = 0.f;
        if (threadIdx.x == 0)
            for (int i = 0; i < V; ++i)
                for (int j = i + 1; j < V; ++j) l += (1.f - simm[i * V + j]) * (1.f - simm[i * V + j]);
        grid_sum_ordered(l * coef, work, loss_out);
    }
    // d loss / d e_i = sum_{j != i} -2 (1 - s_ij) e_j * coef
    for (int v = warp; v < V; v += nw) {
        for (int k = lane; k < D; k += 32) {
            float g = 0.f;
            for (int j = 0; j < V; ++j)
                if (j != v) g += -2.f * (1.f - simm[v * V + j]) * e[j * D + k];
            ge[v * D + k] = g * coef * gscale;
        }
        float dot = 0.f;
        for (int k = lane; k < D; k += 32) dot += e[v * D + k] * ge[v * D + k];
        dot = warp_sum(dot);
        float* gp = grad + ((size_t)v * B + b) * D;
        for (int k = lane; k < D; k += 32) gp[k] = (ge[v * D + k] - e[v * D + k] * dot) / den[v];
    }
}

static int rows_grid(int B) {
    int g = (B + 7) / 8;
    int cap = sm_count() * 4;
    return g < cap ? (g > 0 ? g : 1) : cap;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_dino_loss_parts(int B) { return rows_grid(B); }

int b200_dino_loss_fwd_bwd(const float* s, const float* t, const float* center, const float* t_colmean, int Vs, int Vt,
                           int B, int D, float tau_s, float tau_t, float grad_scale, int variant, float* grad_s,
                           float* part_loss, float* part_colsum, void* stream) {
    B200_REQUIRE(s && t && center && grad_s && part_loss && part_colsum, B200_E_ARG, "dino_loss: null pointer");
    B200_REQUIRE(Vs > 0 && Vt > 0 && B > 0, B200_E_ARG, "dino_loss: non-positive size");
    B200_REQUIRE(variant == 0 || (variant == 1 && t_colmean), B200_E_ARG, "dino_loss: variant 1 needs t_colmean");
    B200_REQUIRE(D % 32 == 0 && D >= 32 && D <= 512, B200_E_SHAPE, "dino_loss: D=%d must be a multiple of 32 in [32,512]", D);
    const int grid = rows_grid(B);
    cudaStream_t st = as_stream(stream);
#define DL_CASE(N)                                                                                                    \
    case N:                                                                                                           \
        dino_loss_kernel<N><<<grid, LOSS_WARPS * 32, 0, st>>>(s, t, center, t_colmean, Vs, Vt, B, tau_s, tau_t, grad_scale, \
                                                              variant, grad_s, part_loss, part_colsum);              \
        break;
    switch (D / 32) {
        DL_CASE(1) DL_CASE(2) DL_CASE(3) DL_CASE(4) DL_CASE(5) DL_CASE(6) DL_CASE(7) DL_CASE(8)
        DL_CASE(9) DL_CASE(10) DL_CASE(11) DL_CASE(12) DL_CASE(16)
        default:
            B200_REQUIRE(false, B200_E_SHAPE, "dino_loss: D=%d not compiled", D);
    }
#undef DL_CASE
    return launch_status("dino_loss_fwd_bwd");
}

int b200_teacher_norm_colmean(const float* t, const float* center, int Vt, int B, int D, float* out, void* stream) {
    B200_REQUIRE(t && center && out && Vt > 0 && B > 0 && D > 0, B200_E_ARG, "teacher_norm_colmean: bad arguments");
    teacher_norm_colmean_kernel<<<Vt, 256, 8 * D * sizeof(float), as_stream(stream)>>>(t, center, B, D, out);
    return launch_status("teacher_norm_colmean");
}

int b200_center_update(float* center, const float* part_colsum, const float* part_loss, int n_parts, int D, int64_t n_rows,
                       float m_c, float one_minus_mc, float* loss_out, float* colsum_out, void* stream) {
    B200_REQUIRE(part_colsum && n_parts > 0 && D > 0 && n_rows > 0 && (center || colsum_out), B200_E_ARG, "center_update: bad arguments");
    center_update_kernel<<<(D + 127) / 128, 128, 0, as_stream(stream)>>>(center, part_colsum, part_loss, n_parts, D,
                                                                         1.0f / (float)n_rows, m_c, one_minus_mc, loss_out,
                                                                         colsum_out);
    return launch_status("center_update");
}

int b200_center_apply(float* center, const float* colsum, int D, int64_t n_rows, float m_c, float one_minus_mc, void* stream) {
    B200_REQUIRE(center && colsum && D > 0 && n_rows > 0, B200_E_ARG, "center_apply: bad arguments");
    center_update_kernel<<<(D + 127) / 128, 128, 0, as_stream(stream)>>>(center, colsum, nullptr, 1, D, 1.0f / (float)n_rows, m_c,
                                                                         one_minus_mc, nullptr, nullptr);
    return launch_status("center_apply");
}

int64_t b200_loss_work_floats(int B) { return (int64_t)(B > 0 ? B : 0) + 1025; }

int b200_mse_align_fwd_bwd(const float* a, const float* b, int B, int D, float grad_scale, float* grad_a, float* grad_b,
                           float* loss_out, float* work, void* stream) {
    B200_REQUIRE(a && b && grad_a && grad_b && loss_out && work && B > 0 && D > 0, B200_E_ARG, "mse_align: bad arguments");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(work, 0, sizeof(float), st);          // the ticket of the fixed-order grid sum
    mse_align_kernel<<<rows_grid(B), 256, 0, st>>>(a, b, B, D, grad_scale, grad_a, grad_b, loss_out, work);
    return launch_status("mse_align_fwd_bwd");
}

int b200_ce_fwd_bwd(const float* logits, const int64_t* labels, int B, int C, float grad_scale, float* grad_logits,
                    float* loss_out, float* work, void* stream) {
    B200_REQUIRE(logits && labels && grad_logits && loss_out && work && B > 0 && C > 0, B200_E_ARG, "ce: bad arguments");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(work, 0, sizeof(float), st);
    int grid = (B + 127) / 128;
    ce_kernel<<<grid, 128, 0, st>>>(logits, labels, B, C, grad_scale, grad_logits, loss_out, work);
    return launch_status("ce_fwd_bwd");
}

int64_t b200_infonce_work_floats(int B, int D) { return (int64_t)4 * B * D + (int64_t)4 * B; }

int b200_infonce_fwd_bwd(const float* a, const float* b, int B, int D, float temperature, float grad_scale, float* grad_a,
                         float* grad_b, float* loss_out, float* work, void* stream) {
    B200_REQUIRE(a && b && grad_a && grad_b && loss_out && work && B > 0, B200_E_ARG, "infonce: bad arguments");
    B200_REQUIRE(D % 16 == 0 && D >= 16 && D <= 256, B200_E_SHAPE, "infonce: D=%d must be a multiple of 16 in [16,256]", D);
    cudaStream_t st = as_stream(stream);
    float* ah = work;
    float* bh = ah + (size_t)B * D;
    float* dah = bh + (size_t)B * D;
    float* dbh = dah + (size_t)B * D;
    float* dena = dbh + (size_t)B * D;
    float* denb = dena + B;
    float* rowsum = denb + B;
    float* colsum = rowsum + B;
    const float inv_temp = 1.0f / temperature;
    const int rg = rows_grid(B), tg = (B + NT - 1) / NT;
    const size_t sm_f = (size_t)2 * D * NT * sizeof(float);
    const size_t sm_g = sm_f + ((size_t)NT * D + (size_t)NT * (NT + 1)) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(infonce_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(infonce_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_done = true;
    }
    cudaMemsetAsync(dah, 0, sizeof(float), st);           // dah is free until the gradient kernels: ticket + partials of the loss sum
    normalize_rows_kernel<<<rg, 256, 0, st>>>(a, B, D, ah, dena);
    normalize_rows_kernel<<<rg, 256, 0, st>>>(b, B, D, bh, denb);
    infonce_rows_kernel<false><<<tg, 256, sm_f, st>>>(ah, bh, B, D, inv_temp, nullptr, nullptr, 0.f, rowsum, 0);
    infonce_rows_kernel<false><<<tg, 256, sm_f, st>>>(bh, ah, B, D, inv_temp, nullptr, nullptr, 0.f, colsum, 0);
    infonce_loss_kernel<<<rg, 256, 0, st>>>(ah, bh, rowsum, colsum, B, D, inv_temp, loss_out, dah);
    const float coef = 1.0f / (float)B;
    infonce_rows_kernel<true><<<tg, 256, sm_g, st>>>(ah, bh, B, D, inv_temp, rowsum, colsum, coef, dah, 0);
    infonce_rows_kernel<true><<<tg, 256, sm_g, st>>>(bh, ah, B, D, inv_temp, colsum, rowsum, coef, dbh, 0);
    normalize_bwd_kernel<<<rg, 256, 0, st>>>(ah, dah, dena, B, D, grad_scale, grad_a);
    normalize_bwd_kernel<<<rg, 256, 0, st>>>(bh, dbh, denb, B, D, grad_scale, grad_b);
    return launch_status("infonce_fwd_bwd");
}

static int64_t pad4l(int64_t n) { return (n + 3) / 4 * 4; }

int64_t b200_ntxent_work_floats(int N, int D) { return (int64_t)2 * N * D + (int64_t)2 * N; }

int b200_ntxent_fwd_bwd(const float* reps, int N, int D, float temperature, float grad_scale, float* grad, float* loss_out, float* work,
                        void* stream) {
    B200_REQUIRE(reps && grad && loss_out && work && N > 1 && N % 2 == 0, B200_E_ARG, "ntxent: bad arguments (N = 2B rows)");
    B200_REQUIRE(D % 16 == 0 && D >= 16 && D <= 256, B200_E_SHAPE, "ntxent: D=%d must be a multiple of 16 in [16,256]", D);
    B200_REQUIRE(temperature > 0.f, B200_E_ARG, "ntxent: temperature must be positive");
    cudaStream_t st = as_stream(stream);
    float* rh = work;
    float* drh = rh + (size_t)N * D;
    float* den = drh + (size_t)N * D;
    float* rowsum = den + N;
    const float inv_temp = 1.0f / temperature;
    const int rg = rows_grid(N), tg = (N + NT - 1) / NT;
    const size_t sm_f = (size_t)2 * D * NT * sizeof(float), sm_g = sm_f + ((size_t)NT * D + (size_t)NT * (NT + 1)) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(infonce_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(infonce_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr = true;
    }
    normalize_rows_kernel<<<rg, 256, 0, st>>>(reps, N, D, rh, den);
    infonce_rows_kernel<false><<<tg, 256, sm_f, st>>>(rh, rh, N, D, inv_temp, nullptr, nullptr, 0.f, rowsum, N / 2);
    cudaMemsetAsync(drh, 0, sizeof(float), st);           // drh is free until the gradient kernel: ticket + partials of the loss sum
    ntxent_loss_kernel<<<rg, 256, 0, st>>>(rh, rowsum, N, D, inv_temp, loss_out, drh);
    // dL/dS_ij = (e_ij / rowsum_i - [j == partner(i)]) / N, and S = R R^T is symmetric in R:
    // dL/dR_hat = (H + H^T) R_hat / temp with (H + H^T)_ij = (e_ij/rs_i + e_ij/rs_j) / N - 2 [j == partner(i)] / N
    infonce_rows_kernel<true><<<tg, 256, sm_g, st>>>(rh, rh, N, D, inv_temp, rowsum, rowsum, 2.0f / (float)N, drh, N / 2);
    normalize_bwd_kernel<<<rg, 256, 0, st>>>(rh, drh, den, N, D, grad_scale, grad);
    return launch_status("ntxent_fwd_bwd");
}

int64_t b200_infonce_tc_work_floats(int B, int D) {
    if (B <= 0 || D <= 0) return 0;
    const int64_t nt = (B + 127) / 128, ldE = (B + 7) / 8 * 8;
    return pad4l(4LL * B * D) + pad4l(4LL * B) + pad4l(2 * nt * B) + pad4l(2LL * B * 2 * D) + pad4l(4LL * B * 3 * D) +
           pad4l((2LL * B * ldE + 2LL * 2 * D * ldE) / 2 + 8) + 64;
}

int b200_infonce_fwd_bwd_tc(const float* a, const float* b, int B, int D, float temperature, float grad_scale, float* grad_a,
                            float* grad_b, float* loss_out, float* work, void* stream) {
    B200_REQUIRE(a && b && grad_a && grad_b && loss_out && work && B > 0, B200_E_ARG, "infonce_tc: bad arguments");
    B200_REQUIRE(D % 32 == 0 && D >= 32 && D <= 1024, B200_E_SHAPE, "infonce_tc: D=%d must be a multiple of 32", D);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(work) & 15) == 0, B200_E_ARG, "infonce_tc: work must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int64_t nt = (B + 127) / 128, ldE = (B + 7) / 8 * 8;
    float* ah = work;
    float* bh = ah + (size_t)B * D;
    float* dah = bh + (size_t)B * D;
    float* dbh = dah + (size_t)B * D;
    float* dena = work + pad4l(4LL * B * D);
    float* denb = dena + B;
    float* rowsum = denb + B;
    float* colsum = rowsum + B;
    float* rowpart = dena + pad4l(4LL * B);
    float* colpart = rowpart + nt * B;
    float* Ta = rowpart + pad4l(2 * nt * B);
    float* Tb = Ta + (size_t)B * 2 * D;
    float* a1 = Ta + pad4l(2LL * B * 2 * D);             // [hi | hi | lo] / [hi | lo | hi] splits of ah and bh, [B][3D] each
    float* a2 = a1 + (size_t)B * 3 * D;
    float* b1 = a2 + (size_t)B * 3 * D;
    float* b2 = b1 + (size_t)B * 3 * D;
    __nv_bfloat16* E = reinterpret_cast<__nv_bfloat16*>(a1 + pad4l(4LL * B * 3 * D));
    __nv_bfloat16* ET = E + (size_t)B * ldE;
    __nv_bfloat16* YTa = ET + (size_t)B * ldE;          // built from ah / rowsum: B operand of the dBh GEMM
    __nv_bfloat16* YTb = YTa + (size_t)2 * D * ldE;     // built from bh / colsum: B operand of the dAh GEMM
    const float inv_temp = 1.0f / temperature;
    const int rg = rows_grid(B);
    cudaMemsetAsync(dah, 0, sizeof(float), st);           // dah is free until the combine kernels: ticket + partials of the loss sum
    normalize_rows_kernel<<<rg, 256, 0, st>>>(a, B, D, ah, dena);
    normalize_rows_kernel<<<rg, 256, 0, st>>>(b, B, D, bh, denb);
    const int sg = (int)(((size_t)B * D + 255) / 256 < (size_t)sm_count() * 8 ? ((size_t)B * D + 255) / 256 : (size_t)sm_count() * 8);
    infonce_split3_kernel<<<sg, 256, 0, st>>>(ah, B, D, a1, a2);
    infonce_split3_kernel<<<sg, 256, 0, st>>>(bh, B, D, b1, b2);
    int rc = launch_gemm_tc_ex(a1, 3 * D, b2, 3 * D, E, ldE, B, B, 3 * D, 1, 5, nullptr, nullptr, 1.f, rowpart, inv_temp, 0, st);
    if (rc) return rc;
    rc = launch_gemm_tc_ex(b1, 3 * D, a2, 3 * D, ET, ldE, B, B, 3 * D, 1, 5, nullptr, nullptr, 1.f, colpart, inv_temp, 0, st);
    if (rc) return rc;
    sum_parts_kernel<<<(B + 255) / 256, 256, 0, st>>>(rowpart, (int)nt, B, rowsum);
    sum_parts_kernel<<<(B + 255) / 256, 256, 0, st>>>(colpart, (int)nt, B, colsum);
    infonce_loss_kernel<<<rg, 256, 0, st>>>(ah, bh, rowsum, colsum, B, D, inv_temp, loss_out, dah);
    const dim3 tg((B + 31) / 32, (D + 31) / 32);
    infonce_build_yt_kernel<<<tg, 256, 0, st>>>(ah, rowsum, B, D, (int)ldE, YTa);
    infonce_build_yt_kernel<<<tg, 256, 0, st>>>(bh, colsum, B, D, (int)ldE, YTb);
    rc = launch_gemm_tc_ex(E, ldE, YTb, ldE, Ta, 2 * D, B, 2 * D, B, 1, 0, nullptr, nullptr, 1.f, nullptr, 0.f, 1, st);
    if (rc) return rc;
    rc = launch_gemm_tc_ex(ET, ldE, YTa, ldE, Tb, 2 * D, B, 2 * D, B, 1, 0, nullptr, nullptr, 1.f, nullptr, 0.f, 1, st);
    if (rc) return rc;
    const float coef = 1.0f / (float)B;
    const int cg = (int)(((size_t)B * D + 255) / 256 < (size_t)sm_count() * 8 ? ((size_t)B * D + 255) / 256 : (size_t)sm_count() * 8);
    infonce_combine_kernel<<<cg, 256, 0, st>>>(Ta, rowsum, bh, B, D, inv_temp, coef, dah);
    infonce_combine_kernel<<<cg, 256, 0, st>>>(Tb, colsum, ah, B, D, inv_temp, coef, dbh);
    normalize_bwd_kernel<<<rg, 256, 0, st>>>(ah, dah, dena, B, D, grad_scale, grad_a);
    normalize_bwd_kernel<<<rg, 256, 0, st>>>(bh, dbh, denb, B, D, grad_scale, grad_b);
    return launch_status("infonce_fwd_bwd_tc");
}

int b200_cosine_consistency_fwd_bwd(const float* emb, int V, int B, int D, float grad_scale, float* grad_emb, float* loss_out,
                                    float* work, void* stream) {
    B200_REQUIRE(emb && grad_emb && loss_out && work && V > 1 && V <= 16 && B > 0 && D > 0, B200_E_ARG, "cosine_consistency: bad arguments");
    cudaStream_t st = as_stream(stream);
    size_t smem = ((size_t)2 * V * D + V + (size_t)V * V) * sizeof(float);
    B200_REQUIRE(smem <= 48 * 1024, B200_E_SHAPE, "cosine_consistency: V*D too large");
    cudaMemsetAsync(work, 0, sizeof(float), st);
    cosine_consistency_kernel<<<B, 128, smem, st>>>(emb, V, B, D, grad_scale, grad_emb, loss_out, work);
    return launch_status("cosine_consistency_fwd_bwd");
}

}  // extern "C"
