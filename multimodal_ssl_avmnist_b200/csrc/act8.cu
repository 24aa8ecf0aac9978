// BatchNorm(train)-apply + ReLU + 2x2 max-pool, forward and backward, on bf16 "act8" activations
// [N][C/8][H][W][8] (the layout the tensor-core convolutions of conv_tc.cu read and write).
// Reference: the BN -> ReLU -> MaxPool2 triplets of models/unimodal.py:130-140, 187-208 and models/dino.py:21-33.
//
// One thread owns one pooled pixel of one channel octet: four 16-byte loads of the 2x2 window (two per row, adjacent),
// all 8 channels in registers, one 16-byte store -- every access is a full, coalesced 16-byte unit.  A block works on
// one (view-call, octet) pair, so the 8 (scale, shift, mean, invstd) tuples are block-uniform and the backward
// reductions are per-thread registers -> warp shuffles -> one double atomicAdd per block and channel.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace b200 {
namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xFFFF0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xFFFF0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xFFFF0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xFFFF0000u);
}
__device__ __forceinline__ void unpack8h(const uint4& u, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = __half22float2(h[i]);
        f[2 * i] = v.x;
        f[2 * i + 1] = v.y;
    }
}
template <bool F16>
__device__ __forceinline__ void unpackz(const uint4& u, float (&f)[8]) {
    if (F16) unpack8h(u, f);
    else unpack8(u, f);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}

// two adjacent 16-byte units (the two columns of a pooling window) as ONE 256-bit access when the pair is 32-byte aligned
// (even W: always): half the load / store instructions and L1 wavefronts of the strided per-thread accesses
__device__ __forceinline__ void ld_pair(const uint4* p, bool wide, uint4& a, uint4& b) {
    if (wide) {
        asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                     : "l"(p));
    } else {
        a = __ldg(p);
        b = __ldg(p + 1);
    }
}
__device__ __forceinline__ void st_pair(uint4* p, bool wide, const uint4& a, const uint4& b) {
    if (wide) {
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
                     "r"(b.y), "r"(b.z), "r"(b.w)
                     : "memory");
    } else {
        p[0] = a;
        p[1] = b;
    }
}

struct Tile {      // work split: grid (chunks, octets, views)
    int n_per_view, C, H, W, HP, WP, P;
    long u0, u1;   // unit range [u0, u1) of this block inside its (view, octet): unit = sample_in_view * HP*WP + pooled pixel
};
__device__ __forceinline__ Tile make_tile(int n_per_view, int C, int H, int W) {
    Tile t;
    t.n_per_view = n_per_view; t.C = C; t.H = H; t.W = W; t.HP = H >> 1; t.WP = W >> 1; t.P = C >> 3;
    const long units = (long)n_per_view * t.HP * t.WP;
    t.u0 = units * blockIdx.x / gridDim.x;
    t.u1 = units * (blockIdx.x + 1) / gridDim.x;
    return t;
}

// out_fmt 0: fp32 NCHW [N][C][HP][WP]; 1: bf16 act8 [N][C/8][HP][WP][8]
template <bool ZF16>
__global__ void __launch_bounds__(256) bn_relu_pool8_fwd_kernel(const uint4* __restrict__ z8, const float* __restrict__ scale,
                                                                const float* __restrict__ shift, void* __restrict__ out, int n_per_view,
                                                                int C, int H, int W, int out_fmt) {
    const Tile t = make_tile(n_per_view, C, H, W);
    const int oct = blockIdx.y, v = blockIdx.z;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = __ldg(scale + v * C + oct * 8 + j);
        b[j] = __ldg(shift + v * C + oct * 8 + j);
    }
    const int hw = t.HP * t.WP;
    const bool wide = ((W & 1) == 0) && ((reinterpret_cast<uintptr_t>(z8) & 31) == 0);
    for (long u = t.u0 + threadIdx.x; u < t.u1; u += blockDim.x) {
        const int s = (int)(u / hw), e = (int)(u - (long)s * hw);
        const int py = e / t.WP, px = e - py * t.WP;
        const long n = (long)v * n_per_view + s;
        const uint4* zp = z8 + ((n * t.P + oct) * H + 2 * py) * W + 2 * px;
        float w0[8], w1[8], w2[8], w3[8], m[8];
        uint4 r0, r1, r2, r3;
        ld_pair(zp, wide, r0, r1);
        ld_pair(zp + W, wide, r2, r3);
        unpackz<ZF16>(r0, w0);
        unpackz<ZF16>(r1, w1);
        unpackz<ZF16>(r2, w2);
        unpackz<ZF16>(r3, w3);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float y = fmaxf(fmaxf(fmaf(a[j], w0[j], b[j]), fmaf(a[j], w1[j], b[j])), fmaxf(fmaf(a[j], w2[j], b[j]), fmaf(a[j], w3[j], b[j])));
            m[j] = fmaxf(y, 0.f);
        }
        if (out_fmt) {
            reinterpret_cast<uint4*>(out)[((n * t.P + oct) * t.HP + py) * t.WP + px] = pack8(m);
        } else {
            float* op = reinterpret_cast<float*>(out) + ((n * C + oct * 8) * t.HP + py) * t.WP + px;
#pragma unroll
            for (int j = 0; j < 8; ++j) op[(long)j * hw] = m[j];
        }
    }
}

// BatchNorm-apply + ReLU on the POOLED extreme e (fp16 act8 [N][C/8][HP][WP][8]) that the convolution's fused max-pool epilogue
// emitted (conv_tc.cu): p = ReLU(a e + b).  Because e = max(z) where gamma >= 0 and min(z) where gamma < 0 and a = gamma * invstd,
// a e + b is exactly the maximum of a z + b over the window: bit-identical to bn_relu_pool8_fwd_kernel on the full-resolution z,
// at a quarter of its reads.  Two adjacent units per thread (256-bit accesses) when the plane size is even.
__global__ void __launch_bounds__(256) bn_relu_apply8_kernel(const uint4* __restrict__ e8, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, void* __restrict__ out, int n_per_view, int C,
                                                             int HP, int WP, int out_fmt) {
    const int oct = blockIdx.y, v = blockIdx.z, P = C >> 3;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = __ldg(scale + v * C + oct * 8 + j);
        b[j] = __ldg(shift + v * C + oct * 8 + j);
    }
    const int hw = HP * WP;
    const bool wide = ((hw & 1) == 0) && ((reinterpret_cast<uintptr_t>(e8) & 31) == 0) && (out_fmt == 0 || (reinterpret_cast<uintptr_t>(out) & 31) == 0);
    const int step = wide ? 2 : 1;
    const long units = (long)n_per_view * hw / step;
    const long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
    for (long uu = u0 + threadIdx.x; uu < u1; uu += blockDim.x) {
        const long u = uu * step;
        const int s = (int)(u / hw), e = (int)(u - (long)s * hw);
        const long n = (long)v * n_per_view + s;
        const long base = (n * P + oct) * hw + e;
        uint4 r0, r1 = make_uint4(0, 0, 0, 0);
        if (wide) ld_pair(e8 + base, true, r0, r1);
        else r0 = __ldg(e8 + base);
        float w0[8], w1[8], m0[8], m1[8];
        unpack8h(r0, w0);
        unpack8h(r1, w1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            m0[j] = fmaxf(fmaf(a[j], w0[j], b[j]), 0.f);
            m1[j] = fmaxf(fmaf(a[j], w1[j], b[j]), 0.f);
        }
        if (out_fmt) {
            uint4* op = reinterpret_cast<uint4*>(out) + base;
            if (wide) st_pair(op, true, pack8(m0), pack8(m1));
            else *op = pack8(m0);
        } else {
            float* op = reinterpret_cast<float*>(out) + (n * C + oct * 8) * hw + e;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                op[(long)j * hw] = m0[j];
                if (wide) op[(long)j * hw + 1] = m1[j];
            }
        }
    }
}

// argmax in PyTorch scan order (first maximum wins)
__device__ __forceinline__ int argmax4(float y0, float y1, float y2, float y3, float& m) {
    int k = 0;
    m = y0;
    if (y1 > m) { m = y1; k = 1; }
    if (y2 > m) { m = y2; k = 2; }
    if (y3 > m) { m = y3; k = 3; }
    return k;
}

// dp_fmt 0: fp32 NCHW [N][C][HP][WP]; 1: bf16 act8.  APPLY = false: sums[view][c] += {sum g, sum g*xhat};
// APPLY = true: dz8 = a*(g_at_argmax - mean(g) - xhat*mean(g*xhat)) for all four window positions.
template <bool APPLY, bool ZF16>
__global__ void __launch_bounds__(256, 3) bn_relu_pool8_bwd_kernel(const uint4* __restrict__ z8, const void* __restrict__ dp,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                double* __restrict__ sums, uint4* __restrict__ dz8, double* __restrict__ dbsum,
                                                                int n_per_view, int C, int H, int W, int dp_fmt) {
    const Tile t = make_tile(n_per_view, C, H, W);
    const int oct = blockIdx.y, v = blockIdx.z;
    // per-channel constants: y = a*z + b decides arg-max / ReLU; reduce: xhat = z*is + nm (nm = -mu*is);
    // apply: dz = a*(dy - k1 - xhat*k2) = ca*z + cb + (a*g at the arg-max), ca = -a*is*k2, cb = -a*(k1 + nm*k2)
    float a[8], b[8], c0[8], c1[8], s1[8], s2[8];
    const float inv_cnt = 1.0f / ((float)n_per_view * (float)H * (float)W);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = v * C + oct * 8 + j;
        a[j] = __ldg(scale + c); b[j] = __ldg(shift + c);
        const float is = __ldg(invstd + c), nm = -__ldg(mean + c) * is;
        s1[j] = s2[j] = 0.f;
        if (APPLY) {
            const float k1 = (float)sums[(size_t)c * 2] * inv_cnt, k2 = (float)sums[(size_t)c * 2 + 1] * inv_cnt;
            c0[j] = -a[j] * is * k2;
            c1[j] = -a[j] * (k1 + nm * k2);
        } else {
            c0[j] = is;
            c1[j] = nm;
        }
    }
    const int hw = t.HP * t.WP;
    const bool wide = ((W & 1) == 0) && ((reinterpret_cast<uintptr_t>(z8) & 31) == 0);
    const bool wide_out = APPLY && ((W & 1) == 0) && ((reinterpret_cast<uintptr_t>(dz8) & 31) == 0);
    for (long u = t.u0 + threadIdx.x; u < t.u1; u += blockDim.x) {
        const int s = (int)(u / hw), e = (int)(u - (long)s * hw);
        const int py = e / t.WP, px = e - py * t.WP;
        const long n = (long)v * n_per_view + s;
        const long zoff = ((n * t.P + oct) * H + 2 * py) * W + 2 * px;
        const uint4* zp = z8 + zoff;
        float w[4][8], g[8];
        {
            uint4 r0, r1, r2, r3;
            ld_pair(zp, wide, r0, r1);
            ld_pair(zp + W, wide, r2, r3);
            unpackz<ZF16>(r0, w[0]);
            unpackz<ZF16>(r1, w[1]);
            unpackz<ZF16>(r2, w[2]);
            unpackz<ZF16>(r3, w[3]);
        }
        if (dp_fmt) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(dp) + ((n * t.P + oct) * t.HP + py) * t.WP + px), g);
        } else {
            const float* gp = reinterpret_cast<const float*>(dp) + ((n * C + oct * 8) * t.HP + py) * t.WP + px;
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = __ldg(gp + (long)j * hw);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float m;
            const int k = argmax4(fmaf(a[j], w[0][j], b[j]), fmaf(a[j], w[1][j], b[j]), fmaf(a[j], w[2][j], b[j]), fmaf(a[j], w[3][j], b[j]), m);
            const float gj = (m > 0.f) ? g[j] : 0.f;
            if (!APPLY) {
                const float zk = (k == 0) ? w[0][j] : (k == 1) ? w[1][j] : (k == 2) ? w[2][j] : w[3][j];
                s1[j] += gj;
                s2[j] += gj * fmaf(zk, c0[j], c1[j]);
            } else {
                const float ag = a[j] * gj;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    w[q][j] = fmaf(c0[j], w[q][j], c1[j]) + ((q == k) ? ag : 0.f);      // dz overwrites z in registers
                    s1[j] += w[q][j];
                }
            }
        }
        if (APPLY) {
            uint4* zo = dz8 + zoff;
            st_pair(zo, wide_out, pack8(w[0]), pack8(w[1]));
            st_pair(zo + W, wide_out, pack8(w[2]), pack8(w[3]));
        }
    }
    if (APPLY && ((H | W) & 1)) {
        // odd planes (7x7 -> floor-pooled 3x3): the last row / column is in no pooling window but still receives the dense
        // BatchNorm terms dz = ca*z + cb
        const int ex = ((W & 1) ? H : 0) + ((H & 1) ? W : 0) - (((H & 1) && (W & 1)) ? 1 : 0);     // positions per (sample, octet)
        const long units = (long)n_per_view * ex;
        const long e0 = units * blockIdx.x / gridDim.x, e1 = units * (blockIdx.x + 1) / gridDim.x;
        for (long u = e0 + threadIdx.x; u < e1; u += blockDim.x) {
            const int s = (int)(u / ex);
            int e = (int)(u - (long)s * ex), y, x;
            if ((W & 1) && e < H) { y = e; x = W - 1; }
            else { e -= (W & 1) ? H : 0; y = H - 1; x = e; }          // last row (without the corner already covered above)
            const long n = (long)v * n_per_view + s;
            const long off = ((n * t.P + oct) * H + y) * W + x;
            float zz[8];
            unpackz<ZF16>(__ldg(z8 + off), zz);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                zz[j] = fmaf(c0[j], zz[j], c1[j]);
                s1[j] += zz[j];
            }
            dz8[off] = pack8(zz);
        }
    }
    if (APPLY) {
        if (dbsum != nullptr) {        // conv bias gradient = sum of dz over the view-call (block-uniform branch)
            __shared__ float redb[8][8];
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x1 = warp_sum(s1[j]);
                if (lane == 0) redb[warp][j] = x1;
            }
            __syncthreads();
            if (threadIdx.x < 8) {
                double acc = 0.0;
                for (int wv = 0; wv < 8; ++wv) acc += (double)redb[wv][threadIdx.x];
                atomicAdd(&dbsum[oct * 8 + threadIdx.x], acc);
            }
        }
    } else {
        __shared__ float red[8][16];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x1 = warp_sum(s1[j]), x2 = warp_sum(s2[j]);
            if (lane == 0) {
                red[warp][j] = x1;
                red[warp][8 + j] = x2;
            }
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            double acc = 0.0;
            for (int wv = 0; wv < 8; ++wv) acc += (double)red[wv][threadIdx.x];
            const int j = threadIdx.x & 7, which = threadIdx.x >> 3;
            atomicAdd(&sums[((size_t)v * C + oct * 8 + j) * 2 + which], acc);
        }
    }
}

// BatchNorm-backward statistics from the POOLED tensors only.  At the arg-max position a*z+b = gamma*xhat + beta equals the
// pooled output p wherever ReLU let it through (p > 0), and the gradient is zero elsewhere, so
//     sum g = sum_{p>0} dp          sum g*xhat = sum_{p>0} dp * (p - beta) / gamma
// -- no read of the 4x larger pre-BatchNorm tensor z and no arg-max search.  p / dp: fp32 NCHW (fmt 0) or bf16 act8 (fmt 1).
__global__ void __launch_bounds__(256) bn_pool8_bwd_reduce_p_kernel(const void* __restrict__ p, const void* __restrict__ dp,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   double* __restrict__ sums, int n_per_view, int C, int HP, int WP, int p_fmt,
                                                                   int dp_fmt) {
    const int oct = blockIdx.y, v = blockIdx.z, P = C >> 3;
    const int hw = HP * WP;
    const long units = (long)n_per_view * hw;
    const long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
    float ig[8], be[8], s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float gj = __ldg(gamma + oct * 8 + j);
        ig[j] = gj != 0.f ? 1.0f / gj : 0.f;
        be[j] = __ldg(beta + oct * 8 + j);
        s1[j] = s2[j] = 0.f;
    }
    auto accumulate = [&](const float* pv, const float* g) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gj = pv[j] > 0.f ? g[j] : 0.f;
            s1[j] += gj;
            s2[j] += gj * ((pv[j] - be[j]) * ig[j]);
        }
    };
    if (p_fmt && dp_fmt) {
        // both tensors act8: four independent 16-byte loads per tensor in flight per thread
        const uint4* pb = reinterpret_cast<const uint4*>(p);
        const uint4* gb = reinterpret_cast<const uint4*>(dp);
        auto addr = [&](long u) {
            const long s = u / hw;
            return (((long)v * n_per_view + s) * P + oct) * hw + (u - s * hw);
        };
        long u = u0 + threadIdx.x;
        const long step = blockDim.x;
        for (; u + 3 * step < u1; u += 4 * step) {
            uint4 rp[4], rg[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const long a = addr(u + q * step);
                rp[q] = __ldg(pb + a);
                rg[q] = __ldg(gb + a);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float pv[8], g[8];
                unpack8(rp[q], pv);
                unpack8(rg[q], g);
                accumulate(pv, g);
            }
        }
        for (; u < u1; u += step) {
            const long a = addr(u);
            float pv[8], g[8];
            unpack8(__ldg(pb + a), pv);
            unpack8(__ldg(gb + a), g);
            accumulate(pv, g);
        }
    } else {
        for (long u = u0 + threadIdx.x; u < u1; u += blockDim.x) {
            const int s = (int)(u / hw), e = (int)(u - (long)s * hw);
            const long n = (long)v * n_per_view + s;
            float pv[8], g[8];
            if (p_fmt) {
                unpack8(__ldg(reinterpret_cast<const uint4*>(p) + (n * P + oct) * hw + e), pv);
            } else {
                const float* pp = reinterpret_cast<const float*>(p) + (n * C + oct * 8) * hw + e;
#pragma unroll
                for (int j = 0; j < 8; ++j) pv[j] = __ldg(pp + (long)j * hw);
            }
            if (dp_fmt) {
                unpack8(__ldg(reinterpret_cast<const uint4*>(dp) + (n * P + oct) * hw + e), g);
            } else {
                const float* gp = reinterpret_cast<const float*>(dp) + (n * C + oct * 8) * hw + e;
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] = __ldg(gp + (long)j * hw);
            }
            accumulate(pv, g);
        }
    }
    __shared__ float red[8][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float x1 = warp_sum(s1[j]), x2 = warp_sum(s2[j]);
        if (lane == 0) {
            red[warp][j] = x1;
            red[warp][8 + j] = x2;
        }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double acc = 0.0;
        for (int wv = 0; wv < 8; ++wv) acc += (double)red[wv][threadIdx.x];
        const int j = threadIdx.x & 7, which = threadIdx.x >> 3;
        atomicAdd(&sums[((size_t)v * C + oct * 8 + j) * 2 + which], acc);
    }
}

__global__ void bias_grad_finalize_kernel(const double* __restrict__ dbsum, float* __restrict__ db, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) db[c] = (float)dbsum[c];
}

// bf16 act8 -> fp32 NCHW (tests / debugging)
__global__ void unpack_act8_kernel(const uint4* __restrict__ x8, float* __restrict__ out, long n_units, int C, int HW) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_units) return;
    const int pix = (int)(i % HW);
    const long no = i / HW;
    const int oct = (int)(no % (C / 8));
    const long n = no / (C / 8);
    float f[8];
    unpack8(__ldg(x8 + i), f);
    float* o = out + ((n * C + oct * 8) * (long)HW) + pix;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[(long)j * HW] = f[j];
}

dim3 tile_grid(int N, int n_per_view, int C, int H, int W) {
    const int views = N / n_per_view, P = C / 8;
    const long units = (long)n_per_view * (H / 2) * (W / 2);
    long chunks = (8L * sm_count() + (long)views * P - 1) / ((long)views * P);
    const long max_chunks = (units + 255) / 256;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    return dim3((unsigned)chunks, (unsigned)P, (unsigned)views);
}

int check_shape(const char* what, int N, int n_per_view, int C, int H, int W) {
    B200_REQUIRE(N > 0 && n_per_view > 0 && N % n_per_view == 0, -2, "%s: N=%d must be a multiple of n_per_view=%d", what, N, n_per_view);
    B200_REQUIRE(C % 8 == 0 && C > 0, -2, "%s: C=%d must be a multiple of 8", what, C);
    B200_REQUIRE(H > 1 && W > 1, -2, "%s: H=%d, W=%d must be at least 2 (odd sizes floor-pool like nn.MaxPool2d)", what, H, W);
    B200_REQUIRE(N / n_per_view <= 65535, -2, "%s: too many view-calls", what);
    return 0;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_bn_relu_pool8_fwd(const void* z8, const float* scale, const float* shift, void* out, int N, int n_per_view, int C, int H, int W,
                           int z_f16, int out_fmt, void* stream) {
    B200_REQUIRE(z8 && scale && shift && out, -1, "bn_relu_pool8_fwd: null pointer");
    int rc = check_shape("bn_relu_pool8_fwd", N, n_per_view, C, H, W);
    if (rc) return rc;
    if (z_f16)
        bn_relu_pool8_fwd_kernel<true><<<tile_grid(N, n_per_view, C, H, W), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(z8), scale,
                                                                                                         shift, out, n_per_view, C, H, W, out_fmt);
    else
        bn_relu_pool8_fwd_kernel<false><<<tile_grid(N, n_per_view, C, H, W), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(z8), scale,
                                                                                                          shift, out, n_per_view, C, H, W, out_fmt);
    return launch_status("bn_relu_pool8_fwd_kernel");
}

int b200_bn_relu_apply8(const void* e8, const float* scale, const float* shift, void* out, int N, int n_per_view, int C, int HP, int WP,
                        int out_fmt, void* stream) {
    B200_REQUIRE(e8 && scale && shift && out, -1, "bn_relu_apply8: null pointer");
    B200_REQUIRE(N > 0 && n_per_view > 0 && N % n_per_view == 0 && C > 0 && C % 8 == 0 && HP > 0 && WP > 0 && N / n_per_view <= 65535, -2,
                 "bn_relu_apply8: bad shape");
    const int views = N / n_per_view, P = C / 8;
    const long units = (long)n_per_view * HP * WP;
    long chunks = (8L * sm_count() + (long)views * P - 1) / ((long)views * P);
    const long max_chunks = (units + 511) / 512;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    bn_relu_apply8_kernel<<<dim3((unsigned)chunks, (unsigned)P, (unsigned)views), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const uint4*>(e8), scale, shift, out, n_per_view, C, HP, WP, out_fmt);
    return launch_status("bn_relu_apply8_kernel");
}

int b200_bn_relu_pool8_bwd_reduce(const void* z8, const void* dp, const float* scale, const float* shift, const float* mean,
                                  const float* invstd, double* sums, int N, int n_per_view, int C, int H, int W, int z_f16, int dp_fmt,
                                  void* stream) {
    B200_REQUIRE(z8 && dp && scale && shift && mean && invstd && sums, -1, "bn_relu_pool8_bwd_reduce: null pointer");
    int rc = check_shape("bn_relu_pool8_bwd_reduce", N, n_per_view, C, H, W);
    if (rc) return rc;
    if (z_f16)
        bn_relu_pool8_bwd_kernel<false, true><<<tile_grid(N, n_per_view, C, H, W), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const uint4*>(z8), dp, scale, shift, mean, invstd, sums, nullptr, nullptr, n_per_view, C, H, W, dp_fmt);
    else
        bn_relu_pool8_bwd_kernel<false, false><<<tile_grid(N, n_per_view, C, H, W), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const uint4*>(z8), dp, scale, shift, mean, invstd, sums, nullptr, nullptr, n_per_view, C, H, W, dp_fmt);
    return launch_status("bn_relu_pool8_bwd_kernel<reduce>");
}

int b200_bn_relu_pool8_bwd_apply(const void* z8, const void* dp, const float* scale, const float* shift, const float* mean,
                                 const float* invstd, const double* sums, void* dz8, double* dbsum, int N, int n_per_view, int C, int H, int W,
                                 int z_f16, int dp_fmt, void* stream) {
    B200_REQUIRE(z8 && dp && scale && shift && mean && invstd && sums && dz8, -1, "bn_relu_pool8_bwd_apply: null pointer");
    int rc = check_shape("bn_relu_pool8_bwd_apply", N, n_per_view, C, H, W);
    if (rc) return rc;
    if (z_f16)
        bn_relu_pool8_bwd_kernel<true, true><<<tile_grid(N, n_per_view, C, H, W), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const uint4*>(z8), dp, scale, shift, mean, invstd, const_cast<double*>(sums), reinterpret_cast<uint4*>(dz8), dbsum,
            n_per_view, C, H, W, dp_fmt);
    else
        bn_relu_pool8_bwd_kernel<true, false><<<tile_grid(N, n_per_view, C, H, W), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const uint4*>(z8), dp, scale, shift, mean, invstd, const_cast<double*>(sums), reinterpret_cast<uint4*>(dz8), dbsum,
            n_per_view, C, H, W, dp_fmt);
    return launch_status("bn_relu_pool8_bwd_kernel<apply>");
}

int b200_bn_pool8_bwd_reduce_p(const void* p, const void* dp, const float* gamma, const float* beta, double* sums, int N, int n_per_view,
                               int C, int HP, int WP, int p_fmt, int dp_fmt, void* stream) {
    B200_REQUIRE(p && dp && gamma && beta && sums, -1, "bn_pool8_bwd_reduce_p: null pointer");
    B200_REQUIRE(N > 0 && n_per_view > 0 && N % n_per_view == 0 && C % 8 == 0 && C > 0 && HP > 0 && WP > 0, -2, "bn_pool8_bwd_reduce_p: bad shape");
    const int views = N / n_per_view, P = C / 8;
    const long units = (long)n_per_view * HP * WP;
    long chunks = (8L * sm_count() + (long)views * P - 1) / ((long)views * P);
    const long max_chunks = (units + 511) / 512;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    bn_pool8_bwd_reduce_p_kernel<<<dim3((unsigned)chunks, (unsigned)P, (unsigned)views), 256, 0, as_stream(stream)>>>(p, dp, gamma, beta, sums,
                                                                                                                  n_per_view, C, HP, WP, p_fmt, dp_fmt);
    return launch_status("bn_pool8_bwd_reduce_p_kernel");
}

int b200_bias_grad_finalize(const double* dbsum, float* db, int C, void* stream) {
    B200_REQUIRE(dbsum && db && C > 0, -1, "bias_grad_finalize: bad arguments");
    bias_grad_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(dbsum, db, C);
    return launch_status("bias_grad_finalize_kernel");
}

int b200_unpack_act8(const void* x8, float* out, int N, int C, int H, int W, void* stream) {
    B200_REQUIRE(x8 && out, -1, "unpack_act8: null pointer");
    B200_REQUIRE(C % 8 == 0 && N > 0, -2, "unpack_act8: C must be a multiple of 8");
    const long units = (long)N * (C / 8) * H * W;
    unpack_act8_kernel<<<(unsigned)((units + 255) / 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(x8), out, units, C, H * W);
    return launch_status("unpack_act8_kernel");
}

}  // extern "C"
