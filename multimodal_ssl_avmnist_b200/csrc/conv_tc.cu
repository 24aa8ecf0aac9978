// Tensor-core convolutions of the AVMNIST encoders (models/unimodal.py:129-140, 186-208; models/dino.py:20-30):
// tcgen05.mma with TMEM accumulators, operands staged by TMA, NO im2col.  Forward, data gradient, weight gradient; first
// layers (C_in = 1) included; the first-layer backward is fused with its BatchNorm / ReLU / max-pool backward.
//
// Activation layout ("act8"): bf16 [N][C/8][H][W][8] -- channel octets are planes, a pixel of a plane is one 16-byte
// unit.  A TMA box load of (8, W+2p, H'+K-1, C/8) starting at (-p, -p) drops a whole zero-padded image (or row band)
// into shared memory (out-of-bounds = 0 is the convolution padding).  With output pixels numbered flat over the PADDED
// pitch, q = y*pitch + x, the input pixel of tap (kh,kw) is q + kh*pitch + kw: for a tile of 128 consecutive q the A
// operand of every tap is the SAME smem image at a shifted start address -- a K-major SWIZZLE_NONE UMMA descriptor
// (8 consecutive pixels x 16 B = one core matrix, SBO = 128 B; the second K chunk is the next channel plane,
// LBO = plane stride, or for C_in = 8 the next kw tap).  Outputs at x >= W_out / y >= H_out are junk and are masked in
// the epilogue.
//
// A UMMA occupies the tensor pipe for its shared-memory operand fetch (26-48 cycles at these shapes) whatever its N, so
// the kernels minimise the NUMBER of UMMAs: adjacent output pixels are packed into N ("x phases", see xph_for / TcCfg),
// with the input slab de-interleaved by phase through strided TMA maps; first layers read the "quad8" image (8 consecutive
// padded pixels per unit = all taps of four output pixels).  See DESIGN.md 4.1 for the measurements behind this.
//
// Kernels: conv_tc_kernel (forward / data gradient: warp 0 = TMA producer, 1-4 MMA-issuer warps, 4 epilogue warps,
// 2-8 TMEM accumulator stages), conv_tc_wgrad_kernel / conv_tc_wgrad_ph_kernel (weight gradients, MN-major operands, one
// TMEM accumulator per (kh, plane)), conv_tc_wgrad_l0_fused_kernel (first layer: BatchNorm-backward-apply in shared memory
// + weight gradient), weight-image / input-image packing kernels.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace b200 {

int encode_tmap_bf16_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                        const uint32_t box[4]) {
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
    cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return -31;
    }
    return 0;
}

namespace {

using namespace umma;

constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }
constexpr int SMEM_EST(int nmma, int npad, int slots, int slot_bytes, int tail) { return round_up(nmma * npad * 32, 128) + slots * slot_bytes + tail + 256 + npad * 4; }
constexpr int nbuf_for(int npadl, int ctas) {        // accumulator stages: as many as 512 TMEM columns per SM allow, at most 8 / 4
    int n = (npadl <= 32 ? 8 : 4) / (ctas > 2 ? 2 : 1);
    while (n > 1 && n * npadl * ctas > 512) n /= 2;
    return n;
}
constexpr int pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// Output-pixel phases packed into the MMA N dimension.  At these channel counts an MMA costs the same ~40-48 cycles of
// shared-memory operand fetch whether N is 16 or 64, so XPH adjacent output pixels x = XPH*xq + ph share ONE pass over the
// input: column (ph, co) of the accumulator is output pixel phase ph, the x taps become kw' = ph + kw in [0, K + XPH - 1)
// (weights zero where kw' - ph is no tap), and row m of a tile is the pixel group xq.  The input slab is therefore held
// de-interleaved in shared memory -- one plane per x phase, loaded by one strided TMA map each -- so that consecutive
// rows m are again consecutive 16-byte units.  5x5, 8 -> 16 channels: 20 MMAs (N = 64) per 512 outputs instead of 60.
// A function of the layer shape only, because the weight image (b200_conv_tc_prep_weights) depends on it.
__host__ __device__ constexpr int xph_for(int cin, int cout, int ks) {
    return (ks == 5 && ((cin == 8 && cout == 16) || (cin == 16 && cout == 8))) ? 4
         : (ks == 5 && ((cin == 16 && cout == 32) || (cin == 32 && cout == 16))) ? 2
         : (ks == 3 && ((cin == 32 && cout == 64) || (cin == 64 && cout == 32))) ? 2 : 1;
}

struct TMaps {
    CUtensorMap m[4];                                           // one per x phase (residue of the global column)
};

constexpr int nbuf_chunked(int tiles, int npadl, int ctas) {     // K-chunked instances: every tile of an item owns a stage until the last chunk
    int m = 512 / (ctas * npadl * tiles);
    while (m > 1 && m * tiles > 8) --m;
    return (m < 1 ? 1 : m) * tiles;
}

template <int CIN_, int COUT_, int NPAD_, int HIN_, int WIN_, int KS_, int PAD_, int BANDS_, int SLOTS_, int CTAS_ = 1, int NSPLIT_ = 1, int KCH_ = 1>
struct TcCfg {
    static constexpr int NSPLIT = NSPLIT_;                      // output channels split over gridDim.z (halves the smem weight image)
    // K chunking (wide layers, SURVEY 8f-4's 3x3 stacks up to 256 channels): the input channel planes of an item arrive in KCH
    // slots of P / KCH planes each; the weights stay resident, the accumulators of ALL tiles of the item live in TMEM until the
    // last chunk has been multiplied in (chunk-outer, tile-inner MMA order)
    static constexpr int KCH = KCH_;
    // First layers (C_in = 1) read the "quad8" image: unit (y, xq) = the 8 padded-row pixels 4*xq .. 4*xq+7, which are exactly
    // the taps kw' = ph + kw < 8 of the four output phases -- only that one plane exists (no other phase planes to load)
    static constexpr int XPH = CIN_ == 1 ? 4 : xph_for(CIN_, COUT_, KS_);       // output x phases packed into N
    static constexpr int XPL = CIN_ == 1 ? 1 : XPH;             // phase planes of the input slab in shared memory
    // accumulator column order: (channel octet, x phase, channel in octet) -- a 16-column chunk then holds two ADJACENT output
    // pixels of one octet (XPH even), i.e. 32 contiguous bytes of the act8 output: one 256-bit store
    __host__ __device__ static constexpr int col_phase(int col) { return (col / 8) % XPH; }
    __host__ __device__ static constexpr int col_channel(int col) { return (col / 8) / XPH * 8 + col % 8; }   // within this CTA's slice
    static constexpr int NPADL = XPH == 1 ? NPAD_ / NSPLIT_ : round_up(XPH * COUT_, 16), COUTL = COUT_ / NSPLIT_;
    static_assert(NSPLIT_ == 1 || (XPH == 1 && COUT_ == NPAD_ && NPADL % 16 == 0), "N split needs C_out == NPAD and 16-channel slices");
    static constexpr int CIN = CIN_, COUT = COUT_, NPAD = NPAD_, HIN = HIN_, WIN = WIN_, KS = KS_, PAD = PAD_, BANDS = BANDS_, SLOTS = SLOTS_;
    static constexpr int CTAS = CTAS_;                          // resident CTAs per SM (independent MMA streams hide the per-UMMA fixed cost)
    static constexpr bool L0 = (CIN == 1);                      // first layer: input is the "shift8" image (unit = x[q..q+7])
    static constexpr int P = L0 ? 1 : CIN / 8;                  // channel planes of the input
    static constexpr int WP = WIN + 2 * PAD;                    // padded pitch (pixels)
    static constexpr int HO = HIN + 2 * PAD - KS + 1, WO = WP - KS + 1;
    static constexpr int HB = HO / BANDS;                       // output rows per band
    static constexpr int HPB = HB + KS - 1;                     // input rows per band slab
    static constexpr int KWX = KS + XPH - 1;                    // x taps including the phase shifts
    static constexpr int WQ = (WP + XPH - 1) / XPH;             // pitch of one phase plane = row pitch of the flat tile index
    static constexpr int WOQ = WO / XPH;                        // valid pixel groups per output row
    static constexpr int Q = (HB - 1) * WQ + WOQ;               // flat tile rows per band (incl. junk columns)
    static constexpr int TILES = (Q + 127) / 128;
    static constexpr int PLANE_BYTES = HPB * WQ * 16;           // one (x phase, channel plane)
    static constexpr int PC = P / KCH;                          // channel planes per slot
    static constexpr int PHASE_BYTES = round_up(PC * PLANE_BYTES, 128);  // stride of a phase-plane group: TMA destinations are 128-byte aligned
    static constexpr int SLOT_BYTES = round_up(XPL * PHASE_BYTES, 128);
    static constexpr int NJ = (KWX + 1) / 2;                    // kw' pairs when CIN == 8
    static constexpr int PH = P / 2;                            // plane pairs when CIN >= 16
    static constexpr int PHC = PC / 2;                          //   ... per slot
    static constexpr int NMMA = L0 ? (KS + 1) / 2 : (CIN == 8) ? KS * NJ : KS * KWX * PH;
    static constexpr int W_BYTES = round_up(NMMA * NPADL * 32, 128);
    static constexpr int MAXPIX = TILES * 128 + (L0 ? KS : KS - 1) * WQ + (KWX - 1) / XPH + 2;   // exclusive bound of units a tile may touch
    static constexpr int TAIL = round_up((MAXPIX > HPB * WQ ? (MAXPIX - HPB * WQ) : 0) * 16, 128) + 128;
    static_assert(XPH == 1 || (WO % XPH == 0 && WIN % XPH == 0 && XPH <= 4 && (CIN != 8 || (XPH % 2 == 0 && KWX % 2 == 0))), "x phases");
    static_assert(!L0 || KWX <= 8, "first layer: all taps of all phases must lie inside one 8-pixel unit");
    static constexpr int IMG_BYTES = SLOTS * SLOT_BYTES + TAIL;
    static constexpr int BAR_OFF = W_BYTES + IMG_BYTES;
    // Fused 2x2 max-pool of the forward layers (ReLU(maxpool(a z + b)) = ReLU(a ext(z) + b), ext = max or min by sign(gamma)): the
    // epilogue keeps the horizontally pooled even rows in a small shared-memory ring ("stash": POOL_R pooled rows) until the odd
    // row below them has been computed, then emits the window extreme -- a quarter of z.  A 128-row tile spans 128 / WQ image rows,
    // so the ring holds every even row a tile can touch plus the ones the warps of the NEXT tile may already be writing.
    // Offered where the horizontal pair lives in one thread (XPH even).  The single-phase 32 -> 64 layers would need 64 lane shuffles
    // per tile on top of their 128 statistics registers: measured 0.17 -> 0.28 / 0.34 ms (profiles/r2e_*), so they keep the z path.
    static constexpr bool POOL_OK = (CIN_ == 1 || COUT_ > CIN_) && NSPLIT_ == 1 && HO % 2 == 0 && HB % 2 == 0 && WO % 2 == 0 && XPH % 2 == 0;
    static constexpr bool BSTAT_OK = CIN_ > COUT_ && NSPLIT_ == 1;          // data-gradient instances (channels shrink): fused BN-backward sums
    static constexpr int POOL_R = 128 / WQ + 3;             // pooled rows of two consecutive tiles (a fast warp may be one tile phase ahead) + margin
    static constexpr int OCTL = COUTL / 8;
    static constexpr int POOL_ROW_UNITS = (WO / 2) * OCTL;       // 16-byte units (one pooled pixel of one channel octet, fp16) per pooled row
    static constexpr int POOL_BYTES = POOL_OK ? POOL_R * POOL_ROW_UNITS * 16 : 0;
    static constexpr int SIGN_OFF = BAR_OFF + 256 + NPADL * 4;   // per accumulator column pair: fp16 sign-bit masks (sign of gamma)
    static constexpr int POOL_OFF = round_up(SIGN_OFF + NPADL * 2, 16);
    static constexpr int SMEM = POOL_OFF + POOL_BYTES;           // the pooling instance
    static constexpr int SMEM_NOPOOL = SIGN_OFF + NPADL * 4;     // bias / beta [NPADL] + 1 / gamma [NPADL] (BSTAT instances)
    static constexpr int NBUF = KCH > 1 ? nbuf_chunked(TILES, NPADL, CTAS) : nbuf_for(NPADL, CTAS);          // TMEM accumulator stages
    static_assert(KCH == 1 || (CIN >= 16 && P % KCH == 0 && PC % 2 == 0 && NBUF % TILES == 0 && NBUF * NPADL * CTAS <= 512), "K chunking");
    static constexpr int TMEM_COLS = pow2_cols(NBUF * NPADL);
    // One issuing thread sustains only ~1 UMMA per 140 cycles at these tile shapes (measured, tools/umma_probe.cu); four
    // concurrent issue streams per SM -- CTAs or warps -- reach the shared-memory operand bandwidth (39-48 cycles per UMMA).
    static constexpr int ISS = CTAS >= 3 ? 1 : (CTAS == 2 ? 2 : 4);     // MMA-issuer warps per CTA
    static constexpr int THREADS = 32 * (1 + ISS + 4);
    // buffer (k * TILES + t) % NBUF must always meet the issuer t % ISS: an issuer that met a buffer only every other use could run two
    // mbarrier phases ahead and pass a parity wait on the stale phase (seen with TILES = 13, ISS = NBUF = 2)
    static_assert(TILES % ISS == 0 || NBUF % TILES == 0, "a TMEM stage must always be filled by the same issuer");
    static_assert(KCH > 1 || NBUF % ISS == 0, "a TMEM stage must always be filled by the same issuer");     // (chunked: NBUF % TILES == 0 gives the same)
    static_assert(TMEM_COLS * CTAS <= 512, "TMEM columns per SM");
    static_assert((SMEM_EST(NMMA, NPADL, SLOTS, SLOT_BYTES, TAIL) + POOL_BYTES + NPADL * 2 + 16 + 1024) * CTAS <= 227 * 1024, "shared memory per SM");
    static_assert(2 * SLOTS + 2 * NBUF <= 24, "barrier area");
    static_assert(CIN == 1 || (CIN % 8 == 0 && (CIN == 8 || CIN % 16 == 0)), "C_in must be 1, 8 or a multiple of 16");
    static_assert(NPAD % 16 == 0 && NPAD >= 16 && NPADL <= 128 && COUT <= NPAD && COUT % 8 == 0, "N tile");
    static_assert(HO % BANDS == 0, "bands must divide the output height");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static_assert(XPL * PHASE_BYTES < (1 << 18) && PLANE_BYTES % 16 == 0, "descriptor range");
};

__device__ __forceinline__ void ld_global_nc_256(const uint4* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// stash / output unit (pooled x * OCTL + octet) of the h-th pooled 16-byte value a thread holds for pixel group xq; -1: padding columns
template <class C>
__device__ __forceinline__ int pool_unit(int h, int xq) {
    if constexpr (C::XPH % 2 == 0) {
        const int col0 = h * 16;                                  // chunk h = pixels (ph0, ph0 + 1) of octet oct0
        if (col0 >= C::XPH * C::COUTL) return -1;
        const int ph0 = C::col_phase(col0), oct0 = C::col_channel(col0) / 8;
        return ((xq * C::XPH + ph0) >> 1) * C::OCTL + oct0;
    } else {
        const int oct = h;                                        // value h = octet h of the pixel pair (xq, xq + 1)
        if (oct * 8 >= C::COUTL) return -1;
        return (xq >> 1) * C::OCTL + oct;
    }
}

// out: fp32 NCHW [N][COUT][HO][WO] (out_bf16 == 0), bf16 act8 [N][COUT/8][HO][WO][8] (1) or the same in fp16 (2: the pre-BatchNorm
// z, which is never an MMA operand -- 11 mantissa bits keep max-pool arg-max ties as rare as on the reference's fp16 autocast path).
// bias may be null (no bias, no statistics); stats: double [views][COUT][2] (sum, sum of squares), accumulated.
// BSTAT (data-gradient instances): the output dx is the gradient dp of the layer BELOW's pooled activation p; the epilogue also
// reads p (same shape, bf16 act8: `pool_out` carries the pointer) and accumulates that layer's BatchNorm-backward sums
// {sum_{p>0} dp, sum_{p>0} dp * (p - beta) / gamma} into `stats` -- exactly what bn_pool8_bwd_reduce_p computes in a pass of its own
// (gamma / beta2 = that layer's BatchNorm weight / bias).
template <class C, bool POOL, bool BSTAT = false>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
conv_tc_kernel(const __grid_constant__ TMaps tmaps, const uint4* __restrict__ wprep, const float* __restrict__ bias,
               void* __restrict__ out, double* __restrict__ stats, int n_per_view, int out_bf16,
               uint4* __restrict__ pool_out, const float* __restrict__ gamma, const float* __restrict__ beta2 = nullptr) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_s = smem;
    uint8_t* img_s = smem + C::W_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);   // full[SLOTS], empty[SLOTS], tfull[NBUF], tempty[NBUF]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::BAR_OFF + 200);
    float* bias_s = reinterpret_cast<float*>(smem + C::BAR_OFF + 256);
    uint32_t* sign_s = reinterpret_cast<uint32_t*>(smem + C::SIGN_OFF);     // [NPADL / 2]: 0x8000 bits where gamma < 0 (two columns per word)
    float* ig_s = reinterpret_cast<float*>(smem + C::SIGN_OFF);             // BSTAT: 1 / gamma per column (bias_s then holds beta)
    uint4* stash = reinterpret_cast<uint4*>(smem + C::POOL_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int view = blockIdx.y, G = gridDim.x, g = blockIdx.x;
    const long items = (long)n_per_view * C::BANDS;
    const int i0 = (int)(items * g / G), i1 = (int)(items * (g + 1) / G);

    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::SLOTS + s); };
    auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * C::SLOTS + b); };
    auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * C::SLOTS + C::NBUF + b); };

    // ---- one-time setup: weights -> smem, zero the image slots (+tail), barriers, TMEM ----
    {
        const uint4* src = wprep;
        uint4* dst = reinterpret_cast<uint4*>(w_s);
        constexpr int NGL = C::NPADL / 8, NG = NGL * C::NSPLIT;       // 8-row groups (128 B = 8 uint4) per K chunk: local / whole
        for (int i = threadIdx.x; i < C::NMMA * 2 * NGL * 8; i += blockDim.x) {
            const int q = i & 7, gl = (i >> 3) % NGL, mc = (i >> 3) / NGL;
            dst[i] = src[(mc * NG + blockIdx.z * NGL + gl) * 8 + q];
        }
        uint4* z = reinterpret_cast<uint4*>(img_s);
        for (int i = threadIdx.x; i < C::IMG_BYTES / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < C::NPADL; i += blockDim.x)     // column (octet, phase, channel): the bias repeats per phase
            bias_s[i] = (bias != nullptr && i < C::XPH * C::COUTL) ? bias[blockIdx.z * C::COUTL + C::col_channel(i)] : 0.f;
        if constexpr (BSTAT) {
            for (int i = threadIdx.x; i < C::NPADL; i += blockDim.x) {
                const bool real = i < C::XPH * C::COUTL;
                const float gm = real ? gamma[blockIdx.z * C::COUTL + C::col_channel(i)] : 0.f;
                ig_s[i] = gm != 0.f ? 1.0f / gm : 0.f;
                bias_s[i] = real ? beta2[blockIdx.z * C::COUTL + C::col_channel(i)] : 0.f;
            }
        }
        if constexpr (POOL) {
            for (int i = threadIdx.x; i < C::NPADL / 2; i += blockDim.x) {
                uint32_t m = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = 2 * i + h;
                    if (gamma != nullptr && col < C::XPH * C::COUTL && gamma[blockIdx.z * C::COUTL + C::col_channel(col)] < 0.f) m |= 0x8000u << (16 * h);
                }
                sign_s[i] = m;
            }
        }
        fence_proxy_async_smem();
    }
    if (warp == 0 && lane == 0) {
        for (int r = 0; r < C::XPL; ++r) prefetch_tmap(&tmaps.m[r]);
        for (int s = 0; s < C::SLOTS; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), C::ISS);
        }
        for (int b = 0; b < C::NBUF; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t img_addr = smem_u32(img_s), w_addr = smem_u32(w_s);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = i0; i < i1; ++i) {
                const int n = view * n_per_view + i / C::BANDS, band = i % C::BANDS;
#pragma unroll 1
                for (int c = 0; c < C::KCH; ++c) {                  // one slot per channel chunk (KCH == 1: the whole item)
                    const int k = (i - i0) * C::KCH + c, slot = k % C::SLOTS, use = k / C::SLOTS;
                    mbar_wait(empty_bar(slot), (use & 1) ^ 1);
                    mbar_expect_tx(full_bar(slot), C::XPL * C::PC * C::PLANE_BYTES);
#pragma unroll
                    for (int rp = 0; rp < C::XPL; ++rp) {
                        // phase plane rp holds the padded columns x' = XPH*i + rp, i.e. the global columns x' - PAD = XPH*(i + a) + r
                        constexpr int X = C::XPH;
                        const int r = ((rp - C::PAD) % X + X) % X, a = (rp - C::PAD - r) / X;
                        tma_load_4d(img_addr + slot * C::SLOT_BYTES + rp * C::PHASE_BYTES, &tmaps.m[r], full_bar(slot), 0, C::L0 ? 0 : a,
                                    band * C::HB - C::PAD, n * C::P + c * C::PC);
                    }
                }
            }
        }
    } else if (warp <= C::ISS) {
        // ===== MMA issuers: warp w takes the tiles t == w-1 (mod ISS) of every item =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(C::NPADL);
            for (int i = i0; i < i1; ++i) {
#pragma unroll 1
              for (int c = 0; c < C::KCH; ++c) {
                const int k = (i - i0) * C::KCH + c, slot = k % C::SLOTS, use = k / C::SLOTS;
                mbar_wait(full_bar(slot), use & 1);
                tc_fence_after_sync();
                const uint32_t slot_addr = img_addr + slot * C::SLOT_BYTES;
                for (int t = warp - 1; t < C::TILES; t += C::ISS) {
                    const uint32_t tcount = (uint32_t)(i - i0) * C::TILES + t;
                    const uint32_t buf = tcount % C::NBUF, u = tcount / C::NBUF;
                    if (c == 0) {
                        mbar_wait(tempty_bar(buf), (u & 1) ^ 1);
                        tc_fence_after_sync();
                    }
                    const uint32_t d = tmem_base + buf * C::NPADL;
                    const uint32_t a0 = slot_addr + t * 2048;
                    int idx = 0;
                    if constexpr (C::L0) {
                        // K = (2 rows kh, kh+1) x (8 shifts = kw taps): one MMA per row pair, second K chunk = next image row
#pragma unroll
                        for (int j = 0; j < C::NMMA; ++j, ++idx) {         // (K + 1) / 2 row pairs
                            const uint64_t ad = smem_desc(a0 + (2 * j * C::WQ) * 16, C::WQ * 16, 128);
                            const uint64_t bd = smem_desc(w_addr + idx * C::NPADL * 32, C::NPADL * 16, 128);
                            mma_bf16(d, ad, bd, idesc, idx > 0);
                        }
                    } else {
#pragma unroll
                    for (int kh = 0; kh < C::KS; ++kh) {
                        if constexpr (C::CIN == 8) {
#pragma unroll
                            for (int j = 0; j < C::NJ; ++j, ++idx) {     // K = taps kw' = 2j, 2j+1: next unit, or the next phase plane
                                const uint64_t ad = smem_desc(a0 + ((2 * j) % C::XPH) * C::PHASE_BYTES + (kh * C::WQ + (2 * j) / C::XPH) * 16,
                                                              C::XPH == 1 ? 16 : C::PHASE_BYTES, 128);
                                const uint64_t bd = smem_desc(w_addr + idx * C::NPADL * 32, C::NPADL * 16, 128);
                                mma_bf16(d, ad, bd, idesc, idx > 0);
                            }
                        } else {
#pragma unroll
                            for (int kw = 0; kw < C::KWX; ++kw) {
#pragma unroll
                                for (int pp = 0; pp < C::PHC; ++pp, ++idx) {
                                    // weight image index of (kh, kw', plane pair c * PHC + pp); the slot holds this chunk's planes only
                                    const int widx = (kh * C::KWX + kw) * C::PH + c * C::PHC + pp;
                                    const uint64_t ad = smem_desc(a0 + (kw % C::XPH) * C::PHASE_BYTES + (kh * C::WQ + kw / C::XPH) * 16 +
                                                                      pp * 2 * C::PLANE_BYTES, C::PLANE_BYTES, 128);
                                    const uint64_t bd = smem_desc(w_addr + widx * C::NPADL * 32, C::NPADL * 16, 128);
                                    mma_bf16(d, ad, bd, idesc, (c > 0 || idx > 0) ? 1u : 0u);
                                }
                            }
                        }
                    }
                    }
                    if (c == C::KCH - 1) mma_commit(tfull_bar(buf));
                }
                mma_commit(empty_bar(slot));
              }
            }
        }
    } else {
        // ===== epilogue warps (TMEM lane quadrant = warp % 4) =====
        const int quad = warp & 3, row = quad * 32 + lane;      // the 4 epilogue warps have consecutive ids: all quadrants covered
        float s1[C::COUTL], s2[C::COUTL];
#pragma unroll
        for (int c = 0; c < C::COUTL; ++c) s1[c] = s2[c] = 0.f;
        const int co0 = blockIdx.z * C::COUTL;                  // first output channel of this CTA's slice
        const bool do_stats = BSTAT ? (stats != nullptr) : ((bias != nullptr) && (stats != nullptr));
        constexpr int NCH = C::NPADL / 16;                       // 16-column chunks of the accumulator
        constexpr int NHP = POOL ? (C::XPH % 2 == 0 ? NCH : 2 * NCH) : 1;      // pooled 16-byte units this thread may own per tile
        uint32_t tcount = 0;
        for (int i = i0; i < i1; ++i) {
            const int n = view * n_per_view + i / C::BANDS, band = i % C::BANDS;
            for (int t = 0; t < C::TILES; ++t, ++tcount) {
                [[maybe_unused]] uint4 hp[NHP];                  // horizontally pooled sign(gamma) * z of this tile row, fp16
                const uint32_t buf = tcount % C::NBUF, u = tcount / C::NBUF;
                const int q = t * 128 + row;
                const int y = q / C::WQ, xq = q - y * C::WQ;
                const bool valid = (y < C::HB) && (xq < C::WOQ);
                const int yy = band * C::HB + y;
                // BSTAT: the p values this row needs are requested BEFORE waiting for the accumulator, so the DRAM latency of the loads
                // overlaps the MMAs of the tile (issued after the wait they serialised the epilogue: 0.32 -> 1.24 ms)
                [[maybe_unused]] uint4 ppre[BSTAT ? NCH : 1][2];
                if constexpr (BSTAT) {
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        const int col0 = cc * 16;
                        ppre[cc][0] = ppre[cc][1] = make_uint4(0, 0, 0, 0);
                        if (valid && col0 < C::XPH * C::COUTL) {
                            const int ph0 = C::col_phase(col0), oct0 = C::col_channel(col0) / 8;
                            const uint4* pin = pool_out;
                            if constexpr (C::XPH % 2 == 0) {     // two adjacent pixels of one octet: 32 contiguous bytes
                                ld_global_nc_256(pin + (((long)n * (C::COUT / 8) + co0 / 8 + oct0) * C::HO + yy) * C::WO + xq * C::XPH + ph0, ppre[cc][0],
                                                 ppre[cc][1]);
                            } else {                             // two octets of one pixel
                                ppre[cc][0] = __ldg(pin + (((long)n * (C::COUT / 8) + co0 / 8 + oct0) * C::HO + yy) * C::WO + xq);
                                if ((oct0 + 1) * 8 < C::COUTL)
                                    ppre[cc][1] = __ldg(pin + (((long)n * (C::COUT / 8) + co0 / 8 + oct0 + 1) * C::HO + yy) * C::WO + xq);
                            }
                        }
                    }
                }
                mbar_wait(tfull_bar(buf), u & 1);
                tc_fence_after_sync();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * C::NPADL;
#pragma unroll
                for (int cc = 0; cc < C::NPADL / 16; ++cc) {
                    uint32_t v[16];
                    tmem_ld16(taddr + cc * 16, v);
                    tmem_ld_wait();
                    if (cc == C::NPADL / 16 - 1) {                // accumulator drained: hand the TMEM buffer back
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar(buf));
                    }
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int col = cc * 16 + j, ch = C::col_channel(col);
                        if constexpr (BSTAT) {
                            f[j] = __uint_as_float(v[j]);
                        } else {
                            f[j] = __uint_as_float(v[j]) + bias_s[col];
                            if (col < C::XPH * C::COUTL && valid) {
                                s1[ch] += f[j];
                                s2[ch] += f[j] * f[j];
                            }
                        }
                    }
                    if constexpr (BSTAT) {
                        // BatchNorm-backward sums of the layer below from (p, dp = this output, as stored: bf16-rounded)
                        const int col0 = cc * 16;
                        if (valid && col0 < C::XPH * C::COUTL) {
                            const uint4 pa = ppre[cc][0], pb = ppre[cc][1];
                            const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int col = col0 + j, ch = C::col_channel(col);
                                const float pv = __uint_as_float((j & 1) ? (pw[j >> 1] & 0xFFFF0000u) : (pw[j >> 1] << 16));
                                const float gr = __bfloat162float(__float2bfloat16_rn(f[j]));
                                if (col < C::XPH * C::COUTL && pv > 0.f) {
                                    s1[ch] += gr;
                                    s2[ch] += gr * ((pv - bias_s[col]) * ig_s[col]);
                                }
                            }
                        }
                    }
                    if constexpr (POOL) {
                        {
                            // t = sign(gamma) * z in fp16 (rounding is monotonic and symmetric: ext(round(z)) == round(ext(z))); max over the pixel pair
                            uint4 a, b;
                            a.x = pack_f16(f[0], f[1]); a.y = pack_f16(f[2], f[3]); a.z = pack_f16(f[4], f[5]); a.w = pack_f16(f[6], f[7]);
                            b.x = pack_f16(f[8], f[9]); b.y = pack_f16(f[10], f[11]); b.z = pack_f16(f[12], f[13]); b.w = pack_f16(f[14], f[15]);
                            const uint4 ma = *reinterpret_cast<const uint4*>(sign_s + cc * 8), mb = *reinterpret_cast<const uint4*>(sign_s + cc * 8 + 4);
                            a.x ^= ma.x; a.y ^= ma.y; a.z ^= ma.z; a.w ^= ma.w;
                            b.x ^= mb.x; b.y ^= mb.y; b.z ^= mb.z; b.w ^= mb.w;
                            if constexpr (C::XPH % 2 == 0) {     // the chunk = two adjacent pixels of one octet: the horizontal pair
                                hp[cc].x = hmax2_u32(a.x, b.x); hp[cc].y = hmax2_u32(a.y, b.y); hp[cc].z = hmax2_u32(a.z, b.z); hp[cc].w = hmax2_u32(a.w, b.w);
                            } else {                             // the chunk = two octets of one pixel: the horizontal partner is the next lane
                                hp[2 * cc].x = hmax2_u32(a.x, __shfl_xor_sync(0xffffffffu, a.x, 1)); hp[2 * cc].y = hmax2_u32(a.y, __shfl_xor_sync(0xffffffffu, a.y, 1));
                                hp[2 * cc].z = hmax2_u32(a.z, __shfl_xor_sync(0xffffffffu, a.z, 1)); hp[2 * cc].w = hmax2_u32(a.w, __shfl_xor_sync(0xffffffffu, a.w, 1));
                                hp[2 * cc + 1].x = hmax2_u32(b.x, __shfl_xor_sync(0xffffffffu, b.x, 1)); hp[2 * cc + 1].y = hmax2_u32(b.y, __shfl_xor_sync(0xffffffffu, b.y, 1));
                                hp[2 * cc + 1].z = hmax2_u32(b.z, __shfl_xor_sync(0xffffffffu, b.z, 1)); hp[2 * cc + 1].w = hmax2_u32(b.w, __shfl_xor_sync(0xffffffffu, b.w, 1));
                            }
                        }
                    }
                    if (valid && (!POOL || out != nullptr)) {
                        if (out_bf16) {
                            uint4 pk[2];
#pragma unroll
                            for (int o = 0; o < 2; ++o) {
                                if (out_bf16 == 2) {
                                    pk[o].x = pack_f16(f[o * 8 + 0], f[o * 8 + 1]);
                                    pk[o].y = pack_f16(f[o * 8 + 2], f[o * 8 + 3]);
                                    pk[o].z = pack_f16(f[o * 8 + 4], f[o * 8 + 5]);
                                    pk[o].w = pack_f16(f[o * 8 + 6], f[o * 8 + 7]);
                                } else {
                                    pk[o].x = pack_bf16(f[o * 8 + 0], f[o * 8 + 1]);
                                    pk[o].y = pack_bf16(f[o * 8 + 2], f[o * 8 + 3]);
                                    pk[o].z = pack_bf16(f[o * 8 + 4], f[o * 8 + 5]);
                                    pk[o].w = pack_bf16(f[o * 8 + 6], f[o * 8 + 7]);
                                }
                            }
                            const int col0 = cc * 16, ph0 = C::col_phase(col0), oct0 = C::col_channel(col0) / 8;
                            if constexpr (C::XPH % 2 == 0) {
                                // the chunk = pixels (ph0, ph0 + 1) of octet oct0: 32 contiguous, 32-byte aligned bytes
                                if (col0 < C::XPH * C::COUTL) {
                                    uint4* dst = reinterpret_cast<uint4*>(out) + (((long)n * (C::COUT / 8) + co0 / 8 + oct0) * C::HO + yy) * C::WO +
                                                 xq * C::XPH + ph0;
                                    st_global_256(dst, pk[0], pk[1]);
                                }
                            } else {
#pragma unroll
                                for (int o = 0; o < 2; ++o) {        // XPH == 1: two octets of the same pixel
                                    const int oct = oct0 + o;
                                    if (oct * 8 < C::COUTL) {
                                        uint4* dst = reinterpret_cast<uint4*>(out) + (((long)n * (C::COUT / 8) + co0 / 8 + oct) * C::HO + yy) * C::WO + xq;
                                        *dst = pk[o];
                                    }
                                }
                            }
                        } else {
                            float* dst = reinterpret_cast<float*>(out) + (((long)n * C::COUT + co0) * C::HO + yy) * C::WO + xq * C::XPH;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int col = cc * 16 + j, ph = C::col_phase(col), ch = C::col_channel(col);
                                if (col < C::XPH * C::COUTL) dst[(long)ch * C::HO * C::WO + ph] = f[j];
                            }
                        }
                    }
                }
                if constexpr (POOL) {
                    {
                        // vertical half of the window: even rows park their pooled values in the ring, the odd row below combines and emits
                        const int slot = (int)(((long)(i - i0) * (C::HB / 2) + (y >> 1)) % C::POOL_R);
                        const bool own = valid && (C::XPH % 2 == 0 || (xq & 1) == 0);
                        uint4* srow = stash + slot * C::POOL_ROW_UNITS;
                        if (own && (y & 1) == 0) {
#pragma unroll
                            for (int h = 0; h < NHP; ++h) {
                                const int u = pool_unit<C>(h, xq);
                                if (u >= 0) srow[u] = hp[h];
                            }
                        }
                        asm volatile("bar.sync 2, 128;" ::: "memory");        // the four epilogue warps
                        if (own && (y & 1) == 1) {
                            const int ypg = (band * C::HB + y) >> 1;
                            uint4 ev[NHP];
#pragma unroll
                            for (int h = 0; h < NHP; ++h) {
                                const int u = pool_unit<C>(h, xq);
                                if (u >= 0) {
                                    const uint4 up = srow[u];
                                    const int col0 = (C::XPH % 2 == 0 ? h : h / 2) * 16 + (C::XPH % 2 == 0 ? 0 : (h & 1) * 8);
                                    const uint4 m = *reinterpret_cast<const uint4*>(sign_s + col0 / 2);
                                    ev[h].x = hmax2_u32(hp[h].x, up.x) ^ m.x; ev[h].y = hmax2_u32(hp[h].y, up.y) ^ m.y;
                                    ev[h].z = hmax2_u32(hp[h].z, up.z) ^ m.z; ev[h].w = hmax2_u32(hp[h].w, up.w) ^ m.w;
                                }
                            }
                            const long plane = (long)(C::HO / 2) * (C::WO / 2);
                            uint4* erow = pool_out + ((long)n * (C::COUT / 8) + co0 / 8) * plane + (long)ypg * (C::WO / 2);
                            if constexpr (C::XPH == 4) {
                                // values (2m, 2m+1) = the two pooled pixels 2 xq, 2 xq + 1 of octet m: 32 contiguous, aligned bytes
#pragma unroll
                                for (int h = 0; h + 1 < NHP; h += 2) {
                                    const int u = pool_unit<C>(h, xq);
                                    if (u >= 0) {
                                        const int xp = u / C::OCTL, oct = u - xp * C::OCTL;
                                        st_global_256(erow + oct * plane + xp, ev[h], ev[h + 1]);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int h = 0; h < NHP; ++h) {
                                    const int u = pool_unit<C>(h, xq);
                                    if (u >= 0) {
                                        const int xp = u / C::OCTL, oct = u - xp * C::OCTL;
                                        erow[oct * plane + xp] = ev[h];
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        if (do_stats) {
#pragma unroll
            for (int c = 0; c < C::COUTL; ++c) {
                const double a = warp_sum((double)s1[c]), b = warp_sum((double)s2[c]);
                if (lane == 0) {
                    atomicAdd(&stats[((long)view * C::COUT + co0 + c) * 2 + 0], a);
                    atomicAdd(&stats[((long)view * C::COUT + co0 + c) * 2 + 1], b);
                }
            }
        }
    }
    // ---- teardown ----
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

// fp32 OIHW weights -> the bf16 byte image the MMA issuer expects: [mma][k chunk (2)][n group][8 rows][8 k].
// flip = 1 prepares the data-gradient convolution: w is the forward weight [CIN][COUT][KS][KS] and the taps are mirrored.
__device__ __forceinline__ void conv_tc_prep_element(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int CIN, int COUT, int NPAD,
                                                     int KS, int XPH, int flip, int e) {
    const int k8 = e & 7, r = (e >> 3) & 7, NG = NPAD / 8;
    const int ng = (e >> 6) % NG, c = (e / (64 * NG)) & 1, m = e / (128 * NG);
    const int ph = ng % XPH, co = ng / XPH * 8 + r;                     // accumulator column = (channel octet, x phase, channel in octet)
    const int KWX = KS + XPH - 1;
    int kh, kw, ci;
    if (CIN == 1) {                 // quad8 first layer: chunk c of MMA m is image row kh = 2m + c, k8 is the tap kw' = ph + kw
        kh = 2 * m + c;
        kw = k8 - ph;
        ci = 0;
        float v1 = 0.f;
        if (kh < KS && kw >= 0 && kw < KS && co < COUT) v1 = w[(co * KS + kh) * KS + kw];
        out[e] = __float2bfloat16_rn(v1);
        return;
    } else if (CIN == 8) {
        const int NJ = (KWX + 1) / 2;
        kh = m / NJ;
        kw = 2 * (m % NJ) + c - ph;
        ci = k8;
    } else {
        const int PH = CIN / 16;
        kh = m / (KWX * PH);
        kw = (m / PH) % KWX - ph;
        ci = (2 * (m % PH) + c) * 8 + k8;
    }
    float v = 0.f;
    if (kw >= 0 && kw < KS && co < COUT) {
        v = flip ? w[((ci * COUT + co) * KS + (KS - 1 - kh)) * KS + (KS - 1 - kw)] : w[((co * CIN + ci) * KS + kh) * KS + kw];
    }
    out[e] = __float2bfloat16_rn(v);
}

__global__ void conv_tc_prep_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int CIN, int COUT, int NPAD,
                                            int KS, int XPH, int flip, int total) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total) conv_tc_prep_element(w, out, CIN, COUT, NPAD, KS, XPH, flip, e);
}

__host__ __device__ inline void conv_tc_weight_geometry(int Cin, int Cout, int K, int* xph, int* npad, int* total) {
    const int x = Cin == 1 ? 4 : xph_for(Cin, Cout, K), kwx = K + x - 1;
    const int np = (x * Cout + 15) / 16 * 16;
    const int nmma = (Cin == 1) ? (K + 1) / 2 : (Cin == 8) ? K * ((kwx + 1) / 2) : K * kwx * (Cin / 16);
    *xph = x;
    *npad = np;
    *total = nmma * np * 16;                  // bf16 elements
}

// all weight images of a step in ONE launch: desc[i] = {w pointer, out pointer, Cin, Cout, K, flip} (int64 each), blockIdx.y = i
__global__ void conv_tc_prep_weights_multi_kernel(const long long* __restrict__ desc) {
    const long long* d = desc + 6 * blockIdx.y;
    const int Cin = (int)d[2], Cout = (int)d[3], K = (int)d[4], flip = (int)d[5];
    int xph, npad, total;
    conv_tc_weight_geometry(Cin, Cout, K, &xph, &npad, &total);
    const float* w = reinterpret_cast<const float*>(d[0]);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d[1]);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x)
        conv_tc_prep_element(w, out, Cin, Cout, npad, K, xph, flip, e);
}

// fp32 NCHW -> bf16 act8 (tests and the hand-over from the fp32 first layer)
__global__ void pack_act8_kernel(const float* __restrict__ x, uint4* __restrict__ out, long n_units, int C, int HW) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;     // one 16-byte unit: (n, octet, pixel)
    if (i >= n_units) return;
    const int pix = (int)(i % HW);
    const long no = i / HW;
    const int oct = (int)(no % (C / 8));
    const long n = no / (C / 8);
    const float* s = x + ((n * C + oct * 8) * (long)HW) + pix;
    uint4 pk;
    pk.x = pack_bf16(s[0], s[(long)HW]);
    pk.y = pack_bf16(s[2L * HW], s[3L * HW]);
    pk.z = pack_bf16(s[4L * HW], s[5L * HW]);
    pk.w = pack_bf16(s[6L * HW], s[7L * HW]);
    out[i] = pk;
}

// fp32 [N][H][W] -> bf16 shift8 [N][H][W+pad][8]: unit (y, xs) = x[y][xs-pad .. xs-pad+7] (zero outside the row): the
// 16-byte "pixel" of the first-layer tensor-core convolutions, whose kw taps are the K dimension.  The left padding
// columns are materialised (their units reach into the row); right / top / bottom padding is TMA zero fill.
__global__ void pack_shift8_kernel(const float* __restrict__ x, uint4* __restrict__ out, long n_units, int W, int pad) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_units) return;
    const int WT = W + pad;
    const int xs = (int)(i % WT);
    const long row = i / WT;
    const float* r = x + row * W;
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int xc = xs - pad + c;
        f[c] = (xc >= 0 && xc < W) ? __ldg(r + xc) : 0.f;
    }
    out[i] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// fp32 [N][H][W] -> bf16 quad8 [N][H][WQ][8], WQ = ceil((W + 2 pad) / 4): unit (y, xq) = padded-row pixels 4*xq .. 4*xq+7
// (padded column c is image column c - pad, zero outside the row).  Each pixel appears in two units: 4 bytes per pixel.
__global__ void pack_quad8_kernel(const float* __restrict__ x, uint4* __restrict__ out, long n_units, int W, int WQ, int pad) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_units) return;
    const int xq = (int)(i % WQ);
    const long row = i / WQ;
    const float* r = x + row * W;
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int xc = 4 * xq - pad + c;
        f[c] = (xc >= 0 && xc < W) ? __ldg(r + xc) : 0.f;
    }
    out[i] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

template <class C>
int launch_conv_tc(const void* x, const void* wprep, const float* bias, void* out, double* stats, int N, int n_per_view, int out_bf16,
                   cudaStream_t st, void* pool_out = nullptr, const float* gamma = nullptr, const float* beta2 = nullptr) {
    if (pool_out != nullptr && beta2 == nullptr && !C::POOL_OK) {
        set_error("conv_tc: this geometry has no fused max-pool epilogue");
        return -5;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_NOPOOL);
        if (e == cudaSuccess && C::POOL_OK) e = cudaFuncSetAttribute(conv_tc_kernel<C, C::POOL_OK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e == cudaSuccess && C::BSTAT_OK)
            e = cudaFuncSetAttribute(conv_tc_kernel<C, false, C::BSTAT_OK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_NOPOOL);
        if (e != cudaSuccess) {
            set_error("conv_tc: cannot set %d bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    TMaps tm;
    constexpr uint64_t WT = C::L0 ? C::WQ : C::WIN;                // quad8 image: WQ units per row, left padding materialised
    for (int r = 0; r < C::XPL; ++r) {                             // map r: the columns x = XPH*i + r of every row
        const uint64_t dims[4] = {8, WT / C::XPL, (uint64_t)C::HIN, (uint64_t)N * C::P};
        const uint64_t strides[3] = {16 * (uint64_t)C::XPL, WT * 16, WT * C::HIN * 16};
        const uint32_t box[4] = {8, (uint32_t)C::WQ, (uint32_t)C::HPB, (uint32_t)C::PC};
        int rc = encode_tmap_bf16_4d(&tm.m[r], reinterpret_cast<const uint8_t*>(x) + 16 * r, dims, strides, box);
        if (rc) return rc;
    }
    for (int r = C::XPL; r < 4; ++r) tm.m[r] = tm.m[0];
    const int views = N / n_per_view;
    int G = sm_count() * C::CTAS / (views * C::NSPLIT);
    if (G < 1) G = 1;
    const long items = (long)n_per_view * C::BANDS;
    if (G > items) G = (int)items;
    if (beta2 != nullptr) {
        if (!C::BSTAT_OK) {
            set_error("conv_tc: this geometry has no fused BatchNorm-backward statistics");
            return -5;
        }
        conv_tc_kernel<C, false, C::BSTAT_OK><<<dim3(G, views, C::NSPLIT), C::THREADS, C::SMEM_NOPOOL, st>>>(
            tm, reinterpret_cast<const uint4*>(wprep), nullptr, out, stats, n_per_view, out_bf16, reinterpret_cast<uint4*>(pool_out), gamma, beta2);
    } else if (pool_out != nullptr)
        conv_tc_kernel<C, C::POOL_OK><<<dim3(G, views, C::NSPLIT), C::THREADS, C::SMEM, st>>>(tm, reinterpret_cast<const uint4*>(wprep), bias, out, stats,
                                                                                              n_per_view, out_bf16, reinterpret_cast<uint4*>(pool_out), gamma);
    else
        conv_tc_kernel<C, false><<<dim3(G, views, C::NSPLIT), C::THREADS, C::SMEM_NOPOOL, st>>>(tm, reinterpret_cast<const uint4*>(wprep), bias, out, stats,
                                                                                                n_per_view, out_bf16, nullptr, nullptr);
    return launch_status("conv_tc_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient: dW[co][ci][kh][kw] = sum over samples and pixels of dz[co][q] * x[ci][q + kh*WP + kw].  The reduction
// (GEMM K) dimension is the flat pixel index, so BOTH operands are MN-major views of act8 smem images (a 16-byte unit = 8
// channels of one pixel; 8 consecutive pixels = the 8 K rows of a core matrix, LBO = 128 B to the next 8 pixels):
//   A (M = 64): the x image of one channel plane; the 8 M units are 8 pixel shifts kw' = 0..7 (SBO = 16 B), start address
//               moved by kh*WP pixels -> rows (kw', ci) of the accumulator, kw' < K are the real taps;
//   B (N = C_out): the dz image, N units = channel planes (SBO = plane stride); dz is loaded with the PADDED pitch, its
//               junk columns are TMA zero fill, so junk pixels contribute nothing.
// One TMEM accumulator [64 x C_out] per (kh, x plane) lives for the whole kernel.  Per-CTA partials go to `work`, reduced
// in a fixed order by wgrad_reduce_kernel.  (The bias gradient sum(dz) is produced by the BatchNorm-backward kernel that
// writes dz.)  First layers (C_in = 1) read the shift8 image: the 8 M units are the image rows kh (SBO = one row) and the 8
// elements of a unit are the kw shifts, so ONE accumulator holds all K*K taps.
constexpr int hbz_for(int hb, int wp) {
    int h = hb;
    while ((h * wp) % 16) ++h;
    return h;
}

template <int CIN_, int COUT_, int HIN_, int WIN_, int KS_, int PAD_, int BANDS_, int SLOTS_, int PSPLIT_, int CTAS_ = 1, int NSPLIT_ = 1>
struct TcWgCfg {
    static constexpr int NSPLIT = NSPLIT_, COUTL = COUT_ / NSPLIT_;     // dz channel planes split over CTAs (fewer TMEM columns per CTA)
    static constexpr int CTAS = CTAS_;                           // resident CTAs per SM
    static constexpr int CIN = CIN_, COUT = COUT_, HIN = HIN_, WIN = WIN_, KS = KS_, PAD = PAD_, BANDS = BANDS_, SLOTS = SLOTS_, PSPLIT = PSPLIT_;
    static constexpr bool L0 = (CIN == 1);                       // first layer over the shift8 image: rows = (kh, kw) in ONE accumulator
    static constexpr int P_IN = L0 ? 1 : CIN / 8, P_OUT = COUT / 8, P_OUTL = COUTL / 8, PI = P_IN / PSPLIT;
    static constexpr int WP = WIN + 2 * PAD;
    static constexpr int HO = HIN + 2 * PAD - KS + 1, WO = WP - KS + 1;
    static constexpr int HB = HO / BANDS, HPB = HB + KS - 1;
    static constexpr int HBZ = hbz_for(HB, WP);
    static constexpr int KSTEPS = HBZ * WP / 16;
    static constexpr int PLANE_X = HPB * WP * 16, PLANE_Z = HBZ * WP * 16;
    static constexpr int X_BYTES = round_up(PI * PLANE_X, 128), Z_BYTES = round_up(P_OUTL * PLANE_Z, 128);
    static constexpr int SLOT_BYTES = X_BYTES + Z_BYTES;
    static constexpr int NACC = L0 ? 1 : KS * PI;
    static constexpr int TMEM_COLS = pow2_cols(NACC * COUTL);
    static constexpr int ONES_OFF = SLOTS * SLOT_BYTES;
    static constexpr int BAR_OFF = ONES_OFF;
    static constexpr int SMEM = BAR_OFF + 256;
    static constexpr int DW = COUT * CIN * KS * KS;
    static constexpr int PART = DW;                                 // floats per CTA partial
    static_assert(P_IN % PSPLIT == 0 && HO % BANDS == 0, "splits");
    static_assert(BANDS == 1 || HBZ == HB, "row bands need HB*WP to be a multiple of 16");
    // MMA-issuer warps; the accumulators (kh, x plane) are dealt out among them, so the count divides NACC where it can: 5
    // accumulators on 4 issuers would leave three of them idle half of the time
    static constexpr int ISS = (L0 || NACC < 4 || CTAS >= 3) ? 1 : (NACC <= 6 ? NACC : 4);
    static constexpr int THREADS = 32 * (1 + ISS + 4);
    static_assert(NACC * COUTL <= 512 && TMEM_COLS * CTAS <= 512 && COUTL % 8 == 0, "TMEM columns");
    static_assert((SMEM + 1024) * CTAS <= 227 * 1024, "shared memory per SM");
    static_assert(KS <= 8 && COUT % 8 == 0 && COUT >= 8 && COUT <= 256, "shape");
    static_assert((KSTEPS * 16 + (L0 ? 7 : KS - 1) * WP + 8 - HPB * WP) * 16 <= Z_BYTES, "x overrun must stay inside the slot");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_z, float* __restrict__ work, int N) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);     // full[SLOTS], empty[SLOTS], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::BAR_OFF + 192);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, g = blockIdx.x, split = blockIdx.z % C::PSPLIT, ns = blockIdx.z / C::PSPLIT;
    const long items = (long)N * C::BANDS;
    const int i0 = (int)(items * g / G), i1 = (int)(items * (g + 1) / G);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::SLOTS + s); };
    const uint32_t done_bar = bar0 + 8u * (2 * C::SLOTS);
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < C::ONES_OFF / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_z);
        for (int s = 0; s < C::SLOTS; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), C::ISS);
        }
        mbar_init(done_bar, C::ISS);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);

    if (warp == 0) {
        if (lane == 0) {
            for (int i = i0; i < i1; ++i) {
                const int k = i - i0, slot = k % C::SLOTS, use = k / C::SLOTS;
                mbar_wait(empty_bar(slot), (use & 1) ^ 1);
                mbar_expect_tx(full_bar(slot), C::PI * C::PLANE_X + C::P_OUTL * C::PLANE_Z);
                const int n = i / C::BANDS, band = i % C::BANDS;
                const uint32_t sa = smem0 + slot * C::SLOT_BYTES;
                tma_load_4d(sa, &tmap_x, full_bar(slot), 0, C::L0 ? 0 : -C::PAD, band * C::HB - C::PAD, n * C::P_IN + split * C::PI);
                tma_load_4d(sa + C::X_BYTES, &tmap_z, full_bar(slot), 0, 0, band * C::HB, n * C::P_OUT + ns * C::P_OUTL);
            }
        }
    } else if (warp <= C::ISS) {
        if (lane == 0) {                     // issuer w owns the accumulators a == w-1 (mod ISS)
            constexpr uint32_t idesc = idesc_bf16(C::COUTL, true, true, 64);
            for (int i = i0; i < i1; ++i) {
                const int k = i - i0, slot = k % C::SLOTS, use = k / C::SLOTS;
                mbar_wait(full_bar(slot), use & 1);
                tc_fence_after_sync();
                const uint32_t xa = smem0 + slot * C::SLOT_BYTES, za = xa + C::X_BYTES;
                for (int ks = 0; ks < C::KSTEPS; ++ks) {
                    const uint32_t acc = (i > i0 || ks > 0) ? 1u : 0u;
                    const uint64_t bd = smem_desc(za + ks * 256, 128, C::PLANE_Z);
                    if constexpr (C::L0) {
                        // M units = 8 image rows kh (SBO = one row), the 8 elements of a unit = the kw shifts
                        const uint64_t ad = smem_desc(xa + ks * 256, 128, C::WP * 16);
                        mma_bf16(tmem_base, ad, bd, idesc, acc);
                    } else {
#pragma unroll
                        for (int kh = 0; kh < C::KS; ++kh) {
#pragma unroll
                            for (int pl = 0; pl < C::PI; ++pl) {
                                if ((kh * C::PI + pl) % C::ISS != warp - 1) continue;
                                const uint64_t ad = smem_desc(xa + pl * C::PLANE_X + (ks * 16 + kh * C::WP) * 16, 128, 16);
                                mma_bf16(tmem_base + (kh * C::PI + pl) * C::COUTL, ad, bd, idesc, acc);
                            }
                        }
                    }
                }
                mma_commit(empty_bar(slot));
            }
            mma_commit(done_bar);
        }
    } else {
        // epilogue: M = 64 accumulators occupy lanes 0-15 of each 32-lane quadrant: row m = quad*16 + lane
        const int quad = warp & 3;
        const int m = quad * 16 + (lane & 15), j = m >> 3, ci8 = m & 7;
        const bool rowok = (lane < 16) && (j < C::KS);
        float* part = work + (long)g * C::PART;
        if (i1 > i0) {
            mbar_wait(done_bar, 0);
            tc_fence_after_sync();
        }
#pragma unroll 1
        for (int a = 0; a < C::NACC; ++a) {
            const int kh = a / C::PI, pl = a % C::PI;
#pragma unroll
            for (int cc = 0; cc < (C::COUTL + 15) / 16; ++cc) {
                uint32_t v[16];
                if (i1 > i0) {
                    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + a * C::COUTL + cc * 16, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int t = 0; t < 16; ++t) v[t] = 0u;
                }
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int co = ns * C::COUTL + cc * 16 + t;
                    if (cc * 16 + t < C::COUTL) {
                        if constexpr (C::L0) {      // row m = (kh = j, kw = ci8)
                            if (lane < 16 && j < C::KS && ci8 < C::KS) part[(co * C::KS + j) * C::KS + ci8] = __uint_as_float(v[t]);
                        } else if (rowok) {
                            const int ci = (split * C::PI + pl) * 8 + ci8;
                            part[((co * C::CIN + ci) * C::KS + kh) * C::KS + j] = __uint_as_float(v[t]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

// dw[e] = sum over CTA partials in a fixed order (deterministic)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ work, int n_parts, int part, float* __restrict__ dw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= part) return;
    float acc = 0.f;
    for (int p = 0; p < n_parts; ++p) acc += work[(long)p * part + e];
    dw[e] = acc;
}

template <class C>
int wgrad_ctas(int N) {
    int G = sm_count() * C::CTAS / (C::PSPLIT * C::NSPLIT);
    const long items = (long)N * C::BANDS;
    if (G > items) G = (int)items;
    return G < 1 ? 1 : G;
}

template <class C>
int launch_conv_tc_wgrad(const void* x, const void* dz, float* dw, float* work, int N, cudaStream_t st, int64_t* need) {
    const int G = wgrad_ctas<C>(N);
    if (need) {
        *need = (int64_t)G * C::PART;
        return 0;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            set_error("conv_tc_wgrad: cannot set %d bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    CUtensorMap tx, tz;
    {
        constexpr uint64_t WT = C::L0 ? C::WIN + C::PAD : C::WIN;
        const uint64_t dims[4] = {8, WT, (uint64_t)C::HIN, (uint64_t)N * C::P_IN};
        const uint64_t strides[3] = {16, WT * 16, WT * C::HIN * 16};
        const uint32_t box[4] = {8, (uint32_t)C::WP, (uint32_t)C::HPB, (uint32_t)C::PI};
        int rc = encode_tmap_bf16_4d(&tx, x, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[4] = {8, (uint64_t)C::WO, (uint64_t)C::HO, (uint64_t)N * C::P_OUT};
        const uint64_t strides[3] = {16, (uint64_t)C::WO * 16, (uint64_t)C::WO * C::HO * 16};
        const uint32_t box[4] = {8, (uint32_t)C::WP, (uint32_t)C::HBZ, (uint32_t)C::P_OUTL};
        int rc = encode_tmap_bf16_4d(&tz, dz, dims, strides, box);
        if (rc) return rc;
    }
    conv_tc_wgrad_kernel<C><<<dim3(G, 1, C::PSPLIT * C::NSPLIT), C::THREADS, C::SMEM, st>>>(tx, tz, work, N);
    int rc = launch_status("conv_tc_wgrad_kernel");
    if (rc) return rc;
    wgrad_reduce_kernel<<<(C::PART + 255) / 256, 256, 0, st>>>(work, G, C::PART, dw);
    return launch_status("wgrad_reduce_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient, one accumulator per filter TAP (wide 3x3 layers of the simple encoders, models/dino.py:20-66).  The
// (8 pixel shifts x 8 channels) M rows of conv_tc_wgrad_kernel use K of 8 shifts: a 3x3 filter wastes 5/8 of every MMA.
// Here the operand roles are swapped and the tap is a start-address shift of the x operand:
//   A (M = 64 / 128 output channels): the dz image, MN-major, M units = channel planes (SBO = plane stride);
//   B (N = 8 * PI input channels):    the x slab, MN-major, N units = channel planes, start moved by (kh*WP + kw) pixels;
//   D_tap[co][ci] += sum over K = flat padded pixel index   =>   dW[co][ci][kh][kw] = D_(kh,kw)[co][ci]
// -- every MMA row and column is a real weight.  KS*KS accumulators of N columns live in TMEM for the whole kernel (9 x 32 = 288
// columns).  dz is loaded PLANE BY PLANE (one TMA each) at a plane stride rounded up to whole K steps, and the gap stays zero, so
// that row bands need no alignment between HB*WP and the K step (the x positions a gap multiplies are finite bf16 values of the slot).
template <int CIN_, int COUT_, int HIN_, int WIN_, int KS_, int PAD_, int BANDS_, int SLOTS_, int PSPLIT_, int NSPLIT_, int ISS_ = 3>
struct TapWgCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, HIN = HIN_, WIN = WIN_, KS = KS_, PAD = PAD_, BANDS = BANDS_, SLOTS = SLOTS_;
    static constexpr int PSPLIT = PSPLIT_, NSPLIT = NSPLIT_, ISS = ISS_;
    static constexpr int P_IN = CIN / 8, PI = P_IN / PSPLIT, P_OUT = COUT / 8, P_OUTL = P_OUT / NSPLIT;
    static constexpr int M = P_OUTL * 8, N = PI * 8, TAPS = KS * KS;
    static constexpr int WP = WIN + 2 * PAD, HO = HIN + 2 * PAD - KS + 1, WO = WP - KS + 1;
    static constexpr int HB = HO / BANDS, HPB = HB + KS - 1;
    static constexpr int KPOS = HB * WP, KSTEPS = (KPOS + 15) / 16;
    static constexpr int PLANE_X = HPB * WP * 16, PLANE_Z = KSTEPS * 256, Z_LOAD = KPOS * 16;   // bytes; Z_LOAD = what TMA writes per dz plane
    static constexpr int X_BYTES = round_up(PI * PLANE_X, 128), Z_BYTES = P_OUTL * PLANE_Z;
    static constexpr int SLOT_BYTES = X_BYTES + Z_BYTES;
    static constexpr int TMEM_COLS = pow2_cols(TAPS * N);
    static constexpr int BAR_OFF = SLOTS * SLOT_BYTES, SMEM = BAR_OFF + 256;
    static constexpr int THREADS = 32 * (1 + ISS + 4);
    static constexpr int PART = COUT * CIN * KS * KS;
    static_assert(CIN % 8 == 0 && COUT % 8 == 0 && P_IN % PSPLIT == 0 && P_OUT % NSPLIT == 0 && HO % BANDS == 0, "splits");
    static_assert((M == 64 || M == 128) && N % 16 == 0 && N >= 16 && N <= 256 && TAPS * N <= 512, "MMA shape / TMEM columns");
    // the last K step of the last tap reads x up to position KSTEPS*16 - 1 + (KS-1)*WP + KS-1 of a plane: past the slab it must land in dz
    static_assert((KSTEPS * 16 + (KS - 1) * WP + KS - 1 - HPB * WP) * 16 <= Z_BYTES, "x overrun must stay inside the slot");
    static_assert(SMEM + 1024 <= 227 * 1024 && 2 * SLOTS + 1 <= 24, "shared memory / barriers");
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
conv_tc_wgrad_tap_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_z, float* __restrict__ work, int N) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);     // full[SLOTS], empty[SLOTS], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::BAR_OFF + 200);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, g = blockIdx.x, ps = blockIdx.z % C::PSPLIT, ns = blockIdx.z / C::PSPLIT;
    const long items = (long)N * C::BANDS;
    const int i0 = (int)(items * g / G), i1 = (int)(items * (g + 1) / G);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::SLOTS + s); };
    const uint32_t done_bar = bar0 + 8u * (2 * C::SLOTS);
    {
        uint4* z = reinterpret_cast<uint4*>(smem);                      // incl. the gaps behind the dz planes, which must stay zero
        for (int i = threadIdx.x; i < C::BAR_OFF / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_z);
        for (int s = 0; s < C::SLOTS; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), C::ISS);
        }
        mbar_init(done_bar, C::ISS);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);

    if (warp == 0) {
        if (lane == 0) {
            for (int i = i0; i < i1; ++i) {
                const int k = i - i0, slot = k % C::SLOTS, use = k / C::SLOTS;
                mbar_wait(empty_bar(slot), (use & 1) ^ 1);
                mbar_expect_tx(full_bar(slot), C::PI * C::PLANE_X + C::P_OUTL * C::Z_LOAD);
                const int n = i / C::BANDS, band = i % C::BANDS;
                const uint32_t sa = smem0 + slot * C::SLOT_BYTES;
                tma_load_4d(sa, &tmap_x, full_bar(slot), 0, -C::PAD, band * C::HB - C::PAD, n * C::P_IN + ps * C::PI);
#pragma unroll 1
                for (int p = 0; p < C::P_OUTL; ++p)
                    tma_load_4d(sa + C::X_BYTES + p * C::PLANE_Z, &tmap_z, full_bar(slot), 0, 0, band * C::HB, n * C::P_OUT + ns * C::P_OUTL + p);
            }
        }
    } else if (warp <= C::ISS) {
        if (lane == 0) {                     // issuer w owns the taps t == w-1 (mod ISS)
            constexpr uint32_t idesc = idesc_bf16(C::N, true, true, C::M);
            for (int i = i0; i < i1; ++i) {
                const int k = i - i0, slot = k % C::SLOTS, use = k / C::SLOTS;
                mbar_wait(full_bar(slot), use & 1);
                tc_fence_after_sync();
                const uint32_t xa = smem0 + slot * C::SLOT_BYTES, za = xa + C::X_BYTES;
#pragma unroll 1
                for (int ks = 0; ks < C::KSTEPS; ++ks) {
                    const uint32_t acc = (i > i0 || ks > 0) ? 1u : 0u;
                    const uint64_t ad = smem_desc(za + ks * 256, 128, C::PLANE_Z);
#pragma unroll
                    for (int t = 0; t < C::TAPS; ++t) {
                        if (t % C::ISS != warp - 1) continue;
                        const uint64_t bd = smem_desc(xa + (ks * 16 + (t / C::KS) * C::WP + (t % C::KS)) * 16, 128, C::PLANE_X);
                        mma_bf16(tmem_base + t * C::N, ad, bd, idesc, acc);
                    }
                }
                mma_commit(empty_bar(slot));
            }
            mma_commit(done_bar);
        }
    } else {
        // epilogue: M = 128: lane = row of the quadrant; M = 64: rows occupy lanes 0-15 of each 32-lane quadrant
        const int quad = warp & 3;
        const int row = C::M == 128 ? quad * 32 + lane : quad * 16 + (lane & 15);
        const bool rowok = C::M == 128 || lane < 16;
        float* part = work + (long)g * C::PART;
        if (i1 > i0) {
            mbar_wait(done_bar, 0);
            tc_fence_after_sync();
        }
        const int co = ns * C::M + row;
#pragma unroll 1
        for (int t = 0; t < C::TAPS; ++t) {
#pragma unroll
            for (int cc = 0; cc < C::N / 16; ++cc) {
                uint32_t v[16];
                if (i1 > i0) {
                    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + t * C::N + cc * 16, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0u;
                }
                if (rowok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int ci = ps * C::N + cc * 16 + j;
                        part[((long)co * C::CIN + ci) * C::TAPS + t] = __uint_as_float(v[j]);
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

template <class C>
int launch_conv_tc_wgrad_tap(const void* x, const void* dz, float* dw, float* work, int N, cudaStream_t st, int64_t* need) {
    int G = sm_count() / (C::PSPLIT * C::NSPLIT);
    const long items = (long)N * C::BANDS;
    if (G > items) G = (int)items;
    if (G < 1) G = 1;
    if (need) {
        *need = (int64_t)G * C::PART;
        return 0;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_tap_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            set_error("conv_tc_wgrad (tap): cannot set %d bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    CUtensorMap tx, tz;
    {
        const uint64_t dims[4] = {8, (uint64_t)C::WIN, (uint64_t)C::HIN, (uint64_t)N * C::P_IN};
        const uint64_t strides[3] = {16, (uint64_t)C::WIN * 16, (uint64_t)C::WIN * C::HIN * 16};
        const uint32_t box[4] = {8, (uint32_t)C::WP, (uint32_t)C::HPB, (uint32_t)C::PI};
        int rc = encode_tmap_bf16_4d(&tx, x, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[4] = {8, (uint64_t)C::WO, (uint64_t)C::HO, (uint64_t)N * C::P_OUT};
        const uint64_t strides[3] = {16, (uint64_t)C::WO * 16, (uint64_t)C::WO * C::HO * 16};
        const uint32_t box[4] = {8, (uint32_t)C::WP, (uint32_t)C::HB, 1};
        int rc = encode_tmap_bf16_4d(&tz, dz, dims, strides, box);
        if (rc) return rc;
    }
    conv_tc_wgrad_tap_kernel<C><<<dim3(G, 1, C::PSPLIT * C::NSPLIT), C::THREADS, C::SMEM, st>>>(tx, tz, work, N);
    int rc = launch_status("conv_tc_wgrad_tap_kernel");
    if (rc) return rc;
    wgrad_reduce_kernel<<<(C::PART + 255) / 256, 256, 0, st>>>(work, G, C::PART, dw);
    return launch_status("wgrad_reduce_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient with output-pixel phases in N (small-channel layers, where the tensor pipe is saturated by N = 16 / 32
// MMAs that do a fraction of its work).  K = flat index of PIXEL GROUPS (y, xq) over the pitch WQ, XPH pixels per group:
//   B (N = XPH*C_out): dz de-interleaved by phase ph = x mod XPH (strided TMA maps), N units = (ph, channel octet) planes;
//   A (M = 64): M units = the 8 taps kw' = ph + kw, elements = 8 input channels.  Unit plane u holds, at (row, xq), the padded
//               pixel (row, XPH*xq + u): the phase plane (u mod XPH) of x shifted by u / XPH groups -- eight TMA loads with
//               the shift in the start coordinate, so that consecutive K rows are consecutive 16-byte units in every plane;
//   D(kh, plane)[(kw', ci)][(ph, co)]   =>   dW[co][ci][kh][kw] = sum_ph D[(kw + ph, ci)][(ph, co)]
// The phase terms go to XPH separate partials per CTA; wgrad_reduce_kernel sums all of them in a fixed order.
template <int CIN_, int COUT_, int HIN_, int WIN_, int KS_, int PAD_, int XPH_, int WQ_, int BANDS_, int SLOTS_, int PSPLIT_>
struct WgPCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, HIN = HIN_, WIN = WIN_, KS = KS_, PAD = PAD_, XPH = XPH_, BANDS = BANDS_, SLOTS = SLOTS_;
    static constexpr int PSPLIT = PSPLIT_, P_IN = CIN / 8, P_OUT = COUT / 8, PI = P_IN / PSPLIT, NTOT = XPH * COUT;
    static constexpr int WP = WIN + 2 * PAD, HO = HIN + 2 * PAD - KS + 1, WO = WP - KS + 1, KWX = KS + XPH - 1;
    static constexpr int WQ = WQ_ ? WQ_ : (WP + XPH - 1) / XPH;                  // pitch of every plane (units)
    static constexpr int HB = HO / BANDS, HPB = HB + KS - 1, HBZ = hbz_for(HB, WQ);
    static constexpr int KSTEPS = HBZ * WQ / 16;
    static constexpr int PLANE_XU = HPB * WQ * 16, PLANE_Z = HBZ * WQ * 16;
    static constexpr int X_BYTES = round_up(8 * PI * PLANE_XU, 128), Z_BYTES = round_up(XPH * P_OUT * PLANE_Z, 128);
    static constexpr int SLOT = X_BYTES + Z_BYTES;
    static constexpr int NACC = KS * PI, TCOLS = pow2_cols(NACC * NTOT);
    static constexpr int ISS = NACC <= 6 ? NACC : 4;
    static constexpr int THREADS = 32 * (1 + ISS + 4);
    static constexpr int PART = COUT * CIN * KS * KS;
    static constexpr int BAR_OFF = SLOTS * SLOT, SMEM = BAR_OFF + 256;
    static_assert(CIN % 8 == 0 && COUT % 8 == 0 && P_IN % PSPLIT == 0 && HO % BANDS == 0 && WO % XPH == 0 && WIN % XPH == 0, "shape");
    static_assert(KWX <= 8 && NTOT <= 256 && NTOT % 16 == 0 && TCOLS <= 512, "taps / N / TMEM");
    static_assert(WQ * XPH >= WP && (BANDS == 1 || HBZ == HB), "pitch / bands");
    static_assert((KSTEPS * 16 + (KS - 1) * WQ - HPB * WQ) * 16 <= Z_BYTES, "x overrun must stay inside the slot");
    static_assert((PI * PLANE_XU) % 128 == 0 && (P_OUT * PLANE_Z) % 128 == 0, "TMA destinations (unit planes, phase planes) must be 128-byte aligned");
    static_assert(SMEM + 1024 <= 227 * 1024 && 2 * SLOTS + 1 <= 24, "shared memory / barriers");
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
conv_tc_wgrad_ph_kernel(const __grid_constant__ TMaps tmaps_x, const __grid_constant__ TMaps tmaps_z, float* __restrict__ work, int N) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);     // full[SLOTS], empty[SLOTS], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::BAR_OFF + 200);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, g = blockIdx.x, split = blockIdx.z;
    const long items = (long)N * C::BANDS;
    const int i0 = (int)(items * g / G), i1 = (int)(items * (g + 1) / G);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (C::SLOTS + s); };
    const uint32_t done_bar = bar0 + 8u * (2 * C::SLOTS);
    {
        uint4* z = reinterpret_cast<uint4*>(smem);                       // unit planes u >= KWX are never loaded: they stay zero
        for (int i = threadIdx.x; i < C::BAR_OFF / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
    }
    if (warp == 0 && lane == 0) {
        for (int r = 0; r < C::XPH; ++r) {
            prefetch_tmap(&tmaps_x.m[r]);
            prefetch_tmap(&tmaps_z.m[r]);
        }
        for (int s = 0; s < C::SLOTS; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), C::ISS);
        }
        mbar_init(done_bar, C::ISS);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<C::TCOLS>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);

    if (warp == 0) {
        if (lane == 0) {
            for (int i = i0; i < i1; ++i) {
                const int k = i - i0, slot = k % C::SLOTS, use = k / C::SLOTS;
                mbar_wait(empty_bar(slot), (use & 1) ^ 1);
                mbar_expect_tx(full_bar(slot), C::KWX * C::PI * C::PLANE_XU + C::XPH * C::P_OUT * C::PLANE_Z);
                const int n = i / C::BANDS, band = i % C::BANDS;
                const uint32_t sa = smem0 + slot * C::SLOT;
#pragma unroll
                for (int u = 0; u < C::KWX; ++u) {   // padded column XPH*xq + u = global column XPH*(xq + a) + r
                    constexpr int X = C::XPH;
                    const int r = ((u - C::PAD) % X + X) % X, a = (u - C::PAD - r) / X;
                    tma_load_4d(sa + u * C::PI * C::PLANE_XU, &tmaps_x.m[r], full_bar(slot), 0, a, band * C::HB - C::PAD,
                                n * C::P_IN + split * C::PI);
                }
#pragma unroll
                for (int ph = 0; ph < C::XPH; ++ph)
                    tma_load_4d(sa + C::X_BYTES + ph * C::P_OUT * C::PLANE_Z, &tmaps_z.m[ph], full_bar(slot), 0, 0, band * C::HB, n * C::P_OUT);
            }
        }
    } else if (warp <= C::ISS) {
        if (lane == 0) {                     // issuer w owns the accumulators a == w-1 (mod ISS)
            constexpr uint32_t idesc = idesc_bf16(C::NTOT, true, true, 64);
            for (int i = i0; i < i1; ++i) {
                const int k = i - i0, slot = k % C::SLOTS, use = k / C::SLOTS;
                mbar_wait(full_bar(slot), use & 1);
                tc_fence_after_sync();
                const uint32_t xa = smem0 + slot * C::SLOT, za = xa + C::X_BYTES;
                for (int ks = 0; ks < C::KSTEPS; ++ks) {
                    const uint32_t acc = (i > i0 || ks > 0) ? 1u : 0u;
                    const uint64_t bd = smem_desc(za + ks * 256, 128, C::PLANE_Z);
#pragma unroll
                    for (int kh = 0; kh < C::KS; ++kh) {
#pragma unroll
                        for (int pl = 0; pl < C::PI; ++pl) {
                            if ((kh * C::PI + pl) % C::ISS != warp - 1) continue;
                            const uint64_t ad = smem_desc(xa + pl * C::PLANE_XU + (ks * 16 + kh * C::WQ) * 16, 128, C::PI * C::PLANE_XU);
                            mma_bf16(tmem_base + (kh * C::PI + pl) * C::NTOT, ad, bd, idesc, acc);
                        }
                    }
                }
                mma_commit(empty_bar(slot));
            }
            mma_commit(done_bar);
        }
    } else {
        // epilogue: M = 64 accumulators occupy lanes 0-15 of each 32-lane quadrant: row m = quad*16 + lane = (tap kw' = u, ci8)
        const int quad = warp & 3;
        const int m = quad * 16 + (lane & 15), u = m >> 3, ci8 = m & 7;
        if (i1 > i0) {
            mbar_wait(done_bar, 0);
            tc_fence_after_sync();
        }
#pragma unroll 1
        for (int a = 0; a < C::NACC; ++a) {
            const int kh = a / C::PI, pl = a % C::PI;
            const int ci = (split * C::PI + pl) * 8 + ci8;
#pragma unroll 1
            for (int cc = 0; cc < C::NTOT / 16; ++cc) {
                uint32_t v[16];
                if (i1 > i0) {
                    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + a * C::NTOT + cc * 16, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int t = 0; t < 16; ++t) v[t] = 0u;
                }
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int col = cc * 16 + t, ph = col / C::COUT, co = col % C::COUT;
                    const int kw = u - ph;
                    if (lane < 16 && kw >= 0 && kw < C::KS)
                        work[((long)g * C::XPH + ph) * C::PART + ((co * C::CIN + ci) * C::KS + kh) * C::KS + kw] = __uint_as_float(v[t]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<C::TCOLS>(tmem_base);
    }
}

template <class C>
int launch_conv_tc_wgrad_ph(const void* x, const void* dz, float* dw, float* work, int N, cudaStream_t st, int64_t* need) {
    int G = sm_count() / C::PSPLIT;
    const long items = (long)N * C::BANDS;
    if (G > items) G = (int)items;
    if (G < 1) G = 1;
    if (need) {
        *need = (int64_t)G * C::XPH * C::PART;
        return 0;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_ph_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            set_error("conv_tc_wgrad: cannot set %d bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    TMaps tx, tz;
    for (int r = 0; r < 4; ++r) {
        const int rr = r < C::XPH ? r : 0;
        {
            const uint64_t dims[4] = {8, (uint64_t)C::WIN / C::XPH, (uint64_t)C::HIN, (uint64_t)N * C::P_IN};
            const uint64_t strides[3] = {16 * (uint64_t)C::XPH, (uint64_t)C::WIN * 16, (uint64_t)C::WIN * C::HIN * 16};
            const uint32_t box[4] = {8, (uint32_t)C::WQ, (uint32_t)C::HPB, (uint32_t)C::PI};
            int rc = encode_tmap_bf16_4d(&tx.m[r], reinterpret_cast<const uint8_t*>(x) + 16 * rr, dims, strides, box);
            if (rc) return rc;
        }
        {
            const uint64_t dims[4] = {8, (uint64_t)C::WO / C::XPH, (uint64_t)C::HO, (uint64_t)N * C::P_OUT};
            const uint64_t strides[3] = {16 * (uint64_t)C::XPH, (uint64_t)C::WO * 16, (uint64_t)C::WO * C::HO * 16};
            const uint32_t box[4] = {8, (uint32_t)C::WQ, (uint32_t)C::HBZ, (uint32_t)C::P_OUT};
            int rc = encode_tmap_bf16_4d(&tz.m[r], reinterpret_cast<const uint8_t*>(dz) + 16 * rr, dims, strides, box);
            if (rc) return rc;
        }
    }
    conv_tc_wgrad_ph_kernel<C><<<dim3(G, 1, C::PSPLIT), C::THREADS, C::SMEM, st>>>(tx, tz, work, N);
    int rc = launch_status("conv_tc_wgrad_ph_kernel");
    if (rc) return rc;
    wgrad_reduce_kernel<<<(C::PART + 255) / 256, 256, 0, st>>>(work, G * C::XPH, C::PART, dw);
    return launch_status("wgrad_reduce_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// First layer, fused backward: BatchNorm-apply / ReLU / max-pool backward + weight gradient in ONE kernel.  The first layer
// needs no data gradient, so its dz (the largest tensor of the backward pass) has a single consumer: instead of writing it
// to HBM and reading it back, the pre-BatchNorm z tile (fp16 act8) is TMA-loaded straight into the position of the dz MMA
// operand, four warps transform it IN PLACE into dz = ca*z + cb + [arg-max] a*g (bf16) from the pooled gradient tile, fence
// it to the async proxy, and the MMA warps accumulate dW.  The same warps accumulate the conv bias gradient sum(dz).
//
// Operands (both MN-major, K = flat index of PIXEL GROUPS xq over the pitch WQ of the quad8 image, 4 pixels per group):
//   A (M = 64): the quad8 x slab; M units = 8 image rows kh (SBO = one row), the 8 elements of a unit = taps kw' = ph + kw;
//   B (N = 4*C_out): dz de-interleaved by output phase ph = x mod 4 (one strided TMA map per phase, like the forward
//               convolution's input planes): N units = (ph, channel octet) planes, SBO = plane stride.
//   D[(kh, kw')][(ph, co)] = sum_xq x[.., 4xq + kw'] dz[co][.., 4xq + ph]   =>   dW[co][kh][kw] = sum_ph D[(kh, kw + ph)][(ph, co)]
// -- a quarter of the K steps of the one-pixel-per-row formulation, at N = 32 / 128 instead of 8 / 32.
//   cst: per (view, channel) constants float4 {a, b, ca, cb} prepared by wgrad_l0_consts_kernel from the BatchNorm tensors.
// (A variant that TMA-loads the z tile densely into a staging area and lets the transform warps scatter dz into the phase
// planes was measured slower -- 0.73 ms with one CTA and 8 transform warps, 0.83 ms with two CTAs, against 0.55 ms -- and removed.)
constexpr int L0F_ISS = 3;                       // MMA-issuer warps (K steps interleaved, one TMEM accumulator each)

template <int COUT_, int HIN_, int WIN_, int KS_, int PAD_, int BANDS_, int SLOTS_, int CTAS_, int NSPLIT_ = 1>
struct L0FCfg {
    static constexpr int COUT = COUT_, HIN = HIN_, WIN = WIN_, KS = KS_, PAD = PAD_, BANDS = BANDS_, SLOTS = SLOTS_, CTAS = CTAS_;
    // output channels split over gridDim.z (32 channels at 112x112: the four dz phase-plane groups of all channels are 119 KB; two
    // 16-channel CTAs per SM double the transform warps and interleave their load / transform / MMA phases)
    static constexpr int NSPLIT = NSPLIT_, COUTL = COUT / NSPLIT;
    static constexpr int TW = 4;                                                 // transform warps
    static constexpr int THREADS = 32 * (1 + L0F_ISS + TW);
    static constexpr int P_OUT = COUT / 8, P_OUTL = COUTL / 8, NTOT = 4 * COUTL;
    static constexpr int WP = WIN + 2 * PAD, WQ = (WP + 3) / 4;                  // quad8 units per row
    static constexpr int HO = HIN + 2 * PAD - KS + 1, WO = WP - KS + 1;
    static constexpr int HB = HO / BANDS, HPB = HB + KS - 1, HBZ = hbz_for(HB, WQ);
    static constexpr int KSTEPS = HBZ * WQ / 16;
    static constexpr int PLANE_X = HPB * WQ * 16, PLANE_Z = HBZ * WQ * 16;       // x slab; one (phase, octet) dz plane
    static constexpr int X_BYTES = round_up(PLANE_X, 128), Z_BYTES = round_up(4 * P_OUTL * PLANE_Z, 128);
    static constexpr int HBP = HB / 2, WOP = WO / 2;                             // pooled rows / columns of a band
    static constexpr int G_BYTES = round_up(P_OUTL * HBP * WOP * 16, 128);
    static constexpr int SLOT = X_BYTES + Z_BYTES + G_BYTES;
    static constexpr int STAGE_BYTES = 4 * COUTL * KS * KS * 4;                   // end-of-kernel staging of the phase terms
    static constexpr int CST_OFF = SLOTS * SLOT, BAR_OFF = CST_OFF + COUTL * 16 + TW * COUTL * 4;
    static constexpr int SMEM = BAR_OFF + 256;
    static constexpr int ACC_COLS = round_up(NTOT, 32), TCOLS = pow2_cols(L0F_ISS * ACC_COLS);
    static constexpr int PART = COUT * KS * KS, PARTL = COUTL * KS * KS;
    static_assert(HO == HIN && WO == WIN && HO % BANDS == 0 && HB % 2 == 0 && WO % 4 == 0 && COUT % (8 * NSPLIT) == 0, "geometry");
    static_assert(KS + 3 <= 8, "all taps of all phases must lie inside one 8-pixel unit");
    static_assert(BANDS == 1 || HBZ == HB, "row bands need HB*WQ to be a multiple of 16");
    static_assert(TCOLS * CTAS <= 512 && NTOT <= 256 && NTOT % 16 == 0, "TMEM columns / N");
    static_assert((SMEM + 1024) * CTAS <= 227 * 1024, "shared memory per SM");
    static_assert(STAGE_BYTES <= SLOT, "staging reuses the first slot");
    static_assert((KSTEPS * 16 + 7 * WQ + 8 - HPB * WQ) * 16 <= Z_BYTES, "x overrun must stay inside the slot");
    static_assert((P_OUTL * PLANE_Z) % 128 == 0, "TMA destinations (phase planes) must be 128-byte aligned");
    static_assert(KSTEPS >= L0F_ISS, "every issuer needs a K step");
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
conv_tc_wgrad_l0_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ TMaps tmaps_z,
                              const __grid_constant__ CUtensorMap tmap_g, const float4* __restrict__ cst, float* __restrict__ work,
                              double* __restrict__ dbsum, int N, int n_per_view) {
    constexpr int HBP = C::HBP, WOP = C::WOP, SLOT = C::SLOT;
    extern __shared__ __align__(1024) uint8_t smem[];
    float4* cst_s = reinterpret_cast<float4*>(smem + C::CST_OFF);
    float* db_s = reinterpret_cast<float*>(smem + C::CST_OFF + C::COUTL * 16);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);     // full[S], ready[S], empty[S], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::BAR_OFF + 200);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, g = blockIdx.x, ns = blockIdx.z;
    const long items = (long)N * C::BANDS;
    const int i0 = (int)(items * g / G), i1 = (int)(items * (g + 1) / G);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto ready_bar = [&](int s) { return bar0 + 8u * (C::SLOTS + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * C::SLOTS + s); };
    const uint32_t done_bar = bar0 + 8u * (3 * C::SLOTS);
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < C::CST_OFF / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < C::TW * C::COUTL; i += blockDim.x) db_s[i] = 0.f;   // [transform warp][channel]
        fence_proxy_async_smem();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        for (int r = 0; r < 4; ++r) prefetch_tmap(&tmaps_z.m[r]);
        prefetch_tmap(&tmap_g);
        for (int s = 0; s < C::SLOTS; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(ready_bar(s), 32 * C::TW);
            mbar_init(empty_bar(s), L0F_ISS);
        }
        mbar_init(done_bar, L0F_ISS);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<C::TCOLS>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t smem0 = smem_u32(smem);

    if (warp == 0) {
        if (lane == 0) {
            for (int i = i0; i < i1; ++i) {
                const int k = i - i0, slot = k % C::SLOTS, use = k / C::SLOTS;
                mbar_wait(empty_bar(slot), (use & 1) ^ 1);
                mbar_expect_tx(full_bar(slot), C::PLANE_X + 4 * C::P_OUTL * C::PLANE_Z + C::P_OUTL * HBP * WOP * 16);
                const int n = i / C::BANDS, band = i % C::BANDS;
                const uint32_t sa = smem0 + slot * SLOT;
                tma_load_4d(sa, &tmap_x, full_bar(slot), 0, 0, band * C::HB - C::PAD, n);
#pragma unroll
                for (int ph = 0; ph < 4; ++ph)      // columns x = 4*xq + ph of every z row -> plane group ph
                    tma_load_4d(sa + C::X_BYTES + ph * C::P_OUTL * C::PLANE_Z, &tmaps_z.m[ph], full_bar(slot), 0, 0, band * C::HB, n * C::P_OUT + ns * C::P_OUTL);
                tma_load_4d(sa + C::X_BYTES + C::Z_BYTES, &tmap_g, full_bar(slot), 0, 0, band * HBP, n * C::P_OUT + ns * C::P_OUTL);
            }
        }
    } else if (warp <= L0F_ISS) {
        if (lane == 0) {
            // issuer w takes the K steps ks = w, w + ISS, ... of every item and owns accumulator w (summed in the epilogue):
            // one issuing thread sustains only ~1 MMA / 140 cycles, the tensor pipe of the SM about four times that
            constexpr uint32_t idesc = idesc_bf16(C::NTOT, true, true, 64);
            const int w = warp - 1;
            const uint32_t acc = tmem_base + w * C::ACC_COLS;
            for (int i = i0; i < i1; ++i) {
                const int k = i - i0, slot = k % C::SLOTS, use = k / C::SLOTS;
                mbar_wait(ready_bar(slot), use & 1);
                tc_fence_after_sync();
                const uint32_t xa = smem0 + slot * SLOT, za = xa + C::X_BYTES;
                for (int ks = w; ks < C::KSTEPS; ks += L0F_ISS) {
                    const uint64_t bd = smem_desc(za + ks * 256, 128, C::PLANE_Z);
                    const uint64_t ad = smem_desc(xa + ks * 256, 128, C::WQ * 16);
                    mma_bf16(acc, ad, bd, idesc, (i > i0 || ks >= L0F_ISS) ? 1u : 0u);
                }
                mma_commit(empty_bar(slot));
            }
            mma_commit(done_bar);
        }
    } else {
        // ===== transform warps: z -> dz in place (then the end-of-kernel epilogue) =====
        constexpr int NT = 32 * C::TW;                                   // transform threads
        const int t = threadIdx.x - 32 * (1 + L0F_ISS);                  // 0..NT-1
        const int tw = warp - (1 + L0F_ISS);                             // transform warp
        int cur_view = -1;
        for (int i = i0; i < i1; ++i) {
            const int k = i - i0, slot = k % C::SLOTS, use = k / C::SLOTS;
            const int n = i / C::BANDS, view = n / n_per_view;
            if (view != cur_view) {                                      // uniform over the four warps
                asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");    // nobody still reads the previous view's constants
                if (t < C::COUTL) cst_s[t] = __ldg(cst + (size_t)view * C::COUT + ns * C::COUTL + t);
                asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
                cur_view = view;
            }
            mbar_wait(full_bar(slot), use & 1);
            uint8_t* zimg = smem + slot * SLOT + C::X_BYTES;
            const uint4* gimg = reinterpret_cast<const uint4*>(smem + slot * SLOT + C::X_BYTES + C::Z_BYTES);
#pragma unroll 1
            for (int o = 0; o < C::P_OUTL; ++o) {
                float4 c4[8];
                float acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    c4[j] = cst_s[o * 8 + j];
                    acc[j] = 0.f;
                }
                for (int e = t; e < HBP * WOP; e += NT) {
                    const int py = e / WOP, px = e - py * WOP;
                    // pooling window: rows 2py, 2py+1; columns 2px, 2px+1 = pixel group px/2, phases 2(px&1) and 2(px&1)+1
                    const int ph0 = 2 * (px & 1);
                    uint4* zp0 = reinterpret_cast<uint4*>(zimg + (size_t)(ph0 * C::P_OUTL + o) * C::PLANE_Z) + (2 * py) * C::WQ + (px >> 1);
                    uint4* zp1 = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(zp0) + (size_t)C::P_OUTL * C::PLANE_Z);
                    const uint4 raw[4] = {zp0[0], zp1[0], zp0[C::WQ], zp1[C::WQ]};
                    const uint4 graw = gimg[o * (HBP * WOP) + e];
                    uint32_t outw[4][4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {                        // channel pairs of the octet
                        const uint32_t gw = (&graw.x)[h];
                        float res[4][2];
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const float4 cc = c4[2 * h + half];
                            float zv[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t zw = (&raw[q].x)[h];
                                const __half2 hz = *reinterpret_cast<const __half2*>(&zw);
                                zv[q] = half ? __high2float(hz) : __low2float(hz);
                            }
                            const float gv = __uint_as_float(half ? (gw & 0xFFFF0000u) : (gw << 16));
                            int kk = 0;
                            float m = fmaf(cc.x, zv[0], cc.y);
#pragma unroll
                            for (int q = 1; q < 4; ++q) {
                                const float y = fmaf(cc.x, zv[q], cc.y);
                                if (y > m) { m = y; kk = q; }
                            }
                            const float ag = (m > 0.f) ? cc.x * gv : 0.f;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                res[q][half] = fmaf(cc.z, zv[q], cc.w) + ((q == kk) ? ag : 0.f);
                                acc[2 * h + half] += res[q][half];
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) outw[q][h] = pack_bf16(res[q][0], res[q][1]);
                    }
                    zp0[0] = make_uint4(outw[0][0], outw[0][1], outw[0][2], outw[0][3]);
                    zp1[0] = make_uint4(outw[1][0], outw[1][1], outw[1][2], outw[1][3]);
                    zp0[C::WQ] = make_uint4(outw[2][0], outw[2][1], outw[2][2], outw[2][3]);
                    zp1[C::WQ] = make_uint4(outw[3][0], outw[3][1], outw[3][2], outw[3][3]);
                }
                if (dbsum != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v = acc[j];
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                        if (lane == 0) db_s[tw * C::COUTL + o * 8 + j] += v;      // single writer per slot: deterministic
                    }
                }
            }
            fence_proxy_async_smem();                                    // generic-proxy writes -> visible to the tensor core
            mbar_arrive(ready_bar(slot));
        }
        // ---- end of kernel: bias-gradient sums and the dW partial of this CTA ----
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        if (t < C::COUTL && dbsum != nullptr) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < C::TW; ++q) v += (double)db_s[q * C::COUTL + t];
            atomicAdd(&dbsum[ns * C::COUTL + t], v);
        }
        const int quad = warp & 3;
        const int m = quad * 16 + (lane & 15), j = m >> 3, e8 = m & 7;   // accumulator row = (image row kh = j, tap kw' = e8)
        float* part = work + (long)g * C::PART;
        float* stage = reinterpret_cast<float*>(smem);                   // [ph][co][kh][kw]: every MMA has completed (done_bar)
        if (i1 > i0 && tw < 4) {                                         // four warps cover the four TMEM lane quadrants
            mbar_wait(done_bar, 0);
            tc_fence_after_sync();
#pragma unroll 1
            for (int cc = 0; cc < C::NTOT / 16; ++cc) {
                float sum[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) sum[q] = 0.f;
#pragma unroll
                for (int w = 0; w < L0F_ISS; ++w) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + w * C::ACC_COLS + cc * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 16; ++q) sum[q] += __uint_as_float(v[q]);
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int col = cc * 16 + q, ph = col / C::COUTL, co = col % C::COUTL;
                    const int kw = e8 - ph;
                    if (lane < 16 && j < C::KS && kw >= 0 && kw < C::KS) stage[((ph * C::COUTL + co) * C::KS + j) * C::KS + kw] = sum[q];
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        for (int idx = t; idx < C::PARTL; idx += NT) {                   // fixed summation order over the four phases
            float v = 0.f;
            if (i1 > i0) {
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) v += stage[ph * C::PARTL + idx];
            }
            part[ns * C::PARTL + idx] = v;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<C::TCOLS>(tmem_base);
    }
}

// {a, b, ca, cb} per (view, channel): y = a*z + b decides arg-max / ReLU; dz = ca*z + cb + [arg-max] a*g
__global__ void wgrad_l0_consts_kernel(const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, const double* __restrict__ sums, float inv_cnt, int n,
                                       float4* __restrict__ cst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const float a = scale[c], is = invstd[c], nm = -mean[c] * is;
    const float k1 = (float)sums[(size_t)c * 2] * inv_cnt, k2 = (float)sums[(size_t)c * 2 + 1] * inv_cnt;
    cst[c] = make_float4(a, shift[c], -a * is * k2, -a * (k1 + nm * k2));
}

template <class C>
int launch_conv_tc_wgrad_l0_fused(const void* x, const void* z, const void* dp, const float* scale, const float* shift, const float* mean,
                                  const float* invstd, const double* sums, float* dw, double* dbsum, float* work, int N, int n_per_view,
                                  cudaStream_t st, int64_t* need) {
    int G = sm_count() * C::CTAS / C::NSPLIT;
    const long items = (long)N * C::BANDS;
    if (G > items) G = (int)items;
    if (G < 1) G = 1;
    const int views = N / n_per_view;
    const int64_t cst_floats = (int64_t)views * C::COUT * 4;
    if (need) {
        *need = (int64_t)G * C::PART + cst_floats + 8;
        return 0;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_l0_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            set_error("conv_tc_wgrad_l0_fused: cannot set %d bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    float4* cst = reinterpret_cast<float4*>(work + (((int64_t)G * C::PART + 3) / 4) * 4);
    wgrad_l0_consts_kernel<<<(views * C::COUT + 127) / 128, 128, 0, st>>>(scale, shift, mean, invstd, sums,
                                                                          1.0f / ((float)n_per_view * C::HO * C::WO), views * C::COUT, cst);
    CUtensorMap tx, tg;
    TMaps tz;
    {
        const uint64_t dims[4] = {8, (uint64_t)C::WQ, (uint64_t)C::HIN, (uint64_t)N};
        const uint64_t strides[3] = {16, (uint64_t)C::WQ * 16, (uint64_t)C::WQ * C::HIN * 16};
        const uint32_t box[4] = {8, (uint32_t)C::WQ, (uint32_t)C::HPB, 1};
        int rc = encode_tmap_bf16_4d(&tx, x, dims, strides, box);
        if (rc) return rc;
    }
    for (int ph = 0; ph < 4; ++ph) {        // fp16 data: same 2-byte elements, no conversion
        // columns 4*i + ph
        const uint64_t dims[4] = {8, (uint64_t)C::WO / 4, (uint64_t)C::HO, (uint64_t)N * C::P_OUT};
        const uint64_t strides[3] = {64, (uint64_t)C::WO * 16, (uint64_t)C::WO * C::HO * 16};
        const uint32_t box[4] = {8, (uint32_t)C::WQ, (uint32_t)C::HBZ, (uint32_t)C::P_OUTL};
        int rc = encode_tmap_bf16_4d(&tz.m[ph], reinterpret_cast<const uint8_t*>(z) + 16 * ph, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[4] = {8, (uint64_t)C::WOP, (uint64_t)(C::HO / 2), (uint64_t)N * C::P_OUT};
        const uint64_t strides[3] = {16, (uint64_t)C::WOP * 16, (uint64_t)C::WOP * (C::HO / 2) * 16};
        const uint32_t box[4] = {8, (uint32_t)C::WOP, (uint32_t)C::HBP, (uint32_t)C::P_OUTL};
        int rc = encode_tmap_bf16_4d(&tg, dp, dims, strides, box);
        if (rc) return rc;
    }
    conv_tc_wgrad_l0_fused_kernel<C><<<dim3(G, 1, C::NSPLIT), C::THREADS, C::SMEM, st>>>(tx, tz, tg, cst, work, dbsum, N, n_per_view);
    int rc = launch_status("conv_tc_wgrad_l0_fused_kernel");
    if (rc) return rc;
    wgrad_reduce_kernel<<<(C::PART + 255) / 256, 256, 0, st>>>(work, G, C::PART, dw);
    return launch_status("wgrad_reduce_kernel");
}

//                           CIN COUT HIN WIN KS PAD BANDS SLOTS PSPLIT
//                   CIN COUT HIN WIN KS PAD XPH WQ BANDS SLOTS PSPLIT   (phases in N)
using WgA1 = WgPCfg<8, 16, 56, 56, 5, 2, 4, 16, 4, 3, 1>;
using WgA2 = WgPCfg<16, 32, 28, 28, 5, 2, 2, 16, 2, 3, 2>;
using WgA3 = TcWgCfg<32, 64, 14, 14, 5, 2, 1, 3, 4>;
using WgI1 = TcWgCfg<32, 64, 14, 14, 5, 0, 1, 4, 4>;
using WgS1 = TcWgCfg<32, 64, 14, 14, 3, 1, 1, 3, 2>;
using WgA0 = TcWgCfg<1, 8, 112, 112, 5, 2, 7, 1, 1, 3>;   // first layers (shift8 image)
using WgI0 = TcWgCfg<1, 32, 28, 28, 5, 2, 1, 3, 1>;
using WgS0 = TcWgCfg<1, 32, 28, 28, 3, 1, 1, 3, 1>;
using WgS2 = TcWgCfg<64, 128, 7, 7, 3, 1, 1, 3, 4, 1, 2>;
// the 3x3 audio stack of the simple multimodal encoders (models/dino.py:43-72): input planes / dz planes split over gridDim.z
//                  CIN COUT HIN WIN KS PAD BANDS SLOTS PSPLIT CTAS NSPLIT
using WgB0 = TcWgCfg<1, 32, 112, 112, 3, 1, 7, 2, 1, 1, 2>;          // (unfused first-layer weight gradient: tests / fp32-dz fallback)
using WgB1 = TcWgCfg<32, 64, 56, 56, 3, 1, 7, 2, 2>;                    // (shift-row formulation, 3 of 8 M rows per plane used: kept for A/B)
using WgB2 = TcWgCfg<64, 128, 28, 28, 3, 1, 1, 2, 4, 1, 4>;
using WgB3 = TcWgCfg<128, 256, 14, 14, 3, 1, 1, 4, 8, 1, 4>;
// one accumulator per tap (every MMA row / column a real weight): the wide 3x3 layers
//                     CIN COUT HIN WIN KS PAD BANDS SLOTS PSPLIT NSPLIT
using TapB1 = TapWgCfg<32, 64, 56, 56, 3, 1, 7, 2, 1, 1>;            // M = 64,  N = 32
using TapB2 = TapWgCfg<64, 128, 28, 28, 3, 1, 4, 2, 2, 1>;           // M = 128, N = 32 (two input-channel halves)
using TapB3 = TapWgCfg<128, 256, 14, 14, 3, 1, 1, 2, 4, 2>;          // M = 128, N = 32
using TapS1 = TapWgCfg<32, 64, 14, 14, 3, 1, 1, 3, 1, 1>;            // image stack of the simple encoders / image_simple
using TapS2 = TapWgCfg<64, 128, 7, 7, 3, 1, 1, 4, 2, 1>;

//                         CIN COUT NPAD HIN  WIN KS PAD BANDS SLOTS
using CfgA1 = TcCfg<8, 16, 16, 56, 56, 5, 2, 2, 2, 2>;   // audio conv2 forward (4 x phases: N = 64)
using CfgA2 = TcCfg<16, 32, 32, 28, 28, 5, 2, 1, 2, 1>;  // audio conv3 forward (2 x phases: N = 64)
using CfgA3 = TcCfg<32, 64, 64, 14, 14, 5, 2, 1, 4>;     // audio conv4 forward (an N split over 2 CTAs/SM measured slower: N = 64 MMAs amortise the fixed cost)
using CfgI1 = TcCfg<32, 64, 64, 14, 14, 5, 0, 1, 4>;     // image conv2 forward (no padding)
using CfgA1d = TcCfg<16, 8, 16, 56, 56, 5, 2, 2, 2, 1>;  // data gradients (C_in/C_out swapped, pad' = K-1-pad); 4 x phases: N = 32
using CfgA2d = TcCfg<32, 16, 16, 28, 28, 5, 2, 1, 2, 1>;  // 2 x phases: N = 32
using CfgA3d = TcCfg<64, 32, 32, 14, 14, 5, 2, 1, 2>;
using CfgI1d = TcCfg<64, 32, 32, 10, 10, 5, 4, 1, 2>;
using CfgA0 = TcCfg<1, 8, 16, 112, 112, 5, 2, 2, 2, 3>;  // first layers on the quad8 image (4 x phases: N = 32 / 128): audio conv1
using CfgI0 = TcCfg<1, 32, 32, 28, 28, 5, 2, 1, 4>;      //   image conv1
using CfgS0 = TcCfg<1, 32, 32, 28, 28, 3, 1, 1, 4>;      //   image_simple conv1
using CfgS2 = TcCfg<64, 128, 128, 7, 7, 3, 1, 1, 3, 2, 2>;   // image_simple conv3 forward (two 64-channel slices) / data gradient
using CfgS2d = TcCfg<128, 64, 64, 7, 7, 3, 1, 1, 2, 1, 1>;
using CfgS1 = TcCfg<32, 64, 64, 14, 14, 3, 1, 1, 4>;     // image_simple conv2 forward / its data gradient
using CfgS1d = TcCfg<64, 32, 32, 14, 14, 3, 1, 1, 4>;
// the 3x3 audio stack of the simple multimodal encoders (models/dino.py:43-72), forward / data gradient.  The wide layers keep their
// weights resident by splitting the output channels over gridDim.z (NSPLIT) and stream the input channel planes in KCH chunks.
//                    CIN COUT NPAD HIN WIN KS PAD BANDS SLOTS CTAS NSPLIT KCH
using CfgB0 = TcCfg<1, 32, 32, 112, 112, 3, 1, 7, 2, 2>;             // quad8 first layer, N = 4 phases x 32 (7 bands: 4 tiles per item)
using CfgB1 = TcCfg<32, 64, 64, 56, 56, 3, 1, 7, 2>;                 // 2 x phases: N = 128
using CfgB1d = TcCfg<64, 32, 32, 56, 56, 3, 1, 7, 2, 1, 1, 2>;       // 2 x phases: N = 64, two chunks of 32 input channels
using CfgB2 = TcCfg<64, 128, 128, 28, 28, 3, 1, 2, 3, 1, 2, 2>;
using CfgB2d = TcCfg<128, 64, 64, 28, 28, 3, 1, 2, 2, 1, 1, 4>;
using CfgB3 = TcCfg<128, 256, 256, 14, 14, 3, 1, 1, 3, 1, 4, 4>;
using CfgB3d = TcCfg<256, 128, 128, 14, 14, 3, 1, 1, 3, 1, 4, 8>;

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_conv_tc_supported(int Cin, int Cout, int H, int W, int K, int pad) {
#define TC_MATCH(CFG) if (Cin == CFG::CIN && Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) return 1;
    TC_MATCH(CfgA1) TC_MATCH(CfgA2) TC_MATCH(CfgA3) TC_MATCH(CfgI1) TC_MATCH(CfgA1d) TC_MATCH(CfgA2d) TC_MATCH(CfgA3d) TC_MATCH(CfgI1d)
    TC_MATCH(CfgS1) TC_MATCH(CfgS1d) TC_MATCH(CfgA0) TC_MATCH(CfgI0) TC_MATCH(CfgS0) TC_MATCH(CfgS2) TC_MATCH(CfgS2d)
    TC_MATCH(CfgB0) TC_MATCH(CfgB1) TC_MATCH(CfgB1d) TC_MATCH(CfgB2) TC_MATCH(CfgB2d) TC_MATCH(CfgB3) TC_MATCH(CfgB3d)
#undef TC_MATCH
    return 0;
}

int64_t b200_conv_tc_weight_bytes(int Cin, int Cout, int K) {
    int xph, npad, total;
    conv_tc_weight_geometry(Cin, Cout, K, &xph, &npad, &total);
    return (int64_t)total * 2;
}

int b200_conv_tc_prep_weights(const float* w, void* out, int Cin, int Cout, int K, int flip, void* stream) {
    B200_REQUIRE(w && out, -1, "conv_tc_prep_weights: null pointer");
    B200_REQUIRE(Cin == 1 || Cin == 8 || (Cin % 16 == 0 && Cin > 0), -2, "conv_tc_prep_weights: C_in must be 1, 8 or a multiple of 16 (got %d)", Cin);
    B200_REQUIRE(!(Cin == 1 && flip), -2, "conv_tc_prep_weights: the first layer has no data gradient");
    B200_REQUIRE(Cout % 8 == 0 && Cout > 0 && Cout <= 256, -2, "conv_tc_prep_weights: C_out must be a multiple of 8, <= 256 (got %d)", Cout);
    const int xph = Cin == 1 ? 4 : xph_for(Cin, Cout, K);
    const int npad = (xph * Cout + 15) / 16 * 16;
    const int total = (int)(b200_conv_tc_weight_bytes(Cin, Cout, K) / 2);
    conv_tc_prep_weights_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(w, reinterpret_cast<__nv_bfloat16*>(out), Cin, Cout, npad, K,
                                                                                      xph, flip, total);
    return launch_status("conv_tc_prep_weights_kernel");
}

int b200_conv_tc_prep_weights_multi(const int64_t* desc_dev, int n, void* stream) {
    B200_REQUIRE(desc_dev && n > 0 && n <= 65535, -1, "conv_tc_prep_weights_multi: bad arguments");
    conv_tc_prep_weights_multi_kernel<<<dim3(16, (unsigned)n), 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(desc_dev));
    return launch_status("conv_tc_prep_weights_multi_kernel");
}

static int g_wgrad_tap = 1;      // 0: the shift-row weight gradient for every geometry (A/B measurements, b200_conv_tc_wgrad_variant)

static int wgrad_tc_dispatch(const void* x, const void* dz, float* dw, float* work, int N, int Cin, int Cout, int H, int W,
                             int K, int pad, cudaStream_t st, int64_t* need) {
#define WG_RUN(CFG) if (Cin == CFG::CIN && Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) \
        return launch_conv_tc_wgrad<CFG>(x, dz, dw, work, N, st, need);
#define WGP_RUN(CFG) if (Cin == CFG::CIN && Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) \
        return launch_conv_tc_wgrad_ph<CFG>(x, dz, dw, work, N, st, need);
    WGP_RUN(WgA1) WGP_RUN(WgA2)
#undef WGP_RUN
#define WGT_RUN(CFG) if (g_wgrad_tap && Cin == CFG::CIN && Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) \
        return launch_conv_tc_wgrad_tap<CFG>(x, dz, dw, work, N, st, need);
    WGT_RUN(TapB1) WGT_RUN(TapB2) WGT_RUN(TapB3) WGT_RUN(TapS1) WGT_RUN(TapS2)
#undef WGT_RUN
    WG_RUN(WgA3) WG_RUN(WgI1) WG_RUN(WgS1) WG_RUN(WgA0) WG_RUN(WgI0) WG_RUN(WgS0) WG_RUN(WgS2) WG_RUN(WgB0) WG_RUN(WgB1) WG_RUN(WgB2) WG_RUN(WgB3)
#undef WG_RUN
    set_error("conv_tc_wgrad: unsupported geometry Cin=%d Cout=%d H=%d W=%d K=%d pad=%d", Cin, Cout, H, W, K, pad);
    return -4;
}

// fused first-layer backward (Cin = 1): geometries of the first layers of the three encoders
//                    COUT HIN  WIN KS PAD BANDS SLOTS CTAS
using WgF_A0 = L0FCfg<8, 112, 112, 5, 2, 7, 2, 2>;
using WgF_I0 = L0FCfg<32, 28, 28, 5, 2, 1, 2, 1>;
using WgF_S0 = L0FCfg<32, 28, 28, 3, 1, 1, 2, 1>;
using WgF_B0 = L0FCfg<32, 112, 112, 3, 1, 7, 1, 2, 2>;   // simple audio stack: two 16-channel CTAs per SM, one 82 KB slot each

static int wgrad_l0_dispatch(const void* x, const void* z, const void* dp, const float* scale, const float* shift, const float* mean,
                             const float* invstd, const double* sums, float* dw, double* dbsum, float* work, int N, int n_per_view, int Cout,
                             int H, int W, int K, int pad, cudaStream_t st, int64_t* need) {
#define WF_RUN(CFG) if (Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) \
        return launch_conv_tc_wgrad_l0_fused<CFG>(x, z, dp, scale, shift, mean, invstd, sums, dw, dbsum, work, N, n_per_view, st, need);
    WF_RUN(WgF_A0) WF_RUN(WgF_I0) WF_RUN(WgF_S0) WF_RUN(WgF_B0)
#undef WF_RUN
    set_error("conv_tc_wgrad_l0_fused: unsupported geometry Cout=%d H=%d W=%d K=%d pad=%d", Cout, H, W, K, pad);
    return -4;
}

int64_t b200_conv_tc_wgrad_l0_fused_work_floats(int N, int n_per_view, int Cout, int H, int W, int K, int pad) {
    int64_t need = 0;
    if (N <= 0 || n_per_view <= 0 || N % n_per_view) return -1;
    int rc = wgrad_l0_dispatch(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, N, n_per_view, Cout, H, W,
                               K, pad, nullptr, &need);
    return rc ? -1 : need;
}

int b200_conv_tc_wgrad_l0_fused(const void* x_quad8, const void* z8, const void* dp8, const float* scale, const float* shift,
                                const float* mean, const float* invstd, const double* sums, float* dw, double* dbsum, float* work, int N,
                                int n_per_view, int Cout, int H, int W, int K, int pad, void* stream) {
    B200_REQUIRE(x_quad8 && z8 && dp8 && scale && shift && mean && invstd && sums && dw && work, -1, "conv_tc_wgrad_l0_fused: null pointer");
    B200_REQUIRE(N > 0 && n_per_view > 0 && N % n_per_view == 0, -2, "conv_tc_wgrad_l0_fused: N=%d must be a multiple of n_per_view=%d", N, n_per_view);
    B200_REQUIRE(((reinterpret_cast<uintptr_t>(x_quad8) | reinterpret_cast<uintptr_t>(z8) | reinterpret_cast<uintptr_t>(dp8) |
                   reinterpret_cast<uintptr_t>(work)) & 15) == 0, -3, "conv_tc_wgrad_l0_fused: pointers must be 16-byte aligned");
    return wgrad_l0_dispatch(x_quad8, z8, dp8, scale, shift, mean, invstd, sums, dw, dbsum, work, N, n_per_view, Cout, H, W, K, pad,
                             as_stream(stream), nullptr);
}

int b200_conv_tc_wgrad_variant(int tap) {
    const int old = g_wgrad_tap;
    if (tap == 0 || tap == 1) g_wgrad_tap = tap;
    return old;
}

int64_t b200_conv_tc_wgrad_work_floats(int N, int Cin, int Cout, int H, int W, int K, int pad) {
    int64_t need = 0;
    int rc = wgrad_tc_dispatch(nullptr, nullptr, nullptr, nullptr, N, Cin, Cout, H, W, K, pad, nullptr, &need);
    return rc ? -1 : need;
}

int b200_conv_tc_wgrad(const void* x_act8, const void* dz_act8, float* dw, float* work, int N, int Cin, int Cout, int H, int W,
                       int K, int pad, void* stream) {
    B200_REQUIRE(x_act8 && dz_act8 && dw && work, -1, "conv_tc_wgrad: null pointer");
    B200_REQUIRE(N > 0, -2, "conv_tc_wgrad: N must be positive");
    B200_REQUIRE(((uintptr_t)x_act8 & 15) == 0 && ((uintptr_t)dz_act8 & 15) == 0, -3, "conv_tc_wgrad: pointers must be 16-byte aligned");
    return wgrad_tc_dispatch(x_act8, dz_act8, dw, work, N, Cin, Cout, H, W, K, pad, as_stream(stream), nullptr);
}

int b200_pack_quad8(const float* x, void* out, int N, int H, int W, int pad, void* stream) {
    B200_REQUIRE(x && out && N > 0 && H > 0 && W > 0 && pad >= 0, -1, "pack_quad8: bad arguments");
    const int WQ = (W + 2 * pad + 3) / 4;
    const long units = (long)N * H * WQ;
    pack_quad8_kernel<<<(unsigned)((units + 255) / 256), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<uint4*>(out), units, W, WQ, pad);
    return launch_status("pack_quad8_kernel");
}

int b200_pack_shift8(const float* x, void* out, int N, int H, int W, int pad, void* stream) {
    B200_REQUIRE(x && out, -1, "pack_shift8: null pointer");
    B200_REQUIRE(N > 0 && H > 0 && W > 0 && pad >= 0, -2, "pack_shift8: bad shape");
    const long px = (long)N * H * (W + pad);
    pack_shift8_kernel<<<(unsigned)((px + 255) / 256), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<uint4*>(out), px, W, pad);
    return launch_status("pack_shift8_kernel");
}

int b200_pack_act8(const float* x, void* out, int N, int C, int H, int W, void* stream) {
    B200_REQUIRE(x && out, -1, "pack_act8: null pointer");
    B200_REQUIRE(C % 8 == 0 && N > 0, -2, "pack_act8: C must be a multiple of 8");
    const long units = (long)N * (C / 8) * H * W;
    pack_act8_kernel<<<(unsigned)((units + 255) / 256), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<uint4*>(out), units, C, H * W);
    return launch_status("pack_act8_kernel");
}

int b200_conv_tc(const void* x_act8, const void* wprep, const float* bias, void* out, double* stats, int N, int n_per_view, int Cin, int Cout,
                 int H, int W, int K, int pad, int out_bf16, void* stream) {
    B200_REQUIRE(x_act8 && wprep && out, -1, "conv_tc: null pointer");
    B200_REQUIRE(N > 0 && n_per_view > 0 && N % n_per_view == 0, -2, "conv_tc: N=%d must be a multiple of n_per_view=%d", N, n_per_view);
    B200_REQUIRE(((uintptr_t)x_act8 & 15) == 0 && ((uintptr_t)wprep & 15) == 0 && ((uintptr_t)out & 15) == 0, -3, "conv_tc: pointers must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
#define TC_RUN(CFG) if (Cin == CFG::CIN && Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) \
        return launch_conv_tc<CFG>(x_act8, wprep, bias, out, stats, N, n_per_view, out_bf16, st);
    TC_RUN(CfgA1) TC_RUN(CfgA2) TC_RUN(CfgA3) TC_RUN(CfgI1) TC_RUN(CfgA1d) TC_RUN(CfgA2d) TC_RUN(CfgA3d) TC_RUN(CfgI1d) TC_RUN(CfgS1) TC_RUN(CfgS1d) TC_RUN(CfgA0) TC_RUN(CfgI0) TC_RUN(CfgS0) TC_RUN(CfgS2) TC_RUN(CfgS2d)
    TC_RUN(CfgB0) TC_RUN(CfgB1) TC_RUN(CfgB1d) TC_RUN(CfgB2) TC_RUN(CfgB2d) TC_RUN(CfgB3) TC_RUN(CfgB3d)
#undef TC_RUN
    set_error("conv_tc: unsupported geometry Cin=%d Cout=%d H=%d W=%d K=%d pad=%d", Cin, Cout, H, W, K, pad);
    return -4;
}

int b200_conv_tc_pool_supported(int Cin, int Cout, int H, int W, int K, int pad) {
#define TC_HAS(CFG) if (Cin == CFG::CIN && Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) return CFG::POOL_OK ? 1 : 0;
    TC_HAS(CfgA1) TC_HAS(CfgA2) TC_HAS(CfgA3) TC_HAS(CfgI1) TC_HAS(CfgS1) TC_HAS(CfgA0) TC_HAS(CfgI0) TC_HAS(CfgS0) TC_HAS(CfgS2)
    TC_HAS(CfgB0) TC_HAS(CfgB1)
#undef TC_HAS
    return 0;
}

int b200_conv_tc_pool(const void* x_act8, const void* wprep, const float* bias, const float* gamma, void* z_out, void* pool_out, double* stats,
                      int N, int n_per_view, int Cin, int Cout, int H, int W, int K, int pad, int z_fmt, void* stream) {
    B200_REQUIRE(x_act8 && wprep && bias && gamma && pool_out && stats, -1, "conv_tc_pool: null pointer");
    B200_REQUIRE(N > 0 && n_per_view > 0 && N % n_per_view == 0, -2, "conv_tc_pool: N=%d must be a multiple of n_per_view=%d", N, n_per_view);
    B200_REQUIRE(((uintptr_t)x_act8 & 15) == 0 && ((uintptr_t)wprep & 15) == 0 && ((uintptr_t)z_out & 15) == 0 && ((uintptr_t)pool_out & 15) == 0, -3,
                 "conv_tc_pool: pointers must be 16-byte aligned");
    B200_REQUIRE(z_out == nullptr || z_fmt == 1 || z_fmt == 2, -2, "conv_tc_pool: z must be bf16 (1) or fp16 (2) act8");
    cudaStream_t st = as_stream(stream);
#define TC_RUN(CFG) if (Cin == CFG::CIN && Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) \
        return launch_conv_tc<CFG>(x_act8, wprep, bias, z_out, stats, N, n_per_view, z_fmt, st, pool_out, gamma);
    TC_RUN(CfgA1) TC_RUN(CfgA2) TC_RUN(CfgA3) TC_RUN(CfgI1) TC_RUN(CfgS1) TC_RUN(CfgA0) TC_RUN(CfgI0) TC_RUN(CfgS0)
    TC_RUN(CfgB0) TC_RUN(CfgB1)
#undef TC_RUN
    set_error("conv_tc_pool: unsupported geometry Cin=%d Cout=%d H=%d W=%d K=%d pad=%d", Cin, Cout, H, W, K, pad);
    return -4;
}

int b200_conv_tc_dgrad_bnstat_supported(int Cin, int Cout, int H, int W, int K, int pad) {
#define TC_HAS(CFG) if (Cin == CFG::CIN && Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) return CFG::BSTAT_OK ? 1 : 0;
    TC_HAS(CfgA1d) TC_HAS(CfgA2d) TC_HAS(CfgA3d) TC_HAS(CfgI1d) TC_HAS(CfgS1d) TC_HAS(CfgS2d)
#undef TC_HAS
    return 0;
}

int b200_conv_tc_dgrad_bnstat(const void* dz_act8, const void* wprep_flip, void* dx_act8, const void* p_act8, const float* gamma, const float* beta,
                              double* sums, int N, int n_per_view, int Cin, int Cout, int H, int W, int K, int pad, void* stream) {
    B200_REQUIRE(dz_act8 && wprep_flip && dx_act8 && p_act8 && gamma && beta && sums, -1, "conv_tc_dgrad_bnstat: null pointer");
    B200_REQUIRE(N > 0 && n_per_view > 0 && N % n_per_view == 0, -2, "conv_tc_dgrad_bnstat: N=%d must be a multiple of n_per_view=%d", N, n_per_view);
    B200_REQUIRE((((uintptr_t)dz_act8 | (uintptr_t)wprep_flip | (uintptr_t)dx_act8) & 15) == 0 && ((uintptr_t)p_act8 & 31) == 0, -3,
                 "conv_tc_dgrad_bnstat: pointers must be 16-byte aligned (p: 32-byte)");
    cudaStream_t st = as_stream(stream);
#define TC_RUN(CFG) if (Cin == CFG::CIN && Cout == CFG::COUT && H == CFG::HIN && W == CFG::WIN && K == CFG::KS && pad == CFG::PAD) \
        return launch_conv_tc<CFG>(dz_act8, wprep_flip, nullptr, dx_act8, sums, N, n_per_view, 1, st, const_cast<void*>(p_act8), gamma, beta);
    TC_RUN(CfgA1d) TC_RUN(CfgA2d) TC_RUN(CfgA3d) TC_RUN(CfgI1d) TC_RUN(CfgS1d) TC_RUN(CfgS2d)
#undef TC_RUN
    set_error("conv_tc_dgrad_bnstat: unsupported geometry Cin=%d Cout=%d H=%d W=%d K=%d pad=%d", Cin, Cout, H, W, K, pad);
    return -4;
}

}  // extern "C"
