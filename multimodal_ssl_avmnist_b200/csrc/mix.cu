// Feature mixing between the per-modality encoders and the fusion MLP of the "simple" multimodal encoder family
// (SURVEY 8f-4): the learnable sigmoid gates of GatedMultiModalEncoder (models/dino.py:237-263) and the pieces of
// CrossModalAttention (models/dino.py:385-405: batch-wide attention softmax((x1 Wq)(x2 Wk)^T / sqrt(D)) (x2 Wv) + x1) that
// are not GEMMs -- the row softmax and its backward -- plus a strided accumulate for the residual / gradient sums.
// The GEMMs themselves are the library's linear kernels (csrc/linear.cu, csrc/gemm_tc.cu).
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ float sigmoidf_(float g) { return 1.0f / (1.0f + expf(-g)); }

// y[m, n] = sigmoid(*gate) * x[m, n]   (forward on the features; backward on the gradient: the same map)
__global__ void __launch_bounds__(256) gate_apply_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy,
                                                         const float* __restrict__ gate, int M, int N) {
    const float s = sigmoidf_(__ldg(gate));
    const int64_t total = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N, n = i - m * N;
        y[m * ldy + n] = s * __ldg(x + m * ldx + n);
    }
}

// sum_{m,n} dy[m,n] x[m,n] in a fixed order (thread-strided partials -> warp -> block -> grid by ticket) -> sum_out[0]
__global__ void __launch_bounds__(256) gate_grad_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ x, int64_t ldx,
                                                        float* __restrict__ work, float* __restrict__ sum_out, int M, int N) {
    const int64_t total = (int64_t)M * N;
    float a = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N, n = i - m * N;
        a = fmaf(__ldg(dy + m * lddy + n), __ldg(x + m * ldx + n), a);
    }
    grid_sum_ordered(block_sum_ordered(warp_sum(a)), work, sum_out);
}

// dgate (+)= sigmoid'(g) * sum
__global__ void gate_grad_finish_kernel(const float* __restrict__ scratch, const float* __restrict__ gate, float* __restrict__ dgate,
                                        int accumulate) {
    const float s = sigmoidf_(gate[0]);
    const float v = scratch[0] * s * (1.0f - s);
    dgate[0] = accumulate ? dgate[0] + v : v;
}

// One block per row: p = softmax(scale * s) in place.  Fixed-order block reductions (deterministic).
__device__ __forceinline__ float block_max_ordered(float v) {
    __shared__ float s_warp[32];
    __shared__ float s_out;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = s_warp[0];
        for (int w = 1; w < nw; ++w) a = fmaxf(a, s_warp[w]);
        s_out = a;
    }
    __syncthreads();
    return s_out;
}
__device__ __forceinline__ float block_sum_bcast(float v) {
    __shared__ float s_out;
    const float a = block_sum_ordered(warp_sum(v));
    if (threadIdx.x == 0) s_out = a;
    __syncthreads();
    const float r = s_out;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ s, int64_t ld, int N, float scale) {
    float* row = s + (int64_t)blockIdx.x * ld;
    float mx = -INFINITY;
    for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, row[n] * scale);
    mx = block_max_ordered(mx);
    float sum = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float e = expf(row[n] * scale - mx);
        row[n] = e;
        sum += e;
    }
    sum = block_sum_bcast(sum);
    const float inv = 1.0f / sum;
    for (int n = threadIdx.x; n < N; n += blockDim.x) row[n] *= inv;
}

// ds = scale * p * (dp - sum_n dp p), in place on dp
__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(float* __restrict__ dp, int64_t lddp, const float* __restrict__ p, int64_t ldp,
                                                               int N, float scale) {
    float* drow = dp + (int64_t)blockIdx.x * lddp;
    const float* prow = p + (int64_t)blockIdx.x * ldp;
    float dot = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) dot = fmaf(drow[n], __ldg(prow + n), dot);
    dot = block_sum_bcast(dot);
    for (int n = threadIdx.x; n < N; n += blockDim.x) drow[n] = scale * __ldg(prow + n) * (drow[n] - dot);
}

__global__ void __launch_bounds__(256) add2d_kernel(float* __restrict__ dst, int64_t ldd, const float* __restrict__ src, int64_t lds, int M,
                                                    int N) {
    const int64_t total = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N, n = i - m * N;
        dst[m * ldd + n] += __ldg(src + m * lds + n);
    }
}

static int ew_grid(int64_t total) {
    int64_t g = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_gate_apply(const float* x, int64_t ldx, float* y, int64_t ldy, const float* gate, int M, int N, void* stream) {
    B200_REQUIRE(x && y && gate && M > 0 && N > 0 && ldx >= N && ldy >= N, B200_E_ARG, "gate_apply: bad arguments");
    gate_apply_kernel<<<ew_grid((int64_t)M * N), 256, 0, as_stream(stream)>>>(x, ldx, y, ldy, gate, M, N);
    return launch_status("gate_apply");
}

int64_t b200_gate_grad_work_floats(void) { return 2 + sm_count(); }

int b200_gate_grad(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* gate, float* dgate, float* work, int M, int N,
                   int accumulate, void* stream) {
    B200_REQUIRE(dy && x && gate && dgate && work && M > 0 && N > 0 && lddy >= N && ldx >= N, B200_E_ARG, "gate_grad: bad arguments");
    cudaStream_t st = as_stream(stream);
    int64_t g = ((int64_t)M * N + 255) / 256;
    if (g > sm_count()) g = sm_count();
    // work[0] = ticket (zero between launches), work[1 .. 1+grid) = block partials, work[1 + sm_count] = the ordered sum
    gate_grad_kernel<<<(int)g, 256, 0, st>>>(dy, lddy, x, ldx, work, work + 1 + sm_count(), M, N);
    int rc = launch_status("gate_grad");
    if (rc) return rc;
    gate_grad_finish_kernel<<<1, 1, 0, st>>>(work + 1 + sm_count(), gate, dgate, accumulate);
    return launch_status("gate_grad_finish");
}

int b200_softmax_rows(float* s, int64_t ld, int M, int N, float scale, void* stream) {
    B200_REQUIRE(s && M > 0 && N > 0 && ld >= N, B200_E_ARG, "softmax_rows: bad arguments");
    softmax_rows_kernel<<<M, 256, 0, as_stream(stream)>>>(s, ld, N, scale);
    return launch_status("softmax_rows");
}

int b200_softmax_rows_bwd(float* dp, int64_t lddp, const float* p, int64_t ldp, int M, int N, float scale, void* stream) {
    B200_REQUIRE(dp && p && M > 0 && N > 0 && lddp >= N && ldp >= N, B200_E_ARG, "softmax_rows_bwd: bad arguments");
    softmax_rows_bwd_kernel<<<M, 256, 0, as_stream(stream)>>>(dp, lddp, p, ldp, N, scale);
    return launch_status("softmax_rows_bwd");
}

int b200_add2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int M, int N, void* stream) {
    B200_REQUIRE(dst && src && M > 0 && N > 0 && ld_dst >= N && ld_src >= N, B200_E_ARG, "add2d: bad arguments");
    add2d_kernel<<<ew_grid((int64_t)M * N), 256, 0, as_stream(stream)>>>(dst, ld_dst, src, ld_src, M, N);
    return launch_status("add2d");
}

}  // extern "C"
