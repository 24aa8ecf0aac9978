// k-nearest-neighbour classification of frozen encoder features: the evaluation the reference runs after
// pre-training, training_structures/dino_train.py:349-368 (sklearn KNeighborsClassifier(n_neighbors=5): Euclidean distance,
// uniform weights, majority vote, ties -> smallest class label) on ~55k x 256 train and ~10k x 256 test features.
//
//   ||a - b||^2 = ||a||^2 - 2 (a.b - ||b||^2 / 2):   score(i, j) = a_i . b_j - ||b_j||^2 / 2   (larger = nearer)
//
// is one exact-fp32 GEMM with a bias (b200_linear_fwd, M = test rows, N = train rows); b200_knn_neg_half_sqnorm builds the bias
// and b200_knn_topk_vote selects the k best columns of every score row (one warp per row: per-lane sorted top-k in registers,
// k rounds of warp arg-max merging) and votes.  Ties between equal scores go to the smaller train index.
#include "common.cuh"

namespace b200 {
namespace {

constexpr int KMAX = 16;

__global__ void __launch_bounds__(256) neg_half_sqnorm_kernel(const float* __restrict__ x, int64_t ldx, int N, int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + warp; r < N; r += gridDim.x * 8) {
        float s = 0.f;
        for (int k = lane; k < D; k += 32) {
            const float v = __ldg(x + (size_t)r * ldx + k);
            s = fmaf(v, v, s);
        }
        s = warp_sum(s);
        if (lane == 0) out[r] = -0.5f * s;
    }
}

// better(a, b): a is a nearer neighbour than b
__device__ __forceinline__ bool better(float sa, int ia, float sb, int ib) { return sa > sb || (sa == sb && ia < ib); }

__global__ void __launch_bounds__(256) knn_topk_vote_kernel(const float* __restrict__ scores, int64_t lds, const int64_t* __restrict__ labels,
                                                            int M, int N, int k, int n_classes, int64_t* __restrict__ pred,
                                                            int32_t* __restrict__ nbr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + warp; r < M; r += gridDim.x * 8) {
        const float* row = scores + (size_t)r * lds;
        float bs[KMAX];
        int bi[KMAX];
#pragma unroll
        for (int t = 0; t < KMAX; ++t) {
            bs[t] = -INFINITY;
            bi[t] = 0x7fffffff;
        }
        float thr_s = -INFINITY;        // the k-th best of this lane so far (everything not better than it is skipped)
        int thr_i = 0x7fffffff;
        for (int j = lane; j < N; j += 32) {
            float s = __ldg(row + j);
            int idx = j;
            if (better(s, idx, thr_s, thr_i)) {
#pragma unroll
                for (int t = 0; t < KMAX; ++t) {            // insertion into the sorted list (register-resident: fully unrolled)
                    if (t < k && better(s, idx, bs[t], bi[t])) {
                        const float ts = bs[t];
                        const int ti = bi[t];
                        bs[t] = s;
                        bi[t] = idx;
                        s = ts;
                        idx = ti;
                    }
                    if (t == k - 1) {
                        thr_s = bs[t];
                        thr_i = bi[t];
                    }
                }
            }
        }
        // merge the 32 sorted lists: k rounds, each takes the best head of all lanes
        int votes = 0;              // lane c counts the votes of class c (n_classes <= 32)
        for (int round = 0; round < k; ++round) {
            float s = bs[0];
            int idx = bi[0];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float so = __shfl_xor_sync(0xffffffffu, s, o);
                const int io = __shfl_xor_sync(0xffffffffu, idx, o);
                if (better(so, io, s, idx)) {
                    s = so;
                    idx = io;
                }
            }
            if (bi[0] == idx && idx != 0x7fffffff) {         // the owning lane pops its head
#pragma unroll
                for (int t = 0; t + 1 < KMAX; ++t) {
                    bs[t] = bs[t + 1];
                    bi[t] = bi[t + 1];
                }
                bs[KMAX - 1] = -INFINITY;
                bi[KMAX - 1] = 0x7fffffff;
            }
            if (idx != 0x7fffffff) {
                const int c = (int)__ldg(labels + idx);
                if (lane == c) ++votes;
                if (nbr != nullptr && lane == 0) nbr[(size_t)r * k + round] = idx;
            } else if (nbr != nullptr && lane == 0) {
                nbr[(size_t)r * k + round] = -1;
            }
        }
        // majority vote, ties -> smallest class label
        int best_v = (lane < n_classes) ? votes : -1, best_c = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int vo = __shfl_xor_sync(0xffffffffu, best_v, o), co = __shfl_xor_sync(0xffffffffu, best_c, o);
            if (vo > best_v || (vo == best_v && co < best_c)) {
                best_v = vo;
                best_c = co;
            }
        }
        if (lane == 0) pred[r] = best_c;
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_knn_neg_half_sqnorm(const float* x, int64_t ldx, int N, int D, float* out, void* stream) {
    B200_REQUIRE(x && out && N > 0 && D > 0 && ldx >= D, B200_E_ARG, "knn_neg_half_sqnorm: bad arguments");
    int grid = (N + 7) / 8;
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    neg_half_sqnorm_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, ldx, N, D, out);
    return launch_status("knn_neg_half_sqnorm");
}

int b200_knn_topk_vote(const float* scores, int64_t lds, const int64_t* labels, int M, int N, int k, int n_classes, int64_t* pred,
                       int32_t* neighbours, void* stream) {
    B200_REQUIRE(scores && labels && pred && M > 0 && N > 0 && lds >= N, B200_E_ARG, "knn_topk_vote: bad arguments");
    B200_REQUIRE(k >= 1 && k <= KMAX && k <= N, B200_E_SHAPE, "knn_topk_vote: k=%d must be in [1, %d] and <= N", k, KMAX);
    B200_REQUIRE(n_classes >= 1 && n_classes <= 32, B200_E_SHAPE, "knn_topk_vote: n_classes=%d must be in [1, 32]", n_classes);
    int grid = (M + 7) / 8;
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    knn_topk_vote_kernel<<<grid, 256, 0, as_stream(stream)>>>(scores, lds, labels, M, N, k, n_classes, pred, neighbours);
    return launch_status("knn_topk_vote");
}

}  // extern "C"
