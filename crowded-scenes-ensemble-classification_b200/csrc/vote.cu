// Ensemble soft vote on the GPU.
//
// Replaces ensemble_predictions (evaluate_ensemble.py:343-370):
//   weights ndarray : summed = np.tensordot(yhats[M,N,C], w[M], axes=(0,0)); argmax(axis=1)
//   "MAXIMUM"       : argmax over the member-major [N, M*C] rows, modulo C
// and the inner loop of grid_search / differential evolution (evaluate_ensemble.py:302-339).
// Arithmetic is fp64 like the reference (its probabilities are float64 values parsed from the
// CSV text), members accumulated in member order with separately rounded multiply and add
// (no FMA contraction); argmax ties go to the lowest index (np.argmax).
#include "common.cuh"

namespace cse {

constexpr int VOTE_MAX_C = 64;
constexpr int VOTE_BLOCK = 128;

template <typename T>
__global__ void __launch_bounds__(VOTE_BLOCK)
vote_kernel(const T* __restrict__ probs, const double* __restrict__ weights, int mode, int M, int N, int C,
            int32_t* __restrict__ pred, double* __restrict__ summed, int rows_per_block) {
  extern __shared__ double tile[];   // [rows_per_block * C]
  const int n0 = blockIdx.x * rows_per_block;
  const int rows = min(rows_per_block, N - n0);
  const int tid = threadIdx.x;
  double acc[VOTE_MAX_C];
  double best = -INFINITY;
  int best_idx = 0;
#pragma unroll 1
  for (int c = 0; c < C; ++c) acc[c] = 0.0;
  for (int m = 0; m < M; ++m) {
    const T* src = probs + ((long long)m * N + n0) * C;
    __syncthreads();
    for (int i = tid; i < rows * C; i += VOTE_BLOCK) tile[i] = (double)src[i];   // coalesced
    __syncthreads();
    if (tid < rows) {
      if (mode == 0) {
        const double w = weights ? weights[m] : 1.0;
        for (int c = 0; c < C; ++c) acc[c] = __dadd_rn(acc[c], __dmul_rn(tile[tid * C + c], w));
      } else {
        for (int c = 0; c < C; ++c) {
          double v = tile[tid * C + c];
          if (v > best) { best = v; best_idx = c; }
        }
      }
    }
  }
  if (tid < rows) {
    if (mode == 0) {
      for (int c = 0; c < C; ++c) {
        if (acc[c] > best) { best = acc[c]; best_idx = c; }
        if (summed) summed[(long long)(n0 + tid) * C + c] = acc[c];
      }
    }
    pred[n0 + tid] = best_idx;
  }
}

int vote_launch(const void* probs, int is_f64, const double* weights, int mode, int M, int N, int C,
                int32_t* pred, double* summed, cudaStream_t st) {
  CSE_REQUIRE(M >= 1 && N >= 0 && C >= 1 && C <= VOTE_MAX_C, "vote: M=%d N=%d C=%d (C <= %d)", M, N, C, VOTE_MAX_C);
  CSE_REQUIRE(mode == 0 || mode == 1, "vote: mode %d", mode);
  if (N == 0) return CSE_OK;
  // the fp64 tile stays inside the 48 KB every kernel gets without an opt-in: C <= 48 -> 128 clips per block, else 64
  const int rpb = (size_t)VOTE_BLOCK * C * sizeof(double) <= 48 * 1024 ? VOTE_BLOCK : VOTE_BLOCK / 2;
  const int blocks = ceil_div(N, rpb);
  const size_t smem = (size_t)rpb * C * sizeof(double);
  if (is_f64)
    vote_kernel<double><<<blocks, VOTE_BLOCK, smem, st>>>((const double*)probs, weights, mode, M, N, C, pred, summed, rpb);
  else
    vote_kernel<float><<<blocks, VOTE_BLOCK, smem, st>>>((const float*)probs, weights, mode, M, N, C, pred, summed, rpb);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

// One block per candidate weight vector; threads stride over clips.
__global__ void __launch_bounds__(256)
vote_search_kernel(const double* __restrict__ probs, const double* __restrict__ weights,
                   const int32_t* __restrict__ labels, int M, int N, int C, int32_t* __restrict__ correct) {
  __shared__ double w[64];
  __shared__ int warp_sums[8];
  const int wi = blockIdx.x;
  for (int m = threadIdx.x; m < M; m += blockDim.x) w[m] = weights[(long long)wi * M + m];
  __syncthreads();
  int hits = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    double best = -INFINITY;
    int best_idx = 0;
    for (int c = 0; c < C; ++c) {
      double s = 0.0;
      for (int m = 0; m < M; ++m) s = __dadd_rn(s, __dmul_rn(probs[((long long)m * N + n) * C + c], w[m]));
      if (s > best) { best = s; best_idx = c; }
    }
    hits += (best_idx == labels[n]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
  if (threadIdx.x % 32 == 0) warp_sums[threadIdx.x / 32] = hits;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += warp_sums[i];
    correct[wi] = t;
  }
}

int vote_search_launch(const double* probs, const double* weights, const int32_t* labels, int W, int M, int N,
                       int C, int32_t* correct, cudaStream_t st) {
  CSE_REQUIRE(W >= 0 && M >= 1 && M <= 64 && N >= 0 && C >= 1, "vote_search: W=%d M=%d N=%d C=%d (M <= 64)", W, M, N, C);
  if (W == 0) return CSE_OK;
  vote_search_kernel<<<W, 256, 0, st>>>(probs, weights, labels, M, N, C, correct);
  CSE_CUDA(cudaGetLastError());
  return CSE_OK;
}

}  // namespace cse
