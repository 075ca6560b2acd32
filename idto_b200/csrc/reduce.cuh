// Deterministic block reductions (fixed shuffle tree, no atomics).
#pragma once
#include <cuda_runtime.h>

namespace idto {

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// All threads of the block must call; `red` is a shared array of >= 32 doubles.  Every thread
// receives the total.
__device__ __forceinline__ double block_sum(double x, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  x = warp_sum(x);
  __syncthreads();  // protect `red` from a previous call
  if (lane == 0) red[wid] = x;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// N sums at once: one pair of barriers instead of N (each value is added up in exactly the order block_sum uses, so
// the results are bit-identical to N separate calls).  `red` holds >= 32 * N doubles.
template <int N>
__device__ __forceinline__ void block_sum_n(double (&x)[N], double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int j = 0; j < N; ++j) x[j] = warp_sum(x[j]);
  __syncthreads();  // protect `red` from a previous call
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < N; ++j) red[j * 32 + wid] = x[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[j * 32 + i];
    x[j] = t;
  }
}

}  // namespace idto
