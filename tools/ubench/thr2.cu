// Throughput of the cross-lane broadcast primitives (B200), per SM, with 1/4/8 warps.
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 512
#define NK 16
__global__ void k_shfl(double* out, long long* clk, int src) {
  unsigned x[NK];
  for (int k = 0; k < NK; ++k) x[k] = threadIdx.x * 7 + k;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) {
    unsigned y[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) y[k] = __shfl_sync(0xffffffffu, x[k], src);
#pragma unroll
    for (int k = 0; k < NK; ++k) x[k] += y[k];
  }
  long long t1 = clock64();
  unsigned s = 0;
  for (int k = 0; k < NK; ++k) s += x[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_redux(double* out, long long* clk, int src) {
  unsigned x[NK];
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < NK; ++k) x[k] = threadIdx.x * 7 + k;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) {
    unsigned y[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) y[k] = __reduce_or_sync(0xffffffffu, lane == src ? x[k] : 0u);
#pragma unroll
    for (int k = 0; k < NK; ++k) x[k] += y[k];
  }
  long long t1 = clock64();
  unsigned s = 0;
  for (int k = 0; k < NK; ++k) s += x[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
template <typename T>
__global__ void k_lds_bcast(double* out, long long* clk, int off) {
  __shared__ __align__(16) double buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = i;
  __syncthreads();
  double acc[NK];
  for (int k = 0; k < NK; ++k) acc[k] = 0;
  int o = off + (threadIdx.x >> 5) * 64;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) {
    T v[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) v[k] = *reinterpret_cast<const T*>(buf + ((o + 2 * k) & 1022));
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[k] += *reinterpret_cast<double*>(&v[k]);
    o += 2;
  }
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < NK; ++k) s += acc[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
template <typename T>
__global__ void k_sts_one(double* out, long long* clk, int p) {
  __shared__ __align__(16) double buf[8 * 64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T v;
  double* vd = reinterpret_cast<double*>(&v);
  vd[0] = threadIdx.x;
  if (sizeof(T) == 16) vd[1] = 1.0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) {
    if (lane == p) {
#pragma unroll
      for (int k = 0; k < NK; ++k) *reinterpret_cast<T*>(buf + w * 64 + 2 * k) = v;
    }
    __syncwarp();
  }
  long long t1 = clock64();
  out[threadIdx.x] = buf[threadIdx.x & 63];
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
template <class F>
void run(const char* name, int threads, F f) {
  double* out;
  long long* clk;
  cudaMalloc(&out, 4096 * 8), cudaMalloc(&clk, 64), cudaMemset(out, 0, 4096 * 8);
  f(out, clk), f(out, clk);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double per = double(h) / (N_IT * NK);
  printf("%-30s warps=%d  %6.2f cycles/op/warp  -> %5.2f ops/clk/SM  (%s)\n", name, threads / 32, per,
         (threads / 32) / per, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out), cudaFree(clk);
}
int main() {
  for (int th : {32, 128, 256}) {
    run("shfl.32 (uniform src)", th, [&](double* o, long long* c) { k_shfl<<<1, th>>>(o, c, 5); });
    run("redux.or broadcast", th, [&](double* o, long long* c) { k_redux<<<1, th>>>(o, c, 5); });
    run("lds.64 broadcast", th, [&](double* o, long long* c) { k_lds_bcast<double><<<1, th>>>(o, c, 0); });
    run("lds.128 broadcast", th, [&](double* o, long long* c) { k_lds_bcast<double2><<<1, th>>>(o, c, 0); });
    run("sts.64 one lane", th, [&](double* o, long long* c) { k_sts_one<double><<<1, th>>>(o, c, 3); });
    run("sts.128 one lane", th, [&](double* o, long long* c) { k_sts_one<double2><<<1, th>>>(o, c, 3); });
  }
  return 0;
}
