// Throughput / round-trip micro-benchmarks for the warp-local LU (B200).  One CTA, nw warps.
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 1024
__global__ void k_shfl_thr(double* out, long long* clk, int src) {
  unsigned x[8];
  for (int k = 0; k < 8; ++k) x[k] = threadIdx.x + k;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N_IT; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = __shfl_sync(0xffffffffu, x[k], src) + 1;
  long long t1 = clock64();
  unsigned s = 0;
  for (int k = 0; k < 8; ++k) s += x[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// broadcast LDS.128: all lanes read the same 16 bytes; 8 independent loads per iteration
__global__ void k_lds128_bcast(double* out, long long* clk, int off) {
  __shared__ __align__(16) double buf[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) buf[i] = i;
  __syncthreads();
  double s = 0;
  int o = off;
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N_IT; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double2 v = *reinterpret_cast<const double2*>(buf + ((o + 2 * k) & 510));
      s += v.x + v.y;
    }
    o += 2;
  }
  long long t1 = clock64();
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// one lane stores 16 B, warp barrier, everyone loads it, dependent FMA: the pivot-row broadcast round trip
__global__ void k_sts_lds_rt(double* out, long long* clk, int p) {
  __shared__ __align__(16) double buf[8 * 4 * 2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double x = out[threadIdx.x] + 1.0, y = x + 1.0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; ++i) {
    double* b = buf + w * 8 + (i & 1) * 4;
    if (lane == p) *reinterpret_cast<double2*>(b) = make_double2(x, y);
    __syncwarp();
    const double2 v = *reinterpret_cast<const double2*>(b);
    x = fma(v.x, 0.5, x), y = fma(v.y, 0.5, y);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + y;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// same through SHFL
__global__ void k_shfl_rt(double* out, long long* clk, int p) {
  double x = out[threadIdx.x] + 1.0, y = x + 1.0;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; ++i) {
    const double vx = __shfl_sync(0xffffffffu, x, p), vy = __shfl_sync(0xffffffffu, y, p);
    x = fma(vx, 0.5, x), y = fma(vy, 0.5, y);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + y;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// straight-line code footprint: NI independent-ish FMAs fully unrolled, executed REP times
template <int NI>
__global__ void k_icache(double* out, long long* clk, double a, int rep) {
  double x[8];
  for (int k = 0; k < 8; ++k) x[k] = out[threadIdx.x] + k;
  long long t0 = clock64();
  for (int r = 0; r < rep; ++r) {
#pragma unroll
    for (int i = 0; i < NI; ++i) x[i & 7] = fma(x[i & 7], a, double(i));
  }
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; ++k) s += x[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
template <class F>
void run(const char* name, int threads, double per, F f) {
  double* out;
  long long* clk;
  cudaMalloc(&out, 4096 * 8), cudaMalloc(&clk, 64), cudaMemset(out, 0, 4096 * 8);
  f(out, clk), f(out, clk);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-34s threads=%4d  %8.2f cycles/op   (%s)\n", name, threads, double(h) / per, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out), cudaFree(clk);
}
int main() {
  for (int th : {32, 64, 128, 256}) {
    run("shfl.32 throughput (per shfl/warp)", th, N_IT * 8.0, [&](double* o, long long* c) { k_shfl_thr<<<1, th>>>(o, c, 5); });
    run("lds.128 broadcast (per lds/warp)", th, N_IT * 8.0, [&](double* o, long long* c) { k_lds128_bcast<<<1, th>>>(o, c, 0); });
    run("sts.128(1 lane)+syncwarp+lds.128+fma", th, N_IT, [&](double* o, long long* c) { k_sts_lds_rt<<<1, th>>>(o, c, 7); });
    run("2x shfl.64 + fma round trip", th, N_IT, [&](double* o, long long* c) { k_shfl_rt<<<1, th>>>(o, c, 7); });
  }
  for (int th : {32, 128}) {
    run("straight-line 512 dfma (per instr)", th, 512.0 * 64, [&](double* o, long long* c) { k_icache<512><<<1, th>>>(o, c, 1.0000001, 64); });
    run("straight-line 2048 dfma", th, 2048.0 * 64, [&](double* o, long long* c) { k_icache<2048><<<1, th>>>(o, c, 1.0000001, 64); });
    run("straight-line 8192 dfma", th, 8192.0 * 64, [&](double* o, long long* c) { k_icache<8192><<<1, th>>>(o, c, 1.0000001, 64); });
    run("straight-line 16384 dfma", th, 16384.0 * 64, [&](double* o, long long* c) { k_icache<16384><<<1, th>>>(o, c, 1.0000001, 64); });
  }
  return 0;
}
