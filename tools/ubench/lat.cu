// Latency micro-benchmarks that size the KKT-sweep design (run on the B200: tools/ubench/run.sh).
// Every test is a dependent chain timed with clock64() by one CTA; prints cycles per operation.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

#define N_IT 2048

__global__ void k_dfma(double* out, long long* clk, double a, double b) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// 8 independent chains per thread: issue-rate bound
__global__ void k_dfma8(double* out, long long* clk, double a, double b) {
  double x[8];
  for (int k = 0; k < 8; ++k) x[k] = out[threadIdx.x] + k;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fma(x[k], a, b);
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; ++k) s += x[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_ddiv(double* out, long long* clk, double a) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; ++i) x = a / x;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_drcp(double* out, long long* clk) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; ++i) x = __drcp_rn(x);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_dmax(double* out, long long* clk, double a) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = fmax(fabs(x), a) * 0.999;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* clk) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) p = nxt[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// store -> barrier -> load round trip (what one pivot step pays at least once)
__global__ void k_bar(double* out, long long* clk) {
  __shared__ double buf[512];
  double x = out[threadIdx.x];
  buf[threadIdx.x] = x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; ++i) {
    buf[threadIdx.x] = x;
    __syncthreads();
    x = buf[(threadIdx.x + 33) % blockDim.x] + 1.0;
    __syncthreads();
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_bar_only(double* out, long long* clk) {
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0, out[0] = 0;
}
__global__ void k_redux(double* out, long long* clk) {
  unsigned x = threadIdx.x * 2654435761u;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = __reduce_max_sync(0xffffffffu, x) + threadIdx.x;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* clk) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 5) & 31) + 1.0;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_ballot(double* out, long long* clk) {
  unsigned x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; ++i) x = __ballot_sync(0xffffffffu, (x >> (threadIdx.x & 7)) & 1) + threadIdx.x;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// cluster barrier round trip (2 CTAs)
__global__ void __cluster_dims__(2, 1, 1) k_cluster(double* out, long long* clk) {
  cg::cluster_group cl = cg::this_cluster();
  long long t0 = clock64();
  for (int i = 0; i < 256; ++i) cl.sync();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = (t1 - t0) * (N_IT / 256), out[0] = 0;
}
// global (L2) dependent load latency
__global__ void k_ldg(const int* __restrict__ chain, double* out, long long* clk) {
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < 256; ++i) p = __ldcg(chain + p);
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) clk[0] = (t1 - t0) * (N_IT / 256);
}

template <class F>
void run(const char* name, int threads, F f) {
  double* out;
  long long* clk;
  cudaMalloc(&out, 4096 * 8);
  cudaMalloc(&clk, 64);
  cudaMemset(out, 0, 4096 * 8);
  f(out, clk);
  f(out, clk);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-28s threads=%4d  %8.1f cycles/op   (%s)\n", name, threads, double(h) / N_IT,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out), cudaFree(clk);
}

int main() {
  for (int th : {32, 96, 256}) {
    run("dfma dependent", th, [&](double* o, long long* c) { k_dfma<<<1, th>>>(o, c, 1.0000001, 1e-9); });
    run("dfma 8 chains (per 8 fma)", th, [&](double* o, long long* c) { k_dfma8<<<1, th>>>(o, c, 1.0000001, 1e-9); });
    run("ddiv dependent", th, [&](double* o, long long* c) { k_ddiv<<<1, th>>>(o, c, 1.7); });
    run("drcp dependent", th, [&](double* o, long long* c) { k_drcp<<<1, th>>>(o, c); });
    run("fabs+fmax+mul dependent", th, [&](double* o, long long* c) { k_dmax<<<1, th>>>(o, c, 0.5); });
    run("lds dependent", th, [&](double* o, long long* c) { k_lds<<<1, th>>>(o, c); });
    run("sts+bar+lds+bar", th, [&](double* o, long long* c) { k_bar<<<1, th>>>(o, c); });
    run("bar.sync only", th, [&](double* o, long long* c) { k_bar_only<<<1, th>>>(o, c); });
    run("redux.max dependent", th, [&](double* o, long long* c) { k_redux<<<1, th>>>(o, c); });
    run("shfl dependent (+dadd)", th, [&](double* o, long long* c) { k_shfl<<<1, th>>>(o, c); });
    run("ballot dependent", th, [&](double* o, long long* c) { k_ballot<<<1, th>>>(o, c); });
  }
  run("cluster.sync (2 CTAs)", 256, [&](double* o, long long* c) { k_cluster<<<2, 256>>>(o, c); });
  {
    int n = 1 << 22;  // 16 MB chain: L2 resident, not L1
    int* h = new int[n];
    for (int i = 0; i < n; ++i) h[i] = int((size_t(i) * 1000003u + 12345u) % n);
    int* d;
    cudaMalloc(&d, size_t(n) * 4);
    cudaMemcpy(d, h, size_t(n) * 4, cudaMemcpyHostToDevice);
    run("ldg.cg dependent (L2)", 32, [&](double* o, long long* c) { k_ldg<<<1, 32>>>(d, o, c); });
  }
  return 0;
}
