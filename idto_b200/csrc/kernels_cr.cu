// Block cyclic reduction of the time-major KKT system (`linear_solver = IDTO_LINSOLVE_CYCLIC_REDUCTION`).
//
// The parallel-in-time alternative to the two-sided sweep of kernels_kkt3.cu, for the same system (reference:
// PentaDiagonalFactorization, optimizer/penta_diagonal_solver.h:124-248; CalcLagrangeMultipliers cc:1371-1396;
// CalcDoglegPoint cc:2137-2140) with the same results layout (pH, lambda).  Two consecutive block rows of the
// block-PENTA-diagonal matrix form one super-row (size m = 2 kb), which makes it block TRI-diagonal:
//     L_s x_{s-1} + D_s x_s + U_s x_{s+1} = r_s,     s = 0 .. S-1,   S = ceil((T + 1) / 2)
// Level h = 1, 2, 4, ...: the rows that are odd multiples of h are eliminated, all of them in parallel,
//     X_s = D_s^-1 [L_s | U_s | r_s]                                   (k_cr_eliminate: one CTA per row, Gauss-Jordan
//                                                                       with partial pivoting inside the block)
// and every surviving row e (a multiple of 2h) absorbs its two neighbours a = e - h, c = e + h
//     D_e -= L_e XU_a + U_e XL_c,  L_e <- -L_e XL_a,  U_e <- -U_e XU_c,  r_e -= L_e Xr_a + U_e Xr_c   (k_cr_update)
// until only row 0 is left; then x_0 = D_0^-1 r_0 and, level by level downwards,
//     x_s = Xr_s - XL_s x_{s-h} - XU_s x_{s+h}                                                      (k_cr_backsub)
// log2(S) + 1 = 6 levels of independent 50 x 50 eliminations instead of 2 x 21 dependent 25 x 25 ones — but every
// elimination is a 50-step pivot chain with 101 right-hand-side columns, and level 0 alone is 10 of them per problem:
// at the batch the benchmark is quoted on (64 problems, every SM already busy with the sweep's two CTAs per problem)
// this costs more than the sweep (DESIGN.md 4.2 has the measured numbers); it is the better order for a handful of
// problems or long horizons, and the independent cross-check of the sweep in the parity tests.
#include <algorithm>

#include "kkt_view.cuh"
#include "solver.h"

namespace idto {

namespace {

constexpr int kCrThreads = 256;

// 1/x to ~1 ulp without the division subroutine (~400 cycles of dependent latency per pivot): MUFU.RCP64H seed + two
// Newton steps, as in kernels_kkt3.cu.
__device__ __forceinline__ double cr_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// Per-problem workspace (doubles): S super-rows x { D, L, U, XL, XU : m x m column-major; r, Xr, x : m }.
struct CrView {
  double *D, *L, *U, *XL, *XU, *r, *Xr, *x;
  int m, S;
};
__host__ __device__ inline size_t cr_doubles_per_problem(int T, int kb) {
  const size_t m = 2 * size_t(kb), S = (size_t(T) + 2) / 2;
  return S * (5 * m * m + 3 * m);
}
__device__ __forceinline__ CrView make_cr_view(const SolverConsts& sc, const SolverBufs& bf, int b, int kb) {
  CrView v;
  v.m = 2 * kb, v.S = (sc.T + 2) / 2;
  const size_t mm = size_t(v.m) * v.m, S = v.S;
  double* base = bf.crw + size_t(b) * cr_doubles_per_problem(sc.T, kb);
  v.D = base, v.L = v.D + S * mm, v.U = v.L + S * mm, v.XL = v.U + S * mm, v.XU = v.XL + S * mm;
  v.r = v.XU + S * mm, v.Xr = v.r + S * v.m, v.x = v.Xr + S * v.m;
  return v;
}

// Block (i, j) of the time-major KKT matrix, zero outside the band, identity for the padding row i = j = T + 1
// (odd number of block rows).
__device__ __forceinline__ double kkt_entry(const KktView& V, int N, int i, int j, int r, int c) {
  if (i > N || j > N) return (i == j && r == c) ? 1.0 : 0.0;
  if (i < 0 || j < 0) return 0.0;
  const int d = i - j;
  if (d > 2 || d < -2) return 0.0;
  return kkt_blk(V, i, j, r, c);
}

}  // namespace

// Super-rows from the scaled Hessian / Jacobian bands: one CTA per (problem, super-row).
__global__ void __launch_bounds__(kCrThreads) k_cr_build(SolverConsts sc, SolverBufs bf, int kb, int force) {
  const int S = (sc.T + 2) / 2, b = blockIdx.x / S, s = blockIdx.x % S;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const KktView V = make_kkt_view(sc, bf, b);
  const CrView W = make_cr_view(sc, bf, b, kb);
  const int m = W.m, N = sc.T, nq = sc.nq;
  const size_t mm = size_t(m) * m;
  double *D = W.D + s * mm, *L = W.L + s * mm, *U = W.U + s * mm;
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int col = e / m, row = e - col * m;
    const int rb = row / kb, r = row - rb * kb, cb = col / kb, c = col - cb * kb;
    const int i = 2 * s + rb;
    D[e] = kkt_entry(V, N, i, 2 * s + cb, r, c);
    L[e] = kkt_entry(V, N, i, 2 * s - 2 + cb, r, c);
    U[e] = kkt_entry(V, N, i, 2 * s + 2 + cb, r, c);
  }
  const double* gs = bf.gs + size_t(b) * sc.n;
  const double* h = bf.st.h + size_t(b) * sc.nh;
  for (int e = threadIdx.x; e < m; e += blockDim.x) {
    const int rb = e / kb, r = e - rb * kb, i = 2 * s + rb;
    double v = 0.0;
    if (i <= N) v = r < nq ? -gs[i * nq + r] : (i >= 1 ? -h[(i - 1) * sc.nu + (r - nq)] : 0.0);
    W.r[size_t(s) * m + e] = v;
  }
}

// X_s = D_s^-1 [L_s | U_s | r_s] for the rows s = h, 3h, 5h, ... (h = 0: row 0 alone, the top of the reduction):
// Gauss-Jordan with implicit partial pivoting on the m x (3m + 1) tableau in shared memory, one barrier per step.
__global__ void __launch_bounds__(kCrThreads) k_cr_eliminate(SolverConsts sc, SolverBufs bf, int kb, int h, int nrows,
                                                             int force) {
  extern __shared__ __align__(16) double Q[];  // [3m + 1][m] column-major: D | L | U | r
  __shared__ int s_ord[64];
  const int b = blockIdx.x / nrows, k = blockIdx.x % nrows;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const CrView W = make_cr_view(sc, bf, b, kb);
  const int m = W.m, s = h == 0 ? 0 : h * (2 * k + 1), tid = threadIdx.x, lane = tid & 31;
  const size_t mm = size_t(m) * m;
  const int Wc = 3 * m + 1;
  for (int e = tid; e < m * m; e += kCrThreads) {
    Q[e] = W.D[s * mm + e];
    Q[mm + e] = W.L[s * mm + e];
    Q[2 * mm + e] = W.U[s * mm + e];
  }
  for (int e = tid; e < m; e += kCrThreads) Q[3 * mm + e] = W.r[size_t(s) * m + e];
  __syncthreads();
  {
    constexpr int NCG = kCrThreads / 64;
    const int rr = tid & 63, cg = tid >> 6;
    const bool rowok = rr < m;
    bool d0 = lane >= m, d1 = lane + 32 >= m;
    bool bad = false;
    for (int c = 0; c < m; ++c) {
      const double a0 = d0 ? 0.0 : Q[c * m + lane], a1 = d1 ? 0.0 : Q[c * m + lane + 32];
      const unsigned k0 = d0 ? 0u : (((unsigned(__double2hiint(a0)) & 0x7fffffc0u) + 64u) | unsigned(63 - lane));
      const unsigned k1 = d1 ? 0u : (((unsigned(__double2hiint(a1)) & 0x7fffffc0u) + 64u) | unsigned(31 - lane));
      const unsigned mx = __reduce_max_sync(0xffffffffu, k0 > k1 ? k0 : k1);
      const int p = 63 - int(mx & 63u);
      bad |= (mx < 128u) | (mx >= 0x7ff00040u);
      d0 |= p == lane, d1 |= p == lane + 32;
      const double inv = cr_rcp(Q[c * m + p]);
      if (tid == 0) s_ord[p] = c;
      if (rowok && rr != p) {
        const double mult = Q[c * m + rr] * inv;
#pragma unroll 4
        for (int j = c + 1 + cg; j < Wc; j += NCG) Q[j * m + rr] = fma(-mult, Q[j * m + p], Q[j * m + rr]);
      }
      __syncthreads();
    }
    if (bad && tid == 0) atomicExch(bf.status, IDTO_ERR_FACTORIZATION);
  }
  // row p of the tableau now holds unknown s_ord[p]: X[s_ord[p]][j] = Q[.][p] / pivot_p
  __shared__ double s_inv[64];
  if (tid < m) s_inv[tid] = cr_rcp(Q[s_ord[tid] * m + tid]);
  __syncthreads();
  for (int e = tid; e < m * (2 * m + 1); e += kCrThreads) {
    const int j = e / m, p = e - j * m, u = s_ord[p];
    const double val = Q[(m + j) * m + p] * s_inv[p];
    if (j < m)
      W.XL[s * mm + size_t(j) * m + u] = val;
    else if (j < 2 * m)
      W.XU[s * mm + size_t(j - m) * m + u] = val;
    else
      W.Xr[size_t(s) * m + u] = val;
  }
  if (h == 0)
    for (int p = tid; p < m; p += kCrThreads) W.x[s_ord[p]] = Q[3 * mm + p] * s_inv[p];  // x_0
}

// Surviving rows e = 0, 2h, 4h, ... absorb their eliminated neighbours e - h and e + h.
__global__ void __launch_bounds__(kCrThreads) k_cr_update(SolverConsts sc, SolverBufs bf, int kb, int h, int nrows,
                                                          int force) {
  extern __shared__ __align__(16) double sm[];  // L_e | U_e
  const int b = blockIdx.x / nrows, k = blockIdx.x % nrows;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const CrView W = make_cr_view(sc, bf, b, kb);
  const int m = W.m, S = W.S, e = 2 * h * k, a = e - h, c = e + h, tid = threadIdx.x;
  const size_t mm = size_t(m) * m;
  const bool ha = a >= 0, hc = c < S;
  double *Le = sm, *Ue = sm + mm;
  for (int i = tid; i < m * m; i += kCrThreads) Le[i] = W.L[e * mm + i], Ue[i] = W.U[e * mm + i];
  __syncthreads();
  const double *XLa = W.XL + (ha ? a : 0) * mm, *XUa = W.XU + (ha ? a : 0) * mm, *Xra = W.Xr + size_t(ha ? a : 0) * m;
  const double *XLc = W.XL + (hc ? c : 0) * mm, *XUc = W.XU + (hc ? c : 0) * mm, *Xrc = W.Xr + size_t(hc ? c : 0) * m;
  for (int i = tid; i < m * m; i += kCrThreads) {
    const int col = i / m, row = i - col * m;
    double dD = 0.0, nL = 0.0, nU = 0.0;
    if (ha) {
      double s0 = 0.0, s1 = 0.0;
      for (int j = 0; j < m; ++j) {
        const double l = Le[j * m + row];
        s0 = fma(l, XUa[size_t(col) * m + j], s0);
        s1 = fma(l, XLa[size_t(col) * m + j], s1);
      }
      dD += s0, nL = -s1;
    }
    if (hc) {
      double s0 = 0.0, s1 = 0.0;
      for (int j = 0; j < m; ++j) {
        const double u = Ue[j * m + row];
        s0 = fma(u, XLc[size_t(col) * m + j], s0);
        s1 = fma(u, XUc[size_t(col) * m + j], s1);
      }
      dD += s0, nU = -s1;
    }
    W.D[e * mm + i] -= dD;
    W.L[e * mm + i] = nL;
    W.U[e * mm + i] = nU;
  }
  for (int row = tid; row < m; row += kCrThreads) {
    double acc = 0.0;
    if (ha)
      for (int j = 0; j < m; ++j) acc = fma(Le[j * m + row], Xra[j], acc);
    if (hc)
      for (int j = 0; j < m; ++j) acc = fma(Ue[j * m + row], Xrc[j], acc);
    W.r[size_t(e) * m + row] -= acc;
  }
}

// x_s = Xr_s - XL_s x_{s-h} - XU_s x_{s+h} for the rows eliminated at level h; the last level (h = 1) also
// scatters the solution into pH (positions) and lambda.
__global__ void __launch_bounds__(64) k_cr_backsub(SolverConsts sc, SolverBufs bf, int kb, int h, int nrows, int force) {
  const int b = blockIdx.x / nrows, k = blockIdx.x % nrows;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const CrView W = make_cr_view(sc, bf, b, kb);
  const int m = W.m, S = W.S, s = h * (2 * k + 1), a = s - h, c = s + h, row = threadIdx.x;
  const size_t mm = size_t(m) * m;
  if (row >= m) return;
  double acc = W.Xr[size_t(s) * m + row];
  const double *xa = W.x + size_t(a) * m, *xc = W.x + size_t(c < S ? c : 0) * m;
  for (int j = 0; j < m; ++j) acc = fma(-W.XL[s * mm + size_t(j) * m + row], xa[j], acc);
  if (c < S)
    for (int j = 0; j < m; ++j) acc = fma(-W.XU[s * mm + size_t(j) * m + row], xc[j], acc);
  W.x[size_t(s) * m + row] = acc;
}

__global__ void __launch_bounds__(kCrThreads) k_cr_emit(SolverConsts sc, SolverBufs bf, int kb, int force) {
  const int b = blockIdx.x;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const CrView W = make_cr_view(sc, bf, b, kb);
  const int nq = sc.nq, N = sc.T;
  double* xq = bf.pH + size_t(b) * sc.n;
  double* lam = bf.lambda + size_t(b) * sc.nh;
  for (int e = threadIdx.x; e < (N + 1) * kb; e += blockDim.x) {
    const int i = e / kb, r = e - i * kb;
    const double v = W.x[e];  // super-rows are consecutive pairs of block rows: x is already in block-row order
    if (r < nq)
      xq[i * nq + r] = v;
    else if (i >= 1)
      lam[(i - 1) * sc.nu + (r - nq)] = v;
  }
}

size_t cr_workspace_doubles(int B, int T, int kb) { return size_t(B) * cr_doubles_per_problem(T, kb); }

bool launch_kkt_cr(int kb, const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  if (!bf.crw || 2 * kb > 64) return false;
  const int m = 2 * kb, S = (sc.T + 2) / 2, f = force ? 1 : 0;
  const size_t elim_smem = size_t(3 * m + 1) * m * sizeof(double), upd_smem = size_t(2) * m * m * sizeof(double);
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_cr_eliminate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_cr_update, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  k_cr_build<<<sc.B * S, kCrThreads, 0, stream>>>(sc, bf, kb, f);
  g_launch_counter += 1;
  int h = 1;
  for (; h < S; h *= 2) {
    const int nel = (S - h + 2 * h - 1) / (2 * h);  // rows h, 3h, 5h, ... < S
    const int nup = (S + 2 * h - 1) / (2 * h);      // rows 0, 2h, 4h, ... < S
    k_cr_eliminate<<<sc.B * nel, kCrThreads, elim_smem, stream>>>(sc, bf, kb, h, nel, f);
    k_cr_update<<<sc.B * nup, kCrThreads, upd_smem, stream>>>(sc, bf, kb, h, nup, f);
    g_launch_counter += 2;
  }
  k_cr_eliminate<<<sc.B, kCrThreads, elim_smem, stream>>>(sc, bf, kb, 0, 1, f);  // x_0
  g_launch_counter += 1;
  for (h /= 2; h >= 1; h /= 2) {
    const int nel = (S - h + 2 * h - 1) / (2 * h);
    k_cr_backsub<<<sc.B * nel, 64, 0, stream>>>(sc, bf, kb, h, nel, f);
    g_launch_counter += 1;
  }
  k_cr_emit<<<sc.B, kCrThreads, 0, stream>>>(sc, bf, kb, f);
  g_launch_counter += 1;
  return true;
}

}  // namespace idto
