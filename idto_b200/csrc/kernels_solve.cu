// Block penta-diagonal factorisation / solves and the Lagrange-multiplier pipeline.
//
// Reference: PentaDiagonalFactorization::Factorize / SolveInPlace (optimizer/penta_diagonal_solver.h:
// 124-248, block Thomas after Benkert & Fischer 2007) and CalcLagrangeMultipliers
// (optimizer/trajectory_optimizer.cc:1371-1396).  This file is the baseline "Thomas" path
// (IDTO_LINSOLVE_THOMAS): same elimination order as the reference, one CTA per problem, with the
// diagonal blocks inverted explicitly by Gauss-Jordan with partial pivoting (the reference keeps a
// PartialPivLU per block) so that every later step is a small dense GEMM.
#include "reduce.cuh"
#include "solver.h"

namespace idto {

namespace {

// C(k x k) = alpha*C0 - A*B, all column-major k x k in shared memory; threads cooperate.
__device__ __forceinline__ void gemm_sub(double* C, const double* C0, const double* A, const double* Bm, int k,
                                         int tid, int nt) {
  for (int e = tid; e < k * k; e += nt) {
    const int c = e / k, r = e % k;
    double acc = 0.0;
    for (int j = 0; j < k; ++j) acc += A[j * k + r] * Bm[c * k + j];
    C[e] = C0[e] - acc;
  }
}
__device__ __forceinline__ void gemm(double* C, const double* A, const double* Bm, int k, int tid, int nt) {
  for (int e = tid; e < k * k; e += nt) {
    const int c = e / k, r = e % k;
    double acc = 0.0;
    for (int j = 0; j < k; ++j) acc += A[j * k + r] * Bm[c * k + j];
    C[e] = acc;
  }
}

}  // namespace

// Factorize (penta_diagonal_solver.h:124-197): for each block row i
//   K_i = B_i - A_i Y_{i-2};  G_i = C_i - A_i Z_{i-2} - K_i Y_{i-1};
//   Y_i = G_i^-1 (D_i - K_i Z_{i-1});  Z_i = G_i^-1 E_i     with D_i = B_{i+1}^T, E_i = A_{i+2}^T.
__global__ void __launch_bounds__(256) k_factor(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_piv;
  __shared__ int s_fail;
  const int b = blockIdx.x;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int nblk = sc.T + 1, k = sc.nq, kk = k * k, tid = threadIdx.x, nt = blockDim.x;
  double* Ym2 = sm;         // Y_{i-2}
  double* Ym1 = Ym2 + kk;   // Y_{i-1}
  double* Zm2 = Ym1 + kk;
  double* Zm1 = Zm2 + kk;
  double* sA = Zm1 + kk;
  double* sK = sA + kk;
  double* sG = sK + kk;     // [G | I] augmented, k x 2k column-major
  double* sY = sG + 2 * kk; // rhs for Y
  double* sE = sY + kk;     // rhs for Z
  double* sT = sE + kk;     // temp
  const size_t base = size_t(b) * nblk * kk;
  for (int e = tid; e < 4 * kk; e += nt) sm[e] = 0.0;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  for (int i = 0; i < nblk; ++i) {
    const double* gA = bf.SA + base + size_t(i) * kk;
    const double* gB = bf.SB + base + size_t(i) * kk;
    const double* gC = bf.SC + base + size_t(i) * kk;
    for (int e = tid; e < kk; e += nt) {
      const int c = e / k, r = e % k;
      sA[e] = gA[e];
      sK[e] = gB[e];
      sG[e] = gC[e];
      sG[kk + e] = (r == c) ? 1.0 : 0.0;
      sY[e] = (i < nblk - 1) ? gB[kk + r * k + c] : 0.0;      // D_i = B_{i+1}^T
      sE[e] = (i < nblk - 2) ? gA[2 * kk + r * k + c] : 0.0;  // E_i = A_{i+2}^T
    }
    __syncthreads();
    gemm_sub(sT, sK, sA, Ym2, k, tid, nt);  // K_i = B_i - A_i Y_{i-2}
    gemm_sub(sG, sG, sA, Zm2, k, tid, nt);  // G_i = C_i - A_i Z_{i-2}   (in place: own element only)
    __syncthreads();
    for (int e = tid; e < kk; e += nt) sK[e] = sT[e];
    __syncthreads();
    gemm_sub(sG, sG, sK, Ym1, k, tid, nt);  // G_i -= K_i Y_{i-1}
    gemm_sub(sY, sY, sK, Zm1, k, tid, nt);  // Yrhs = D_i - K_i Z_{i-1}
    __syncthreads();
    // Gauss-Jordan with partial pivoting on [G | I]
    for (int c = 0; c < k; ++c) {
      if (tid < 32) {
        double best = -1.0;
        int bi = c;
        for (int r = c + tid; r < k; r += 32) {
          const double x = fabs(sG[c * k + r]);
          if (x > best) best = x, bi = r;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
        }
        if (tid == 0) {
          s_piv = bi;
          if (!(best > 0.0)) s_fail = 1;
        }
      }
      __syncthreads();
      const int p = s_piv;
      if (p != c)
        for (int j = tid; j < 2 * k; j += nt) {
          const double x = sG[j * k + c];
          sG[j * k + c] = sG[j * k + p];
          sG[j * k + p] = x;
        }
      __syncthreads();
      const double inv = 1.0 / sG[c * k + c];
      __syncthreads();
      for (int j = tid; j < 2 * k; j += nt) sG[j * k + c] *= inv;
      __syncthreads();
      // eliminate column c from every other row; the multiplier column is read before it is zeroed
      for (int e = tid; e < 2 * kk; e += nt) {
        const int j = e / k, r = e % k;
        if (r != c && j != c) sG[e] -= sG[c * k + r] * sG[j * k + c];
      }
      __syncthreads();
      for (int r = tid; r < k; r += nt)
        if (r != c) sG[c * k + r] = 0.0;
      __syncthreads();
    }
    double* sGinv = sG + kk;
    gemm(sT, sGinv, sY, k, tid, nt);   // Y_i
    gemm(sA, sGinv, sE, k, tid, nt);   // Z_i (sA is free now)
    __syncthreads();
    double* oK = bf.FK + base + size_t(i) * kk;
    double* oG = bf.FG + base + size_t(i) * kk;
    double* oY = bf.FY + base + size_t(i) * kk;
    double* oZ = bf.FZ + base + size_t(i) * kk;
    for (int e = tid; e < kk; e += nt) {
      oK[e] = sK[e];
      oG[e] = sGinv[e];
      oY[e] = sT[e];
      oZ[e] = sA[e];
      Ym2[e] = Ym1[e], Zm2[e] = Zm1[e];
    }
    __syncthreads();
    for (int e = tid; e < kk; e += nt) Ym1[e] = sT[e], Zm1[e] = sA[e];
    __syncthreads();
  }
  if (tid == 0 && s_fail) atomicExch(bf.status, IDTO_ERR_FACTORIZATION);
}

// SolveInPlace (penta_diagonal_solver.h:199-248) for `ncols` right-hand sides of problem b stored as
// X[b][col][n].  mode 0: RHS already in X.  mode 1: RHS column (t,u) = row (t,u) of J~ (generated on
// the fly from the Jm/Jt/Jp bands).  Each CTA owns `cpb` columns; thread = (column, row in block).
__global__ void __launch_bounds__(256) k_penta_solve(SolverConsts sc, SolverBufs bf, double* Xbase, int ncols,
                                                     int cpb, int mode, int force) {
  extern __shared__ __align__(16) double sm[];
  const int ctas_per_b = (ncols + cpb - 1) / cpb;
  const int b = blockIdx.x / ctas_per_b, c0 = (blockIdx.x % ctas_per_b) * cpb;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int nblk = sc.T + 1, k = sc.nq, kk = k * k, n = sc.n, tid = threadIdx.x, nt = blockDim.x;
  double* sM0 = sm;            // A_i  / Y_i
  double* sM1 = sM0 + kk;      // K_i  / Z_i
  double* sM2 = sM1 + kk;      // G_i^-1
  double* x1 = sM2 + kk;       // [cpb][k] r_{i-1} / x_{i+1}
  double* x2 = x1 + cpb * k;   // r_{i-2} / x_{i+2}
  double* tt = x2 + cpb * k;   // temp
  const int lc = tid / k, r = tid % k;
  const bool act = lc < cpb && (c0 + lc) < ncols;
  const int col = c0 + lc;
  double* X = Xbase + (size_t(b) * ncols + (act ? col : 0)) * n;
  const size_t fb = size_t(b) * nblk * kk;
  const int nu = sc.nu;
  const int ct = mode == 1 && act ? col / nu : 0, cu = mode == 1 && act ? col % nu : 0;
  const double* Jm = bf.Jm + (size_t(b) * sc.T + ct) * nu * k + size_t(cu) * k;
  const double* Jt = bf.Jt + (size_t(b) * sc.T + ct) * nu * k + size_t(cu) * k;
  const double* Jp = bf.Jp + (size_t(b) * sc.T + ct) * nu * k + size_t(cu) * k;
  for (int e = tid; e < 2 * cpb * k; e += nt) x1[e] = 0.0;
  __syncthreads();
  // forward sweep: r_i = G_i^-1 (b_i - A_i r_{i-2} - K_i r_{i-1})
  for (int i = 0; i < nblk; ++i) {
    for (int e = tid; e < kk; e += nt) {
      sM0[e] = bf.SA[fb + size_t(i) * kk + e];
      sM1[e] = bf.FK[fb + size_t(i) * kk + e];
      sM2[e] = bf.FG[fb + size_t(i) * kk + e];
    }
    __syncthreads();
    if (act) {
      double rhs;
      if (mode == 1)
        rhs = (i == ct - 1) ? Jm[r] : (i == ct) ? Jt[r] : (i == ct + 1) ? Jp[r] : 0.0;
      else
        rhs = X[size_t(i) * k + r];
      double acc = rhs;
      for (int j = 0; j < k; ++j) acc -= sM0[j * k + r] * x2[lc * k + j];
      for (int j = 0; j < k; ++j) acc -= sM1[j * k + r] * x1[lc * k + j];
      tt[lc * k + r] = acc;
    }
    __syncthreads();
    double out = 0.0;
    if (act) {
      for (int j = 0; j < k; ++j) out += sM2[j * k + r] * tt[lc * k + j];
      X[size_t(i) * k + r] = out;
    }
    __syncthreads();
    if (act) x2[lc * k + r] = x1[lc * k + r];
    __syncthreads();
    if (act) x1[lc * k + r] = out;
    __syncthreads();
  }
  // backward sweep: x_i = r_i - Y_i x_{i+1} - Z_i x_{i+2}
  for (int e = tid; e < 2 * cpb * k; e += nt) x1[e] = 0.0;
  __syncthreads();
  for (int i = nblk - 1; i >= 0; --i) {
    for (int e = tid; e < kk; e += nt) {
      sM0[e] = bf.FY[fb + size_t(i) * kk + e];
      sM1[e] = bf.FZ[fb + size_t(i) * kk + e];
    }
    __syncthreads();
    double out = 0.0;
    if (act) {
      out = X[size_t(i) * k + r];
      for (int j = 0; j < k; ++j) out -= sM0[j * k + r] * x1[lc * k + j];
      for (int j = 0; j < k; ++j) out -= sM1[j * k + r] * x2[lc * k + j];
      X[size_t(i) * k + r] = out;
    }
    __syncthreads();
    if (act) x2[lc * k + r] = x1[lc * k + r];
    __syncthreads();
    if (act) x1[lc * k + r] = out;
    __syncthreads();
  }
}

// S = J~ (H~^-1 J~^T) and rhs = h - (H~^-1 J~^T)^T g~ (cc:1395); one CTA per (b, column c).
__global__ void __launch_bounds__(128) k_schur(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double red[32];
  const int b = blockIdx.x / sc.nh, c = blockIdx.x % sc.nh;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int n = sc.n, nh = sc.nh, nu = sc.nu, k = sc.nq, T = sc.T, tid = threadIdx.x, nt = blockDim.x;
  const double* X = bf.X + (size_t(b) * nh + c) * n;
  for (int e = tid; e < n; e += nt) sm[e] = X[e];
  __syncthreads();
  for (int row = tid; row < nh; row += nt) {
    const int t = row / nu, u = row % nu;
    const size_t jb = (size_t(b) * T + t) * nu * k + size_t(u) * k;
    double acc = 0.0;
    if (t > 1)
      for (int j = 0; j < k; ++j) acc += bf.Jm[jb + j] * sm[(t - 1) * k + j];
    if (t > 0)
      for (int j = 0; j < k; ++j) acc += bf.Jt[jb + j] * sm[t * k + j];
    for (int j = 0; j < k; ++j) acc += bf.Jp[jb + j] * sm[(t + 1) * k + j];
    bf.S[size_t(b) * nh * nh + size_t(c) * nh + row] = acc;
  }
  double part = 0.0;
  const double* gs = bf.gs + size_t(b) * n;
  for (int e = tid; e < n; e += nt) part += sm[e] * gs[e];
  part = block_sum(part, red);
  if (tid == 0) bf.rhs[size_t(b) * nh + c] = bf.st.h[size_t(b) * nh + c] - part;
}

// lambda = S^-1 rhs by LDL^T with diagonal pivoting (stands in for Eigen's LDLT, cc:1395); one CTA
// per problem, S kept full-symmetric in global memory (L2 resident).
__global__ void __launch_bounds__(512) k_ldlt(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_p;
  __shared__ double red[32];
  __shared__ int redi[32];
  const int b = blockIdx.x;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int n = sc.nh, tid = threadIdx.x, nt = blockDim.x;
  double* A = bf.S + size_t(b) * n * n;  // column-major, A(i,j) = A[j*n+i]
  double* y = sm;                        // [n]
  double* colk = sm + n;                 // [n] scaled column
  int* perm = reinterpret_cast<int*>(sm + 2 * n);
  for (int i = tid; i < n; i += nt) perm[i] = i;
  __syncthreads();
  for (int kx = 0; kx < n; ++kx) {
    // pivot: first largest |A(i,i)|, i >= kx
    double best = -1.0;
    int bi = kx;
    for (int i = kx + tid; i < n; i += nt) {
      const double x = fabs(A[size_t(i) * n + i]);
      if (x > best) best = x, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
    }
    if ((tid & 31) == 0) red[tid >> 5] = best, redi[tid >> 5] = bi;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < (nt + 31) / 32; ++w)
        if (red[w] > best || (red[w] == best && redi[w] < bi)) best = red[w], bi = redi[w];
      s_p = bi;
    }
    __syncthreads();
    const int p = s_p;
    if (p != kx) {
      for (int j = tid; j < n; j += nt) {  // swap rows kx, p
        const double x = A[size_t(j) * n + kx];
        A[size_t(j) * n + kx] = A[size_t(j) * n + p];
        A[size_t(j) * n + p] = x;
      }
      __syncthreads();
      for (int i = tid; i < n; i += nt) {  // swap columns kx, p
        const double x = A[size_t(kx) * n + i];
        A[size_t(kx) * n + i] = A[size_t(p) * n + i];
        A[size_t(p) * n + i] = x;
      }
      if (tid == 0) {
        const int x = perm[kx];
        perm[kx] = perm[p];
        perm[p] = x;
      }
      __syncthreads();
    }
    const double d = A[size_t(kx) * n + kx];
    if (d == 0.0) continue;  // uniform across the CTA
    for (int i = kx + 1 + tid; i < n; i += nt) {
      const double l = A[size_t(kx) * n + i] / d;
      A[size_t(kx) * n + i] = l;
      colk[i] = l;
    }
    __syncthreads();
    const int m = n - kx - 1;
    for (int e = tid; e < m * m; e += nt) {
      const int j = kx + 1 + e / m, i = kx + 1 + e % m;
      A[size_t(j) * n + i] -= colk[i] * (colk[j] * d);
    }
    __syncthreads();
  }
  // solve: y = P b; L y = y; D; L^T; un-permute
  const double* rhs = bf.rhs + size_t(b) * n;
  for (int i = tid; i < n; i += nt) y[i] = rhs[perm[i]];
  __syncthreads();
  for (int kx = 0; kx < n; ++kx) {
    const double yk = y[kx];
    for (int i = kx + 1 + tid; i < n; i += nt) y[i] -= A[size_t(kx) * n + i] * yk;
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) {
    const double d = A[size_t(i) * n + i];
    y[i] = d != 0.0 ? y[i] / d : 0.0;
  }
  __syncthreads();
  for (int kx = n - 1; kx >= 0; --kx) {
    // y[kx] -= sum_{i>kx} L(i,kx) y[i]
    double part = 0.0;
    for (int i = kx + 1 + tid; i < n; i += nt) part += A[size_t(kx) * n + i] * y[i];
    part = block_sum(part, red);
    if (tid == 0) y[kx] -= part;
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) bf.lambda[size_t(b) * n + perm[i]] = y[i];
}

// gm = g~ + J~^T lambda (cc:1442) and merit = L + h.lambda (cc:1418); one CTA per problem.
__global__ void __launch_bounds__(256) k_merit(SolverConsts sc, SolverBufs bf, int force) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int n = sc.n, nh = sc.nh, nu = sc.nu, k = sc.nq, T = sc.T, tid = threadIdx.x, nt = blockDim.x;
  const double* gs = bf.gs + size_t(b) * n;
  double* gm = bf.gm + size_t(b) * n;
  if (!sc.eq || nh == 0) {
    for (int e = tid; e < n; e += nt) gm[e] = gs[e];
    if (tid == 0) bf.merit[b] = bf.st.cost[b];
    return;
  }
  const double* lam = bf.lambda + size_t(b) * nh;
  for (int e = tid; e < n; e += nt) {
    const int s = e / k, c = e % k;
    double acc = 0.0;
    if (s >= 1)  // rows (s-1, u): Jp[s-1]
      for (int u = 0; u < nu; ++u) acc += bf.Jp[((size_t(b) * T + (s - 1)) * nu + u) * k + c] * lam[(s - 1) * nu + u];
    if (s < T)  // rows (s, u): Jt[s]
      for (int u = 0; u < nu; ++u) acc += bf.Jt[((size_t(b) * T + s) * nu + u) * k + c] * lam[s * nu + u];
    if (s + 1 < T)  // rows (s+1, u): Jm[s+1]
      for (int u = 0; u < nu; ++u) acc += bf.Jm[((size_t(b) * T + (s + 1)) * nu + u) * k + c] * lam[(s + 1) * nu + u];
    gm[e] = gs[e] + acc;
  }
  double part = 0.0;
  const double* h = bf.st.h + size_t(b) * nh;
  for (int e = tid; e < nh; e += nt) part += h[e] * lam[e];
  part = block_sum(part, red);
  if (tid == 0) bf.merit[b] = bf.st.cost[b] + part;
}

static int solve_cpb(const SolverConsts& sc, int ncols) {
  int cpb = 256 / sc.nq;
  if (cpb > ncols) cpb = ncols;
  if (sc.nu > 0 && cpb > sc.nu) cpb = (cpb / sc.nu) * sc.nu;  // keep the columns of one time step together
  return cpb < 1 ? 1 : cpb;
}

void launch_penta_solve(const SolverConsts& sc, const SolverBufs& bf, double* X, int ncols, int mode, bool force,
                        cudaStream_t stream) {
  const int cpb = solve_cpb(sc, ncols);
  const int ctas = (ncols + cpb - 1) / cpb;
  const int smem = (3 * sc.nq * sc.nq + 3 * cpb * sc.nq) * 8;
  g_launch_counter += 1;
  k_penta_solve<<<sc.B * ctas, 256, smem, stream>>>(sc, bf, X, ncols, cpb, mode, force ? 1 : 0);
}

void launch_factor(const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  const int smem = 12 * sc.nq * sc.nq * 8;
  g_launch_counter += 1;
  k_factor<<<sc.B, 256, smem, stream>>>(sc, bf, force ? 1 : 0);
}

void launch_lagrange(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                     cudaStream_t stream) {
  (void)dm;
  if (sc.eq && sc.nh > 0) {
    launch_penta_solve(sc, bf, bf.X, sc.nh, 1, force, stream);
    g_launch_counter += 2;
    k_schur<<<sc.B * sc.nh, 128, sc.n * 8, stream>>>(sc, bf, force ? 1 : 0);
    k_ldlt<<<sc.B, 512, (2 * sc.nh + sc.nh / 2 + 2) * 8, stream>>>(sc, bf, force ? 1 : 0);
  }
  g_launch_counter += 1;
  k_merit<<<sc.B, 256, 0, stream>>>(sc, bf, force ? 1 : 0);
}

}  // namespace idto
