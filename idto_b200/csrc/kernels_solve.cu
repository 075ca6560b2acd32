// Block penta-diagonal elimination: Newton step and Lagrange multipliers in one banded pass.
//
// Reference: PentaDiagonalFactorization::Factorize / SolveInPlace (optimizer/penta_diagonal_solver.h:
// 124-248, block Thomas after Benkert & Fischer 2007), CalcLagrangeMultipliers
// (optimizer/trajectory_optimizer.cc:1371-1396) and the Gauss-Newton step of CalcDoglegPoint
// (cc:2137-2140).  The reference forms H~^-1 J~^T column by column (n_unact*T solves), a dense
// Schur complement S = J~ H~^-1 J~^T, an LDLT of S, and then a second factorisation of H~ for pH.
// Those are the KKT conditions (cc:2117-2123)
//        [ H~  J~^T ] [ x ]   [ -g~ ]        lambda = S^-1 (h - J~ H~^-1 g~),
//        [ J~   0   ] [ l ] = [ -h  ]        x = -H~^-1 (g~ + J~^T lambda) = Delta * pH .
// Ordering the unknowns time-major as blocks (q_i, lambda_{i-1}) keeps the KKT matrix block
// penta-diagonal (h_{i-1} depends on q_{i-2}, q_{i-1}, q_i) with block size nq + n_unact and puts the
// dominant d tau_{i-1}/d q_i block on the diagonal, so the same block-Thomas recurrence solves for x and
// lambda in ONE sweep over T+1 blocks.  With equality constraints off the block size is nq and the
// system is H~ x = -g~.  Diagonal blocks are eliminated by Gauss-Jordan with partial pivoting applied
// to the augmented right-hand sides [D_i - K_i Z_{i-1} | E_i | r_i] (the reference uses PartialPivLU
// solves; no explicit inverse is ever formed — explicit inverses lose the cond(H) ~ 1e10 acrobot case).
// One CTA per problem; Y_i, Z_i, r_i go to HBM for the backward sweep.
#include "reduce.cuh"
#include "solver.h"

namespace idto {

namespace {

struct KktView {
  const double *SA, *SB, *SC;  // scaled Hessian lower bands of problem b: [T+1][nq*nq] column-major
  const double *Jm, *Jt, *Jp;  // scaled Jacobian bands of problem b: [T][nu*nq] row-major (u, c)
  int nq, nu, T, eq;
};

// Entry (r, c) of the lower-band blocks of the time-major KKT matrix (see file header).
__device__ __forceinline__ double kkt_C(const KktView& V, int i, int r, int c) {
  const int nq = V.nq;
  if (r < nq && c < nq) return V.SC[size_t(i) * nq * nq + c * nq + r];
  if (r >= nq && c >= nq) return (i == 0 && r == c) ? 1.0 : 0.0;  // dummy lambda_{-1}
  if (i == 0) return 0.0;
  const int u = (r >= nq ? r : c) - nq, cc = (r >= nq ? c : r);
  return V.Jp[(size_t(i - 1) * V.nu + u) * nq + cc];
}
__device__ __forceinline__ double kkt_B(const KktView& V, int i, int r, int c) {  // block (i, i-1), i >= 1
  const int nq = V.nq;
  if (c >= nq) return 0.0;
  if (r < nq) return V.SB[size_t(i) * nq * nq + c * nq + r];
  return V.Jt[(size_t(i - 1) * V.nu + (r - nq)) * nq + c];
}
__device__ __forceinline__ double kkt_A(const KktView& V, int i, int r, int c) {  // block (i, i-2), i >= 2
  const int nq = V.nq;
  if (c >= nq) return 0.0;
  if (r < nq) return V.SA[size_t(i) * nq * nq + c * nq + r];
  return V.Jm[(size_t(i - 1) * V.nu + (r - nq)) * nq + c];
}

}  // namespace

__global__ void __launch_bounds__(256) k_kkt_solve(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_piv;
  __shared__ int s_fail;
  const int b = blockIdx.x;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int nblk = sc.T + 1, nq = sc.nq, nu = sc.eq ? sc.nu : 0, kb = nq + nu, kk = kb * kb;
  const int W = 3 * kb + 1, tid = threadIdx.x, nt = blockDim.x;
  KktView V;
  V.SA = bf.SA + size_t(b) * nblk * nq * nq, V.SB = bf.SB + size_t(b) * nblk * nq * nq;
  V.SC = bf.SC + size_t(b) * nblk * nq * nq;
  V.Jm = bf.Jm + size_t(b) * sc.T * sc.nu * nq, V.Jt = bf.Jt + size_t(b) * sc.T * sc.nu * nq;
  V.Jp = bf.Jp + size_t(b) * sc.T * sc.nu * nq;
  V.nq = nq, V.nu = sc.nu, V.T = sc.T, V.eq = sc.eq;
  const double* gs = bf.gs + size_t(b) * sc.n;
  const double* h = bf.st.h + size_t(b) * sc.nh;

  double* Ym2 = sm;
  double* Ym1 = Ym2 + kk;
  double* Zm2 = Ym1 + kk;
  double* Zm1 = Zm2 + kk;
  double* sA = Zm1 + kk;
  double* sK = sA + kk;
  double* M = sK + kk;        // kb x W, column-major: [G | Yrhs | E | r]
  double* rm1 = M + kb * W;   // r_{i-1}
  double* rm2 = rm1 + kb;     // r_{i-2}
  double* mult = rm2 + kb;    // multipliers of the current elimination step
  double* FY = bf.FY + size_t(b) * nblk * kk;
  double* FZ = bf.FZ + size_t(b) * nblk * kk;
  double* Fr = bf.X + size_t(b) * nblk * kb;  // r_i  ([B][(T+1)*(nq+nu)] doubles)
  for (int e = tid; e < 4 * kk; e += nt) sm[e] = 0.0;
  for (int e = tid; e < 2 * kb; e += nt) rm1[e] = 0.0;
  if (tid == 0) s_fail = 0;
  __syncthreads();

  for (int i = 0; i < nblk; ++i) {
    // ---- load A_i, B_i, and the augmented matrix [C_i | D_i | E_i | b_i] ----------------------
    for (int e = tid; e < kk; e += nt) {
      const int c = e / kb, r = e % kb;
      sA[e] = i >= 2 ? kkt_A(V, i, r, c) : 0.0;
      sK[e] = i >= 1 ? kkt_B(V, i, r, c) : 0.0;
      M[e] = kkt_C(V, i, r, c);
      M[kk + e] = i < nblk - 1 ? kkt_B(V, i + 1, c, r) : 0.0;      // D_i = B_{i+1}^T
      M[2 * kk + e] = i < nblk - 2 ? kkt_A(V, i + 2, c, r) : 0.0;  // E_i = A_{i+2}^T
    }
    for (int r = tid; r < kb; r += nt)
      M[3 * kk + r] = r < nq ? -gs[i * nq + r] : (i >= 1 ? -h[(i - 1) * sc.nu + (r - nq)] : 0.0);
    __syncthreads();
    // ---- K_i = B_i - A_i Y_{i-2};  G_i = C_i - A_i Z_{i-2};  r -= A_i r_{i-2} ------------------
    for (int e = tid; e < kk; e += nt) {
      const int c = e / kb, r = e % kb;
      double k1 = 0.0, g1 = 0.0;
      for (int j = 0; j < kb; ++j) {
        const double a = sA[j * kb + r];
        k1 += a * Ym2[c * kb + j];
        g1 += a * Zm2[c * kb + j];
      }
      sK[e] -= k1;
      M[e] -= g1;
    }
    for (int r = tid; r < kb; r += nt) {
      double acc = 0.0;
      for (int j = 0; j < kb; ++j) acc += sA[j * kb + r] * rm2[j];
      M[3 * kk + r] -= acc;
    }
    __syncthreads();
    // ---- G_i -= K_i Y_{i-1};  Yrhs = D_i - K_i Z_{i-1};  r -= K_i r_{i-1} ----------------------
    for (int e = tid; e < kk; e += nt) {
      const int c = e / kb, r = e % kb;
      double g1 = 0.0, y1 = 0.0;
      for (int j = 0; j < kb; ++j) {
        const double kx = sK[j * kb + r];
        g1 += kx * Ym1[c * kb + j];
        y1 += kx * Zm1[c * kb + j];
      }
      M[e] -= g1;
      M[kk + e] -= y1;
    }
    for (int r = tid; r < kb; r += nt) {
      double acc = 0.0;
      for (int j = 0; j < kb; ++j) acc += sK[j * kb + r] * rm1[j];
      M[3 * kk + r] -= acc;
    }
    __syncthreads();
    // ---- Gauss-Jordan with partial pivoting on [G | Yrhs | E | r] -------------------------------
    for (int c = 0; c < kb; ++c) {
      if (tid < 32) {
        double best = -1.0;
        int bi = c;
        for (int r = c + tid; r < kb; r += 32) {
          const double x = fabs(M[c * kb + r]);
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
      // swap rows c <-> p (columns >= c) and capture the multipliers of the swapped column
      for (int j = c + tid; j < W; j += nt) {
        const double x = M[j * kb + c], y = M[j * kb + p];
        M[j * kb + c] = y;
        M[j * kb + p] = x;
      }
      __syncthreads();
      const double piv = M[c * kb + c];
      for (int r = tid; r < kb; r += nt) mult[r] = (r == c) ? 0.0 : M[c * kb + r] / piv;
      __syncthreads();
      // row_r -= mult_r * row_c for r != c, then scale row c; only columns > c matter from here on
      const int ncols = W - c - 1;
      for (int e = tid; e < ncols * kb; e += nt) {
        const int j = c + 1 + e / kb, r = e % kb;
        const double pc = M[j * kb + c];
        if (r != c)
          M[j * kb + r] -= mult[r] * pc;
      }
      __syncthreads();
      for (int j = c + 1 + tid; j < W; j += nt) M[j * kb + c] /= piv;
      // (column c itself is implicitly e_c from now on; it is never read again)
      __syncthreads();
    }
    // ---- M = [I | Y_i | Z_i | r_i]: store and shift -------------------------------------------------
    for (int e = tid; e < kk; e += nt) {
      const double y = M[kk + e], z = M[2 * kk + e];
      FY[size_t(i) * kk + e] = y;
      FZ[size_t(i) * kk + e] = z;
      Ym2[e] = Ym1[e], Zm2[e] = Zm1[e];
    }
    for (int r = tid; r < kb; r += nt) {
      Fr[size_t(i) * kb + r] = M[3 * kk + r];
      rm2[r] = rm1[r];
    }
    __syncthreads();
    for (int e = tid; e < kk; e += nt) Ym1[e] = M[kk + e], Zm1[e] = M[2 * kk + e];
    for (int r = tid; r < kb; r += nt) rm1[r] = M[3 * kk + r];
    __syncthreads();
  }
  // ---- backward sweep: x_i = r_i - Y_i x_{i+1} - Z_i x_{i+2} (penta_diagonal_solver.h:229-247) -----
  double* x1 = rm1;  // x_{i+1}
  double* x2 = rm2;  // x_{i+2}
  for (int e = tid; e < 2 * kb; e += nt) rm1[e] = 0.0;
  __syncthreads();
  double* xq = bf.pH + size_t(b) * sc.n;          // x = -H~^-1 gm  (Delta * pH)
  double* lam = bf.lambda + size_t(b) * sc.nh;
  for (int i = nblk - 1; i >= 0; --i) {
    for (int e = tid; e < kk; e += nt) sA[e] = FY[size_t(i) * kk + e], sK[e] = FZ[size_t(i) * kk + e];
    __syncthreads();
    double out = 0.0;
    if (tid < kb) {
      out = Fr[size_t(i) * kb + tid];
      for (int j = 0; j < kb; ++j) out -= sA[j * kb + tid] * x1[j];
      for (int j = 0; j < kb; ++j) out -= sK[j * kb + tid] * x2[j];
      if (tid < nq)
        xq[i * nq + tid] = out;
      else if (i >= 1)
        lam[(i - 1) * sc.nu + (tid - nq)] = out;
    }
    __syncthreads();
    if (tid < kb) x2[tid] = x1[tid];
    __syncthreads();
    if (tid < kb) x1[tid] = out;
    __syncthreads();
  }
  if (tid == 0 && s_fail) atomicExch(bf.status, IDTO_ERR_FACTORIZATION);
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

void launch_factor(const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  (void)sc, (void)bf, (void)force, (void)stream;  // fused into launch_lagrange (k_kkt_solve)
}

void launch_lagrange(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                     cudaStream_t stream) {
  (void)dm;
  const int kb = sc.nq + (sc.eq ? sc.nu : 0);
  const int smem = (6 * kb * kb + kb * (3 * kb + 1) + 3 * kb) * 8;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_kkt_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  g_launch_counter += 2;
  k_kkt_solve<<<sc.B, 256, smem, stream>>>(sc, bf, force ? 1 : 0);
  k_merit<<<sc.B, 256, 0, stream>>>(sc, bf, force ? 1 : 0);
}

}  // namespace idto
