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

constexpr int kMaxOwnCols = 13;  // ceil((3*32+1)/8) columns of [G|Y|Z|r] per warp

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

// Thread mapping: 8 warps; lane = block row r, warp w owns the columns j = w, w+8, ... of every
// matrix, so column-major shared-memory accesses are conflict-free and a Gauss-Jordan step needs a
// single CTA barrier (every warp recomputes the pivot search redundantly from the finished column).
template <int KB>
__global__ void __launch_bounds__(256) k_kkt_solve(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_fail;
  const int b = blockIdx.x;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  constexpr int kb = KB, kk = KB * KB, W = 3 * KB + 1;
  constexpr int NCK = (KB + 7) / 8;  // own columns of a KB-wide matrix
  constexpr int NCW = (W + 7) / 8;   // own columns of the augmented matrix
  const int nblk = sc.T + 1, nq = sc.nq;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool row = lane < kb;
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
  double* FY = bf.FY + size_t(b) * nblk * kk;
  double* FZ = bf.FZ + size_t(b) * nblk * kk;
  double* Fr = bf.X + size_t(b) * nblk * kb;  // r_i  ([B][(T+1)*(nq+nu)] doubles)
  for (int e = tid; e < 4 * kk; e += blockDim.x) sm[e] = 0.0;
  for (int e = tid; e < 2 * kb; e += blockDim.x) rm1[e] = 0.0;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  const int r = row ? lane : 0;

  for (int i = 0; i < nblk; ++i) {
    // ---- load A_i, B_i and the augmented matrix [C_i | D_i | E_i | b_i] ------------------------------
    {
      double va[NCK], vb[NCK], vc[NCK], vd[NCK], ve[NCK];
#pragma unroll
      for (int mm = 0; mm < NCK; ++mm) {
        const int c = wid + 8 * mm;
        if (c < kb && row) {
          va[mm] = i >= 2 ? kkt_A(V, i, r, c) : 0.0;
          vb[mm] = i >= 1 ? kkt_B(V, i, r, c) : 0.0;
          vc[mm] = kkt_C(V, i, r, c);
          vd[mm] = i < nblk - 1 ? kkt_B(V, i + 1, c, r) : 0.0;  // D_i = B_{i+1}^T
          ve[mm] = i < nblk - 2 ? kkt_A(V, i + 2, c, r) : 0.0;  // E_i = A_{i+2}^T
        }
      }
#pragma unroll
      for (int mm = 0; mm < NCK; ++mm) {
        const int c = wid + 8 * mm;
        if (c < kb && row) {
          const int e = c * kb + r;
          sA[e] = va[mm], sK[e] = vb[mm], M[e] = vc[mm], M[kk + e] = vd[mm], M[2 * kk + e] = ve[mm];
        }
      }
      if (wid == 0 && row)
        M[3 * kk + r] = r < nq ? -gs[i * nq + r] : (i >= 1 ? -h[(i - 1) * sc.nu + (r - nq)] : 0.0);
    }
    __syncthreads();
    // ---- K_i = B_i - A_i Y_{i-2};  G_i = C_i - A_i Z_{i-2};  r -= A_i r_{i-2}  (A_i has nq nonzero columns)
    if (row) {
#pragma unroll
      for (int mm = 0; mm < NCK; ++mm) {
        const int c = wid + 8 * mm;
        if (c < kb) {
          double k1 = 0.0, g1 = 0.0;
#pragma unroll 5
          for (int j = 0; j < nq; ++j) {
            const double a = sA[j * kb + r];
            k1 += a * Ym2[c * kb + j];
            g1 += a * Zm2[c * kb + j];
          }
          sK[c * kb + r] -= k1;
          M[c * kb + r] -= g1;
        }
      }
      if (wid == 0) {
        double acc = 0.0;
        for (int j = 0; j < nq; ++j) acc += sA[j * kb + r] * rm2[j];
        M[3 * kk + r] -= acc;
      }
    }
    __syncthreads();
    // ---- G_i -= K_i Y_{i-1};  Yrhs = D_i - K_i Z_{i-1};  r -= K_i r_{i-1} ----------------------------
    if (row) {
#pragma unroll
      for (int mm = 0; mm < NCK; ++mm) {
        const int c = wid + 8 * mm;
        if (c < kb) {
          double g1 = 0.0, y1 = 0.0;
#pragma unroll
          for (int j = 0; j < kb; ++j) {
            const double kx = sK[j * kb + r];
            g1 += kx * Ym1[c * kb + j];
            y1 += kx * Zm1[c * kb + j];
          }
          M[c * kb + r] -= g1;
          M[kk + c * kb + r] -= y1;
        }
      }
      if (wid == 0) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < kb; ++j) acc += sK[j * kb + r] * rm1[j];
        M[3 * kk + r] -= acc;
      }
    }
    __syncthreads();
    // ---- Gauss-Jordan with partial pivoting on [G | Yrhs | E | r]: one barrier per step -------------
#pragma unroll 1
    for (int c = 0; c < kb; ++c) {
      // pivot = first row >= c with the largest |M[r, c]|: warp max over the two 32-bit halves of the
      // (non-negative) double, then the lowest matching lane
      const double cand = (row && lane >= c) ? fabs(M[c * kb + lane]) : -1.0;
      const unsigned hi = cand >= 0.0 ? unsigned(__double2hiint(cand)) + 1u : 0u;
      const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
      const unsigned lo = (hi == mhi && cand >= 0.0) ? unsigned(__double2loint(cand)) : 0u;
      const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
      const unsigned hit = __ballot_sync(0xffffffffu, hi == mhi && lo == mlo && cand >= 0.0);
      int p = hit ? __ffs(hit) - 1 : c;
      const double best = __shfl_sync(0xffffffffu, cand, p);
      if (!(best > 0.0)) {
        if (tid == 0) s_fail = 1;
        p = c;
      }
      // one reciprocal per step (not one division per element): row c is scaled by 1/piv, the
      // multipliers are M[src, c] / piv
      const double inv = 1.0 / (best > 0.0 ? M[c * kb + p] : 1.0);
      const int src = (lane == c) ? p : (lane == p ? c : lane);  // row swap c <-> p, applied on the fly
      const int srcr = row ? src : 0;
      const double m = M[c * kb + srcr] * inv;
      // own columns j = wid + 8*mm strictly right of c: gather, one warp barrier, scatter
      double xs[NCW], pc[NCW];
#pragma unroll
      for (int mm = 0; mm < NCW; ++mm) {
        const int j = wid + 8 * mm;
        if (j > c && j < W) {
          xs[mm] = M[j * kb + srcr];
          pc[mm] = M[j * kb + p];
        }
      }
      __syncwarp();
      if (row) {
#pragma unroll
        for (int mm = 0; mm < NCW; ++mm) {
          const int j = wid + 8 * mm;
          if (j > c && j < W) M[j * kb + lane] = (lane == c) ? pc[mm] * inv : xs[mm] - m * pc[mm];
        }
      }
      __syncthreads();
    }
    // ---- M = [I | Y_i | Z_i | r_i]: store and shift ----------------------------------------------------
    if (row) {
#pragma unroll
      for (int mm = 0; mm < NCK; ++mm) {
        const int c = wid + 8 * mm;
        if (c < kb) {
          const int e = c * kb + r;
          const double y = M[kk + e], z = M[2 * kk + e];
          FY[size_t(i) * kk + e] = y;
          FZ[size_t(i) * kk + e] = z;
          Ym2[e] = Ym1[e], Zm2[e] = Zm1[e];
          Ym1[e] = y, Zm1[e] = z;  // (c, r) is owned by this thread in all four arrays
        }
      }
      if (wid == 0) {
        const double rr = M[3 * kk + r];
        Fr[size_t(i) * kb + r] = rr;
        rm2[r] = rm1[r];
        rm1[r] = rr;
      }
    }
    __syncthreads();
  }
  if (tid == 0 && s_fail) atomicExch(bf.status, IDTO_ERR_FACTORIZATION);
  // ---- backward sweep: x_i = r_i - Y_i x_{i+1} - Z_i x_{i+2} (penta_diagonal_solver.h:229-247) -------
  // All warps stage Y_i, Z_i from L2 into shared memory (coalesced, many loads in flight); warp 0
  // owns x (lane r holds x[r]) and broadcasts it with shuffles.
  double* xq = bf.pH + size_t(b) * sc.n;  // x = -H~^-1 gm  (Delta * pH)
  double* lam = bf.lambda + size_t(b) * sc.nh;
  double x1 = 0.0, x2 = 0.0;  // x_{i+1}[lane], x_{i+2}[lane]
  for (int i = nblk - 1; i >= 0; --i) {
    for (int e = tid; e < kk; e += blockDim.x) sA[e] = FY[size_t(i) * kk + e], sK[e] = FZ[size_t(i) * kk + e];
    __syncthreads();
    if (wid == 0) {
      double out = row ? Fr[size_t(i) * kb + lane] : 0.0;
      double accy = 0.0, accz = 0.0;
#pragma unroll
      for (int j = 0; j < kb; ++j) {
        const double xj1 = __shfl_sync(0xffffffffu, x1, j), xj2 = __shfl_sync(0xffffffffu, x2, j);
        accy += sA[j * kb + r] * xj1, accz += sK[j * kb + r] * xj2;
      }
      out = (out - accy) - accz;
      if (row) {
        if (lane < nq)
          xq[i * nq + lane] = out;
        else if (i >= 1)
          lam[(i - 1) * sc.nu + (lane - nq)] = out;
      }
      x2 = x1, x1 = row ? out : 0.0;
    }
    __syncthreads();
  }
}

template <int KB>
static void launch_kkt_kb(const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  const int smem = (6 * KB * KB + KB * (3 * KB + 1) + 2 * KB) * 8;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_kkt_solve<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  k_kkt_solve<KB><<<sc.B, 256, smem, stream>>>(sc, bf, force ? 1 : 0);
}

template <int KB>
static void launch_kkt_dispatch(int kb, const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  if (kb == KB) {
    launch_kkt_kb<KB>(sc, bf, force, stream);
  } else if constexpr (KB < 32) {
    launch_kkt_dispatch<KB + 1>(kb, sc, bf, force, stream);
  }
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
  g_launch_counter += 2;
  launch_kkt_dispatch<1>(kb, sc, bf, force, stream);
  k_merit<<<sc.B, 256, 0, stream>>>(sc, bf, force ? 1 : 0);
}

}  // namespace idto
