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
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>

#include "kkt_view.cuh"
#include "reduce.cuh"
#include "solver.h"

// Phase timing of the twisted sweep (debug builds only: make EXTRA=-DIDTO_KKT_TIMING).
#ifdef IDTO_KKT_TIMING
#define KT_DECL long long kt_t = clock64(), kt_acc[12] = {0};
#define KT(i) { const long long kt_n = clock64(); kt_acc[i] += kt_n - kt_t; kt_t = kt_n; }
#define KT_PRINT(b, dir)                                                                                    \
  if (threadIdx.x == 0 && (b) == 0)                                                                         \
    printf("kkt dir %d: load %lld gemm2 %lld gemm1 %lld lu %lld backsub %lld store %lld csync %lld iface_asm " \
           "%lld iface_gj %lld csync2 %lld final_backsub %lld\n", dir, kt_acc[0], kt_acc[1], kt_acc[2],      \
           kt_acc[3], kt_acc[4], kt_acc[5], kt_acc[6], kt_acc[7], kt_acc[8], kt_acc[9], kt_acc[10]);
#else
#define KT_DECL
#define KT(i)
#define KT_PRINT(b, dir)
#endif

namespace idto {

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
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_kkt_solve<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  k_kkt_solve<KB><<<sc.B, 256, smem, stream>>>(sc, bf, force ? 1 : 0);
}

template <int KB>
static bool launch_kkt_dispatch(int kb, const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  if (kb == KB) {
    launch_kkt_kb<KB>(sc, bf, force, stream);
    return true;
  } else if constexpr (KB < 32) {
    return launch_kkt_dispatch<KB + 1>(kb, sc, bf, force, stream);
  }
  return false;
}

// ---- two-sided ("twisted") block elimination: a cluster of two CTAs per problem ---------------------
// CTA 0 eliminates block rows 0..m top-down, CTA 1 eliminates rows N..m+1 bottom-up (the matrix is
// symmetric, so the reversed sweep is the same recurrence with Blk(i, i+1) = B_{i+1}^T and
// Blk(i, i+2) = A_{i+2}^T as its "back" blocks).  The two chains meet in a 2kb x 2kb interface system for
// (x_m, x_{m+1}); then both CTAs back-substitute their halves concurrently.  This halves the sequential
// chain of (T+1)*kb pivot steps that bounds the single-sweep kernel.
template <int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256)
    k_kkt_twisted(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_fail;
  __shared__ int s_piv;
  const int b = blockIdx.x >> 1, dir = blockIdx.x & 1;
  if (!force && !bf.ctl[b].derivs_dirty) return;  // same decision in both CTAs of the cluster
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  constexpr int kb = KB, kk = KB * KB, W = 3 * KB + 1;
  constexpr int NCK = (KB + 7) / 8, NCW = (W + 7) / 8;
  const int nblk = sc.T + 1, N = sc.T, nq = sc.nq;
  const int mid = N / 2;                                   // forward: rows 0..mid, backward: rows N..mid+1
  const int nsteps = dir == 0 ? mid + 1 : N - mid;
  const int sgn = dir == 0 ? 1 : -1, first = dir == 0 ? 0 : N;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool row = lane < kb;
  const int r = row ? lane : 0;
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
  double* M = sK + kk;       // kb x W: [G | Yrhs | E | r]; later reused for the interface system
  double* rm1 = M + kb * W;
  double* rm2 = rm1 + kb;
  double* FY = bf.FY + size_t(b) * nblk * kk;
  double* FZ = bf.FZ + size_t(b) * nblk * kk;
  double* Fr = bf.X + size_t(b) * nblk * kb;
  double* xq = bf.pH + size_t(b) * sc.n;
  double* lam = bf.lambda + size_t(b) * sc.nh;
  double* xint = bf.tmp2 + size_t(b) * sc.n;  // interface solution (x_mid, x_mid+1): 2*kb doubles (n >= 2*kb)
  for (int e = tid; e < 4 * kk; e += blockDim.x) sm[e] = 0.0;
  for (int e = tid; e < 2 * kb; e += blockDim.x) rm1[e] = 0.0;
  if (tid == 0) s_fail = 0;
  __syncthreads();
  KT_DECL
  // inner dimension of the products with the "back-2"/"back-1" blocks: the lower-band blocks A_i, B_i have
  // only nq non-zero columns; their transposes (reversed sweep) do not
  const int jn = dir == 0 ? nq : kb;

  for (int n = 0; n < nsteps; ++n) {
    const int i = first + sgn * n;
    const bool hb1 = n >= 1, hb2 = n >= 2;                          // back neighbours exist
    const bool hf1 = dir == 0 ? (i + 1 <= N) : (i - 1 >= 0);         // front neighbours exist in the matrix
    const bool hf2 = dir == 0 ? (i + 2 <= N) : (i - 2 >= 0);
    {
      double va[NCK], vb[NCK], vc[NCK], vd[NCK], ve[NCK];
#pragma unroll
      for (int mm = 0; mm < NCK; ++mm) {
        const int c = wid + 8 * mm;
        if (c < kb && row) {
          va[mm] = hb2 ? kkt_blk(V, i, i - 2 * sgn, r, c) : 0.0;
          vb[mm] = hb1 ? kkt_blk(V, i, i - sgn, r, c) : 0.0;
          vc[mm] = kkt_C(V, i, r, c);
          vd[mm] = hf1 ? kkt_blk(V, i, i + sgn, r, c) : 0.0;
          ve[mm] = hf2 ? kkt_blk(V, i, i + 2 * sgn, r, c) : 0.0;
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
    KT(0)
    if (row) {
#pragma unroll
      for (int mm = 0; mm < NCK; ++mm) {
        const int c = wid + 8 * mm;
        if (c < kb) {
          double k1 = 0.0, g1 = 0.0;
#pragma unroll 5
          for (int j = 0; j < jn; ++j) {
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
        for (int j = 0; j < jn; ++j) acc += sA[j * kb + r] * rm2[j];
        M[3 * kk + r] -= acc;
      }
    }
    __syncthreads();
    KT(1)
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
    KT(2)
    // LU with partial pivoting (unit-diagonal U: the pivot row is normalised), applied to the
    // right-hand sides [Yrhs | E | r] as well; then a barrier-free, column-parallel back-substitution.
    // (Gauss-Jordan is not backward stable for the bottom-up chain, whose diagonal blocks reach
    // cond ~ 1e10; LU with partial pivoting is.)
#pragma unroll 1
    for (int c = 0; c < kb; ++c) {
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
      const double inv = 1.0 / (best > 0.0 ? M[c * kb + p] : 1.0);
      const int src = (lane == c) ? p : (lane == p ? c : lane);  // row swap c <-> p, applied on the fly
      const bool low = row && lane >= c;                         // rows above the pivot keep their U entries
      const int srcr = low ? src : 0;
      const double m = M[c * kb + srcr];                         // multiplier w.r.t. the normalised pivot row
      double xs[NCW], pc[NCW];
#pragma unroll
      for (int mm = 0; mm < NCW; ++mm) {
        const int j = wid + 8 * mm;
        if (j > c && j < W) {
          xs[mm] = M[j * kb + srcr];
          pc[mm] = M[j * kb + p] * inv;
        }
      }
      __syncwarp();
      if (low) {
#pragma unroll
        for (int mm = 0; mm < NCW; ++mm) {
          const int j = wid + 8 * mm;
          if (j > c && j < W) M[j * kb + lane] = (lane == c) ? pc[mm] : xs[mm] - m * pc[mm];
        }
      }
      __syncthreads();
    }
    KT(3)
    // back-substitution with the unit upper-triangular U = M[:, 0:kb]: one thread per right-hand side
    if (tid < W - kb) {
      double* col = M + (kb + tid) * kb;
      double x[KB];
#pragma unroll
      for (int rr = KB - 1; rr >= 0; --rr) {
        double acc = col[rr];
#pragma unroll
        for (int k2 = KB - 1; k2 > rr; --k2) acc -= M[k2 * kb + rr] * x[k2];
        x[rr] = acc;
      }
#pragma unroll
      for (int rr = 0; rr < KB; ++rr) col[rr] = x[rr];
    }
    __syncthreads();
    KT(4)
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
          Ym1[e] = y, Zm1[e] = z;
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
    KT(5)
  }
  if (tid == 0 && s_fail) atomicExch(bf.status, IDTO_ERR_FACTORIZATION);
  __threadfence();
  cluster.sync();  // both chains (and their Y, Z, r in HBM) are complete
  KT(6)

  // ---- interface system for u = (x_mid, x_mid+1), solved by CTA 0 ----------------------------------------
  //   (I - Z_m Z'_{m+2}) x_m + (Y_m - Z_m Y'_{m+2}) x_{m+1} = r_m - Z_m r'_{m+2}
  //   (Y'_{m+1} - Z'_{m+1} Y_{m-1}) x_m + (I - Z'_{m+1} Z_{m-1}) x_{m+1} = r'_{m+1} - Z'_{m+1} r_{m-1}
  // (primes: the bottom-up chain; a missing neighbour contributes zero blocks)
  constexpr int n2 = 2 * KB, W2 = 2 * KB + 1;
  if (dir == 0) {
    double* Q = M;  // n2 x W2 column-major (fits: 2kb*(2kb+1) <= kb*(3kb+1) + slack from sA/sK? use sm base)
    Q = sm;         // reuse the whole dynamic shared memory (6kk + kb*W doubles >= n2*W2)
    __syncthreads();
    const int m0 = mid;
    auto blkY = [&](int i) { return FY + size_t(i) * kk; };
    auto blkZ = [&](int i) { return FZ + size_t(i) * kk; };
    const bool has_m2 = m0 + 2 <= N, has_mm1 = m0 - 1 >= 0;
    for (int e = tid; e < n2 * W2; e += blockDim.x) {
      const int c = e / n2, rr = e % n2;
      double val = 0.0;
      if (c < n2) {
        const int bc = c / kb, cc = c % kb, br = rr / kb, r1 = rr % kb;
        if (br == 0 && bc == 0) {          // I - Z_m Z'_{m+2}
          val = (r1 == cc) ? 1.0 : 0.0;
          if (has_m2)
            for (int j = 0; j < kb; ++j) val -= blkZ(m0)[j * kb + r1] * blkZ(m0 + 2)[cc * kb + j];
        } else if (br == 0 && bc == 1) {   // Y_m - Z_m Y'_{m+2}
          val = blkY(m0)[cc * kb + r1];
          if (has_m2)
            for (int j = 0; j < kb; ++j) val -= blkZ(m0)[j * kb + r1] * blkY(m0 + 2)[cc * kb + j];
        } else if (br == 1 && bc == 0) {   // Y'_{m+1} - Z'_{m+1} Y_{m-1}
          val = blkY(m0 + 1)[cc * kb + r1];
          if (has_mm1)
            for (int j = 0; j < kb; ++j) val -= blkZ(m0 + 1)[j * kb + r1] * blkY(m0 - 1)[cc * kb + j];
        } else {                           // I - Z'_{m+1} Z_{m-1}
          val = (r1 == cc) ? 1.0 : 0.0;
          if (has_mm1)
            for (int j = 0; j < kb; ++j) val -= blkZ(m0 + 1)[j * kb + r1] * blkZ(m0 - 1)[cc * kb + j];
        }
      } else {
        const int br = rr / kb, r1 = rr % kb;
        if (br == 0) {
          val = Fr[size_t(m0) * kb + r1];
          if (has_m2)
            for (int j = 0; j < kb; ++j) val -= blkZ(m0)[j * kb + r1] * Fr[size_t(m0 + 2) * kb + j];
        } else {
          val = Fr[size_t(m0 + 1) * kb + r1];
          if (has_mm1)
            for (int j = 0; j < kb; ++j) val -= blkZ(m0 + 1)[j * kb + r1] * Fr[size_t(m0 - 1) * kb + j];
        }
      }
      Q[e] = val;
    }
    __syncthreads();
    KT(7)
    // Gauss-Jordan with partial pivoting on the n2 x (n2+1) system; warp 0 searches (two rows per lane)
    for (int c = 0; c < n2; ++c) {
      if (tid < 32) {
        double best = -1.0;
        int bi = c;
        for (int rr = c + lane; rr < n2; rr += 32) {
          const double x = fabs(Q[c * n2 + rr]);
          if (x > best) best = x, bi = rr;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
        }
        if (tid == 0) {
          s_piv = bi;
          if (!(best > 0.0)) atomicExch(bf.status, IDTO_ERR_FACTORIZATION);
        }
      }
      __syncthreads();
      const int p = s_piv;
      if (p != c)
        for (int j = c + tid; j < W2; j += blockDim.x) {
          const double x = Q[j * n2 + c];
          Q[j * n2 + c] = Q[j * n2 + p];
          Q[j * n2 + p] = x;
        }
      __syncthreads();
      const double inv = 1.0 / Q[c * n2 + c];
      const int ncols = W2 - c - 1;
      for (int e = tid; e < ncols * n2; e += blockDim.x) {
        const int j = c + 1 + e / n2, rr = e % n2;
        if (rr != c) Q[j * n2 + rr] -= (Q[c * n2 + rr] * inv) * Q[j * n2 + c];
      }
      __syncthreads();
      for (int j = c + 1 + tid; j < W2; j += blockDim.x) Q[j * n2 + c] *= inv;
      __syncthreads();
    }
    for (int e = tid; e < n2; e += blockDim.x) xint[e] = Q[n2 * n2 + e];
    __threadfence();
    KT(8)
  }
  cluster.sync();  // the interface solution is visible to both CTAs
  KT(9)

  // ---- back-substitution of each half, warp 0 of each CTA ------------------------------------------------
  //   top half:    x_i = r_i - Y_i x_{i+1} - Z_i x_{i+2},  i = mid-1 .. 0
  //   bottom half: x_i = r_i - Y_i x_{i-1} - Z_i x_{i-2},  i = mid+2 .. N
  __syncthreads();
  double x1, x2;  // the two most recent solution blocks (lane r holds component r)
  if (dir == 0) x1 = row ? xint[lane] : 0.0, x2 = row ? xint[kb + lane] : 0.0;        // x_mid, x_mid+1
  else x1 = row ? xint[kb + lane] : 0.0, x2 = row ? xint[lane] : 0.0;                 // x_mid+1, x_mid
  auto emit = [&](int i, double out) {
    if (wid == 0 && row) {
      if (lane < nq)
        xq[i * nq + lane] = out;
      else if (i >= 1)
        lam[(i - 1) * sc.nu + (lane - nq)] = out;
    }
  };
  emit(dir == 0 ? mid : mid + 1, x1);
  const int istart = dir == 0 ? mid - 1 : mid + 2, iend = dir == 0 ? -1 : N + 1;
  for (int i = istart; i != iend; i -= sgn) {
    for (int e = tid; e < kk; e += blockDim.x) sA[e] = FY[size_t(i) * kk + e], sK[e] = FZ[size_t(i) * kk + e];
    __syncthreads();
    double out = 0.0;
    if (wid == 0) {
      out = row ? Fr[size_t(i) * kb + lane] : 0.0;
      double accy = 0.0, accz = 0.0;
#pragma unroll
      for (int j = 0; j < kb; ++j) {
        const double xj1 = __shfl_sync(0xffffffffu, x1, j), xj2 = __shfl_sync(0xffffffffu, x2, j);
        accy += sA[j * kb + r] * xj1, accz += sK[j * kb + r] * xj2;
      }
      out = (out - accy) - accz;
      emit(i, out);
      x2 = x1, x1 = row ? out : 0.0;
    }
    __syncthreads();
  }
  KT(10)
  KT_PRINT(b, dir)
}

template <int KB>
static void launch_kkt_twisted_kb(const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  const int smem = std::max((6 * KB * KB + KB * (3 * KB + 1) + 2 * KB), 2 * KB * (2 * KB + 1)) * 8;
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_kkt_twisted<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  k_kkt_twisted<KB><<<2 * sc.B, 256, smem, stream>>>(sc, bf, force ? 1 : 0);
}

template <int KB>
static bool launch_kkt_twisted_dispatch(int kb, const SolverConsts& sc, const SolverBufs& bf, bool force,
                                        cudaStream_t stream) {
  if (kb == KB) {
    launch_kkt_twisted_kb<KB>(sc, bf, force, stream);
    return true;
  } else if constexpr (KB < 32) {
    return launch_kkt_twisted_dispatch<KB + 1>(kb, sc, bf, force, stream);
  }
  return false;
}

// No KKT kernel exists for this block size (idto_solver_create rejects such models; this is the backstop):
// raise the sticky status so that the call fails instead of consuming a stale solution.
__global__ void k_kkt_unsupported(int* status) { atomicExch(status, IDTO_ERR_UNSUPPORTED); }


// ---- SolverParameters::LinearSolverType::kDenseLdlt (cc:2088-2093) ---------------------------------------------
// The reference's debugging alternative to the block penta-diagonal solver for the Gauss-Newton step:
// pH = H~.ldlt().solve(-gm / Delta).  Here: an LDL^T factorisation that knows nothing about the block structure
// (H~ stored as a plain band matrix of half bandwidth 3 nq - 1, one CTA per problem), run AFTER the KKT sweep, whose
// lambda it keeps; it overwrites x = -H~^-1 gm (= pH Delta).  A cross-check, not a fast path.
__global__ void __launch_bounds__(256) k_dense_ldlt(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double xs[];  // [n]
  __shared__ double s_d;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int nq = sc.nq, n = sc.n, bw = 3 * nq - 1, ld = bw + 1, kk = nq * nq, nblk = sc.T + 1;
  double* A = bf.S + size_t(b) * n * ld;  // A[i * ld + (i - j)] = H~(i, j), 0 <= i - j <= bw
  const double* SA = bf.SA + size_t(b) * nblk * kk;
  const double* SB = bf.SB + size_t(b) * nblk * kk;
  const double* SC = bf.SC + size_t(b) * nblk * kk;
  for (int e = tid; e < n * ld; e += nt) {
    const int i = e / ld, j = i - (e - i * ld);
    double val = 0.0;
    if (j >= 0) {
      const int bi = i / nq, ri = i - bi * nq, bj = j / nq, cj = j - bj * nq, d = bi - bj;
      const double* blk = d == 0 ? SC : (d == 1 ? SB : (d == 2 ? SA : nullptr));
      if (blk) val = blk[size_t(bi) * kk + cj * nq + ri];
    }
    A[e] = val;
  }
  for (int e = tid; e < n; e += nt) xs[e] = -bf.gm[size_t(b) * n + e];
  __syncthreads();
  const int ta = tid & 63, tb = tid >> 6;  // row offset, column-offset group (4 groups)
  for (int k = 0; k < n; ++k) {
    if (tid == 0) {
      s_d = A[size_t(k) * ld];
      if (!(s_d > 0.0)) atomicExch(bf.status, IDTO_ERR_FACTORIZATION);
    }
    __syncthreads();
    const double inv = 1.0 / s_d;
    const int m = min(bw, n - 1 - k);
    // trailing update inside the band: A(i, j) -= (A(i, k) / d) A(j, k),  k < j <= i <= k + m
    for (int a = ta + 1; a <= m; a += 64) {
      const double lik = A[size_t(k + a) * ld + a] * inv;
      for (int c = tb + 1; c <= a; c += 4) A[size_t(k + a) * ld + (a - c)] -= lik * A[size_t(k + c) * ld + c];
    }
    __syncthreads();
    for (int a = tid + 1; a <= m; a += nt) A[size_t(k + a) * ld + a] *= inv;  // L(k + a, k)
  }
  __syncthreads();
  // L y = b,  D z = y,  L^T x = z  (axpy forms)
  for (int k = 0; k < n; ++k) {
    const double xk = xs[k];
    const int m = min(bw, n - 1 - k);
    __syncthreads();
    for (int a = tid + 1; a <= m; a += nt) xs[k + a] -= A[size_t(k + a) * ld + a] * xk;
    __syncthreads();
  }
  for (int e = tid; e < n; e += nt) xs[e] /= A[size_t(e) * ld];
  __syncthreads();
  for (int k = n - 1; k >= 0; --k) {
    const double xk = xs[k];
    const int m = min(bw, k);
    __syncthreads();
    for (int a = tid + 1; a <= m; a += nt) xs[k - a] -= A[size_t(k) * ld + a] * xk;
    __syncthreads();
  }
  for (int e = tid; e < n; e += nt) bf.pH[size_t(b) * n + e] = xs[e];
}

void launch_factor(const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  (void)sc, (void)bf, (void)force, (void)stream;  // fused into launch_lagrange (k_kkt_solve)
}

void launch_lagrange(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                     cudaStream_t stream, bool with_gm) {
  (void)dm;
  const int kb = sc.nq + (sc.eq ? sc.nu : 0);
  g_launch_counter += 1;
  // two-sided elimination needs at least 4 block rows per half to pay off; 2*kb interface unknowns must fit n
  bool launched;
  if (sc.linear_solver == IDTO_LINSOLVE_CYCLIC_REDUCTION && bf.crw)
    launched = launch_kkt_cr(kb, sc, bf, force, stream);
  else if (sc.linear_solver != IDTO_LINSOLVE_THOMAS && sc.T + 1 >= 8)
    launched = launch_kkt_v3(kb, sc, bf, force, stream) || launch_kkt_tw2(kb, sc, bf, force, stream) ||
               launch_kkt_twisted_dispatch<1>(kb, sc, bf, force, stream);
  else
    launched = launch_kkt_dispatch<1>(kb, sc, bf, force, stream);
  if (!launched) k_kkt_unsupported<<<1, 1, 0, stream>>>(bf.status);
  const bool dense = sc.linear_solver == IDTO_LINSOLVE_DENSE_LDLT && bf.S;
  (void)with_gm;
  launch_gm_matvec(sc, bf, force, stream);  // gm, merit, gHg, g.g
  if (dense) {
    static bool attr_set[kMaxDevices] = {};
    if (first_use_on_device(attr_set))
      cudaFuncSetAttribute(k_dense_ldlt, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    g_launch_counter += 1;
    k_dense_ldlt<<<sc.B, 256, sc.n * sizeof(double), stream>>>(sc, bf, force ? 1 : 0);
  }
}

}  // namespace idto
