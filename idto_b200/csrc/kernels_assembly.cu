// Gradient (cc:1021-1081), Gauss-Newton Hessian bands (cc:1093-1165), scale factors (cc:1225-1255),
// scaled Hessian / gradient (cc:1181-1223, penta_diagonal_matrix.cc:221-257) and the scaled
// equality-constraint Jacobian bands (cc:1292-1334).  Diagonal cost weights (every reference example;
// dense weights are rejected at solver creation).  Velocity partials (cc:962-973) are never
// materialised: dvt_dqt[t] = N+_t/dt and dvt_dqm[t] = -N+_t/dt are applied in place.
#include <algorithm>

#include "reduce.cuh"
#include "solver.h"

namespace idto {

namespace {

constexpr int kAsmThreads = 160;
constexpr int kTile = 4;  // every thread accumulates a kTile x kTile tile of one product

// The eight nv x nq blocks a row needs (column-major, the layout they have in HBM) ...
enum Blk { kP0 = 0, kT0, kM1, kP1, kT1, kP2, kN0, kN1, kNumBlk };
// ... and the eight products X^T W Y over the nv rows.  The first five are symmetric (lower tiles only).
//   C_t     = Qq' + N0'N0 + P0'P0 + T0'T0 + M1'M1 + N1'N1
//   B_{t+1} = P1'T0 + T1'M1 - N1'N1          A_{t+2} = P2'M1
struct Prod {
  int left, right, sym;
};
__device__ __constant__ Prod kProds[8] = {{kN0, kN0, 1}, {kP0, kP0, 1}, {kT0, kT0, 1}, {kM1, kM1, 1},
                                          {kN1, kN1, 1}, {kP1, kT0, 0}, {kT1, kM1, 0}, {kP2, kM1, 0}};

}  // namespace

// One CTA per (b, t), t = 0..T: g_t, C_t, B_{t+1}, A_{t+2} and the scale factors D_t.
//
// The blocks go global -> shared as 1-D bulk TMA copies (they are contiguous in HBM and keep their
// layout); every thread then owns a 4x4 output tile of one product and walks the nv rows two at a
// time with 16-byte shared-memory loads: 8 LDS.128 per 32 FMA instead of the 2 LDS.64 per FMA of a
// thread-per-entry mapping (the first version was shared-memory-pipe bound at 163 us).  The row
// weights 2 dt R / 2 dt Qv / 2 Qf_v are applied to the left operand in registers, in the reference's
// operation order ((X^T W) Y, cc:1103-1165); the products land in shared memory and are summed in the
// reference's order.
__global__ void __launch_bounds__(kAsmThreads) k_assemble(SolverConsts sc, SolverBufs bf, int force, int alias) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.x / (sc.T + 1), t = blockIdx.x % (sc.T + 1);
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int T = sc.T, nq = sc.nq, nv = sc.nv, blk = nv * nq, tid = threadIdx.x, nt = blockDim.x;
  const double dt = sc.dt;
  const size_t pb = size_t(b) * T * blk;  // partial blocks of problem b
  double* g = bf.g + size_t(b) * sc.n + size_t(t) * nq;
  double* HA = bf.HA + (size_t(b) * (T + 1)) * nq * nq;
  double* HB = bf.HB + (size_t(b) * (T + 1)) * nq * nq;
  double* HC = bf.HC + (size_t(b) * (T + 1)) * nq * nq;
  double* D = bf.D + size_t(b) * sc.n + size_t(t) * nq;

  if (t == 0) {  // cc:1044, cc:1124: g_0 = 0, C_0 = I (B_1, A_2 keep their constructor zeros)
    for (int e = tid; e < nq * nq; e += nt) HC[e] = (e / nq == e % nq) ? 1.0 : 0.0;
    for (int e = tid; e < nq; e += nt) {
      g[e] = 0.0;
      if (sc.scaling) {
        const double hd = 1.0;
        switch (sc.scaling_method) {
          case IDTO_SCALING_SQRT: D[e] = fmin(1.0, 1 / sqrt(hd)); break;
          case IDTO_SCALING_ADAPTIVE_SQRT: D[e] = fmin(D[e], 1 / sqrt(hd)); break;
          case IDTO_SCALING_DOUBLE_SQRT: D[e] = fmin(1.0, 1 / sqrt(sqrt(hd))); break;
          default: D[e] = fmin(D[e], 1 / sqrt(sqrt(hd))); break;
        }
      }
    }
    return;
  }
  // ---- stage the blocks this row needs ----------------------------------------------------------
  const int blkp = (blk + 1) & ~1;            // block stride in shared memory (16-byte aligned)
  const int ntl = (nq + kTile - 1) / kTile;   // tiles per side
  const int npad = ntl * kTile;               // padded side of a product
  // `alias`: every thread has at most one tile, so the products can stay in registers until all reads of the
  // staged blocks are done and then overwrite them: 27 KB instead of 48 KB per CTA, twice the CTAs per SM
  double* sblk = sm;                          // [kNumBlk][blkp]
  const int nstage = kNumBlk * blkp, nprod = 8 * npad * npad;
  double* part = alias ? sm : sm + nstage;    // [8][npad * npad] products, column-major
  double* gpart = sm + (alias ? (nstage > nprod ? nstage : nprod) : nstage + nprod);  // [6][nq] gradient terms
  const double* Np = bf.st.Nplus + size_t(b) * (T + 1) * blk;
  const double* src[kNumBlk];
  src[kP0] = bf.dqp + pb + size_t(t - 1) * blk;
  src[kN0] = Np + size_t(t) * blk;
  src[kT0] = t < T ? bf.dqt + pb + size_t(t) * blk : nullptr;
  src[kP1] = t < T ? bf.dqp + pb + size_t(t) * blk : nullptr;
  src[kN1] = t < T ? Np + size_t(t + 1) * blk : nullptr;
  src[kM1] = t < T - 1 ? bf.dqm + pb + size_t(t + 1) * blk : nullptr;
  src[kT1] = t < T - 1 ? bf.dqt + pb + size_t(t + 1) * blk : nullptr;
  src[kP2] = t < T - 1 ? bf.dqp + pb + size_t(t + 1) * blk : nullptr;
  const bool bulk = (blk & 1) == 0;  // cp.async.bulk moves multiples of 16 bytes between 16-byte aligned addresses
  if (bulk) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      int nblk = 0;
#pragma unroll
      for (int k = 0; k < kNumBlk; ++k) nblk += src[k] != nullptr;
      mbar_arrive_expect_tx(&bar, uint32_t(nblk * blk * 8));
#pragma unroll
      for (int k = 0; k < kNumBlk; ++k)
        if (src[k]) tma_bulk_g2s(sblk + k * blkp, src[k], uint32_t(blk * 8), &bar);
    }
    mbar_wait(&bar, 0);
  } else {
#pragma unroll
    for (int k = 0; k < kNumBlk; ++k)
      if (src[k])
        for (int e = tid; e < blk; e += nt) sblk[k * blkp + e] = src[k][e];
    __syncthreads();
  }
  const double two_dt = 2 * dt, inv_dt = 1 / dt;
  const double* Qvn = (t == T - 1) ? sc.Qfv : sc.Qv;  // weight of the v_{t+1} term (cc:1054-1061, 1132-1147)
  const double Qvn_s = (t == T - 1) ? 2.0 : two_dt;
  // row weight of the left operand of product k, and the 1/dt that turns N+ into a velocity partial
  auto weight = [&](int k, int r) {
    if (k == 0) return t < T ? sc.Qv[r] * two_dt : sc.Qfv[r] * 2.0;
    if (k == 4) return Qvn[r] * Qvn_s;
    return sc.R[r] * two_dt;
  };
  auto live = [&](int k) {  // products that exist at this t
    if (t == T) return k <= 1;
    if (t == T - 1) return k != 3 && k != 6 && k != 7;
    return true;
  };

  // ---- products: one 4x4 tile per thread-task -----------------------------------------------------
  const int nsym = ntl * (ntl + 1) / 2, nfull = ntl * ntl;
  const int ntask = 5 * nsym + 3 * nfull;
  const bool even = (nv & 1) == 0;
  double acc[kTile][kTile];
  int held_k = -1, held_ti = 0, held_tj = 0;
  auto store_tile = [&](int k, int ti, int tj) {
    double* out = part + k * npad * npad;
#pragma unroll
    for (int u = 0; u < kTile; ++u)
#pragma unroll
      for (int v = 0; v < kTile; ++v) out[(tj + v * ntl) * npad + ti + u * ntl] = acc[u][v];
  };
  for (int task = tid; task < ntask; task += nt) {
    int k, ti, tj;
    if (task < 5 * nsym) {
      k = task / nsym;
      int rem = task - k * nsym;  // lower-triangle tile index -> (ti >= tj)
      ti = 0;
      while (rem >= ti + 1) rem -= ti + 1, ++ti;
      tj = rem;
    } else {
      const int rem0 = task - 5 * nsym;
      k = 5 + rem0 / nfull;
      const int rem = rem0 - (k - 5) * nfull;
      ti = rem % ntl, tj = rem / ntl;
    }
    if (!live(k)) continue;
    const Prod pr = kProds[k];
    const double* A = sblk + pr.left * blkp;
    const double* C = sblk + pr.right * blkp;
    const double sa = k == 0 || k == 4 ? inv_dt : 1.0;  // N+ / dt on both sides of the velocity products
    int ia[kTile], jc[kTile];
#pragma unroll
    for (int u = 0; u < kTile; ++u) {
      // INTERLEAVED tiles: thread tile (ti, tj) owns rows ti + u ntl and columns tj + v ntl.  Threads of a warp with
      // consecutive ti then read consecutive columns of a block (stride nv = 18 doubles: bank offset 4 words per
      // thread, conflict-free 16-byte loads); contiguous tiles put them 4 nv doubles apart, on two bank groups
      // (ncu: 56 % of the shared-memory wavefronts of this kernel were bank conflicts).
      ia[u] = min(ti + u * ntl, nq - 1) * nv;  // clamped: the padding rows are computed and discarded
      jc[u] = min(tj + u * ntl, nq - 1) * nv;
    }
#pragma unroll
    for (int u = 0; u < kTile; ++u)
#pragma unroll
      for (int v = 0; v < kTile; ++v) acc[u][v] = 0.0;
    if (even) {
      for (int r = 0; r < nv; r += 2) {
        const double w0 = weight(k, r), w1 = weight(k, r + 1);
        double2 a[kTile], c[kTile];
#pragma unroll
        for (int u = 0; u < kTile; ++u) {
          a[u] = *reinterpret_cast<const double2*>(A + ia[u] + r);
          c[u] = *reinterpret_cast<const double2*>(C + jc[u] + r);
        }
#pragma unroll
        for (int u = 0; u < kTile; ++u) {
          const double a0 = (a[u].x * sa) * w0, a1 = (a[u].y * sa) * w1;
#pragma unroll
          for (int v = 0; v < kTile; ++v) {
            acc[u][v] = fma(a0, c[v].x * sa, acc[u][v]);
            acc[u][v] = fma(a1, c[v].y * sa, acc[u][v]);
          }
        }
      }
    } else {
      for (int r = 0; r < nv; ++r) {
        const double w0 = weight(k, r);
#pragma unroll
        for (int u = 0; u < kTile; ++u) {
          const double a0 = (A[ia[u] + r] * sa) * w0;
#pragma unroll
          for (int v = 0; v < kTile; ++v) acc[u][v] = fma(a0, C[jc[v] + r] * sa, acc[u][v]);
        }
      }
    }
    if (alias)
      held_k = k, held_ti = ti, held_tj = tj;  // (single trip: ntask <= blockDim)
    else
      store_tile(k, ti, tj);
  }
  // ---- gradient terms (cc:1021-1081): six mat-vecs, one (term, column) per thread-task -----------
  const double* q = bf.st.q + (size_t(b) * (T + 1) + t) * nq;
  const double* qn = bf.q_nom + (size_t(b) * (T + 1) + t) * nq;
  const double* v = bf.st.v + (size_t(b) * (T + 1) + t) * nv;
  const double* vn = bf.v_nom + (size_t(b) * (T + 1) + t) * nv;
  const double* tau = bf.st.tau + size_t(b) * T * nv;
  for (int task = tid; task < 5 * nq; task += nt) {
    const int term = task / nq, j = task - term * nq;
    double x = 0.0;
    if (term == 0) {  // (Qv' v~_t)^T dvt_dqt[t]
      const double* N0 = sblk + kN0 * blkp + j * nv;
      if (t < T)
        for (int r = 0; r < nv; ++r) x += ((v[r] - vn[r]) * (sc.Qv[r] * two_dt)) * (N0[r] * inv_dt);
      else
        for (int r = 0; r < nv; ++r) x += ((v[r] - vn[r]) * (sc.Qfv[r] * 2)) * (N0[r] * inv_dt);
    } else if (term == 1) {  // (Q' v~_{t+1})^T dvt_dqm[t+1]
      const double* N1 = sblk + kN1 * blkp + j * nv;
      if (t < T)
        for (int r = 0; r < nv; ++r) x += ((v[nv + r] - vn[nv + r]) * (Qvn[r] * Qvn_s)) * (-(N1[r] * inv_dt));
    } else if (term == 2) {  // (R' tau_{t-1})^T dtau_dqp[t-1]
      const double* P0 = sblk + kP0 * blkp + j * nv;
      for (int r = 0; r < nv; ++r) x += (tau[(t - 1) * nv + r] * (sc.R[r] * two_dt)) * P0[r];
    } else if (term == 3) {  // (R' tau_t)^T dtau_dqt[t]
      const double* T0 = sblk + kT0 * blkp + j * nv;
      if (t < T)
        for (int r = 0; r < nv; ++r) x += (tau[t * nv + r] * (sc.R[r] * two_dt)) * T0[r];
    } else {  // (R' tau_{t+1})^T dtau_dqm[t+1]
      const double* M1 = sblk + kM1 * blkp + j * nv;
      if (t < T - 1)
        for (int r = 0; r < nv; ++r) x += (tau[(t + 1) * nv + r] * (sc.R[r] * two_dt)) * M1[r];
    }
    gpart[term * nq + j] = x;
  }
  __syncthreads();
  if (alias) {  // all reads of the staged blocks are done: the tiles may overwrite them
    if (held_k >= 0) store_tile(held_k, held_ti, held_tj);
    __syncthreads();
  }

  // ---- Hessian bands: sum the products in the reference's order; MakeSymmetric -------------------
  auto P = [&](int k, int i, int j) { return part[k * npad * npad + j * npad + i]; };
  // symmetric products: only the tiles with ti >= tj are computed; entry (i, j) lies in tile (i % ntl, j % ntl)
  auto S = [&](int k, int i, int j) { return (i % ntl) >= (j % ntl) ? P(k, i, j) : P(k, j, i); };
  auto Psym = [&](int k, int i, int j) { return i >= j ? S(k, i, j) : S(k, j, i); };
  for (int e = tid; e < nq * nq; e += nt) {
    const int j = e / nq, i = e % nq;  // column-major: entry (i, j)
    const int il = i >= j ? i : j, jl = i >= j ? j : i;  // penta_diagonal_matrix.cc:74-76: upper of C := lower
    if (t < T) {
      double c = (i == j) ? sc.Qq[i] * two_dt : 0.0;
      c += S(0, il, jl);
      c += S(1, il, jl);
      c += S(2, il, jl);
      if (t < T - 1) c += S(3, il, jl);
      c += S(4, il, jl);
      HC[size_t(t) * nq * nq + e] = c;
      double bb = P(5, i, j);
      if (t < T - 1) bb += P(6, i, j);
      bb += -Psym(4, i, j);  // dvt_dqt[t+1]^T Q dvt_dqm[t+1] = -(N/dt)^T Q (N/dt)
      HB[size_t(t + 1) * nq * nq + e] = bb;
      if (t < T - 1) HA[size_t(t + 2) * nq * nq + e] = P(7, i, j);
    } else {  // cc:1157-1161
      double c = (i == j) ? sc.Qfq[i] * 2 : 0.0;
      c += S(0, il, jl);
      c += S(1, il, jl);
      HC[size_t(t) * nq * nq + e] = c;
    }
  }
  // ---- scale factors (cc:1235-1254) and gradient ----------------------------------------------------
  for (int e = tid; e < nq; e += nt) {
    if (sc.scaling) {
      double hd = t < T ? sc.Qq[e] * two_dt : sc.Qfq[e] * 2;
      hd += S(0, e, e);
      hd += S(1, e, e);
      if (t < T) {
        hd += S(2, e, e);
        if (t < T - 1) hd += S(3, e, e);
        hd += S(4, e, e);
      }
      switch (sc.scaling_method) {
        case IDTO_SCALING_SQRT: D[e] = fmin(1.0, 1 / sqrt(hd)); break;
        case IDTO_SCALING_ADAPTIVE_SQRT: D[e] = fmin(D[e], 1 / sqrt(hd)); break;
        case IDTO_SCALING_DOUBLE_SQRT: D[e] = fmin(1.0, 1 / sqrt(sqrt(hd))); break;
        default: D[e] = fmin(D[e], 1 / sqrt(sqrt(hd))); break;
      }
    }
    double gj;
    if (t < T) {
      gj = (q[e] - qn[e]) * (sc.Qq[e] * two_dt);
      gj += gpart[0 * nq + e];
      gj += gpart[1 * nq + e];
      gj += gpart[2 * nq + e];
      gj += gpart[3 * nq + e];
      if (t != T - 1) gj += gpart[4 * nq + e];
    } else {  // cc:1074-1080
      gj = gpart[2 * nq + e];
      gj += (q[e] - qn[e]) * (sc.Qfq[e] * 2);
      gj += gpart[0 * nq + e];
    }
    g[e] = gj;
  }
}


// Dense cost weights (ProblemDefinition's MatrixXd, problem_definition.h:38-52): the general form of k_assemble,
// one thread per entry, no tiling — every shipped example has diagonal weights and takes k_assemble.  Same
// expressions and summation order as the reference: gradient cc:1021-1081, Hessian cc:1093-1165 (X^T (W Y) with
// W = 2 dt Q or 2 Q_f), MakeSymmetric (penta_diagonal_matrix.cc:64-105), scale factors cc:1235-1254.
__global__ void __launch_bounds__(128) k_assemble_dense(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x / (sc.T + 1), t = blockIdx.x % (sc.T + 1);
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int T = sc.T, nq = sc.nq, nv = sc.nv, blk = nv * nq, tid = threadIdx.x, nt = blockDim.x;
  const double dt = sc.dt, two_dt = 2 * dt, inv_dt = 1 / dt;
  const size_t pb = size_t(b) * T * blk;
  double* g = bf.g + size_t(b) * sc.n + size_t(t) * nq;
  double* HA = bf.HA + (size_t(b) * (T + 1)) * nq * nq;
  double* HB = bf.HB + (size_t(b) * (T + 1)) * nq * nq;
  double* HC = bf.HC + (size_t(b) * (T + 1)) * nq * nq;
  double* D = bf.D + size_t(b) * sc.n + size_t(t) * nq;
  auto scale_factor = [&](int e, double hd) {
    switch (sc.scaling_method) {
      case IDTO_SCALING_SQRT: D[e] = fmin(1.0, 1 / sqrt(hd)); break;
      case IDTO_SCALING_ADAPTIVE_SQRT: D[e] = fmin(D[e], 1 / sqrt(hd)); break;
      case IDTO_SCALING_DOUBLE_SQRT: D[e] = fmin(1.0, 1 / sqrt(sqrt(hd))); break;
      default: D[e] = fmin(D[e], 1 / sqrt(sqrt(hd))); break;
    }
  };
  if (t == 0) {  // cc:1044, cc:1124: g_0 = 0, C_0 = I
    for (int e = tid; e < nq * nq; e += nt) HC[e] = (e / nq == e % nq) ? 1.0 : 0.0;
    for (int e = tid; e < nq; e += nt) {
      g[e] = 0.0;
      if (sc.scaling) scale_factor(e, 1.0);
    }
    return;
  }
  double* sblk = sm;                    // [kNumBlk][blk]   (velocity blocks already divided by dt)
  double* wy = sblk + kNumBlk * blk;    // [5][blk]  W Y for Y = N0, P0, T0, M1, N1
  double* prod = wy + 5 * blk;          // [8][nq * nq]
  double* wvec = prod + 8 * nq * nq;    // [6][max(nq, nv)]  weighted error vectors of the gradient
  const int mx = nq > nv ? nq : nv;
  const double* Np = bf.st.Nplus + size_t(b) * (T + 1) * blk;
  const double* src[kNumBlk];
  src[kP0] = bf.dqp + pb + size_t(t - 1) * blk;
  src[kN0] = Np + size_t(t) * blk;
  src[kT0] = t < T ? bf.dqt + pb + size_t(t) * blk : nullptr;
  src[kP1] = t < T ? bf.dqp + pb + size_t(t) * blk : nullptr;
  src[kN1] = t < T ? Np + size_t(t + 1) * blk : nullptr;
  src[kM1] = t < T - 1 ? bf.dqm + pb + size_t(t + 1) * blk : nullptr;
  src[kT1] = t < T - 1 ? bf.dqt + pb + size_t(t + 1) * blk : nullptr;
  src[kP2] = t < T - 1 ? bf.dqp + pb + size_t(t + 1) * blk : nullptr;
  for (int k = 0; k < kNumBlk; ++k)
    for (int e = tid; e < blk; e += nt) {
      double x = src[k] ? src[k][e] : 0.0;
      if (k == kN0 || k == kN1) x *= inv_dt;  // dvt_dqt = N+ / dt; dvt_dqm = -N+ / dt (sign applied below)
      sblk[k * blk + e] = x;
    }
  // weights of this row (cc:1103-1107): W0 for the v_t term, W4 for the v_{t+1} term, R' for the input terms
  const double* W0 = t < T ? sc.QvM : sc.QfvM;
  const double s0 = t < T ? two_dt : 2.0;
  const double* W4 = (t == T - 1) ? sc.QfvM : sc.QvM;
  const double s4 = (t == T - 1) ? 2.0 : two_dt;
  __syncthreads();
  // W Y: wy[k][r][j] = sum_s W[r][s] Y[s][j]
  const int ysrc[5] = {kN0, kP0, kT0, kM1, kN1};
  for (int e = tid; e < 5 * blk; e += nt) {
    const int k = e / blk, rem = e - k * blk, j = rem / nv, r = rem - j * nv;
    const double* W = k == 0 ? W0 : (k == 4 ? W4 : sc.RM);
    const double ws = k == 0 ? s0 : (k == 4 ? s4 : two_dt);
    const double* Y = sblk + ysrc[k] * blk + j * nv;
    double acc = 0.0;
    for (int s = 0; s < nv; ++s) acc += (W[size_t(s) * nv + r] * ws) * Y[s];
    wy[e] = acc;
  }
  __syncthreads();
  // products X^T (W Y): N0'WN0, P0'WP0, T0'WT0, M1'WM1, N1'WN1, P1'WT0, T1'WM1, P2'WM1
  const int xl[8] = {kN0, kP0, kT0, kM1, kN1, kP1, kT1, kP2}, yr[8] = {0, 1, 2, 3, 4, 2, 3, 3};
  for (int e = tid; e < 8 * nq * nq; e += nt) {
    const int k = e / (nq * nq), rem = e - k * nq * nq, j = rem / nq, i = rem - j * nq;
    const double* X = sblk + xl[k] * blk + i * nv;
    const double* Z = wy + yr[k] * blk + j * nv;
    double acc = 0.0;
    for (int r = 0; r < nv; ++r) acc += X[r] * Z[r];
    prod[e] = acc;
  }
  // weighted error vectors of the gradient (row vectors e^T W, cc:1049-1069)
  const double* q = bf.st.q + (size_t(b) * (T + 1) + t) * nq;
  const double* qn = bf.q_nom + (size_t(b) * (T + 1) + t) * nq;
  const double* v = bf.st.v + (size_t(b) * (T + 1) + t) * nv;
  const double* vn = bf.v_nom + (size_t(b) * (T + 1) + t) * nv;
  const double* tau = bf.st.tau + size_t(b) * T * nv;
  for (int e = tid; e < 6 * mx; e += nt) {
    const int k = e / mx, j = e - k * mx;
    double acc = 0.0;
    if (k == 0 && j < nq) {  // q~^T Qq' (Qf_q' at t = T)
      const double* W = t < T ? sc.QqM : sc.QfqM;
      for (int i = 0; i < nq; ++i) acc += (q[i] - qn[i]) * (W[size_t(j) * nq + i] * (t < T ? two_dt : 2.0));
    } else if (k == 1 && j < nv) {  // v~_t^T W0
      for (int i = 0; i < nv; ++i) acc += (v[i] - vn[i]) * (W0[size_t(j) * nv + i] * s0);
    } else if (k == 2 && j < nv && t < T) {  // v~_{t+1}^T W4
      for (int i = 0; i < nv; ++i) acc += (v[nv + i] - vn[nv + i]) * (W4[size_t(j) * nv + i] * s4);
    } else if (k >= 3 && j < nv) {  // tau_{t-1}, tau_t, tau_{t+1} times R'
      const int tt = t - 1 + (k - 3);
      if (tt < T)
        for (int i = 0; i < nv; ++i) acc += tau[tt * nv + i] * (sc.RM[size_t(j) * nv + i] * two_dt);
    }
    wvec[e] = acc;
  }
  __syncthreads();
  auto P = [&](int k, int i, int j) { return prod[(k * nq + j) * nq + i]; };
  const double* QqM = t < T ? sc.QqM : sc.QfqM;
  const double sq = t < T ? two_dt : 2.0;
  for (int e = tid; e < nq * nq; e += nt) {
    const int j = e / nq, i = e % nq;
    const int il = i >= j ? i : j, jl = i >= j ? j : i;  // penta_diagonal_matrix.cc:74-76: upper of C := lower
    double c = QqM[size_t(jl) * nq + il] * sq;
    c += P(0, il, jl);
    c += P(1, il, jl);
    if (t < T) {
      c += P(2, il, jl);
      if (t < T - 1) c += P(3, il, jl);
      c += P(4, il, jl);  // dvt_dqm' W dvt_dqm: the two signs cancel
      double bb = P(5, i, j);
      if (t < T - 1) bb += P(6, i, j);
      bb += -P(4, i, j);  // dvt_dqt[t+1]^T W dvt_dqm[t+1] = -(N/dt)^T W (N/dt)
      HB[size_t(t + 1) * nq * nq + e] = bb;
      if (t < T - 1) HA[size_t(t + 2) * nq * nq + e] = P(7, i, j);
    }
    HC[size_t(t) * nq * nq + e] = c;
  }
  for (int e = tid; e < nq; e += nt) {
    if (sc.scaling) {
      double hd = QqM[size_t(e) * nq + e] * sq;
      hd += P(0, e, e);
      hd += P(1, e, e);
      if (t < T) {
        hd += P(2, e, e);
        if (t < T - 1) hd += P(3, e, e);
        hd += P(4, e, e);
      }
      scale_factor(e, hd);
    }
    auto dotcol = [&](int k, int blkid, double sign) {  // (weighted row vector k) . column e of a block
      const double* X = sblk + blkid * blk + e * nv;
      double acc = 0.0;
      for (int r = 0; r < nv; ++r) acc += wvec[k * mx + r] * (sign * X[r]);
      return acc;
    };
    double gj;
    if (t < T) {
      gj = wvec[e];
      gj += dotcol(1, kN0, 1.0);
      gj += dotcol(2, kN1, -1.0);
      gj += dotcol(3, kP0, 1.0);
      gj += dotcol(4, kT0, 1.0);
      if (t != T - 1) gj += dotcol(5, kM1, 1.0);
    } else {  // cc:1074-1080
      gj = dotcol(3, kP0, 1.0);
      gj += wvec[e];
      gj += dotcol(1, kN0, 1.0);
    }
    g[e] = gj;
  }
}

// One CTA per (b, t): scaled bands H~ = D H D (lower bands), g~ = D g, J~ bands = rows of the ID
// partials of the unactuated dofs times D.  With scaling off D == 1 and this is a copy.
__global__ void __launch_bounds__(128) k_scale(SolverConsts sc, SolverBufs bf, int force) {
  const int b = blockIdx.x / (sc.T + 1), t = blockIdx.x % (sc.T + 1);
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int T = sc.T, nq = sc.nq, nv = sc.nv, nu = sc.nu, tid = threadIdx.x, nt = blockDim.x;
  const double* D = bf.D + size_t(b) * sc.n;
  const size_t hb = (size_t(b) * (T + 1) + t) * nq * nq;
#pragma unroll 4
  for (int e = tid; e < nq * nq; e += nt) {
    const int c = e / nq, r = e % nq;
    const double dr = D[t * nq + r];
    bf.SC[hb + e] = dr * bf.HC[hb + e] * D[t * nq + c];
    bf.SB[hb + e] = t >= 1 ? dr * bf.HB[hb + e] * D[(t - 1) * nq + c] : 0.0;
    bf.SA[hb + e] = t >= 2 ? dr * bf.HA[hb + e] * D[(t - 2) * nq + c] : 0.0;
  }
  for (int e = tid; e < nq; e += nt) {
    const size_t idx = size_t(b) * sc.n + size_t(t) * nq + e;
    bf.gs[idx] = D[t * nq + e] * bf.g[idx];
  }
  if (t < T && nu > 0) {
    const size_t pb = (size_t(b) * T + t) * nv * nq, jb = (size_t(b) * T + t) * nu * nq;
#pragma unroll 2
    for (int e = tid; e < nu * nq; e += nt) {
      const int u = e / nq, c = e % nq, row = sc.unact[u];
      bf.Jp[jb + e] = bf.dqp[pb + size_t(c) * nv + row] * D[(t + 1) * nq + c];
      bf.Jt[jb + e] = t > 0 ? bf.dqt[pb + size_t(c) * nv + row] * D[t * nq + c] : 0.0;
      bf.Jm[jb + e] = t > 1 ? bf.dqm[pb + size_t(c) * nv + row] * D[(t - 1) * nq + c] : 0.0;
    }
  }
}

void launch_assemble(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                     cudaStream_t stream) {
  (void)dm;
  const int blkp = (sc.nv * sc.nq + 1) & ~1, npad = (sc.nq + kTile - 1) / kTile * kTile;
  const int ntl = npad / kTile, ntask = 5 * (ntl * (ntl + 1) / 2) + 3 * ntl * ntl;
  const int alias = ntask <= kAsmThreads ? 1 : 0;
  const int nstage = kNumBlk * blkp, nprod = 8 * npad * npad;
  const int smem = ((alias ? std::max(nstage, nprod) : nstage + nprod) + 6 * sc.nq) * 8;
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  g_launch_counter += 2;
  if (sc.dense_w) {
    const int mx = std::max(sc.nq, sc.nv);
    const int dsmem = ((kNumBlk + 5) * sc.nv * sc.nq + 8 * sc.nq * sc.nq + 6 * mx) * 8;
    static bool dattr[kMaxDevices] = {};
    if (first_use_on_device(dattr))
      cudaFuncSetAttribute(k_assemble_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_assemble_dense<<<sc.B*(sc.T + 1), 128, dsmem, stream>>>(sc, bf, force ? 1 : 0);
  } else
    k_assemble<<<sc.B*(sc.T + 1), kAsmThreads, smem, stream>>>(sc, bf, force ? 1 : 0, alias);
  k_scale<<<sc.B*(sc.T + 1), 128, 0, stream>>>(sc, bf, force ? 1 : 0);
}

}  // namespace idto
