// Gradient (cc:1021-1081), Gauss-Newton Hessian bands (cc:1093-1165), scale factors (cc:1225-1255),
// scaled Hessian / gradient (cc:1181-1223, penta_diagonal_matrix.cc:221-257) and the scaled
// equality-constraint Jacobian bands (cc:1292-1334).  Diagonal cost weights (every reference example;
// dense weights are rejected at solver creation).  Velocity partials (cc:962-973) are never
// materialised: dvt_dqt[t] = N+_t/dt and dvt_dqm[t] = -N+_t/dt are applied in place.
#include "reduce.cuh"
#include "solver.h"

namespace idto {

namespace {

// sum_r Aw(r,i) * C(r,j) over nv rows, where Aw(r,i) = A(r,i) * (w_r * ws) was staged once (same
// operation order as the reference's (A^T W) C products); column-major nv x nq in shared memory.
__device__ __forceinline__ double wdot(const double* Aw, const double* C, int nv, int ld, int i, int j) {
  double acc = 0.0;
  const double* a = Aw + i * ld;
  const double* c = C + j * ld;
#pragma unroll 6
  for (int r = 0; r < nv; ++r) acc += a[r] * c[r];
  return acc;
}

}  // namespace

// One CTA per (b, t), t = 0..T: g_t, C_t, B_{t+1}, A_{t+2} and the scale factors D_t.
__global__ void __launch_bounds__(128) k_assemble(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x / (sc.T + 1), t = blockIdx.x % (sc.T + 1);
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int T = sc.T, nq = sc.nq, nv = sc.nv, blk = nv * nq, tid = threadIdx.x, nt = blockDim.x;
  const int ld = nv | 1, sblk = ld * nq;  // odd leading dimension in shared memory: conflict-free columns
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
  // stage the blocks this row needs: P_{t-1}, Tt_t, M_{t+1}, P_t, Tt_{t+1}, P_{t+1}, N_t, N_{t+1}
  double* sP0 = sm;             // dqp[t-1]
  double* sT0 = sP0 + sblk;      // dqt[t]
  double* sM1 = sT0 + sblk;      // dqm[t+1]
  double* sP1 = sM1 + sblk;      // dqp[t]
  double* sT1 = sP1 + sblk;      // dqt[t+1]
  double* sP2 = sT1 + sblk;      // dqp[t+1]
  double* sN0 = sP2 + sblk;      // N+_t / dt
  double* sN1 = sN0 + sblk;      // N+_{t+1} / dt
  double* wP0 = sN1 + sblk;      // weighted copies (left operands): R' P_{t-1}, R' Tt_t, R' M_{t+1}, ...
  double* wT0 = wP0 + sblk;
  double* wM1 = wT0 + sblk;
  double* wP1 = wM1 + sblk;
  double* wT1 = wP1 + sblk;
  double* wP2 = wT1 + sblk;
  double* wN0 = wP2 + sblk;      // Qv' N+_t/dt  (Qf_v' at t = T)
  double* wN1 = wN0 + sblk;      // Qv' (or Qf_v') N+_{t+1}/dt
  double* sC = wN1 + sblk;       // C_t staging (for the diagonal)
  const double* Np = bf.st.Nplus + size_t(b) * (T + 1) * blk;
  for (int e = tid; e < blk; e += nt) {
    const int se = (e / nv) * ld + e % nv;
    sP0[se] = bf.dqp[pb + size_t(t - 1) * blk + e];
    sN0[se] = Np[size_t(t) * blk + e] * (1 / dt);
    if (t < T) {
      sT0[se] = bf.dqt[pb + size_t(t) * blk + e];
      sP1[se] = bf.dqp[pb + size_t(t) * blk + e];
      sN1[se] = Np[size_t(t + 1) * blk + e] * (1 / dt);
    }
    if (t < T - 1) {
      sM1[se] = bf.dqm[pb + size_t(t + 1) * blk + e];
      sT1[se] = bf.dqt[pb + size_t(t + 1) * blk + e];
      sP2[se] = bf.dqp[pb + size_t(t + 1) * blk + e];
    }
  }
  __syncthreads();
  const double two_dt = 2 * dt;
  const double* Qvn = (t == T - 1) ? sc.Qfv : sc.Qv;  // weight of the v_{t+1} term (cc:1054-1061, 1132-1147)
  const double Qvn_s = (t == T - 1) ? 2.0 : two_dt;
  for (int e0 = tid; e0 < blk; e0 += nt) {
    const int r = e0 % nv, e = (e0 / nv) * ld + r;
    const double wr = sc.R[r] * two_dt;
    wP0[e] = sP0[e] * wr;
    wN0[e] = sN0[e] * (t < T ? sc.Qv[r] * two_dt : sc.Qfv[r] * 2.0);
    if (t < T) {
      wT0[e] = sT0[e] * wr;
      wP1[e] = sP1[e] * wr;
      wN1[e] = sN1[e] * (Qvn[r] * Qvn_s);
    }
    if (t < T - 1) {
      wM1[e] = sM1[e] * wr;
      wT1[e] = sT1[e] * wr;
      wP2[e] = sP2[e] * wr;
    }
  }
  __syncthreads();

  // ---- Hessian bands --------------------------------------------------------------------------
  for (int e = tid; e < nq * nq; e += nt) {
    const int j = e / nq, i = e % nq;  // column-major: entry (i, j)
    if (t < T) {
      double c = (i == j) ? sc.Qq[i] * two_dt : 0.0;
      c += wdot(wN0, sN0, nv, ld, i, j);
      c += wdot(wP0, sP0, nv, ld, i, j);
      c += wdot(wT0, sT0, nv, ld, i, j);
      if (t < T - 1) {
        c += wdot(wM1, sM1, nv, ld, i, j);
        c += wdot(wN1, sN1, nv, ld, i, j);
      } else {
        c += wdot(wN1, sN1, nv, ld, i, j);
      }
      sC[e] = c;
      // B_{t+1}: dg_t/dq_{t+1}
      double bb = wdot(wP1, sT0, nv, ld, i, j);
      if (t < T - 1) bb += wdot(wT1, sM1, nv, ld, i, j);
      bb += -wdot(wN1, sN1, nv, ld, i, j);  // dvt_dqt[t+1]^T Q dvt_dqm[t+1] = -(N/dt)^T Q (N/dt)
      HB[size_t(t + 1) * nq * nq + e] = bb;
      if (t < T - 1) HA[size_t(t + 2) * nq * nq + e] = wdot(wP2, sM1, nv, ld, i, j);
    } else {  // cc:1157-1161
      double c = (i == j) ? sc.Qfq[i] * 2 : 0.0;
      c += wdot(wN0, sN0, nv, ld, i, j);
      c += wdot(wP0, sP0, nv, ld, i, j);
      sC[e] = c;
    }
  }
  __syncthreads();
  // MakeSymmetric (penta_diagonal_matrix.cc:74-76): strictly-upper of C := lower
  for (int e = tid; e < nq * nq; e += nt) {
    const int j = e / nq, i = e % nq;
    HC[size_t(t) * nq * nq + e] = (i >= j) ? sC[e] : sC[i * nq + j];
  }
  // ---- scale factors (cc:1235-1254) ------------------------------------------------------------
  if (sc.scaling) {
    for (int e = tid; e < nq; e += nt) {
      const double hd = sC[e * nq + e];
      switch (sc.scaling_method) {
        case IDTO_SCALING_SQRT: D[e] = fmin(1.0, 1 / sqrt(hd)); break;
        case IDTO_SCALING_ADAPTIVE_SQRT: D[e] = fmin(D[e], 1 / sqrt(hd)); break;
        case IDTO_SCALING_DOUBLE_SQRT: D[e] = fmin(1.0, 1 / sqrt(sqrt(hd))); break;
        default: D[e] = fmin(D[e], 1 / sqrt(sqrt(hd))); break;
      }
    }
  }
  // ---- gradient ---------------------------------------------------------------------------------
  const double* q = bf.st.q + (size_t(b) * (T + 1) + t) * nq;
  const double* qn = bf.q_nom + (size_t(b) * (T + 1) + t) * nq;
  const double* v = bf.st.v + (size_t(b) * (T + 1) + t) * nv;
  const double* vn = bf.v_nom + (size_t(b) * (T + 1) + t) * nv;
  const double* tau = bf.st.tau + size_t(b) * T * nv;
  for (int j = tid; j < nq; j += nt) {
    double gj;
    if (t < T) {
      gj = (q[j] - qn[j]) * (sc.Qq[j] * two_dt);
      double x = 0.0;
      for (int r = 0; r < nv; ++r) x += ((v[r] - vn[r]) * (sc.Qv[r] * two_dt)) * sN0[j * ld + r];
      gj += x;
      x = 0.0;
      for (int r = 0; r < nv; ++r) x += ((v[nv + r] - vn[nv + r]) * (Qvn[r] * Qvn_s)) * (-sN1[j * ld + r]);
      gj += x;
      x = 0.0;
      for (int r = 0; r < nv; ++r) x += (tau[(t - 1) * nv + r] * (sc.R[r] * two_dt)) * sP0[j * ld + r];
      gj += x;
      x = 0.0;
      for (int r = 0; r < nv; ++r) x += (tau[t * nv + r] * (sc.R[r] * two_dt)) * sT0[j * ld + r];
      gj += x;
      if (t != T - 1) {
        x = 0.0;
        for (int r = 0; r < nv; ++r) x += (tau[(t + 1) * nv + r] * (sc.R[r] * two_dt)) * sM1[j * ld + r];
        gj += x;
      }
    } else {  // cc:1074-1080
      double x = 0.0;
      for (int r = 0; r < nv; ++r) x += (tau[(T - 1) * nv + r] * (sc.R[r] * two_dt)) * sP0[j * ld + r];
      gj = x;
      gj += (q[j] - qn[j]) * (sc.Qfq[j] * 2);
      x = 0.0;
      for (int r = 0; r < nv; ++r) x += ((v[r] - vn[r]) * (sc.Qfv[r] * 2)) * sN0[j * ld + r];
      gj += x;
    }
    g[j] = gj;
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
  const int smem = (16 * (sc.nv | 1) * sc.nq + sc.nq * sc.nq) * 8;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  g_launch_counter += 2;
  k_assemble<<<sc.B*(sc.T + 1), 128, smem, stream>>>(sc, bf, force ? 1 : 0);
  k_scale<<<sc.B*(sc.T + 1), 128, 0, stream>>>(sc, bf, force ? 1 : 0);
}

}  // namespace idto
