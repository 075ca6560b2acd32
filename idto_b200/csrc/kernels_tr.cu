// Trust-region scalars kept on the device: dogleg point (cc:2108-2202), trust ratio (cc:1979-2035),
// step acceptance, Delta update, convergence checks and stats (cc:2494-2625).  One CTA per problem; the
// host never reads a scalar back during a solve.
#include "reduce.cuh"
#include "solver.h"

namespace idto {

namespace {

// y = H~ x for the symmetric block penta-diagonal matrix given by its lower bands
// (PentaDiagonalMatrix::MultiplyBy, penta_diagonal_matrix.cc:181-207; D_i = B_{i+1}^T, E_i = A_{i+2}^T).
__device__ __forceinline__ void penta_matvec(const SolverConsts& sc, const double* SA, const double* SB,
                                             const double* SC, const double* x, double* y, int tid, int nt) {
  const int k = sc.nq, kk = k * k, nblk = sc.T + 1;
  for (int e = tid; e < sc.n; e += nt) {
    const int i = e / k, r = e % k;
    double acc = 0.0;
    const double* C = SC + size_t(i) * kk;
    for (int c = 0; c < k; ++c) acc += C[c * k + r] * x[i * k + c];
    if (i >= 1) {
      const double* Bm = SB + size_t(i) * kk;
      for (int c = 0; c < k; ++c) acc += Bm[c * k + r] * x[(i - 1) * k + c];
    }
    if (i >= 2) {
      const double* Am = SA + size_t(i) * kk;
      for (int c = 0; c < k; ++c) acc += Am[c * k + r] * x[(i - 2) * k + c];
    }
    if (i < nblk - 1) {
      const double* Bn = SB + size_t(i + 1) * kk;  // D_i(r,c) = B_{i+1}(c,r)
      for (int c = 0; c < k; ++c) acc += Bn[r * k + c] * x[(i + 1) * k + c];
    }
    if (i < nblk - 2) {
      const double* An = SA + size_t(i + 2) * kk;
      for (int c = 0; c < k; ++c) acc += An[r * k + c] * x[(i + 2) * k + c];
    }
    y[e] = acc;
  }
}

}  // namespace

// Dogleg part 1: Hg = H~ gm, gHg, g.g.
__global__ void __launch_bounds__(256) k_dogleg_pre(SolverConsts sc, SolverBufs bf) {
  __shared__ double red[32];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, n = sc.n;
  if (!bf.ctl[b].active) return;
  const size_t hb = size_t(b) * (sc.T + 1) * sc.nq * sc.nq;
  const double* gm = bf.gm + size_t(b) * n;
  double* Hg = bf.tmp1 + size_t(b) * n;
  penta_matvec(sc, bf.SA + hb, bf.SB + hb, bf.SC + hb, gm, Hg, tid, nt);
  __syncthreads();
  double gHg = 0.0, gg = 0.0;
  for (int e = tid; e < n; e += nt) gHg += gm[e] * Hg[e], gg += gm[e] * gm[e];
  gHg = block_sum(gHg, red);
  gg = block_sum(gg, red);
  if (tid == 0) bf.red[b * 8 + 0] = gHg, bf.red[b * 8 + 1] = gg;
}

// Dogleg part 2: pH = H~^-1(-gm/Delta) = x/Delta with x from k_kkt_solve (cc:2139-2140; the solve is
// linear in the right-hand side, so x is computed once per derivative update and reused by rejected
// steps); branch logic, dq, dqH, logging scalars, and the
// scratch trajectory q + dq (cc:1991-1993).
__global__ void __launch_bounds__(256) k_dogleg_post(SolverConsts sc, SolverBufs bf) {
  __shared__ double red[32];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, n = sc.n;
  ProbCtl* ctl = bf.ctl + b;
  if (!ctl->active) return;
  const double Delta = ctl->Delta;
  const double gHg = bf.red[b * 8 + 0], gg = bf.red[b * 8 + 1];
  const double* gm = bf.gm + size_t(b) * n;
  const double* pH = bf.pH + size_t(b) * n;
  const double* D = bf.D + size_t(b) * n;
  double* dq = bf.dq + size_t(b) * n;
  double* dqH = bf.dqH + size_t(b) * n;
  // pU = -(g.g / gHg) g / Delta  (cc:2157)
  double pU2 = 0.0, pH2 = 0.0, a = 0.0, bq = 0.0;
  for (int e = tid; e < n; e += nt) {
    const double pu = -(gg / gHg) * gm[e] / Delta, ph = pH[e] / Delta;
    pU2 += pu * pu, pH2 += ph * ph;
    const double d = ph - pu;
    a += d * d, bq += pu * d;
  }
  pU2 = block_sum(pU2, red), pH2 = block_sum(pH2, red), a = block_sum(a, red), bq = block_sum(bq, red);
  const double pUn = sqrt(pU2), pHn = sqrt(pH2);
  int active;
  double s = 0.0;
  int branch;
  if (1.0 <= pUn) {  // cc:2160-2168
    branch = 0, active = 1;
  } else if (1.0 >= pHn) {  // cc:2171-2178
    branch = 1, active = 0;
  } else {  // cc:2192-2201 + SolveDoglegQuadratic cc:2037-2066
    branch = 2, active = 1;
    const double bb = 2 * bq, cc = pU2 - 1.0;
    if (a < 2.220446049250313e-16) {
      s = -cc / bb;
    } else {
      const double b_tilde = bb / a, c_tilde = cc / a;
      s = (-b_tilde + sqrt(b_tilde * b_tilde - 4 * c_tilde)) / 2;
    }
  }
  double dq2 = 0.0, dqH2 = 0.0, gdq = 0.0, q2 = 0.0;
  const double* q = bf.st.q + size_t(b) * n;
  double* qs = bf.sc.q + size_t(b) * n;
  for (int e = tid; e < n; e += nt) {
    const double pu = -(gg / gHg) * gm[e] / Delta, ph = pH[e] / Delta;
    double x;
    if (branch == 0)
      x = (Delta / pUn) * pu;
    else if (branch == 1)
      x = ph * Delta;
    else
      x = (pu + s * (ph - pu)) * Delta;
    if (sc.scaling) x = D[e] * x;
    const double xh = ph * Delta;  // cc:2152 (dqH is NOT rescaled by D in the reference)
    dq[e] = x, dqH[e] = xh;
    dq2 += x * x, dqH2 += xh * xh;
    gdq += sc.scaling ? gm[e] * ((1.0 / D[e]) * x) : gm[e] * x;  // cc:2518-2523
    q2 += q[e] * q[e];
    qs[e] = q[e] + x;  // scratch_state.set_q(q); AddToQ(dq)
  }
  dq2 = block_sum(dq2, red), dqH2 = block_sum(dqH2, red), gdq = block_sum(gdq, red), q2 = block_sum(q2, red);
  __syncthreads();
  if (sc.normalize_quat) {  // cc:1993 + cc:2691-2707
    for (int idx = tid; idx < (sc.T + 1) * sc.nquat; idx += nt) {
      const int t = idx / sc.nquat, qi = sc.quat_starts[idx % sc.nquat];
      double* x = qs + size_t(t) * sc.nq + qi;
      const double nrm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
      x[0] /= nrm, x[1] /= nrm, x[2] /= nrm, x[3] /= nrm;
    }
  }
  if (tid == 0) {
    ctl->tr_active = active;
    ctl->dq_norm = sqrt(dq2), ctl->dqH_norm = sqrt(dqH2), ctl->q_norm = sqrt(q2);
    ctl->dL_dq = gdq / bf.st.cost[b];
    ctl->gnorm = sqrt(gg);
  }
}

// Trust ratio (cc:1979-2035), acceptance (cc:2550-2553), stats (cc:2586-2598), convergence
// (cc:2601-2611, 2653-2689) and the Delta update (cc:2613-2622).  `commit`=0 only evaluates rho.
__global__ void __launch_bounds__(256) k_trust_update(SolverConsts sc, SolverBufs bf, int commit) {
  __shared__ double red[32];
  __shared__ int s_accept;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, n = sc.n, nh = sc.nh;
  ProbCtl* ctl = bf.ctl + b;
  if (!ctl->active) return;
  const size_t hb = size_t(b) * (sc.T + 1) * sc.nq * sc.nq;
  const double* dq = bf.dq + size_t(b) * n;
  const double* D = bf.D + size_t(b) * n;
  const double* gm = bf.gm + size_t(b) * n;
  double* dqs = bf.tmp1 + size_t(b) * n;
  double* Hdq = bf.tmp2 + size_t(b) * n;
  for (int e = tid; e < n; e += nt) dqs[e] = sc.scaling ? (1.0 / D[e]) * dq[e] : dq[e];
  __syncthreads();
  penta_matvec(sc, bf.SA + hb, bf.SB + hb, bf.SC + hb, dqs, Hdq, tid, nt);
  __syncthreads();
  double ht = 0.0, gt = 0.0, hl = 0.0, h2 = 0.0;
  for (int e = tid; e < n; e += nt) ht += dqs[e] * Hdq[e], gt += gm[e] * dqs[e];
  if (sc.eq)
    for (int e = tid; e < nh; e += nt) hl += bf.sc.h[size_t(b) * nh + e] * bf.lambda[size_t(b) * nh + e];
  for (int e = tid; e < nh; e += nt) {
    const double x = bf.st.h[size_t(b) * nh + e];
    h2 += x * x;
  }
  ht = block_sum(ht, red), gt = block_sum(gt, red), hl = block_sum(hl, red), h2 = block_sum(h2, red);
  const double merit_k = bf.merit[b];
  const double merit_kp = bf.sc.cost[b] + hl;
  const double predicted = -gt - 0.5 * ht;
  const double actual = merit_k - merit_kp;
  const double eps = 10 * 2.220446049250313e-16 / sc.dt / sc.dt;
  double rho = (predicted < eps && actual < eps) ? 0.5 : actual / predicted;
  if (tid == 0) {
    ctl->rho = rho;
    ctl->hnorm = sqrt(h2);
    s_accept = (rho > 0.0) ? 1 : 0;
  }
  __syncthreads();
  if (!commit) return;
  const int accept = s_accept;
  const double cost_k = bf.st.cost[b];
  if (tid == 0) {
    const int it = ctl->iters;
    if (it < bf.stats_cap) {
      double* st = bf.stats + (size_t(b) * bf.stats_cap + it) * IDTO_NUM_STATS;
      st[0] = cost_k, st[1] = ctl->Delta, st[2] = ctl->q_norm, st[3] = ctl->dq_norm, st[4] = ctl->dqH_norm;
      st[5] = rho, st[6] = ctl->gnorm, st[7] = ctl->dL_dq, st[8] = ctl->hnorm, st[9] = merit_k;
    }
    ctl->iters = it + 1;
  }
  if (accept) {
    // state.AddToQ(dq): the scratch trajectory already holds q+dq and everything derived from it;
    // the reference recomputes the same numbers from scratch (cc:1989-1990 TODO) — we adopt them.
    const int T = sc.T, nq = sc.nq, nv = sc.nv;
    for (int e = tid; e < (T + 1) * nq; e += nt) bf.st.q[size_t(b) * (T + 1) * nq + e] = bf.sc.q[size_t(b) * (T + 1) * nq + e];
    for (int e = tid; e < (T + 1) * nv; e += nt) bf.st.v[size_t(b) * (T + 1) * nv + e] = bf.sc.v[size_t(b) * (T + 1) * nv + e];
    for (int e = tid; e < T * nv; e += nt) {
      bf.st.a[size_t(b) * T * nv + e] = bf.sc.a[size_t(b) * T * nv + e];
      bf.st.tau[size_t(b) * T * nv + e] = bf.sc.tau[size_t(b) * T * nv + e];
    }
    for (int e = tid; e < (T + 1) * nv * nq; e += nt)
      bf.st.Nplus[size_t(b) * (T + 1) * nv * nq + e] = bf.sc.Nplus[size_t(b) * (T + 1) * nv * nq + e];
    for (int e = tid; e < nh; e += nt) bf.st.h[size_t(b) * nh + e] = bf.sc.h[size_t(b) * nh + e];
  }
  // Convergence (cc:2601-2611) needs EvalMeritFunctionGradient of the NEW state, i.e. the
  // derivative pipeline of the next iteration: it is marked pending here and evaluated by
  // k_conv_check right after that pipeline (before anything of the next iteration is recorded).
  __syncthreads();
  if (tid == 0) {
    if (accept) {
      bf.st.cost[b] = bf.sc.cost[b];
      ctl->derivs_dirty = 1;
      if (sc.check_convergence) ctl->pending = 1;
    } else {
      ctl->derivs_dirty = 0;
    }
    double Delta = ctl->Delta;
    if (rho < 0.25)
      Delta *= 0.25;
    else if (rho > 0.75 && ctl->tr_active)
      Delta = fmin(2 * Delta, sc.Delta_max);
    ctl->Delta = Delta;
  }
}

// VerifyConvergenceCriteria (cc:2653-2689) for the step accepted in the previous iteration.
__global__ void __launch_bounds__(256) k_conv_check(SolverConsts sc, SolverBufs bf) {
  __shared__ double red[32];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, n = sc.n;
  ProbCtl* ctl = bf.ctl + b;
  if (!ctl->active || !ctl->pending) return;
  const double* gm = bf.gm + size_t(b) * n;
  const double* dq = bf.dq + size_t(b) * n;
  const double* q = bf.st.q + size_t(b) * n;
  double gdq = 0.0, dq2 = 0.0, q2 = 0.0;
  for (int e = tid; e < n; e += nt) gdq += gm[e] * dq[e], dq2 += dq[e] * dq[e], q2 += q[e] * q[e];
  gdq = block_sum(gdq, red), dq2 = block_sum(dq2, red), q2 = block_sum(q2, red);
  if (tid == 0) {
    const double cost = bf.st.cost[b];
    int reason = 0;
    if (fabs(ctl->prev_cost - cost) < sc.tol[1] + sc.tol[0] * cost) reason |= 1;
    if (fabs(gdq) < sc.tol[3] + sc.tol[2] * cost) reason |= 2;
    if (sqrt(dq2) < sc.tol[5] + sc.tol[4] * sqrt(q2)) reason |= 4;
    ctl->prev_cost = cost;
    ctl->reason = reason;
    ctl->pending = 0;
    if (reason != 0) ctl->active = 0;
  }
}

__global__ void k_clear_dirty(SolverConsts sc, SolverBufs bf) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < sc.B) bf.ctl[b].derivs_dirty = 0;
}

void launch_conv_check(const SolverConsts& sc, const SolverBufs& bf, cudaStream_t stream) {
  k_conv_check<<<sc.B, 256, 0, stream>>>(sc, bf);
  g_launch_counter += 1;
}
void launch_clear_dirty(const SolverConsts& sc, const SolverBufs& bf, cudaStream_t stream) {
  k_clear_dirty<<<(sc.B + 127) / 128, 128, 0, stream>>>(sc, bf);
  g_launch_counter += 1;
}

void launch_dogleg(const SolverConsts& sc, const SolverBufs& bf, cudaStream_t stream) {
  g_launch_counter += 2;
  k_dogleg_pre<<<sc.B, 256, 0, stream>>>(sc, bf);
  k_dogleg_post<<<sc.B, 256, 0, stream>>>(sc, bf);
}

void launch_trust_update(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool commit,
                         cudaStream_t stream) {
  (void)dm;
  g_launch_counter += 1;
  k_trust_update<<<sc.B, 256, 0, stream>>>(sc, bf, commit ? 1 : 0);
}

}  // namespace idto
