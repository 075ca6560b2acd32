// Trust-region scalars kept on the device: dogleg point (cc:2108-2202), trust ratio (cc:1979-2035),
// step acceptance, Delta update, convergence checks and stats (cc:2494-2625).  One CTA per problem; the
// host never reads a scalar back during a solve.
#include "reduce.cuh"
#include "solver.h"

namespace idto {

namespace {

constexpr int kRowThreads = 128;

// Block row i of y = H~ x for the symmetric block penta-diagonal matrix given by its lower bands
// (PentaDiagonalMatrix::MultiplyBy, penta_diagonal_matrix.cc:181-207; D_i = B_{i+1}^T, E_i = A_{i+2}^T).
// The five blocks of the row are first staged in shared memory `hs` [5][k*k] by all threads (independent,
// coalesced loads: one L2 round trip instead of one per multiply-add chain step); xs[5][k] holds the blocks
// i-2..i+2 of x (zeros outside the matrix); thread-task (term, row) computes one k-long dot product into
// terms[5][k]; the caller sums the five terms in the reference's order.
__device__ __forceinline__ void stage_row_blocks(const SolverConsts& sc, const double* SA, const double* SB,
                                                 const double* SC, int i, double* hs, int tid, int nt) {
  const int k = sc.nq, kk = k * k, nblk = sc.T + 1;
  const double* src[5] = {SC + size_t(i) * kk, i >= 1 ? SB + size_t(i) * kk : nullptr,
                          i >= 2 ? SA + size_t(i) * kk : nullptr, i < nblk - 1 ? SB + size_t(i + 1) * kk : nullptr,
                          i < nblk - 2 ? SA + size_t(i + 2) * kk : nullptr};
#pragma unroll
  for (int term = 0; term < 5; ++term) {
    const double* p = src[term];
#pragma unroll 8
    for (int e = tid; e < kk; e += nt) hs[term * kk + e] = p ? p[e] : 0.0;
  }
}
__device__ __forceinline__ void block_row_matvec(const SolverConsts& sc, const double* hs, const double* xs,
                                                 double* terms, int tid, int nt) {
  const int k = sc.nq, kk = k * k;
  for (int task = tid; task < 5 * k; task += nt) {
    const int term = task / k, r = task - term * k;
    const double* Hb = hs + term * kk;
    // terms 0..4 multiply x_i, x_{i-1}, x_{i-2}, x_{i+1}, x_{i+2}; the last two use the transposed block
    const double* x = xs + (term == 0 ? 2 : (term == 1 ? 1 : (term == 2 ? 0 : (term == 3 ? 3 : 4)))) * k;
    double acc = 0.0;
    if (term < 3) {
      for (int c = 0; c < k; ++c) acc += Hb[c * k + r] * x[c];
    } else {
      for (int c = 0; c < k; ++c) acc += Hb[r * k + c] * x[c];
    }
    terms[task] = acc;
  }
}

// Sum of the per-block-row partials part[j][q], j < nblk, by warp 0 (all 32 lanes call): lane j takes rows
// j, j+32, ... (independent L2 loads), then a fixed shuffle tree — deterministic, and one L2 round trip
// instead of nblk dependent ones (the serial sum was a 12 us tail on every launch).
__device__ __forceinline__ double partials_sum(const double* part, int nblk, int q, int lane) {
  double x = 0.0;
  for (int j = lane; j < nblk; j += 32) x += __ldcg(part + 4 * j + q);
  return warp_sum(x);
}

// Deterministic sum over the first k <= 32 lanes of warp 0 (fixed shuffle tree).
__device__ __forceinline__ double lanes_sum(double x) { return warp_sum(x); }

// "Last CTA of the problem" election (threadfence reduction): returns true in every thread of the CTA that
// arrives last among the nblk block-row CTAs of problem b; that CTA then owns the per-problem epilogue.
__device__ __forceinline__ bool last_block_of_problem(int* cnt, int nblk) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(cnt, 1);
    s_last = done == nblk - 1;
    if (s_last) *cnt = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

}  // namespace

// gm = g~ + J~^T lambda (cc:1442), merit = L + h.lambda (cc:1418) and the dogleg scalars gHg = gm.H~ gm,
// g.g (cc:2157): one CTA per (problem, block row) — the per-problem version of this mat-vec ran 46 us on 64
// SMs.  Every CTA rebuilds the five blocks of gm its row touches (a few hundred multiply-adds), writes
// its own block of gm and its partial sums; the last CTA of a problem adds them up in block order.
__global__ void __launch_bounds__(kRowThreads) k_gm_matvec(SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) double dyn[];
  __shared__ double xs[5 * 32], terms[5 * 32];
  const int nblk = sc.T + 1, b = blockIdx.x / nblk, i = blockIdx.x % nblk;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  const int n = sc.n, nh = sc.nh, nu = sc.nu, k = sc.nq, T = sc.T, tid = threadIdx.x, nt = blockDim.x;
  const size_t hb = size_t(b) * nblk * k * k;
  const double* gs = bf.gs + size_t(b) * n;
  const double* lam = bf.lambda + size_t(b) * nh;
  const bool eq = sc.eq && nh > 0;
  double* hs = dyn;  // [5][k*k] blocks of the row
  stage_row_blocks(sc, bf.SA + hb, bf.SB + hb, bf.SC + hb, i, hs, tid, nt);
  for (int e = tid; e < 5 * k; e += nt) {
    const int s = i - 2 + e / k, c = e % k;
    double val = 0.0;
    if (s >= 0 && s < nblk) {
      double acc = 0.0;
      if (eq) {
        if (s >= 1)  // rows (s-1, u): Jp[s-1]
#pragma unroll 6
          for (int u = 0; u < nu; ++u) acc += bf.Jp[((size_t(b) * T + (s - 1)) * nu + u) * k + c] * lam[(s - 1) * nu + u];
        if (s < T)  // rows (s, u): Jt[s]
#pragma unroll 6
          for (int u = 0; u < nu; ++u) acc += bf.Jt[((size_t(b) * T + s) * nu + u) * k + c] * lam[s * nu + u];
        if (s + 1 < T)  // rows (s+1, u): Jm[s+1]
#pragma unroll 6
          for (int u = 0; u < nu; ++u) acc += bf.Jm[((size_t(b) * T + (s + 1)) * nu + u) * k + c] * lam[(s + 1) * nu + u];
      }
      val = eq ? gs[s * k + c] + acc : gs[s * k + c];
      if (s == i) bf.gm[size_t(b) * n + s * k + c] = val;
    }
    xs[e] = val;
  }
  __syncthreads();
  block_row_matvec(sc, hs, xs, terms, tid, nt);
  __syncthreads();
  if (tid < 32) {
    double gHg = 0.0, gg = 0.0, hl = 0.0;
    if (tid < k) {
      const double y = (((terms[tid] + terms[k + tid]) + terms[2 * k + tid]) + terms[3 * k + tid]) + terms[4 * k + tid];
      const double x = xs[2 * k + tid];
      gHg = x * y, gg = x * x;
    }
    if (eq && i < T && tid < nu) hl = bf.st.h[size_t(b) * nh + i * nu + tid] * lam[i * nu + tid];
    gHg = lanes_sum(gHg), gg = lanes_sum(gg), hl = lanes_sum(hl);
    if (tid == 0) {
      double* pp = bf.part + (size_t(b) * nblk + i) * 4;
      pp[0] = gHg, pp[1] = gg, pp[2] = hl;
    }
  }
  if (last_block_of_problem(bf.cnt + b, nblk) && tid < 32) {
    const double* pp = bf.part + size_t(b) * nblk * 4;
    const double gHg = partials_sum(pp, nblk, 0, tid), gg = partials_sum(pp, nblk, 1, tid);
    const double hl = partials_sum(pp, nblk, 2, tid);
    if (tid == 0) {
      bf.red[b * 8 + 0] = gHg, bf.red[b * 8 + 1] = gg;
      bf.merit[b] = eq ? bf.st.cost[b] + hl : bf.st.cost[b];
    }
  }
}

// Dogleg part 2: pH = H~^-1(-gm/Delta) = x/Delta with x from k_kkt_solve (cc:2139-2140; the solve is
// linear in the right-hand side, so x is computed once per derivative update and reused by rejected
// steps); branch logic, dq, dqH, logging scalars, and the
// scratch trajectory q + dq (cc:1991-1993).
__global__ void __launch_bounds__(256) k_dogleg_post(SolverConsts sc, SolverBufs bf) {
  __shared__ double red[5 * 32];
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
  double* dqs = bf.tmp1 + size_t(b) * n;
  // pU = -(g.g / gHg) g / Delta  (cc:2157)
  double pU2 = 0.0, pH2 = 0.0, a = 0.0, bq = 0.0, xg = 0.0;
#pragma unroll 4
  for (int e = tid; e < n; e += nt) {
    const double pu = -(gg / gHg) * gm[e] / Delta, ph = pH[e] / Delta;
    pU2 += pu * pu, pH2 += ph * ph;
    const double d = ph - pu;
    a += d * d, bq += pu * d;
    xg += ph * gm[e];
  }
  {
    double v[5] = {pU2, pH2, a, bq, xg};  // (xg = gm . pH)
    block_sum_n(v, red);
    pU2 = v[0], pH2 = v[1], a = v[2], bq = v[3], xg = v[4];
  }
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
#pragma unroll 4
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
    dqs[e] = sc.scaling ? (1.0 / D[e]) * x : x;  // s = D^-1 dq for the trust ratio (cc:2008)
    dq2 += x * x, dqH2 += xh * xh;
    gdq += sc.scaling ? gm[e] * ((1.0 / D[e]) * x) : gm[e] * x;  // cc:2518-2523
    q2 += q[e] * q[e];
    qs[e] = q[e] + x;  // scratch_state.set_q(q); AddToQ(dq)
  }
  {
    double v[4] = {dq2, dqH2, gdq, q2};
    block_sum_n(v, red);
    dq2 = v[0], dqH2 = v[1], gdq = v[2], q2 = v[3];
  }
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
    // Model terms of the trust ratio for s = D^-1 dq = Delta y (cc:2008-2017): y is a combination of pU and pH, and
    //   pU^T H~ pU = pU^T H~ pH = g.g^2 / (gHg Delta^2)   (pU = -(g.g / gHg) gm / Delta; H~ pH = -gm / Delta),
    //   pH^T H~ pH = -(gm . pH) / Delta,
    // H~ pH = -gm / Delta being what the KKT sweep solved (residual 1e-9 relative: tests/test_gpu_parity.py).
    const double cuu = gg * gg / (gHg * Delta * Delta), chh = -xg / Delta;
    double yHy;
    if (branch == 0) {
      yHy = cuu / pU2;
    } else if (branch == 1) {
      yHy = chh;
    } else {
      const double w = 1.0 - s;
      yHy = w * w * cuu + 2.0 * s * w * cuu + s * s * chh;
    }
    ctl->ht = Delta * Delta * yHy, ctl->gt = gdq;  // gm . s, summed above like the reference does (cc:2518-2523)
  }
}

// Trust ratio (cc:1979-2035), acceptance (cc:2550-2553), stats (cc:2586-2598), convergence
// (cc:2601-2611, 2653-2689) and the Delta update (cc:2613-2622).  `commit`=0 only evaluates rho.
// One CTA per (problem, block row) computes its rows of H~ s, s = D^-1 dq, and the partial sums of
// s.H~s, gm.s, h(q+dq).lambda and |h|^2; the last CTA of a problem owns the scalar logic and the commit.
__global__ void __launch_bounds__(kRowThreads) k_trust_update(SolverConsts sc, SolverBufs bf, int commit, int kNearStride) {
  __shared__ double xs[5 * 32], terms[5 * 32];
  __shared__ double red[32];
  __shared__ int s_accept;
  const int nblk = sc.T + 1, b = blockIdx.x / nblk, i = blockIdx.x % nblk;
  const int tid = threadIdx.x, nt = blockDim.x, n = sc.n, nh = sc.nh, k = sc.nq, nu = sc.nu;
  ProbCtl* ctl = bf.ctl + b;
  if (!ctl->active) return;
  const size_t hb = size_t(b) * nblk * k * k;
  const double* dqs = bf.tmp1 + size_t(b) * n;
  const double* gm = bf.gm + size_t(b) * n;
  extern __shared__ __align__(16) double dyn[];
  stage_row_blocks(sc, bf.SA + hb, bf.SB + hb, bf.SC + hb, i, dyn, tid, nt);
  for (int e = tid; e < 5 * k; e += nt) {
    const int s = i - 2 + e / k;
    xs[e] = (s >= 0 && s < nblk) ? dqs[s * k + e % k] : 0.0;
  }
  __syncthreads();
  block_row_matvec(sc, dyn, xs, terms, tid, nt);
  __syncthreads();
  if (tid < 32) {
    double ht = 0.0, gt = 0.0, hl = 0.0, h2 = 0.0;
    if (tid < k) {
      const double y = (((terms[tid] + terms[k + tid]) + terms[2 * k + tid]) + terms[3 * k + tid]) + terms[4 * k + tid];
      const double x = xs[2 * k + tid];
      ht = x * y, gt = gm[i * k + tid] * x;
    }
    if (i < sc.T && tid < nu) {
      const size_t e = size_t(b) * nh + i * nu + tid;
      if (sc.eq) hl = bf.sc.h[e] * bf.lambda[e];
      h2 = bf.st.h[e] * bf.st.h[e];
    }
    ht = lanes_sum(ht), gt = lanes_sum(gt), hl = lanes_sum(hl), h2 = lanes_sum(h2);
    if (tid == 0) {
      double* pp = bf.part + (size_t(b) * nblk + i) * 4;
      pp[0] = ht, pp[1] = gt, pp[2] = hl, pp[3] = h2;
    }
  }
  if (!last_block_of_problem(bf.cnt + b, nblk)) return;
  if (tid < 32) {
    const double* pp = bf.part + size_t(b) * nblk * 4;
    for (int qq = 0; qq < 4; ++qq) {
      const double x = partials_sum(pp, nblk, qq, tid);
      if (tid == 0) red[qq] = x;
    }
  }
  __syncthreads();
  const double ht = red[0], gt = red[1], hl = red[2], h2 = red[3];
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
#pragma unroll 4
    for (int e = tid; e < (T + 1) * nq; e += nt) bf.st.q[size_t(b) * (T + 1) * nq + e] = bf.sc.q[size_t(b) * (T + 1) * nq + e];
#pragma unroll 4
    for (int e = tid; e < (T + 1) * nv; e += nt) bf.st.v[size_t(b) * (T + 1) * nv + e] = bf.sc.v[size_t(b) * (T + 1) * nv + e];
#pragma unroll 4
    for (int e = tid; e < T * nv; e += nt) {
      bf.st.a[size_t(b) * T * nv + e] = bf.sc.a[size_t(b) * T * nv + e];
      bf.st.tau[size_t(b) * T * nv + e] = bf.sc.tau[size_t(b) * T * nv + e];
    }
    // N+ is a constant pattern except for the 3x4 block of every quaternion joint (k_traj): rows v0..v0+2,
    // columns q0..q0+3, with v0 = q0 - (number of quaternion joints before it)
    for (int e = tid; e < (T + 1) * sc.nquat * 12; e += nt) {
      const int t = e / (sc.nquat * 12), rem = e % (sc.nquat * 12), j = rem / 12, c = (rem % 12) / 3, r = rem % 3;
      const int q0 = sc.quat_starts[j];
      int v0 = q0;
      for (int jj = 0; jj < sc.nquat; ++jj) v0 -= sc.quat_starts[jj] < q0;
      const size_t o = (size_t(b) * (T + 1) + t) * nv * nq + size_t(q0 + c) * nv + v0 + r;
      bf.st.Nplus[o] = bf.sc.Nplus[o];
    }
    for (int e = tid; e < nh; e += nt) bf.st.h[size_t(b) * nh + e] = bf.sc.h[size_t(b) * nh + e];
    if (bf.st.near)  // near lists of the adopted poses (pruned contact models)
      for (int e = tid; e < T * kNearStride; e += nt)
        bf.st.near[size_t(b) * T * kNearStride + e] = bf.sc.near[size_t(b) * T * kNearStride + e];
  }
  // Convergence (cc:2601-2611) needs EvalMeritFunctionGradient of the NEW state, i.e. the
  // derivative pipeline of the next iteration: it is marked pending here and evaluated by
  // k_conv_check right after that pipeline (before anything of the next iteration is recorded).
  __syncthreads();
  if (tid == 0) {
    if (accept) {
      bf.st.cost[b] = bf.sc.cost[b];
      ctl->derivs_dirty = 1;
      ctl->stash_sel ^= 1;  // the scratch evaluation's per-body records are now those of the state
      if (sc.check_convergence) ctl->pending = 1;
    } else {
      ctl->derivs_dirty = 0;
    }
    double Delta = ctl->Delta;
    ctl->Delta_prev = Delta;
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
    if (reason != 0) {
      ctl->active = 0;
      // the reference leaves its loop BEFORE the trust-region update of a converged step (cc:2608-2622); that
      // update already ran here (the check needs the next derivative pipeline): undo it
      ctl->Delta = ctl->Delta_prev;
    }
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

void launch_gm_matvec(const SolverConsts& sc, const SolverBufs& bf, bool force, cudaStream_t stream) {
  g_launch_counter += 1;
  const int smem = 5 * sc.nq * sc.nq * 8;
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_gm_matvec, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_trust_update, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  }
  k_gm_matvec<<<sc.B*(sc.T + 1), kRowThreads, smem, stream>>>(sc, bf, force ? 1 : 0);
}

void launch_dogleg(const SolverConsts& sc, const SolverBufs& bf, cudaStream_t stream) {
  g_launch_counter += 1;  // Hg, gHg, g.g come with gm (k_gm_matvec, launched with the KKT sweep)
  k_dogleg_post<<<sc.B, 256, 0, stream>>>(sc, bf);
}

void launch_trust_update(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool commit,
                         cudaStream_t stream) {
  g_launch_counter += 1;
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_trust_update, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  }
  k_trust_update<<<sc.B*(sc.T + 1), kRowThreads, 5 * sc.nq * sc.nq * 8, stream>>>(sc, bf, commit ? 1 : 0,
                                                                                near_stride(dm.nact));
}

}  // namespace idto
