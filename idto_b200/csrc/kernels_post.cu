// Scalar tail of a trust-region iteration, one CTA per problem (after the inverse dynamics of the scratch trajectory):
// rho (cc:1979-2035), acceptance (cc:2550-2553), stats (cc:2586-2598), the commit of an accepted step, the
// trust-region update (cc:2613-2622).
//
// The model terms of the trust ratio, s.H~s and gm.s (cc:2008-2017), come from k_dogleg_post (kernels_tr.cu), which
// forms them from dot products it has anyway: every dogleg step is a combination of the Cauchy direction and the
// Gauss-Newton step x = -H~^-1 gm of the KKT sweep, and H~ x = -gm holds to the sweep's residual (1e-9 relative,
// tests/test_gpu_parity.py), so no second pass over H~ is needed.  k_trust_update (kernels_tr.cu), which multiplies
// H~ s out explicitly like the reference, remains behind IDTO_TRUST_MATVEC=1; tests/test_more_options_gpu.py
// compares the two.  Replaces a 2624-CTA row-parallel kernel with a last-CTA election (41 us) by 64 small CTAs.
#include <cstdlib>

#include "reduce.cuh"
#include "solver.h"

namespace idto {

// Trust ratio (cc:1979-2035), acceptance (cc:2550-2553), stats (cc:2586-2598), commit, Delta update (cc:2613-2622).
// `commit` = 0 only evaluates rho.  One CTA per problem; s.H~s and gm.s come from k_post.
__global__ void __launch_bounds__(256) k_trust_final(SolverConsts sc, SolverBufs bf, int commit, int near_stride_) {
  __shared__ double red[2 * 32];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, nh = sc.nh;
  ProbCtl* ctl = bf.ctl + b;
  if (!ctl->active) return;
  double hl = 0.0, h2 = 0.0;
  for (int e = tid; e < nh; e += nt) {
    const size_t ge = size_t(b) * nh + e;
    if (sc.eq) hl += bf.sc.h[ge] * bf.lambda[ge];
    h2 += bf.st.h[ge] * bf.st.h[ge];
  }
  {
    double v[2] = {hl, h2};
    block_sum_n(v, red);
    hl = v[0], h2 = v[1];
  }
  const double ht = ctl->ht, gt = ctl->gt;  // k_dogleg_post
  const double merit_k = bf.merit[b];
  const double merit_kp = bf.sc.cost[b] + hl;
  const double predicted = -gt - 0.5 * ht;
  const double actual = merit_k - merit_kp;
  const double eps = 10 * 2.220446049250313e-16 / sc.dt / sc.dt;
  const double rho = (predicted < eps && actual < eps) ? 0.5 : actual / predicted;
  const int accept = rho > 0.0 ? 1 : 0;
  if (tid == 0) {
    ctl->rho = rho;
    ctl->hnorm = sqrt(h2);
  }
  if (!commit) return;
  const double cost_k = bf.st.cost[b];
  if (tid == 0) {
    const int it = ctl->iters;
    if (it < bf.stats_cap) {
      double* st = bf.stats + (size_t(b) * bf.stats_cap + it) * IDTO_NUM_STATS;
      st[0] = cost_k, st[1] = ctl->Delta, st[2] = ctl->q_norm, st[3] = ctl->dq_norm, st[4] = ctl->dqH_norm;
      st[5] = rho, st[6] = ctl->gnorm, st[7] = ctl->dL_dq, st[8] = sqrt(h2), st[9] = merit_k;
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
      for (int e = tid; e < T * near_stride_; e += nt)
        bf.st.near[size_t(b) * T * near_stride_ + e] = bf.sc.near[size_t(b) * T * near_stride_ + e];
  }
  // Convergence (cc:2601-2611) needs EvalMeritFunctionGradient of the NEW state, i.e. the derivative pipeline of
  // the next iteration: it is marked pending here and evaluated by k_conv_check right after that pipeline.
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

bool trust_final_enabled() {
  static const bool off = std::getenv("IDTO_TRUST_MATVEC") != nullptr;  // the explicit H~ s mat-vec of kernels_tr.cu
  return !off;
}

void launch_trust_final(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool commit,
                        cudaStream_t stream) {
  g_launch_counter += 1;
  k_trust_final<<<sc.B, 256, 0, stream>>>(sc, bf, commit ? 1 : 0, near_stride(dm.nact));
}

}  // namespace idto
