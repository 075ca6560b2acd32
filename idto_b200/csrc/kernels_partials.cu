// ID-partials kernel family: d tau_{t-1,t,t+1} / d q_t by finite differences.
//
// Reference: CalcInverseDynamicsPartialsFiniteDiff (optimizer/trajectory_optimizer.cc:426-563) and
// CalcInverseDynamicsPartialsCentralDiff / ...WrtQtCentralDiff (cc:565-885, incl. the 4th-order
// stencil).  One G-lane group per (problem b, time step t, position index i); the group runs the
// 2 (forward), 6 (central) or 12 (4th order) inverse-dynamics evaluations of that column and writes
// column i of dtau_dqp[t-1], dtau_dqt[t], dtau_dqm[t+1].  Evaluations that share q (tau_t and
// tau_{t+1} are evaluated at the unperturbed q_{t+1}, q_{t+2}: cc:790-798, 817-824) share one
// position phase.  Perturbations propagate through the FIXED N+_t, N+_{t+1} columns (cc:516-520).
#include "dynamics.cuh"

namespace idto {

namespace {

// Column i of N+(q) restricted to the owner body's velocity slots (zero elsewhere).
__device__ __forceinline__ void nplus_col(int jtype, int local, const double* qb, double* n6) {
#pragma unroll
  for (int j = 0; j < 6; ++j) n6[j] = 0.0;
  if (jtype == IDTO_JOINT_QUAT_FLOATING) {
    if (local < 4) {
      const V3 c = quat_nplus_col(qb, local);
      n6[0] = c.x, n6[1] = c.y, n6[2] = c.z;
    } else {
#pragma unroll
      for (int j = 3; j < 6; ++j)
        if (j == local - 1) n6[j] = 1.0;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j == local) n6[j] = 1.0;
  }
}

template <int N>
__device__ __forceinline__ void load_seg(const double* __restrict__ src, int n, double* dst) {
#pragma unroll
  for (int j = 0; j < N; ++j)
    if (j < n) dst[j] = src[j];
}

}  // namespace

template <int G, int METHOD>
__global__ void __launch_bounds__(128) k_partials(DevModel dm, SolverConsts sc, SolverBufs bf, int force) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int groups = blockDim.x / G, grp = threadIdx.x / G, k = threadIdx.x % G;
  const int T = sc.T, nq = sc.nq, nv = sc.nv;
  const int item = blockIdx.x * groups + grp;
  const bool in_range = item < sc.B * T * nq;
  const int b = in_range ? item / (T * nq) : 0;
  const int rem = in_range ? item % (T * nq) : 0;
  const int t = rem / nq + 1, i = rem % nq;
  const bool live = in_range && (force || bf.ctl[b].derivs_dirty);
  if (!__syncthreads_or(live ? 1 : 0)) return;  // whole CTA belongs to clean problems

  int* si = reinterpret_cast<int*>(smem);
  double* sd = reinterpret_cast<double*>(smem + dm.itab_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + dm.itab_bytes + dm.dtab_bytes);
  double* gbase = reinterpret_cast<double*>(smem + model_smem_bytes(dm));
  stage_model(dm, si, sd, bar);
  const SModel M = make_smodel(dm, si, sd);
  const GroupSmem S = make_group_smem(dm, gbase + size_t(grp) * group_smem_doubles(dm));

  const bool body = k < M.nb;
  const int owner = M.qowner[i];
  int jt = 0, q0 = 0, v0 = 0, nqb = 0, nvb = 0;
  if (body) jt = M.jtype[k], q0 = M.qs[k], v0 = M.vs[k], nqb = joint_nq(jt), nvb = joint_nv(jt);
  const bool is_owner = body && (k == owner);
  const int local = i - q0;  // index of q_i inside the owner's joint (valid on the owner lane)

  const double* qB = bf.st.q + size_t(b) * (T + 1) * nq;
  const double* vB = bf.st.v + size_t(b) * (T + 1) * nv;
  const double* aB = bf.st.a + size_t(b) * T * nv;

  // step size (cc:504-511 / 709-716), computed on the owner lane and broadcast to the group
  const double eps = 1.4901161193847656e-08;  // sqrt(2^-52)
  double dq = 0.0;
  double nt[6], ntp[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) nt[j] = 0.0, ntp[j] = 0.0;
  double qt_own[7] = {1, 0, 0, 0, 0, 0, 0};
  if (body) load_seg<7>(qB + size_t(t) * nq + q0, nqb, qt_own);
  if (is_owner) {
    const double qi = qB[size_t(t) * nq + i];
    dq = eps * fmax(1.0, fabs(qi));
    const double temp = __dadd_rn(qi, dq);
    dq = __dadd_rn(temp, -qi);
    nplus_col(jt, local, qt_own, nt);
    if (t < T) {
      double qtp[7] = {1, 0, 0, 0, 0, 0, 0};
      load_seg<7>(qB + size_t(t + 1) * nq + q0, nqb, qtp);
      nplus_col(jt, local, qtp, ntp);
    }
  }
  dq = __shfl_sync(0xffffffffu, dq, (threadIdx.x & 31) / G * G + owner);
  const double dv = dq / sc.dt, da = dv / sc.dt;

  constexpr int NK = METHOD == IDTO_GRAD_CENTRAL4 ? 4 : (METHOD == IDTO_GRAD_CENTRAL ? 2 : 1);
  const double mult[4] = {1.0, -1.0, 2.0, -2.0};
  LaneKin L;
  double qb[7], vb[6], ab[6], tk[NK][6], v_un[6], a_un[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) v_un[j] = 0.0, a_un[j] = 0.0;

  auto finish = [&](double* __restrict__ dst_block, const double* __restrict__ tau_base) {
    // dst_block: start of the nv x nq block; writes column i rows [v0, v0+nvb)
    if (!(live && body)) return;
    double* dst = dst_block + size_t(i) * nv + v0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      if (j < nvb) {
        double val;
        if (METHOD == IDTO_GRAD_FORWARD) {
          val = (tk[0][j] - tau_base[v0 + j]) / dq;  // cc:531, 539
        } else if (METHOD == IDTO_GRAD_CENTRAL) {
          val = 0.5 * (tk[0][j] - tk[1 % NK][j]) / dq;  // cc:785
        } else {
          val = 2.0 / 3.0 * (tk[0][j] - tk[1 % NK][j]) / dq - 1.0 / 12.0 * (tk[2 % NK][j] - tk[3 % NK][j]) / dq;
        }
        dst[j] = val;
      }
    }
  };

  // ---- tau[t-1] = ID(q_t^e, v_t^e, a_{t-1}^e)   (cc:526-531, 763-787) -------------------------
  if (body) {
    load_seg<6>(vB + size_t(t) * nv + v0, nvb, v_un);
    load_seg<6>(aB + size_t(t - 1) * nv + v0, nvb, a_un);
  }
#pragma unroll
  for (int kk = 0; kk < NK; ++kk) {
    const double m = mult[kk];
#pragma unroll
    for (int j = 0; j < 7; ++j) qb[j] = qt_own[j];
    if (is_owner) {
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (j == local) qb[j] += m * dq;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) vb[j] = v_un[j] + m * dv * nt[j], ab[j] = a_un[j] + m * da * nt[j];
    PositionPhase<G>(M, S, sc, k, qb, &L);
    VelocityPhase<G>(M, S, sc, k, L, vb, ab, true, tk[kk]);
  }
  finish(bf.dqp + (size_t(b) * T + (t - 1)) * nv * nq, bf.st.tau + (size_t(b) * T + (t - 1)) * nv);

  // ---- tau[t] = ID(q_{t+1}, v_{t+1}^e, a_t^e)   (cc:533-540, 788-814) ---------------------------
  if (t < T) {  // uniform across the group
    if (body) {
      load_seg<7>(qB + size_t(t + 1) * nq + q0, nqb, qb);
      load_seg<6>(vB + size_t(t + 1) * nv + v0, nvb, v_un);
      load_seg<6>(aB + size_t(t) * nv + v0, nvb, a_un);
    }
  }
  // NB: groups of one warp may disagree on (t < T); every lane still executes the phases (they
  // contain warp-level barriers) and `finish` is predicated instead.
  {
    PositionPhase<G>(M, S, sc, k, qb, &L);
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
      const double m = mult[kk];
#pragma unroll
      for (int j = 0; j < 6; ++j) vb[j] = v_un[j] - m * dv * ntp[j], ab[j] = a_un[j] - m * da * (ntp[j] + nt[j]);
      VelocityPhase<G>(M, S, sc, k, L, vb, ab, true, tk[kk]);
    }
    if (t < T) finish(bf.dqt + (size_t(b) * T + t) * nv * nq, bf.st.tau + (size_t(b) * T + t) * nv);
  }

  // ---- tau[t+1] = ID(q_{t+2}, v_{t+2}, a_{t+1}^e)   (cc:552-561, 815-839) -----------------------
  if (t < T - 1) {
    if (body) {
      load_seg<7>(qB + size_t(t + 2) * nq + q0, nqb, qb);
      load_seg<6>(vB + size_t(t + 2) * nv + v0, nvb, v_un);
      load_seg<6>(aB + size_t(t + 1) * nv + v0, nvb, a_un);
    }
  }
  {
    PositionPhase<G>(M, S, sc, k, qb, &L);
    if (METHOD == IDTO_GRAD_FORWARD) {
      // dtau_dqm[t+1] = M(q_{t+2}) N+_{t+1} / dt^2 (cc:556-561): one bias-free evaluation with a = N+ col.
      VelocityPhase<G>(M, S, sc, k, L, v_un, ntp, false, tk[0]);
      if (t < T - 1 && live && body) {
        double* dst = bf.dqm + (size_t(b) * T + (t + 1)) * nv * nq + size_t(i) * nv + v0;
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (j < nvb) dst[j] = 1 / sc.dt / sc.dt * tk[0][j];
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < NK; ++kk) {
        const double m = mult[kk];
#pragma unroll
        for (int j = 0; j < 6; ++j) ab[j] = a_un[j] + m * da * ntp[j];
        VelocityPhase<G>(M, S, sc, k, L, v_un, ab, true, tk[kk]);
      }
      if (t < T - 1) finish(bf.dqm + (size_t(b) * T + (t + 1)) * nv * nq, nullptr);
    }
  }
}

int partials_smem_bytes(const DevModel& dm, int threads) {
  return model_smem_bytes(dm) + (threads / dm.group) * group_smem_doubles(dm) * 8;
}

template <int G, int METHOD>
static void launch_partials_gm(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                               cudaStream_t stream) {
  const int threads = 128, groups = threads / G;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_partials<G, METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  const int items = sc.B * sc.T * sc.nq;
  const int grid = (items + groups - 1) / groups;
  g_launch_counter += 1;
  k_partials<G, METHOD><<<grid, threads, partials_smem_bytes(dm, threads), stream>>>(dm, sc, bf, force ? 1 : 0);
}

template <int G>
static void launch_partials_g(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                              cudaStream_t stream) {
  switch (sc.method) {
    case IDTO_GRAD_FORWARD: launch_partials_gm<G, IDTO_GRAD_FORWARD>(dm, sc, bf, force, stream); break;
    case IDTO_GRAD_CENTRAL: launch_partials_gm<G, IDTO_GRAD_CENTRAL>(dm, sc, bf, force, stream); break;
    default: launch_partials_gm<G, IDTO_GRAD_CENTRAL4>(dm, sc, bf, force, stream); break;
  }
}

void launch_partials(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                     cudaStream_t stream) {
  switch (dm.group) {
    case 2: launch_partials_g<2>(dm, sc, bf, force, stream); break;
    case 4: launch_partials_g<4>(dm, sc, bf, force, stream); break;
    case 8: launch_partials_g<8>(dm, sc, bf, force, stream); break;
    case 16: launch_partials_g<16>(dm, sc, bf, force, stream); break;
    default: launch_partials_g<32>(dm, sc, bf, force, stream); break;
  }
}

}  // namespace idto
