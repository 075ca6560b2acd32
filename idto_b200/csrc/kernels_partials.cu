// ID-partials kernel family: d tau_{t-1,t,t+1} / d q_t by finite differences.
//
// Reference: CalcInverseDynamicsPartialsFiniteDiff (optimizer/trajectory_optimizer.cc:426-563) and
// CalcInverseDynamicsPartialsCentralDiff / ...WrtQtCentralDiff (cc:565-885, incl. the 4th-order
// stencil).  A CTA owns `slots` (problem b, time step t) pairs; inside a slot one G-lane group per
// position index i runs the inverse-dynamics evaluations of column i and writes column i of
// dtau_dqp[t-1], dtau_dqt[t], dtau_dqm[t+1]:
//   A  tau[t-1] at q_t +- dq e_i (v_t, a_{t-1} perturbed through the FIXED N+_t column, cc:516-520):
//      one position phase + one velocity phase per stencil point;
//   B  tau[t] at the UNPERTURBED q_{t+1} (cc:790-798): the position phase (poses, contact geometry)
//      is computed once per slot and shared by all nq groups; one velocity phase per stencil point;
//   C  tau[t+1] depends on q_t only through a_{t+1}, linearly: d tau_{t+1}/d q_t = M(q_{t+2}) N+_{t+1}/dt^2.
//      The reference's forward-difference path uses exactly this (cc:552-561); its central-difference
//      path differences two full evaluations (cc:815-839), which is the same number plus O(eps/dq)
//      cancellation noise.  All methods use the exact form here: one bias-free velocity phase on the
//      slot-shared pose of q_{t+2}.
#include "dynamics.cuh"

namespace idto {

namespace {

template <int N>
__device__ __forceinline__ void load_seg(const double* __restrict__ src, int n, double* dst) {
#pragma unroll
  for (int j = 0; j < N; ++j)
    if (j < n) dst[j] = src[j];
}

// x[0..5] += coef * (column of N+ restricted to the owner's velocity slots)
__device__ __forceinline__ void add_col(double* x, double coef, bool quatcol, int sl, V3 n3) {
  if (quatcol) {
    x[0] += coef * n3.x, x[1] += coef * n3.y, x[2] += coef * n3.z;
  } else {
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (j == sl) x[j] += coef;
  }
}

}  // namespace

struct PartialsLayout {
  int slots, groups, threads, smem_bytes;
};

static PartialsLayout partials_layout(const DevModel& dm, int nq) {
  PartialsLayout L;
  const int per_slot = nq * dm.group;
  L.slots = per_slot >= 256 ? 1 : 256 / per_slot;
  const int used = L.slots * per_slot;
  L.threads = (used + 31) / 32 * 32;
  L.groups = L.threads / dm.group;
  L.smem_bytes = model_smem_bytes(dm) +
                 8 * (L.slots * 2 * pos_smem_doubles(dm) + L.groups * (pos_smem_doubles(dm) + vel_smem_doubles(dm)));
  return L;
}

#ifndef IDTO_PARTIALS_MINB
#define IDTO_PARTIALS_MINB 2
#endif
template <int G, int METHOD>
__global__ void __launch_bounds__(320, IDTO_PARTIALS_MINB) k_partials(DevModel dm, SolverConsts sc, SolverBufs bf, int slots,
                                                     int force) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int T = sc.T, nq = sc.nq, nv = sc.nv;
  const int ngroups = blockDim.x / G, g = threadIdx.x / G, k = threadIdx.x % G;
  const int slot = g / nq, i = g % nq;
  const int sg = blockIdx.x * slots + slot;
  const bool valid = (slot < slots) && (sg < sc.B * T);
  const int b = valid ? sg / T : 0;
  const int t = valid ? sg % T + 1 : 1;
  const bool live = valid && (force || bf.ctl[b].derivs_dirty);
  if (!__syncthreads_or(live ? 1 : 0)) return;  // whole CTA belongs to clean problems

  int* si = reinterpret_cast<int*>(smem);
  double* sd = reinterpret_cast<double*>(smem + dm.itab_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + dm.itab_bytes + dm.dtab_bytes);
  double* base = reinterpret_cast<double*>(smem + model_smem_bytes(dm));
  stage_model(dm, si, sd, bar);
  const SModel M = make_smodel(dm, si, sd);
  const int pd = pos_smem_doubles(dm), vd = vel_smem_doubles(dm);
  const int sslot = valid ? slot : 0;
  const PosSmem PB = make_pos_smem(dm, base + size_t(sslot) * 2 * pd);       // pose of q_{t+1}, slot-shared
  const PosSmem PC = make_pos_smem(dm, base + size_t(sslot) * 2 * pd + pd);  // pose of q_{t+2}, slot-shared
  double* gbase = base + size_t(slots) * 2 * pd + size_t(g) * (pd + vd);
  const PosSmem PA = make_pos_smem(dm, gbase);
  const VelSmem S = make_vel_smem(dm, gbase + pd);
  (void)ngroups;

  const bool body = k < M.nb;
  const int owner = M.qowner[i];
  int jt = 0, q0 = 0, v0 = 0, nqb = 0, nvb = 0;
  if (body) jt = M.jtype[k], q0 = M.qs[k], v0 = M.vs[k], nqb = joint_nq(jt), nvb = joint_nv(jt);
  const bool is_owner = body && (k == owner);
  const int local = i - q0;  // index of q_i inside the owner's joint (meaningful on the owner lane)
  const bool quatcol = is_owner && jt == IDTO_JOINT_QUAT_FLOATING && local < 4;
  const int sl = (jt == IDTO_JOINT_QUAT_FLOATING) ? local - 1 : local;  // velocity slot of a unit column

  const double* qB = bf.st.q + size_t(b) * (T + 1) * nq;
  const double* vB = bf.st.v + size_t(b) * (T + 1) * nv;
  const double* aB = bf.st.a + size_t(b) * T * nv;

  double qb[7] = {1, 0, 0, 0, 0, 0, 0};
  // ---- phase 0: slot-shared position phases of the unperturbed q_{t+1} (group 0) and q_{t+2} -------
  {
    const int rounds = nq == 1 ? 2 : 1;
    for (int rd = 0; rd < rounds; ++rd) {
      const bool doB = valid && i == 0 && rd == 0 && t < T;
      const bool doC = valid && t < T - 1 && ((nq == 1) ? (rd == 1) : (i == 1));
      if (__any_sync(0xffffffffu, doB || doC)) {
        const int tt = doB ? t + 1 : (doC ? t + 2 : t);
        if (body) load_seg<7>(qB + size_t(tt) * nq + q0, nqb, qb);
        PositionPhase<G>(M, doB ? PB : (doC ? PC : PA), sc, k, qb);
      }
    }
  }

  // ---- step size (cc:504-511 / 709-716) on the owner lane, broadcast to the group ------------------
  const double eps = 1.4901161193847656e-08;  // sqrt(2^-52)
  double dq = 0.0;
  V3 nt3 = {0, 0, 0}, ntp3 = {0, 0, 0};
  double qt_own[7] = {1, 0, 0, 0, 0, 0, 0};
  if (body) load_seg<7>(qB + size_t(t) * nq + q0, nqb, qt_own);
  if (is_owner) {
    const double qi = qB[size_t(t) * nq + i];
    dq = eps * fmax(1.0, fabs(qi));
    const double temp = __dadd_rn(qi, dq);
    dq = __dadd_rn(temp, -qi);
    if (quatcol) {
      nt3 = quat_nplus_col(qt_own, local);
      if (t < T) {
        load_seg<7>(qB + size_t(t + 1) * nq + q0, nqb, qb);
        ntp3 = quat_nplus_col(qb, local);
      }
    }
  }
  dq = __shfl_sync(0xffffffffu, dq, (threadIdx.x & 31) / G * G + owner);
  const double dv = dq / sc.dt, da = dv / sc.dt;

  constexpr int NK = METHOD == IDTO_GRAD_CENTRAL4 ? 4 : (METHOD == IDTO_GRAD_CENTRAL ? 2 : 1);
  double vb[6], ab[6], v_un[6], a_un[6], tau[6], hold[6], d1[6], d2[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) v_un[j] = 0.0, a_un[j] = 0.0, hold[j] = 0.0, d1[j] = 0.0, d2[j] = 0.0;

  // combine the stencil points of one column and write rows [v0, v0+nvb) of column i
  auto emit = [&](double* __restrict__ dst_block, const double* __restrict__ tau_base, bool ok) {
    if (!(ok && live && body)) return;
    double* dst = dst_block + size_t(i) * nv + v0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      if (j < nvb) {
        double val;
        if (METHOD == IDTO_GRAD_FORWARD)
          val = (hold[j] - tau_base[v0 + j]) / dq;  // cc:531, 539
        else if (METHOD == IDTO_GRAD_CENTRAL)
          val = 0.5 * d1[j] / dq;  // cc:785
        else
          val = 2.0 / 3.0 * d1[j] / dq - 1.0 / 12.0 * d2[j] / dq;  // cc:782-783
        dst[j] = val;
      }
    }
  };
  auto stash = [&](int kk) {  // kk: 0 = +dq, 1 = -dq, 2 = +2dq, 3 = -2dq
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      if (kk & 1) {
        const double df = hold[j] - tau[j];
        if (kk == 1) d1[j] = df; else d2[j] = df;
      } else {
        hold[j] = tau[j];
      }
    }
  };

  // ---- A: tau[t-1] = ID(q_t^e, v_t^e, a_{t-1}^e)   (cc:526-531, 763-787) -------------------------
  if (body) {
    load_seg<6>(vB + size_t(t) * nv + v0, nvb, v_un);
    load_seg<6>(aB + size_t(t - 1) * nv + v0, nvb, a_un);
  }
#pragma unroll 1
  for (int kk = 0; kk < NK; ++kk) {
    const double m = ((kk & 1) ? -1.0 : 1.0) * ((kk >= 2) ? 2.0 : 1.0);
#pragma unroll
    for (int j = 0; j < 7; ++j) qb[j] = qt_own[j];
#pragma unroll
    for (int j = 0; j < 6; ++j) vb[j] = v_un[j], ab[j] = a_un[j];
    if (is_owner) {
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (j == local) qb[j] += m * dq;
      add_col(vb, m * dv, quatcol, sl, nt3);
      add_col(ab, m * da, quatcol, sl, nt3);
    }
    PositionPhase<G>(M, PA, sc, k, qb);
    VelocityPhase<G>(M, PA, S, sc, k, vb, ab, true, tau);
    stash(kk);
  }
  emit(bf.dqp + (size_t(b) * T + (t - 1)) * nv * nq, bf.st.tau + (size_t(b) * T + (t - 1)) * nv, true);

  __syncthreads();  // slot-shared poses of q_{t+1}, q_{t+2} are complete

  // ---- B: tau[t] = ID(q_{t+1}, v_{t+1}^e, a_t^e)   (cc:533-540, 788-814) ---------------------------
  if (body && t < T) {
    load_seg<6>(vB + size_t(t + 1) * nv + v0, nvb, v_un);
    load_seg<6>(aB + size_t(t) * nv + v0, nvb, a_un);
  }
  // Groups of one warp may disagree on (t < T); every lane executes the phases (warp-level barriers
  // inside) on whatever pose the slot holds and `emit` is predicated instead.
#pragma unroll 1
  for (int kk = 0; kk < NK; ++kk) {
    const double m = ((kk & 1) ? -1.0 : 1.0) * ((kk >= 2) ? 2.0 : 1.0);
#pragma unroll
    for (int j = 0; j < 6; ++j) vb[j] = v_un[j], ab[j] = a_un[j];
    if (is_owner) {
      add_col(vb, -(m * dv), quatcol, sl, ntp3);                  // v[t+1] -= m dv N+_{t+1}[:,i]
      if (quatcol) {
        ab[0] -= m * da * (ntp3.x + nt3.x), ab[1] -= m * da * (ntp3.y + nt3.y), ab[2] -= m * da * (ntp3.z + nt3.z);
      } else {
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (j == sl) ab[j] -= m * da * (1.0 + 1.0);             // a[t] -= m da (N+_{t+1} + N+_t)[:,i]
      }
    }
    VelocityPhase<G>(M, (t < T) ? PB : PA, S, sc, k, vb, ab, true, tau);
    stash(kk);
  }
  emit(bf.dqt + (size_t(b) * T + t) * nv * nq, bf.st.tau + (size_t(b) * T + t) * nv, t < T);

  // ---- C: dtau_dqm[t+1] = M(q_{t+2}) N+_{t+1} / dt^2   (cc:552-561) ---------------------------------
#pragma unroll
  for (int j = 0; j < 6; ++j) ab[j] = 0.0, vb[j] = 0.0;
  if (is_owner) add_col(ab, 1.0, quatcol, sl, ntp3);
  VelocityPhase<G>(M, (t < T - 1) ? PC : PA, S, sc, k, vb, ab, false, tau);
  if (t < T - 1 && live && body) {
    double* dst = bf.dqm + (size_t(b) * T + (t + 1)) * nv * nq + size_t(i) * nv + v0;
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (j < nvb) dst[j] = 1 / sc.dt / sc.dt * tau[j];
  }
}

int partials_smem_bytes(const DevModel& dm, int nq) { return partials_layout(dm, nq).smem_bytes; }

template <int G, int METHOD>
static void launch_partials_gm(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                               cudaStream_t stream) {
  const PartialsLayout L = partials_layout(dm, sc.nq);
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_partials<G, METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  }
  const int grid = (sc.B * sc.T + L.slots - 1) / L.slots;
  g_launch_counter += 1;
  k_partials<G, METHOD><<<grid, L.threads, L.smem_bytes, stream>>>(dm, sc, bf, L.slots, force ? 1 : 0);
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
  if (use_chain_kernels(dm)) {
    launch_partials_chain(dm, sc, bf, force, stream);
    return;
  }
  switch (dm.group) {
    case 2: launch_partials_g<2>(dm, sc, bf, force, stream); break;
    case 4: launch_partials_g<4>(dm, sc, bf, force, stream); break;
    case 8: launch_partials_g<8>(dm, sc, bf, force, stream); break;
    case 16: launch_partials_g<16>(dm, sc, bf, force, stream); break;
    default: launch_partials_g<32>(dm, sc, bf, force, stream); break;
  }
}

}  // namespace idto
