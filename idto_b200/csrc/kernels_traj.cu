// Trajectory-level cache entries: N+(q), v, a (cc:178-202, cc:1633-1647), tau (cc:204-226),
// cost (cc:136-176) and the equality-constraint violations h (cc:1267-1279).
#include "dynamics.cuh"
#include "reduce.cuh"

namespace idto {

namespace {

}  // namespace

// One thread per (problem, step, joint): v_t, the quaternion block of N+(q_t) (the rest of N+ is the
// identity pattern preset at creation) and a_t = (v_{t+1} - v_t)/dt (cc:193-202) with v_{t+1} recomputed in
// place — bit-identical to the stored one — so that no thread waits for another.  (The first version
// walked a whole problem in one 64-thread CTA: 27 us.)
__global__ void __launch_bounds__(128) k_traj(DevModel dm, SolverConsts sc, TrajBuf tb, const double* __restrict__ v_init,
                                              const ProbCtl* __restrict__ ctl, int force) {
  const int T = sc.T, nq = sc.nq, nv = sc.nv;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= sc.B * (T + 1) * dm.nb) return;
  const int k = idx % dm.nb, t = (idx / dm.nb) % (T + 1), b = idx / (dm.nb * (T + 1));
  if (!force && !ctl[b].traj_dirty) return;
  const int jt = (dm.itab + dm.o_jtype)[k], q0 = (dm.itab + dm.o_qs)[k], v0 = (dm.itab + dm.o_vs)[k];
  const int njv = jt == IDTO_JOINT_QUAT_FLOATING ? 6 : (jt == IDTO_JOINT_PLANAR ? 3 : 1);
  const double* qt = tb.q + (size_t(b) * (T + 1) + t) * nq + q0;
  const double* vi = v_init + size_t(b) * nv + v0;
  double vt[6], vn[6];
  V3 col[4];
  joint_velocity(sc, jt, qt, vi, t, vt, col);
  double* v = tb.v + (size_t(b) * (T + 1) + t) * nv + v0;
  for (int j = 0; j < njv; ++j) v[j] = vt[j];
  if (jt == IDTO_JOINT_QUAT_FLOATING) {
    double* Np = tb.Nplus + (size_t(b) * (T + 1) + t) * nv * nq;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double* dst = Np + size_t(q0 + c) * nv + v0;
      dst[0] = col[c].x, dst[1] = col[c].y, dst[2] = col[c].z;
    }
  }
  if (t < T) {
    joint_velocity(sc, jt, qt + nq, vi, t + 1, vn, col);
    double* a = tb.a + (size_t(b) * T + t) * nv + v0;
    for (int j = 0; j < njv; ++j) a[j] = (vn[j] - vt[j]) / sc.dt;
  }
}

// tau_t = ID(q_{t+1}, v_{t+1}, a_t): one G-lane group per (b, t).
template <int G>
__global__ void __launch_bounds__(128) k_tau(DevModel dm, SolverConsts sc, TrajBuf tb,
                                             const ProbCtl* __restrict__ ctl, int force) {
  extern __shared__ __align__(16) unsigned char smem[];
  int* si = reinterpret_cast<int*>(smem);
  double* sd = reinterpret_cast<double*>(smem + dm.itab_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + dm.itab_bytes + dm.dtab_bytes);
  double* gbase = reinterpret_cast<double*>(smem + model_smem_bytes(dm));
  stage_model(dm, si, sd, bar);
  const SModel M = make_smodel(dm, si, sd);
  const int groups = blockDim.x / G, grp = threadIdx.x / G, k = threadIdx.x % G;
  double* gslab = gbase + size_t(grp) * (pos_smem_doubles(dm) + vel_smem_doubles(dm));
  const PosSmem P = make_pos_smem(dm, gslab);
  const VelSmem S = make_vel_smem(dm, gslab + pos_smem_doubles(dm));
  const int T = sc.T, nq = sc.nq, nv = sc.nv;
  const int item = blockIdx.x * groups + grp;
  const bool in_range = item < sc.B * T;
  const int b = in_range ? item / T : 0, t = in_range ? item % T : 0;
  const bool live = in_range && (force || ctl[b].traj_dirty);
  const bool body = k < M.nb;
  double qb[7] = {1, 0, 0, 0, 0, 0, 0}, vb[6] = {0, 0, 0, 0, 0, 0}, ab[6] = {0, 0, 0, 0, 0, 0}, taub[6];
  int jt = 0, q0 = 0, v0 = 0;
  if (body) {
    jt = M.jtype[k], q0 = M.qs[k], v0 = M.vs[k];
    const double* q = tb.q + (size_t(b) * (T + 1) + t + 1) * nq + q0;
    const double* v = tb.v + (size_t(b) * (T + 1) + t + 1) * nv + v0;
    const double* a = tb.a + (size_t(b) * T + t) * nv + v0;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j < joint_nq(jt)) qb[j] = q[j];
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (j < joint_nv(jt)) vb[j] = v[j], ab[j] = a[j];
  }

  PositionPhase<G>(M, P, sc, k, qb);
  VelocityPhase<G>(M, P, S, sc, k, vb, ab, true, taub);
  if (live && body) {
    double* tau = tb.tau + (size_t(b) * T + t) * nv + v0;
#pragma unroll
    for (int j = 0; j < 6; ++j)
      if (j < joint_nv(jt)) tau[j] = taub[j];
  }
}

// Cost (cc:148-176, diagonal weights) and h (cc:1274-1278); one CTA per problem, one thread per
// (step, term) — position, velocity and input cost of a step are independent sums.
__global__ void __launch_bounds__(256) k_cost(SolverConsts sc, TrajBuf tb, const double* __restrict__ q_nom,
                                              const double* __restrict__ v_nom, ProbCtl* __restrict__ ctl, int force,
                                              int clear_flag) {
  __shared__ double red[32];
  __shared__ double part[3 * 128];
  const int b = blockIdx.x;
  if (!force && !ctl[b].traj_dirty) return;
  const int T = sc.T, nq = sc.nq, nv = sc.nv, tid = threadIdx.x, nt = blockDim.x;
  const double* q = tb.q + size_t(b) * (T + 1) * nq;
  const double* v = tb.v + size_t(b) * (T + 1) * nv;
  const double* tau = tb.tau + size_t(b) * T * nv;
  const double* qn = q_nom + size_t(b) * (T + 1) * nq;
  const double* vn = v_nom + size_t(b) * (T + 1) * nv;
  double run = 0.0, term = 0.0;
  for (int t0 = 0; t0 <= T; t0 += 128) {  // 128 steps per pass
    __syncthreads();
    for (int task = tid; task < 3 * 128; task += nt) {
      const int what = task / 128, t = t0 + task % 128;
      double c = 0.0;
      if (t <= T) {
        if (what == 0) {
#pragma unroll 4
          for (int i = 0; i < nq; ++i) {
            const double e = q[t * nq + i] - qn[t * nq + i];
            c += e * (t < T ? sc.Qq[i] : sc.Qfq[i]) * e;
          }
        } else if (what == 1) {
#pragma unroll 4
          for (int i = 0; i < nv; ++i) {
            const double e = v[t * nv + i] - vn[t * nv + i];
            c += e * (t < T ? sc.Qv[i] : sc.Qfv[i]) * e;
          }
        } else if (t < T) {
#pragma unroll 4
          for (int i = 0; i < nv; ++i) c += tau[t * nv + i] * sc.R[i] * tau[t * nv + i];
        }
      }
      part[task] = c;
    }
    __syncthreads();
    const int t = t0 + tid;
    if (tid < 128 && t <= T) {
      if (t < T)
        run += (part[tid] + part[128 + tid]) + part[256 + tid];
      else
        term += part[tid] + part[128 + tid];
    }
  }
  run = block_sum(run, red);
  term = block_sum(term, red);
  if (tid == 0) tb.cost[b] = run * sc.dt + term;
  for (int idx = tid; idx < sc.nh; idx += nt) {
    const int t = idx / sc.nu, j = idx % sc.nu;
    tb.h[size_t(b) * sc.nh + idx] = tau[t * nv + sc.unact[j]];
  }
  if (clear_flag && tid == 0) ctl[b].traj_dirty = 0;
}

template <int G>
static void launch_tau_g(const DevModel& dm, const SolverConsts& sc, const TrajBuf& tb, const ProbCtl* ctl,
                         bool force, cudaStream_t stream) {
  const int threads = 128, groups = threads / G;
  const int smem = model_smem_bytes(dm) + groups * (pos_smem_doubles(dm) + vel_smem_doubles(dm)) * 8;
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_tau<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  const int grid = (sc.B * sc.T + groups - 1) / groups;
  g_launch_counter += 1;
  k_tau<G><<<grid, threads, smem, stream>>>(dm, sc, tb, ctl, force ? 1 : 0);
}

void launch_traj(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool scratch, bool force,
                 cudaStream_t stream) {
  if (use_chain_kernels(dm)) return;  // k_tau_chain computes v, a, N+ of its own item (launch_tau)
  const TrajBuf& tb = scratch ? bf.sc : bf.st;
  g_launch_counter += 1;
  k_traj<<<(sc.B * (sc.T + 1) * dm.nb + 127) / 128, 128, 0, stream>>>(dm, sc, tb, bf.v_init, bf.ctl, force ? 1 : 0);
}

void launch_tau(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool scratch, bool force,
                cudaStream_t stream) {
  const TrajBuf& tb = scratch ? bf.sc : bf.st;
  if (use_chain_kernels(dm)) {  // one launch: N+, v, a, tau, the per-body records, cost and h
    launch_tau_chain(dm, sc, bf, scratch, force, stream);
    return;
  }
  switch (dm.group) {
    case 2: launch_tau_g<2>(dm, sc, tb, bf.ctl, force, stream); break;
    case 4: launch_tau_g<4>(dm, sc, tb, bf.ctl, force, stream); break;
    case 8: launch_tau_g<8>(dm, sc, tb, bf.ctl, force, stream); break;
    case 16: launch_tau_g<16>(dm, sc, tb, bf.ctl, force, stream); break;
    default: launch_tau_g<32>(dm, sc, tb, bf.ctl, force, stream); break;
  }
  g_launch_counter += 1;
  k_cost<<<sc.B, 256, 0, stream>>>(sc, tb, bf.q_nom, bf.v_nom, bf.ctl, force ? 1 : 0, scratch ? 0 : 1);
}

}  // namespace idto
