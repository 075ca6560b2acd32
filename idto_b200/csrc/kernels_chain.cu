#include <cstdlib>
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
// Chain-lane versions of the inverse-dynamics kernels (tau and the ID partials); see dynamics_chain.cuh
// for the decomposition and kernels_partials.cu for the finite-difference scheme they share:
//   A  tau[t-1] at q_t +- dq e_i: full evaluations (pose computed on the fly, nothing but the contact
//      geometry is kept per evaluation);
//   B  tau[t] at the unperturbed q_{t+1}: body poses + contact geometry computed once per (b,t) slot and
//      shared by the nq groups of the slot;
//   C  dtau_{t+1}/dq_t = M(q_{t+2}) N+_{t+1} / dt^2 (cc:552-561): one bias-free evaluation on the shared pose.
#include "dynamics_chain.cuh"

namespace idto {

namespace {

struct ChainLayout {
  int slots, groups, threads, smem_bytes;
  int nsplit, ncolc;  // the columns of a (b,t) slot are dealt to nsplit CTAs, ncolc columns each
};

// shared memory: tables | per slot 2 shared poses | per group: eval scratch + private pose + 3 tau rows
// tau rows: +dq, -dq, and (+dq) - (-dq) for the 4th-order stencil
__host__ __device__ inline int cgroup_doubles(const DevModel& dm, int nv, int ntau) {
  const int n = ceval_doubles(dm) + cpose_private_doubles(dm) + ntau * nv;
  return n + ((dm.cg_res - n) & 15);  // groups of one warp must not start in the same bank
}
// Padding groups write the shared poses they compute to a dummy slot — except for pruned models (large poses,
// shared memory is what limits them), whose padding groups shadow slot 0 of the CTA instead (same values).
__host__ __device__ inline int dummy_pose_slots(const DevModel& dm) { return dm.prune ? 0 : 1; }
inline int chain_smem_bytes(const DevModel& dm, int nv, int ntau, int slots, int ncolc) {
  return model_smem_bytes(dm) + 8 * nv +
         8 * ((slots + dummy_pose_slots(dm)) * 2 * cpose_doubles(dm) + (slots * ncolc + 1) * cgroup_doubles(dm, nv, ntau));
}

// Padding groups (thread count rounded up to a warp) share one dummy area, so shared memory is sized by
// the groups that do real work: slots * ncolc + 1.  A model whose evaluations are too large for all the
// columns of even one slot (allegro hand: 23 columns x 7.4 KB) splits the columns of a slot over several CTAs;
// each of them computes the two shared poses of the slot for itself.
ChainLayout chain_layout(const DevModel& dm, int ncol, int nv, int ntau) {
  ChainLayout L;
  const int budget = 226 * 1024;
  static const int max_slots = [] {  // tuning knob; 3 slots measured fastest on the quadruped
    const char* e = std::getenv("IDTO_CHAIN_SLOTS");
    return e ? std::max(1, std::atoi(e)) : 64;
  }();
  static const int min_split = [] {  // tuning knob: at least this many CTAs per slot
    const char* e = std::getenv("IDTO_CHAIN_NSPLIT");
    return e ? std::max(1, std::atoi(e)) : 1;
  }();
  for (int nsplit = std::min(min_split, ncol); nsplit <= ncol; ++nsplit) {
    const int ncolc = (ncol + nsplit - 1) / nsplit;
    if (nsplit > min_split && (ncol + nsplit - 2) / (nsplit - 1) == ncolc) continue;  // same chunk size, more CTAs
    const int per_slot = ncolc * dm.cgroup;
    int best = 0;
    for (int s = 1; s <= 64; ++s) {
      const int threads = (s * per_slot + 31) / 32 * 32;
      if (threads <= 320 && chain_smem_bytes(dm, nv, ntau, s, ncolc) <= budget && s <= max_slots) best = s;
    }
    if (best == 0 && ncolc > 1) continue;
    if (best == 0) best = 1;  // (rejected at solver creation: chain_min_smem_bytes)
    L.nsplit = nsplit, L.ncolc = ncolc;
    L.slots = best;
    L.threads = (best * per_slot + 31) / 32 * 32;
    L.groups = best * ncolc + 1;
    L.smem_bytes = chain_smem_bytes(dm, nv, ntau, best, ncolc);
    break;
  }
  return L;
}

}  // namespace

template <int CG, int NLEV, int METHOD>
__global__ void __launch_bounds__(320, 1) k_partials_chain(DevModel dm, SolverConsts sc, SolverBufs bf, int slots,
                                                           int nsplit, int force) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int T = sc.T, nq = sc.nq, nv = sc.nv;
  const int g = threadIdx.x / CG, c = threadIdx.x % CG;
  // columns differentiated here (the path columns go to kernels_path.cu): this CTA takes chunk `chunk` of
  // ncol columns (nsplit == 1: all of them) for each of its slots
  const int ncol = (dm.nfull + nsplit - 1) / nsplit, chunk = blockIdx.x % nsplit;
  const int slot = g / ncol, ii = g % ncol;
  const int icol = chunk * ncol + ii;
  const int i = (dm.itab + dm.o_fullcols)[icol < dm.nfull ? icol : dm.nfull - 1];
  const int sg = (blockIdx.x / nsplit) * slots + slot;
  const bool valid_slot = (slot < slots) && (sg < sc.B * T);
  const bool valid = valid_slot && icol < dm.nfull;
  const int sgx = valid_slot ? sg : (blockIdx.x / nsplit) * slots;  // padding groups shadow the CTA's first slot
  const int b = sgx / T;
  const int t = sgx % T + 1;
  const bool live = valid && (force || bf.ctl[b].derivs_dirty);
  if (!__syncthreads_or(live ? 1 : 0)) return;

  int* si = reinterpret_cast<int*>(smem);
  double* sd = reinterpret_cast<double*>(smem + dm.itab_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + dm.itab_bytes + dm.dtab_bytes);
  double* base = reinterpret_cast<double*>(smem + model_smem_bytes(dm));
  stage_model(dm, si, sd, bar);
  const CModel C = make_cmodel(dm, si, sd);
  double* zrow = base;  // nv zeros
  for (int e = threadIdx.x; e < nv; e += blockDim.x) zrow[e] = 0.0;
  base += nv;
  constexpr int NTAU = METHOD == IDTO_GRAD_CENTRAL4 ? 3 : 2;
  const int pd = cpose_doubles(dm), gd = cgroup_doubles(dm, nv, NTAU);
  // padding groups write their (discarded) poses to a dummy slot, or (same values) to the slot they shadow
  const int sslot = valid_slot ? slot : (dummy_pose_slots(dm) ? slots : 0);
  const PoseSmem PB = make_cpose(dm, base + size_t(sslot) * 2 * pd);
  const PoseSmem PC = make_cpose(dm, base + size_t(sslot) * 2 * pd + pd);
  const int gidx = slot < slots ? g : slots * ncol;  // padding groups share one dummy area
  double* gbase = base + size_t(slots + dummy_pose_slots(dm)) * 2 * pd + size_t(gidx) * gd;
  const EvalSmem S = make_ceval(dm, gbase);
  const PoseSmem PA = make_cpose_private(dm, gbase + ceval_doubles(dm));
  double* T0 = gbase + ceval_doubles(dm) + cpose_private_doubles(dm);
  double* T1 = T0 + nv;
  double* T2 = NTAU > 2 ? T1 + nv : T1;

  const double* qB = bf.st.q + size_t(b) * (T + 1) * nq;
  const double* vB = bf.st.v + size_t(b) * (T + 1) * nv;
  const double* aB = bf.st.a + size_t(b) * T * nv;
  const int tp1 = t < T ? t + 1 : t, tp2 = t < T - 1 ? t + 2 : t;  // clamped rows for predicated-off work
  // pruned models: near lists of the unperturbed poses of this problem, entry j for q_{j+1} (k_tau_chain)
  const int kNearStride = near_stride(dm.nact);
  const int* near_list = dm.prune ? bf.st.near + size_t(b) * T * kNearStride : nullptr;

  // ---- column bookkeeping (every lane computes it: no shuffles needed) ------------------------------------
  const int owner = C.M.qowner[i];
  const int ojt = C.M.jtype[owner], oq0 = C.M.qs[owner];
  const int local = i - oq0;
  const bool quatcol = ojt == IDTO_JOINT_QUAT_FLOATING && local < 4;
  const int sl = (ojt == IDTO_JOINT_QUAT_FLOATING) ? local - 1 : local;
  const double eps = 1.4901161193847656e-08;  // sqrt(2^-52)
  const double qi = qB[size_t(t) * nq + i];
  double dq = eps * fmax(1.0, fabs(qi));
  {
    const double temp = __dadd_rn(qi, dq);  // make dq representable (cc:506-508)
    dq = __dadd_rn(temp, -qi);
  }
  const double dv = dq / sc.dt, da = dv / sc.dt;
  V3 nt3 = {0, 0, 0}, ntp3 = {0, 0, 0};
  if (quatcol) {
    nt3 = quat_nplus_col(qB + size_t(t) * nq + oq0, local);
    ntp3 = quat_nplus_col(qB + size_t(tp1) * nq + oq0, local);
  }
  Perturb pt;
  pt.owner = owner, pt.local = local, pt.sl = sl, pt.quatcol = quatcol;
  __syncthreads();  // zrow

  // ---- phase 0: shared poses of the slot: q_{t+1} by its group 0, q_{t+2} by its group 1 ------------------------
  // (every warp-level primitive of chain_eval is scoped to the calling group, so exactly one group writes each
  // shared pose; the other groups wait at the barrier below.  Earlier builds let every lane of the warps holding
  // groups 0 and 1 run these evaluations and store the same values redundantly: racecheck hazards.)
  const unsigned gm = group_mask<CG>();
  {
    const int ncol_here = min(ncol, dm.nfull - chunk * ncol);  // columns this CTA really has (last chunk: fewer)
    const bool slot_live = valid_slot && (force || bf.ctl[b].derivs_dirty);
    if (slot_live && ii < 2 && ii < ncol_here) {
      Perturb none = pt;
      none.owner = -1;
      const bool second = ii == 1;
      // the pose of q_{t+2} only serves the bias-free evaluations of phase C: no contact geometry
      PairWalk pw;
      pw.near_in = near_list ? near_list + size_t(tp1 - 1) * kNearStride : nullptr;
      pw.skip = second;
      chain_eval<CG, NLEV, kEvalPoseOnly>(C, sc, second ? PC : PB, S, c, qB + size_t(second ? tp2 : tp1) * nq, vB, aB,
                                          none, T0, nullptr, pw);
      if (ncol_here == 1) {  // a single column: its group computes both poses
        pw.skip = true;
        chain_eval<CG, NLEV, kEvalPoseOnly>(C, sc, PC, S, c, qB + size_t(tp2) * nq, vB, aB, none, T0, nullptr, pw);
      }
    }
  }

  constexpr int NK = METHOD == IDTO_GRAD_CENTRAL4 ? 4 : (METHOD == IDTO_GRAD_CENTRAL ? 2 : 1);
  // combine the stencil points of one column: rows r = c, c+CG, ... of column i
  auto emit = [&](double* __restrict__ dst_block, const double* __restrict__ tau_base, bool ok) {
    __syncwarp(gm);
    if (ok && live)
      for (int r = c; r < nv; r += CG) {
        double val;
        if (METHOD == IDTO_GRAD_FORWARD)
          val = (T0[r] - tau_base[r]) / dq;  // cc:531, 539
        else if (METHOD == IDTO_GRAD_CENTRAL)
          val = 0.5 * (T0[r] - T1[r]) / dq;  // cc:785
        else
          val = 2.0 / 3.0 * T2[r] / dq - 1.0 / 12.0 * (T0[r] - T1[r]) / dq;  // cc:782-783 (T2 = tau+ - tau-)
        dst_block[size_t(i) * nv + r] = val;
      }
    __syncwarp(gm);
  };
  auto stash_d1 = [&]() {  // CD4: keep tau(+dq) - tau(-dq) while T0/T1 are reused for +-2dq
    __syncwarp(gm);
    for (int r = c; r < nv; r += CG) T2[r] = T0[r] - T1[r];
    __syncwarp(gm);
  };

  // ---- A: tau[t-1] = ID(q_t^e, v_t^e, a_{t-1}^e)   (cc:526-531, 763-787) --------------------------------
  // (groups without a live column — padding lanes of the last warp, slots beyond the batch, problems whose
  // derivatives are current — skip the evaluations: nothing they would compute is kept)
  PairWalk pwA;  // the perturbed pose can only activate pairs of the near list of q_t
  pwA.near_in = near_list ? near_list + size_t(t - 1) * kNearStride : nullptr;
#pragma unroll 1
  for (int kk = 0; kk < (live ? NK : 0); ++kk) {
    const double m = ((kk & 1) ? -1.0 : 1.0) * ((kk >= 2) ? 2.0 : 1.0);
    if (bf.act_fd) pwA.act_out = bf.act_fd + ((((size_t(b) * T + (t - 1)) * nq + i) * 4 + kk) * dm.np);
    pt.dq = m * dq, pt.cv = m * dv, pt.ca = m * da, pt.uv = 1.0, pt.ua = 1.0, pt.nv3 = nt3, pt.na3 = nt3;
    chain_eval<CG, NLEV, kEvalFull>(C, sc, PA, S, c, qB + size_t(t) * nq, vB + size_t(t) * nv,
                                    aB + size_t(t - 1) * nv, pt, (kk & 1) ? T1 : T0, nullptr, pwA);
    if (METHOD == IDTO_GRAD_CENTRAL4 && kk == 1) stash_d1();
  }
  emit(bf.dqp + (size_t(b) * T + (t - 1)) * nv * nq, bf.st.tau + (size_t(b) * T + (t - 1)) * nv, true);

  __syncthreads();  // shared poses complete
  if (!live) return;  // (no block-wide barrier below)

  // ---- B: tau[t] = ID(q_{t+1}, v_{t+1}^e, a_t^e)   (cc:533-540, 788-814) ---------------------------------
  const int tb = t < T ? t : T - 1;  // a row index that exists even when the result is discarded
#pragma unroll 1
  for (int kk = 0; kk < NK; ++kk) {
    const double m = ((kk & 1) ? -1.0 : 1.0) * ((kk >= 2) ? 2.0 : 1.0);
    pt.dq = 0.0, pt.cv = -(m * dv), pt.ca = -(m * da), pt.uv = 1.0, pt.ua = 1.0 + 1.0;
    pt.nv3 = ntp3, pt.na3 = ntp3 + nt3;
    chain_eval<CG, NLEV, kEvalSharedPose>(C, sc, PB, S, c, qB + size_t(tp1) * nq, vB + size_t(tp1) * nv,
                                          aB + size_t(tb) * nv, pt, (kk & 1) ? T1 : T0);
    if (METHOD == IDTO_GRAD_CENTRAL4 && kk == 1) stash_d1();
  }
  emit(bf.dqt + (size_t(b) * T + tb) * nv * nq, bf.st.tau + (size_t(b) * T + tb) * nv, t < T);

  // ---- C: dtau_dqm[t+1] = M(q_{t+2}) N+_{t+1} / dt^2   (cc:552-561) ---------------------------------------
  pt.dq = 0.0, pt.cv = 0.0, pt.ca = 1.0, pt.uv = 1.0, pt.ua = 1.0, pt.nv3 = ntp3, pt.na3 = ntp3;
  chain_eval<CG, NLEV, kEvalSharedPoseNoBias>(C, sc, PC, S, c, qB + size_t(tp2) * nq, zrow, zrow, pt, T0);
  __syncwarp(gm);
  if (t < T - 1) {
    double* dst = bf.dqm + (size_t(b) * T + (t + 1)) * nv * nq + size_t(i) * nv;
    for (int r = c; r < nv; r += CG) dst[r] = 1 / sc.dt / sc.dt * T0[r];
  }
}

// One CG-lane group per (b, t): everything the trajectory-level cache holds for its step, in one launch —
//   N+(q_{t+1}), v_{t+1} = N+ (q_{t+1} - q_t) / dt, a_t = (v_{t+1} - v_t) / dt   (cc:178-202, 1633-1647; the item t = 0
//   also v_0 = v_init and N+(q_0)),   tau_t = ID(q_{t+1}, v_{t+1}, a_t)  (cc:204-245),   h (cc:1274-1278),
//   its terms of the cost (cc:148-176); the group of a problem that finishes last adds the terms up in step order.
// (Separate k_traj / k_cost launches around the latency-bound tau kernel cost 9 + 11 us and two launch gaps.)
// STASH: additionally write the per-body records of the evaluation for the path columns (kernels_path.cu) into
// the half of `stash` that belongs to this trajectory: ctl[b].stash_sel for the state, the other one for the
// scratch trajectory (k_trust_final flips stash_sel when it adopts the scratch trajectory).
template <int CG, int NLEV, bool STASH = false>
__global__ void __launch_bounds__(128) k_tau_chain(DevModel dm, SolverConsts sc, TrajBuf tb, double* __restrict__ stash,
                                                   size_t stash_half, int scratch, ProbCtl* __restrict__ ctl, int force,
                                                   int* act_base, const double* __restrict__ v_init,
                                                   const double* __restrict__ q_nom, const double* __restrict__ v_nom,
                                                   double* __restrict__ part, int* __restrict__ cnt, int clear_flag) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int T = sc.T, nq = sc.nq, nv = sc.nv;
  const int groups = blockDim.x / CG, grp = threadIdx.x / CG, c = threadIdx.x % CG;
  const int item = blockIdx.x * groups + grp;
  const bool in_range = item < sc.B * T;
  const int b = in_range ? item / T : 0, t = in_range ? item % T : 0;
  const bool live = in_range && (force || ctl[b].traj_dirty);
  if (!__syncthreads_or(live ? 1 : 0)) return;  // nothing stale in this CTA (e.g. the evaluation after an accepted step)
  int* si = reinterpret_cast<int*>(smem);
  double* sd = reinterpret_cast<double*>(smem + dm.itab_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + dm.itab_bytes + dm.dtab_bytes);
  double* base = reinterpret_cast<double*>(smem + model_smem_bytes(dm));
  stage_model(dm, si, sd, bar);
  const CModel C = make_cmodel(dm, si, sd);
  const int gd = cgroup_doubles(dm, nv, 3);
  double* gbase = base + size_t(grp) * gd;
  const EvalSmem S = make_ceval(dm, gbase);
  const PoseSmem PA = make_cpose_private(dm, gbase + ceval_doubles(dm));
  double* T0 = gbase + ceval_doubles(dm) + cpose_private_doubles(dm);
  double* vrow = T0 + nv;    // v_{t+1}
  double* arow = vrow + nv;  // a_t
  Perturb none;
  none.owner = -1, none.local = 0, none.sl = 0, none.quatcol = false;
  none.dq = none.cv = none.ca = 0.0, none.uv = none.ua = 1.0, none.nv3 = none.na3 = {0, 0, 0};
  if (!live) return;  // (chain_eval's warp-level primitives are scoped to the group: idle groups simply leave)
  const unsigned gm = group_mask<CG>();
  const double* qrow = tb.q + (size_t(b) * (T + 1) + t + 1) * nq;  // q_{t+1}
  // ---- N+, v, a of this item: the arithmetic of k_traj (kernels_traj.cu), joint by joint ---------------------------
  for (int k = c; k < dm.nb; k += CG) {
    const int jt = C.M.jtype[k], q0 = C.M.qs[k], v0 = C.M.vs[k];
    const int njv = jt == IDTO_JOINT_QUAT_FLOATING ? 6 : (jt == IDTO_JOINT_PLANAR ? 3 : 1);
    const double* vi = v_init + size_t(b) * nv + v0;
    double vt[6], vn[6];
    V3 col[4];
    joint_velocity(sc, jt, qrow - nq + q0, vi, t, vt, col);  // v_t (and N+(q_t))
    if (t == 0) {
      double* v = tb.v + (size_t(b) * (T + 1)) * nv + v0;
      for (int j = 0; j < njv; ++j) v[j] = vt[j];
      if (jt == IDTO_JOINT_QUAT_FLOATING) {
        double* Np = tb.Nplus + (size_t(b) * (T + 1)) * nv * nq;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          double* dst = Np + size_t(q0 + cc) * nv + v0;
          dst[0] = col[cc].x, dst[1] = col[cc].y, dst[2] = col[cc].z;
        }
      }
    }
    joint_velocity(sc, jt, qrow + q0, vi, t + 1, vn, col);  // v_{t+1} and N+(q_{t+1})
    double* v = tb.v + (size_t(b) * (T + 1) + t + 1) * nv + v0;
    double* a = tb.a + (size_t(b) * T + t) * nv + v0;
    for (int j = 0; j < njv; ++j) {
      const double aj = (vn[j] - vt[j]) / sc.dt;
      v[j] = vn[j], a[j] = aj;
      vrow[v0 + j] = vn[j], arow[v0 + j] = aj;
    }
    if (jt == IDTO_JOINT_QUAT_FLOATING) {
      double* Np = tb.Nplus + (size_t(b) * (T + 1) + t + 1) * nv * nq;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        double* dst = Np + size_t(q0 + cc) * nv + v0;
        dst[0] = col[cc].x, dst[1] = col[cc].y, dst[2] = col[cc].z;
      }
    }
  }
  __syncwarp(gm);
  // ---- tau_t ---------------------------------------------------------------------------------------------------
  double* rec = STASH ? stash + size_t(ctl[b].stash_sel ^ scratch) * stash_half + (size_t(b) * T + t) * dm.nb * kStashDoubles
                      : nullptr;
  PairWalk pw;
  if (dm.prune && tb.near) {
    // what a finite-difference perturbation of one coordinate (|dq| <= 2 sqrt(eps) max(1,|q_i|), cc:506, 763) can
    // move a geometry centre by: |dq| times a lever arm bounded by the chain lengths plus the joint travel
    double qmax = 1.0;
    for (int e = 0; e < nq; ++e) qmax = fmax(qmax, fabs(qrow[e]));
    pw.near_out = tb.near + (size_t(b) * T + t) * near_stride(dm.nact);
    pw.margin = 4e-7 * qmax * (dm.reach + dm.nb * qmax);
  }
  if (act_base) pw.act_out = act_base + (size_t(b) * T + t) * dm.np;
  chain_eval<CG, NLEV, kEvalFull, STASH>(C, sc, PA, S, c, qrow, vrow, arow, none, T0, rec, pw);
  __syncwarp(gm);
  double* tau = tb.tau + (size_t(b) * T + t) * nv;
  for (int r = c; r < nv; r += CG) tau[r] = T0[r];
  // h (cc:1274-1278)
  for (int j = c; j < sc.nu; j += CG) tb.h[size_t(b) * sc.nh + t * sc.nu + j] = T0[sc.unact[j]];
  // ---- cost terms of this item (cc:148-176, diagonal weights): tau_t, and q, v of step t+1 (t = 0: also step 0) -----
  double* pp = part + size_t(b) * (T + 1) * 4;  // [T+1][4]: position, velocity, input term of every step
  for (int task = c; task < (t == 0 ? 5 : 3); task += CG) {
    double cst = 0.0;
    // e^T W e with a dense W (problem_definition.h:38-52; column-major), accumulated row by row like Eigen's
    // (e^T W) e
    auto quad = [&](const double* W, int nn, auto err) {
      double s2 = 0.0;
      for (int j = 0; j < nn; ++j) {
        double row = 0.0;
        for (int i = 0; i < nn; ++i) row += err(i) * W[size_t(j) * nn + i];
        s2 += row * err(j);
      }
      return s2;
    };
    if (task == 2) {
      if (sc.dense_w)
        cst = quad(sc.RM, nv, [&](int i) { return T0[i]; });
      else
        for (int i = 0; i < nv; ++i) cst += T0[i] * sc.R[i] * T0[i];
      pp[t * 4 + 2] = cst;
    } else {
      const int tt = task < 2 ? t + 1 : 0;
      if (task == 0 || task == 3) {  // tasks 0, 3: position
        const double* q = tb.q + (size_t(b) * (T + 1) + tt) * nq;
        const double* qn = q_nom + (size_t(b) * (T + 1) + tt) * nq;
        if (sc.dense_w)
          cst = quad(tt < T ? sc.QqM : sc.QfqM, nq, [&](int i) { return q[i] - qn[i]; });
        else
          for (int i = 0; i < nq; ++i) {
            const double e = q[i] - qn[i];
            cst += e * (tt < T ? sc.Qq[i] : sc.Qfq[i]) * e;
          }
        pp[tt * 4 + 0] = cst;
      } else {  // tasks 1, 4: velocity (v_0 = v_init)
        const double* v = tt == 0 ? v_init + size_t(b) * nv : vrow;
        const double* vn = v_nom + (size_t(b) * (T + 1) + tt) * nv;
        if (sc.dense_w)
          cst = quad(tt < T ? sc.QvM : sc.QfvM, nv, [&](int i) { return v[i] - vn[i]; });
        else
          for (int i = 0; i < nv; ++i) {
            const double e = v[i] - vn[i];
            cst += e * (tt < T ? sc.Qv[i] : sc.Qfv[i]) * e;
          }
        pp[tt * 4 + 1] = cst;
      }
    }
  }
  // ---- the group that finishes last adds the terms up (fixed order: the result does not depend on the schedule) ----
  __threadfence();
  __syncwarp(gm);
  int last = 0;
  if (c == 0) last = atomicAdd(cnt + b, 1) == T - 1 ? 1 : 0;
  last = __shfl_sync(gm, last, (threadIdx.x & 31) / CG * CG);
  if (last && c == 0) {
    cnt[b] = 0;
    __threadfence();
    double run = 0.0;
    for (int tt = 0; tt < T; ++tt) run += (__ldcg(pp + tt * 4) + __ldcg(pp + tt * 4 + 1)) + __ldcg(pp + tt * 4 + 2);
    const double term = __ldcg(pp + T * 4) + __ldcg(pp + T * 4 + 1);
    tb.cost[b] = run * sc.dt + term;
    if (clear_flag) ctl[b].traj_dirty = 0;
  }
}

// A side stream (and the two events of a fork / join) per launching stream and device, created on first use and
// kept for the life of the process.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream& side_stream_of(cudaStream_t parent) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, SideStream> all;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  SideStream& e = all[std::make_pair(dev, parent)];
  if (!e.s) {  // (may happen while the calling thread captures a graph: creating objects is not a stream operation)
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    cudaThreadExchangeStreamCaptureMode(&mode);
    cudaStreamCreateWithFlags(&e.s, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&e.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e.join, cudaEventDisableTiming);
    cudaThreadExchangeStreamCaptureMode(&mode);
  }
  return e;
}

// ---- dispatch on (chain group size, padded tree depth) ------------------------------------------------------
template <int CG, int NLEV>
static void launch_partials_chain_cl(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                                     cudaStream_t stream) {
  // The path columns (kernels_path.cu; the base records come from the trajectory stage) and the full columns are
  // independent: the two kernels run side by side, the second one on a side stream forked from and joined back into
  // `stream`.  The full-column kernel goes first so that its CTAs (one per SM, the long pole) are placed first; one
  // path CTA then fits into the registers and shared memory each of them leaves.
  static const bool side_by_side = [] {
    const char* e = std::getenv("IDTO_PATH_CONCURRENT");
    return !e || std::atoi(e) != 0;
  }();
  if (dm.nfull == 0 || dm.npath == 0 || !side_by_side) {
    launch_partials_path(dm, sc, bf, force, stream);
    if (dm.nfull == 0) return;
  }
  const bool fork = side_by_side && dm.npath > 0;
  const ChainLayout L = chain_layout(dm, dm.nfull, sc.nv, sc.method == IDTO_GRAD_CENTRAL4 ? 3 : 2);
  const int grid = (sc.B * sc.T + L.slots - 1) / L.slots * L.nsplit;
  g_launch_counter += 1;
#define IDTO_LAUNCH_PC(METHOD)                                                                                   \
  {                                                                                                              \
    static bool attr_set[kMaxDevices] = {};                                                                                \
    if (first_use_on_device(attr_set)) {                                                                                             \
      cudaFuncSetAttribute(k_partials_chain<CG, NLEV, METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                           227 * 1024);                                                                          \
    }                                                                                                            \
    k_partials_chain<CG, NLEV, METHOD><<<grid, L.threads, L.smem_bytes, stream>>>(dm, sc, bf, L.slots, L.nsplit,  \
                                                                                  force);                        \
  }
  SideStream* side = fork ? &side_stream_of(stream) : nullptr;
  if (side) cudaEventRecord(side->fork, stream);
  switch (sc.method) {
    case IDTO_GRAD_FORWARD: IDTO_LAUNCH_PC(IDTO_GRAD_FORWARD) break;
    case IDTO_GRAD_CENTRAL: IDTO_LAUNCH_PC(IDTO_GRAD_CENTRAL) break;
    default: IDTO_LAUNCH_PC(IDTO_GRAD_CENTRAL4) break;
  }
#undef IDTO_LAUNCH_PC
  if (side) {
    cudaStreamWaitEvent(side->s, side->fork, 0);
    launch_partials_path(dm, sc, bf, force, side->s);
    cudaEventRecord(side->join, side->s);
    cudaStreamWaitEvent(stream, side->join, 0);
  }
}

template <int CG, int NLEV>
static void launch_tau_chain_cl(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool scratch,
                                bool force, cudaStream_t stream) {
  static const int tau_threads = [] {  // tuning knob
    const char* e = std::getenv("IDTO_TAU_THREADS");
    return e ? std::max(32, std::min(128, std::atoi(e) / 32 * 32)) : 64;
  }();
  const int threads = std::max(tau_threads, CG), groups = threads / CG;  // 2560 (b,t) items x CG lanes: small CTAs reach every SM
  const int smem = model_smem_bytes(dm) + groups * cgroup_doubles(dm, sc.nv, 3) * 8;
  static bool attr_set[kMaxDevices] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(k_tau_chain<CG, NLEV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(k_tau_chain<CG, NLEV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  }
  g_launch_counter += 1;
  const TrajBuf& tb = scratch ? bf.sc : bf.st;
  const int grid = (sc.B * sc.T + groups - 1) / groups;
  if (bf.stash)
    k_tau_chain<CG, NLEV, true><<<grid, threads, smem, stream>>>(dm, sc, tb, bf.stash, bf.stash_half, scratch ? 1 : 0,
                                                                 bf.ctl, force, scratch ? nullptr : bf.act_base,
                                                                 bf.v_init, bf.q_nom, bf.v_nom, bf.part, bf.cnt,
                                                                 scratch ? 0 : 1);
  else
    k_tau_chain<CG, NLEV><<<grid, threads, smem, stream>>>(dm, sc, tb, nullptr, 0, 0, bf.ctl, force,
                                                           scratch ? nullptr : bf.act_base, bf.v_init, bf.q_nom,
                                                           bf.v_nom, bf.part, bf.cnt, scratch ? 0 : 1);
}

// Instantiated (lanes per evaluation, padded tree depth) pairs; anything else falls back to the
// one-lane-per-body kernels (chain_supported() is false).
static int chain_key(const DevModel& dm) {
  const int nl = dm.nlevels <= 2 ? 2 : (dm.nlevels <= 4 ? 4 : 8);
  return dm.cgroup * 16 + nl;
}
#define IDTO_CHAIN_DISPATCH(FN, ...)              \
  switch (chain_key(dm)) {                        \
    case 1 * 16 + 2: FN<1, 2>(__VA_ARGS__); break; \
    case 1 * 16 + 4: FN<1, 4>(__VA_ARGS__); break; \
    case 1 * 16 + 8: FN<1, 8>(__VA_ARGS__); break; \
    case 2 * 16 + 2: FN<2, 2>(__VA_ARGS__); break; \
    case 2 * 16 + 4: FN<2, 4>(__VA_ARGS__); break; \
    case 4 * 16 + 4: FN<4, 4>(__VA_ARGS__); break; \
    case 8 * 16 + 4: FN<8, 4>(__VA_ARGS__); break; \
    default: break;                               \
  }

// Shared memory one CTA of the ID kernels needs for this model (a single slot is the minimum).  The contact-pair
// scratch is sized by dm.nact: every candidate pair for small models, the compacted-list capacity for models
// with more candidates than that (allegro hand: 188).
int chain_min_smem_bytes(const DevModel& dm, int nv, int method) {
  const int ntau = method == IDTO_GRAD_CENTRAL4 ? 3 : 2;
  const int partials = chain_smem_bytes(dm, nv, ntau, 1, 1);  // one slot, one column per CTA
  const int tau = model_smem_bytes(dm) + (64 / dm.cgroup) * cgroup_doubles(dm, nv, 3) * 8;
  return partials > tau ? partials : tau;
}

// Capacity of the per-evaluation active-pair list for a model with more than kMaxActivePairs candidates: the
// largest even count up to kDefaultPairSlots for which ALL full columns of a (b,t) slot still share one CTA (a
// larger list splits the columns over several CTAs, each with few warps and its own copy of the shared poses: the
// allegro hand's ID partials ran 2x slower with 64 slots than with 32); if not even kMaxActivePairs slots fit that
// way the columns are split anyway and the list gets kDefaultPairSlots.
int chain_fit_pair_slots(DevModel dm, int nv) {
  const int hi = std::min(dm.npp, kDefaultPairSlots);
  for (int n = hi; n >= kMaxActivePairs; n -= 2) {
    dm.nact = n;
    if (chain_smem_bytes(dm, nv, 2, 1, std::max(dm.nfull, 1)) <= 226 * 1024) return n;
  }
  return hi;
}

bool chain_supported(const DevModel& dm) {
  if (!dm.chain_ok) return false;
  switch (chain_key(dm)) {
    case 1 * 16 + 2: case 1 * 16 + 4: case 1 * 16 + 8: case 2 * 16 + 2: case 2 * 16 + 4: case 4 * 16 + 4:
    case 8 * 16 + 4: return true;
    default: return false;
  }
}

void launch_partials_chain(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                           cudaStream_t stream) {
  IDTO_CHAIN_DISPATCH(launch_partials_chain_cl, dm, sc, bf, force, stream)
}
void launch_tau_chain(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool scratch, bool force,
                      cudaStream_t stream) {
  IDTO_CHAIN_DISPATCH(launch_tau_chain_cl, dm, sc, bf, scratch, force, stream)
}

}  // namespace idto
