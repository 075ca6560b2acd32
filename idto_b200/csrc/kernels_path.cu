// ID partials of the "path" columns by single-lane subtree evaluations.
//
// Reference: the same finite-difference scheme as kernels_chain.cu (CalcInverseDynamicsPartialsFiniteDiff /
// ...CentralDiff, optimizer/trajectory_optimizer.cc:426-885).  Perturbing q_t[i] (and v, a through the FIXED
// N+ column) changes the kinematics of the owner joint's SUBTREE only; every other body keeps its pose,
// velocity, acceleration and own inertial force bit for bit, and the perturbed tau differs from the base one
// only in the rows of the subtree joints and of the owner's ancestors (the rows whose subtree sum of forces
// contains a changed body).  The reference (and the full-evaluation kernel) recompute all 13 bodies of the
// quadruped for each of the 4446 evaluations of a derivative update; for a knee column one body changes.
//
// For columns whose owner is a 1-dof joint with a simple chain below it and only world-anchored contact
// partners (DevModel::npath: 12 of the quadruped's 19 columns) one THREAD evaluates a perturbed
// tau: it restarts the outward pass at the owner from the parent's state, walks down the chain (pose,
// velocity, acceleration, inertial + gravity force, contact), then walks the inward pass up to the root,
// adding the unchanged siblings' subtree totals.  Everything unchanged comes from the per-body records of
// the base evaluations E_t = ID(q_{t+1}, v_{t+1}, a_t) that k_tau_chain wrote to HBM (48 doubles per
// body; 12.8 MB for 64 x 40 x 13 bodies, L2 resident): A perturbs E_{t-1}, B perturbs E_t, C (mass-matrix
// column) needs the pose of E_{t+1}.  The arithmetic per body and the order of every accumulation are
// those of chain_eval, so the unaffected rows are exactly zero and the affected ones agree with the full
// evaluation to the last bits.  Shared memory holds only the baked tables: occupancy is bounded by registers.
#include "dynamics_chain.cuh"

namespace idto {

namespace {

constexpr int kMaxDown = 4;  // bodies in the owner's subtree chain (checked at model creation)
constexpr int kMaxUp = 7;    // ancestors of the owner
constexpr int kMaxRows = kMaxDown + 6 + (kMaxUp - 1);  // affected rows: subtree joints + ancestors' joints

struct GModel {  // the baked tables (staged in shared memory by one bulk TMA copy per CTA)
  const int *parent, *jtype, *qs, *vs, *nchild, *child, *flags, *gbody, *gtype, *pA, *pB, *gslot;
  const double *XPF, *RMB, *axis, *mass, *com, *inertia, *damping, *gdims, *XBG, *XWGs;
  int nb, nbp, nq, nv, ng, np;
  V3 g;
};
__device__ __forceinline__ GModel make_gmodel(const DevModel& dm, const int* si, const double* sd) {
  GModel M;
  M.parent = si + dm.o_parent, M.jtype = si + dm.o_jtype, M.qs = si + dm.o_qs, M.vs = si + dm.o_vs;
  M.nchild = si + dm.o_nchild, M.child = si + dm.o_child, M.flags = si + dm.o_flags;
  M.gbody = si + dm.o_gbody, M.gtype = si + dm.o_gtype, M.pA = si + dm.o_pA, M.pB = si + dm.o_pB;
  M.gslot = si + dm.o_gslot;
  M.XPF = sd + dm.o_XPF, M.RMB = sd + dm.o_RMB, M.axis = sd + dm.o_axis, M.mass = sd + dm.o_mass;
  M.com = sd + dm.o_com, M.inertia = sd + dm.o_inertia, M.damping = sd + dm.o_damping;
  M.gdims = sd + dm.o_gdims, M.XBG = sd + dm.o_XBG, M.XWGs = sd + dm.o_XWGs;
  M.nb = dm.nb, M.nbp = dm.nbp, M.nq = dm.nq, M.nv = dm.nv, M.ng = dm.ng, M.np = dm.np;
  M.g = {dm.gx, dm.gy, dm.gz};
  return M;
}

struct PathPlan {
  int owner, ndown, nup;
  int down[kMaxDown];  // owner, its child, ... leaf
  int up[kMaxUp];      // parent of the owner, ... root
};

__device__ __forceinline__ M3 unstash_R(const double* rec, int off) {
  M3 R;
#pragma unroll
  for (int e = 0; e < 9; ++e) R.m[e] = rec[off + e];
  return R;
}

// One perturbed evaluation.  `st`: records of the base evaluation whose inputs are perturbed ([nb][48]).
// rows[] receives tau of the affected joints: first the subtree joints (one row each), then the ancestors'
// joints in path order (1, 3 or 6 rows each).
template <int MODE>
__device__ __forceinline__ void path_eval(const GModel& M, const SolverConsts& sc, const PathPlan& pl,
                                          const double* __restrict__ st, const double* __restrict__ q,
                                          const double* __restrict__ v, const double* __restrict__ a,
                                          const Perturb& pt, double* rows, int* act_out = nullptr) {
  constexpr bool kPose = MODE == kEvalFull;
  constexpr bool kBias = MODE == kEvalFull || MODE == kEvalSharedPose;
  // ---- outward pass over the subtree chain ----------------------------------------------------------------
  BodyState par;
  if (pl.nup > 0) {
    const double* rp = st + size_t(pl.up[0]) * kStashDoubles;
    par.R = unstash_R(rp, kStR), par.p = unstash_V(rp, kStP);
    if (kBias) {
      par.w = unstash_V(rp, kStW), par.v = unstash_V(rp, kStV), par.al = unstash_V(rp, kStAl), par.ac = unstash_V(rp, kStAc);
    } else {  // bias-free: only the owner's joint accelerates
      par.w = par.v = par.al = par.ac = {0, 0, 0};
    }
  } else {
    par.R = identity3();
    par.p = par.w = par.v = par.al = par.ac = {0, 0, 0};
  }
  double Fpre[kMaxDown][6], Pd[kMaxDown][3], AXd[kMaxDown][3];
#pragma unroll 1
  for (int d = 0; d < pl.ndown; ++d) {
    const int b = pl.down[d];
    const int parent = M.parent[b], jtype = M.jtype[b], q0 = M.qs[b], v0 = M.vs[b];
    const V3 axis = load_V(M.axis, M.nbp, b);
    M3 R_WB, R_WF;
    V3 p_WB;
    if (kPose) {
      double qb0 = q[q0];
      if (b == pt.owner) qb0 += pt.dq;
      const M3 R_PF = load_R(M.XPF, M.nbp, b);
      const V3 p_PF = load_V(M.XPF + 9 * M.nbp, M.nbp, b);
      M3 R_FM = identity3();
      V3 p_FM = {0, 0, 0};
      if (jtype == IDTO_JOINT_REVOLUTE)
        R_FM = axis_angle_R(axis, qb0);
      else
        p_FM = qb0 * axis;
      M3 R_PB = mul(R_PF, R_FM);
      if (!(M.flags[b] & 1)) R_PB = mul(R_PB, load_R(M.RMB, M.nbp, b));
      const V3 p_PB = p_PF + mul(R_PF, p_FM);
      R_WB = mul(par.R, R_PB);
      p_WB = par.p + mul(par.R, p_PB);
      R_WF = mul(par.R, R_PF);
    } else {
      const double* rb = st + size_t(b) * kStashDoubles;
      R_WB = unstash_R(rb, kStR), R_WF = unstash_R(rb, kStRF), p_WB = unstash_V(rb, kStP);
    }
    // joint velocity / acceleration with the finite-difference perturbation (1-dof joints only here)
    double vb[6] = {0, 0, 0, 0, 0, 0}, ab[6] = {0, 0, 0, 0, 0, 0};
    if (kBias) vb[0] = v[v0], ab[0] = a[v0];  // bias-free: M(q) times the unit acceleration of the owner only
    if (b == pt.owner) vb[0] += pt.cv * pt.uv, ab[0] += pt.ca * pt.ua;
    V3 wF, vF, w_rel = {0, 0, 0}, v_rel = {0, 0, 0};
    if (kBias) {
      hinge_map(jtype, axis, vb, &wF, &vF);
      w_rel = mul(R_WF, wF), v_rel = mul(R_WF, vF);
    }
    hinge_map(jtype, axis, ab, &wF, &vF);
    const V3 al_rel = mul(R_WF, wF), a_rel = mul(R_WF, vF);
    V3 w, vv, al, ac;
    if (parent >= 0) {
      const V3 r = p_WB - par.p;
      w = par.w + w_rel;
      vv = par.v + cross(par.w, r) + v_rel;
      al = par.al + cross(par.w, w_rel) + al_rel;
      ac = par.ac + cross(par.al, r) + cross(par.w, cross(par.w, r)) + 2.0 * cross(par.w, v_rel) + a_rel;
    } else {
      w = w_rel, vv = v_rel, al = al_rel, ac = a_rel;
    }
    par.R = R_WB, par.p = p_WB, par.w = w, par.v = vv, par.al = al, par.ac = ac;
    // own spatial force about Bo in W: inertial - gravity
    const double m = M.mass[b];
    const V3 cm = mul(R_WB, load_V(M.com, M.nbp, b));
    const double* I = M.inertia;
    const M3 IB = {{I[b], I[3 * M.nbp + b], I[4 * M.nbp + b], I[3 * M.nbp + b], I[M.nbp + b], I[5 * M.nbp + b],
                    I[4 * M.nbp + b], I[5 * M.nbp + b], I[2 * M.nbp + b]}};
    const V3 Iw = mul(R_WB, mul(IB, tmul(R_WB, w)));
    const V3 Ial = mul(R_WB, mul(IB, tmul(R_WB, al)));
    V3 f = m * (ac + cross(al, cm) + cross(w, cross(w, cm)));
    V3 t = Ial + cross(w, Iw) + m * cross(cm, ac);
    if (kBias) {
      const V3 fg = m * M.g;
      t = t - cross(cm, fg), f = f - fg;
      rows[d] = M.damping[v0] * vb[0];  // generalized applied force: joint damping (cc:232)
    } else {
      rows[d] = 0.0;
    }
    // contact forces on this body in pair order (cc:272-384); the partner geometry is world-anchored
    if (kBias && M.gslot[b] >= 0) {
      V3 Ft = {0, 0, 0}, Ff = {0, 0, 0};
      for (int ip = 0; ip < M.np; ++ip) {
        const int gA = M.pA[ip], gB = M.pB[ip];
        const int bA = M.gbody[gA], bB = M.gbody[gB];
        if (bA != b && bB != b) continue;
        M3 R_WGa, R_WGb;
        V3 p_WGa, p_WGb;
        if (bA >= 0) {
          R_WGa = mul(R_WB, load_R(M.XBG, M.ng, gA)), p_WGa = p_WB + mul(R_WB, load_V(M.XBG + 9 * M.ng, M.ng, gA));
        } else {
          R_WGa = load_R(M.XWGs, M.ng, gA), p_WGa = load_V(M.XWGs + 9 * M.ng, M.ng, gA);
        }
        if (bB >= 0) {
          R_WGb = mul(R_WB, load_R(M.XBG, M.ng, gB)), p_WGb = p_WB + mul(R_WB, load_V(M.XBG + 9 * M.ng, M.ng, gB));
        } else {
          R_WGb = load_R(M.XWGs, M.ng, gB), p_WGb = load_V(M.XWGs + 9 * M.ng, M.ng, gB);
        }
        const V3 dimA = load_V(M.gdims, M.ng, gA), dimB = load_V(M.gdims, M.ng, gB);
        double distance;
        V3 p_ACa, p_BCb, nhat_BA_W;
        if (M.gtype[gA] == IDTO_GEOM_SPHERE) {
          const PointDist pd = point_to_shape(M.gtype[gB], dimB, R_WGb, p_WGb, p_WGa);
          distance = pd.distance - dimA.x;
          p_BCb = pd.p_GN;
          nhat_BA_W = pd.grad_W;
          p_ACa = (-dimA.x) * tmul(R_WGa, pd.grad_W);
        } else {
          const PointDist pd = point_to_shape(M.gtype[gA], dimA, R_WGa, p_WGa, p_WGb);
          distance = pd.distance - dimB.x;
          p_ACa = pd.p_GN;
          nhat_BA_W = -pd.grad_W;
          p_BCb = (-dimB.x) * tmul(R_WGb, pd.grad_W);
        }
        double fn_c = 0.0;
        if (distance <= sc.threshold) {
          const double exponent = -distance / sc.sigma;
          fn_c = exponent >= 37 ? -sc.k * distance : sc.sigma * sc.k * log(1 + exp(exponent));
        }
        if (act_out) act_out[ip] = fn_c > 0.0 ? 1 : 0;  // debug trace (idto_debug_pair_trace)
        if (!(fn_c > 0.0)) continue;
        const V3 nhat = -nhat_BA_W;
        const V3 p_WC = 0.5 * ((mul(R_WGa, p_ACa) + p_WGa) + (mul(R_WGb, p_BCb) + p_WGb));
        V3 v_Ac = {0, 0, 0}, v_Bc = {0, 0, 0};
        if (bA >= 0) v_Ac = vv + cross(w, p_WC - p_WB);
        if (bB >= 0) v_Bc = vv + cross(w, p_WC - p_WB);
        const V3 v_AcBc = v_Bc - v_Ac;
        const double vn = dot(nhat, v_AcBc);
        const V3 vt = v_AcBc - vn * nhat;
        double dissipation_factor = 0.0;
        const double s = vn / sc.vd;
        if (s < 0) {
          dissipation_factor = 1 - s;
        } else if (s < 2) {
          dissipation_factor = (s - 2) * (s - 2) / 4;
        }
        const double fn = fn_c * dissipation_factor;
        const V3 that_regularized = (-1.0 / sqrt(sc.vs * sc.vs + dot(vt, vt))) * vt;
        const V3 ft_BC = (sc.mu * fn) * that_regularized;
        const V3 f_BC = fn * nhat + ft_BC;
        const V3 pc = p_WC - p_WB;
        if (bA == b) Ft = Ft + cross(pc, -f_BC), Ff = Ff - f_BC;
        if (bB == b) Ft = Ft + cross(pc, f_BC), Ff = Ff + f_BC;
      }
      t = t - Ft, f = f - Ff;
    }
    Fpre[d][0] = t.x, Fpre[d][1] = t.y, Fpre[d][2] = t.z, Fpre[d][3] = f.x, Fpre[d][4] = f.y, Fpre[d][5] = f.z;
    Pd[d][0] = p_WB.x, Pd[d][1] = p_WB.y, Pd[d][2] = p_WB.z;
    const V3 ax = mul(R_WF, axis);
    AXd[d][0] = ax.x, AXd[d][1] = ax.y, AXd[d][2] = ax.z;
  }
  // ---- inward pass: the subtree chain ... -------------------------------------------------------------------
  V3 tc = {0, 0, 0}, fc = {0, 0, 0}, pc = {0, 0, 0};  // running subtree total and its reference point
#pragma unroll 1
  for (int d = pl.ndown - 1; d >= 0; --d) {
    V3 Tt = {Fpre[d][0], Fpre[d][1], Fpre[d][2]}, Tf = {Fpre[d][3], Fpre[d][4], Fpre[d][5]};
    const V3 p_WB = {Pd[d][0], Pd[d][1], Pd[d][2]};
    if (d + 1 < pl.ndown) {
      const V3 rc = pc - p_WB;
      Tt = Tt + tc + cross(rc, fc);
      Tf = Tf + fc;
    }
    const V3 ax = {AXd[d][0], AXd[d][1], AXd[d][2]};
    rows[d] += M.jtype[pl.down[d]] == IDTO_JOINT_REVOLUTE ? dot(ax, Tt) : dot(ax, Tf);
    tc = Tt, fc = Tf, pc = p_WB;
  }
  // ---- ... then the ancestors: own part from the base records, unchanged siblings' totals from the base ------
  int nr = pl.ndown, pathchild = pl.owner;
#pragma unroll 1
  for (int u = 0; u < pl.nup; ++u) {
    const int ab = pl.up[u];
    const double* ra = st + size_t(ab) * kStashDoubles;
    const V3 p_WB = unstash_V(ra, kStP);
    V3 Tt = {0, 0, 0}, Tf = {0, 0, 0};
    if (kBias) Tt = unstash_V(ra, kStFpre), Tf = unstash_V(ra, kStFpre + 3);
    const int nchild = M.nchild[ab];
    for (int ci = 0; ci < nchild; ++ci) {
      const int ch = M.child[ci * M.nbp + ab];
      if (ch == pathchild) {
        const V3 rc = pc - p_WB;
        Tt = Tt + tc + cross(rc, fc);
        Tf = Tf + fc;
      } else if (kBias) {
        const double* rch = st + size_t(ch) * kStashDoubles;
        const V3 tb = unstash_V(rch, kStFtot), fb = unstash_V(rch, kStFtot + 3);
        const V3 rc = unstash_V(rch, kStP) - p_WB;
        Tt = Tt + tb + cross(rc, fb);
        Tf = Tf + fb;
      }
    }
    const int jtype = M.jtype[ab], v0 = M.vs[ab];
    if (jtype == IDTO_JOINT_REVOLUTE || jtype == IDTO_JOINT_PRISMATIC) {
      const double damp = kBias ? M.damping[v0] * v[v0] : 0.0;
      const V3 ax = unstash_V(ra, kStAX);
      rows[nr++] = damp + (jtype == IDTO_JOINT_REVOLUTE ? dot(ax, Tt) : dot(ax, Tf));
    } else {
      const M3 R_WF = unstash_R(ra, kStRF);
      const V3 tF = tmul(R_WF, Tt), fF = tmul(R_WF, Tf);
      if (jtype == IDTO_JOINT_PLANAR) {
        rows[nr] = (kBias ? M.damping[v0] * v[v0] : 0.0) + fF.x;
        rows[nr + 1] = (kBias ? M.damping[v0 + 1] * v[v0 + 1] : 0.0) + fF.y;
        rows[nr + 2] = (kBias ? M.damping[v0 + 2] * v[v0 + 2] : 0.0) + tF.z;
        nr += 3;
      } else {
        const double r6[6] = {tF.x, tF.y, tF.z, fF.x, fF.y, fF.z};
#pragma unroll
        for (int j = 0; j < 6; ++j) rows[nr + j] = (kBias ? M.damping[v0 + j] * v[v0 + j] : 0.0) + r6[j];
        nr += 6;
      }
    }
    tc = Tt, fc = Tf, pc = p_WB, pathchild = ab;
  }
}

}  // namespace

// One thread per (phase A/B/C, path column, problem, step t = 1..T, stencil point).
// At most 176 registers per thread, so that one CTA of this kernel (128 x 176 = 22 528 registers, 7 KB of shared
// memory) fits next to a CTA of k_partials_chain (256 x 168 = 43 008 registers, 221 KB) on the same SM: the two kernels are
// independent and run concurrently (kernels_chain.cu), this one — L2-latency bound — in the issue slots the other —
// a chain of dependent fp64 operations — leaves idle.
#ifndef IDTO_PATH_MAXREG
#define IDTO_PATH_MAXREG 176
#endif
template <int METHOD>
__global__ void __maxnreg__(IDTO_PATH_MAXREG) k_partials_path(DevModel dm, SolverConsts sc, SolverBufs bf, int force) {
  const int T = sc.T, nq = sc.nq, nv = sc.nv, np = dm.npath;
  // consecutive threads take consecutive (problem, step) items of the SAME column and phase: a warp runs one
  // code path with one trip count (mixing the phases in a warp serialised three instantiations: 523 us)
  extern __shared__ __align__(16) unsigned char smem[];
  int* si = reinterpret_cast<int*>(smem);
  double* sd = reinterpret_cast<double*>(smem + dm.itab_bytes);
  stage_model(dm, si, sd, reinterpret_cast<uint64_t*>(smem + dm.itab_bytes + dm.dtab_bytes));
  // NK adjacent lanes share one (phase, column, item) and evaluate one stencil point each (+dq, -dq, +2dq, -2dq):
  // twice / four times the warps for the same registers per thread, which is what hides the L2 latency of the
  // records; the stencil is combined with shuffles.  The mass-matrix phase C has a single evaluation (lane 0).
  constexpr int NK = METHOD == IDTO_GRAD_CENTRAL4 ? 4 : (METHOD == IDTO_GRAD_CENTRAL ? 2 : 1);
  const long idx = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nitem = long(sc.B) * T;
  if (idx >= nitem * 3 * np * NK) return;
  const int kk = int(idx % NK);
  const long gidx = idx / NK;
  const int item = int(gidx % nitem), ci = int((gidx / nitem) % np), phase = int(gidx / (nitem * np));
  const int t = item % T + 1, b = item / T;
  if (!force && !bf.ctl[b].derivs_dirty) return;
  if ((phase == 1 && t >= T) || (phase == 2 && (t >= T - 1 || kk > 0))) return;
  const GModel M = make_gmodel(dm, si, sd);
  const int i = (si + dm.o_pathcols)[ci];
  PathPlan pl;
  pl.owner = (si + dm.o_qowner)[i];
  pl.ndown = 0;
  for (int k = pl.owner;;) {
    pl.down[pl.ndown++] = k;
    if (M.nchild[k] == 0 || pl.ndown == kMaxDown) break;
    k = M.child[k];  // child 0
  }
  pl.nup = 0;
  for (int k = M.parent[pl.owner]; k >= 0 && pl.nup < kMaxUp; k = M.parent[k]) pl.up[pl.nup++] = k;

  const double* qB = bf.st.q + size_t(b) * (T + 1) * nq;
  const double* vB = bf.st.v + size_t(b) * (T + 1) * nv;
  const double* aB = bf.st.a + size_t(b) * T * nv;
  const double* stB = bf.stash + size_t(bf.ctl[b].stash_sel) * bf.stash_half + size_t(b) * T * dm.nb * kStashDoubles;
  auto rec_of = [&](int tt) { return stB + size_t(tt) * dm.nb * kStashDoubles; };
  const double eps = 1.4901161193847656e-08;  // sqrt(2^-52)
  const double qi = qB[size_t(t) * nq + i];
  double dq = eps * fmax(1.0, fabs(qi));
  {
    const double temp = __dadd_rn(qi, dq);  // make dq representable (cc:506-508)
    dq = __dadd_rn(temp, -qi);
  }
  const double dv = dq / sc.dt, da = dv / sc.dt;
  Perturb pt;
  pt.owner = pl.owner, pt.local = 0, pt.sl = 0, pt.quatcol = false;
  pt.nv3 = pt.na3 = {0, 0, 0};
  double R0[kMaxRows], R1[kMaxRows], R2[kMaxRows], R3[kMaxRows];
  for (int r = 0; r < kMaxRows; ++r) R0[r] = 0.0;
  double* dst;
  const double* tau_base;
  const double m = ((kk & 1) ? -1.0 : 1.0) * ((kk >= 2) ? 2.0 : 1.0);  // this lane's stencil point
  if (phase == 0) {  // A: tau[t-1] = ID(q_t^e, v_t^e, a_{t-1}^e)   (cc:526-531, 763-787)
    pt.dq = m * dq, pt.cv = m * dv, pt.ca = m * da, pt.uv = 1.0, pt.ua = 1.0;
    int* act = bf.act_fd ? bf.act_fd + ((((size_t(b) * T + (t - 1)) * nq + i) * 4 + kk) * dm.np) : nullptr;
    path_eval<kEvalFull>(M, sc, pl, rec_of(t - 1), qB + size_t(t) * nq, vB + size_t(t) * nv, aB + size_t(t - 1) * nv, pt, R0,
                         act);
    dst = bf.dqp + (size_t(b) * T + (t - 1)) * nv * nq;
    tau_base = bf.st.tau + (size_t(b) * T + (t - 1)) * nv;
  } else if (phase == 1) {  // B: tau[t] = ID(q_{t+1}, v_{t+1}^e, a_t^e)   (cc:533-540, 788-814)
    pt.dq = 0.0, pt.cv = -(m * dv), pt.ca = -(m * da), pt.uv = 1.0, pt.ua = 1.0 + 1.0;
    path_eval<kEvalSharedPose>(M, sc, pl, rec_of(t), qB + size_t(t + 1) * nq, vB + size_t(t + 1) * nv, aB + size_t(t) * nv,
                               pt, R0);
    dst = bf.dqt + (size_t(b) * T + t) * nv * nq;
    tau_base = bf.st.tau + (size_t(b) * T + t) * nv;
  } else {  // C: dtau_dqm[t+1] = M(q_{t+2}) N+_{t+1} / dt^2   (cc:552-561)
    pt.dq = 0.0, pt.cv = 0.0, pt.ca = 1.0, pt.uv = 1.0, pt.ua = 1.0;
    path_eval<kEvalSharedPoseNoBias>(M, sc, pl, rec_of(t + 1), qB + size_t(t + 2) * nq, nullptr, nullptr, pt, R0);
    dst = bf.dqm + (size_t(b) * T + (t + 1)) * nv * nq;
    tau_base = nullptr;
  }
  if (NK > 1 && phase < 2) {  // the other stencil points sit in the next NK-1 lanes (same warp: NK divides 32)
    const unsigned mask = __activemask();
#pragma unroll
    for (int r = 0; r < kMaxRows; ++r) {
      R1[r] = __shfl_down_sync(mask, R0[r], 1);
      if (NK > 2) R2[r] = __shfl_down_sync(mask, R0[r], 2), R3[r] = __shfl_down_sync(mask, R0[r], 3);
    }
    if (kk > 0) return;
  }
  // column i of the block: zero except for the affected rows
  double* col = dst + size_t(i) * nv;
  for (int r = 0; r < nv; ++r) col[r] = 0.0;
  int nr = 0;
  auto put = [&](int row) {
    double val;
    if (phase == 2)
      val = 1 / sc.dt / sc.dt * R0[nr];
    else if (METHOD == IDTO_GRAD_FORWARD)
      val = (R0[nr] - tau_base[row]) / dq;  // cc:531, 539
    else if (METHOD == IDTO_GRAD_CENTRAL)
      val = 0.5 * (R0[nr] - R1[nr]) / dq;  // cc:785
    else
      val = 2.0 / 3.0 * (R0[nr] - R1[nr]) / dq - 1.0 / 12.0 * (R2[nr] - R3[nr]) / dq;  // cc:782-783
    col[row] = val;
    ++nr;
  };
  for (int d = 0; d < pl.ndown; ++d) put(M.vs[pl.down[d]]);
  for (int u = 0; u < pl.nup; ++u) {
    const int jt = M.jtype[pl.up[u]], v0 = M.vs[pl.up[u]];
    const int n = jt == IDTO_JOINT_QUAT_FLOATING ? 6 : (jt == IDTO_JOINT_PLANAR ? 3 : 1);
    for (int j = 0; j < n; ++j) put(v0 + j);
  }
}

void launch_partials_path(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool force,
                          cudaStream_t stream) {
  if (dm.npath == 0) return;
  const int nk = sc.method == IDTO_GRAD_CENTRAL4 ? 4 : (sc.method == IDTO_GRAD_CENTRAL ? 2 : 1);
  const long n = long(sc.B) * sc.T * 3 * dm.npath * nk;
  const int grid = int((n + 127) / 128);
  const int smem = model_smem_bytes(dm);
  g_launch_counter += 1;
  switch (sc.method) {
    case IDTO_GRAD_FORWARD: k_partials_path<IDTO_GRAD_FORWARD><<<grid, 128, smem, stream>>>(dm, sc, bf, force ? 1 : 0); break;
    case IDTO_GRAD_CENTRAL: k_partials_path<IDTO_GRAD_CENTRAL><<<grid, 128, smem, stream>>>(dm, sc, bf, force ? 1 : 0); break;
    default: k_partials_path<IDTO_GRAD_CENTRAL4><<<grid, 128, smem, stream>>>(dm, sc, bf, force ? 1 : 0); break;
  }
}

}  // namespace idto
