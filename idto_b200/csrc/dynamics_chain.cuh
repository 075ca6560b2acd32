// Chain-lane inverse dynamics: lane = kinematic chain, step = tree level.
//
// Same physics as dynamics.cuh (CalcInverseDynamicsSingleTimeStep + CalcContactForceContribution,
// optimizer/trajectory_optimizer.cc:228-386) with a different work decomposition: the tree is cut into
// vertex-disjoint root-to-leaf paths ("chains": a body's first child continues its chain, other
// children start new ones); a group of CG lanes (CG = 1, 2, 4, ... >= #chains) evaluates one inverse
// dynamics, lane c walking chain c level by level with the parent's state in registers.  At level l a
// lane holds at most one body, so all lanes of a group do useful work at every step (Mini Cheetah:
// 13 bodies in 4 lanes x 4 levels = 81 % of the lane-steps, against 13 of 64 for one lane per body) and a
// warp carries 32/CG independent evaluations.  Cross-chain parents (branch points) are fetched with
// warp shuffles from the owner lane; the backward (force) pass exchanges per-body totals through a
// small per-evaluation shared-memory stash.
#pragma once
#include "dynamics.cuh"

namespace idto {

struct CModel {  // baked tables in shared memory (superset of SModel)
  SModel M;
  const int *levbody, *levcross, *plane, *gslot, *gdyn, *bchain;
  const double* XWGs;
  int ngb, ngd, nlev;
  int nact, prune;  // pair slots per evaluation; prune: slots hold the compacted ACTIVE pairs, not every candidate
};
__device__ __forceinline__ CModel make_cmodel(const DevModel& dm, const int* si, const double* sd) {
  CModel C;
  C.M = make_smodel(dm, si, sd);
  C.levbody = si + dm.o_levbody, C.levcross = si + dm.o_levcross, C.plane = si + dm.o_plane;
  C.gslot = si + dm.o_gslot, C.gdyn = si + dm.o_gdyn, C.bchain = si + dm.o_bchain;
  C.XWGs = sd + dm.o_XWGs;
  C.ngb = dm.ngb, C.ngd = dm.ngd, C.nlev = dm.nlevels;
  C.nact = dm.nact, C.prune = dm.prune;
  return C;
}

// Pose of one configuration q (body poses, geometry poses, contact-pair geometry).  Private to an
// evaluation, or shared by all evaluations of a (b,t) slot when q is unperturbed.
struct PoseSmem {
  double* RWB;  // [9][nb]
  double* PWB;  // [3][nb]
  double* RWF;  // [9][nb]
  double* GP;   // [12][ngd] world pose of geometries on moving bodies
  double* PG;   // [7][nact]  nhat(3) p_WC(3) fn_c(1); pruned models: + [nact] candidate index of each slot, [1] count
};
// Per-evaluation scratch.
struct EvalSmem {
  double* F;    // [6][nb] spatial force about Bo in W (own, then subtree total)
  double* P;    // [3][nb] p_WB  (copy; the pose may be shared)
  double* AX;   // [3][nb] joint axis in W (1-dof joints)
  double* BV;   // [6][ngb] w, v of geometry-carrying bodies
  double* PF;   // [3][nact] contact force of each pair slot
};
__host__ __device__ inline int cpair_doubles(const DevModel& dm) { return 7 * dm.nact + (dm.prune ? dm.nact + 2 : 0); }
__host__ __device__ inline int cpose_doubles(const DevModel& dm) { return 21 * dm.nb + 12 * (dm.ngd > 0 ? dm.ngd : 1) + cpair_doubles(dm); }
// private pose part of a full evaluation: geometry poses + pair geometry only
__host__ __device__ inline int cpose_private_doubles(const DevModel& dm) { return 12 * (dm.ngd > 0 ? dm.ngd : 1) + cpair_doubles(dm); }
__device__ __forceinline__ PoseSmem make_cpose_private(const DevModel& dm, double* b) {
  const int ngd = dm.ngd > 0 ? dm.ngd : 1;
  return {nullptr, nullptr, nullptr, b, b + 12 * ngd};
}
__host__ __device__ inline int ceval_doubles(const DevModel& dm) { return 12 * dm.nb + 6 * (dm.ngb > 0 ? dm.ngb : 1) + 3 * dm.nact; }
__device__ __forceinline__ PoseSmem make_cpose(const DevModel& dm, double* b) {
  const int ngd = dm.ngd > 0 ? dm.ngd : 1;
  return {b, b + 9 * dm.nb, b + 12 * dm.nb, b + 21 * dm.nb, b + 21 * dm.nb + 12 * ngd};
}
__device__ __forceinline__ EvalSmem make_ceval(const DevModel& dm, double* b) {
  const int ngb = dm.ngb > 0 ? dm.ngb : 1;
  return {b, b + 6 * dm.nb, b + 9 * dm.nb, b + 12 * dm.nb, b + 12 * dm.nb + 6 * ngb};
}

struct BodyState {
  M3 R;
  V3 p, w, v, al, ac;
};
__device__ __forceinline__ double shfl_d(unsigned wm, double x, int src) { return __shfl_sync(wm, x, src); }
__device__ __forceinline__ V3 shfl_v(unsigned wm, V3 a, int src) {
  return {shfl_d(wm, a.x, src), shfl_d(wm, a.y, src), shfl_d(wm, a.z, src)};
}
// lanes of the CG-lane group that contains the calling lane: every warp-level primitive of chain_eval is scoped to
// it, so a group can run (or skip) an evaluation on its own — no lane ever executes an evaluation only to keep a
// neighbour's shuffles alive (those redundant evaluations wrote shared poses twice: racecheck hazards)
template <int CG>
__device__ __forceinline__ unsigned group_mask() {
  if (CG == 32) return 0xffffffffu;
  return ((1u << (CG & 31)) - 1u) << ((threadIdx.x & 31) / CG * CG);
}

// Geometry of candidate pair ip at the pose in `Po` (cc:272-320): signed distance, contact normal (from A into B,
// negated: the direction of the force on B), contact point (midpoint of the witness points).
struct PairGeom {
  double distance;
  V3 nhat, p_WC;
};
__device__ __forceinline__ PairGeom pair_geometry(const CModel& C, const PoseSmem& Po, int ip) {
  const SModel& M = C.M;
  const int gA = M.pA[ip], gB = M.pB[ip];
  M3 R_WGa, R_WGb;
  V3 p_WGa, p_WGb;
  if (M.gbody[gA] >= 0) {
    R_WGa = load_R(Po.GP, C.ngd, C.gdyn[gA]), p_WGa = load_V(Po.GP + 9 * C.ngd, C.ngd, C.gdyn[gA]);
  } else {
    R_WGa = load_R(C.XWGs, M.ng, gA), p_WGa = load_V(C.XWGs + 9 * M.ng, M.ng, gA);
  }
  if (M.gbody[gB] >= 0) {
    R_WGb = load_R(Po.GP, C.ngd, C.gdyn[gB]), p_WGb = load_V(Po.GP + 9 * C.ngd, C.ngd, C.gdyn[gB]);
  } else {
    R_WGb = load_R(C.XWGs, M.ng, gB), p_WGb = load_V(C.XWGs + 9 * M.ng, M.ng, gB);
  }
  const V3 dimA = load_V(M.gdims, M.ng, gA), dimB = load_V(M.gdims, M.ng, gB);
  PairGeom r;
  V3 p_ACa, p_BCb, nhat_BA_W;
  if (M.gtype[gA] == IDTO_GEOM_SPHERE) {
    const PointDist d = point_to_shape(M.gtype[gB], dimB, R_WGb, p_WGb, p_WGa);
    r.distance = d.distance - dimA.x;
    p_BCb = d.p_GN;
    nhat_BA_W = d.grad_W;
    p_ACa = (-dimA.x) * tmul(R_WGa, d.grad_W);
  } else {
    const PointDist d = point_to_shape(M.gtype[gA], dimA, R_WGa, p_WGa, p_WGb);
    r.distance = d.distance - dimB.x;
    p_ACa = d.p_GN;
    nhat_BA_W = -d.grad_W;
    p_BCb = (-dimB.x) * tmul(R_WGb, d.grad_W);
  }
  r.nhat = -nhat_BA_W;
  r.p_WC = 0.5 * ((mul(R_WGa, p_ACa) + p_WGa) + (mul(R_WGb, p_BCb) + p_WGb));
  return r;
}
// candidate index held by slot `is` of a pruned pair list (clamped: the padding groups of a CTA share one
// scratch area, so what they read back may be another group's list)
__device__ __forceinline__ int slot_pair(const double* ids, int is, int np) {
  const int ip = int(ids[is]);
  return ip < 0 ? 0 : (ip < np ? ip : np - 1);
}
// compliant normal force at signed distance d (cc:349-359); 0 outside the activation distance (cc:268-275)
__device__ __forceinline__ double contact_fn_c(const SolverConsts& sc, double distance) {
  if (!(distance <= sc.threshold)) return 0.0;
  const double exponent = -distance / sc.sigma;
  return exponent >= 37 ? -sc.k * distance : sc.sigma * sc.k * log(1 + exp(exponent));
}

// Perturbation of the inputs of one evaluation (finite differencing of column i): applied to the owner
// body's joint only.  q[local] += dq; v += cv * Ncol_v; a += ca * Ncol_a, where a column of N+ is either
// a unit vector at velocity slot `sl` (scaled by uv / ua: 1, or 2 for N+_{t+1} + N+_t) or the quaternion
// block column (nv3 / na3).
struct Perturb {
  int owner;  // body index, -1: none
  int local, sl;
  bool quatcol;
  double dq, cv, ca, uv, ua;
  V3 nv3, na3;
};

enum { kEvalFull = 0, kEvalSharedPose = 1, kEvalSharedPoseNoBias = 2, kEvalPoseOnly = 3 };

// How one evaluation visits the contact pairs (all fields group-uniform).
struct PairWalk {
  const int* near_in = nullptr;  // pruned models: walk this near list instead of every candidate
  int* near_out = nullptr;       // pruned models: write the near list of this (unperturbed) pose
  double margin = 0.0;           // near_out: distance <= threshold + margin
  bool skip = false;             // no contact geometry wanted (pose used by bias-free evaluations only)
  int* act_out = nullptr;        // debug trace (idto_debug_pair_trace): [np] flags, 1 for every candidate pair whose
                                 // force this evaluation computes (the caller zero-fills it); may differ per group
};

// One inverse-dynamics evaluation by a group of CG lanes (all CG lanes of the group must call, converged).
//   MODE kEvalFull:           pose from q (written to `Po`), velocities, forces -> tau
//   MODE kEvalSharedPose:     pose read from `Po` (computed earlier), velocities, forces -> tau
//   MODE kEvalSharedPoseNoBias: tau = M(q) a only (no gravity / damping / contact / velocity terms)
//   MODE kEvalPoseOnly:       pose + contact geometry of q into `Po`, nothing else
// q, v, a: global rows of this (problem, time).  tau_out: [nv] doubles (shared or global).
// STASH: additionally write the per-body record of this evaluation (kStash* layout below) to `stash`
// ([nb][kStashDoubles] doubles in global memory) for the single-lane path evaluations of kernels_path.cu.
constexpr int kStashDoubles = 48;
enum { kStR = 0, kStP = 9, kStRF = 12, kStW = 21, kStV = 24, kStAl = 27, kStAc = 30, kStFpre = 33, kStFtot = 39,
       kStAX = 45 };
__device__ __forceinline__ void stash_V(double* rec, int off, V3 x) { rec[off] = x.x, rec[off + 1] = x.y, rec[off + 2] = x.z; }
__device__ __forceinline__ V3 unstash_V(const double* rec, int off) { return {rec[off], rec[off + 1], rec[off + 2]}; }

template <int CG, int NLEV, int MODE, bool STASH = false>
__device__ __forceinline__ void chain_eval(const CModel& C, const SolverConsts& sc, const PoseSmem& Po,
                                           const EvalSmem& S, int c, const double* __restrict__ q,
                                           const double* __restrict__ v, const double* __restrict__ a,
                                           const Perturb& pt, double* tau_out, double* stash = nullptr,
                                           const PairWalk pw = PairWalk()) {
  const SModel& M = C.M;
  const int nb = M.nb;
  const int gbase = (threadIdx.x & 31) / CG * CG;
  const unsigned wm = group_mask<CG>();
  constexpr bool kPose = MODE == kEvalFull || MODE == kEvalPoseOnly;
  constexpr bool kStorePose = MODE == kEvalPoseOnly;  // body poses are kept only when they will be shared
  constexpr bool kDyn = MODE != kEvalPoseOnly;
  constexpr bool kBias = MODE == kEvalFull || MODE == kEvalSharedPose;
  BodyState prev;
  prev.R = identity3();
  prev.p = prev.w = prev.v = prev.al = prev.ac = {0, 0, 0};

  // (not unrolled: nothing is indexed by the level, and the unrolled body overflows the instruction cache)
#pragma unroll 1
  for (int l = 0; l < NLEV; ++l) {
    const int b = (l < C.nlev && c < CG) ? C.levbody[l * CG + c] : -1;
    BodyState par = prev;
    if (l < C.nlev && C.levcross[l]) {  // group-uniform: some body of this level hangs off another chain
      const int src = (b >= 0 && C.plane[b] >= 0 && C.plane[b] != c) ? gbase + C.plane[b] : (threadIdx.x & 31);
#pragma unroll
      for (int e = 0; e < 9; ++e) par.R.m[e] = shfl_d(wm, prev.R.m[e], src);
      par.p = shfl_v(wm, prev.p, src);
      if (kDyn) {
        par.w = shfl_v(wm, prev.w, src), par.v = shfl_v(wm, prev.v, src);
        par.al = shfl_v(wm, prev.al, src), par.ac = shfl_v(wm, prev.ac, src);
      }
    }
    if (b >= 0) {
      const int parent = M.parent[b], jtype = M.jtype[b], q0 = M.qs[b], v0 = M.vs[b];
      if (parent < 0) {
        par.R = identity3();
        par.p = par.w = par.v = par.al = par.ac = {0, 0, 0};
      }
      const V3 axis = load_V(M.axis, M.nbp, b);
      M3 R_WB, R_WF;
      V3 p_WB;
      if (kPose) {
        double qb[7] = {1, 0, 0, 0, 0, 0, 0};
        const int nqb = joint_nq(jtype);
#pragma unroll
        for (int j = 0; j < 7; ++j)
          if (j < nqb) qb[j] = q[q0 + j];
        if (b == pt.owner) {
#pragma unroll
          for (int j = 0; j < 7; ++j)
            if (j == pt.local) qb[j] += pt.dq;
        }
        const M3 R_PF = load_R(M.XPF, M.nbp, b);
        const V3 p_PF = load_V(M.XPF + 9 * M.nbp, M.nbp, b);
        M3 R_FM = identity3();
        V3 p_FM = {0, 0, 0};
        switch (jtype) {
          case IDTO_JOINT_REVOLUTE: R_FM = axis_angle_R(axis, qb[0]); break;
          case IDTO_JOINT_PRISMATIC: p_FM = qb[0] * axis; break;
          case IDTO_JOINT_PLANAR: {
            double s, cs;
            sincos(qb[2], &s, &cs);
            R_FM = {{cs, -s, 0, s, cs, 0, 0, 0, 1}};
            p_FM = {qb[0], qb[1], 0};
          } break;
          default:
            R_FM = quat_to_R(qb[0], qb[1], qb[2], qb[3]);
            p_FM = {qb[4], qb[5], qb[6]};
            break;
        }
        M3 R_PB = mul(R_PF, R_FM);
        if (!(M.flags[b] & 1)) R_PB = mul(R_PB, load_R(M.RMB, M.nbp, b));
        const V3 p_PB = p_PF + mul(R_PF, p_FM);
        R_WB = mul(par.R, R_PB);
        p_WB = par.p + mul(par.R, p_PB);
        R_WF = mul(par.R, R_PF);
        if (kStorePose) {
#pragma unroll
          for (int e = 0; e < 9; ++e) Po.RWB[e * nb + b] = R_WB.m[e], Po.RWF[e * nb + b] = R_WF.m[e];
          store_V(Po.PWB, nb, b, p_WB);
        }
        // world pose of the geometries carried by this body (scan: few geometries per model)
        if (C.gslot[b] >= 0)
          for (int gi = 0; gi < M.ng; ++gi)
            if (M.gbody[gi] == b) {
              const int gs = C.gdyn[gi];
              const M3 R_WG = mul(R_WB, load_R(M.XBG, M.ng, gi));
              const V3 p_WG = p_WB + mul(R_WB, load_V(M.XBG + 9 * M.ng, M.ng, gi));
#pragma unroll
              for (int e = 0; e < 9; ++e) Po.GP[e * C.ngd + gs] = R_WG.m[e];
              store_V(Po.GP + 9 * C.ngd, C.ngd, gs, p_WG);
            }
      } else {
        R_WB = load_R(Po.RWB, nb, b);
        R_WF = load_R(Po.RWF, nb, b);
        p_WB = load_V(Po.PWB, nb, b);
      }
      prev.R = R_WB, prev.p = p_WB;
      if (kDyn) {
        // joint velocities / accelerations of this body, with the finite-difference perturbation
        double vb[6] = {0, 0, 0, 0, 0, 0}, ab[6] = {0, 0, 0, 0, 0, 0};
        const int nvb = joint_nv(jtype);
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (j < nvb) {
            if (kBias) vb[j] = v[v0 + j];
            ab[j] = a[v0 + j];
          }
        if (b == pt.owner) {
          if (pt.quatcol) {
            vb[0] += pt.cv * pt.nv3.x, vb[1] += pt.cv * pt.nv3.y, vb[2] += pt.cv * pt.nv3.z;
            ab[0] += pt.ca * pt.na3.x, ab[1] += pt.ca * pt.na3.y, ab[2] += pt.ca * pt.na3.z;
          } else {
#pragma unroll
            for (int j = 0; j < 6; ++j)
              if (j == pt.sl) vb[j] += pt.cv * pt.uv, ab[j] += pt.ca * pt.ua;
          }
        }
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
        prev.w = w, prev.v = vv, prev.al = al, prev.ac = ac;
        if (kBias && C.gslot[b] >= 0) {
          store_V(S.BV, C.ngb, C.gslot[b], w);
          store_V(S.BV + 3 * C.ngb, C.ngb, C.gslot[b], vv);
        }
        // own spatial force about Bo in W: inertial - gravity (contact is added in the backward pass)
        const double m = M.mass[b];
        const V3 cm = mul(R_WB, load_V(M.com, M.nbp, b));
        const double* I = M.inertia;
        const M3 IB = {{I[b], I[3 * M.nbp + b], I[4 * M.nbp + b], I[3 * M.nbp + b], I[M.nbp + b],
                        I[5 * M.nbp + b], I[4 * M.nbp + b], I[5 * M.nbp + b], I[2 * M.nbp + b]}};
        const V3 Iw = mul(R_WB, mul(IB, tmul(R_WB, w)));
        const V3 Ial = mul(R_WB, mul(IB, tmul(R_WB, al)));
        V3 f = m * (ac + cross(al, cm) + cross(w, cross(w, cm)));
        V3 t = Ial + cross(w, Iw) + m * cross(cm, ac);
        if (kBias) {
          const V3 fg = m * M.g;
          t = t - cross(cm, fg), f = f - fg;
          // generalized applied forces: joint damping (cc:232); the projection is added in the backward pass
#pragma unroll
          for (int j = 0; j < 6; ++j)
            if (j < nvb) tau_out[v0 + j] = M.damping[v0 + j] * vb[j];
        } else {
#pragma unroll
          for (int j = 0; j < 6; ++j)
            if (j < nvb) tau_out[v0 + j] = 0.0;
        }
        store_V(S.F, nb, b, t);
        store_V(S.F + 3 * nb, nb, b, f);
        store_V(S.P, nb, b, p_WB);
        store_V(S.AX, nb, b, mul(R_WF, axis));
        if (STASH) {
          double* rec = stash + size_t(b) * kStashDoubles;
#pragma unroll
          for (int e = 0; e < 9; ++e) rec[kStR + e] = R_WB.m[e], rec[kStRF + e] = R_WF.m[e];
          stash_V(rec, kStP, p_WB), stash_V(rec, kStW, w), stash_V(rec, kStV, vv);
          stash_V(rec, kStAl, al), stash_V(rec, kStAc, ac), stash_V(rec, kStAX, mul(R_WF, axis));
        }
      }
    }
  }
  __syncwarp(wm);

  // ---- contact pairs: geometry (cc:272-320, 349-359) and forces (cc:322-373), one pair per lane ---------
  // Pair slots: slot == candidate index when every candidate has one (C.prune == 0); otherwise the active pairs
  // (distance <= threshold, exactly the pairs the reference visits, cc:272-275) compacted in candidate order,
  // which keeps the order the forces are summed in (cc:376-384).
  const int pst = C.nact;
  double* const ids = Po.PG + 7 * pst;  // pruned: candidate index of each slot, then the slot count
  if ((kPose || kBias) && M.np > 0) {
    if (kPose) {
      if (C.prune) {
        // candidates of this evaluation: the near list of the unperturbed pose, or all of them
        const bool walk_near = pw.near_in != nullptr;
        const int ncand = pw.skip ? 0 : (walk_near ? min(max(pw.near_in[0], 0), pst) : M.np);
        const int nmax = ncand;  // group-uniform (near_in and skip are)
        const unsigned gmask = CG == 32 ? 0xffffffffu : ((1u << (CG & 31)) - 1u), below = (1u << c) - 1u;
        int count = 0, ncount = 0;
        for (int k0 = 0; k0 < nmax; k0 += CG) {
          const int k = k0 + c;
          int ip = -1;
          if (k < ncand) ip = walk_near ? min(max(pw.near_in[1 + k], 0), M.np - 1) : k;
          PairGeom pg;
          bool act = false, nearby = false;
          if (ip >= 0) {
            pg = pair_geometry(C, Po, ip);
            act = pg.distance <= sc.threshold;
            nearby = pg.distance <= sc.threshold + pw.margin;
          }
          const unsigned grp = (__ballot_sync(wm, act) >> gbase) & gmask;
          const unsigned ngrp = (__ballot_sync(wm, nearby) >> gbase) & gmask;
          const int pos = count + __popc(grp & below), npos = ncount + __popc(ngrp & below);
          if (act && pos < pst) {
            store_V(Po.PG, pst, pos, pg.nhat);
            store_V(Po.PG + 3 * pst, pst, pos, pg.p_WC);
            Po.PG[6 * pst + pos] = contact_fn_c(sc, pg.distance);
            ids[pos] = double(ip);
          }
          if (nearby && pw.near_out && npos < pst) pw.near_out[1 + npos] = ip;
          count += __popc(grp), ncount += __popc(ngrp);
        }
        if (c == 0) {
          ids[pst] = double(count < pst ? count : pst);
          if (pw.near_out) pw.near_out[0] = ncount < pst ? ncount : pst;
          if (count > pst || (pw.near_out && ncount > pst)) atomicExch(sc.status, IDTO_ERR_CONTACT_OVERFLOW);
        }
        __syncwarp(wm);
        if (pw.act_out) {  // what the force pass will walk: the compacted list, read back from shared memory
          const int cnt = min(int(ids[pst]), pst);
          for (int is = c; is < cnt; is += CG) pw.act_out[slot_pair(ids, is, M.np)] = 1;
        }
      } else if (!pw.skip) {
        for (int ip = c; ip < M.np; ip += CG) {
          const PairGeom pg = pair_geometry(C, Po, ip);
          store_V(Po.PG, pst, ip, pg.nhat);
          store_V(Po.PG + 3 * pst, pst, ip, pg.p_WC);
          const double fn_c = contact_fn_c(sc, pg.distance);
          Po.PG[6 * pst + ip] = fn_c;
          if (pw.act_out) pw.act_out[ip] = fn_c > 0.0 ? 1 : 0;  // the force pass skips slots with fn_c == 0
        }
      }
    }
    if (kBias) {
      const int nslot = C.prune ? min(int(ids[pst]), pst) : M.np;
      for (int is = c; is < nslot; is += CG) {
        const int ip = C.prune ? slot_pair(ids, is, M.np) : is;
        const int bA = M.gbody[M.pA[ip]], bB = M.gbody[M.pB[ip]];
        const double fn_c = Po.PG[6 * pst + is];
        V3 f_BC = {0, 0, 0};
        if (fn_c > 0.0) {
          const V3 nhat = load_V(Po.PG, pst, is), p_WC = load_V(Po.PG + 3 * pst, pst, is);
          V3 v_Ac = {0, 0, 0}, v_Bc = {0, 0, 0};
          if (bA >= 0) {
            const int gs = C.gslot[bA];
            v_Ac = load_V(S.BV + 3 * C.ngb, C.ngb, gs) + cross(load_V(S.BV, C.ngb, gs), p_WC - load_V(S.P, nb, bA));
          }
          if (bB >= 0) {
            const int gs = C.gslot[bB];
            v_Bc = load_V(S.BV + 3 * C.ngb, C.ngb, gs) + cross(load_V(S.BV, C.ngb, gs), p_WC - load_V(S.P, nb, bB));
          }
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
          f_BC = fn * nhat + ft_BC;
        }
        store_V(S.PF, pst, is, f_BC);
      }
    }
    __syncwarp(wm);
  }
  if (!kDyn) return;

  // ---- backward pass: subtree totals level by level, projection onto the joint (RNEA inward pass) -------
#pragma unroll 1
  for (int l = NLEV - 1; l >= 0; --l) {
    const int b = (l < C.nlev && c < CG) ? C.levbody[l * CG + c] : -1;
    if (b >= 0) {
      V3 Tt = load_V(S.F, nb, b), Tf = load_V(S.F + 3 * nb, nb, b);
      const V3 p_WB = load_V(S.P, nb, b);
      if (kBias && C.gslot[b] >= 0) {  // contact forces on this body in pair order (cc:376-384)
        V3 Ft = {0, 0, 0}, Ff = {0, 0, 0};
        const int nslot = C.prune ? min(int(ids[pst]), pst) : M.np;
        for (int is = 0; is < nslot; ++is)
          if (Po.PG[6 * pst + is] > 0.0) {
            const int ip = C.prune ? slot_pair(ids, is, M.np) : is;
            const int bA = M.gbody[M.pA[ip]], bB = M.gbody[M.pB[ip]];
            if (bA == b || bB == b) {
              const V3 f = load_V(S.PF, pst, is);
              const V3 pc = load_V(Po.PG + 3 * pst, pst, is) - p_WB;
              if (bA == b) Ft = Ft + cross(pc, -f), Ff = Ff - f;
              if (bB == b) Ft = Ft + cross(pc, f), Ff = Ff + f;
            }
          }
        Tt = Tt - Ft, Tf = Tf - Ff;
      }
      if (STASH) stash_V(stash + size_t(b) * kStashDoubles, kStFpre, Tt), stash_V(stash + size_t(b) * kStashDoubles, kStFpre + 3, Tf);
      const int nchild = M.nchild[b];
      for (int ci = 0; ci < nchild; ++ci) {
        const int ch = M.child[ci * M.nbp + b];
        const V3 tc = load_V(S.F, nb, ch), fc = load_V(S.F + 3 * nb, nb, ch);
        const V3 rc = load_V(S.P, nb, ch) - p_WB;
        Tt = Tt + tc + cross(rc, fc);
        Tf = Tf + fc;
      }
      store_V(S.F, nb, b, Tt);
      store_V(S.F + 3 * nb, nb, b, Tf);
      if (STASH) stash_V(stash + size_t(b) * kStashDoubles, kStFtot, Tt), stash_V(stash + size_t(b) * kStashDoubles, kStFtot + 3, Tf);
      const int jtype = M.jtype[b], v0 = M.vs[b];
      if (jtype == IDTO_JOINT_REVOLUTE) {
        tau_out[v0] += dot(load_V(S.AX, nb, b), Tt);
      } else if (jtype == IDTO_JOINT_PRISMATIC) {
        tau_out[v0] += dot(load_V(S.AX, nb, b), Tf);
      } else {
        // multi-dof joints hang off the world here (checked at model creation): R_WF == R_PF
        const M3 R_WF = (MODE == kEvalFull) ? load_R(M.XPF, M.nbp, b) : load_R(Po.RWF, nb, b);
        const V3 tF = tmul(R_WF, Tt), fF = tmul(R_WF, Tf);
        if (jtype == IDTO_JOINT_PLANAR) {
          tau_out[v0] += fF.x, tau_out[v0 + 1] += fF.y, tau_out[v0 + 2] += tF.z;
        } else {
          tau_out[v0] += tF.x, tau_out[v0 + 1] += tF.y, tau_out[v0 + 2] += tF.z;
          tau_out[v0 + 3] += fF.x, tau_out[v0 + 4] += fF.y, tau_out[v0 + 5] += fF.z;
        }
      }
    }
    __syncwarp(wm);
  }
}

}  // namespace idto
