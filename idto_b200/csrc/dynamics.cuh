// Group-cooperative inverse dynamics with the compliant contact model.
//
// One inverse-dynamics evaluation tau = ID(q, v, a) — the unit of work of the whole hot path
// (reference: CalcInverseDynamicsSingleTimeStep, optimizer/trajectory_optimizer.cc:228-245, and
// CalcContactForceContribution, cc:247-386) — is executed by a group of G lanes of one warp, one
// lane per moving body of the baked tree (G = 2..32, several groups per warp).  The recursion over
// the tree runs level by level; a lane keeps its own body's state in registers and publishes what
// other bodies need (pose, spatial velocity/acceleration, spatial force) in a shared-memory SoA
// slab private to the group.  The baked tables live in shared memory (staged once per CTA by TMA).
//
// An evaluation is split in two phases so that finite differencing can reuse work:
//   PositionPhase(q)      body poses + contact geometry (signed distance, normal, contact point,
//                         compliant normal force) — everything that depends on q only;
//   VelocityPhase(v, a)   spatial velocities/accelerations, contact dissipation + friction, RNEA
//                         forces, projection onto the joint axes.
// Drake conventions restated here are listed in SURVEY.md Appendix B.
#pragma once
#include "solver.h"

namespace idto {

struct SModel {  // baked tables in shared memory
  const int *parent, *jtype, *qs, *vs, *level, *nchild, *child, *flags, *qowner, *gbody, *gtype, *pA, *pB;
  const double *XPF, *RMB, *axis, *mass, *com, *inertia, *damping, *gdims, *XBG;
  int nb, nbp, nq, nv, ng, np, npp, nlevels;
  V3 g;
};

__device__ __forceinline__ SModel make_smodel(const DevModel& dm, const int* si, const double* sd) {
  SModel M;
  M.parent = si + dm.o_parent, M.jtype = si + dm.o_jtype, M.qs = si + dm.o_qs, M.vs = si + dm.o_vs;
  M.level = si + dm.o_level, M.nchild = si + dm.o_nchild, M.child = si + dm.o_child, M.flags = si + dm.o_flags;
  M.qowner = si + dm.o_qowner, M.gbody = si + dm.o_gbody, M.gtype = si + dm.o_gtype;
  M.pA = si + dm.o_pA, M.pB = si + dm.o_pB;
  M.XPF = sd + dm.o_XPF, M.RMB = sd + dm.o_RMB, M.axis = sd + dm.o_axis, M.mass = sd + dm.o_mass;
  M.com = sd + dm.o_com, M.inertia = sd + dm.o_inertia, M.damping = sd + dm.o_damping;
  M.gdims = sd + dm.o_gdims, M.XBG = sd + dm.o_XBG;
  M.nb = dm.nb, M.nbp = dm.nbp, M.nq = dm.nq, M.nv = dm.nv, M.ng = dm.ng, M.np = dm.np, M.npp = dm.npp;
  M.nlevels = dm.nlevels;
  M.g = {dm.gx, dm.gy, dm.gz};
  return M;
}

// Stage both tables into shared memory with two bulk TMA copies on one mbarrier.  All threads call.
__device__ __forceinline__ void stage_model(const DevModel& dm, int* si, double* sd, uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(dm.itab_bytes + dm.dtab_bytes));
    tma_bulk_g2s(si, dm.itab, static_cast<uint32_t>(dm.itab_bytes), bar);
    tma_bulk_g2s(sd, dm.dtab, static_cast<uint32_t>(dm.dtab_bytes), bar);
  }
  mbar_wait(bar, 0);
}

__host__ __device__ inline int model_smem_bytes(const DevModel& dm) { return dm.itab_bytes + dm.dtab_bytes + 16; }
// ------------------------------------------------------------------ mobilizer maps
// Rotation matrix of a possibly non-unit quaternion (2/|q|^2 form).
__device__ __forceinline__ M3 quat_to_R(double w, double x, double y, double z) {
  const double two_over_n2 = 2.0 / (w * w + x * x + y * y + z * z);
  const double sx = two_over_n2 * x, sy = two_over_n2 * y, sz = two_over_n2 * z;
  const double swx = sx * w, swy = sy * w, swz = sz * w;
  const double sxx = sx * x, sxy = sy * x, sxz = sz * x;
  const double syy = sy * y, syz = sz * y, szz = sz * z;
  return {{1 - syy - szz, sxy - swz, sxz + swy, sxy + swz, 1 - sxx - szz, syz - swx, sxz - swy, syz + swx,
           1 - sxx - syy}};
}
__device__ __forceinline__ M3 axis_angle_R(V3 a, double angle) {
  double s, c;
  sincos(angle, &s, &c);
  const V3 sa = s * a, ca = (1.0 - c) * a;
  M3 R;
  double tmp;
  tmp = ca.x * a.y;
  R.m[3] = tmp + sa.z, R.m[1] = tmp - sa.z;
  tmp = ca.x * a.z;
  R.m[6] = tmp - sa.y, R.m[2] = tmp + sa.y;
  tmp = ca.y * a.z;
  R.m[7] = tmp + sa.x, R.m[5] = tmp - sa.x;
  R.m[0] = ca.x * a.x + c, R.m[4] = ca.y * a.y + c, R.m[8] = ca.z * a.z + c;
  return R;
}
// Column `col` (0..3) of the 3x4 block N+(q) = L(2q~)^T (I - q~ q~^T)/|q| of the quaternion
// floating mobilizer (reference call site cc:1645; evaluated at non-unit quaternions, cc:514).
__device__ __forceinline__ V3 quat_nplus_col(const double* q, int col) {
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double qt0 = q[0] / n, qt1 = q[1] / n, qt2 = q[2] / n, qt3 = q[3] / n;
  const double s = 2.0 * qt0, x = 2.0 * qt1, y = 2.0 * qt2, z = 2.0 * qt3;
  const double qc = col == 0 ? qt0 : (col == 1 ? qt1 : (col == 2 ? qt2 : qt3));
  // G[:, col] = (e_col - q~ q~_col)/|q|
  const double g0 = ((col == 0 ? 1.0 : 0.0) - qt0 * qc) / n, g1 = ((col == 1 ? 1.0 : 0.0) - qt1 * qc) / n;
  const double g2 = ((col == 2 ? 1.0 : 0.0) - qt2 * qc) / n, g3 = ((col == 3 ? 1.0 : 0.0) - qt3 * qc) / n;
  // L(2q~)^T rows: [-x s -z y; -y z s -x; -z -y x s]
  V3 r;
  r.x = ((-x * g0 + s * g1) + -z * g2) + y * g3;
  r.y = ((-y * g0 + z * g1) + s * g2) + -x * g3;
  r.z = ((-z * g0 + -y * g1) + x * g2) + s * g3;
  return r;
}

// Velocity of joint k at step t: v_t = N+(q_t)(q_t - q_{t-1})/dt, v_0 = v_init (cc:178-191).  For a
// quaternion joint the 3x4 block of N+ is returned in `col` (cc:1633-1647).
__device__ __forceinline__ void joint_velocity(const SolverConsts& sc, int jt, const double* qt, const double* vinit,
                                               int t, double* vt, V3* col) {
  const int nq = sc.nq;
  if (jt == IDTO_JOINT_QUAT_FLOATING) {
#pragma unroll
    for (int c = 0; c < 4; ++c) col[c] = quat_nplus_col(qt, c);
    if (t == 0) {
      for (int j = 0; j < 6; ++j) vt[j] = vinit[j];
    } else {
      const double* qm = qt - nq;
      V3 acc = {0, 0, 0};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double d = qt[c] - qm[c];
        acc.x += col[c].x * d, acc.y += col[c].y * d, acc.z += col[c].z * d;
      }
      vt[0] = acc.x / sc.dt, vt[1] = acc.y / sc.dt, vt[2] = acc.z / sc.dt;
      for (int j = 0; j < 3; ++j) vt[3 + j] = (qt[4 + j] - qm[4 + j]) / sc.dt;
    }
  } else {
    const int n = jt == IDTO_JOINT_PLANAR ? 3 : 1;
    for (int j = 0; j < n; ++j) vt[j] = t == 0 ? vinit[j] : (qt[j] - qt[j - nq]) / sc.dt;
  }
}


__device__ __forceinline__ void hinge_map(int jtype, V3 axis, const double* x, V3* wF, V3* vF) {
  switch (jtype) {
    case IDTO_JOINT_REVOLUTE: *wF = x[0] * axis, *vF = {0, 0, 0}; break;
    case IDTO_JOINT_PRISMATIC: *wF = {0, 0, 0}, *vF = x[0] * axis; break;
    case IDTO_JOINT_PLANAR: *wF = {0, 0, x[2]}, *vF = {x[0], x[1], 0}; break;
    default: *wF = {x[0], x[1], x[2]}, *vF = {x[3], x[4], x[5]}; break;
  }
}

// ------------------------------------------------------------------ signed distance closed forms
struct PointDist {
  double distance;
  V3 p_GN, grad_W;
};
__device__ __forceinline__ PointDist point_to_sphere(double r, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p_GQ = tmul(R_WG, p_WQ - p_WG);
  const double dist = sqrt(dot(p_GQ, p_GQ));
  const V3 grad_G = dist > 1e-14 ? (1.0 / dist) * p_GQ : V3{1, 0, 0};
  return {dist - r, r * grad_G, mul(R_WG, grad_G)};
}
__device__ __forceinline__ PointDist point_to_box(V3 size, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p = tmul(R_WG, p_WQ - p_WG);
  const double hx = 0.5 * size.x, hy = 0.5 * size.y, hz = 0.5 * size.z;
  V3 pn = {fmin(fmax(p.x, -hx), hx), fmin(fmax(p.y, -hy), hy), fmin(fmax(p.z, -hz), hz)};
  V3 grad;
  if (pn.x != p.x || pn.y != p.y || pn.z != p.z) {
    const V3 d = p - pn;
    const double nrm = sqrt(dot(d, d));
    grad = {d.x / nrm, d.y / nrm, d.z / nrm};
  } else {
    const double dx = hx - fabs(p.x), dy = hy - fabs(p.y), dz = hz - fabs(p.z);
    grad = {0, 0, 0};
    if (dx <= dy && dx <= dz) {
      const double sg = p.x >= 0 ? 1.0 : -1.0;
      pn.x = sg * hx, grad.x = sg;
    } else if (dy <= dz) {
      const double sg = p.y >= 0 ? 1.0 : -1.0;
      pn.y = sg * hy, grad.y = sg;
    } else {
      const double sg = p.z >= 0 ? 1.0 : -1.0;
      pn.z = sg * hz, grad.z = sg;
    }
  }
  const V3 grad_W = mul(R_WG, grad);
  const V3 p_WN = mul(R_WG, pn) + p_WG;
  return {dot(grad_W, p_WQ - p_WN), pn, grad_W};
}
// Capsule (radius r, length L of the cylindrical part, axis Gz): distance to the segment minus r.
__device__ __forceinline__ PointDist point_to_capsule(double r, double L, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p = tmul(R_WG, p_WQ - p_WG);
  const double pz = fmin(fmax(p.z, -0.5 * L), 0.5 * L);
  const V3 d = {p.x, p.y, p.z - pz};
  const double dist = sqrt(dot(d, d));
  const V3 grad_G = dist > 1e-14 ? (1.0 / dist) * d : V3{1, 0, 0};
  const V3 p_GN = V3{0, 0, pz} + r * grad_G;
  return {dist - r, p_GN, mul(R_WG, grad_G)};
}
// Solid cylinder (radius r, length L, axis Gz): the 2-D box rule on (rho, z) with half sizes (r, L/2).
__device__ __forceinline__ PointDist point_to_cylinder(double r, double L, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p = tmul(R_WG, p_WQ - p_WG);
  const double h = 0.5 * L, rho = sqrt(p.x * p.x + p.y * p.y);
  const double ux = rho > 1e-14 ? p.x / rho : 1.0, uy = rho > 1e-14 ? p.y / rho : 0.0;
  double ca, cb, ga, gb;
  if (rho > r || fabs(p.z) > h) {
    ca = fmin(rho, r), cb = fmin(fmax(p.z, -h), h);
    const double da = rho - ca, db = p.z - cb, nrm = sqrt(da * da + db * db);
    ga = da / nrm, gb = db / nrm;
  } else {
    const double da = r - rho, db = h - fabs(p.z);
    if (da <= db) {
      ca = r, cb = p.z, ga = 1.0, gb = 0.0;
    } else {
      const double sg = p.z >= 0 ? 1.0 : -1.0;
      ca = rho, cb = sg * h, ga = 0.0, gb = sg;
    }
  }
  const V3 p_GN = {ca * ux, ca * uy, cb};
  const V3 grad_W = mul(R_WG, V3{ga * ux, ga * uy, gb});
  const V3 p_WN = mul(R_WG, p_GN) + p_WG;
  return {dot(grad_W, p_WQ - p_WN), p_GN, grad_W};
}
// Half space z <= 0 of frame G: the height above the boundary plane; the witness point is the foot of the
// perpendicular, the gradient the plane normal Gz in W.
__device__ __forceinline__ PointDist point_to_half_space(const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p = tmul(R_WG, p_WQ - p_WG);
  return {p.z, V3{p.x, p.y, 0.0}, V3{R_WG.m[2], R_WG.m[5], R_WG.m[8]}};
}
// Signed distance from the query point to the shape `type` (dims as in idto_model_desc::geom_dims).
__device__ __forceinline__ PointDist point_to_shape(int type, V3 dims, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  switch (type) {
    case IDTO_GEOM_SPHERE: return point_to_sphere(dims.x, R_WG, p_WG, p_WQ);
    case IDTO_GEOM_CAPSULE: return point_to_capsule(dims.x, dims.y, R_WG, p_WG, p_WQ);
    case IDTO_GEOM_CYLINDER: return point_to_cylinder(dims.x, dims.y, R_WG, p_WG, p_WQ);
    case IDTO_GEOM_HALF_SPACE: return point_to_half_space(R_WG, p_WG, p_WQ);
    default: return point_to_box(dims, R_WG, p_WG, p_WQ);
  }
}

// ------------------------------------------------------------------ shared-memory slabs
// Position data of one evaluation (SoA, body-minor): written by PositionPhase, read-only afterwards.
// It may be private to a group (perturbed q) or shared by all groups of a (b,t) slot (unperturbed q).
struct PosSmem {
  double* R_WB;  // [9][nbp]
  double* p_WB;  // [3][nbp]
  double* R_WF;  // [9][nbp]
  double* pg;    // [7][npp]: nhat(3) p_WC(3) fn_c(1)
};
// Velocity-level scratch, always private to the group.
struct VelSmem {
  double* wv;  // [12][nbp]: w(3) v(3) alpha(3) acc(3); the total spatial force F(6) aliases alpha/acc
  double* pf;  // [3][npp]: contact force f_BC of each pair
};
__host__ __device__ inline int pos_smem_doubles(const DevModel& dm) { return 21 * dm.nbp + 7 * dm.npp; }
__host__ __device__ inline int vel_smem_doubles(const DevModel& dm) { return 12 * dm.nbp + 3 * dm.npp; }
__device__ __forceinline__ PosSmem make_pos_smem(const DevModel& dm, double* base) {
  return {base, base + 9 * dm.nbp, base + 12 * dm.nbp, base + 21 * dm.nbp};
}
__device__ __forceinline__ VelSmem make_vel_smem(const DevModel& dm, double* base) {
  return {base, base + 12 * dm.nbp};
}

__device__ __forceinline__ M3 load_R(const double* base, int stride, int k) {
  M3 R;
#pragma unroll
  for (int e = 0; e < 9; ++e) R.m[e] = base[e * stride + k];
  return R;
}
__device__ __forceinline__ V3 load_V(const double* base, int stride, int k) {
  return {base[k], base[stride + k], base[2 * stride + k]};
}
__device__ __forceinline__ void store_V(double* base, int stride, int k, V3 v) {
  base[k] = v.x, base[stride + k] = v.y, base[2 * stride + k] = v.z;
}
__device__ __forceinline__ void body_pose(const PosSmem& P, int nbp, int body, M3* R, V3* p) {
  if (body >= 0) {
    *R = load_R(P.R_WB, nbp, body);
    *p = load_V(P.p_WB, nbp, body);
  } else {
    *R = identity3();
    *p = {0, 0, 0};
  }
}

// Position phase.  All lanes of the warp must call (contains __syncwarp); `k` is the lane's body
// index inside its group of G lanes; lanes with k >= nb idle but help with contact pairs.
// qb: the lane's own joint positions (up to 7).  Results go to P.
template <int G>
__device__ __forceinline__ void PositionPhase(const SModel& M, const PosSmem& P, const SolverConsts& sc,
                                              int k, const double* qb) {
  const bool body = k < M.nb;
  constexpr int nbp = G;  // SoA stride == group size (tables and slabs are padded to it)
  int parent = -1, level = -1;
  M3 R_PB = identity3(), R_PF = identity3();
  V3 p_PB = {0, 0, 0};
  if (body) {
    const int jtype = M.jtype[k];
    parent = M.parent[k], level = M.level[k];
    R_PF = load_R(M.XPF, nbp, k);
    const V3 p_PF = load_V(M.XPF + 9 * nbp, nbp, k);
    M3 R_FM = identity3();
    V3 p_FM = {0, 0, 0};
    const V3 axis = load_V(M.axis, nbp, k);
    switch (jtype) {
      case IDTO_JOINT_REVOLUTE: R_FM = axis_angle_R(axis, qb[0]); break;
      case IDTO_JOINT_PRISMATIC: p_FM = qb[0] * axis; break;
      case IDTO_JOINT_PLANAR: {
        double s, c;
        sincos(qb[2], &s, &c);
        R_FM = {{c, -s, 0, s, c, 0, 0, 0, 1}};
        p_FM = {qb[0], qb[1], 0};
      } break;
      default:
        R_FM = quat_to_R(qb[0], qb[1], qb[2], qb[3]);
        p_FM = {qb[4], qb[5], qb[6]};
        break;
    }
    R_PB = mul(R_PF, R_FM);
    if (!(M.flags[k] & 1)) R_PB = mul(R_PB, load_R(M.RMB, nbp, k));
    p_PB = p_PF + mul(R_PF, p_FM);
  }
#pragma unroll 1
  for (int l = 0; l < M.nlevels; ++l) {
    if (body && level == l) {
      M3 R_WP;
      V3 p_WP;
      body_pose(P, nbp, parent, &R_WP, &p_WP);
      const M3 R_WB = mul(R_WP, R_PB);
      const V3 p_WB = p_WP + mul(R_WP, p_PB);
      const M3 R_WF = mul(R_WP, R_PF);
#pragma unroll
      for (int e = 0; e < 9; ++e) P.R_WB[e * nbp + k] = R_WB.m[e], P.R_WF[e * nbp + k] = R_WF.m[e];
      store_V(P.p_WB, nbp, k, p_WB);
    }
    __syncwarp();
  }
  // Contact geometry (cc:272-320 + the position-only part of the force law, cc:349-359).
#pragma unroll 1
  for (int ip = k; ip < M.np; ip += G) {
    const int gA = M.pA[ip], gB = M.pB[ip];
    M3 R_WA, R_WBd;
    V3 p_WA, p_WBd;
    body_pose(P, nbp, M.gbody[gA], &R_WA, &p_WA);
    body_pose(P, nbp, M.gbody[gB], &R_WBd, &p_WBd);
    const M3 R_WGa = mul(R_WA, load_R(M.XBG, M.ng, gA));
    const V3 p_WGa = p_WA + mul(R_WA, load_V(M.XBG + 9 * M.ng, M.ng, gA));
    const M3 R_WGb = mul(R_WBd, load_R(M.XBG, M.ng, gB));
    const V3 p_WGb = p_WBd + mul(R_WBd, load_V(M.XBG + 9 * M.ng, M.ng, gB));
    const V3 dimA = load_V(M.gdims, M.ng, gA), dimB = load_V(M.gdims, M.ng, gB);
    double distance;
    V3 p_ACa, p_BCb, nhat_BA_W;
    if (M.gtype[gA] == IDTO_GEOM_SPHERE) {
      const PointDist d = point_to_shape(M.gtype[gB], dimB, R_WGb, p_WGb, p_WGa);
      distance = d.distance - dimA.x;
      p_BCb = d.p_GN;
      nhat_BA_W = d.grad_W;
      p_ACa = (-dimA.x) * tmul(R_WGa, d.grad_W);
    } else {
      const PointDist d = point_to_shape(M.gtype[gA], dimA, R_WGa, p_WGa, p_WGb);
      distance = d.distance - dimB.x;
      p_ACa = d.p_GN;
      nhat_BA_W = -d.grad_W;
      p_BCb = (-dimB.x) * tmul(R_WGb, d.grad_W);
    }
    double fn_c = 0.0;  // 0 marks "pair not reported by the distance query" (cc:279)
    if (distance <= sc.threshold) {
      const double exponent = -distance / sc.sigma;  // cc:350-359
      fn_c = exponent >= 37 ? -sc.k * distance : sc.sigma * sc.k * log(1 + exp(exponent));
    }
    const V3 nhat = -nhat_BA_W;                                                          // cc:283
    const V3 p_WC = 0.5 * ((mul(R_WGa, p_ACa) + p_WGa) + (mul(R_WGb, p_BCb) + p_WGb));  // cc:309-316
    store_V(P.pg, M.npp, ip, nhat);
    store_V(P.pg + 3 * M.npp, M.npp, ip, p_WC);
    P.pg[6 * M.npp + ip] = fn_c;
  }
  __syncwarp();
}

// Velocity phase: returns the lane's generalized forces tau_b[0..nv_b).  with_bias=false evaluates
// M(q) * a only (no gravity, damping, contact or velocity terms): used for d tau_{t+1}/d q_t =
// M(q_{t+2}) N+_{t+1} / dt^2 (cc:556-561).
template <int G>
__device__ __forceinline__ void VelocityPhase(const SModel& M, const PosSmem& P, const VelSmem& S,
                                              const SolverConsts& sc, int k, const double* vb, const double* ab,
                                              bool with_bias, double* tau_b) {
  const bool body = k < M.nb;
  constexpr int nbp = G;
  int jtype = 0, parent = -1, level = -1;
  V3 axis = {0, 0, 1};
  V3 w = {0, 0, 0}, v = {0, 0, 0}, al = {0, 0, 0}, ac = {0, 0, 0};
  V3 w_rel = {0, 0, 0}, v_rel = {0, 0, 0}, al_rel = {0, 0, 0}, a_rel = {0, 0, 0}, r = {0, 0, 0};
  if (body) {
    jtype = M.jtype[k], parent = M.parent[k], level = M.level[k];
    axis = load_V(M.axis, nbp, k);
    const M3 R_WF = load_R(P.R_WF, nbp, k);
    V3 wF, vF;
    if (with_bias) {
      hinge_map(jtype, axis, vb, &wF, &vF);
      w_rel = mul(R_WF, wF), v_rel = mul(R_WF, vF);
    }
    hinge_map(jtype, axis, ab, &wF, &vF);
    al_rel = mul(R_WF, wF), a_rel = mul(R_WF, vF);
    if (parent >= 0) r = load_V(P.p_WB, nbp, k) - load_V(P.p_WB, nbp, parent);
  }
#pragma unroll 1
  for (int l = 0; l < M.nlevels; ++l) {
    if (body && level == l) {
      if (parent >= 0) {
        const V3 wp = load_V(S.wv, nbp, parent), vp = load_V(S.wv + 3 * nbp, nbp, parent);
        const V3 alp = load_V(S.wv + 6 * nbp, nbp, parent), acp = load_V(S.wv + 9 * nbp, nbp, parent);
        w = wp + w_rel;
        v = vp + cross(wp, r) + v_rel;
        al = alp + cross(wp, w_rel) + al_rel;
        ac = acp + cross(alp, r) + cross(wp, cross(wp, r)) + 2.0 * cross(wp, v_rel) + a_rel;
      } else {
        w = w_rel, v = v_rel, al = al_rel, ac = a_rel;
      }
      store_V(S.wv, nbp, k, w);
      store_V(S.wv + 3 * nbp, nbp, k, v);
      store_V(S.wv + 6 * nbp, nbp, k, al);
      store_V(S.wv + 9 * nbp, nbp, k, ac);
    }
    __syncwarp();
  }
  // Contact forces (velocity-dependent part, cc:322-373), one pair per lane.
  if (with_bias && M.np > 0) {
#pragma unroll 1
    for (int ip = k; ip < M.np; ip += G) {
      const double fn_c = P.pg[6 * M.npp + ip];
      V3 f_BC = {0, 0, 0};
      if (fn_c > 0.0) {
        const int bA = M.gbody[M.pA[ip]], bB = M.gbody[M.pB[ip]];
        const V3 nhat = load_V(P.pg, M.npp, ip), p_WC = load_V(P.pg + 3 * M.npp, M.npp, ip);
        V3 v_Ac = {0, 0, 0}, v_Bc = {0, 0, 0};
        if (bA >= 0)
          v_Ac = load_V(S.wv + 3 * nbp, nbp, bA) + cross(load_V(S.wv, nbp, bA), p_WC - load_V(P.p_WB, nbp, bA));
        if (bB >= 0)
          v_Bc = load_V(S.wv + 3 * nbp, nbp, bB) + cross(load_V(S.wv, nbp, bB), p_WC - load_V(P.p_WB, nbp, bB));
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
      store_V(S.pf, M.npp, ip, f_BC);
    }
    __syncwarp();
  }
  V3 Tt = {0, 0, 0}, Tf = {0, 0, 0};
  V3 p_WB = {0, 0, 0};
  int nchild = 0;
  if (body) {
    // applied forces: gravity (cc:232) + contact in pair order (cc:376-384)
    V3 Ft = {0, 0, 0}, Ff = {0, 0, 0};
    const double m = M.mass[k];
    const M3 R_WB = load_R(P.R_WB, nbp, k);
    p_WB = load_V(P.p_WB, nbp, k);
    const V3 c = mul(R_WB, load_V(M.com, nbp, k));
    if (with_bias) {
      const V3 fg = m * M.g;
      Ft = cross(c, fg), Ff = fg;
#pragma unroll 1
      for (int ip = 0; ip < M.np; ++ip) {
        if (P.pg[6 * M.npp + ip] > 0.0) {
          const int bA = M.gbody[M.pA[ip]], bB = M.gbody[M.pB[ip]];
          if (bA == k || bB == k) {
            const V3 f = load_V(S.pf, M.npp, ip);
            const V3 pc = load_V(P.pg + 3 * M.npp, M.npp, ip) - p_WB;
            if (bA == k) Ft = Ft + cross(pc, -f), Ff = Ff - f;
            if (bB == k) Ft = Ft + cross(pc, f), Ff = Ff + f;
          }
        }
      }
    }
    // inertial forces about Bo in W
    const double* I = M.inertia;
    const M3 IB = {{I[k], I[3 * nbp + k], I[4 * nbp + k], I[3 * nbp + k], I[nbp + k], I[5 * nbp + k],
                    I[4 * nbp + k], I[5 * nbp + k], I[2 * nbp + k]}};
    const V3 Iw = mul(R_WB, mul(IB, tmul(R_WB, w)));
    const V3 Ial = mul(R_WB, mul(IB, tmul(R_WB, al)));
    const V3 f = m * (ac + cross(al, c) + cross(w, cross(w, c)));
    const V3 t = Ial + cross(w, Iw) + m * cross(c, ac);
    Tt = t - Ft, Tf = f - Ff;
    nchild = M.nchild[k];
  }
  __syncwarp();  // every lane is done reading alpha/acc of its parent: F may now alias them
  if (body) {
    store_V(S.wv + 6 * nbp, nbp, k, Tt);
    store_V(S.wv + 9 * nbp, nbp, k, Tf);
  }
  __syncwarp();
#pragma unroll 1
  for (int l = M.nlevels - 2; l >= 0; --l) {
    if (body && level == l && nchild > 0) {
      for (int ci = 0; ci < nchild; ++ci) {
        const int c = M.child[ci * nbp + k];
        const V3 tc = load_V(S.wv + 6 * nbp, nbp, c), fc = load_V(S.wv + 9 * nbp, nbp, c);
        const V3 rc = load_V(P.p_WB, nbp, c) - p_WB;
        Tt = Tt + tc + cross(rc, fc);
        Tf = Tf + fc;
      }
      store_V(S.wv + 6 * nbp, nbp, k, Tt);
      store_V(S.wv + 9 * nbp, nbp, k, Tf);
    }
    __syncwarp();
  }
  if (body) {
    const M3 R_WF = load_R(P.R_WF, nbp, k);
    const V3 tF = tmul(R_WF, Tt), fF = tmul(R_WF, Tf);
    switch (jtype) {
      case IDTO_JOINT_REVOLUTE: tau_b[0] = dot(axis, tF); break;
      case IDTO_JOINT_PRISMATIC: tau_b[0] = dot(axis, fF); break;
      case IDTO_JOINT_PLANAR: tau_b[0] = fF.x, tau_b[1] = fF.y, tau_b[2] = tF.z; break;
      default: tau_b[0] = tF.x, tau_b[1] = tF.y, tau_b[2] = tF.z, tau_b[3] = fF.x, tau_b[4] = fF.y, tau_b[5] = fF.z; break;
    }
    if (with_bias) {
      const int vs = M.vs[k];
      const int nvb = jtype == IDTO_JOINT_QUAT_FLOATING ? 6 : (jtype == IDTO_JOINT_PLANAR ? 3 : 1);
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (j < nvb) tau_b[j] += M.damping[vs + j] * vb[j];
    }
  }
}

__device__ __forceinline__ int joint_nq(int jtype) {
  return jtype == IDTO_JOINT_QUAT_FLOATING ? 7 : (jtype == IDTO_JOINT_PLANAR ? 3 : 1);
}
__device__ __forceinline__ int joint_nv(int jtype) {
  return jtype == IDTO_JOINT_QUAT_FLOATING ? 6 : (jtype == IDTO_JOINT_PLANAR ? 3 : 1);
}

}  // namespace idto
