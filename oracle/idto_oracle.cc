// =============================================================================
// oracle/idto_oracle.cc — CPU restatement of IDTO's Gauss-Newton hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product
// (idto_b200/csrc) never links, imports or calls it.
//
// What it is: a plain C++17 (no Eigen, no Drake) restatement of the reference
// algorithm, function by function, each citing the reference file:line it
// follows ("cc" = optimizer/trajectory_optimizer.cc, paths relative to the
// reference root).  The reference itself cannot be compiled here: every
// optimizer header includes Drake v1.30.0 (README.md:77-78), which is neither
// vendored nor installed, and Eigen is absent.
//
// Parity status:
//   * IDTO-side arithmetic (cost, gradient, Hessian, scaling, penta-diagonal
//     factorisation, Lagrange multipliers, dogleg, trust ratio, TR loop) is
//     PINNED by the reference's own known-answer tests, re-run against this
//     file in tests/test_oracle_kat.py (pendulum closed forms, CalcCost golden
//     vector, penta-diagonal properties, spinner end-to-end q_T golden).
//   * The Drake boundary (RNEA for multi-body trees, planar/quaternion
//     mobilizers, N+(q) at non-unit quaternions, signed-distance witness points,
//     pair ordering): PARITY UNPINNED except through the spinner end-to-end
//     golden (python_bindings/test/trajectory_optimizer_test.py:84-85) and the
//     pendulum closed forms.  It is a restatement of Drake's published
//     algorithms plus physical self-checks (tests/test_oracle_physics.py).
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#endif

#include "../include/idto_b200.h"

namespace {

constexpr int kMaxBodies = 64;
constexpr int kMaxPairs = 256;

// ----------------------------------------------------------------------------- small 3-vector algebra
struct V3 {
  double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
struct M3 {
  double m[9];  // row-major
};
inline V3 operator*(const M3& R, V3 v) {
  return {R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z,
          R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z};
}
inline V3 tmul(const M3& R, V3 v) {  // R^T v
  return {R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z,
          R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z};
}
inline M3 operator*(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
inline M3 identity3() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }

// ----------------------------------------------------------------------------- model (deep copy of the desc)
struct Model {
  int nb, nq, nv;
  std::vector<int> parent, jtype, qs, vs, actuated;
  std::vector<M3> R_PF, R_MB;
  std::vector<V3> p_PF, axis, com;
  std::vector<double> damping, mass, inertia;  // inertia [nb][6]
  V3 g;
  int ng, np;
  std::vector<int> gbody, gtype, pA, pB;
  std::vector<V3> gdims, p_BG;
  std::vector<M3> R_BG;
  std::vector<int> unactuated;
  std::vector<int> quat_starts;

  explicit Model(const idto_model_desc& d) {
    nb = d.nbodies, nq = d.nq, nv = d.nv;
    parent.assign(d.parent, d.parent + nb);
    jtype.assign(d.joint_type, d.joint_type + nb);
    qs.assign(d.q_start, d.q_start + nb);
    vs.assign(d.v_start, d.v_start + nb);
    actuated.assign(d.actuated, d.actuated + nv);
    damping.assign(d.damping, d.damping + nv);
    mass.assign(d.mass, d.mass + nb);
    inertia.assign(d.inertia, d.inertia + 6 * nb);
    R_PF.resize(nb), R_MB.resize(nb), p_PF.resize(nb), axis.resize(nb), com.resize(nb);
    for (int k = 0; k < nb; ++k) {
      std::memcpy(R_PF[k].m, d.X_PF + 12 * k, 9 * sizeof(double));
      p_PF[k] = {d.X_PF[12 * k + 9], d.X_PF[12 * k + 10], d.X_PF[12 * k + 11]};
      std::memcpy(R_MB[k].m, d.R_MB + 9 * k, 9 * sizeof(double));
      axis[k] = {d.axis[3 * k], d.axis[3 * k + 1], d.axis[3 * k + 2]};
      com[k] = {d.com[3 * k], d.com[3 * k + 1], d.com[3 * k + 2]};
      if (jtype[k] == IDTO_JOINT_QUAT_FLOATING) quat_starts.push_back(qs[k]);
    }
    g = {d.gravity[0], d.gravity[1], d.gravity[2]};
    ng = d.ngeoms, np = d.npairs;
    gbody.assign(d.geom_body, d.geom_body + ng);
    gtype.assign(d.geom_type, d.geom_type + ng);
    gdims.resize(ng), p_BG.resize(ng), R_BG.resize(ng);
    for (int k = 0; k < ng; ++k) {
      gdims[k] = {d.geom_dims[3 * k], d.geom_dims[3 * k + 1], d.geom_dims[3 * k + 2]};
      std::memcpy(R_BG[k].m, d.X_BG + 12 * k, 9 * sizeof(double));
      p_BG[k] = {d.X_BG[12 * k + 9], d.X_BG[12 * k + 10], d.X_BG[12 * k + 11]};
    }
    pA.assign(d.pair_geomA, d.pair_geomA + np);
    pB.assign(d.pair_geomB, d.pair_geomB + np);
    // cc:63-72: unactuated dofs = rows of B summing to zero; if B is empty, fully actuated.
    bool any = false;
    for (int i = 0; i < nv; ++i) any |= (actuated[i] != 0);
    if (any)
      for (int i = 0; i < nv; ++i)
        if (!actuated[i]) unactuated.push_back(i);
  }
};

struct ContactParams {
  double k, sigma, vd, vs, mu;
};

// ----------------------------------------------------------------------------- mobilizer kinematics (Drake conventions, Appendix B)
// Rotation matrix of a possibly non-unit quaternion (Drake RotationMatrix(Quaternion): 2/|q|^2 form).
inline M3 quat_to_R(double w, double x, double y, double z) {
  const double two_over_n2 = 2.0 / (w * w + x * x + y * y + z * z);
  const double sx = two_over_n2 * x, sy = two_over_n2 * y, sz = two_over_n2 * z;
  const double swx = sx * w, swy = sy * w, swz = sz * w;
  const double sxx = sx * x, sxy = sy * x, sxz = sz * x;
  const double syy = sy * y, syz = sz * y, szz = sz * z;
  return {{1 - syy - szz, sxy - swz, sxz + swy, sxy + swz, 1 - sxx - szz, syz - swx, sxz - swy,
           syz + swx, 1 - sxx - syy}};
}

// Rotation by `angle` about unit `a` (Eigen AngleAxis::toRotationMatrix operation order).
inline M3 axis_angle_R(V3 a, double angle) {
  const double s = std::sin(angle), c = std::cos(angle);
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

// N+(q) 3x4 block of the quaternion floating mobilizer, evaluated at a possibly non-unit
// quaternion: N+ = L(2 q~)^T (I - q~ q~^T)/|q|  (Drake QuaternionFloatingMobilizer::
// QuaternionRateToAngularVelocityMatrix; SURVEY.md App. B "highest-risk", parity unpinned).
inline void quat_nplus(const double* q, double N[3][4]) {
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double qt[4] = {q[0] / n, q[1] / n, q[2] / n, q[3] / n};
  const double s = 2.0 * qt[0], x = 2.0 * qt[1], y = 2.0 * qt[2], z = 2.0 * qt[3];
  // L(2q~)^T, rows = angular velocity components, cols = (w,x,y,z)
  const double LT[3][4] = {{-x, s, -z, y}, {-y, z, s, -x}, {-z, -y, x, s}};
  double G[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) G[i][j] = ((i == j ? 1.0 : 0.0) - qt[i] * qt[j]) / n;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      double acc = 0;
      for (int k = 0; k < 4; ++k) acc += LT[i][k] * G[k][j];
      N[i][j] = acc;
    }
}

// ----------------------------------------------------------------------------- signed distance (Drake closed forms)
struct DistResult {
  double distance;
  V3 p_ACa, p_BCb;  // witness points in the geometry frames
  V3 nhat_BA_W;
};

struct PointDist {
  double distance;
  V3 p_GN;    // nearest point on dG in G
  V3 grad_W;  // gradient of the distance field at Q, in W
};

// Drake point_distance::DistanceToPoint for a sphere G at X_WG, query point p_WQ.
inline PointDist point_to_sphere(double r, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p_GQ = tmul(R_WG, p_WQ - p_WG);
  const double dist = std::sqrt(dot(p_GQ, p_GQ));
  const double tol = 1e-14;  // below this the direction is arbitrary: use Gx
  const V3 grad_G = dist > tol ? (1.0 / dist) * p_GQ : V3{1, 0, 0};
  return {dist - r, r * grad_G, R_WG * grad_G};
}

// Drake point_distance::DistanceToPoint for a box G (full sizes `size`).
inline PointDist point_to_box(V3 size, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p = tmul(R_WG, p_WQ - p_WG);
  const double h[3] = {0.5 * size.x, 0.5 * size.y, 0.5 * size.z};
  const double pq[3] = {p.x, p.y, p.z};
  double pn[3], grad[3] = {0, 0, 0};
  bool outside = false;
  for (int i = 0; i < 3; ++i) {
    pn[i] = std::min(std::max(pq[i], -h[i]), h[i]);
    if (pn[i] != pq[i]) outside = true;
  }
  if (outside) {
    double d[3] = {pq[0] - pn[0], pq[1] - pn[1], pq[2] - pn[2]};
    const double nrm = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int i = 0; i < 3; ++i) grad[i] = d[i] / nrm;
  } else {
    // inside (or on the boundary): nearest face
    int axis = 0;
    double best = std::numeric_limits<double>::infinity();
    for (int i = 0; i < 3; ++i) {
      const double di = h[i] - std::fabs(pq[i]);
      if (di < best) best = di, axis = i;
    }
    const double sgn = pq[axis] >= 0 ? 1.0 : -1.0;
    pn[axis] = sgn * h[axis];
    grad[axis] = sgn;
  }
  const V3 p_GN = {pn[0], pn[1], pn[2]};
  const V3 grad_W = R_WG * V3{grad[0], grad[1], grad[2]};
  const V3 p_WN = R_WG * p_GN + p_WG;
  return {dot(grad_W, p_WQ - p_WN), p_GN, grad_W};
}

// Drake point_distance::DistanceToPoint for a capsule G (radius r, length L of the cylindrical part, axis Gz):
// distance to the segment [-L/2, L/2] on Gz minus r.  Not used by a BASELINE config; parity unpinned.
inline PointDist point_to_capsule(double r, double L, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p = tmul(R_WG, p_WQ - p_WG);
  const double pz = std::min(std::max(p.z, -0.5 * L), 0.5 * L);
  const V3 d = {p.x, p.y, p.z - pz};
  const double dist = std::sqrt(dot(d, d));
  const V3 grad_G = dist > 1e-14 ? (1.0 / dist) * d : V3{1, 0, 0};
  const V3 p_GN = V3{0, 0, pz} + r * grad_G;
  return {dist - r, p_GN, R_WG * grad_G};
}

// Drake point_distance::DistanceToPoint for a solid cylinder G (radius r, length L, axis Gz): the 2-D box
// rule on (rho, z) with half sizes (r, L/2); inside, the nearer of the side and the caps (side first on ties,
// like the first-axis rule of the box); on the axis the radial direction is Gx.
inline PointDist point_to_cylinder(double r, double L, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p = tmul(R_WG, p_WQ - p_WG);
  const double h = 0.5 * L, rho = std::sqrt(p.x * p.x + p.y * p.y);
  const double ux = rho > 1e-14 ? p.x / rho : 1.0, uy = rho > 1e-14 ? p.y / rho : 0.0;
  double ca, cb, ga, gb;
  if (rho > r || std::fabs(p.z) > h) {
    ca = std::min(rho, r), cb = std::min(std::max(p.z, -h), h);
    const double da = rho - ca, db = p.z - cb, nrm = std::sqrt(da * da + db * db);
    ga = da / nrm, gb = db / nrm;
  } else {
    const double da = r - rho, db = h - std::fabs(p.z);
    if (da <= db) {
      ca = r, cb = p.z, ga = 1.0, gb = 0.0;
    } else {
      const double sg = p.z >= 0 ? 1.0 : -1.0;
      ca = rho, cb = sg * h, ga = 0.0, gb = sg;
    }
  }
  const V3 p_GN = {ca * ux, ca * uy, cb};
  const V3 grad_W = R_WG * V3{ga * ux, ga * uy, gb};
  const V3 p_WN = R_WG * p_GN + p_WG;
  return {dot(grad_W, p_WQ - p_WN), p_GN, grad_W};
}

// drake::geometry::HalfSpace: the region z <= 0 of its frame, so the signed distance is the z coordinate of the
// query point in G, the nearest point its projection onto the plane z = 0, the gradient the normal Gz.
inline PointDist point_to_half_space(const M3& R_WG, V3 p_WG, V3 p_WQ) {
  const V3 p = tmul(R_WG, p_WQ - p_WG);
  const V3 normal_W = R_WG * V3{0.0, 0.0, 1.0};
  return {p.z, V3{p.x, p.y, 0.0}, normal_W};
}

// Signed distance from the query point to the shape `type` (dims as in idto_model_desc::geom_dims).
inline PointDist point_to_shape(int type, V3 dims, const M3& R_WG, V3 p_WG, V3 p_WQ) {
  switch (type) {
    case IDTO_GEOM_SPHERE: return point_to_sphere(dims.x, R_WG, p_WG, p_WQ);
    case IDTO_GEOM_CAPSULE: return point_to_capsule(dims.x, dims.y, R_WG, p_WG, p_WQ);
    case IDTO_GEOM_CYLINDER: return point_to_cylinder(dims.x, dims.y, R_WG, p_WG, p_WQ);
    case IDTO_GEOM_HALF_SPACE: return point_to_half_space(R_WG, p_WG, p_WQ);
    default: return point_to_box(dims, R_WG, p_WG, p_WQ);
  }
}

// ----------------------------------------------------------------------------- inverse dynamics workspace
struct Kin {
  M3 R_WB[kMaxBodies], R_WF[kMaxBodies];
  V3 p_WB[kMaxBodies];
  V3 w[kMaxBodies], v[kMaxBodies];          // V_WB
  V3 alpha[kMaxBodies], acc[kMaxBodies];    // A_WB
  V3 w_rel[kMaxBodies], v_rel[kMaxBodies];  // across-mobilizer velocity in W
  V3 Ft[kMaxBodies], Ff[kMaxBodies];        // applied spatial forces at Bo, in W (torque, force)
  V3 Tt[kMaxBodies], Tf[kMaxBodies];        // total (inertial - applied + children)
};

struct Optimizer;

// Position kinematics for all moving bodies (Drake: EvalBodyPoseInWorld).
static void PositionKinematics(const Model& M, const double* q, Kin* K) {
  for (int k = 0; k < M.nb; ++k) {
    const int p = M.parent[k];
    M3 R_WP = identity3();
    V3 p_WP = {0, 0, 0};
    if (p >= 0) R_WP = K->R_WB[p], p_WP = K->p_WB[p];
    const M3 R_WF = R_WP * M.R_PF[k];
    const V3 p_WF = p_WP + R_WP * M.p_PF[k];
    M3 R_FM = identity3();
    V3 p_FM = {0, 0, 0};
    const double* qb = q + M.qs[k];
    switch (M.jtype[k]) {
      case IDTO_JOINT_REVOLUTE: R_FM = axis_angle_R(M.axis[k], qb[0]); break;
      case IDTO_JOINT_PRISMATIC: p_FM = qb[0] * M.axis[k]; break;
      case IDTO_JOINT_PLANAR: {
        const double s = std::sin(qb[2]), c = std::cos(qb[2]);
        R_FM = {{c, -s, 0, s, c, 0, 0, 0, 1}};
        p_FM = {qb[0], qb[1], 0};
      } break;
      case IDTO_JOINT_QUAT_FLOATING:
        R_FM = quat_to_R(qb[0], qb[1], qb[2], qb[3]);
        p_FM = {qb[4], qb[5], qb[6]};
        break;
    }
    K->R_WF[k] = R_WF;
    K->R_WB[k] = (R_WF * R_FM) * M.R_MB[k];
    K->p_WB[k] = p_WF + R_WF * p_FM;
  }
}

// Across-mobilizer spatial velocity / acceleration in F: V_FM_F = H_F * x, x in R^{nv_b}.
inline void HingeMap(const Model& M, int k, const double* x, V3* w_F, V3* v_F) {
  switch (M.jtype[k]) {
    case IDTO_JOINT_REVOLUTE: *w_F = x[0] * M.axis[k], *v_F = {0, 0, 0}; break;
    case IDTO_JOINT_PRISMATIC: *w_F = {0, 0, 0}, *v_F = x[0] * M.axis[k]; break;
    case IDTO_JOINT_PLANAR: *w_F = {0, 0, x[2]}, *v_F = {x[0], x[1], 0}; break;
    default: *w_F = {x[0], x[1], x[2]}, *v_F = {x[3], x[4], x[5]}; break;
  }
}

static void VelocityKinematics(const Model& M, const double* v, Kin* K) {
  for (int k = 0; k < M.nb; ++k) {
    const int p = M.parent[k];
    V3 wF, vF;
    HingeMap(M, k, v + M.vs[k], &wF, &vF);
    K->w_rel[k] = K->R_WF[k] * wF;
    K->v_rel[k] = K->R_WF[k] * vF;
    if (p >= 0) {
      const V3 r = K->p_WB[k] - K->p_WB[p];
      K->w[k] = K->w[p] + K->w_rel[k];
      K->v[k] = K->v[p] + cross(K->w[p], r) + K->v_rel[k];
    } else {
      K->w[k] = K->w_rel[k];
      K->v[k] = K->v_rel[k];
    }
  }
}

// cc:247-386 CalcContactForceContribution: adds contact spatial forces to K->Ft/Ff.
static void ContactForces(const Model& M, const ContactParams& cp, Kin* K, int* pair_active) {
  const double k = cp.k, sigma = cp.sigma, dissipation_velocity = cp.vd, vs = cp.vs, mu = cp.mu;
  // cc:266-269: distance beyond which contact forces vanish.
  const double eps = std::sqrt(std::numeric_limits<double>::epsilon());
  const double threshold = -sigma * std::log(std::exp(eps / (sigma * k)) - 1.0);
  for (int ip = 0; ip < M.np; ++ip) {
    const int gA = M.pA[ip], gB = M.pB[ip];
    const int bA = M.gbody[gA], bB = M.gbody[gB];
    const M3 R_WA = bA >= 0 ? K->R_WB[bA] : identity3();
    const V3 p_WA = bA >= 0 ? K->p_WB[bA] : V3{0, 0, 0};
    const M3 R_WB = bB >= 0 ? K->R_WB[bB] : identity3();
    const V3 p_WB = bB >= 0 ? K->p_WB[bB] : V3{0, 0, 0};
    // cc:301-312 geometry poses in world
    const M3 R_WGa = R_WA * M.R_BG[gA];
    const V3 p_WGa = p_WA + R_WA * M.p_BG[gA];
    const M3 R_WGb = R_WB * M.R_BG[gB];
    const V3 p_WGb = p_WB + R_WB * M.p_BG[gB];

    // ComputeSignedDistancePairwiseClosestPoints (Drake closed forms; cc:279).
    DistResult pr;
    if (M.gtype[gA] == IDTO_GEOM_SPHERE) {
      // sphere A vs shape B: distance from B to A's centre, minus rA.
      PointDist d = point_to_shape(M.gtype[gB], M.gdims[gB], R_WGb, p_WGb, p_WGa);
      const double rA = M.gdims[gA].x;
      pr.distance = d.distance - rA;
      pr.p_BCb = d.p_GN;
      pr.nhat_BA_W = d.grad_W;
      pr.p_ACa = (-rA) * tmul(R_WGa, d.grad_W);
    } else {
      // shape A (box, capsule, cylinder) vs sphere B: roles swapped, then results swapped back.
      PointDist d = point_to_shape(M.gtype[gA], M.gdims[gA], R_WGa, p_WGa, p_WGb);
      const double rB = M.gdims[gB].x;
      pr.distance = d.distance - rB;
      pr.p_ACa = d.p_GN;
      pr.nhat_BA_W = -d.grad_W;
      pr.p_BCb = (-rB) * tmul(R_WGb, d.grad_W);
    }
    if (pair_active) pair_active[ip] = pr.distance <= threshold;
    if (!(pr.distance <= threshold)) continue;  // pair not reported by the query (cc:279)

    const V3 nhat = -pr.nhat_BA_W;  // cc:283
    const V3 p_WCa = R_WGa * pr.p_ACa + p_WGa;  // cc:309
    const V3 p_WCb = R_WGb * pr.p_BCb + p_WGb;  // cc:312
    const V3 p_WC = 0.5 * (p_WCa + p_WCb);      // cc:316
    const V3 p_AC = p_WC - p_WA, p_BC = p_WC - p_WB;  // cc:319-320
    const V3 wA = bA >= 0 ? K->w[bA] : V3{0, 0, 0}, vA = bA >= 0 ? K->v[bA] : V3{0, 0, 0};
    const V3 wB = bB >= 0 ? K->w[bB] : V3{0, 0, 0}, vB = bB >= 0 ? K->v[bB] : V3{0, 0, 0};
    const V3 v_Ac = vA + cross(wA, p_AC), v_Bc = vB + cross(wB, p_BC);  // cc:327-328
    const V3 v_AcBc = v_Bc - v_Ac;                                       // cc:331-332
    const double vn = dot(nhat, v_AcBc);                                 // cc:335
    const V3 vt = v_AcBc - vn * nhat;                                    // cc:336
    double dissipation_factor = 0.0;                                     // cc:339-345
    const double s = vn / dissipation_velocity;
    if (s < 0) {
      dissipation_factor = 1 - s;
    } else if (s < 2) {
      dissipation_factor = (s - 2) * (s - 2) / 4;
    }
    double compliant_fn;  // cc:349-359
    const double exponent = -pr.distance / sigma;
    if (exponent >= 37) {
      compliant_fn = -k * pr.distance;
    } else {
      compliant_fn = sigma * k * std::log(1 + std::exp(exponent));
    }
    const double fn = compliant_fn * dissipation_factor;  // cc:360
    const V3 that_regularized = (-1.0 / std::sqrt(vs * vs + dot(vt, vt))) * vt;  // cc:368-369
    const V3 ft_BC = (mu * fn) * that_regularized;                               // cc:370
    const V3 f_BC = fn * nhat + ft_BC;                                           // cc:373
    // cc:376-384: shift to body origins and accumulate (A gets -f, B gets +f).
    if (bA >= 0) {
      K->Ft[bA] = K->Ft[bA] + cross(p_AC, -f_BC);
      K->Ff[bA] = K->Ff[bA] - f_BC;
    }
    if (bB >= 0) {
      K->Ft[bB] = K->Ft[bB] + cross(p_BC, f_BC);
      K->Ff[bB] = K->Ff[bB] + f_BC;
    }
  }
}

// cc:228-245 CalcInverseDynamicsSingleTimeStep: tau = M(q) a + C(q,v) v - tau_app - J^T F_app with
// F_app = force elements (gravity, joint damping; cc:232) + contact (cc:240).
// `with_bias=false` drops gravity/damping/contact and velocity terms (used for M(q)*a products).
static void InverseDynamics(const Model& M, const ContactParams& cp, const double* q, const double* v,
                            const double* a, double* tau, bool with_bias = true,
                            int* pair_active = nullptr) {
  Kin K;
  PositionKinematics(M, q, &K);
  std::vector<double> vz;
  if (!with_bias) {
    vz.assign(M.nv, 0.0);
    v = vz.data();
  }
  VelocityKinematics(M, v, &K);
  // accelerations (RNEA outward pass)
  for (int k = 0; k < M.nb; ++k) {
    const int p = M.parent[k];
    V3 aF_w, aF_v;
    HingeMap(M, k, a + M.vs[k], &aF_w, &aF_v);
    const V3 al_rel = K.R_WF[k] * aF_w, a_rel = K.R_WF[k] * aF_v;
    if (p >= 0) {
      const V3 r = K.p_WB[k] - K.p_WB[p];
      const V3 wp = K.w[p];
      K.alpha[k] = K.alpha[p] + cross(wp, K.w_rel[k]) + al_rel;
      K.acc[k] = K.acc[p] + cross(K.alpha[p], r) + cross(wp, cross(wp, r)) +
                 2.0 * cross(wp, K.v_rel[k]) + a_rel;
    } else {
      K.alpha[k] = al_rel;
      K.acc[k] = a_rel;
    }
  }
  // applied forces: cc:232 CalcForceElementsContribution (gravity as body forces at Bo)
  for (int k = 0; k < M.nb; ++k) {
    K.Ft[k] = {0, 0, 0}, K.Ff[k] = {0, 0, 0};
    if (with_bias) {
      const V3 c = K.R_WB[k] * M.com[k];
      const V3 fg = M.mass[k] * M.g;
      K.Ft[k] = cross(c, fg);
      K.Ff[k] = fg;
    }
  }
  if (with_bias && M.np > 0) ContactForces(M, cp, &K, pair_active);
  // body inertial forces about Bo, world frame
  for (int k = 0; k < M.nb; ++k) {
    const M3& R = K.R_WB[k];
    const double* I = &M.inertia[6 * k];
    const M3 IB = {{I[0], I[3], I[4], I[3], I[1], I[5], I[4], I[5], I[2]}};
    const V3 c = R * M.com[k];
    const double m = M.mass[k];
    const V3 w = K.w[k], al = K.alpha[k], ac = K.acc[k];
    const V3 Iw = R * (IB * tmul(R, w));
    const V3 Ial = R * (IB * tmul(R, al));
    const V3 f = m * (ac + cross(al, c) + cross(w, cross(w, c)));
    const V3 t = Ial + cross(w, Iw) + m * cross(c, ac);
    K.Tt[k] = t - K.Ft[k];
    K.Tf[k] = f - K.Ff[k];
  }
  // inward pass
  for (int k = M.nb - 1; k >= 0; --k) {
    const V3 tF = tmul(K.R_WF[k], K.Tt[k]), fF = tmul(K.R_WF[k], K.Tf[k]);
    double* tb = tau + M.vs[k];
    switch (M.jtype[k]) {
      case IDTO_JOINT_REVOLUTE: tb[0] = dot(M.axis[k], tF); break;
      case IDTO_JOINT_PRISMATIC: tb[0] = dot(M.axis[k], fF); break;
      case IDTO_JOINT_PLANAR: tb[0] = fF.x, tb[1] = fF.y, tb[2] = tF.z; break;
      default: tb[0] = tF.x, tb[1] = tF.y, tb[2] = tF.z, tb[3] = fF.x, tb[4] = fF.y, tb[5] = fF.z; break;
    }
    const int p = M.parent[k];
    if (p >= 0) {
      const V3 r = K.p_WB[k] - K.p_WB[p];
      K.Tt[p] = K.Tt[p] + K.Tt[k] + cross(r, K.Tf[k]);
      K.Tf[p] = K.Tf[p] + K.Tf[k];
    }
  }
  // generalized applied forces: joint damping -d*v (cc:232) is subtracted.
  if (with_bias)
    for (int i = 0; i < M.nv; ++i) tau[i] += M.damping[i] * v[i];
}

// Drake MakeQDotToVelocityMap: N+(q), nv x nq column-major.
static void CalcNplusSingle(const Model& M, const double* q, double* N) {
  std::fill(N, N + M.nv * M.nq, 0.0);
  for (int k = 0; k < M.nb; ++k) {
    const int qs = M.qs[k], vs = M.vs[k];
    if (M.jtype[k] == IDTO_JOINT_QUAT_FLOATING) {
      double B[3][4];
      quat_nplus(q + qs, B);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) N[(qs + j) * M.nv + vs + i] = B[i][j];
      for (int i = 0; i < 3; ++i) N[(qs + 4 + i) * M.nv + vs + 3 + i] = 1.0;
    } else {
      const int n = M.jtype[k] == IDTO_JOINT_PLANAR ? 3 : 1;
      for (int i = 0; i < n; ++i) N[(qs + i) * M.nv + vs + i] = 1.0;
    }
  }
}

// ----------------------------------------------------------------------------- dense helpers (column-major)
using Vec = std::vector<double>;
struct Mat {
  int r = 0, c = 0;
  Vec d;
  Mat() {}
  Mat(int r_, int c_, double v = 0.0) : r(r_), c(c_), d(size_t(r_) * c_, v) {}
  double& operator()(int i, int j) { return d[size_t(j) * r + i]; }
  double operator()(int i, int j) const { return d[size_t(j) * r + i]; }
};
static Mat matmul(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  for (int j = 0; j < B.c; ++j)
    for (int k = 0; k < A.c; ++k) {
      const double b = B(k, j);
      for (int i = 0; i < A.r; ++i) C(i, j) += A(i, k) * b;
    }
  return C;
}
static Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < A.c; ++j) T(j, i) = A(i, j);
  return T;
}
static Mat scaled(const Mat& A, double s) {
  Mat B = A;
  for (auto& x : B.d) x *= s;
  return B;
}
static void add_in(Mat& A, const Mat& B) {
  for (size_t i = 0; i < A.d.size(); ++i) A.d[i] += B.d[i];
}
static Mat AtBC(const Mat& A, const Mat& B, const Mat& C) { return matmul(matmul(transpose(A), B), C); }

// Partial-pivot LU (Eigen::PartialPivLU semantics: first max |a_ik| in the column).
struct LU {
  int n = 0;
  Vec lu;
  std::vector<int> piv;
  void compute(const Mat& A) {
    n = A.r;
    lu = A.d;
    piv.resize(n);
    for (int k = 0; k < n; ++k) {
      int p = k;
      double best = std::fabs(lu[size_t(k) * n + k]);
      for (int i = k + 1; i < n; ++i) {
        const double x = std::fabs(lu[size_t(k) * n + i]);
        if (x > best) best = x, p = i;
      }
      piv[k] = p;
      if (p != k)
        for (int j = 0; j < n; ++j) std::swap(lu[size_t(j) * n + k], lu[size_t(j) * n + p]);
      const double d = lu[size_t(k) * n + k];
      if (d != 0.0)
        for (int i = k + 1; i < n; ++i) lu[size_t(k) * n + i] /= d;
      for (int j = k + 1; j < n; ++j) {
        const double u = lu[size_t(j) * n + k];
        for (int i = k + 1; i < n; ++i) lu[size_t(j) * n + i] -= lu[size_t(k) * n + i] * u;
      }
    }
  }
  void solve_in_place(double* b, int ncols, int ldb) const {
    for (int c = 0; c < ncols; ++c) {
      double* x = b + size_t(c) * ldb;
      for (int k = 0; k < n; ++k)
        if (piv[k] != k) std::swap(x[k], x[piv[k]]);
      for (int k = 0; k < n; ++k)
        for (int i = k + 1; i < n; ++i) x[i] -= lu[size_t(k) * n + i] * x[k];
      for (int k = n - 1; k >= 0; --k) {
        x[k] /= lu[size_t(k) * n + k];
        for (int i = 0; i < k; ++i) x[i] -= lu[size_t(k) * n + i] * x[k];
      }
    }
  }
};

// Symmetric solve standing in for Eigen's LDLT (cc:1395): LDL^T with diagonal pivoting.
static bool ldlt_solve(Mat A, Vec* b) {
  const int n = A.r;
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = std::fabs(A(k, k));
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A(i, i)) > best) best = std::fabs(A(i, i)), p = i;
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(A(k, j), A(p, j));
      for (int i = 0; i < n; ++i) std::swap(A(i, k), A(i, p));
      std::swap(perm[k], perm[p]);
    }
    const double d = A(k, k);
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; ++i) A(i, k) /= d;
    // Keep the whole trailing block symmetric-valid (both triangles) so that later symmetric
    // row/column swaps move correct values.
    for (int j = k + 1; j < n; ++j) {
      const double ljd = A(j, k) * d;
      for (int i = k + 1; i < n; ++i) A(i, j) -= A(i, k) * ljd;
    }
    for (int j = k + 1; j < n; ++j) A(k, j) = A(j, k);
  }
  Vec y(n);
  for (int i = 0; i < n; ++i) y[i] = (*b)[perm[i]];
  for (int k = 0; k < n; ++k)
    for (int i = k + 1; i < n; ++i) y[i] -= A(i, k) * y[k];
  for (int k = 0; k < n; ++k) y[k] = A(k, k) != 0.0 ? y[k] / A(k, k) : 0.0;
  for (int k = n - 1; k >= 0; --k)
    for (int i = k + 1; i < n; ++i) y[k] -= A(i, k) * y[i];
  for (int i = 0; i < n; ++i) (*b)[perm[i]] = y[i];
  return true;
}

// ----------------------------------------------------------------------------- PentaDiagonalMatrix (optimizer/penta_diagonal_matrix.{h,cc})
struct Penta {
  int nblocks = 0, k = 0;
  std::vector<Mat> A, B, C, D, E;
  bool symmetric = false;
  Penta() {}
  Penta(int nb, int bs) : nblocks(nb), k(bs) {  // penta_diagonal_matrix.cc:12-22
    A.assign(nb, Mat(bs, bs)), B = A, C = A, D = A, E = A;
    symmetric = true;
  }
  void MakeSymmetric() {  // penta_diagonal_matrix.cc:64-105
    for (int i = 0; i < nblocks; ++i)
      for (int r = 0; r < k; ++r)
        for (int c = r + 1; c < k; ++c) C[i](r, c) = C[i](c, r);
    if (nblocks >= 2) {
      for (int i = 0; i < nblocks - 1; ++i) D[i] = transpose(B[i + 1]);
      D[nblocks - 1] = Mat(k, k);
    }
    if (nblocks >= 3) {
      for (int i = 0; i < nblocks - 2; ++i) E[i] = transpose(A[i + 2]);
      E[nblocks - 1] = Mat(k, k);
      E[nblocks - 2] = Mat(k, k);
    }
    symmetric = true;
  }
  void MultiplyBy(const Vec& v, Vec* out) const {  // penta_diagonal_matrix.cc:181-207
    out->assign(size_t(nblocks) * k, 0.0);
    auto acc = [&](const Mat& Mx, int src, int dst) {
      for (int c = 0; c < k; ++c) {
        const double x = v[size_t(src) * k + c];
        for (int r = 0; r < k; ++r) (*out)[size_t(dst) * k + r] += Mx(r, c) * x;
      }
    };
    for (int i = 0; i < nblocks; ++i) {
      acc(C[i], i, i);
      if (i >= 1) acc(B[i], i - 1, i);
      if (i >= 2) acc(A[i], i - 2, i);
      if (i < nblocks - 1) acc(D[i], i + 1, i);
      if (i < nblocks - 2) acc(E[i], i + 2, i);
    }
  }
  void ExtractDiagonal(Vec* d) const {  // penta_diagonal_matrix.cc:210-218
    d->resize(size_t(nblocks) * k);
    for (int i = 0; i < nblocks; ++i)
      for (int r = 0; r < k; ++r) (*d)[size_t(i) * k + r] = C[i](r, r);
  }
  void ScaleByDiagonal(const Vec& s) {  // penta_diagonal_matrix.cc:221-257
    auto sc = [&](Mat& Mx, int ri, int ci) {
      for (int c = 0; c < k; ++c)
        for (int r = 0; r < k; ++r) Mx(r, c) = s[size_t(ri) * k + r] * Mx(r, c) * s[size_t(ci) * k + c];
    };
    for (int i = 0; i < nblocks; ++i) {
      sc(C[i], i, i);
      if (i >= 1) sc(B[i], i, i - 1);
      if (i >= 2) sc(A[i], i, i - 2);
    }
    if (nblocks >= 2)
      for (int i = 0; i < nblocks - 1; ++i) D[i] = transpose(B[i + 1]);
    if (nblocks >= 3)
      for (int i = 0; i < nblocks - 2; ++i) E[i] = transpose(A[i + 2]);
  }
  Mat MakeDense() const {  // penta_diagonal_matrix.cc:148-169
    Mat Mx(nblocks * k, nblocks * k);
    auto put = [&](const Mat& Bk, int bi, int bj) {
      for (int c = 0; c < k; ++c)
        for (int r = 0; r < k; ++r) Mx(bi * k + r, bj * k + c) = Bk(r, c);
    };
    for (int i = 0; i < nblocks; ++i) {
      if (i >= 2) put(A[i], i, i - 2);
      if (i >= 1) put(B[i], i, i - 1);
      put(C[i], i, i);
      if (i < nblocks - 1) put(D[i], i, i + 1);
      if (i < nblocks - 2) put(E[i], i, i + 2);
    }
    return Mx;
  }
};

// PentaDiagonalFactorization (optimizer/penta_diagonal_solver.h:117-248): block Thomas.
struct PentaFactorization {
  int n = 0, k = 0;
  std::vector<Mat> A, K, Ym2, Zm2;
  std::vector<LU> Ginv;
  bool ok = false;
  explicit PentaFactorization(const Penta& M) {  // :124-197
    n = M.nblocks, k = M.k;
    A = M.A;
    K.assign(n, Mat(k, k));
    Ym2.assign(n + 2, Mat(k, k));
    Zm2.assign(n + 2, Mat(k, k));
    Ginv.resize(n);
    for (int i = 0; i < n; ++i) {
      Mat Ki = M.B[i], Gi = M.C[i];
      const Mat AY = matmul(M.A[i], Ym2[i]), AZ = matmul(M.A[i], Zm2[i]);
      for (size_t e = 0; e < Ki.d.size(); ++e) Ki.d[e] -= AY.d[e], Gi.d[e] -= AZ.d[e];  // :166-167
      const Mat KY = matmul(Ki, Ym2[i + 1]), KZ = matmul(Ki, Zm2[i + 1]);
      Mat Yi = M.D[i];
      for (size_t e = 0; e < Gi.d.size(); ++e) Gi.d[e] -= KY.d[e], Yi.d[e] -= KZ.d[e];  // :173-177
      Ginv[i].compute(Gi);                                                                // :181
      Ginv[i].solve_in_place(Yi.d.data(), k, k);                                          // :188
      Mat Zi = M.E[i];
      Ginv[i].solve_in_place(Zi.d.data(), k, k);  // :193
      K[i] = Ki, Ym2[i + 2] = Yi, Zm2[i + 2] = Zi;
    }
    ok = true;
  }
  void SolveInPlace(double* b) const {  // :199-248
    Vec r(size_t(k) * (n + 2), 0.0);
    std::copy(b, b + size_t(n) * k, r.begin() + 2 * k);
    for (int i = 0; i < n; ++i) {
      double* rim2 = &r[size_t(i) * k];
      double* rim1 = &r[size_t(i + 1) * k];
      double* ri = &r[size_t(i + 2) * k];
      for (int c = 0; c < k; ++c)
        for (int rr = 0; rr < k; ++rr) ri[rr] -= A[i](rr, c) * rim2[c];
      for (int c = 0; c < k; ++c)
        for (int rr = 0; rr < k; ++rr) ri[rr] -= K[i](rr, c) * rim1[c];
      Ginv[i].solve_in_place(ri, 1, k);
    }
    std::copy(r.begin() + 2 * k, r.end(), b);
    if (n >= 2) {
      int i = n - 2;
      for (int c = 0; c < k; ++c)
        for (int rr = 0; rr < k; ++rr) b[size_t(i) * k + rr] -= Ym2[i + 2](rr, c) * b[size_t(i + 1) * k + c];
    }
    for (int i = n - 3; i >= 0; --i) {
      for (int c = 0; c < k; ++c)
        for (int rr = 0; rr < k; ++rr) b[size_t(i) * k + rr] -= Ym2[i + 2](rr, c) * b[size_t(i + 1) * k + c];
      for (int c = 0; c < k; ++c)
        for (int rr = 0; rr < k; ++rr) b[size_t(i) * k + rr] -= Zm2[i + 2](rr, c) * b[size_t(i + 2) * k + c];
    }
  }
};

// ----------------------------------------------------------------------------- state (optimizer/trajectory_optimizer_state.h)
struct State {
  int T = 0, nq = 0, nv = 0;
  std::vector<Vec> q;
  // cache entries + dirty flags (state.h:333-350)
  std::vector<Vec> v, a, tau, Nplus;
  std::vector<Mat> dqm, dqt, dqp;
  double cost = 0, merit = 0;
  Vec g, D, gs, h, lambda, gm;
  Penta H, Hs;
  Mat J;
  bool nplus_ok = false, traj_ok = false, id_ok = false, derivs_ok = false, cost_ok = false, g_ok = false,
       H_ok = false, D_ok = false, Hs_ok = false, gs_ok = false, h_ok = false, J_ok = false,
       lambda_ok = false, merit_ok = false, gm_ok = false;
  void init(int T_, int nq_, int nv_, int nh) {
    T = T_, nq = nq_, nv = nv_;
    q.assign(T + 1, Vec(nq, 0.0));
    v.assign(T + 1, Vec(nv, 0.0));
    a.assign(T, Vec(nv, 0.0));
    tau.assign(T, Vec(nv, 0.0));
    Nplus.assign(T + 1, Vec(size_t(nv) * nq, 0.0));
    // inverse_dynamics_partials.h:32-42
    dqm.assign(T, Mat(nv, nq)), dqt = dqm, dqp = dqm;
    for (auto& x : dqm[0].d) x = std::numeric_limits<double>::quiet_NaN();
    g.assign(size_t(T + 1) * nq, 0.0);
    D.assign(size_t(T + 1) * nq, 1.0);  // state.h:68
    gs = g, gm = g;
    h.assign(nh, 0.0), lambda.assign(nh, 0.0);
    H = Penta(T + 1, nq), Hs = Penta(T + 1, nq);
    J = Mat(nh, (T + 1) * nq);
    invalidate();
  }
  void invalidate() {  // state.h:333-350
    nplus_ok = traj_ok = id_ok = derivs_ok = cost_ok = g_ok = H_ok = D_ok = Hs_ok = gs_ok = h_ok = J_ok =
        lambda_ok = merit_ok = gm_ok = false;
  }
  void set_q(const std::vector<Vec>& qq) { q = qq, invalidate(); }
  void AddToQ(const Vec& dq) {  // state.h:269-275
    for (int t = 0; t <= T; ++t)
      for (int i = 0; i < nq; ++i) q[t][i] += dq[size_t(t) * nq + i];
    invalidate();
  }
  double norm() const {
    double s = 0;
    for (auto& qt : q)
      for (double x : qt) s += x * x;
    return std::sqrt(s);
  }
};

inline double vdot(const Vec& a, const Vec& b) {
  double s = 0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}
inline double vnorm(const Vec& a) { return std::sqrt(vdot(a, a)); }

// ----------------------------------------------------------------------------- TrajectoryOptimizer<double>
struct Optimizer {
  Model M;
  idto_params P;
  ContactParams cp;
  int T, nq, nv;
  double dt;
  Vec q_init, v_init;
  Mat Qq, Qv, Qfq, Qfv, R;
  std::vector<Vec> q_nom, v_nom;
  int num_threads = 1;

  // WarmStart (optimizer/warm_start.h:23-76)
  State state, scratch;
  double Delta;
  Vec dq, dqH;
  bool tr_active = false;
  double last_rho = 0;

  Optimizer(const idto_model_desc& md, const idto_problem_desc& pd, const idto_params& p) : M(md), P(p) {
    T = pd.num_steps, dt = pd.time_step, nq = M.nq, nv = M.nv;
    cp = {p.contact_stiffness, p.smoothing_factor, p.dissipation_velocity, p.stiction_velocity,
          p.friction_coefficient};
    q_init.assign(pd.q_init, pd.q_init + nq);
    v_init.assign(pd.v_init, pd.v_init + nv);
    auto mk = [](const double* src, int n) {
      Mat A(n, n);
      std::copy(src, src + size_t(n) * n, A.d.begin());
      return A;
    };
    Qq = mk(pd.Qq, nq), Qv = mk(pd.Qv, nv), Qfq = mk(pd.Qf_q, nq), Qfv = mk(pd.Qf_v, nv), R = mk(pd.R, nv);
    q_nom.resize(T + 1), v_nom.resize(T + 1);
    for (int t = 0; t <= T; ++t) {
      q_nom[t].assign(pd.q_nom + size_t(t) * nq, pd.q_nom + size_t(t + 1) * nq);
      v_nom[t].assign(pd.v_nom + size_t(t) * nv, pd.v_nom + size_t(t + 1) * nv);
    }
    const int nh = num_eq();
    state.init(T, nq, nv, nh), scratch.init(T, nq, nv, nh);
    Delta = p.Delta0;
    dq.assign(size_t(T + 1) * nq, 0.0), dqH = dq;
  }
  int num_eq() const { return int(M.unactuated.size()) * T; }

  // ---- trajectory data ------------------------------------------------------
  const std::vector<Vec>& EvalNplus(State& s) {  // cc:1633-1647
    if (!s.nplus_ok) {
      for (int t = 0; t <= T; ++t) CalcNplusSingle(M, s.q[t].data(), s.Nplus[t].data());
      s.nplus_ok = true;
    }
    return s.Nplus;
  }
  void CalcCacheTrajectoryData(State& s) {  // cc:1501-1520
    const auto& N = EvalNplus(s);
    s.v[0] = v_init;  // cc:187
    for (int t = 1; t <= T; ++t) {  // cc:188-190
      Vec dqt(nq);
      for (int i = 0; i < nq; ++i) dqt[i] = s.q[t][i] - s.q[t - 1][i];
      for (int r = 0; r < nv; ++r) {
        double acc = 0;
        for (int c = 0; c < nq; ++c) acc += N[t][size_t(c) * nv + r] * dqt[c];
        s.v[t][r] = acc / dt;
      }
    }
    for (int t = 0; t < T; ++t)  // cc:199-201
      for (int i = 0; i < nv; ++i) s.a[t][i] = (s.v[t + 1][i] - s.v[t][i]) / dt;
    s.traj_ok = true;
  }
  const std::vector<Vec>& EvalV(State& s) {
    if (!s.traj_ok) CalcCacheTrajectoryData(s);
    return s.v;
  }
  const std::vector<Vec>& EvalA(State& s) {
    if (!s.traj_ok) CalcCacheTrajectoryData(s);
    return s.a;
  }
  const std::vector<Vec>& EvalTau(State& s) {  // cc:204-226
    if (!s.id_ok) {
      EvalA(s);
#if defined(_OPENMP)
#pragma omp parallel for num_threads(num_threads)
#endif
      for (int t = 0; t < T; ++t)  // "all terms implicit": (q[t+1], v[t+1], a[t]) cc:220-224
        InverseDynamics(M, cp, s.q[t + 1].data(), s.v[t + 1].data(), s.a[t].data(), s.tau[t].data());
      s.id_ok = true;
    }
    return s.tau;
  }
  static double quad(const Mat& Q, const Vec& x) {
    double s = 0;
    for (int j = 0; j < Q.c; ++j) {
      double acc = 0;
      for (int i = 0; i < Q.r; ++i) acc += x[i] * Q(i, j);
      s += acc * x[j];
    }
    return s;
  }
  double EvalCost(State& s) {  // cc:126-176
    if (!s.cost_ok) {
      const auto& v = EvalV(s);
      const auto& tau = EvalTau(s);
      double cost = 0;
      Vec qe(nq), ve(nv);
      for (int t = 0; t < T; ++t) {
        for (int i = 0; i < nq; ++i) qe[i] = s.q[t][i] - q_nom[t][i];
        for (int i = 0; i < nv; ++i) ve[i] = v[t][i] - v_nom[t][i];
        cost += quad(Qq, qe);
        cost += quad(Qv, ve);
        cost += quad(R, tau[t]);
      }
      cost *= dt;
      for (int i = 0; i < nq; ++i) qe[i] = s.q[T][i] - q_nom[T][i];
      for (int i = 0; i < nv; ++i) ve[i] = v[T][i] - v_nom[T][i];
      cost += quad(Qfq, qe);
      cost += quad(Qfv, ve);
      s.cost = cost, s.cost_ok = true;
    }
    return s.cost;
  }

  // ---- inverse dynamics partials -------------------------------------------
  // Perturbation size (cc:504-511, 709-716).  `volatile` keeps the compiler from folding
  // (q+dq)-q back to dq.
  static double StepSize(double qi) {
    const double eps = std::sqrt(std::numeric_limits<double>::epsilon());
    double dq = eps * std::max(1.0, std::fabs(qi));
    volatile double temp = qi + dq;
    dq = temp - qi;
    return dq;
  }
  void CalcPartialsForwardDiff(State& s) {  // cc:426-563
    const auto& q = s.q;
    const auto& v = EvalV(s);
    const auto& a = EvalA(s);
    const auto& tau = EvalTau(s);
    const auto& N = EvalNplus(s);
#if defined(_OPENMP)
#pragma omp parallel for num_threads(num_threads)
#endif
    for (int t = 1; t <= T; ++t) {
      Vec q_eps = q[t], v_eps_t(nv), v_eps_tp(nv), a_eps_tm(nv), a_eps_t(nv), tau_eps(nv);
      for (int i = 0; i < nq; ++i) {
        v_eps_t = v[t], a_eps_tm = a[t - 1];
        if (t < T) v_eps_tp = v[t + 1], a_eps_t = a[t];
        const double dq_i = StepSize(q_eps[i]);
        const double dv_i = dq_i / dt, da_i = dv_i / dt;
        q_eps[i] += dq_i;
        for (int r = 0; r < nv; ++r) {
          const double nt = N[t][size_t(i) * nv + r];
          v_eps_t[r] += dv_i * nt;
          a_eps_tm[r] += da_i * nt;
          if (t < T) {
            const double ntp = N[t + 1][size_t(i) * nv + r];
            v_eps_tp[r] -= dv_i * ntp;
            a_eps_t[r] -= da_i * (ntp + nt);
          }
        }
        InverseDynamics(M, cp, q_eps.data(), v_eps_t.data(), a_eps_tm.data(), tau_eps.data());
        for (int r = 0; r < nv; ++r) s.dqp[t - 1](r, i) = (tau_eps[r] - tau[t - 1][r]) / dq_i;  // cc:531
        if (t < T) {
          InverseDynamics(M, cp, q[t + 1].data(), v_eps_tp.data(), a_eps_t.data(), tau_eps.data());
          for (int r = 0; r < nv; ++r) s.dqt[t](r, i) = (tau_eps[r] - tau[t][r]) / dq_i;  // cc:539
        }
        q_eps[i] = q[t][i];
      }
      if (t < T - 1) {  // cc:556-561: dtau_dqm[t+1] = M(q[t+2]) N+[t+1] / dt^2
        Mat Mm(nv, nv);
        Vec e(nv, 0.0), col(nv);
        for (int j = 0; j < nv; ++j) {
          e[j] = 1.0;
          InverseDynamics(M, cp, q[t + 2].data(), nullptr, e.data(), col.data(), /*with_bias=*/false);
          e[j] = 0.0;
          for (int r = 0; r < nv; ++r) Mm(r, j) = col[r];
        }
        for (int c = 0; c < nq; ++c)
          for (int r = 0; r < nv; ++r) {
            double acc = 0;
            for (int j = 0; j < nv; ++j) acc += Mm(r, j) * N[t + 1][size_t(c) * nv + j];
            s.dqm[t + 1](r, c) = 1 / dt / dt * acc;
          }
      }
    }
  }
  void CalcPartialsCentralDiff(State& s, bool fourth_order) {  // cc:565-885
    const auto& q = s.q;
    const auto& v = EvalV(s);
    const auto& a = EvalA(s);
    const auto& N = EvalNplus(s);
    EvalTau(s);
#if defined(_OPENMP)
#pragma omp parallel for num_threads(num_threads)
#endif
    for (int t = 1; t <= T; ++t) {
      const int nk = fourth_order ? 4 : 2;
      const double mult[4] = {1.0, -1.0, 2.0, -2.0};  // ep, em, epp, emm
      Vec qk(nq), vk(nv), ak(nv);
      std::vector<Vec> tm(4, Vec(nv)), tt(4, Vec(nv)), tp(4, Vec(nv));
      for (int i = 0; i < nq; ++i) {
        const double dq = StepSize(q[t][i]);
        const double dv = dq / dt, da = dv / dt;
        for (int kk = 0; kk < nk; ++kk) {
          const double m = mult[kk];
          // tau[t-1] = ID(q[t]^e, v[t]^e, a[t-1]^e)   cc:763-787
          qk = q[t], vk = v[t], ak = a[t - 1];
          qk[i] += m * dq;
          for (int r = 0; r < nv; ++r) {
            const double nt = N[t][size_t(i) * nv + r];
            vk[r] += m * dv * nt;
            ak[r] += m * da * nt;
          }
          InverseDynamics(M, cp, qk.data(), vk.data(), ak.data(), tm[kk].data());
          if (t < T) {  // tau[t] = ID(q[t+1], v[t+1]^e, a[t]^e)   cc:788-814
            vk = v[t + 1], ak = a[t];
            for (int r = 0; r < nv; ++r) {
              const double nt = N[t][size_t(i) * nv + r], ntp = N[t + 1][size_t(i) * nv + r];
              vk[r] -= m * dv * ntp;
              ak[r] -= m * da * (ntp + nt);
            }
            InverseDynamics(M, cp, q[t + 1].data(), vk.data(), ak.data(), tt[kk].data());
          }
          if (t < T - 1) {  // tau[t+1] = ID(q[t+2], v[t+2], a[t+1]^e)   cc:815-839
            ak = a[t + 1];
            for (int r = 0; r < nv; ++r) ak[r] += m * da * N[t + 1][size_t(i) * nv + r];
            InverseDynamics(M, cp, q[t + 2].data(), v[t + 2].data(), ak.data(), tp[kk].data());
          }
        }
        auto diff = [&](const std::vector<Vec>& x, int r) {
          if (fourth_order)  // cc:782-783
            return 2.0 / 3.0 * (x[0][r] - x[1][r]) / dq - 1.0 / 12.0 * (x[2][r] - x[3][r]) / dq;
          return 0.5 * (x[0][r] - x[1][r]) / dq;  // cc:785
        };
        for (int r = 0; r < nv; ++r) {
          s.dqp[t - 1](r, i) = diff(tm, r);
          if (t < T) s.dqt[t](r, i) = diff(tt, r);
          if (t < T - 1) s.dqm[t + 1](r, i) = diff(tp, r);
        }
      }
    }
  }
  void EvalDerivs(State& s) {  // cc:1587-1604 + cc:388-424
    if (s.derivs_ok) return;
    switch (P.gradients_method) {
      case IDTO_GRAD_FORWARD: CalcPartialsForwardDiff(s); break;
      case IDTO_GRAD_CENTRAL: CalcPartialsCentralDiff(s, false); break;
      default: CalcPartialsCentralDiff(s, true); break;
    }
    s.derivs_ok = true;  // velocity partials (cc:962-973) are +-N+/dt, used in place below
  }
  Mat dvt_dqt(State& s, int t) {  // cc:968
    Mat X(nv, nq);
    X.d = EvalNplus(s)[t];
    return scaled(X, 1 / dt);
  }
  Mat dvt_dqm(State& s, int t) {  // cc:970
    Mat X(nv, nq);
    X.d = EvalNplus(s)[t];
    return scaled(X, -1 / dt);
  }

  // ---- gradient (cc:1021-1081) ---------------------------------------------
  const Vec& EvalGradient(State& s) {
    if (s.g_ok) return s.g;
    const auto& v = EvalV(s);
    const auto& tau = EvalTau(s);
    EvalDerivs(s);
    auto rowvec_times = [](const Vec& x, const Mat& A) {  // x^T A
      Vec out(A.c, 0.0);
      for (int j = 0; j < A.c; ++j)
        for (int i = 0; i < A.r; ++i) out[j] += x[i] * A(i, j);
      return out;
    };
    std::fill(s.g.begin(), s.g.end(), 0.0);
    Vec qe(nq), ve(nv), vep(nv);
    for (int t = 1; t < T; ++t) {
      double* gt = &s.g[size_t(t) * nq];
      for (int i = 0; i < nq; ++i) qe[i] = s.q[t][i] - q_nom[t][i];
      for (int i = 0; i < nv; ++i) ve[i] = v[t][i] - v_nom[t][i], vep[i] = v[t + 1][i] - v_nom[t + 1][i];
      Vec x = rowvec_times(qe, scaled(Qq, 2 * dt));
      for (int i = 0; i < nq; ++i) gt[i] = x[i];
      x = rowvec_times(rowvec_times(ve, scaled(Qv, 2 * dt)), dvt_dqt(s, t));
      for (int i = 0; i < nq; ++i) gt[i] += x[i];
      if (t == T - 1)
        x = rowvec_times(rowvec_times(vep, scaled(Qfv, 2)), dvt_dqm(s, t + 1));
      else
        x = rowvec_times(rowvec_times(vep, scaled(Qv, 2 * dt)), dvt_dqm(s, t + 1));
      for (int i = 0; i < nq; ++i) gt[i] += x[i];
      x = rowvec_times(rowvec_times(tau[t - 1], scaled(R, 2 * dt)), s.dqp[t - 1]);
      for (int i = 0; i < nq; ++i) gt[i] += x[i];
      x = rowvec_times(rowvec_times(tau[t], scaled(R, 2 * dt)), s.dqt[t]);
      for (int i = 0; i < nq; ++i) gt[i] += x[i];
      if (t != T - 1) {
        x = rowvec_times(rowvec_times(tau[t + 1], scaled(R, 2 * dt)), s.dqm[t + 1]);
        for (int i = 0; i < nq; ++i) gt[i] += x[i];
      }
    }
    double* gT = &s.g[size_t(T) * nq];
    Vec x = rowvec_times(rowvec_times(tau[T - 1], scaled(R, 2 * dt)), s.dqp[T - 1]);
    for (int i = 0; i < nq; ++i) gT[i] = x[i];
    for (int i = 0; i < nq; ++i) qe[i] = s.q[T][i] - q_nom[T][i];
    for (int i = 0; i < nv; ++i) ve[i] = v[T][i] - v_nom[T][i];
    x = rowvec_times(qe, scaled(Qfq, 2));
    for (int i = 0; i < nq; ++i) gT[i] += x[i];
    x = rowvec_times(rowvec_times(ve, scaled(Qfv, 2)), dvt_dqt(s, T));
    for (int i = 0; i < nq; ++i) gT[i] += x[i];
    s.g_ok = true;
    return s.g;
  }

  // ---- Hessian (cc:1093-1165) ----------------------------------------------
  const Penta& EvalHessian(State& s) {
    if (s.H_ok) return s.H;
    EvalDerivs(s);
    const Mat Qq_ = scaled(Qq, 2 * dt), Qv_ = scaled(Qv, 2 * dt), R_ = scaled(R, 2 * dt);
    const Mat Qfq_ = scaled(Qfq, 2), Qfv_ = scaled(Qfv, 2);
    Penta& H = s.H;
    H = Penta(T + 1, nq);
    for (int i = 0; i < nq; ++i) H.C[0](i, i) = 1.0;  // cc:1124
    for (int t = 1; t < T; ++t) {
      Mat C = Qq_;
      const Mat dvt = dvt_dqt(s, t), dvm_p = dvt_dqm(s, t + 1);
      add_in(C, AtBC(dvt, Qv_, dvt));
      add_in(C, AtBC(s.dqp[t - 1], R_, s.dqp[t - 1]));
      add_in(C, AtBC(s.dqt[t], R_, s.dqt[t]));
      if (t < T - 1) {
        add_in(C, AtBC(s.dqm[t + 1], R_, s.dqm[t + 1]));
        add_in(C, AtBC(dvm_p, Qv_, dvm_p));
      } else {
        add_in(C, AtBC(dvm_p, Qfv_, dvm_p));
      }
      H.C[t] = C;
      Mat B = AtBC(s.dqp[t], R_, s.dqt[t]);
      const Mat dvt_p = dvt_dqt(s, t + 1);
      if (t < T - 1) {
        add_in(B, AtBC(s.dqt[t + 1], R_, s.dqm[t + 1]));
        add_in(B, AtBC(dvt_p, Qv_, dvm_p));
      } else {
        add_in(B, AtBC(dvt_p, Qfv_, dvm_p));
      }
      H.B[t + 1] = B;
      if (t < T - 1) H.A[t + 2] = AtBC(s.dqp[t + 1], R_, s.dqm[t + 1]);
    }
    Mat CT = Qfq_;
    const Mat dvT = dvt_dqt(s, T);
    add_in(CT, AtBC(dvT, Qfv_, dvT));
    add_in(CT, AtBC(s.dqp[T - 1], R_, s.dqp[T - 1]));
    H.C[T] = CT;
    H.MakeSymmetric();
    s.H_ok = true;
    return s.H;
  }

  // ---- scaling (cc:1181-1265) ----------------------------------------------
  const Vec& EvalScaleFactors(State& s) {
    if (s.D_ok) return s.D;
    Vec diag;
    EvalHessian(s).ExtractDiagonal(&diag);
    for (size_t i = 0; i < s.D.size(); ++i) {
      switch (P.scaling_method) {
        case IDTO_SCALING_SQRT: s.D[i] = std::min(1.0, 1 / std::sqrt(diag[i])); break;
        case IDTO_SCALING_ADAPTIVE_SQRT: s.D[i] = std::min(s.D[i], 1 / std::sqrt(diag[i])); break;
        case IDTO_SCALING_DOUBLE_SQRT: s.D[i] = std::min(1.0, 1 / std::sqrt(std::sqrt(diag[i]))); break;
        default: s.D[i] = std::min(s.D[i], 1 / std::sqrt(std::sqrt(diag[i]))); break;
      }
    }
    s.D_ok = true;
    return s.D;
  }
  const Penta& EvalScaledHessian(State& s) {
    if (!P.scaling) return EvalHessian(s);
    if (!s.Hs_ok) {
      s.Hs = EvalHessian(s);
      s.Hs.ScaleByDiagonal(EvalScaleFactors(s));
      s.Hs_ok = true;
    }
    return s.Hs;
  }
  const Vec& EvalScaledGradient(State& s) {
    if (!P.scaling) return EvalGradient(s);
    if (!s.gs_ok) {
      const Vec& g = EvalGradient(s);
      const Vec& D = EvalScaleFactors(s);
      for (size_t i = 0; i < g.size(); ++i) s.gs[i] = D[i] * g[i];
      s.gs_ok = true;
    }
    return s.gs;
  }

  // ---- equality constraints (cc:1267-1345) ---------------------------------
  const Vec& EvalH(State& s) {
    if (!s.h_ok) {
      const auto& tau = EvalTau(s);
      const int nu = int(M.unactuated.size());
      for (int t = 0; t < T; ++t)
        for (int j = 0; j < nu; ++j) s.h[size_t(t) * nu + j] = tau[t][M.unactuated[j]];
      s.h_ok = true;
    }
    return s.h;
  }
  const Mat& EvalJ(State& s) {
    if (s.J_ok) return s.J;
    EvalDerivs(s);
    const int nu = int(M.unactuated.size());
    std::fill(s.J.d.begin(), s.J.d.end(), 0.0);
    for (int t = 0; t < T; ++t)
      for (int i = 0; i < nu; ++i) {
        const int row = t * nu + i, u = M.unactuated[i];
        for (int c = 0; c < nq; ++c) {
          s.J(row, (t + 1) * nq + c) = s.dqp[t](u, c);
          if (t > 0) s.J(row, t * nq + c) = s.dqt[t](u, c);
          if (t > 1) s.J(row, (t - 1) * nq + c) = s.dqm[t](u, c);
        }
      }
    if (P.scaling) {  // cc:1330-1333
      const Vec& D = EvalScaleFactors(s);
      for (int c = 0; c < s.J.c; ++c)
        for (int r = 0; r < s.J.r; ++r) s.J(r, c) *= D[c];
    }
    s.J_ok = true;
    return s.J;
  }
  const Vec& EvalLambda(State& s) {  // cc:1371-1396
    if (s.lambda_ok) return s.lambda;
    const Penta& H = EvalScaledHessian(s);
    const Vec& g = EvalScaledGradient(s);
    const Vec& h = EvalH(s);
    const Mat& J = EvalJ(s);
    Mat HinvJT = transpose(J);
    PentaFactorization Hlu(H);
    for (int i = 0; i < HinvJT.c; ++i) Hlu.SolveInPlace(&HinvJT.d[size_t(i) * HinvJT.r]);
    const Mat S = matmul(J, HinvJT);
    Vec rhs = h;
    for (int i = 0; i < HinvJT.c; ++i) {
      double acc = 0;
      for (int r = 0; r < HinvJT.r; ++r) acc += HinvJT(r, i) * g[r];
      rhs[i] -= acc;
    }
    ldlt_solve(S, &rhs);
    s.lambda = rhs;
    s.lambda_ok = true;
    return s.lambda;
  }
  double EvalMerit(State& s) {  // cc:1411-1433
    if (!P.equality_constraints) return EvalCost(s);
    if (!s.merit_ok) {
      s.merit = EvalCost(s) + vdot(EvalH(s), EvalLambda(s));
      s.merit_ok = true;
    }
    return s.merit;
  }
  const Vec& EvalMeritGradient(State& s) {  // cc:1435-1456
    if (!P.equality_constraints) return EvalScaledGradient(s);
    if (!s.gm_ok) {
      const Vec& g = EvalScaledGradient(s);
      const Vec& lam = EvalLambda(s);
      const Mat& J = EvalJ(s);
      s.gm = g;
      for (int c = 0; c < J.c; ++c) {
        double acc = 0;
        for (int r = 0; r < J.r; ++r) acc += J(r, c) * lam[r];
        s.gm[c] += acc;
      }
      s.gm_ok = true;
    }
    return s.gm;
  }

  // ---- dogleg (cc:2108-2202) -----------------------------------------------
  static double SolveDoglegQuadratic(double a, double b, double c) {  // cc:2037-2066
    double s;
    if (a < std::numeric_limits<double>::epsilon()) {
      s = -c / b;
    } else {
      const double b_tilde = b / a, c_tilde = c / a;
      const double determinant = b_tilde * b_tilde - 4 * c_tilde;
      s = (-b_tilde + std::sqrt(determinant)) / 2;
    }
    return s;
  }
  bool CalcDoglegPoint(State& s, double Del, Vec* dq_out, Vec* dqH_out) {
    const Penta& H = EvalScaledHessian(s);
    const Vec& g = EvalMeritGradient(s);
    const size_t n = g.size();
    Vec Hg;
    H.MultiplyBy(g, &Hg);
    const double gHg = vdot(g, Hg);
    Vec pH(n);
    for (size_t i = 0; i < n; ++i) pH[i] = -g[i] / Del;
    PentaFactorization Hlu(H);  // cc:2083 (factor again)
    Hlu.SolveInPlace(pH.data());
    for (size_t i = 0; i < n; ++i) (*dqH_out)[i] = pH[i] * Del;
    Vec pU(n);
    const double gg = vdot(g, g);
    for (size_t i = 0; i < n; ++i) pU[i] = -(gg / gHg) * g[i] / Del;
    const Vec* D = P.scaling ? &EvalScaleFactors(s) : nullptr;
    const double pUn = vnorm(pU);
    if (1.0 <= pUn) {
      for (size_t i = 0; i < n; ++i) (*dq_out)[i] = (Del / pUn) * pU[i];
      if (D)
        for (size_t i = 0; i < n; ++i) (*dq_out)[i] = (*D)[i] * (*dq_out)[i];
      return true;
    }
    if (1.0 >= vnorm(pH)) {
      for (size_t i = 0; i < n; ++i) (*dq_out)[i] = pH[i] * Del;
      if (D)
        for (size_t i = 0; i < n; ++i) (*dq_out)[i] = (*D)[i] * (*dq_out)[i];
      return false;
    }
    double a = 0, b = 0, c = 0;
    for (size_t i = 0; i < n; ++i) {
      const double d = pH[i] - pU[i];
      a += d * d, b += pU[i] * d, c += pU[i] * pU[i];
    }
    b *= 2, c -= 1.0;
    const double sq = SolveDoglegQuadratic(a, b, c);
    for (size_t i = 0; i < n; ++i) (*dq_out)[i] = (pU[i] + sq * (pH[i] - pU[i])) * Del;
    if (D)
      for (size_t i = 0; i < n; ++i) (*dq_out)[i] = (*D)[i] * (*dq_out)[i];
    return true;
  }

  void NormalizeQuaternions(State& s) {  // cc:2691-2707
    for (int qs : M.quat_starts)
      for (int t = 0; t <= T; ++t) {
        double* x = &s.q[t][qs];
        const double n = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
        for (int i = 0; i < 4; ++i) x[i] /= n;
      }
  }

  // ---- trust ratio (cc:1979-2035) ------------------------------------------
  double CalcTrustRatio(State& s, const Vec& dq_, State& sc) {
    const double merit_k = EvalMerit(s);
    const Vec& g_tilde_k = EvalMeritGradient(s);
    const Penta& H_k = EvalScaledHessian(s);
    sc.set_q(s.q);
    sc.AddToQ(dq_);
    if (P.normalize_quaternions) NormalizeQuaternions(sc);
    double merit_kp = EvalCost(sc);
    if (P.equality_constraints) merit_kp += vdot(EvalH(sc), EvalLambda(s));
    Vec dq_scaled = dq_;
    if (P.scaling) {
      const Vec& D = EvalScaleFactors(s);
      for (size_t i = 0; i < dq_.size(); ++i) dq_scaled[i] = (1.0 / D[i]) * dq_[i];
    }
    Vec Hdq;
    H_k.MultiplyBy(dq_scaled, &Hdq);
    const double hessian_term = 0.5 * vdot(dq_scaled, Hdq);
    const double gradient_term = vdot(g_tilde_k, dq_scaled);
    const double predicted_reduction = -gradient_term - hessian_term;
    const double actual_reduction = merit_k - merit_kp;
    const double eps = 10 * std::numeric_limits<double>::epsilon() / dt / dt;
    if ((predicted_reduction < eps) && (actual_reduction < eps)) return 0.5;
    return actual_reduction / predicted_reduction;
  }

  int VerifyConvergenceCriteria(State& s, double previous_cost, const Vec& dq_) {  // cc:2653-2689
    int reason = 0;
    const double cost = EvalCost(s);
    if (std::fabs(previous_cost - cost) < P.tol_abs_cost_reduction + P.tol_rel_cost_reduction * cost) reason |= 1;
    const Vec& g = EvalMeritGradient(s);
    if (std::fabs(vdot(g, dq_)) < P.tol_abs_gradient_along_dq + P.tol_rel_gradient_along_dq * cost) reason |= 2;
    if (vnorm(dq_) < P.tol_abs_state_change + P.tol_rel_state_change * s.norm()) reason |= 4;
    return reason;
  }

  // ---- SolveFromWarmStart (cc:2449-2651) -----------------------------------
  int Solve(int max_iterations, double* stats, int* reason_out) {
    const double Delta_max = P.Delta_max, eta = 0.0;
    int k = 0;
    double previous_cost = EvalCost(state);
    if (reason_out) *reason_out = 0;
    while (k < max_iterations) {
      tr_active = CalcDoglegPoint(state, Delta, &dq, &dqH);
      const Vec& g = EvalMeritGradient(state);
      const Vec& h = EvalH(state);
      const double cost = EvalCost(state);
      const double merit = EvalMerit(state);
      const double q_norm = state.norm();
      double dL_dq;
      if (P.scaling) {
        const Vec& D = EvalScaleFactors(state);
        double acc = 0;
        for (size_t i = 0; i < g.size(); ++i) acc += g[i] * ((1.0 / D[i]) * dq[i]);
        dL_dq = acc / cost;
      } else {
        dL_dq = vdot(g, dq) / cost;
      }
      const double rho = CalcTrustRatio(state, dq, scratch);
      last_rho = rho;
      const double gnorm = vnorm(g), hnorm = vnorm(h);  // before q is updated (cc:2509-2510)
      if (rho > eta) {
        state.AddToQ(dq);
        if (P.normalize_quaternions) NormalizeQuaternions(state);
      }
      if (stats) {
        double* st = stats + size_t(k) * IDTO_NUM_STATS;
        st[0] = cost, st[1] = Delta, st[2] = q_norm, st[3] = vnorm(dq), st[4] = vnorm(dqH), st[5] = rho,
        st[6] = gnorm, st[7] = dL_dq, st[8] = hnorm, st[9] = merit;
      }
      int reason = 0;
      if (P.check_convergence && (rho > eta)) {
        reason = VerifyConvergenceCriteria(state, previous_cost, dq);
        previous_cost = EvalCost(state);
        if (reason_out) *reason_out = reason;
      }
      if (reason != 0) {
        ++k;  // this iteration's stats were recorded
        return k;
      }
      if (rho < 0.25) {
        Delta *= 0.25;
      } else if ((rho > 0.75) && tr_active) {
        Delta = std::min(2 * Delta, Delta_max);
      }
      ++k;
    }
    return k;
  }
};

inline long flatten(const std::vector<Vec>& x, double* out) {
  long n = 0;
  for (auto& v : x) {
    if (out) std::copy(v.begin(), v.end(), out + n);
    n += long(v.size());
  }
  return n;
}
inline long flatten(const std::vector<Mat>& x, double* out) {
  long n = 0;
  for (auto& m : x) {
    if (out) std::copy(m.d.begin(), m.d.end(), out + n);
    n += long(m.d.size());
  }
  return n;
}
inline long flatten(const Vec& x, double* out) {
  if (out) std::copy(x.begin(), x.end(), out);
  return long(x.size());
}


}  // namespace

// ============================================================================= C entry points (ctypes)
extern "C" {

void oracle_params_default(idto_params* p) {  // solver_parameters.h:64-167
  std::memset(p, 0, sizeof(*p));
  p->max_iterations = 100, p->gradients_method = IDTO_GRAD_FORWARD, p->normalize_quaternions = 0;
  p->contact_stiffness = 100, p->dissipation_velocity = 0.1, p->stiction_velocity = 0.05;
  p->friction_coefficient = 0.5, p->smoothing_factor = 0.1, p->scaling = 1;
  p->scaling_method = IDTO_SCALING_DOUBLE_SQRT, p->equality_constraints = 1;
  p->Delta0 = 1e-1, p->Delta_max = 1e5, p->check_convergence = 0, p->linear_solver = IDTO_LINSOLVE_THOMAS;
}

void* oracle_create(const idto_model_desc* m, const idto_problem_desc* pd, const idto_params* p) {
  if (m->nbodies > kMaxBodies || m->npairs > kMaxPairs) return nullptr;
  return new Optimizer(*m, *pd, *p);
}
void oracle_destroy(void* h) { delete static_cast<Optimizer*>(h); }
void oracle_set_num_threads(void* h, int n) { static_cast<Optimizer*>(h)->num_threads = n; }
int oracle_max_threads() {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}
int oracle_num_unactuated(void* h) { return int(static_cast<Optimizer*>(h)->M.unactuated.size()); }
void oracle_unactuated_dofs(void* h, int* out) {
  auto& u = static_cast<Optimizer*>(h)->M.unactuated;
  std::copy(u.begin(), u.end(), out);
}

void oracle_set_q(void* h, const double* q) {
  auto* o = static_cast<Optimizer*>(h);
  std::vector<Vec> qq(o->T + 1, Vec(o->nq));
  for (int t = 0; t <= o->T; ++t) std::copy(q + size_t(t) * o->nq, q + size_t(t + 1) * o->nq, qq[t].begin());
  o->state.set_q(qq);
}
void oracle_reset_initial_conditions(void* h, const double* q0, const double* v0) {  // h:463-468
  auto* o = static_cast<Optimizer*>(h);
  o->q_init.assign(q0, q0 + o->nq), o->v_init.assign(v0, v0 + o->nv);
  o->state.invalidate(), o->scratch.invalidate();
}
void oracle_update_nominal_trajectory(void* h, const double* qn, const double* vn) {  // h:477-483
  auto* o = static_cast<Optimizer*>(h);
  for (int t = 0; t <= o->T; ++t) {
    o->q_nom[t].assign(qn + size_t(t) * o->nq, qn + size_t(t + 1) * o->nq);
    o->v_nom[t].assign(vn + size_t(t) * o->nv, vn + size_t(t + 1) * o->nv);
  }
  o->state.invalidate(), o->scratch.invalidate();
}
void oracle_set_delta(void* h, double d) { static_cast<Optimizer*>(h)->Delta = d; }
double oracle_get_delta(void* h) { return static_cast<Optimizer*>(h)->Delta; }

// Evaluates every cache entry of the hot path for the current q (and dogleg + trust ratio with
// the current Delta).  `stage`: 0 trajectory, 1 +derivatives, 2 +assembly, 3 +dogleg, 4 +trust ratio.
void oracle_eval(void* h, int stage) {
  auto* o = static_cast<Optimizer*>(h);
  State& s = o->state;
  o->EvalNplus(s), o->EvalV(s), o->EvalTau(s), o->EvalCost(s), o->EvalH(s);
  if (stage < 1) return;
  o->EvalDerivs(s);
  if (stage < 2) return;
  o->EvalGradient(s), o->EvalHessian(s);
  if (o->P.scaling) o->EvalScaleFactors(s), o->EvalScaledHessian(s), o->EvalScaledGradient(s);
  if (o->P.equality_constraints) o->EvalJ(s), o->EvalLambda(s);
  o->EvalMerit(s), o->EvalMeritGradient(s);
  if (stage < 3) return;
  o->tr_active = o->CalcDoglegPoint(s, o->Delta, &o->dq, &o->dqH);
  if (stage < 4) return;
  o->last_rho = o->CalcTrustRatio(s, o->dq, o->scratch);
}

// Same field names/layouts as idto_get (include/idto_b200.h).  Returns the number of doubles;
// out may be NULL to query the size.
long oracle_get(void* h, const char* name, double* out) {
  auto* o = static_cast<Optimizer*>(h);
  State& s = o->state;
  const std::string f(name);
  auto scalar = [&](double x) {
    if (out) out[0] = x;
    return 1L;
  };
  if (f == "q") return flatten(s.q, out);
  if (f == "v") return flatten(s.v, out);
  if (f == "a") return flatten(s.a, out);
  if (f == "tau") return flatten(s.tau, out);
  if (f == "Nplus") return flatten(s.Nplus, out);
  if (f == "cost") return scalar(s.cost);
  if (f == "h") return flatten(s.h, out);
  if (f == "dtau_dqm") return flatten(s.dqm, out);
  if (f == "dtau_dqt") return flatten(s.dqt, out);
  if (f == "dtau_dqp") return flatten(s.dqp, out);
  if (f == "g") return flatten(s.g, out);
  if (f == "H_A") return flatten(s.H.A, out);
  if (f == "H_B") return flatten(s.H.B, out);
  if (f == "H_C") return flatten(s.H.C, out);
  if (f == "D") return flatten(s.D, out);
  if (f == "Hs_A") return flatten(o->P.scaling ? s.Hs.A : s.H.A, out);
  if (f == "Hs_B") return flatten(o->P.scaling ? s.Hs.B : s.H.B, out);
  if (f == "Hs_C") return flatten(o->P.scaling ? s.Hs.C : s.H.C, out);
  if (f == "gs") return flatten(o->P.scaling ? s.gs : s.g, out);
  if (f == "J") return flatten(s.J.d, out);  // dense (nu*T) x ((T+1) nq), column-major
  if (f == "lambda") return flatten(s.lambda, out);
  if (f == "merit") return scalar(o->P.equality_constraints ? s.merit : s.cost);
  if (f == "gm") return flatten(o->P.equality_constraints ? s.gm : (o->P.scaling ? s.gs : s.g), out);
  if (f == "dq") return flatten(o->dq, out);
  if (f == "dqH") return flatten(o->dqH, out);
  if (f == "dq_active") return scalar(o->tr_active ? 1.0 : 0.0);
  if (f == "rho") return scalar(o->last_rho);
  if (f == "delta") return scalar(o->Delta);
  return -1;
}

// SolveFromWarmStart.  stats: [max_iterations][IDTO_NUM_STATS].  Returns iterations run.
int oracle_solve(void* h, int max_iterations, int* reason_out, double* stats) {
  return static_cast<Optimizer*>(h)->Solve(max_iterations, stats, reason_out);
}

// Solution (cc:2636-2638): q, v, tau of the current state.
void oracle_solution(void* h, double* q, double* v, double* tau) {
  auto* o = static_cast<Optimizer*>(h);
  o->EvalV(o->state), o->EvalTau(o->state);
  if (q) flatten(o->state.q, q);
  if (v) flatten(o->state.v, v);
  if (tau) flatten(o->state.tau, tau);
}

// ---- stand-alone pieces for the Drake-free known-answer tests ---------------------------------
// Single inverse-dynamics evaluation (cc:228-245).  pair_active[npairs] may be NULL.
void oracle_inverse_dynamics(void* h, const double* q, const double* v, const double* a, double* tau,
                             int* pair_active) {
  auto* o = static_cast<Optimizer*>(h);
  InverseDynamics(o->M, o->cp, q, v, a, tau, true, pair_active);
}
void oracle_mass_matrix(void* h, const double* q, double* Mout) {  // column-major nv x nv
  auto* o = static_cast<Optimizer*>(h);
  Vec e(o->nv, 0.0);
  for (int j = 0; j < o->nv; ++j) {
    e[j] = 1.0;
    InverseDynamics(o->M, o->cp, q, nullptr, e.data(), Mout + size_t(j) * o->nv, false);
    e[j] = 0.0;
  }
}
void oracle_body_poses(void* h, const double* q, double* R_WB, double* p_WB) {
  auto* o = static_cast<Optimizer*>(h);
  Kin K;
  PositionKinematics(o->M, q, &K);
  for (int k = 0; k < o->M.nb; ++k) {
    std::memcpy(R_WB + 9 * k, K.R_WB[k].m, 9 * sizeof(double));
    p_WB[3 * k] = K.p_WB[k].x, p_WB[3 * k + 1] = K.p_WB[k].y, p_WB[3 * k + 2] = K.p_WB[k].z;
  }
}

// Penta-diagonal pieces (optimizer/test/penta_diagonal_solver_test.cc).  Bands are
// [nblocks][k*k] column-major blocks; A,B,C lower bands (symmetric matrix is completed).
static Penta make_penta(int nb, int k, const double* A, const double* B, const double* C) {
  Penta P(nb, k);
  for (int i = 0; i < nb; ++i) {
    std::copy(A + size_t(i) * k * k, A + size_t(i + 1) * k * k, P.A[i].d.begin());
    std::copy(B + size_t(i) * k * k, B + size_t(i + 1) * k * k, P.B[i].d.begin());
    std::copy(C + size_t(i) * k * k, C + size_t(i + 1) * k * k, P.C[i].d.begin());
  }
  P.MakeSymmetric();
  return P;
}
void oracle_penta_multiply(int nb, int k, const double* A, const double* B, const double* C, const double* x,
                           double* y) {
  Penta P = make_penta(nb, k, A, B, C);
  Vec xv(x, x + size_t(nb) * k), yv;
  P.MultiplyBy(xv, &yv);
  std::copy(yv.begin(), yv.end(), y);
}
void oracle_penta_solve(int nb, int k, const double* A, const double* B, const double* C, double* b, int nrhs) {
  Penta P = make_penta(nb, k, A, B, C);
  PentaFactorization F(P);
  for (int c = 0; c < nrhs; ++c) F.SolveInPlace(b + size_t(c) * nb * k);
}
void oracle_penta_dense(int nb, int k, const double* A, const double* B, const double* C, double* out) {
  Penta P = make_penta(nb, k, A, B, C);
  Mat D = P.MakeDense();
  std::copy(D.d.begin(), D.d.end(), out);
}
void oracle_penta_scale(int nb, int k, double* A, double* B, double* C, const double* s) {
  Penta P = make_penta(nb, k, A, B, C);
  P.ScaleByDiagonal(Vec(s, s + size_t(nb) * k));
  for (int i = 0; i < nb; ++i) {
    std::copy(P.A[i].d.begin(), P.A[i].d.end(), A + size_t(i) * k * k);
    std::copy(P.B[i].d.begin(), P.B[i].d.end(), B + size_t(i) * k * k);
    std::copy(P.C[i].d.begin(), P.C[i].d.end(), C + size_t(i) * k * k);
  }
}

// Test hook: signed distance, nearest surface point (geometry frame) and gradient (world) from the point p_WQ to
// a shape at pose (R_WG row-major, p_WG).  out = [distance, p_GN(3), grad_W(3)].
void oracle_point_distance(int type, const double* dims, const double* R_WG, const double* p_WG, const double* p_WQ,
                           double* out) {
  M3 R;
  for (int e = 0; e < 9; ++e) R.m[e] = R_WG[e];
  const PointDist d = point_to_shape(type, V3{dims[0], dims[1], dims[2]}, R, V3{p_WG[0], p_WG[1], p_WG[2]},
                                     V3{p_WQ[0], p_WQ[1], p_WQ[2]});
  out[0] = d.distance, out[1] = d.p_GN.x, out[2] = d.p_GN.y, out[3] = d.p_GN.z;
  out[4] = d.grad_W.x, out[5] = d.grad_W.y, out[6] = d.grad_W.z;
}

}  // extern "C"
