// idto_b200_drake.hpp — the one-time bake from a REAL Drake plant.
//
//   #define IDTO_HAVE_DRAKE 1   (and have Drake 1.30's headers on the include path)
//   #include "idto_b200_drake.hpp"
//   auto baked = idto::optimizer::BakeFromPlant(plant, scene_graph.model_inspector());
//   idto::optimizer::TrajectoryOptimizer<double> opt(&diagram_standin, &baked, prob, params);
//
// Replaces what the reference's TrajectoryOptimizer constructor and its hot loop ask the plant for
// (optimizer/trajectory_optimizer.cc:43-72: MakeActuationMatrix; :228-245 CalcInverseDynamics / force elements;
// :272-326 geometry queries: GetFrameId, GetPoseInFrame, GetBodyFromFrameId, the pairs reported by
// ComputeSignedDistancePairwiseClosestPoints, i.e. the collision-filtered candidates; :1645 MakeQDotToVelocityMap)
// by ONE pass over the plant at setup that fills the tables of idto_model_desc (include/idto_b200.h):
//   * moving bodies = bodies whose inboard joint has dofs, in the plant's dof order (velocity_start);
//   * bodies welded (directly or through other welded bodies) to a moving body, or to the world, are merged into it:
//     composite mass / centre of mass / rotational inertia, their geometries re-posed in the merged body;
//   * the merged body's frame is the joint's child frame M (so R_MB = identity and p_MoBo = 0 by construction);
//   * geometries in GetAllGeometryIds(Role::kProximity) order, candidate pairs = GetCollisionCandidates() (the default
//     filters are already applied by the plant), A < B.
// Only Drake's public API is used, and of the Eigen types only operator()(i[, j]) / size() / rows() / cols(), so that
// this header also compiles against the minimal stand-in headers under tests/cpp/drake_stub (the CI check that keeps
// it from rotting on machines without Drake; tests/test_cpp_api.py).
#pragma once
#include "idto_b200.hpp"

#if defined(IDTO_HAVE_DRAKE) && IDTO_HAVE_DRAKE
#include <algorithm>
#include <array>
#include <functional>
#include <map>

#include <drake/geometry/scene_graph_inspector.h>
#include <drake/geometry/shape_specification.h>
#include <drake/multibody/plant/multibody_plant.h>
#include <drake/multibody/tree/planar_joint.h>
#include <drake/multibody/tree/prismatic_joint.h>
#include <drake/multibody/tree/revolute_joint.h>

namespace idto {
namespace optimizer {

namespace bake_detail {
using Mat3 = std::array<std::array<double, 3>, 3>;
using Vec3 = std::array<double, 3>;
struct Pose {
  Mat3 R{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
  Vec3 p{0, 0, 0};
};
template <class RigidTransformT>
Pose to_pose(const RigidTransformT& X) {
  Pose o;
  const auto R = X.rotation().matrix();
  const auto t = X.translation();
  for (int i = 0; i < 3; ++i) {
    o.p[i] = t(i);
    for (int j = 0; j < 3; ++j) o.R[i][j] = R(i, j);
  }
  return o;
}
inline Vec3 apply(const Pose& X, const Vec3& v) {
  Vec3 o;
  for (int i = 0; i < 3; ++i) o[i] = X.p[i] + X.R[i][0] * v[0] + X.R[i][1] * v[1] + X.R[i][2] * v[2];
  return o;
}
inline Pose compose(const Pose& A, const Pose& B) {
  Pose o;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o.R[i][j] = A.R[i][0] * B.R[0][j] + A.R[i][1] * B.R[1][j] + A.R[i][2] * B.R[2][j];
  o.p = apply(A, B.p);
  return o;
}
inline void append(std::vector<double>* out, const Pose& X) {  // 12 doubles: R row-major, then p
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out->push_back(X.R[i][j]);
  for (int i = 0; i < 3; ++i) out->push_back(X.p[i]);
}
}  // namespace bake_detail

inline MultibodyPlant BakeFromPlant(const drake::multibody::MultibodyPlant<double>& plant,
                                    const drake::geometry::SceneGraphInspector<double>& inspector) {
  using namespace bake_detail;
  using drake::multibody::BodyIndex;
  using drake::multibody::Joint;
  using drake::multibody::JointIndex;
  auto context = plant.CreateDefaultContext();
  const int nall = plant.num_bodies();
  // inboard joint of every body (ephemeral floating joints included: every non-world body has one after Finalize)
  std::vector<const Joint<double>*> inboard(nall, nullptr);
  for (JointIndex ji : plant.GetJointIndices()) {
    const Joint<double>& J = plant.get_joint(ji);
    inboard[int(J.child_body().index())] = &J;
  }
  // moving bodies in dof order
  std::vector<int> moving;
  for (int b = 1; b < nall; ++b) {
    if (!inboard[b]) throw std::runtime_error("BakeFromPlant: body without an inboard joint (plant not finalized?)");
    if (inboard[b]->num_velocities() > 0) moving.push_back(b);
  }
  std::sort(moving.begin(), moving.end(),
            [&](int a, int b) { return inboard[a]->velocity_start() < inboard[b]->velocity_start(); });
  std::vector<int> mov_index(nall, -1);
  for (size_t k = 0; k < moving.size(); ++k) mov_index[moving[k]] = int(k);
  // anchor of a body: the moving body it is rigidly attached to (-1: the world)
  std::function<int(int)> anchor = [&](int b) -> int {
    if (b == 0) return -1;
    if (mov_index[b] >= 0) return mov_index[b];
    return anchor(int(inboard[b]->parent_body().index()));
  };
  // pose of a frame in the reference frame of anchor k: the child frame M of k's inboard joint, or the world
  auto pose_in_anchor = [&](int k, const drake::multibody::Frame<double>& frame) {
    const drake::multibody::Frame<double>& ref = k < 0 ? plant.world_frame() : inboard[moving[k]]->frame_on_child();
    return to_pose(plant.CalcRelativeTransform(*context, ref, frame));
  };

  MultibodyPlant::Tables t;
  t.nq = plant.num_positions(), t.nv = plant.num_velocities();
  const int nb = int(moving.size());
  t.damping.assign(t.nv, 0.0);
  for (int k = 0; k < nb; ++k) {
    const Joint<double>& J = *inboard[moving[k]];
    const std::string& type = J.type_name();
    int jt;
    Vec3 axis{0, 0, 1};
    if (type == "revolute") {
      jt = IDTO_JOINT_REVOLUTE;
      const auto a = dynamic_cast<const drake::multibody::RevoluteJoint<double>&>(J).revolute_axis();
      axis = {a(0), a(1), a(2)};
    } else if (type == "prismatic") {
      jt = IDTO_JOINT_PRISMATIC;
      const auto a = dynamic_cast<const drake::multibody::PrismaticJoint<double>&>(J).translation_axis();
      axis = {a(0), a(1), a(2)};
    } else if (type == "planar") {
      jt = IDTO_JOINT_PLANAR;
    } else if (type == "quaternion_floating") {
      jt = IDTO_JOINT_QUAT_FLOATING;
    } else {
      throw std::runtime_error("BakeFromPlant: joint type '" + type + "' is not on the CUDA path");
    }
    const int kp = anchor(int(J.parent_body().index()));
    t.parent.push_back(kp), t.joint_type.push_back(jt);
    t.q_start.push_back(J.position_start()), t.v_start.push_back(J.velocity_start());
    append(&t.X_PF, pose_in_anchor(kp, J.frame_on_parent()));
    for (int e = 0; e < 9; ++e) t.R_MB.push_back(e % 4 == 0 ? 1.0 : 0.0);  // merged body frame := M
    for (int i = 0; i < 3; ++i) t.axis.push_back(axis[i]);
    const auto damp = J.default_damping_vector();
    for (int i = 0; i < int(damp.size()); ++i) t.damping[J.velocity_start() + i] = damp(i);
  }
  // composite inertia of every moving body about the origin of its frame M, expressed in M
  t.mass.assign(nb, 0.0), t.com.assign(3 * nb, 0.0), t.inertia.assign(6 * nb, 0.0);
  std::vector<Mat3> Io(nb, Mat3{});
  for (int b = 1; b < nall; ++b) {
    const int k = anchor(b);
    if (k < 0) continue;  // welded to the world: no dynamics
    const drake::multibody::RigidBody<double>& body = plant.get_body(BodyIndex(b));
    const Pose X_MB = pose_in_anchor(k, body.body_frame());
    const double m = body.default_mass();
    const auto cB = body.default_com();
    const auto IB = body.default_rotational_inertia().CopyToFullMatrix3();  // about Bo, in B
    const Vec3 c_B{cB(0), cB(1), cB(2)};
    const Vec3 c_M = apply(X_MB, c_B);
    // about the centre of mass (parallel axis), re-expressed in M, then about Mo
    Mat3 Icm{};
    const double c2 = c_B[0] * c_B[0] + c_B[1] * c_B[1] + c_B[2] * c_B[2];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Icm[i][j] = IB(i, j) - m * ((i == j ? c2 : 0.0) - c_B[i] * c_B[j]);
    const double d2 = c_M[0] * c_M[0] + c_M[1] * c_M[1] + c_M[2] * c_M[2];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double r = 0.0;
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 3; ++c) r += X_MB.R[i][a] * Icm[a][c] * X_MB.R[j][c];
        Io[k][i][j] += r + m * ((i == j ? d2 : 0.0) - c_M[i] * c_M[j]);
      }
    for (int i = 0; i < 3; ++i) t.com[3 * k + i] += m * c_M[i];
    t.mass[k] += m;
  }
  for (int k = 0; k < nb; ++k) {
    for (int i = 0; i < 3; ++i) t.com[3 * k + i] = t.mass[k] > 0 ? t.com[3 * k + i] / t.mass[k] : 0.0;
    const Mat3& I = Io[k];
    const double six[6] = {I[0][0], I[1][1], I[2][2], I[0][1], I[0][2], I[1][2]};
    for (int e = 0; e < 6; ++e) t.inertia[6 * k + e] = six[e];
  }
  const auto g = plant.gravity_field().gravity_vector();
  t.gravity = {g(0), g(1), g(2)};
  // cc:63-72: a dof is actuated if some column of the actuation matrix drives it
  t.actuated.assign(t.nv, 0);
  const auto Bm = plant.MakeActuationMatrix();
  for (int i = 0; i < int(Bm.rows()); ++i)
    for (int j = 0; j < int(Bm.cols()); ++j)
      if (Bm(i, j) != 0.0) t.actuated[i] = 1;
  // proximity geometries in registration order, posed in their merged body
  const std::vector<drake::geometry::GeometryId> ids = inspector.GetAllGeometryIds(drake::geometry::Role::kProximity);
  std::map<drake::geometry::GeometryId, int> gindex;
  for (const drake::geometry::GeometryId id : ids) {
    const drake::multibody::RigidBody<double>* body = plant.GetBodyFromFrameId(inspector.GetFrameId(id));
    if (!body) throw std::runtime_error("BakeFromPlant: a proximity geometry is not attached to a plant body");
    const int k = anchor(int(body->index()));
    const Pose X_RG = compose(pose_in_anchor(k, body->body_frame()), to_pose(inspector.GetPoseInFrame(id)));
    const drake::geometry::Shape& shape = inspector.GetShape(id);
    const std::string type(shape.type_name());
    int gt;
    Vec3 dims{0, 0, 0};
    if (type == "Sphere") {
      gt = IDTO_GEOM_SPHERE, dims[0] = dynamic_cast<const drake::geometry::Sphere&>(shape).radius();
    } else if (type == "Box") {
      const auto& bx = dynamic_cast<const drake::geometry::Box&>(shape);
      gt = IDTO_GEOM_BOX, dims = {bx.width(), bx.depth(), bx.height()};
    } else if (type == "Capsule") {
      const auto& c = dynamic_cast<const drake::geometry::Capsule&>(shape);
      gt = IDTO_GEOM_CAPSULE, dims = {c.radius(), c.length(), 0};
    } else if (type == "Cylinder") {
      const auto& c = dynamic_cast<const drake::geometry::Cylinder&>(shape);
      gt = IDTO_GEOM_CYLINDER, dims = {c.radius(), c.length(), 0};
    } else if (type == "HalfSpace") {
      gt = IDTO_GEOM_HALF_SPACE;
    } else {
      throw std::runtime_error("BakeFromPlant: shape '" + type + "' has no closed-form signed distance on the CUDA path");
    }
    gindex[id] = int(t.geom_body.size());
    t.geom_body.push_back(k), t.geom_type.push_back(gt);
    for (int i = 0; i < 3; ++i) t.geom_dims.push_back(dims[i]);
    append(&t.X_BG, X_RG);
  }
  std::vector<std::pair<int, int>> pairs;
  for (const auto& pr : inspector.GetCollisionCandidates()) {
    const auto a = gindex.find(pr.first), b = gindex.find(pr.second);
    if (a == gindex.end() || b == gindex.end()) continue;
    pairs.emplace_back(std::min(a->second, b->second), std::max(a->second, b->second));
  }
  std::sort(pairs.begin(), pairs.end());
  for (const auto& pr : pairs) t.pair_geomA.push_back(pr.first), t.pair_geomB.push_back(pr.second);
  return MultibodyPlant::FromTables(t, plant.time_step());
}

}  // namespace optimizer
}  // namespace idto
#endif  // IDTO_HAVE_DRAKE
