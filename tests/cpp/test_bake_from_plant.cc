// BakeFromPlant (include/idto_b200_drake.hpp) compiled against the Drake API stand-in (tests/cpp/drake_stub):
// a plant rebuilt from a baked table set must bake back to the same tables, welded bodies must be merged with the
// right composite inertia and their geometry re-posed.
//   usage: test_bake_from_plant <model.txt> [...]     (tables from BakedModel.save_txt; identity R_MB models)
#define IDTO_HAVE_DRAKE 1
#include <cmath>
#include <cstdio>

#include "idto_b200_drake.hpp"

using namespace idto::optimizer;
namespace dm = drake::multibody;
namespace dg = drake::geometry;
using drake::math::RigidTransform;
using drake::stub::Matrix3d;
using drake::stub::Vector3d;

#define CHECK(c)                                                 \
  do {                                                           \
    if (!(c)) {                                                  \
      std::printf("CHECK failed: %s (line %d)\n", #c, __LINE__); \
      return 1;                                                  \
    }                                                            \
  } while (0)

static RigidTransform<double> pose12(const double* x) {
  Matrix3d R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R(i, j) = x[3 * i + j];
  return RigidTransform<double>(R, Vector3d(x[9], x[10], x[11]));
}

static bool close(const double* a, const double* b, int n, double tol = 1e-12) {
  for (int i = 0; i < n; ++i)
    if (!(std::fabs(a[i] - b[i]) <= tol * std::fmax(1.0, std::fabs(b[i])))) {
      std::printf("  mismatch at %d: %.17g vs %.17g\n", i, a[i], b[i]);
      return false;
    }
  return true;
}

static int roundtrip(const char* path, bool split_last_body) {
  const double dt = 0.05;
  const MultibodyPlant ref = MultibodyPlant::LoadBaked(path, dt);
  const idto_model_desc& d = ref.desc();
  dm::MultibodyPlant<double> plant(dt);
  dg::SceneGraphInspector<double> inspector;
  std::vector<const dm::RigidBody<double>*> body(d.nbodies);
  int extra_geom_body = -1;
  for (int k = 0; k < d.nbodies; ++k) {
    double mass = d.mass[k];
    Vector3d com(d.com[3 * k], d.com[3 * k + 1], d.com[3 * k + 2]);
    Matrix3d I;
    const double* s = d.inertia + 6 * k;
    I(0, 0) = s[0], I(1, 1) = s[1], I(2, 2) = s[2];
    I(0, 1) = I(1, 0) = s[3], I(0, 2) = I(2, 0) = s[4], I(1, 2) = I(2, 1) = s[5];
    const bool split = split_last_body && k == d.nbodies - 1;
    if (split) {  // half of the body stays, the other half becomes a separate body welded at an offset frame
      mass *= 0.5;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) I(i, j) *= 0.5;
    }
    body[k] = &plant.StubAddBody(mass, com, I);
    const dm::RigidBody<double>& parent = d.parent[k] < 0 ? plant.world_body() : *body[d.parent[k]];
    const dm::Frame<double>& F = plant.StubAddFrame(parent, pose12(d.X_PF + 12 * k));
    const Vector3d axis(d.axis[3 * k], d.axis[3 * k + 1], d.axis[3 * k + 2]);
    const int nvb = d.joint_type[k] == IDTO_JOINT_QUAT_FLOATING ? 6 : (d.joint_type[k] == IDTO_JOINT_PLANAR ? 3 : 1);
    std::vector<double> damp(d.damping + d.v_start[k], d.damping + d.v_start[k] + nvb);
    switch (d.joint_type[k]) {
      case IDTO_JOINT_REVOLUTE: plant.StubAddJoint<dm::RevoluteJoint>(F, body[k]->body_frame(), axis, damp); break;
      case IDTO_JOINT_PRISMATIC: plant.StubAddJoint<dm::PrismaticJoint>(F, body[k]->body_frame(), axis, damp); break;
      case IDTO_JOINT_PLANAR: plant.StubAddJoint<dm::PlanarJoint>(F, body[k]->body_frame(), axis, damp); break;
      default: plant.StubAddJoint<dm::QuaternionFloatingJoint>(F, body[k]->body_frame(), axis, damp); break;
    }
    if (split) {
      // the welded half: its body frame W2 sits at X_BW2 (a quarter turn about z, shifted), and carries the same
      // mass distribution expressed in ITS frame, so that the composite equals the original body
      Matrix3d Rz;
      Rz(0, 0) = 0, Rz(0, 1) = -1, Rz(1, 0) = 1, Rz(1, 1) = 0;
      const RigidTransform<double> X_BW2(Rz, Vector3d(0.1, -0.2, 0.3));
      const RigidTransform<double> X_W2B = X_BW2.inverse();
      // com and inertia of the same material expressed in W2: c' = X_W2B c; I'_o' = R I_cm R^T + m(|c'|^2 1 - c'c'^T)
      const auto R = X_W2B.rotation().matrix();
      const auto p = X_W2B.translation();
      Vector3d c2;
      for (int i = 0; i < 3; ++i) c2(i) = p(i) + R(i, 0) * com(0) + R(i, 1) * com(1) + R(i, 2) * com(2);
      const double cc = com(0) * com(0) + com(1) * com(1) + com(2) * com(2);
      const double c2c2 = c2(0) * c2(0) + c2(1) * c2(1) + c2(2) * c2(2);
      Matrix3d Icm, I2;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Icm(i, j) = I(i, j) - mass * ((i == j ? cc : 0.0) - com(i) * com(j));
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double r = 0;
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) r += R(i, a) * Icm(a, b) * R(j, b);
          I2(i, j) = r + mass * ((i == j ? c2c2 : 0.0) - c2(i) * c2(j));
        }
      const dm::RigidBody<double>& half = plant.StubAddBody(mass, c2, I2);
      const dm::Frame<double>& Fw = plant.StubAddFrame(*body[k], X_BW2);
      plant.StubAddJoint<dm::WeldJoint>(Fw, half.body_frame(), Vector3d(0, 0, 1), {});
      extra_geom_body = int(half.index());
    }
  }
  drake::stub::MatrixXd B(d.nv, 0);
  {
    int nu = 0;
    for (int i = 0; i < d.nv; ++i) nu += d.actuated[i] != 0;
    B = drake::stub::MatrixXd(d.nv, nu);
    int u = 0;
    for (int i = 0; i < d.nv; ++i)
      if (d.actuated[i]) B(i, u++) = 1.0;
  }
  plant.StubSetActuation(B);
  plant.mutable_gravity_field().set_gravity_vector(Vector3d(d.gravity[0], d.gravity[1], d.gravity[2]));
  std::vector<dg::GeometryId> gid(d.ngeoms);
  for (int g = 0; g < d.ngeoms; ++g) {
    const int k = d.geom_body[g];
    const double* dims = d.geom_dims + 3 * g;
    std::shared_ptr<dg::Shape> shape;
    switch (d.geom_type[g]) {
      case IDTO_GEOM_SPHERE: shape = std::make_shared<dg::Sphere>(dims[0]); break;
      case IDTO_GEOM_BOX: shape = std::make_shared<dg::Box>(dims[0], dims[1], dims[2]); break;
      case IDTO_GEOM_CAPSULE: shape = std::make_shared<dg::Capsule>(dims[0], dims[1]); break;
      case IDTO_GEOM_CYLINDER: shape = std::make_shared<dg::Cylinder>(dims[0], dims[1]); break;
      default: shape = std::make_shared<dg::HalfSpace>(); break;
    }
    RigidTransform<double> X_BG = pose12(d.X_BG + 12 * g);
    int frame_body = k < 0 ? 0 : int(body[k]->index());
    if (split_last_body && k == d.nbodies - 1 && extra_geom_body >= 0) {
      // register the geometry on the welded half instead: same world placement, pose expressed in W2
      Matrix3d Rz;
      Rz(0, 0) = 0, Rz(0, 1) = -1, Rz(1, 0) = 1, Rz(1, 1) = 0;
      X_BG = RigidTransform<double>(Rz, Vector3d(0.1, -0.2, 0.3)).inverse() * X_BG;
      frame_body = extra_geom_body;
    }
    gid[g] = inspector.StubRegister(dg::FrameId(frame_body), X_BG, shape);
  }
  // an illustration-only geometry must be ignored
  inspector.StubRegister(dg::FrameId(0), RigidTransform<double>(), std::make_shared<dg::Sphere>(1.0), dg::Role::kIllustration);
  for (int p = 0; p < d.npairs; ++p) inspector.StubAddCandidate(gid[d.pair_geomB[p]], gid[d.pair_geomA[p]]);

  const MultibodyPlant baked = BakeFromPlant(plant, inspector);
  const idto_model_desc& o = baked.desc();
  CHECK(o.nbodies == d.nbodies && o.nq == d.nq && o.nv == d.nv && o.ngeoms == d.ngeoms && o.npairs == d.npairs);
  CHECK(baked.time_step() == dt);
  for (int k = 0; k < d.nbodies; ++k) {
    CHECK(o.parent[k] == d.parent[k] && o.joint_type[k] == d.joint_type[k]);
    CHECK(o.q_start[k] == d.q_start[k] && o.v_start[k] == d.v_start[k]);
  }
  for (int i = 0; i < d.nv; ++i) CHECK(o.actuated[i] == d.actuated[i]);
  for (int g = 0; g < d.ngeoms; ++g) CHECK(o.geom_body[g] == d.geom_body[g] && o.geom_type[g] == d.geom_type[g]);
  for (int p = 0; p < d.npairs; ++p) CHECK(o.pair_geomA[p] == d.pair_geomA[p] && o.pair_geomB[p] == d.pair_geomB[p]);
  CHECK(close(o.X_PF, d.X_PF, 12 * d.nbodies) && close(o.R_MB, d.R_MB, 9 * d.nbodies));
  CHECK(close(o.damping, d.damping, d.nv) && close(o.gravity, d.gravity, 3));
  CHECK(close(o.mass, d.mass, d.nbodies) && close(o.com, d.com, 3 * d.nbodies, 1e-11));
  CHECK(close(o.inertia, d.inertia, 6 * d.nbodies, 1e-11));
  CHECK(close(o.geom_dims, d.geom_dims, 3 * d.ngeoms) && close(o.X_BG, d.X_BG, 12 * d.ngeoms, 1e-11));
  for (int k = 0; k < d.nbodies; ++k)
    if (d.joint_type[k] <= IDTO_JOINT_PRISMATIC) CHECK(close(o.axis + 3 * k, d.axis + 3 * k, 3));
  std::printf("%s%s: %d bodies, %d geometries, %d pairs round-trip\n", path, split_last_body ? " (welded half)" : "",
              d.nbodies, d.ngeoms, d.npairs);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  for (int a = 1; a < argc; ++a) {
    if (roundtrip(argv[a], false)) return 1;
    if (roundtrip(argv[a], true)) return 1;
  }
  std::printf("BakeFromPlant test OK\n");
  return 0;
}
