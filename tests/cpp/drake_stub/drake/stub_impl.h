// Minimal stand-in for the subset of Drake 1.30's public API that include/idto_b200_drake.hpp (BakeFromPlant) uses.
// TEST INFRASTRUCTURE: same namespaces, class and method names and signatures as Drake (plus a few `Stub*` builder
// methods Drake does not have), so that BakeFromPlant compiles and runs on machines without Drake.  It models a
// finalized plant at its default configuration: every relative transform BakeFromPlant asks for is between
// rigidly connected frames, i.e. independent of q.
#pragma once
#include <array>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace drake {
namespace stub {
struct Vector3d {
  double v[3]{0, 0, 0};
  Vector3d() = default;
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  double operator()(int i) const { return v[i]; }
  double& operator()(int i) { return v[i]; }
  int size() const { return 3; }
};
struct Matrix3d {
  double m[3][3]{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double operator()(int i, int j) const { return m[i][j]; }
  double& operator()(int i, int j) { return m[i][j]; }
};
struct VectorXd {
  std::vector<double> d;
  int size() const { return int(d.size()); }
  double operator()(int i) const { return d[i]; }
};
struct MatrixXd {
  int r{0}, c{0};
  std::vector<double> d;
  MatrixXd() = default;
  MatrixXd(int r_, int c_) : r(r_), c(c_), d(size_t(r_) * c_, 0.0) {}
  int rows() const { return r; }
  int cols() const { return c; }
  double operator()(int i, int j) const { return d[size_t(j) * r + i]; }
  double& operator()(int i, int j) { return d[size_t(j) * r + i]; }
};
template <class Tag>
class Index {
 public:
  Index() = default;
  explicit Index(int v) : v_(v) {}
  operator int() const { return v_; }

 private:
  int v_{-1};
};
}  // namespace stub

namespace math {
class RotationMatrixd {
 public:
  RotationMatrixd() = default;
  explicit RotationMatrixd(const stub::Matrix3d& R) : R_(R) {}
  const stub::Matrix3d& matrix() const { return R_; }

 private:
  stub::Matrix3d R_;
};
template <class T>
class RigidTransform {
 public:
  RigidTransform() = default;
  RigidTransform(const stub::Matrix3d& R, const stub::Vector3d& p) : R_(R), p_(p) {}
  const RotationMatrixd& rotation() const { return R_; }
  const stub::Vector3d& translation() const { return p_; }
  RigidTransform operator*(const RigidTransform& o) const {
    stub::Matrix3d R;
    stub::Vector3d p;
    const stub::Matrix3d &A = R_.matrix(), &B = o.R_.matrix();
    for (int i = 0; i < 3; ++i) {
      p(i) = p_(i);
      for (int j = 0; j < 3; ++j) {
        R(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
        p(i) += A(i, j) * o.p_(j);
      }
    }
    return RigidTransform(R, p);
  }
  RigidTransform inverse() const {
    stub::Matrix3d R;
    stub::Vector3d p;
    const stub::Matrix3d& A = R_.matrix();
    for (int i = 0; i < 3; ++i) {
      p(i) = 0;
      for (int j = 0; j < 3; ++j) R(i, j) = A(j, i), p(i) -= A(j, i) * p_(j);
    }
    return RigidTransform(R, p);
  }

 private:
  RotationMatrixd R_;
  stub::Vector3d p_;
};
}  // namespace math

namespace systems {
template <class T>
class Context {};
}  // namespace systems

namespace geometry {
struct GeometryIdTag {};
struct FrameIdTag {};
class GeometryId {
 public:
  GeometryId() = default;
  explicit GeometryId(int v) : v_(v) {}
  bool operator<(const GeometryId& o) const { return v_ < o.v_; }
  bool operator==(const GeometryId& o) const { return v_ == o.v_; }
  int get_value() const { return v_; }

 private:
  int v_{-1};
};
class FrameId {
 public:
  FrameId() = default;
  explicit FrameId(int v) : v_(v) {}
  bool operator<(const FrameId& o) const { return v_ < o.v_; }
  int get_value() const { return v_; }

 private:
  int v_{-1};
};
enum class Role { kUnassigned, kProximity, kIllustration, kPerception };
class Shape {
 public:
  virtual ~Shape() = default;
  virtual std::string type_name() const = 0;
};
class Sphere final : public Shape {
 public:
  explicit Sphere(double r) : r_(r) {}
  double radius() const { return r_; }
  std::string type_name() const override { return "Sphere"; }

 private:
  double r_;
};
class Box final : public Shape {
 public:
  Box(double w, double d, double h) : w_(w), d_(d), h_(h) {}
  double width() const { return w_; }
  double depth() const { return d_; }
  double height() const { return h_; }
  std::string type_name() const override { return "Box"; }

 private:
  double w_, d_, h_;
};
class Capsule final : public Shape {
 public:
  Capsule(double r, double l) : r_(r), l_(l) {}
  double radius() const { return r_; }
  double length() const { return l_; }
  std::string type_name() const override { return "Capsule"; }

 private:
  double r_, l_;
};
class Cylinder final : public Shape {
 public:
  Cylinder(double r, double l) : r_(r), l_(l) {}
  double radius() const { return r_; }
  double length() const { return l_; }
  std::string type_name() const override { return "Cylinder"; }

 private:
  double r_, l_;
};
class HalfSpace final : public Shape {
 public:
  std::string type_name() const override { return "HalfSpace"; }
};
template <class T>
class SceneGraphInspector {
 public:
  std::vector<GeometryId> GetAllGeometryIds(std::optional<Role> role = std::nullopt) const {
    std::vector<GeometryId> out;
    for (const auto& g : geoms_)
      if (!role || g.second.role == *role) out.push_back(g.first);
    return out;  // (std::map: ascending ids, like Drake's documented stable order)
  }
  FrameId GetFrameId(GeometryId id) const { return geoms_.at(id).frame; }
  const math::RigidTransform<double>& GetPoseInFrame(GeometryId id) const { return geoms_.at(id).X_FG; }
  const Shape& GetShape(GeometryId id) const { return *geoms_.at(id).shape; }
  std::set<std::pair<GeometryId, GeometryId>> GetCollisionCandidates() const { return candidates_; }
  // ---- stub-only builders
  GeometryId StubRegister(FrameId frame, const math::RigidTransform<double>& X_FG, std::shared_ptr<Shape> shape,
                          Role role = Role::kProximity) {
    GeometryId id(int(geoms_.size()) + 100);
    geoms_[id] = Geom{frame, X_FG, std::move(shape), role};
    return id;
  }
  void StubAddCandidate(GeometryId a, GeometryId b) { candidates_.insert(a < b ? std::make_pair(a, b) : std::make_pair(b, a)); }

 private:
  struct Geom {
    FrameId frame;
    math::RigidTransform<double> X_FG;
    std::shared_ptr<Shape> shape;
    Role role;
  };
  std::map<GeometryId, Geom> geoms_;
  std::set<std::pair<GeometryId, GeometryId>> candidates_;
};
}  // namespace geometry

namespace multibody {
struct BodyTag {};
struct JointTag {};
using BodyIndex = stub::Index<BodyTag>;
using JointIndex = stub::Index<JointTag>;
template <class T>
class RigidBody;
template <class T>
class Frame {
 public:
  Frame(const RigidBody<T>* body, const math::RigidTransform<double>& X_BF) : body_(body), X_BF_(X_BF) {}
  const RigidBody<T>& body() const { return *body_; }
  math::RigidTransform<T> CalcPoseInBodyFrame(const systems::Context<T>&) const { return X_BF_; }
  const math::RigidTransform<double>& StubPoseInBody() const { return X_BF_; }

 private:
  const RigidBody<T>* body_;
  math::RigidTransform<double> X_BF_;
};
template <class T>
class RotationalInertia {
 public:
  RotationalInertia() = default;
  explicit RotationalInertia(const stub::Matrix3d& I) : I_(I) {}
  stub::Matrix3d CopyToFullMatrix3() const { return I_; }

 private:
  stub::Matrix3d I_;
};
template <class T>
class RigidBody {
 public:
  RigidBody(BodyIndex index, double mass, const stub::Vector3d& com, const stub::Matrix3d& I_BBo_B)
      : index_(index), mass_(mass), com_(com), I_(I_BBo_B), frame_(this, math::RigidTransform<double>()) {}
  BodyIndex index() const { return index_; }
  double default_mass() const { return mass_; }
  const stub::Vector3d& default_com() const { return com_; }
  RotationalInertia<double> default_rotational_inertia() const { return RotationalInertia<double>(I_); }
  const Frame<T>& body_frame() const { return frame_; }

 private:
  BodyIndex index_;
  double mass_;
  stub::Vector3d com_;
  stub::Matrix3d I_;
  Frame<T> frame_;
};
template <class T>
class Joint {
 public:
  Joint(const Frame<T>* F, const Frame<T>* M, int nq, int nv, std::vector<double> damping)
      : F_(F), M_(M), nq_(nq), nv_(nv), damping_(std::move(damping)) {}
  virtual ~Joint() = default;
  virtual const std::string& type_name() const = 0;
  const RigidBody<T>& parent_body() const { return F_->body(); }
  const RigidBody<T>& child_body() const { return M_->body(); }
  const Frame<T>& frame_on_parent() const { return *F_; }
  const Frame<T>& frame_on_child() const { return *M_; }
  int num_positions() const { return nq_; }
  int num_velocities() const { return nv_; }
  int position_start() const { return qs_; }
  int velocity_start() const { return vs_; }
  stub::VectorXd default_damping_vector() const { return stub::VectorXd{damping_}; }
  void StubSetStarts(int qs, int vs) { qs_ = qs, vs_ = vs; }

 private:
  const Frame<T>*F_, *M_;
  int nq_, nv_, qs_{0}, vs_{0};
  std::vector<double> damping_;
};
#define IDTO_STUB_JOINT(NAME, STR, NQ, NV)                                                              \
  template <class T>                                                                                   \
  class NAME final : public Joint<T> {                                                                 \
   public:                                                                                             \
    NAME(const Frame<T>* F, const Frame<T>* M, const stub::Vector3d& axis, std::vector<double> damping) \
        : Joint<T>(F, M, NQ, NV, std::move(damping)), axis_(axis) {}                                    \
    const std::string& type_name() const override {                                                    \
      static const std::string s = STR;                                                                \
      return s;                                                                                        \
    }                                                                                                  \
    const stub::Vector3d& revolute_axis() const { return axis_; }                                      \
    const stub::Vector3d& translation_axis() const { return axis_; }                                   \
                                                                                                       \
   private:                                                                                            \
    stub::Vector3d axis_;                                                                              \
  };
IDTO_STUB_JOINT(RevoluteJoint, "revolute", 1, 1)
IDTO_STUB_JOINT(PrismaticJoint, "prismatic", 1, 1)
IDTO_STUB_JOINT(PlanarJoint, "planar", 3, 3)
IDTO_STUB_JOINT(QuaternionFloatingJoint, "quaternion_floating", 7, 6)
IDTO_STUB_JOINT(WeldJoint, "weld", 0, 0)
#undef IDTO_STUB_JOINT

template <class T>
class UniformGravityFieldElement {
 public:
  const stub::Vector3d& gravity_vector() const { return g_; }
  void set_gravity_vector(const stub::Vector3d& g) { g_ = g; }

 private:
  stub::Vector3d g_{0, 0, -9.81};
};

template <class T>
class MultibodyPlant {
 public:
  explicit MultibodyPlant(double time_step) : dt_(time_step) {
    bodies_.push_back(std::make_unique<RigidBody<T>>(BodyIndex(0), 0.0, stub::Vector3d(), stub::Matrix3d()));
  }
  double time_step() const { return dt_; }
  int num_positions() const { return nq_; }
  int num_velocities() const { return nv_; }
  int num_bodies() const { return int(bodies_.size()); }
  std::vector<JointIndex> GetJointIndices() const {
    std::vector<JointIndex> out;
    for (size_t i = 0; i < joints_.size(); ++i) out.push_back(JointIndex(int(i)));
    return out;
  }
  const Joint<T>& get_joint(JointIndex i) const { return *joints_[int(i)]; }
  const RigidBody<T>& get_body(BodyIndex i) const { return *bodies_[int(i)]; }
  const RigidBody<T>& world_body() const { return *bodies_[0]; }
  const Frame<T>& world_frame() const { return bodies_[0]->body_frame(); }
  std::unique_ptr<systems::Context<T>> CreateDefaultContext() const { return std::make_unique<systems::Context<T>>(); }
  // X_AB at the default configuration (identity across every joint)
  math::RigidTransform<T> CalcRelativeTransform(const systems::Context<T>&, const Frame<T>& A, const Frame<T>& B) const {
    return world_pose(A).inverse() * world_pose(B);
  }
  stub::MatrixXd MakeActuationMatrix() const { return B_; }
  const UniformGravityFieldElement<T>& gravity_field() const { return gravity_; }
  UniformGravityFieldElement<T>& mutable_gravity_field() { return gravity_; }
  const RigidBody<T>* GetBodyFromFrameId(geometry::FrameId id) const {
    const int b = id.get_value();
    return b >= 0 && b < num_bodies() ? bodies_[b].get() : nullptr;
  }
  geometry::FrameId GetBodyFrameIdOrThrow(BodyIndex b) const { return geometry::FrameId(int(b)); }
  // ---- stub-only builders
  const RigidBody<T>& StubAddBody(double mass, const stub::Vector3d& com, const stub::Matrix3d& I) {
    bodies_.push_back(std::make_unique<RigidBody<T>>(BodyIndex(int(bodies_.size())), mass, com, I));
    return *bodies_.back();
  }
  const Frame<T>& StubAddFrame(const RigidBody<T>& body, const math::RigidTransform<double>& X_BF) {
    frames_.push_back(std::make_unique<Frame<T>>(&body, X_BF));
    return *frames_.back();
  }
  template <template <class> class JointT>
  const Joint<T>& StubAddJoint(const Frame<T>& F, const Frame<T>& M, const stub::Vector3d& axis,
                               std::vector<double> damping) {
    auto j = std::make_unique<JointT<T>>(&F, &M, axis, std::move(damping));
    j->StubSetStarts(nq_, nv_);
    nq_ += j->num_positions(), nv_ += j->num_velocities();
    joints_.push_back(std::move(j));
    return *joints_.back();
  }
  void StubSetActuation(const stub::MatrixXd& B) { B_ = B; }

 private:
  math::RigidTransform<T> world_pose(const Frame<T>& f) const {
    const RigidBody<T>& body = f.body();
    math::RigidTransform<T> X_WB;
    if (int(body.index()) != 0) {
      const Joint<T>* in = nullptr;
      for (const auto& j : joints_)
        if (int(j->child_body().index()) == int(body.index())) in = j.get();
      if (!in) throw std::runtime_error("stub: body without inboard joint");
      // X_WB = X_WF * X_FM(q0 = identity) * X_MB
      X_WB = world_pose(in->frame_on_parent()) * in->frame_on_child().StubPoseInBody().inverse();
    }
    return X_WB * f.StubPoseInBody();
  }
  double dt_;
  int nq_{0}, nv_{0};
  std::vector<std::unique_ptr<RigidBody<T>>> bodies_;
  std::vector<std::unique_ptr<Frame<T>>> frames_;
  std::vector<std::unique_ptr<Joint<T>>> joints_;
  stub::MatrixXd B_;
  UniformGravityFieldElement<T> gravity_;
};
}  // namespace multibody
}  // namespace drake
