// idto_b200.hpp — header-only C++ mirror of IDTO's optimizer API over the C ABI (idto_b200.h).
//
// Same class / struct / enum / method names as the reference so that code written against
//   optimizer/problem_definition.h:24-59, optimizer/solver_parameters.h:14-167,
//   optimizer/convergence_criteria_tolerances.h:8-39, optimizer/trajectory_optimizer_solution.h:16-185,
//   optimizer/warm_start.h:23-76, optimizer/trajectory_optimizer.h:41-483
// reads the same.  Drake/Eigen types are replaced by minimal stand-ins: VectorXd / MatrixXd (column-major
// storage, the subset of the Eigen interface the API needs) and MultibodyPlant = baked tables + time step
// (what `BakeFromPlant(plant, inspector)` would fill from a real Drake plant).
// Aborts (DRAKE_DEMAND) become std::runtime_error.
#pragma once
#include <cmath>
#include <cstdio>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "idto_b200.h"

namespace idto {
namespace optimizer {

class VectorXd {
 public:
  VectorXd() = default;
  explicit VectorXd(int n, double v = 0.0) : d_(n, v) {}
  VectorXd(std::initializer_list<double> l) : d_(l) {}
  int size() const { return int(d_.size()); }
  void resize(int n) { d_.resize(n); }
  double& operator[](int i) { return d_[i]; }
  double operator[](int i) const { return d_[i]; }
  double& operator()(int i) { return d_[i]; }
  double operator()(int i) const { return d_[i]; }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  bool operator==(const VectorXd& o) const { return d_ == o.d_; }
  double norm() const {
    double s = 0;
    for (double x : d_) s += x * x;
    return std::sqrt(s);
  }

 private:
  std::vector<double> d_;
};

class MatrixXd {
 public:
  MatrixXd() = default;
  MatrixXd(int r, int c) : r_(r), c_(c), d_(size_t(r) * c, 0.0) {}
  static MatrixXd Identity(int r, int c) {
    MatrixXd m(r, c);
    for (int i = 0; i < r && i < c; ++i) m(i, i) = 1.0;
    return m;
  }
  static MatrixXd Diagonal(const std::vector<double>& d) {
    MatrixXd m(int(d.size()), int(d.size()));
    for (size_t i = 0; i < d.size(); ++i) m(int(i), int(i)) = d[i];
    return m;
  }
  MatrixXd operator*(double s) const {
    MatrixXd m = *this;
    for (double& x : m.d_) x *= s;
    return m;
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  double& operator()(int i, int j) { return d_[size_t(j) * r_ + i]; }
  double operator()(int i, int j) const { return d_[size_t(j) * r_ + i]; }
  const double* data() const { return d_.data(); }

 private:
  int r_ = 0, c_ = 0;
  std::vector<double> d_;
};
inline MatrixXd operator*(double s, const MatrixXd& m) { return m * s; }

// ---- optimizer/trajectory_optimizer_solution.h:16-33 ------------------------------------------------
enum SolverFlag { kSuccess, kLinesearchMaxIters, kFactorizationFailed, kMaxIterationsReached };
enum ConvergenceReason : int {
  kNoConvergenceCriteriaSatisfied = 0b000,
  kCostReductionCriterionSatisfied = 0b001,
  kGradientCriterionSatisfied = 0b010,
  kSateCriterionSatisfied = 0b100  // (sic) the reference's spelling is API
};
inline std::string DecodeConvergenceReasons(ConvergenceReason reason) {  // trajectory_optimizer_solution.cc:8-27
  if (reason == kNoConvergenceCriteriaSatisfied) return "no convergence criterion satisfied";
  std::string r;
  if (reason & kCostReductionCriterionSatisfied) r = "cost reduction";
  if (reason & kGradientCriterionSatisfied) r += std::string(r.empty() ? "" : ", ") + "gradient";
  if (reason & kSateCriterionSatisfied) r += std::string(r.empty() ? "" : ", ") + "state change";
  return r;
}

// ---- optimizer/solver_parameters.h ---------------------------------------------------------------------
enum LinesearchMethod { kArmijo, kBacktracking };
enum SolverMethod { kLinesearch, kTrustRegion };
enum GradientsMethod { kForwardDifferences, kCentralDifferences, kCentralDifferences4, kAutoDiff, kNoGradients };
enum ScalingMethod { kSqrt, kAdaptiveSqrt, kDoubleSqrt, kAdaptiveDoubleSqrt };

struct ConvergenceCriteriaTolerances {
  double rel_cost_reduction{0.0}, abs_cost_reduction{0.0}, rel_gradient_along_dq{0.0};
  double abs_gradient_along_dq{0.0}, rel_state_change{0.0}, abs_state_change{0.0};
};

struct SolverParameters {
  // (kCyclicReduction: this build's parallel-in-time order of kPentaDiagonalLu's elimination; not in the reference)
  enum LinearSolverType { kDenseLdlt, kPentaDiagonalLu, kCyclicReduction };
  bool check_convergence = false;
  ConvergenceCriteriaTolerances convergence_tolerances;
  SolverMethod method{kTrustRegion};
  LinesearchMethod linesearch_method{kArmijo};
  int max_iterations{100};
  int max_linesearch_iterations{50};
  GradientsMethod gradients_method{kForwardDifferences};
  LinearSolverType linear_solver{kPentaDiagonalLu};
  bool normalize_quaternions{false};
  bool verbose{true};
  double contact_stiffness{100};
  double dissipation_velocity{0.1};
  double stiction_velocity{0.05};
  double friction_coefficient{0.5};
  double smoothing_factor{0.1};
  bool exact_hessian{false};
  bool scaling{true};
  ScalingMethod scaling_method{kDoubleSqrt};
  bool equality_constraints{true};
  double Delta0{1e-1};
  double Delta_max{1e5};
  int num_threads{1};
};

// ---- optimizer/problem_definition.h:24-59 ---------------------------------------------------------------
struct ProblemDefinition {
  int num_steps{0};
  VectorXd q_init, v_init;
  MatrixXd Qq, Qv, Qf_q, Qf_v, R;
  std::vector<VectorXd> q_nom, v_nom;
};

template <typename T>
struct TrajectoryOptimizerSolution {
  std::vector<VectorXd> q, v, tau;
};

template <typename T>
struct TrajectoryOptimizerStats {
  ConvergenceReason convergence_reason{kNoConvergenceCriteriaSatisfied};
  double solve_time{0};
  std::vector<double> iteration_times;
  std::vector<T> iteration_costs;
  std::vector<int> linesearch_iterations;
  std::vector<double> linesearch_alphas;
  std::vector<T> trust_region_radii, gradient_norms, q_norms, dq_norms, dqH_norms, trust_ratios, dL_dqs, h_norms, merits;
  void push_data(double iter_time, T iter_cost, int linesearch_iters, double alpha, double delta, T q_norm, T dq_norm,
                 T dqH_norm, T trust_ratio, T grad_norm, T dL_dq, T h_norm, T merit) {
    iteration_times.push_back(iter_time), iteration_costs.push_back(iter_cost);
    linesearch_iterations.push_back(linesearch_iters), linesearch_alphas.push_back(alpha);
    trust_region_radii.push_back(delta), q_norms.push_back(q_norm), dq_norms.push_back(dq_norm);
    dqH_norms.push_back(dqH_norm), trust_ratios.push_back(trust_ratio), gradient_norms.push_back(grad_norm);
    dL_dqs.push_back(dL_dq), h_norms.push_back(h_norm), merits.push_back(merit);
  }
  bool is_empty() const {
    return iteration_times.empty() && iteration_costs.empty() && linesearch_iterations.empty() &&
           linesearch_alphas.empty() && trust_region_radii.empty() && q_norms.empty() && dq_norms.empty() &&
           dqH_norms.empty() && trust_ratios.empty() && gradient_norms.empty() && dL_dqs.empty() &&
           h_norms.empty() && merits.empty();
  }
  void SaveToCsv(std::string fname) const {
    std::ofstream f(fname);
    f << "iter, time, cost, ls_iters, alpha, delta, q_norm, dq_norm, dqH_norm, trust_ratio, grad_norm, dL_dq, "
         "h_norm, merit\n";
    for (size_t i = 0; i < iteration_times.size(); ++i)
      f << i << ", " << iteration_times[i] << ", " << iteration_costs[i] << ", " << linesearch_iterations[i] << ", "
        << linesearch_alphas[i] << ", " << trust_region_radii[i] << ", " << q_norms[i] << ", " << dq_norms[i] << ", "
        << dqH_norms[i] << ", " << trust_ratios[i] << ", " << gradient_norms[i] << ", " << dL_dqs[i] << ", "
        << h_norms[i] << ", " << merits[i] << "\n";
  }
};

// ---- MultibodyPlant stand-in: baked tables (idto_b200/bake.py `save_txt`) + discrete time step ----------
class MultibodyPlant {
 public:
  // The baked tables as plain vectors (field names and layouts of idto_model_desc, include/idto_b200.h): what
  // idto_b200/bake.py writes from URDF / SDF and what BakeFromPlant (idto_b200_drake.hpp) fills from a Drake plant.
  struct Tables {
    int nq{0}, nv{0};
    std::vector<int> parent, joint_type, q_start, v_start, actuated, geom_body, geom_type, pair_geomA, pair_geomB;
    std::vector<double> X_PF, R_MB, axis, damping, mass, com, inertia, gravity, geom_dims, X_BG;
  };
  static MultibodyPlant FromTables(const Tables& t, double time_step) {
    MultibodyPlant p;
    p.dt_ = time_step;
    const int nb = int(t.parent.size()), ng = int(t.geom_body.size()), np = int(t.pair_geomA.size());
    if (int(t.joint_type.size()) != nb || int(t.q_start.size()) != nb || int(t.v_start.size()) != nb ||
        int(t.X_PF.size()) != 12 * nb || int(t.R_MB.size()) != 9 * nb || int(t.axis.size()) != 3 * nb ||
        int(t.mass.size()) != nb || int(t.com.size()) != 3 * nb || int(t.inertia.size()) != 6 * nb ||
        int(t.damping.size()) != t.nv || int(t.actuated.size()) != t.nv || int(t.gravity.size()) != 3 ||
        int(t.geom_type.size()) != ng || int(t.geom_dims.size()) != 3 * ng || int(t.X_BG.size()) != 12 * ng ||
        int(t.pair_geomB.size()) != np)
      throw std::runtime_error("MultibodyPlant::FromTables: inconsistent table sizes");
    auto pad = [](auto v) { if (v.empty()) v.resize(1); return v; };
    p.parent_ = t.parent, p.jt_ = t.joint_type, p.qs_ = t.q_start, p.vs_ = t.v_start, p.act_ = pad(t.actuated);
    p.gb_ = pad(t.geom_body), p.gt_ = pad(t.geom_type), p.pa_ = pad(t.pair_geomA), p.pb_ = pad(t.pair_geomB);
    p.xpf_ = t.X_PF, p.rmb_ = t.R_MB, p.axis_ = t.axis, p.damp_ = pad(t.damping), p.mass_ = t.mass, p.com_ = t.com;
    p.inertia_ = t.inertia, p.grav_ = t.gravity, p.gd_ = pad(t.geom_dims), p.xbg_ = pad(t.X_BG);
    idto_model_desc& d = p.d_;
    d.nbodies = nb, d.nq = t.nq, d.nv = t.nv, d.ngeoms = ng, d.npairs = np;
    p.bind();
    return p;
  }
  MultibodyPlant() = default;
  MultibodyPlant(const MultibodyPlant& o) { *this = o; }
  MultibodyPlant& operator=(const MultibodyPlant& o) {  // the descriptor points into this object's own vectors
    d_ = o.d_, dt_ = o.dt_;
    parent_ = o.parent_, jt_ = o.jt_, qs_ = o.qs_, vs_ = o.vs_, act_ = o.act_, gb_ = o.gb_, gt_ = o.gt_, pa_ = o.pa_, pb_ = o.pb_;
    xpf_ = o.xpf_, rmb_ = o.rmb_, axis_ = o.axis_, damp_ = o.damp_, mass_ = o.mass_, com_ = o.com_;
    inertia_ = o.inertia_, grav_ = o.grav_, gd_ = o.gd_, xbg_ = o.xbg_;
    bind();
    return *this;
  }
  static MultibodyPlant LoadBaked(const std::string& path, double time_step) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open baked model " + path);
    MultibodyPlant p;
    p.dt_ = time_step;
    idto_model_desc& d = p.d_;
    f >> d.nbodies >> d.nq >> d.nv >> d.ngeoms >> d.npairs;
    auto ri = [&](std::vector<int>& v, int n) { v.resize(n > 0 ? n : 1); for (int i = 0; i < n; ++i) f >> v[i]; };
    auto rd = [&](std::vector<double>& v, int n) { v.resize(n > 0 ? n : 1); for (int i = 0; i < n; ++i) f >> v[i]; };
    const int nb = d.nbodies, ng = d.ngeoms, np = d.npairs;
    ri(p.parent_, nb), ri(p.jt_, nb), ri(p.qs_, nb), ri(p.vs_, nb), ri(p.act_, d.nv), ri(p.gb_, ng), ri(p.gt_, ng);
    ri(p.pa_, np), ri(p.pb_, np);
    rd(p.xpf_, 12 * nb), rd(p.rmb_, 9 * nb), rd(p.axis_, 3 * nb), rd(p.damp_, d.nv), rd(p.mass_, nb);
    rd(p.com_, 3 * nb), rd(p.inertia_, 6 * nb), rd(p.grav_, 3), rd(p.gd_, 3 * ng), rd(p.xbg_, 12 * ng);
    if (!f) throw std::runtime_error("malformed baked model " + path);
    p.bind();
    return p;
  }
  double time_step() const { return dt_; }
  int num_positions() const { return d_.nq; }
  int num_velocities() const { return d_.nv; }
  const idto_model_desc& desc() const { return d_; }

 private:
  void bind() {
    idto_model_desc& d = d_;
    d.parent = parent_.data(), d.joint_type = jt_.data(), d.q_start = qs_.data(), d.v_start = vs_.data();
    d.actuated = act_.data(), d.geom_body = gb_.data(), d.geom_type = gt_.data();
    d.pair_geomA = pa_.data(), d.pair_geomB = pb_.data();
    d.X_PF = xpf_.data(), d.R_MB = rmb_.data(), d.axis = axis_.data(), d.damping = damp_.data();
    d.mass = mass_.data(), d.com = com_.data(), d.inertia = inertia_.data();
    d.geom_dims = gd_.data(), d.X_BG = xbg_.data();
    for (int i = 0; i < 3; ++i) d.gravity[i] = grav_[i];
  }
  idto_model_desc d_{};
  double dt_{0};
  std::vector<int> parent_, jt_, qs_, vs_, act_, gb_, gt_, pa_, pb_;
  std::vector<double> xpf_, rmb_, axis_, damp_, mass_, com_, inertia_, grav_, gd_, xbg_;
};
template <typename T>
class Diagram {};  // signature compatibility only

inline void ThrowOnError(int rc, const char* what) {
  if (rc != IDTO_OK) throw std::runtime_error(std::string(what) + ": " + idto_last_error());
}

// ---- optimizer/inverse_dynamics_partials.h:20-85, velocity_partials.h:19-39, penta_diagonal_matrix.h ------------
struct InverseDynamicsPartials {
  std::vector<MatrixXd> dtau_dqm, dtau_dqt, dtau_dqp;  // [num_steps] of nv x nq (dqm[0] = NaN, dqt[0] = dqm[1] = 0)
};
struct VelocityPartials {
  std::vector<MatrixXd> dvt_dqt, dvt_dqm;  // [num_steps + 1] of nv x nq (dvt_dqm[0] = NaN)
};
struct PentaDiagonalMatrix {  // symmetric: D_i = B_{i+1}^T, E_i = A_{i+2}^T (penta_diagonal_matrix.cc:64-105)
  std::vector<MatrixXd> A, B, C, D, E;
  int block_rows() const { return int(C.size()); }
  int block_size() const { return C.empty() ? 0 : C[0].rows(); }
};
// optimizer/trajectory_optimizer_state.h: q and everything computed from it.  Here: a view of the device-resident
// cache of one WarmStart; the Eval* accessors of TrajectoryOptimizer compute what is stale and mirror it to the host.
class TrajectoryOptimizerState {
 public:
  explicit TrajectoryOptimizerState(idto_solver_t s = nullptr) : s_(s) {}
  idto_solver_t handle() const { return s_; }

 private:
  idto_solver_t s_;
};

// ---- optimizer/warm_start.h:23-76 ---------------------------------------------------------------------------
class WarmStart {
 public:
  WarmStart(idto_model_t model, const idto_problem_desc& pd, const idto_params& p, int T, int nq,
            const std::vector<VectorXd>& q_guess)
      : T_(T), nq_(nq) {
    ThrowOnError(idto_solver_create(model, &pd, &p, 1, &s_), "idto_solver_create");
    state = TrajectoryOptimizerState(s_);
    set_q(q_guess);
  }
  ~WarmStart() { if (s_) idto_solver_destroy(s_); }
  WarmStart(const WarmStart&) = delete;
  void set_q(const std::vector<VectorXd>& q_guess) {
    std::vector<double> q(size_t(T_ + 1) * nq_);
    for (int t = 0; t <= T_; ++t)
      for (int i = 0; i < nq_; ++i) q[size_t(t) * nq_ + i] = q_guess[t][i];
    ThrowOnError(idto_set_q(s_, q.data()), "idto_set_q");
  }
  std::vector<VectorXd> get_q() const { return series("q", T_ + 1, nq_); }
  double Delta() const {
    double d = 0;
    ThrowOnError(idto_get(s_, "delta", &d), "idto_get");
    return d;
  }
  std::vector<VectorXd> series(const char* name, int rows, int cols) const {
    std::vector<double> buf(size_t(rows) * cols);
    ThrowOnError(idto_get(s_, name, buf.data()), "idto_get");
    std::vector<VectorXd> out(rows, VectorXd(cols));
    for (int t = 0; t < rows; ++t)
      for (int i = 0; i < cols; ++i) out[t][i] = buf[size_t(t) * cols + i];
    return out;
  }
  // The MPC shell between two re-solves (ModelPredictiveController::UpdateAbstractState,
  // examples/mpc_controller.cc:43-98), on the device: the stored solution is spline-shifted by `elapsed`
  // seconds into the new guess, q_nom moves with q0 for the selected coordinates, (q0, v0) become the initial
  // conditions.  `q_nom_relative_to_q_init` may be empty (no shift).
  void AdvanceFromMeasuredState(double elapsed, const VectorXd& q0, const VectorXd& v0,
                                const std::vector<bool>& q_nom_relative_to_q_init = {}) {
    std::vector<double> sel(q_nom_relative_to_q_init.begin(), q_nom_relative_to_q_init.end());
    if (q0.size() != nq_ || (!sel.empty() && int(sel.size()) != nq_))
      throw std::runtime_error("AdvanceFromMeasuredState: wrong sizes");
    ThrowOnError(idto_mpc_advance(s_, &elapsed, q0.data(), v0.data(), sel.empty() ? nullptr : sel.data()),
                 "idto_mpc_advance");
  }
  idto_solver_t handle() const { return s_; }
  int version{0};
  TrajectoryOptimizerState state;  // warm_start.h:63

 private:
  idto_solver_t s_{nullptr};
  int T_, nq_;
};

// ---- optimizer/trajectory_optimizer.h:41-483 (T = double only) -----------------------------------------------------
template <typename T>
class TrajectoryOptimizer {
 public:
  TrajectoryOptimizer(const Diagram<T>* diagram, const MultibodyPlant* plant, const ProblemDefinition& prob,
                      const SolverParameters& params = SolverParameters{})
      : diagram_(diagram), plant_(plant), prob_(prob), params_(params) {
    const int Tn = prob.num_steps;
    if (int(prob.q_nom.size()) != Tn + 1 || int(prob.v_nom.size()) != Tn + 1)  // cc:75-76
      throw std::runtime_error("q_nom and v_nom must have num_steps + 1 entries");
    for (int t = 0; t <= Tn; ++t)
      if (prob.q_nom[t].size() != plant->num_positions() || prob.v_nom[t].size() != plant->num_velocities())
        throw std::runtime_error("q_nom / v_nom entries have the wrong size");  // cc:79-82
    if (params.method != kTrustRegion) throw std::runtime_error("only SolverMethod::kTrustRegion is on the CUDA path");
    if (params.gradients_method > kCentralDifferences4)
      throw std::runtime_error("autodiff gradients need Drake scalars; use forward/central differences");
    ThrowOnError(idto_model_create(&plant->desc(), &model_), "idto_model_create");
    const int nu = idto_model_num_unactuated(model_);
    unactuated_dofs_.resize(nu);
    if (nu > 0) idto_model_unactuated_dofs(model_, unactuated_dofs_.data());
  }
  ~TrajectoryOptimizer() { if (model_) idto_model_destroy(model_); }
  TrajectoryOptimizer(const TrajectoryOptimizer&) = delete;

  double time_step() const { return plant_->time_step(); }
  int num_steps() const { return prob_.num_steps; }
  const std::vector<int>& unactuated_dofs() const { return unactuated_dofs_; }
  int num_equality_constraints() const { return int(unactuated_dofs_.size()) * num_steps(); }
  const MultibodyPlant& plant() const { return *plant_; }
  const SolverParameters& params() const { return params_; }
  const ProblemDefinition& prob() const { return prob_; }

  std::unique_ptr<WarmStart> CreateWarmStart(const std::vector<VectorXd>& q_guess) const {  // cc:1353-1361
    if (int(q_guess.size()) != num_steps() + 1 || q_guess[0].size() != plant_->num_positions())
      throw std::runtime_error("CreateWarmStart: q_guess must be (num_steps + 1) x nq");
    Flat f(prob_, *plant_);
    const idto_params p = c_params();
    auto ws = std::make_unique<WarmStart>(model_, f.desc, p, num_steps(), plant_->num_positions(), q_guess);
    ws->version = version_;
    return ws;
  }

  SolverFlag Solve(const std::vector<VectorXd>& q_guess, TrajectoryOptimizerSolution<T>* solution,
                   TrajectoryOptimizerStats<T>* stats, ConvergenceReason* reason = nullptr) const {  // cc:2213-2234
    if (!(q_guess[0] == prob_.q_init)) throw std::runtime_error("Solve: q_guess[0] must equal q_init");
    if (!stats->is_empty()) throw std::runtime_error("Solve: stats must be empty");
    auto ws = CreateWarmStart(q_guess);
    return SolveFromWarmStart(ws.get(), solution, stats, reason);
  }

  SolverFlag SolveFromWarmStart(WarmStart* ws, TrajectoryOptimizerSolution<T>* solution,
                                TrajectoryOptimizerStats<T>* stats, ConvergenceReason* reason = nullptr) const {
    const int Tn = num_steps(), nq = plant_->num_positions(), nv = plant_->num_velocities();
    if (ws->version != version_) {  // ResetInitialConditions / UpdateNominalTrajectory since creation
      Flat f(prob_, *plant_);
      ThrowOnError(idto_reset_initial_conditions(ws->handle(), f.q_init.data(), f.v_init.data()), "reset");
      ThrowOnError(idto_update_nominal_trajectory(ws->handle(), f.q_nom.data(), f.v_nom.data()), "nominal");
      ws->version = version_;
    }
    const int K = params_.max_iterations;
    int iters = 0, why = 0;
    std::vector<double> st(size_t(K > 0 ? K : 1) * IDTO_NUM_STATS);
    ThrowOnError(idto_solve(ws->handle(), K, &iters, &why, st.data()), "idto_solve");
    for (int k = 0; k < iters; ++k) {
      const double* r = &st[size_t(k) * IDTO_NUM_STATS];
      stats->push_data(0.0, r[0], 0, NAN, r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9]);
    }
    ThrowOnError(idto_eval_trajectory(ws->handle()), "idto_eval_trajectory");
    solution->q = ws->series("q", Tn + 1, nq);
    solution->v = ws->series("v", Tn + 1, nv);
    solution->tau = ws->series("tau", Tn, nv);
    if (reason) *reason = static_cast<ConvergenceReason>(why);
    stats->convergence_reason = static_cast<ConvergenceReason>(why);
    return (why == 0 && iters == K) ? kMaxIterationsReached : kSuccess;  // cc:2648-2650
  }

  // ---- evaluators (trajectory_optimizer.h:125-453): compute what is stale on the device, mirror to the host ------
  std::vector<VectorXd> EvalV(const TrajectoryOptimizerState& st) const { return rows(st, 0, "v", num_steps() + 1, nv()); }
  std::vector<VectorXd> EvalA(const TrajectoryOptimizerState& st) const { return rows(st, 0, "a", num_steps(), nv()); }
  std::vector<VectorXd> EvalTau(const TrajectoryOptimizerState& st) const { return rows(st, 0, "tau", num_steps(), nv()); }
  double EvalCost(const TrajectoryOptimizerState& st) const { return flat(st, 0, "cost")[0]; }
  std::vector<MatrixXd> EvalNplus(const TrajectoryOptimizerState& st) const {
    return blocks(flat(st, 0, "Nplus"), num_steps() + 1, nv(), nq());
  }
  InverseDynamicsPartials EvalInverseDynamicsPartials(const TrajectoryOptimizerState& st) const {
    InverseDynamicsPartials p;
    p.dtau_dqm = blocks(flat(st, 1, "dtau_dqm"), num_steps(), nv(), nq());
    p.dtau_dqt = blocks(flat(st, 1, "dtau_dqt"), num_steps(), nv(), nq());
    p.dtau_dqp = blocks(flat(st, 1, "dtau_dqp"), num_steps(), nv(), nq());
    return p;
  }
  VelocityPartials EvalVelocityPartials(const TrajectoryOptimizerState& st) const {  // cc:962-973
    VelocityPartials p;
    p.dvt_dqt = blocks(flat(st, 0, "dvt_dqt"), num_steps() + 1, nv(), nq());
    p.dvt_dqm = blocks(flat(st, 0, "dvt_dqm"), num_steps() + 1, nv(), nq());
    return p;
  }
  VectorXd EvalGradient(const TrajectoryOptimizerState& st) const { return vec(flat(st, 2, "g")); }
  VectorXd EvalScaledGradient(const TrajectoryOptimizerState& st) const { return vec(flat(st, 2, "gs")); }
  VectorXd EvalScaleFactors(const TrajectoryOptimizerState& st) const { return vec(flat(st, 2, "D")); }
  PentaDiagonalMatrix EvalHessian(const TrajectoryOptimizerState& st) const { return penta(st, "H_A", "H_B", "H_C"); }
  PentaDiagonalMatrix EvalScaledHessian(const TrajectoryOptimizerState& st) const { return penta(st, "Hs_A", "Hs_B", "Hs_C"); }
  VectorXd EvalEqualityConstraintViolations(const TrajectoryOptimizerState& st) const { return vec(flat(st, 0, "h")); }
  MatrixXd EvalEqualityConstraintJacobian(const TrajectoryOptimizerState& st) const {  // dense (nu T) x n, cc:1292-1334
    const std::vector<double> J = flat(st, 2, "J");
    MatrixXd out(num_equality_constraints(), (num_steps() + 1) * nq());
    for (int j = 0; j < out.cols(); ++j)
      for (int i = 0; i < out.rows(); ++i) out(i, j) = J[size_t(j) * out.rows() + i];
    return out;
  }
  VectorXd EvalLagrangeMultipliers(const TrajectoryOptimizerState& st) const { return vec(flat(st, 2, "lambda")); }
  double EvalMeritFunction(const TrajectoryOptimizerState& st) const { return flat(st, 2, "merit")[0]; }
  VectorXd EvalMeritFunctionGradient(const TrajectoryOptimizerState& st) const { return vec(flat(st, 2, "gm")); }

  void ResetInitialConditions(const VectorXd& q_init, const VectorXd& v_init) {  // h:463-468
    if (q_init.size() != plant_->num_positions() || v_init.size() != plant_->num_velocities())
      throw std::runtime_error("ResetInitialConditions: wrong sizes");
    prob_.q_init = q_init, prob_.v_init = v_init;
    ++version_;
  }
  void UpdateNominalTrajectory(const std::vector<VectorXd>& q_nom, const std::vector<VectorXd>& v_nom) {  // h:477-483
    if (int(q_nom.size()) != num_steps() + 1 || q_nom[0].size() != plant_->num_positions())
      throw std::runtime_error("UpdateNominalTrajectory: wrong sizes");
    prob_.q_nom = q_nom, prob_.v_nom = v_nom;
    ++version_;
  }

 private:
  int nq() const { return plant_->num_positions(); }
  int nv() const { return plant_->num_velocities(); }
  // stage: 0 trajectory-level entries, 1 + ID partials, 2 + gradient / Hessian / constraints / multipliers
  std::vector<double> flat(const TrajectoryOptimizerState& st, int stage, const char* name) const {
    idto_solver_t s = st.handle();
    if (!s) throw std::runtime_error("Eval*: empty TrajectoryOptimizerState");
    ThrowOnError(stage == 0 ? idto_eval_trajectory(s) : (stage == 1 ? idto_eval_derivatives(s) : idto_eval_assembly(s)),
                 "idto_eval");
    const long n = idto_field_size(s, name);
    if (n < 0) throw std::runtime_error(std::string("unknown cache entry ") + name);
    std::vector<double> buf(size_t(n > 0 ? n : 1));
    ThrowOnError(idto_get(s, name, buf.data()), "idto_get");
    buf.resize(size_t(n));
    return buf;
  }
  static VectorXd vec(const std::vector<double>& x) {
    VectorXd v(int(x.size()));
    for (size_t i = 0; i < x.size(); ++i) v[int(i)] = x[i];
    return v;
  }
  std::vector<VectorXd> rows(const TrajectoryOptimizerState& st, int stage, const char* name, int n, int w) const {
    const std::vector<double> x = flat(st, stage, name);
    std::vector<VectorXd> out(n, VectorXd(w));
    for (int t = 0; t < n; ++t)
      for (int i = 0; i < w; ++i) out[t][i] = x[size_t(t) * w + i];
    return out;
  }
  static std::vector<MatrixXd> blocks(const std::vector<double>& x, int n, int r, int c) {  // column-major blocks
    std::vector<MatrixXd> out(n, MatrixXd(r, c));
    for (int t = 0; t < n; ++t)
      for (int j = 0; j < c; ++j)
        for (int i = 0; i < r; ++i) out[t](i, j) = x[(size_t(t) * c + j) * r + i];
    return out;
  }
  PentaDiagonalMatrix penta(const TrajectoryOptimizerState& st, const char* a, const char* b, const char* c) const {
    PentaDiagonalMatrix H;
    const int nb = num_steps() + 1, k = nq();
    H.A = blocks(flat(st, 2, a), nb, k, k), H.B = blocks(flat(st, 2, b), nb, k, k), H.C = blocks(flat(st, 2, c), nb, k, k);
    H.D.assign(nb, MatrixXd(k, k)), H.E.assign(nb, MatrixXd(k, k));
    for (int i = 0; i < nb; ++i)
      for (int r = 0; r < k; ++r)
        for (int cc = 0; cc < k; ++cc) {
          if (i + 1 < nb) H.D[i](r, cc) = H.B[i + 1](cc, r);
          if (i + 2 < nb) H.E[i](r, cc) = H.A[i + 2](cc, r);
        }
    return H;
  }
  struct Flat {  // ProblemDefinition -> idto_problem_desc (column-major matrices, flattened trajectories)
    std::vector<double> q_init, v_init, q_nom, v_nom;
    idto_problem_desc desc{};
    Flat(const ProblemDefinition& p, const MultibodyPlant& plant) {
      const int nq = plant.num_positions(), nv = plant.num_velocities(), Tn = p.num_steps;
      q_init.assign(p.q_init.data(), p.q_init.data() + nq), v_init.assign(p.v_init.data(), p.v_init.data() + nv);
      for (int t = 0; t <= Tn; ++t) {
        q_nom.insert(q_nom.end(), p.q_nom[t].data(), p.q_nom[t].data() + nq);
        v_nom.insert(v_nom.end(), p.v_nom[t].data(), p.v_nom[t].data() + nv);
      }
      desc.num_steps = Tn, desc.time_step = plant.time_step();
      desc.q_init = q_init.data(), desc.v_init = v_init.data(), desc.q_nom = q_nom.data(), desc.v_nom = v_nom.data();
      desc.Qq = p.Qq.data(), desc.Qv = p.Qv.data(), desc.Qf_q = p.Qf_q.data(), desc.Qf_v = p.Qf_v.data(), desc.R = p.R.data();
    }
  };
  idto_params c_params() const { return ToC(params_); }

 public:
  using PublicFlat = Flat;
  static idto_params ToC(const SolverParameters& s) {
    idto_params p;
    idto_params_default(&p);
    p.max_iterations = s.max_iterations, p.gradients_method = int(s.gradients_method);
    p.normalize_quaternions = s.normalize_quaternions, p.contact_stiffness = s.contact_stiffness;
    p.dissipation_velocity = s.dissipation_velocity, p.stiction_velocity = s.stiction_velocity;
    p.friction_coefficient = s.friction_coefficient, p.smoothing_factor = s.smoothing_factor;
    p.scaling = s.scaling, p.scaling_method = int(s.scaling_method), p.equality_constraints = s.equality_constraints;
    p.Delta0 = s.Delta0, p.Delta_max = s.Delta_max, p.check_convergence = s.check_convergence;
    if (s.linear_solver == SolverParameters::kDenseLdlt) p.linear_solver = IDTO_LINSOLVE_DENSE_LDLT;  // cc:2088-2093
    if (s.linear_solver == SolverParameters::kCyclicReduction) p.linear_solver = IDTO_LINSOLVE_CYCLIC_REDUCTION;
    const ConvergenceCriteriaTolerances& t = s.convergence_tolerances;
    p.tol_rel_cost_reduction = t.rel_cost_reduction, p.tol_abs_cost_reduction = t.abs_cost_reduction;
    p.tol_rel_gradient_along_dq = t.rel_gradient_along_dq, p.tol_abs_gradient_along_dq = t.abs_gradient_along_dq;
    p.tol_rel_state_change = t.rel_state_change, p.tol_abs_state_change = t.abs_state_change;
    return p;
  }

 private:
  const Diagram<T>* diagram_{nullptr};
  const MultibodyPlant* plant_{nullptr};
  ProblemDefinition prob_;
  const SolverParameters params_;
  idto_model_t model_{nullptr};
  std::vector<int> unactuated_dofs_;
  int version_{0};
};

// ---- batch of independent solves over several GPUs, in one process ------------------------------
// The reference's threading contract: independent optimizers / WarmStarts may run on different threads
// (trajectory_optimizer.h:1096 shared context).  Here: one model + one batched solver per device, one
// host thread per device for every call, contiguous batch slices, no collective — the in-process twin of
// `torchrun bench.py --gpus N`.  Host arrays are indexed by GLOBAL problem.
class MultiGpuBatch {
 public:
  MultiGpuBatch(const MultibodyPlant& plant, const ProblemDefinition& prob, const SolverParameters& params, int batch,
                int num_devices = 0)
      : nq_(plant.num_positions()), nv_(plant.num_velocities()), T_(prob.num_steps), B_(batch) {
    int ndev = idto_device_count();
    if (ndev == 0) throw std::runtime_error("MultiGpuBatch: no CUDA device (there is no CPU fallback)");
    if (num_devices > 0 && num_devices < ndev) ndev = num_devices;
    if (ndev > batch) ndev = batch;
    TrajectoryOptimizer<double> proto_check(nullptr, &plant, prob, params);  // size checks, like a single optimizer
    shards_.resize(ndev);
    run([&](int d) {
      Shard& sh = shards_[d];
      sh.b0 = int(size_t(batch) * d / ndev), sh.b1 = int(size_t(batch) * (d + 1) / ndev);
      ThrowOnError(idto_set_device(d), "idto_set_device");
      ThrowOnError(idto_model_create(&plant.desc(), &sh.model), "idto_model_create");
      typename TrajectoryOptimizer<double>::PublicFlat f(prob, plant);
      const idto_params p = TrajectoryOptimizer<double>::ToC(params);
      ThrowOnError(idto_solver_create(sh.model, &f.desc, &p, sh.b1 - sh.b0, &sh.solver), "idto_solver_create");
    });
  }
  ~MultiGpuBatch() {
    for (Shard& sh : shards_) {
      if (sh.solver) idto_solver_destroy(sh.solver);
      if (sh.model) idto_model_destroy(sh.model);
    }
  }
  MultiGpuBatch(const MultiGpuBatch&) = delete;
  int num_devices() const { return int(shards_.size()); }
  int batch() const { return B_; }
  // q: [batch][(T+1) nq]
  void set_q(const std::vector<double>& q) {
    const size_t w = size_t(T_ + 1) * nq_;
    if (q.size() != w * B_) throw std::runtime_error("MultiGpuBatch::set_q: wrong size");
    run([&](int d) { ThrowOnError(idto_set_q(shards_[d].solver, q.data() + w * shards_[d].b0), "idto_set_q"); });
  }
  void ResetInitialConditions(const std::vector<double>& q_init, const std::vector<double>& v_init) {
    if (q_init.size() != size_t(B_) * nq_ || v_init.size() != size_t(B_) * nv_)
      throw std::runtime_error("MultiGpuBatch::ResetInitialConditions: wrong sizes");
    run([&](int d) {
      const Shard& sh = shards_[d];
      ThrowOnError(idto_reset_initial_conditions(sh.solver, q_init.data() + size_t(sh.b0) * nq_,
                                                 v_init.data() + size_t(sh.b0) * nv_), "idto_reset_initial_conditions");
    });
  }
  // SolveFromWarmStart for every problem: iterations run per problem; stats [batch][max_iterations][IDTO_NUM_STATS]
  std::vector<int> Solve(int max_iterations, std::vector<double>* stats = nullptr) {
    std::vector<int> iters(B_, 0);
    const size_t w = size_t(max_iterations > 0 ? max_iterations : 1) * IDTO_NUM_STATS;
    if (stats) stats->assign(w * B_, 0.0);
    run([&](int d) {
      const Shard& sh = shards_[d];
      ThrowOnError(idto_solve(sh.solver, max_iterations, iters.data() + sh.b0, nullptr,
                              stats ? stats->data() + w * sh.b0 : nullptr), "idto_solve");
    });
    return iters;
  }
  // field of every problem, concatenated in global order (names: idto_get)
  std::vector<double> Get(const char* name) {
    const long n = idto_field_size(shards_[0].solver, name);
    if (n < 0) throw std::runtime_error(std::string("unknown cache entry ") + name);
    std::vector<double> out(size_t(n) * B_);
    run([&](int d) {
      const Shard& sh = shards_[d];
      ThrowOnError(idto_eval_trajectory(sh.solver), "idto_eval_trajectory");
      ThrowOnError(idto_get(sh.solver, name, out.data() + size_t(n) * sh.b0), "idto_get");
    });
    return out;
  }

 private:
  struct Shard {
    idto_model_t model{nullptr};
    idto_solver_t solver{nullptr};
    int b0{0}, b1{0};
  };
  template <class F>
  void run(F f) {  // one host thread per device; exceptions are re-thrown on the caller's thread
    std::vector<std::thread> th;
    std::vector<std::string> err(shards_.size());
    for (int d = 0; d < int(shards_.size()); ++d)
      th.emplace_back([&, d] {
        try {
          f(d);
        } catch (const std::exception& e) {
          err[d] = e.what()[0] ? e.what() : "error";
        }
      });
    for (auto& t : th) t.join();
    for (const std::string& e : err)
      if (!e.empty()) throw std::runtime_error(e);
  }
  std::vector<Shard> shards_;
  int nq_, nv_, T_, B_;
};

}  // namespace optimizer
}  // namespace idto
