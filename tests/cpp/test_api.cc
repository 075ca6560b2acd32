// C++ twin of python_bindings/test/trajectory_optimizer_test.py written against include/idto_b200.hpp:
// same class names and call sequence as code using the reference's optimizer/*.h.
//   usage: test_api <spinner.txt>      (baked table from BakedModel.save_txt)
#include <cstdio>
#include <cstdlib>

#include "idto_b200.hpp"

using namespace idto::optimizer;

#define CHECK(c)                                                   \
  do {                                                             \
    if (!(c)) {                                                    \
      std::printf("CHECK failed: %s (line %d)\n", #c, __LINE__);   \
      return 1;                                                    \
    }                                                              \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const double time_step = 0.05;
  MultibodyPlant plant = MultibodyPlant::LoadBaked(argv[1], time_step);
  Diagram<double>* diagram = nullptr;

  ProblemDefinition problem;
  problem.num_steps = 40;
  problem.q_init = VectorXd{0.3, 1.5, 0.0};
  problem.v_init = VectorXd{0.0, 0.0, 0.0};
  problem.Qq = 1.0 * MatrixXd::Identity(3, 3);
  problem.Qv = 0.1 * MatrixXd::Identity(3, 3);
  problem.R = MatrixXd::Diagonal({0.1, 0.1, 1e3});
  problem.Qf_q = 10 * MatrixXd::Identity(3, 3);
  problem.Qf_v = 0.1 * MatrixXd::Identity(3, 3);
  for (int i = 0; i <= problem.num_steps; ++i) {
    problem.q_nom.push_back(VectorXd{0.3, 1.5, 2.0});
    problem.v_nom.push_back(VectorXd{0.0, 0.0, 0.0});
  }
  SolverParameters params;
  params.max_iterations = 200;
  params.scaling = true, params.equality_constraints = true;
  params.Delta0 = 1e1, params.Delta_max = 1e5;
  params.contact_stiffness = 200, params.dissipation_velocity = 0.1, params.smoothing_factor = 0.01;
  params.friction_coefficient = 0.5, params.stiction_velocity = 0.05, params.verbose = false;
  std::vector<VectorXd> q_guess(problem.num_steps + 1, VectorXd{0.3, 1.5, 0.0});

  TrajectoryOptimizer<double> opt(diagram, &plant, problem, params);
  CHECK(opt.time_step() == time_step && opt.num_steps() == 40);
  CHECK(opt.unactuated_dofs().size() == 1 && opt.unactuated_dofs()[0] == 2);
  CHECK(opt.num_equality_constraints() == 40);

  TrajectoryOptimizerSolution<double> solution;
  TrajectoryOptimizerStats<double> stats;
  ConvergenceReason reason;
  const SolverFlag flag = opt.Solve(q_guess, &solution, &stats, &reason);
  CHECK(flag == SolverFlag::kMaxIterationsReached);
  CHECK(solution.q.size() == 41 && solution.v.size() == 41 && solution.tau.size() == 40);
  CHECK(stats.iteration_costs.size() == 200 && !stats.is_empty());
  const double e0 = solution.q[40][0] - 0.287, e1 = solution.q[40][1] - 1.497, e2 = solution.q[40][2] - 1.995;
  std::printf("q_T = %.5f %.5f %.5f  cost %.6f\n", solution.q[40][0], solution.q[40][1], solution.q[40][2],
              stats.iteration_costs.back());
  CHECK(std::sqrt(e0 * e0 + e1 * e1 + e2 * e2) < 1e-3);  // python_bindings/test/trajectory_optimizer_test.py:84-85

  // warm start: 10 x SolveFromWarmStart(max_iterations = 1) appends stats, Delta persists (warm_start_test.py)
  SolverParameters p1 = params;
  p1.max_iterations = 1;
  TrajectoryOptimizer<double> opt1(diagram, &plant, problem, p1);
  auto ws = opt1.CreateWarmStart(q_guess);
  TrajectoryOptimizerStats<double> st1;
  for (int k = 0; k < 10; ++k) opt1.SolveFromWarmStart(ws.get(), &solution, &st1);
  CHECK(st1.iteration_costs.size() == 10);
  for (int k = 0; k < 10; ++k) CHECK(std::fabs(st1.iteration_costs[k] - stats.iteration_costs[k]) < 1e-8);
  CHECK(ws->Delta() > 0 && ws->get_q().size() == 41);

  // MPC mutators (h:463-483)
  opt1.ResetInitialConditions(VectorXd{0.31, 1.49, 0.01}, VectorXd{0.0, 0.0, 0.0});
  std::vector<VectorXd> qn = problem.q_nom;
  qn[10] = VectorXd{0.4, 1.6, 2.1};
  opt1.UpdateNominalTrajectory(qn, problem.v_nom);
  CHECK(opt1.prob().q_nom[10][2] == 2.1 && opt1.prob().q_init[0] == 0.31);
  opt1.SolveFromWarmStart(ws.get(), &solution, &st1);
  CHECK(st1.iteration_costs.size() == 11);
  CHECK(DecodeConvergenceReasons(kNoConvergenceCriteriaSatisfied) == "no convergence criterion satisfied");
  bool threw = false;
  try {
    TrajectoryOptimizerStats<double> dirty = stats;
    opt.Solve(q_guess, &solution, &dirty);
  } catch (const std::runtime_error&) {
    threw = true;  // stats must be empty (cc:2225)
  }
  CHECK(threw);

  // MPC shell on the device (examples/mpc_controller.cc:43-98): with zero elapsed time the shifted guess is
  // the stored solution at its knots, with q_0 replaced by the measured state
  const std::vector<VectorXd> before = ws->get_q();
  ws->AdvanceFromMeasuredState(0.0, VectorXd{0.32, 1.48, 0.02}, VectorXd{0.0, 0.0, 0.0}, {false, false, true});
  const std::vector<VectorXd> after = ws->get_q();
  CHECK(after[0][0] == 0.32 && after[0][2] == 0.02);
  for (int t = 1; t <= 40; ++t)
    for (int i = 0; i < 3; ++i) CHECK(std::fabs(after[t][i] - before[t][i]) < 1e-12);

  // Eval* accessors on the warm start's state (trajectory_optimizer.h:125-453)
  {
    const TrajectoryOptimizerState& st = ws->state;
    const std::vector<VectorXd> v = opt1.EvalV(st), tau = opt1.EvalTau(st);
    CHECK(v.size() == 41 && tau.size() == 40 && v[0].size() == 3);
    const VelocityPartials vp = opt1.EvalVelocityPartials(st);  // cc:962-973: N+ = I for 1-dof joints
    CHECK(vp.dvt_dqt.size() == 41 && std::fabs(vp.dvt_dqt[3](1, 1) - 1.0 / time_step) < 1e-12);
    CHECK(std::isnan(vp.dvt_dqm[0](0, 0)) && std::fabs(vp.dvt_dqm[3](2, 2) + 1.0 / time_step) < 1e-12);
    CHECK(vp.dvt_dqt[3](0, 1) == 0.0);
    const InverseDynamicsPartials idp = opt1.EvalInverseDynamicsPartials(st);
    CHECK(idp.dtau_dqp.size() == 40 && idp.dtau_dqp[0].rows() == 3 && std::isnan(idp.dtau_dqm[0](0, 0)));
    const PentaDiagonalMatrix H = opt1.EvalHessian(st);
    CHECK(H.block_rows() == 41 && H.block_size() == 3 && H.C[0](0, 0) == 1.0 && H.C[0](1, 0) == 0.0);  // C_0 = I
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) CHECK(H.D[4](r, c) == H.B[5](c, r) && H.E[4](r, c) == H.A[6](c, r));
    const VectorXd g = opt1.EvalGradient(st), lam = opt1.EvalLagrangeMultipliers(st);
    CHECK(g.size() == 41 * 3 && g[0] == 0.0 && lam.size() == 40);
    const MatrixXd J = opt1.EvalEqualityConstraintJacobian(st);
    CHECK(J.rows() == 40 && J.cols() == 123);
    CHECK(std::fabs(opt1.EvalMeritFunction(st) - (opt1.EvalCost(st) + [&] {
            const VectorXd h = opt1.EvalEqualityConstraintViolations(st);
            double s = 0;
            for (int i = 0; i < 40; ++i) s += h[i] * lam[i];
            return s;
          }())) < 1e-9 * std::fmax(1.0, std::fabs(opt1.EvalCost(st))));
  }

  // kDenseLdlt (cc:2088-2093) through the C++ API: same iterates as the penta-diagonal solver
  {
    SolverParameters pd = params;
    pd.max_iterations = 5, pd.linear_solver = SolverParameters::kDenseLdlt;
    TrajectoryOptimizer<double> optd(diagram, &plant, problem, pd);
    TrajectoryOptimizerSolution<double> sd;
    TrajectoryOptimizerStats<double> std_;
    optd.Solve(q_guess, &sd, &std_);
    for (int k = 0; k < 5; ++k)
      CHECK(std::fabs(std_.iteration_costs[k] - stats.iteration_costs[k]) < 1e-6 * std::fabs(stats.iteration_costs[k]));
  }

  // kCyclicReduction (this build's parallel-in-time order of the same elimination): same iterates too
  {
    SolverParameters pc = params;
    pc.max_iterations = 5, pc.linear_solver = SolverParameters::kCyclicReduction;
    TrajectoryOptimizer<double> optc(diagram, &plant, problem, pc);
    TrajectoryOptimizerSolution<double> sc_;
    TrajectoryOptimizerStats<double> stc;
    optc.Solve(q_guess, &sc_, &stc);
    for (int k = 0; k < 5; ++k)
      CHECK(std::fabs(stc.iteration_costs[k] - stats.iteration_costs[k]) < 1e-6 * std::fabs(stats.iteration_costs[k]));
  }

  // a batch over every GPU of the process (one host thread per device, no collective): each problem equals the
  // single solve above
  {
    SolverParameters pb = params;
    pb.max_iterations = 4;
    const int B = 5;
    MultiGpuBatch mb(plant, problem, pb, B);
    std::vector<double> q0(size_t(B) * 41 * 3);
    for (int b = 0; b < B; ++b)
      for (int t = 0; t <= 40; ++t)
        for (int i = 0; i < 3; ++i) q0[(size_t(b) * 41 + t) * 3 + i] = q_guess[t][i];
    mb.set_q(q0);
    std::vector<double> st;
    const std::vector<int> it = mb.Solve(4, &st);
    CHECK(mb.num_devices() >= 1 && int(it.size()) == B);
    for (int b = 0; b < B; ++b) {
      CHECK(it[b] == 4);
      for (int k = 0; k < 4; ++k)
        CHECK(std::fabs(st[(size_t(b) * 4 + k) * IDTO_NUM_STATS] - stats.iteration_costs[k]) < 1e-9 * stats.iteration_costs[k]);
    }
    CHECK(mb.Get("q").size() == size_t(B) * 41 * 3);
    std::printf("MultiGpuBatch: %d problems on %d device(s)\n", B, mb.num_devices());
  }
  std::printf("C++ API test OK\n");
  return 0;
}
