/* idto_b200 — C ABI of the B200-native IDTO hot path.
 *
 * The reference (ToyotaResearchInstitute/idto) has no FFI/plugin boundary: its
 * boundary is the C++ class `TrajectoryOptimizer<double>` + the pybind11 module
 * `pyidto`.  This header is the thin C-ABI CUDA layer that the
 * host C++ / Python mirrors of that class sit on.  Entry points are at the
 * granularity of the reference's cache entries (optimizer/trajectory_optimizer_state.h:333-350):
 * each `idto_eval_*`/`idto_get` pair replaces one `Eval*` accessor, and
 * `idto_solve` replaces `SolveFromWarmStart` (optimizer/trajectory_optimizer.cc:2449-2651)
 * for a batch of independent warm starts.
 *
 * Conventions: plain pointers and sizes, all matrices column-major fp64 unless
 * stated, `int` status return (0 = ok, <0 = error — replaces the reference's
 * DRAKE_DEMAND aborts), caller-owned opaque handles, no hidden global state.
 * Every handle is bound to the CUDA device current at creation time.
 */
#ifndef IDTO_B200_H_
#define IDTO_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
enum {
  IDTO_OK = 0,
  IDTO_ERR_INVALID_ARG = -1,   /* size/shape check failed (reference: DRAKE_DEMAND) */
  IDTO_ERR_UNSUPPORTED = -2,   /* geometry pair / joint / option outside the baked set */
  IDTO_ERR_CUDA = -3,          /* CUDA runtime error; see idto_last_error() */
  IDTO_ERR_FACTORIZATION = -4, /* penta-diagonal / dense factorisation failed */
  IDTO_ERR_NO_DEVICE = -5,     /* no CUDA device: there is NO CPU fallback */
  IDTO_ERR_CONTACT_OVERFLOW = -6 /* more contact pairs active at once than an evaluation's list holds (models with
                                  * > 32 candidate pairs: 64 slots, or IDTO_MAX_ACTIVE_PAIRS at model creation) */
};

/* ------------------------------------------------------- baked model tables
 * Stand-in for the one-time query of Drake's MultibodyPlant + SceneGraph
 * (reference call sites: trajectory_optimizer.cc:51-72, :272-326, :1645).
 * Bodies are the *moving* mobilized bodies in Drake's depth-first dof order;
 * welded bodies are merged into their parent at bake time.  Parent -1 = world.
 */
enum { IDTO_JOINT_REVOLUTE = 0, IDTO_JOINT_PRISMATIC = 1, IDTO_JOINT_PLANAR = 2,
       IDTO_JOINT_QUAT_FLOATING = 3 };
/* geom_dims: sphere (radius, -, -); box (full sizes x, y, z); capsule / cylinder (radius, length, -), axis = Gz;
 * half space (-, -, -): the region z <= 0 of its frame G, outward normal +Gz (drake::geometry::HalfSpace) */
enum { IDTO_GEOM_SPHERE = 0, IDTO_GEOM_BOX = 1, IDTO_GEOM_CAPSULE = 2, IDTO_GEOM_CYLINDER = 3,
       IDTO_GEOM_HALF_SPACE = 4 };

typedef struct {
  int nbodies, nq, nv;
  const int* parent;       /* [nbodies] */
  const int* joint_type;   /* [nbodies] */
  const int* q_start;      /* [nbodies] */
  const int* v_start;      /* [nbodies] */
  const double* X_PF;      /* [nbodies][12]: R row-major (9) then p (3); F fixed on parent body P (or world) */
  const double* R_MB;      /* [nbodies][9]: rotation from body frame B to mobilized frame M (row-major R_MB); p_MoBo = 0 */
  const double* axis;      /* [nbodies][3] unit axis in F (revolute/prismatic) */
  const double* damping;   /* [nv] viscous joint damping */
  const double* mass;      /* [nbodies] */
  const double* com;       /* [nbodies][3] p_BoBcm_B */
  const double* inertia;   /* [nbodies][6] I_BBo_B: xx yy zz xy xz yz */
  double gravity[3];       /* world frame, default (0,0,-9.81) */
  const int* actuated;     /* [nv] 1 if some actuator drives this dof */
  int ngeoms;
  const int* geom_body;    /* [ngeoms] moving-body index or -1 (world) */
  const int* geom_type;    /* [ngeoms] */
  const double* geom_dims; /* [ngeoms][3]: sphere r,0,0 ; box full sizes */
  const double* X_BG;      /* [ngeoms][12] pose of geometry in its (merged) body */
  int npairs;              /* candidate pairs after default filtering, sorted by (idA,idB) */
  const int* pair_geomA;   /* [npairs] registration index, A < B */
  const int* pair_geomB;   /* [npairs] */
} idto_model_desc;

/* ---------------------------------------------------------- problem + params
 * Mirrors ProblemDefinition (optimizer/problem_definition.h:24-59) and the
 * hot-path subset of SolverParameters (optimizer/solver_parameters.h:64-167).
 */
typedef struct {
  int num_steps;           /* T */
  double time_step;        /* plant.time_step() */
  const double* q_init;    /* [nq] */
  const double* v_init;    /* [nv] */
  const double* Qq;        /* [nq*nq] col-major */
  const double* Qv;        /* [nv*nv] */
  const double* Qf_q;      /* [nq*nq] */
  const double* Qf_v;      /* [nv*nv] */
  const double* R;         /* [nv*nv] */
  const double* q_nom;     /* [(T+1)*nq] */
  const double* v_nom;     /* [(T+1)*nv] */
} idto_problem_desc;

enum { IDTO_GRAD_FORWARD = 0, IDTO_GRAD_CENTRAL = 1, IDTO_GRAD_CENTRAL4 = 2 };
enum { IDTO_SCALING_SQRT = 0, IDTO_SCALING_ADAPTIVE_SQRT = 1,
       IDTO_SCALING_DOUBLE_SQRT = 2, IDTO_SCALING_ADAPTIVE_DOUBLE_SQRT = 3 };
/* single top-down block-Thomas sweep, or two-sided ("twisted") elimination by a 2-CTA cluster */
/* TWISTED / THOMAS: the block penta-diagonal solver (reference: kPentaDiagonalLu) as a two-sided or a single
 * top-down sweep; DENSE_LDLT: the reference's debugging cross-check kDenseLdlt (trajectory_optimizer.cc:2088-2093):
 * the Gauss-Newton step is re-solved by a structure-agnostic LDL^T of H~; CYCLIC_REDUCTION: the same block system
 * (kPentaDiagonalLu) eliminated by parallel block cyclic reduction instead of a sweep (same results to the
 * conditioning of the system; slower than the sweep for T <= 60 at batch 64). */
enum { IDTO_LINSOLVE_THOMAS = 0, IDTO_LINSOLVE_TWISTED = 1, IDTO_LINSOLVE_DENSE_LDLT = 2,
       IDTO_LINSOLVE_CYCLIC_REDUCTION = 3 };

typedef struct {
  int max_iterations;         /* 100 */
  int gradients_method;       /* IDTO_GRAD_FORWARD */
  int normalize_quaternions;  /* 0 */
  double contact_stiffness;   /* 100 */
  double dissipation_velocity;/* 0.1 */
  double stiction_velocity;   /* 0.05 */
  double friction_coefficient;/* 0.5 */
  double smoothing_factor;    /* 0.1 */
  int scaling;                /* 1 */
  int scaling_method;         /* IDTO_SCALING_DOUBLE_SQRT */
  int equality_constraints;   /* 1 */
  double Delta0;              /* 1e-1 */
  double Delta_max;           /* 1e5 */
  int check_convergence;      /* 0 */
  double tol_rel_cost_reduction, tol_abs_cost_reduction;
  double tol_rel_gradient_along_dq, tol_abs_gradient_along_dq;
  double tol_rel_state_change, tol_abs_state_change;
  int linear_solver;          /* IDTO_LINSOLVE_TWISTED (build-specific; reference: kPentaDiagonalLu) */
} idto_params;

/* Fills `p` with the reference defaults (solver_parameters.h:64-167). */
void idto_params_default(idto_params* p);

/* ------------------------------------------------------------------ handles */
typedef struct idto_model_s* idto_model_t;
typedef struct idto_solver_s* idto_solver_t;

const char* idto_last_error(void);
int idto_device_count(void);
/* Makes CUDA device `device` current for the calling thread (models are created on the current device): lets a
 * host program without the CUDA runtime headers drive one solver per GPU (one host thread + stream
 * per device, no collective). */
int idto_set_device(int device);

/* Copies the tables to the current device.  Replaces the one-time plant query
 * in the TrajectoryOptimizer constructor (trajectory_optimizer.cc:43-72). */
int idto_model_create(const idto_model_desc* desc, idto_model_t* out);
int idto_model_destroy(idto_model_t m);
int idto_model_num_unactuated(idto_model_t m);
/* out[num_unactuated]: velocity indices of unactuated dofs (trajectory_optimizer.cc:63-72). */
int idto_model_unactuated_dofs(idto_model_t m, int* out);

/* A solver lives on the device that was current when its model was created; every entry point taking a
 * solver makes that device current for the calling thread (one host thread per device may drive several
 * solvers concurrently; one solver is not re-entrant, like the reference's TrajectoryOptimizer).
 *
 * A solver = one TrajectoryOptimizer bound to `batch` independent WarmStarts
 * (state + scratch_state + Delta + dq + dqH, optimizer/warm_start.h:23-76),
 * all device resident.  `prob` is the template problem; per-batch-element
 * q_init/v_init/q_nom/v_nom are set with the calls below. */
int idto_solver_create(idto_model_t m, const idto_problem_desc* prob,
                       const idto_params* params, int batch, idto_solver_t* out);
int idto_solver_destroy(idto_solver_t s);

/* Number of internal sub-batch streams (1..16; default 4 when batch >= 32, else 1): the batch is cut
 * into contiguous sub-batches whose kernels overlap (latency-bound KKT sweeps of one sub-batch with
 * the throughput-bound ID-partials kernels of another).  Results do not depend on it. */
int idto_solver_set_substreams(idto_solver_t s, int n);
/* stream: a cudaStream_t passed as void* (0 = default stream). */
int idto_solver_set_stream(idto_solver_t s, void* stream);

/* WarmStart::set_q (warm_start.h:55). q: host [batch][(T+1)*nq]. Invalidates all caches. */
int idto_set_q(idto_solver_t s, const double* q_host);
/* Marks every cache entry stale without changing q (== WarmStart::set_q(get_q()); what an MPC
 * re-solve does after shifting its guess, examples/mpc_controller.cc:56-59). */
int idto_invalidate(idto_solver_t s);
/* ResetInitialConditions (trajectory_optimizer.h:463-468): host [batch][nq], [batch][nv]. */
int idto_reset_initial_conditions(idto_solver_t s, const double* q_init, const double* v_init);
/* UpdateNominalTrajectory (trajectory_optimizer.h:477-483): host [batch][(T+1)*nq], [batch][(T+1)*nv]. */
int idto_update_nominal_trajectory(idto_solver_t s, const double* q_nom, const double* v_nom);
/* WarmStart::Delta; host [batch]. */
int idto_set_delta(idto_solver_t s, const double* delta);
int idto_get_delta(idto_solver_t s, double* delta);

/* Cache-entry evaluators: compute (if stale) on the device; results stay there.
 *   idto_eval_trajectory : EvalNplus, EvalV, EvalA, EvalTau, EvalCost, EvalEqualityConstraintViolations
 *   idto_eval_derivatives: EvalInverseDynamicsPartials (+ velocity partials, implicit ±N+/dt)
 *   idto_eval_assembly   : EvalGradient, EvalHessian, EvalScaleFactors, EvalScaledHessian/Gradient,
 *                          EvalEqualityConstraintJacobian, EvalLagrangeMultipliers,
 *                          EvalMeritFunction, EvalMeritFunctionGradient
 *   idto_eval_dogleg     : CalcDoglegPoint (dq, dqH) with the current Delta
 *   idto_eval_trust_ratio: CalcTrustRatio for the current dq (uses the scratch state)
 */
int idto_eval_trajectory(idto_solver_t s);
int idto_eval_derivatives(idto_solver_t s);
int idto_eval_assembly(idto_solver_t s);
int idto_eval_dogleg(idto_solver_t s);
int idto_eval_trust_ratio(idto_solver_t s);

/* Device → host mirror of a named cache entry for all batch elements.
 * `out` must hold idto_field_size(field) * batch doubles.  Names:
 *  q v a tau Nplus cost h dtau_dqm dtau_dqt dtau_dqp g H_A H_B H_C D Hs_A Hs_B Hs_C gs
 *  J lambda merit gm dq dqH dq_active rho delta q_nom v_nom
 *  dvt_dqt dvt_dqm  (EvalVelocityPartials, velocity_partials.h:19-39: +N+_t/dt, -N+_t/dt with [0] = NaN;
 *                    never materialised on the device, formed on the host from Nplus)
 *  pair_active pair_active_fd  (only after idto_debug_pair_trace, see below)
 * Layouts follow the reference containers: per time step, column-major blocks. */
long idto_field_size(idto_solver_t s, const char* field);
int idto_get(idto_solver_t s, const char* field, double* out);

/* SolveFromWarmStart for every batch element (trajectory_optimizer.cc:2449-2651):
 * runs `max_iterations` trust-region iterations (or until convergence) entirely
 * on the device.  Stats are appended per batch element, 13 series in push_data
 * order (trajectory_optimizer_solution.h:95-111) minus the two linesearch ones:
 *   stats[b][iter][k], k = 0 cost, 1 delta, 2 q_norm, 3 dq_norm, 4 dqH_norm,
 *   5 trust_ratio, 6 grad_norm, 7 dL_dq, 8 h_norm, 9 merit
 * iters_out[b] = iterations actually run, reason_out[b] = ConvergenceReason bits.
 * Any of the output pointers may be NULL. */
#define IDTO_NUM_STATS 10
int idto_solve(idto_solver_t s, int max_iterations, int* iters_out, int* reason_out,
               double* stats_out /* [batch][max_iterations][IDTO_NUM_STATS] */);

/* Same, but asynchronous on the solver's stream with caller-pinned host
 * buffers: H2D of the guess/initial conditions, the iterations, and D2H of the
 * solution are all enqueued; call idto_synchronize() to wait.  This is the
 * end-to-end MPC re-solve call (examples/mpc_controller.cc:43-85).
 * Any input pointer may be NULL to keep the device-resident value.
 * iters_out (pinned host [batch], may be NULL) receives the number of iterations recorded per problem:
 * with check_convergence a problem may stop early, and rows of stats_out beyond iters_out[b] are not
 * written by this call.
 * Host buffers may be pageable or pinned.  Pinned (page-locked, mapped) buffers are read and written by kernels
 * directly through their device aliases — no staging copies; pageable ones go through cudaMemcpyAsync.  Either way
 * they must not be touched before idto_synchronize(), and a buffer that is passed again at the same address must
 * still be pinned (or still be pageable): the call sequence of a given argument set is captured once (CUDA graph)
 * and replayed. */
int idto_resolve_async(idto_solver_t s, int max_iterations,
                       const double* q_guess, const double* q_init, const double* v_init,
                       const double* q_nom, const double* v_nom,
                       double* q_out, double* v_out, double* tau_out,
                       int* iters_out, double* stats_out);
int idto_synchronize(idto_solver_t s);
/* MPC shell between two re-solves, on the device (ModelPredictiveController::UpdateAbstractState,
 * examples/mpc_controller.cc:43-98; python_examples/mpc_utils.py:183-217), for every batch element b:
 *   q_guess_i = spline(elapsed[b] + i dt), i = 1..T, where spline is the C2 cubic (not-a-knot) interpolant of
 *               the current q_0..q_T (PiecewisePolynomial::CubicWithContinuousSecondDerivatives,
 *               mpc_controller.cc:129-137), evaluation clamped to [0, T dt];  q_guess_0 = q0[b];
 *   q_nom_t  += q_nom_selector o (q0[b] - q_nom_0)   (mpc_controller.cc:62-69; NULL: no shift);
 *   ResetInitialConditions(q0[b], v0[b]); all cache entries stale.
 * Host pointers: elapsed [batch], q0 [batch][nq], v0 [batch][nv], q_nom_selector [nq] (0/1 as doubles).
 * Asynchronous on the solver's stream; follow with idto_resolve_async(..., all inputs NULL). */
int idto_mpc_advance(idto_solver_t s, const double* elapsed, const double* q0, const double* v0,
                     const double* q_nom_selector);
/* One MPC re-plan in ONE call: idto_mpc_advance(elapsed, q0, v0, q_nom_selector) followed by
 * idto_resolve_async(max_iterations, no new inputs, outputs) — the whole of
 * ModelPredictiveController::UpdateAbstractState (examples/mpc_controller.cc:43-85).  Same results as the two calls;
 * with pinned input buffers the advance is part of the re-solve's captured graph (one launch from the host). */
int idto_mpc_resolve_async(idto_solver_t s, const double* elapsed, const double* q0, const double* v0,
                           const double* q_nom_selector, int max_iterations,
                           double* q_out, double* v_out, double* tau_out, int* iters_out, double* stats_out);
/* Stream-ordering point without a host wait: everything enqueued so far (on the internal sub-batch
 * streams too) is ordered before whatever the caller enqueues next on the solver's stream, e.g. a
 * CUDA event record. */
int idto_fence(idto_solver_t s);
/* Benchmark hygiene: overwrite `bytes` of caller-provided device scratch (larger than L2) in the
 * solver's stream order, so that the next re-solve starts with a cold L2. */
int idto_flush_l2(idto_solver_t s, void* scratch, size_t bytes);

/* Debug trace of the contact-pair indexing (trajectory_optimizer.cc:272-279: the pairs
 * ComputeSignedDistancePairwiseClosestPoints(threshold) returns are the ones a force is computed for).
 * After enabling, the inverse-dynamics kernels record which candidate pairs each evaluation applies a
 * force for, and idto_get serves
 *   "pair_active"    [batch][T][npairs]        tau_t of the current trajectory: 1 = force computed, 0 = not
 *   "pair_active_fd" [batch][T][nq][4][npairs] the perturbed evaluations of tau_{t-1} at q_t + m dq e_i, stencil
 *                    point m = +1, -1, +2, -2 (the ones the gradients_method uses); -1 = the evaluation did not
 *                    visit the pair (subtree-only evaluations keep the base force of pairs outside the subtree)
 * (as doubles).  Tracing forces every evaluation to run and disables the sub-batch streams: debugging only. */
int idto_debug_pair_trace(idto_solver_t s, int enable);

/* Number of kernel launches issued by this solver since creation. */
long idto_launch_count(idto_solver_t s);

/* Timing hook for bench.py: average device time (ms) of the ID-partials kernel
 * over the launches since the last reset, measured with CUDA events on the
 * solver's stream. */
int idto_profile_enable(idto_solver_t s, int enable);
int idto_profile_read(idto_solver_t s, const char* kernel, double* total_ms, long* launches);

#ifdef __cplusplus
}
#endif
#endif /* IDTO_B200_H_ */
