// Host/device shared declarations of the idto_b200 CUDA layer.
//
// Data layout in HBM (all fp64, batch-major; one "problem" = one WarmStart of the reference,
// optimizer/warm_start.h:23-76).  T = num_steps, n = (T+1)*nq, nh = nu*T.
//   q     [B][T+1][nq]      v [B][T+1][nv]     a, tau [B][T][nv]     Nplus [B][T+1][nv*nq] (col-major)
//   dtau_dqm/dqt/dqp [B][T][nv*nq] (col-major nv x nq blocks, inverse_dynamics_partials.h:20-85)
//   g, D, gs, gm, dq, dqH, pH ... [B][n]
//   H bands A,B,C (unscaled) and scaled copies [B][T+1][nq*nq] (col-major blocks, lower bands only;
//     D_i = B_{i+1}^T and E_i = A_{i+2}^T are never materialised)
//   J~ as three row bands Jm,Jt,Jp [B][T][nu*nq] (row-major nu x nq: the rows of the ID partials of
//     the unactuated dofs times D), instead of the reference's dense (nu*T) x n matrix
//   factor: K, Ginv, Y, Z [B][T+1][nq*nq];  X = H~^-1 J~^T [B][nh][n];  S [B][nh*nh]
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/idto_b200.h"
#include <atomic>

#include "common.cuh"

namespace idto {

constexpr int kMaxChildren = 8;
constexpr int kMaxGroup = 32;
constexpr int kMaxLevels = 8;    // tree depth supported by the chain-lane kernels
// Contact pairs one inverse-dynamics evaluation can have ACTIVE (signed distance <= threshold, cc:268-275) at the
// same time when the model has more candidate pairs than this: each evaluation then compacts the active pairs,
// in candidate order, into a list of this capacity (dynamics_chain.cuh) instead of keeping every candidate.
constexpr int kMaxActivePairs = 32;   // models with more candidate pairs than this compact their active pairs
constexpr int kDefaultPairSlots = 64; // ... into DevModel::nact slots per evaluation (IDTO_MAX_ACTIVE_PAIRS overrides)
// Near list of a pose (pruned models): [0] count, [1..] candidate indices, ascending, of the pairs whose signed
// distance at the UNPERTURBED pose is within the activation distance plus a margin that bounds what a finite-
// difference perturbation can move (k_tau_chain writes it while it walks every candidate anyway; the perturbed
// evaluations of the ID partials walk it instead of the candidates).
__host__ __device__ inline int near_stride(int nact) { return nact + 1; }

// Baked model on the device: two SoA tables (ints, doubles) copied to shared memory by TMA.
struct DevModel {
  const int* itab;
  const double* dtab;
  int itab_bytes, dtab_bytes;  // multiples of 16
  int nb, nbp, nq, nv, ng, np, npp, nlevels, group;  // nbp/npp: padded strides; group: lanes per evaluation
  int nact, prune;  // per-evaluation pair slots: npp, or (np > kMaxActivePairs: prune = 1) the capacity of the
                    // compacted active list (kDefaultPairSlots unless IDTO_MAX_ACTIVE_PAIRS says otherwise)
  double reach;     // bound on the lever arm of any joint on any geometry centre at q = 0 (near-list margin)
  // chain decomposition (lane = kinematic chain, step = tree level): cgroup lanes per evaluation
  int cgroup, nchains, ngb, ngd;  // ngb: bodies carrying geometry, ngd: geometries on moving bodies
  int chain_ok;                   // the chain-lane kernels support this model
  int cg_res;                     // per-group smem stride is padded to cg_res (mod 16 doubles): bank spread
  // Column split of the ID partials: "path" columns (1-dof joint whose subtree is a simple chain touching only
  // world-anchored geometry) are differentiated by single-lane subtree evaluations (kernels_path.cu), the rest
  // ("full" columns) by full evaluations (kernels_chain.cu).  npath + nfull == nq.
  int npath, nfull, o_pathcols, o_fullcols;
  double gx, gy, gz;
  // int table offsets (in ints)
  int o_parent, o_jtype, o_qs, o_vs, o_level, o_nchild, o_child, o_flags, o_qowner, o_gbody, o_gtype, o_pA, o_pB;
  int o_levbody, o_levcross, o_plane, o_gslot, o_gdyn, o_bchain;  // chain tables
  // double table offsets (in doubles)
  int o_XPF, o_RMB, o_axis, o_mass, o_com, o_inertia, o_damping, o_gdims, o_XBG;
  int o_XWGs;  // world pose of world-anchored geometries [12][ng]
};

// Constants of one TrajectoryOptimizer (problem + params), uniform across the batch.
struct SolverConsts {
  int B, T, nq, nv, nu, n, nh;
  double dt;
  int method, scaling, scaling_method, eq, normalize_quat, check_convergence, linear_solver;
  double k, sigma, vd, vs, mu, threshold;  // contact (cc:257-269)
  double Delta_max;
  double tol[6];
  const double *Qq, *Qv, *Qfq, *Qfv, *R;  // diagonals, device [nq]/[nv]
  // ProblemDefinition carries dense matrices (problem_definition.h:38-52); every example's are diagonal.  When one
  // of them is not, dense_w = 1 and the full column-major matrices below feed the general (slower) cost /
  // gradient / Hessian code paths (k_assemble_dense, the dense branch of the cost terms).
  int dense_w;
  const double *QqM, *QvM, *QfqM, *QfvM, *RM;
  const int* unact;                       // device [nu]
  const int* quat_starts;                 // device [nquat]
  int nquat;
  int* status;                            // == SolverBufs::status (sticky error flag), for device functions
};

// One TrajectoryOptimizerState's trajectory-level cache (state.h:37-351) for the batch.
struct TrajBuf {
  double *q, *v, *a, *tau, *Nplus, *cost, *h;
  int* near;  // [B][T][near_stride(nact)] near list of pose q_{t+1} (pruned models only, else null)
};

// Per-problem trust-region control block (device resident; host never reads it mid-solve).
struct ProbCtl {
  double Delta, rho, prev_cost, merit, cost_kp, gnorm, hnorm, dq_norm, dqH_norm, q_norm, dL_dq;
  double ht, gt;      // s.H~s and gm.s of the current dogleg step (k_dogleg_post), for k_trust_final (cc:2008-2017)
  double Delta_prev;  // Delta before the last update: restored when that step turns out to have converged (cc:2601-2622)
  int traj_dirty;    // q changed: v, a, tau, cost, h stale
  int derivs_dirty;  // partials / g / H / D / factor / lambda stale
  int active;        // still iterating (not converged)
  int tr_active;     // dogleg hit the trust-region boundary
  int iters;         // iterations recorded so far in this idto_solve call
  int reason;        // ConvergenceReason bits
  int pending;       // convergence check of the last accepted step still to be evaluated
  int stash_sel;     // which half of SolverBufs::stash holds the records of the STATE trajectory (the other half
                     // receives those of the scratch trajectory; an accepted step flips it)
};

struct SolverBufs {
  TrajBuf st, sc;  // state, scratch_state
  const double *q_init, *v_init, *q_nom, *v_nom;
  double *dqm, *dqt, *dqp;
  double *g, *D, *gs, *gm, *lambda, *merit;
  double *HA, *HB, *HC, *SA, *SB, *SC;  // unscaled / scaled lower bands
  double *Jm, *Jt, *Jp;                 // scaled constraint Jacobian bands
  double *FK, *FG, *FY, *FZ;            // block-Thomas factor: K_i, G_i^-1, Y_i, Z_i
  double *X, *S, *rhs;                  // Lagrange multiplier workspace
  double *pH, *dq, *dqH, *tmp1, *tmp2;
  double* red;                          // [B][8] reduction scratch (gHg, gg, ...)
  double* stash;                        // [2][B][T][nb][48] per-body records of the base evaluations (kernels_path.cu)
  size_t stash_half;                    // doubles between the two halves (full batch, also in sub-batch views)
  double* part;                         // [B][T+1][4] per-block-row partial sums of the row-parallel mat-vecs
  int* cnt;                             // [B] arrival counters of the "last CTA of the problem" election
  ProbCtl* ctl;
  double* stats;  // [B][stats_cap][IDTO_NUM_STATS]
  int stats_cap;
  int* status;  // [1] sticky device-side error flag (factorisation failure, active-pair overflow)
  double* crw;              // workspace of the cyclic-reduction solver (kernels_cr.cu), nullptr unless selected
  const double* spline_cp;  // [T+1] modified super-diagonal of the not-a-knot spline's Thomas recurrence (kernels_mpc.cu)
  // Debug trace of the contact pairs each inverse-dynamics evaluation applies forces for (idto_debug_pair_trace;
  // null otherwise): act_base [B][T][np] for tau_t of the state trajectory, act_fd [B][T][nq][4][np] for the
  // perturbed evaluations of tau_{t-1} at q_t +- dq e_i (stencil point kk = 0..3: +dq, -dq, +2dq, -2dq), -1 where
  // an evaluation did not visit the pair (path columns: pairs outside the owner's subtree keep the base forces).
  int *act_base, *act_fd;
};

// ---- kernel launchers (each in its own .cu; all asynchronous on `stream`) ---------------------
void launch_traj(const DevModel& dm, const SolverConsts& sc, const SolverBufs& b, bool scratch, bool force,
                 cudaStream_t stream);
void launch_tau(const DevModel& dm, const SolverConsts& sc, const SolverBufs& b, bool scratch, bool force,
                cudaStream_t stream);
void launch_partials(const DevModel& dm, const SolverConsts& sc, const SolverBufs& b, bool force,
                     cudaStream_t stream);
void launch_assemble(const DevModel& dm, const SolverConsts& sc, const SolverBufs& b, bool force,
                     cudaStream_t stream);
void launch_factor(const SolverConsts& sc, const SolverBufs& b, bool force, cudaStream_t stream);
// KKT sweep (lambda, x = -H~^-1 gm) and, with_gm, the merit gradient / merit / Cauchy scalars that follow it
void launch_lagrange(const DevModel& dm, const SolverConsts& sc, const SolverBufs& b, bool force,
                     cudaStream_t stream, bool with_gm = true);
// kernels_post.cu: scalar tail of the iteration (the model terms of the trust ratio come from k_dogleg_post)
bool trust_final_enabled();
void launch_trust_final(const DevModel& dm, const SolverConsts& sc, const SolverBufs& b, bool commit,
                        cudaStream_t stream);
// register-resident two-sided KKT sweep (kernels_kkt2.cu); false if this block size is not instantiated
bool launch_kkt_tw2(int kb, const SolverConsts& sc, const SolverBufs& b, bool force, cudaStream_t stream);
// third generation (kernels_kkt3.cu): single-warp LU + column-per-thread triangular solves; same contract
bool launch_kkt_v3(int kb, const SolverConsts& sc, const SolverBufs& b, bool force, cudaStream_t stream);
// block cyclic reduction of the same system (kernels_cr.cu): linear_solver = IDTO_LINSOLVE_CYCLIC_REDUCTION
bool launch_kkt_cr(int kb, const SolverConsts& sc, const SolverBufs& b, bool force, cudaStream_t stream);
size_t cr_workspace_doubles(int B, int T, int kb);
void launch_conv_check(const SolverConsts& sc, const SolverBufs& b, cudaStream_t stream);
void launch_clear_dirty(const SolverConsts& sc, const SolverBufs& b, cudaStream_t stream);
void launch_gm_matvec(const SolverConsts& sc, const SolverBufs& b, bool force, cudaStream_t stream);
void launch_dogleg(const SolverConsts& sc, const SolverBufs& b, cudaStream_t stream);
void launch_trust_update(const DevModel& dm, const SolverConsts& sc, const SolverBufs& b, bool commit,
                         cudaStream_t stream);
// MPC shell: spline-shifted guess, shifted nominal trajectory, new initial conditions (kernels_mpc.cu); all
// pointers are device pointers ([B] elapsed seconds, [B][nq] q0, [B][nv] v0, [nq] selector or null)
int launch_mpc_advance(const SolverConsts& sc, const SolverBufs& b, const double* elapsed, const double* q0,
                       const double* v0, const double* selector, double* q_init, double* v_init, double* q_nom,
                       cudaStream_t stream);
int partials_smem_bytes(const DevModel& dm, int nq);
bool chain_supported(const DevModel& dm);
int chain_fit_pair_slots(DevModel dm, int nv);  // default capacity of the active-pair list of a pruned model
int chain_min_smem_bytes(const DevModel& dm, int nv, int method);  // one-slot CTA of the chain-lane ID kernels
void launch_partials_chain(const DevModel& dm, const SolverConsts& sc, const SolverBufs& b, bool force,
                           cudaStream_t stream);
// single-lane subtree evaluations for the path columns (records of the base evaluations: bf.stash)
void launch_partials_path(const DevModel& dm, const SolverConsts& sc, const SolverBufs& b, bool force,
                          cudaStream_t stream);
// tau of the state (scratch = false) or scratch trajectory; with path columns it also writes the per-body records
// of the evaluation into the state's (scratch's) half of bf.stash
void launch_tau_chain(const DevModel& dm, const SolverConsts& sc, const SolverBufs& bf, bool scratch, bool force,
                      cudaStream_t stream);
bool use_chain_kernels(const DevModel& dm);  // chain-lane kernels unless IDTO_DYNAMICS=group or unsupported
extern std::atomic<long> g_launch_counter;  // kernels launched by this library (all solvers; solvers of different devices run on different host threads)

}  // namespace idto
