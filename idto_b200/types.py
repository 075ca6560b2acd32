"""ctypes mirrors of the plain structs in include/idto_b200.h and Python mirrors of the
reference's value types (ProblemDefinition, SolverParameters, Solution, Stats).

Field names/defaults follow optimizer/problem_definition.h:24-59,
optimizer/solver_parameters.h:64-167 and optimizer/trajectory_optimizer_solution.h:16-185.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

_D = ctypes.POINTER(ctypes.c_double)

GRAD_FORWARD, GRAD_CENTRAL, GRAD_CENTRAL4 = 0, 1, 2
SCALING_SQRT, SCALING_ADAPTIVE_SQRT, SCALING_DOUBLE_SQRT, SCALING_ADAPTIVE_DOUBLE_SQRT = 0, 1, 2, 3
LINSOLVE_THOMAS, LINSOLVE_TWISTED, LINSOLVE_DENSE_LDLT, LINSOLVE_CYCLIC_REDUCTION = 0, 1, 2, 3
NUM_STATS = 10
STAT_NAMES = ("cost", "delta", "q_norm", "dq_norm", "dqH_norm", "trust_ratio", "grad_norm", "dL_dq",
              "h_norm", "merit")


class ProblemDesc(ctypes.Structure):
    _fields_ = [("num_steps", ctypes.c_int), ("time_step", ctypes.c_double), ("q_init", _D), ("v_init", _D),
                ("Qq", _D), ("Qv", _D), ("Qf_q", _D), ("Qf_v", _D), ("R", _D), ("q_nom", _D), ("v_nom", _D)]


class Params(ctypes.Structure):
    _fields_ = [("max_iterations", ctypes.c_int), ("gradients_method", ctypes.c_int),
                ("normalize_quaternions", ctypes.c_int), ("contact_stiffness", ctypes.c_double),
                ("dissipation_velocity", ctypes.c_double), ("stiction_velocity", ctypes.c_double),
                ("friction_coefficient", ctypes.c_double), ("smoothing_factor", ctypes.c_double),
                ("scaling", ctypes.c_int), ("scaling_method", ctypes.c_int),
                ("equality_constraints", ctypes.c_int), ("Delta0", ctypes.c_double),
                ("Delta_max", ctypes.c_double), ("check_convergence", ctypes.c_int),
                ("tol_rel_cost_reduction", ctypes.c_double), ("tol_abs_cost_reduction", ctypes.c_double),
                ("tol_rel_gradient_along_dq", ctypes.c_double), ("tol_abs_gradient_along_dq", ctypes.c_double),
                ("tol_rel_state_change", ctypes.c_double), ("tol_abs_state_change", ctypes.c_double),
                ("linear_solver", ctypes.c_int)]


@dataclass
class ConvergenceCriteriaTolerances:
    """optimizer/convergence_criteria_tolerances.h:8-39."""
    rel_cost_reduction: float = 0.0
    abs_cost_reduction: float = 0.0
    rel_gradient_along_dq: float = 0.0
    abs_gradient_along_dq: float = 0.0
    rel_state_change: float = 0.0
    abs_state_change: float = 0.0


@dataclass
class SolverParameters:
    """optimizer/solver_parameters.h:64-167 (hot-path subset; same names and defaults)."""
    max_iterations: int = 100
    gradients_method: int = GRAD_FORWARD
    normalize_quaternions: bool = False
    verbose: bool = True
    contact_stiffness: float = 100.0
    dissipation_velocity: float = 0.1
    stiction_velocity: float = 0.05
    friction_coefficient: float = 0.5
    smoothing_factor: float = 0.1
    scaling: bool = True
    scaling_method: int = SCALING_DOUBLE_SQRT
    equality_constraints: bool = True
    Delta0: float = 1e-1
    Delta_max: float = 1e5
    num_threads: int = 1
    check_convergence: bool = False
    convergence_tolerances: ConvergenceCriteriaTolerances = field(default_factory=ConvergenceCriteriaTolerances)
    linear_solver: int = LINSOLVE_TWISTED

    def to_c(self) -> Params:
        t = self.convergence_tolerances
        return Params(int(self.max_iterations), int(self.gradients_method), int(self.normalize_quaternions),
                      float(self.contact_stiffness), float(self.dissipation_velocity),
                      float(self.stiction_velocity), float(self.friction_coefficient),
                      float(self.smoothing_factor), int(self.scaling), int(self.scaling_method),
                      int(self.equality_constraints), float(self.Delta0), float(self.Delta_max),
                      int(self.check_convergence), t.rel_cost_reduction, t.abs_cost_reduction,
                      t.rel_gradient_along_dq, t.abs_gradient_along_dq, t.rel_state_change,
                      t.abs_state_change, int(self.linear_solver))


@dataclass
class ProblemDefinition:
    """optimizer/problem_definition.h:24-59.  `time_step` is the plant's (plant.time_step())."""
    num_steps: int = 0
    q_init: np.ndarray = None
    v_init: np.ndarray = None
    Qq: np.ndarray = None
    Qv: np.ndarray = None
    Qf_q: np.ndarray = None
    Qf_v: np.ndarray = None
    R: np.ndarray = None
    q_nom: list = None
    v_nom: list = None

    def to_c(self, time_step: float, nq: int, nv: int):
        """Returns (ProblemDesc, keepalive).  Size checks replace cc:74-82's DRAKE_DEMANDs."""
        T = int(self.num_steps)
        keep = {}

        def arr(name, x, shape):
            a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
            if a.shape != shape:
                raise ValueError(f"ProblemDefinition.{name}: expected shape {shape}, got {a.shape}")
            keep[name] = a
            return a.ctypes.data_as(_D)

        def mat(name, x, n):
            a = np.asarray(x, dtype=np.float64)
            if a.shape != (n, n):
                raise ValueError(f"ProblemDefinition.{name}: expected shape {(n, n)}, got {a.shape}")
            a = np.asfortranarray(a)  # column-major like Eigen
            keep[name] = a
            return a.ctypes.data_as(_D)

        d = ProblemDesc(T, float(time_step), arr("q_init", self.q_init, (nq,)), arr("v_init", self.v_init, (nv,)),
                        mat("Qq", self.Qq, nq), mat("Qv", self.Qv, nv), mat("Qf_q", self.Qf_q, nq),
                        mat("Qf_v", self.Qf_v, nv), mat("R", self.R, nv),
                        arr("q_nom", np.asarray(self.q_nom, dtype=np.float64), (T + 1, nq)),
                        arr("v_nom", np.asarray(self.v_nom, dtype=np.float64), (T + 1, nv)))
        return d, keep


@dataclass
class TrajectoryOptimizerSolution:
    """optimizer/trajectory_optimizer_solution.h:46-56."""
    q: list = field(default_factory=list)
    v: list = field(default_factory=list)
    tau: list = field(default_factory=list)


@dataclass
class TrajectoryOptimizerStats:
    """optimizer/trajectory_optimizer_solution.h:58-185 (13 per-iteration series + solve_time)."""
    solve_time: float = 0.0
    iteration_times: list = field(default_factory=list)
    iteration_costs: list = field(default_factory=list)
    linesearch_iterations: list = field(default_factory=list)
    linesearch_alphas: list = field(default_factory=list)
    trust_region_radii: list = field(default_factory=list)
    gradient_norms: list = field(default_factory=list)
    q_norms: list = field(default_factory=list)
    dq_norms: list = field(default_factory=list)
    dqH_norms: list = field(default_factory=list)
    trust_ratios: list = field(default_factory=list)
    dL_dqs: list = field(default_factory=list)
    h_norms: list = field(default_factory=list)
    merits: list = field(default_factory=list)

    def push_data(self, iter_time, iter_cost, linesearch_iters, alpha, delta, q_norm, dq_norm, dqH_norm,
                  trust_ratio, grad_norm, dL_dq, h_norm, merit):
        self.iteration_times.append(iter_time)
        self.iteration_costs.append(iter_cost)
        self.linesearch_iterations.append(linesearch_iters)
        self.linesearch_alphas.append(alpha)
        self.trust_region_radii.append(delta)
        self.q_norms.append(q_norm)
        self.dq_norms.append(dq_norm)
        self.dqH_norms.append(dqH_norm)
        self.trust_ratios.append(trust_ratio)
        self.gradient_norms.append(grad_norm)
        self.dL_dqs.append(dL_dq)
        self.h_norms.append(h_norm)
        self.merits.append(merit)

    def push_row(self, iter_time, row):
        """row = one stats[b][iter][:] record of idto_solve (IDTO_NUM_STATS entries)."""
        cost, delta, q_norm, dq_norm, dqH_norm, rho, gnorm, dL_dq, hnorm, merit = [float(x) for x in row]
        self.push_data(iter_time, cost, 0, float("nan"), delta, q_norm, dq_norm, dqH_norm, rho, gnorm, dL_dq,
                       hnorm, merit)

    def is_empty(self):
        return all(len(x) == 0 for x in (
            self.iteration_times, self.iteration_costs, self.linesearch_iterations, self.linesearch_alphas,
            self.trust_region_radii, self.q_norms, self.dq_norms, self.dqH_norms, self.trust_ratios,
            self.gradient_norms, self.dL_dqs, self.h_norms, self.merits))

    def SaveToCsv(self, fname):
        with open(fname, "w") as f:
            f.write("iter, time, cost, ls_iters, alpha, delta, q_norm, dq_norm, dqH_norm, "
                    "trust_ratio, grad_norm, dL_dq, h_norm, merit\n")
            for i in range(len(self.iteration_times)):
                f.write(", ".join(str(x) for x in (
                    i, self.iteration_times[i], self.iteration_costs[i], self.linesearch_iterations[i],
                    self.linesearch_alphas[i], self.trust_region_radii[i], self.q_norms[i], self.dq_norms[i],
                    self.dqH_norms[i], self.trust_ratios[i], self.gradient_norms[i], self.dL_dqs[i],
                    self.h_norms[i], self.merits[i])) + "\n")
