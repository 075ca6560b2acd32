"""YAML front-end of the reference's examples.

`TrajOptExampleParams` mirrors examples/yaml_config.h:24-218 field for field (same names, same defaults),
`SetProblemDefinition` / `SetSolverParameters` mirror TrajOptExample::SetProblemDefinition and
::SetSolverParameters (examples/example_base.cc:377-543), including their error behaviour (unknown
option strings raise RuntimeError with the reference's message).  With these the reference's example
configs (examples/*/*.yaml) drive the CUDA path unmodified:

    opts = load_yaml("examples/hopper/hopper.yaml")
    prob = SetProblemDefinition(opts, baked_model)
    params = SetSolverParameters(opts)
    solver = capi.BatchSolver(capi.Model(baked_model), opts.time_step, prob, params, batch)
    solver.set_q(MakeInitialGuess(opts))

Options that select code outside the hot path (method: linesearch, gradients_method: autodiff,
exact_hessian) parse like in the reference and are recorded on the returned
parameters; the CUDA solver rejects them at creation (IDTO_ERR_UNSUPPORTED), it never falls back.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields

import numpy as np

from .types import (LINSOLVE_DENSE_LDLT, LINSOLVE_TWISTED, GRAD_CENTRAL, GRAD_CENTRAL4, GRAD_FORWARD, SCALING_ADAPTIVE_DOUBLE_SQRT, SCALING_ADAPTIVE_SQRT,
                    SCALING_DOUBLE_SQRT, SCALING_SQRT, ConvergenceCriteriaTolerances, ProblemDefinition,
                    SolverParameters)


def _vec():
    return field(default_factory=lambda: np.zeros(0))


@dataclass
class TrajOptExampleParams:
    """examples/yaml_config.h:24-218."""
    q_init: np.ndarray = _vec()
    v_init: np.ndarray = _vec()
    q_nom_start: np.ndarray = _vec()
    q_nom_end: np.ndarray = _vec()
    q_guess: np.ndarray = _vec()
    Qq: np.ndarray = _vec()
    Qv: np.ndarray = _vec()
    R: np.ndarray = _vec()
    Qfq: np.ndarray = _vec()
    Qfv: np.ndarray = _vec()
    time_step: float = 0.0
    num_steps: int = 0
    max_iters: int = 0
    tolerances: ConvergenceCriteriaTolerances = field(default_factory=ConvergenceCriteriaTolerances)
    linesearch: str = "armijo"
    method: str = "trust_region"
    gradients_method: str = "forward_differences"
    linear_solver: str = "pentadiagonal_lu"
    play_optimal_trajectory: bool = True
    play_initial_guess: bool = False
    play_target_trajectory: bool = False
    linesearch_plot_every_iteration: bool = False
    print_debug_data: bool = False
    save_solver_stats_csv: bool = True
    contact_stiffness: float = 100.0
    dissipation_velocity: float = 0.1
    smoothing_factor: float = 1.0
    stiction_velocity: float = 0.05
    friction_coefficient: float = 0.5
    save_contour_data: bool = False
    contour_q1_min: float = 0.0
    contour_q1_max: float = 1.0
    contour_q2_min: float = 0.0
    contour_q2_max: float = 1.0
    save_lineplot_data: bool = False
    lineplot_q_min: float = 0.0
    lineplot_q_max: float = 1.0
    verbose: bool = True
    normalize_quaternions: bool = False
    exact_hessian: bool = False
    scaling: bool = True
    mpc: bool = False
    mpc_iters: int = 1
    controller_frequency: float = 30.0
    sim_time: float = 10.0
    sim_time_step: float = 1e-3
    sim_realtime_rate: float = 1.0
    Kp: np.ndarray = _vec()
    Kd: np.ndarray = _vec()
    feed_forward: bool = True
    scaling_method: str = "double_sqrt"
    equality_constraints: bool = True
    Delta_max: float = 1e5
    Delta0: float = 1e-1
    num_threads: int = 1
    q_nom_relative_to_q_init: np.ndarray = field(default_factory=lambda: np.zeros(0, bool))
    save_mpc_result_as_static_html: bool = False
    static_html_filename: str = "/tmp/meshcat_recording.html"


_VECTORS = {"q_init", "v_init", "q_nom_start", "q_nom_end", "q_guess", "Qq", "Qv", "R", "Qfq", "Qfv", "Kp", "Kd"}


def from_dict(d: dict) -> TrajOptExampleParams:
    """Like drake::yaml::LoadYamlFile with its defaults: unknown keys are an error, missing keys keep the
    struct's default."""
    known = {f.name: f for f in fields(TrajOptExampleParams)}
    out = TrajOptExampleParams()
    for k, v in d.items():
        if k not in known:
            raise RuntimeError(f"YAML node has an unknown key '{k}' (not a field of TrajOptExampleParams)")
        if k in _VECTORS:
            v = np.asarray(v, float).reshape(-1)
        elif k == "q_nom_relative_to_q_init":
            v = np.asarray(v, bool).reshape(-1)
        elif k == "tolerances":
            t = ConvergenceCriteriaTolerances()
            for tk, tv in (v or {}).items():
                if not hasattr(t, tk):
                    raise RuntimeError(f"unknown convergence tolerance '{tk}'")
                setattr(t, tk, float(tv))
            v = t
        else:
            typ = type(getattr(out, k))
            v = typ(v) if typ in (int, float, str) else (bool(v) if typ is bool else v)
        setattr(out, k, v)
    return out


def load_yaml(path: str) -> TrajOptExampleParams:
    import yaml
    with open(path) as f:
        return from_dict(yaml.safe_load(f) or {})


def load_yaml_string(text: str) -> TrajOptExampleParams:
    import yaml
    return from_dict(yaml.safe_load(text) or {})


def MakeLinearInterpolation(start, end, N):
    """examples/example_base.cc:309-316: N points from start to end inclusive."""
    start, end = np.asarray(start, float), np.asarray(end, float)
    return [start + (end - start) * (i / (N - 1.0)) for i in range(N)]


def MakeInitialGuess(options: TrajOptExampleParams):
    """examples/example_base.cc:110-111, 140-141: linear interpolation from q_init to q_guess."""
    return MakeLinearInterpolation(options.q_init, options.q_guess, options.num_steps + 1)


def _relative(options):
    rel = np.asarray(options.q_nom_relative_to_q_init, bool)
    if rel.size == 0:  # not specified: the nominal trajectory is not relative to the initial condition
        rel = np.zeros(np.asarray(options.q_init).size, bool)
    return rel


def SetProblemDefinition(options: TrajOptExampleParams, model) -> ProblemDefinition:
    """examples/example_base.cc:377-425.  `model` is the baked model (it stands in for the plant: only the
    quaternion locations are read, for NormalizeQuaternions)."""
    q_init = np.asarray(options.q_init, float).copy()
    v_init = np.asarray(options.v_init, float).copy()
    rel = _relative(options).astype(float)
    q_nom = MakeLinearInterpolation(options.q_nom_start + rel * q_init, options.q_nom_end + rel * q_init,
                                    options.num_steps + 1)
    v_nom = [v_init.copy()]
    for t in range(1, options.num_steps + 1):
        if q_init.size == v_init.size:  # no quaternion DoFs: v_nom from q_nom
            v_nom.append((q_nom[t] - q_nom[t - 1]) / options.time_step)
        else:
            v_nom.append(v_init.copy())
    for q0 in model.quat_q_starts:  # NormalizeQuaternions(plant, &q_nom / &q_init)
        for q in q_nom:
            q[q0:q0 + 4] /= np.linalg.norm(q[q0:q0 + 4])
        q_init[q0:q0 + 4] /= np.linalg.norm(q_init[q0:q0 + 4])
    return ProblemDefinition(num_steps=int(options.num_steps), q_init=q_init, v_init=v_init,
                             Qq=np.diag(options.Qq), Qv=np.diag(options.Qv), Qf_q=np.diag(options.Qfq),
                             Qf_v=np.diag(options.Qfv), R=np.diag(options.R), q_nom=q_nom, v_nom=v_nom)


def SetSolverParameters(options: TrajOptExampleParams) -> SolverParameters:
    """examples/example_base.cc:427-543."""
    if options.linesearch not in ("backtracking", "armijo"):
        raise RuntimeError(f"Unknown linesearch method '{options.linesearch}'")
    gm = {"forward_differences": GRAD_FORWARD, "central_differences": GRAD_CENTRAL,
          "central_differences4": GRAD_CENTRAL4, "autodiff": None}
    if options.gradients_method not in gm:
        raise RuntimeError(f"Unknown gradient method '{options.gradients_method}'")
    if options.method not in ("linesearch", "trust_region"):
        raise RuntimeError(f"Unknown solver method '{options.method}'")
    if options.linear_solver not in ("pentadiagonal_lu", "dense_ldlt"):
        raise RuntimeError(f"Unknown linear solver '{options.linear_solver}'")
    sm = {"sqrt": SCALING_SQRT, "adaptive_sqrt": SCALING_ADAPTIVE_SQRT, "double_sqrt": SCALING_DOUBLE_SQRT,
          "adaptive_double_sqrt": SCALING_ADAPTIVE_DOUBLE_SQRT}
    if options.scaling_method not in sm:
        raise RuntimeError(f"Unknown scaling method '{options.scaling_method}'")
    t = options.tolerances
    p = SolverParameters(
        max_iterations=int(options.max_iters),
        gradients_method=gm[options.gradients_method] if gm[options.gradients_method] is not None else GRAD_FORWARD,
        normalize_quaternions=bool(options.normalize_quaternions), verbose=bool(options.verbose),
        contact_stiffness=float(options.contact_stiffness), dissipation_velocity=float(options.dissipation_velocity),
        stiction_velocity=float(options.stiction_velocity), friction_coefficient=float(options.friction_coefficient),
        smoothing_factor=float(options.smoothing_factor), scaling=bool(options.scaling),
        scaling_method=sm[options.scaling_method], equality_constraints=bool(options.equality_constraints),
        Delta0=float(options.Delta0), Delta_max=float(options.Delta_max), num_threads=int(options.num_threads),
        linear_solver=LINSOLVE_DENSE_LDLT if options.linear_solver == "dense_ldlt" else LINSOLVE_TWISTED,
        convergence_tolerances=t)
    # check_convergence keeps its default (False): the reference's SetSolverParameters copies the tolerances
    # (example_base.cc:427-543) and never enables the check; only its unit tests do
    # recorded for the caller; anything outside the CUDA hot path is rejected at solver creation
    p.unsupported = [name for name, bad in (("method: linesearch", options.method == "linesearch"),
                                            ("gradients_method: autodiff", options.gradients_method == "autodiff"),
                                            ("exact_hessian", bool(options.exact_hessian))) if bad]
    p.q_nom_relative_to_q_init = _relative(options)
    return p
