"""Problem/solver definitions of the BASELINE.json configs (SURVEY.md §8d "Synthetic inputs").

Values restate the reference's example configs with `num_steps` overridden to the BASELINE
horizon:
  acrobot      examples/acrobot/acrobot.yaml:7-30
  spinner      python_bindings/test/trajectory_optimizer_test.py:29-62 (== examples/spinner/spinner.yaml)
  hopper       examples/hopper/hopper.yaml:7-31,68-80
  mini_cheetah python_examples/mini_cheetah_mpc.py:33-97 (the twin that loads
               models/mini_cheetah_with_ground.urdf, the "4 feet" benchmark model)
Nominal trajectories follow examples/example_base.cc:394-424 (linear interpolation of q_nom;
v_nom from finite differences when nq == nv, else v_init).
"""
from __future__ import annotations

import numpy as np

from .bake import BakedModel, load_model
from .types import (GRAD_CENTRAL, GRAD_FORWARD, ProblemDefinition, SolverParameters)


def _interp(a, b, n):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return [a + (b - a) * (i / (n - 1)) for i in range(n)]


def _make(model: BakedModel, T, dt, q_init, v_init, q_nom_start, q_nom_end, Qq, Qv, R, Qfq, Qfv, relative=None):
    q_init = np.asarray(q_init, float)
    v_init = np.asarray(v_init, float)
    rel = np.zeros(model.nq) if relative is None else np.asarray(relative, float)
    qs = np.asarray(q_nom_start, float) + rel * q_init
    qe = np.asarray(q_nom_end, float) + rel * q_init
    q_nom = _interp(qs, qe, T + 1)
    v_nom = [v_init.copy()]
    for t in range(1, T + 1):
        if model.nq == model.nv:
            v_nom.append((q_nom[t] - q_nom[t - 1]) / dt)
        else:
            v_nom.append(v_init.copy())
    for qs_ in model.quat_q_starts:  # example_base.cc:424 NormalizeQuaternions(q_nom)
        for q in q_nom:
            q[qs_:qs_ + 4] /= np.linalg.norm(q[qs_:qs_ + 4])
    return ProblemDefinition(num_steps=T, q_init=q_init, v_init=v_init, Qq=np.diag(Qq).astype(float),
                             Qv=np.diag(Qv).astype(float), Qf_q=np.diag(Qfq).astype(float),
                             Qf_v=np.diag(Qfv).astype(float), R=np.diag(R).astype(float),
                             q_nom=q_nom, v_nom=v_nom)


def acrobot(T=40, gradients_method=GRAD_FORWARD):
    m = load_model("acrobot")
    prob = _make(m, T, 0.05, [0, 0], [0, 0], [3.1415, 0], [3.1415, 0], [1, 1], [1, 1], [1e3, 0.1],
                 [100, 100], [1, 1])
    params = SolverParameters(max_iterations=100, scaling=False, equality_constraints=True, Delta0=1e3,
                              gradients_method=gradients_method, verbose=False)
    guess = _interp([0, 0], [0, 0], T + 1)
    return m, 0.05, prob, params, guess


def spinner(T=40, gradients_method=GRAD_FORWARD, max_iterations=200):
    m = load_model("spinner")
    prob = _make(m, T, 0.05, [0.3, 1.5, 0.0], [0, 0, 0], [0.3, 1.5, 2.0], [0.3, 1.5, 2.0], [1, 1, 1],
                 [0.1, 0.1, 0.1], [0.1, 0.1, 1e3], [10, 10, 10], [0.1, 0.1, 0.1])
    # python_bindings/test/trajectory_optimizer_test.py:38-45 uses a constant nominal with v_nom = 0
    prob.v_nom = [np.zeros(3) for _ in range(T + 1)]
    params = SolverParameters(max_iterations=max_iterations, scaling=True, equality_constraints=True,
                              Delta0=1e1, Delta_max=1e5, contact_stiffness=200, dissipation_velocity=0.1,
                              smoothing_factor=0.01, friction_coefficient=0.5, stiction_velocity=0.05,
                              gradients_method=gradients_method, verbose=False)
    guess = [np.array([0.3, 1.5, 0.0]) for _ in range(T + 1)]
    return m, 0.05, prob, params, guess


def spinner_capsule(T=40, gradients_method=GRAD_FORWARD, max_iterations=200):
    """models/spinner_capsule.urdf (finger of spheres against a capsule-shaped spinner) with the spinner example's
    cost and contact parameters (examples/spinner/spinner.yaml); exercises the sphere-capsule closed form."""
    m = load_model("spinner_capsule")
    prob = _make(m, T, 0.05, [0.2, 1.5, 0.0], [0, 0, 0], [0.2, 1.5, 0.0], [0.2, 1.5, 2.0], [1, 1, 1],
                 [0.1, 0.1, 0.1], [0.1, 0.1, 1e3], [10, 10, 10], [0.1, 0.1, 0.1])
    params = SolverParameters(max_iterations=max_iterations, scaling=True, equality_constraints=True,
                              Delta0=1e1, Delta_max=1e5, contact_stiffness=200, dissipation_velocity=0.1,
                              smoothing_factor=0.01, friction_coefficient=0.5, stiction_velocity=0.05,
                              gradients_method=gradients_method, verbose=False)
    guess = [np.array([0.2, 1.5, 0.0]) for _ in range(T + 1)]
    return m, 0.05, prob, params, guess


def hopper(T=50, gradients_method=GRAD_FORWARD, max_iterations=200, model="hopper"):
    m = load_model(model)  # "hopper_half_space": the same ground registered as a HalfSpace
    q0 = [0.61, 0.0, 0.3, -0.5, 0.2]
    qe = [0.61, -0.5, 0.3, -0.5, 0.2]
    prob = _make(m, T, 0.05, q0, [0] * 5, q0, qe, [1.0] * 5, [0.1] * 5, [1e2, 1e2, 1e2, 0.1, 0.1],
                 [10] * 5, [1.0] * 5)
    params = SolverParameters(max_iterations=max_iterations, scaling=True, equality_constraints=True,
                              Delta0=1e-3, contact_stiffness=800, dissipation_velocity=0.1,
                              smoothing_factor=0.01, friction_coefficient=1.0, stiction_velocity=0.05,
                              gradients_method=gradients_method, verbose=False)
    guess = _interp(q0, qe, T + 1)
    return m, 0.05, prob, params, guess


_CHEETAH_Q0 = [1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.29] + [0.0, -0.8, 1.6] * 4


def mini_cheetah(T=40, gradients_method=GRAD_CENTRAL, max_iterations=1):
    m = load_model("mini_cheetah")
    qe = list(_CHEETAH_Q0)
    qe[4] = 0.4
    prob = _make(m, T, 0.05, _CHEETAH_Q0, [0.0] * 18, _CHEETAH_Q0, qe,
                 [10] * 4 + [10] * 3 + [0] * 12, [1] * 3 + [1] * 3 + [0.1] * 12,
                 [100] * 3 + [100] * 3 + [0.01] * 12, [10] * 4 + [10] * 3 + [1] * 12,
                 [1] * 3 + [1] * 3 + [0.1] * 12)
    params = SolverParameters(max_iterations=max_iterations, scaling=True, equality_constraints=True,
                              contact_stiffness=2000, dissipation_velocity=0.1, smoothing_factor=0.01,
                              friction_coefficient=1.0, stiction_velocity=0.5,
                              gradients_method=gradients_method, verbose=False)
    guess = [np.array(_CHEETAH_Q0) for _ in range(T + 1)]
    return m, 0.05, prob, params, guess


_ALLEGRO_Q0 = [-0.1, 1.0, 1.0, 1.0, 0.6, 1.9, 1.0, 1.0, 0.0, 0.7, 1.0, 1.0, 0.1, 1.0, 1.0, 1.0,
               1.0, 0.0, 0.0, 0.0, -0.06, 0.0, 0.07]


def allegro_hand(T=60, gradients_method=GRAD_FORWARD, max_iterations=1):
    """examples/allegro_hand/allegro_hand.yaml:8-125 (in-hand rotation of a free ball: 16 finger joints + the
    ball's quaternion and position; 188 candidate contact pairs), num_steps overridden to the BASELINE horizon."""
    m = load_model("allegro_hand")
    qn = list(_ALLEGRO_Q0)
    qn[16:20] = [0.7, 0.0, 0.0, -0.7]
    prob = _make(m, T, 0.05, _ALLEGRO_Q0, [0.0] * 22, qn, qn,
                 [1e-2] * 16 + [1e1] * 7, [1e-3] * 16 + [1e0] * 6, [1e-1] * 16 + [1e3] * 6,
                 [1e0] * 16 + [1e2] * 7, [1e1] * 22)
    params = SolverParameters(max_iterations=max_iterations, scaling=True, equality_constraints=True,
                              contact_stiffness=100, dissipation_velocity=0.1, smoothing_factor=0.001,
                              friction_coefficient=1.0, stiction_velocity=0.1,
                              gradients_method=gradients_method, verbose=False)
    guess = [np.array(_ALLEGRO_Q0) for _ in range(T + 1)]
    return m, 0.05, prob, params, guess


def allegro_hand_upside_down(T=40, gradients_method=GRAD_FORWARD, max_iterations=1):
    """examples/allegro_hand/allegro_hand_upside_down.yaml:1-106 (`allegro_hand --upside_down`, allegro_hand.cc:94-97:
    gravity reversed, the ball hangs in the finger tips and is to be turned a quarter turn)."""
    m = load_model("allegro_hand_upside_down")
    q0 = [-0.2, 1.4, 0.6, 0.7, 0.3, 1.5, 1.0, 1.0, 0.0, 0.7, 1.0, 1.0, 0.1, 1.0, 1.0, 1.0,
          1.0, 0.0, 0.0, 0.0, -0.06, 0.0, 0.07]
    qe = list(q0)
    qe[16:20] = [0.7, 0.0, 0.0, -0.7]
    prob = _make(m, T, 0.05, q0, [0.0] * 22, q0, qe,
                 [1e-2] * 16 + [1e1] * 7, [1e-3] * 16 + [1e0] * 6, [1e-1] * 16 + [1e3] * 6,
                 [1e-2] * 16 + [1e1] * 7, [1e-3] * 16 + [1e0] * 6)
    params = SolverParameters(max_iterations=max_iterations, scaling=True, equality_constraints=True,
                              contact_stiffness=100, dissipation_velocity=0.01, smoothing_factor=0.001,
                              friction_coefficient=1.0, stiction_velocity=0.03,
                              gradients_method=gradients_method, verbose=False)
    guess = [np.array(q0) for _ in range(T + 1)]
    return m, 0.05, prob, params, guess


def pendulum(T=20, dt=0.05, gradients_method=GRAD_FORWARD):
    """optimizer/test/trajectory_optimizer_test.cc:434-490 (PendulumSwingup)."""
    m = load_model("pendulum")
    prob = ProblemDefinition(num_steps=T, q_init=np.array([0.1]), v_init=np.array([0.0]),
                             Qq=0.0 * np.eye(1), Qv=0.1 * np.eye(1), Qf_q=1000 * np.eye(1),
                             Qf_v=1 * np.eye(1), R=0.01 * np.eye(1),
                             q_nom=[np.array([np.pi]) for _ in range(T + 1)],
                             v_nom=[np.array([0.0]) for _ in range(T + 1)])
    params = SolverParameters(max_iterations=20, gradients_method=gradients_method, verbose=False,
                              check_convergence=True)
    params.convergence_tolerances.rel_cost_reduction = 1e-5
    params.convergence_tolerances.abs_cost_reduction = 1e-5
    guess = [np.array([0.1]) for _ in range(T + 1)]
    return m, dt, prob, params, guess


def perturbed_batch(model: BakedModel, prob: ProblemDefinition, batch: int):
    """Batch element b: rng(b) perturbation of the initial condition:
    joint angles +N(0,0.02^2), base xyz +N(0,0.01^2), base quaternion rotated by a rotation vector
    ~N(0,0.02^2) and renormalised, v_init ~N(0,0.1^2); q_guess_t = q_init for all t."""
    T, nq, nv = prob.num_steps, model.nq, model.nv
    q_init = np.zeros((batch, nq))
    v_init = np.zeros((batch, nv))
    for b in range(batch):
        rng = np.random.default_rng(b)
        q = np.array(prob.q_init, float)
        dq = rng.normal(0.0, 0.02, nq)
        for k in range(model.nbodies):
            qs = int(model.q_start[k])
            if model.joint_type[k] == 3:  # quaternion floating
                rv = rng.normal(0.0, 0.02, 3)
                ang = np.linalg.norm(rv)
                ax = rv / ang if ang > 0 else np.array([1.0, 0, 0])
                dw, dv_ = np.cos(ang / 2), np.sin(ang / 2) * ax
                w, vv = q[qs], q[qs + 1:qs + 4]
                qn = np.concatenate([[dw * w - dv_ @ vv], dw * vv + w * dv_ + np.cross(dv_, vv)])
                q[qs:qs + 4] = qn / np.linalg.norm(qn)
                q[qs + 4:qs + 7] += rng.normal(0.0, 0.01, 3)
            else:
                n = 3 if model.joint_type[k] == 2 else 1
                q[qs:qs + n] += dq[qs:qs + n]
        q_init[b] = q
        v_init[b] = rng.normal(0.0, 0.1, nv)
    q_guess = np.repeat(q_init[:, None, :], T + 1, axis=1)
    return q_init, v_init, q_guess
