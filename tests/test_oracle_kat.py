"""Pins the CPU oracle against the reference's own Drake-free known-answer tests
(SURVEY.md §4 / §8c).  Citations: optimizer/test/trajectory_optimizer_test.cc ("tt"),
optimizer/test/penta_diagonal_solver_test.cc ("pt"), python_bindings/test/*.py.
"""
import copy

import numpy as np
import pytest

from idto_b200 import problems
from idto_b200.bake import load_model
from idto_b200.types import (GRAD_CENTRAL, GRAD_CENTRAL4, GRAD_FORWARD, ProblemDefinition, SolverParameters)

EPS = np.finfo(float).eps


def rel_close(a, b, tol):
    """utils/eigen_matrix_compare.h:92-98 relative rule: |a-b| <= tol*max(1,|a|,|b|) element-wise."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))


def pend_problem(T, q_init=0.0, v_init=0.0, **kw):
    d = dict(num_steps=T, q_init=np.array([q_init]), v_init=np.array([v_init]), Qq=np.zeros((1, 1)),
             Qv=np.zeros((1, 1)), Qf_q=np.zeros((1, 1)), Qf_v=np.zeros((1, 1)), R=np.zeros((1, 1)),
             q_nom=[np.zeros(1)] * (T + 1), v_nom=[np.zeros(1)] * (T + 1))
    d.update(kw)
    return ProblemDefinition(**d)


def test_spinner_end_to_end_golden(oracle_mod):
    """python_bindings/test/trajectory_optimizer_test.py:84-85: the one numeric anchor that crosses
    the Drake boundary with contact (3 revolute joints, one sphere-sphere pair)."""
    m, dt, prob, params, guess = problems.spinner()
    o = oracle_mod.Oracle(m, dt, prob, params)
    o.set_q(guess)
    k, _, stats = o.solve(200)
    assert k == 200
    q, v, tau = o.solution()
    assert np.linalg.norm(q[-1] - np.array([0.287, 1.497, 1.995])) < 1e-3
    assert o.unactuated_dofs() == [2]  # tt:1501


def test_warm_start_equals_one_shot(oracle_mod):
    """python_bindings/test/warm_start_test.py:165-182: 10x SolveFromWarmStart(max_iterations=1) ==
    Solve(max_iterations=10) to 1e-8 on q, v, costs, Delta, |g|."""
    m, dt, prob, params, guess = problems.spinner(max_iterations=10)
    a = oracle_mod.Oracle(m, dt, prob, params)
    a.set_q(guess)
    _, _, sa = a.solve(10)
    b = oracle_mod.Oracle(m, dt, prob, params)
    b.set_q(guess)
    sb = np.concatenate([b.solve(1)[2] for _ in range(10)])
    qa, va, _ = a.solution()
    qb, vb, _ = b.solution()
    assert np.max(np.abs(qa - qb)) < 1e-8 and np.max(np.abs(va - vb)) < 1e-8
    for col in (0, 1, 6):  # cost, Delta, |g|
        assert np.max(np.abs(sa[:, col] - sb[:, col])) < 1e-8


def test_pendulum_inverse_dynamics_analytic(oracle_mod):
    """tt:1314-1385 PendulumCalcInverseDynamics: tau = m l^2 a + b v + m g l sin(q), eps relative."""
    T, dt = 5, 1e-2
    m = load_model("pendulum")
    o = oracle_mod.Oracle(m, dt, pend_problem(T, 0.0, 0.1), SolverParameters())
    q = np.array([[0.0 + 0.6 * t] for t in range(T + 1)])
    q[0] = 0.0
    o.set_q(q)
    o.eval(0)
    v = o.get("v")
    tau = o.get("tau")
    ml2, b, mgl = 0.25, 0.1, 9.81 * 0.5
    for t in range(T):
        a_t = (v[t + 1] - v[t]) / dt
        gt = ml2 * a_t + b * v[t + 1] + mgl * np.sin(q[t + 1, 0])
        assert rel_close(tau[t], gt, 10 * EPS)


@pytest.mark.parametrize("method,tol", [(GRAD_FORWARD, 100), (GRAD_CENTRAL, 1), (GRAD_CENTRAL4, 1)])
def test_pendulum_dtau_dq_analytic(oracle_mod, method, tol):
    """tt:1058-1150 PendulumDtauDq closed-form partials incl. m g l cos q, sqrt(eps) relative
    (the reference runs this with the default forward differences)."""
    T, dt = 5, 1e-2
    m = load_model("pendulum")
    o = oracle_mod.Oracle(m, dt, pend_problem(T, 0.0, 0.1), SolverParameters(gradients_method=method))
    q = np.array([[0.6 * t] for t in range(T + 1)])
    o.set_q(q)
    o.eval(1)
    dqm, dqt, dqp = o.get("dtau_dqm"), o.get("dtau_dqt"), o.get("dtau_dqp")
    ml2, b, mgl = 0.25, 0.1, 9.81 * 0.5
    k = tol * np.sqrt(EPS)
    for t in range(1, T):
        assert rel_close(dqp[t], ml2 / dt / dt + b / dt + mgl * np.cos(q[t + 1, 0]), k)
        assert rel_close(dqt[t], -2 * ml2 / dt / dt - b / dt, k)
        assert rel_close(dqm[t], 0.0 if t == 1 else ml2 / dt / dt, k)
    assert np.isnan(dqm[0]) and dqt[0] == 0.0  # inverse_dynamics_partials.h:38-42


def test_cost_from_state_golden_vector(oracle_mod):
    """tt:1155-1245 CalcCostFromState: 11 hard-coded q values -> cost vs hand formula, 100 eps."""
    T, dt = 10, 5e-2
    m = load_model("pendulum")
    m.gravity = np.zeros(3)
    prob = pend_problem(T, Qv=0.1 * np.eye(1), Qf_q=10 * np.eye(1), Qf_v=np.eye(1), R=np.eye(1),
                        q_nom=[np.array([np.pi])] * (T + 1), v_nom=[np.array([-0.1])] * (T + 1))
    q = np.array([0.0000000000000000000000000, 0.0950285641187840757204697, 0.2659896360172592788551071,
                  0.4941147113506765831125733, 0.7608818755930255584019051, 1.0479359055822168311777887,
                  1.3370090901260500704239575, 1.6098424281109515732168802, 1.8481068641834854648919872,
                  2.0333242222438583368671061, 2.1467874956452459578315484])
    o = oracle_mod.Oracle(m, dt, prob, SolverParameters())
    o.set_q(q.reshape(-1, 1))
    o.eval(0)
    L = o.get("cost")[0]
    ml2, b = 0.25, 0.1
    L_gt, vt = 0.0, 0.0
    for t in range(T):
        if t > 0:
            vt = (q[t] - q[t - 1]) / dt
        vp = (q[t + 1] - q[t]) / dt
        ut = ml2 * (vp - vt) / dt + b * vp
        L_gt += dt * (vt + 0.1) * 0.1 * (vt + 0.1) + dt * ut * ut
    vt = (q[T] - q[T - 1]) / dt
    L_gt += (q[T] - np.pi) * 10 * (q[T] - np.pi) + (vt + 0.1) ** 2
    assert abs(L - L_gt) < 100 * EPS * max(1.0, abs(L_gt))


def test_gradient_matches_finite_difference_of_cost(oracle_mod):
    """tt:1000-1056 CalcGradientPendulum: g vs central FD of the cost (dt=1e-3 there; sqrt(eps)/dt)."""
    T, dt = 5, 1e-2
    m = load_model("pendulum")
    prob = pend_problem(T, 0.1, 0.0, Qq=0.1 * np.eye(1), Qv=0.2 * np.eye(1), Qf_q=0.3 * np.eye(1),
                        Qf_v=0.4 * np.eye(1), R=0.5 * np.eye(1), q_nom=[np.array([np.pi])] * (T + 1),
                        v_nom=[np.array([0.2])] * (T + 1))
    o = oracle_mod.Oracle(m, dt, prob, SolverParameters(gradients_method=GRAD_CENTRAL))
    q = np.array([[0.1 + 0.05 * t * t] for t in range(T + 1)])
    o.set_q(q)
    o.eval(2)
    g = o.get("g")
    assert np.all(g[:1] == 0)  # cc:1044
    h = np.cbrt(EPS)
    for t in range(1, T + 1):
        qp, qm = q.copy(), q.copy()
        qp[t] += h
        qm[t] -= h
        o.set_q(qp); o.eval(0); Lp = o.get("cost")[0]
        o.set_q(qm); o.eval(0); Lm = o.get("cost")[0]
        assert abs(g[t] - (Lp - Lm) / (2 * h)) < 1e-5 * max(1.0, abs(g[t]))


def test_trust_ratio_is_one_for_quadratic_problem(oracle_mod):
    """tt:369-428 TrustRatio: pendulum without gravity => exactly quadratic cost => rho = 1 (sqrt eps)."""
    T, dt = 5, 1e-2
    m = load_model("pendulum")
    m.gravity = np.zeros(3)
    prob = pend_problem(T, 0.0, 0.0, Qq=0.1 * np.eye(1), Qv=0.2 * np.eye(1), Qf_q=0.3 * np.eye(1),
                        Qf_v=0.4 * np.eye(1), R=0.5 * np.eye(1), q_nom=[np.array([np.pi])] * (T + 1),
                        v_nom=[np.array([0.0])] * (T + 1))
    params = SolverParameters(scaling=False, gradients_method=GRAD_CENTRAL)
    o = oracle_mod.Oracle(m, dt, prob, params)
    o.set_q(np.array([[0.0 + 0.1 * t] for t in range(T + 1)]))
    for delta in (1e-3, 1e-1, 1e3):
        o.set_delta(delta)
        o.eval(4)
        assert abs(o.get("rho")[0] - 1.0) < np.sqrt(EPS) * 10


def test_dogleg_point_norms(oracle_mod):
    """tt:285-361 DoglegPoint: |dq| == Delta when the TR constraint is active, < Delta for a huge TR."""
    T, dt = 5, 1e-2
    m = load_model("pendulum")
    prob = pend_problem(T, 0.0, 0.0, Qq=0.1 * np.eye(1), Qv=0.2 * np.eye(1), Qf_q=0.3 * np.eye(1),
                        Qf_v=0.4 * np.eye(1), R=0.5 * np.eye(1), q_nom=[np.array([0.1])] * (T + 1),
                        v_nom=[np.array([0.0])] * (T + 1))
    o = oracle_mod.Oracle(m, dt, prob, SolverParameters(scaling=False))
    o.set_q(np.array([[0.0 + 0.1 * t] for t in range(T + 1)]))
    norms = []
    for delta, active in ((1e-3, True), (1e-1, True), (1e6, False)):  # small, medium, large (tt:340-361)
        o.set_delta(delta)
        o.eval(3)
        dq = o.get("dq")
        assert bool(o.get("dq_active")[0]) == active
        if active:
            assert abs(np.linalg.norm(dq) - delta) < EPS / dt * 100
        else:
            assert np.linalg.norm(dq) < delta
            assert np.allclose(dq, o.get("dqH"))
        norms.append(np.linalg.norm(dq))
    assert norms[0] < norms[1] < norms[2]


def test_pendulum_swingup_converges(oracle_mod):
    """tt:434-490 PendulumSwingup: q_T ~= pi within 1e-3."""
    m, dt, prob, params, guess = problems.pendulum()
    params = copy.deepcopy(params)
    params.max_iterations, params.check_convergence = 100, False
    o = oracle_mod.Oracle(m, dt, prob, params)
    o.set_q(guess)
    o.solve(100)
    q, _, _ = o.solution()
    assert abs(q[-1, 0] - np.pi) < 1e-3


def test_fd_vs_cd_partials_with_contact(oracle_mod):
    """tt:183-279 ContactGradientMethods (spinner_sphere.urdf, dt=1, T=2) minus the autodiff legs:
    tau is method independent; FD and CD partials agree to the sum of their test tolerances."""
    m = load_model("spinner_sphere")
    T, dt = 2, 1.0
    prob = ProblemDefinition(num_steps=T, q_init=np.array([0.2, 1.5, 0.0]), v_init=np.zeros(3), Qq=np.eye(3),
                             Qv=np.eye(3), Qf_q=np.eye(3), Qf_v=np.eye(3), R=np.eye(3),
                             q_nom=[np.zeros(3)] * (T + 1), v_nom=[np.zeros(3)] * (T + 1))
    q = np.array([[0.2, 1.5, 0.0], [0.25, 1.45, 0.02], [0.3, 1.4, 0.05]])
    out = {}
    for meth in (GRAD_FORWARD, GRAD_CENTRAL, GRAD_CENTRAL4):
        o = oracle_mod.Oracle(m, dt, prob, SolverParameters(gradients_method=meth))
        o.set_q(q)
        o.eval(1)
        out[meth] = (o.get("tau"), o.get("dtau_dqp"), o.get("dtau_dqt")[9:], o.get("dtau_dqm")[9:])
    assert np.array_equal(out[GRAD_FORWARD][0], out[GRAD_CENTRAL][0])
    for k in (1, 2):
        assert rel_close(out[GRAD_FORWARD][k], out[GRAD_CENTRAL][k], 110 * np.sqrt(EPS))
        assert rel_close(out[GRAD_CENTRAL4][k], out[GRAD_CENTRAL][k], 20 * np.sqrt(EPS))


def test_equality_constraints_and_scaling_invariances(oracle_mod):
    """tt:1637-1750 EqualityConstraintsAndScaling (hopper, no ground): sizes, J*D == J~, lambda equal
    with/without scaling and equal to the dense formula, merit equal, D*gm == gm~, rho equal & > 0.6."""
    m = load_model("hopper_no_ground")
    T, dt = 5, 1e-2
    prob = ProblemDefinition(
        num_steps=T, q_init=np.array([0.0, 0.6, 0.3, -0.5, 0.2]), v_init=np.array([1.0, -0.2, 0.1, -0.3, 0.4]),
        Qq=0.1 * np.eye(5), Qv=0.2 * np.eye(5), Qf_q=0.3 * np.eye(5), Qf_v=0.4 * np.eye(5), R=0.01 * np.eye(5),
        q_nom=[np.array([0.5, 0.5, 0.3, -0.4, 0.1])] * (T + 1), v_nom=[np.array([0.01, 0.0, 0.2, 0.1, -0.1])] * (T + 1))
    q = np.array([prob.q_init + dt * t * prob.v_init for t in range(T + 1)])
    res = {}
    for scaling in (False, True):
        o = oracle_mod.Oracle(m, dt, prob, SolverParameters(scaling=scaling, gradients_method=GRAD_CENTRAL,
                                                            Delta0=1e1))
        o.set_q(q)
        o.eval(4)
        nh, n = 3 * T, (T + 1) * 5
        res[scaling] = dict(h=o.get("h"), J=o.get("J").reshape(n, nh).T, lam=o.get("lambda"), D=o.get("D"),
                            merit=o.get("merit")[0], gm=o.get("gm"), rho=o.get("rho")[0], g=o.get("gs"),
                            Hs=[o.get("Hs_A"), o.get("Hs_B"), o.get("Hs_C")])
        assert o.unactuated_dofs() == [0, 1, 2] and res[scaling]["h"].size == nh  # tt:1598
    a, b = res[False], res[True]
    assert np.all(a["D"] == 1.0)
    assert np.allclose(a["J"] * b["D"][None, :], b["J"], rtol=0, atol=1e-12 * np.max(np.abs(a["J"])))
    scale = max(1.0, np.max(np.abs(a["lam"])))
    assert np.max(np.abs(a["lam"] - b["lam"])) < 1e-6 * scale  # conditioning of J H^-1 J^T, not eps
    assert abs(a["merit"] - b["merit"]) < 1e-8 * max(1, abs(a["merit"]))
    assert np.allclose(b["D"] * a["gm"], b["gm"], rtol=1e-6, atol=1e-6 * np.max(np.abs(b["gm"])))
    assert a["rho"] > 0.6 and b["rho"] > 0.6


# ----------------------------------------------------------------------------- penta-diagonal (pt)
def _random_spd_penta(nb, k, seed):
    rng = np.random.default_rng(seed)
    n = nb * k
    # SPD banded: M = P^T P + I with P block-bidiagonal+1 => penta-diagonal
    P = np.zeros((n, n))
    for i in range(nb):
        for j in range(max(0, i - 2), i + 1):
            P[i * k:(i + 1) * k, j * k:(j + 1) * k] = rng.uniform(-1, 1, (k, k))
    M = P @ P.T + np.eye(n)
    A = np.zeros((nb, k, k)); B = np.zeros((nb, k, k)); C = np.zeros((nb, k, k))
    for i in range(nb):
        C[i] = M[i * k:(i + 1) * k, i * k:(i + 1) * k]
        if i >= 1: B[i] = M[i * k:(i + 1) * k, (i - 1) * k:i * k]
        if i >= 2: A[i] = M[i * k:(i + 1) * k, (i - 2) * k:(i - 1) * k]
    return M, A, B, C


def _cm(X):  # [nb][k][k] row-major numpy -> column-major blocks
    return np.ascontiguousarray(np.transpose(X, (0, 2, 1)))


@pytest.mark.parametrize("nb,k", [(5, 3), (21, 2), (41, 19), (3, 1), (2, 4), (1, 3)])
def test_penta_multiply_solve_dense(oracle_mod, nb, k):
    """pt:20-40 MultiplyBy vs dense; pt:125-257 solves vs dense (eps*size relative)."""
    L = oracle_mod.lib()
    P = oracle_mod._p
    M, A, B, C = _random_spd_penta(nb, k, nb * 100 + k)
    a, b, c = _cm(A), _cm(B), _cm(C)
    n = nb * k
    dense = np.zeros((n, n))
    L.oracle_penta_dense(nb, k, P(a), P(b), P(c), P(dense))
    assert np.array_equal(dense.T, M)
    x = np.random.default_rng(1).uniform(-1, 1, n)
    y = np.zeros(n)
    L.oracle_penta_multiply(nb, k, P(a), P(b), P(c), P(x), P(y))
    assert rel_close(y, M @ x, EPS * n * 10)
    rhs = (M @ x).copy()
    L.oracle_penta_solve(nb, k, P(a), P(b), P(c), P(rhs), 1)
    assert np.max(np.abs(rhs - x)) < EPS * n * np.linalg.cond(M)


def test_penta_identity_and_scale(oracle_mod):
    """pt:109-123 SolveIdentity exact; pt:345-371 ScaleByDiagonal == D*M*D (eps)."""
    L, P = oracle_mod.lib(), oracle_mod._p
    nb, k = 6, 3
    Z = np.zeros((nb, k, k)); I = np.tile(np.eye(k), (nb, 1, 1))
    b = np.arange(nb * k, dtype=float)
    rhs = b.copy()
    L.oracle_penta_solve(nb, k, P(Z.copy()), P(Z.copy()), P(I.copy()), P(rhs), 1)
    assert np.array_equal(rhs, b)
    M, A, B, C = _random_spd_penta(nb, k, 3)
    a, bb, c = _cm(A), _cm(B), _cm(C)
    s = np.random.default_rng(2).uniform(0.1, 2.0, nb * k)
    L.oracle_penta_scale(nb, k, P(a), P(bb), P(c), P(s))
    dense = np.zeros((nb * k, nb * k))
    L.oracle_penta_dense(nb, k, P(a), P(bb), P(c), P(dense))
    assert rel_close(dense.T, s[:, None] * M * s[None, :], 4 * EPS)
