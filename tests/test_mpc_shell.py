"""MPC shell (examples/mpc_controller.cc:43-137): the restated spline shift is pinned against scipy's
not-a-knot CubicSpline (CPU), and the device-side idto_mpc_advance against the restatement (GPU)."""
import numpy as np
import pytest

from oracle import mpc_shell


@pytest.mark.parametrize("N", [2, 3, 4, 5, 10, 40])
def test_spline_restatement_matches_scipy_not_a_knot(N):
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(N)
    h = 0.05
    y = rng.normal(size=(N + 1, 3)).cumsum(axis=0)
    cs = CubicSpline(np.arange(N + 1) * h, y, bc_type="not-a-knot")
    M = mpc_shell.not_a_knot_second_derivatives(y, h)
    for t in np.linspace(0.0, N * h, 57):
        assert np.allclose(mpc_shell.spline_value(y, M, h, t), cs(t), rtol=1e-10, atol=1e-10)
    # clamping outside the domain, like PiecewisePolynomial::value
    assert np.allclose(mpc_shell.spline_value(y, M, h, N * h + 1.0), y[-1], atol=1e-12)
    assert np.allclose(mpc_shell.spline_value(y, M, h, -1.0), y[0], atol=1e-12)


def test_shifted_guess_and_nominal():
    rng = np.random.default_rng(0)
    T, nq, dt = 12, 4, 0.05
    q = rng.normal(size=(T + 1, nq)).cumsum(axis=0)
    q0 = rng.normal(size=nq)
    g = mpc_shell.shifted_guess(q, dt, 0.0, q0)
    assert np.allclose(g[1:], q[1:], atol=1e-12) and np.array_equal(g[0], q0)  # zero elapsed time: knots
    g = mpc_shell.shifted_guess(q, dt, dt, q0)
    assert np.allclose(g[1:-1], q[2:], atol=1e-12) and np.allclose(g[-1], q[-1], atol=1e-12)  # one full step
    sel = np.array([1.0, 0.0, 1.0, 0.0])
    qn = rng.normal(size=(T + 1, nq))
    out = mpc_shell.shifted_nominal(qn, q0, sel)
    assert np.allclose(out[:, 1], qn[:, 1]) and np.allclose(out[0, [0, 2]], q0[[0, 2]])


@pytest.mark.gpu
@pytest.mark.parametrize("name,T", [("hopper", 50), ("mini_cheetah", 40), ("spinner", 4), ("acrobot", 3)])
def test_device_advance_matches_restatement(name, T):
    from idto_b200 import capi, problems
    from idto_b200.types import GRAD_FORWARD
    m, dt, prob, params, guess = getattr(problems, name)(T=T, gradients_method=GRAD_FORWARD)
    B = 3
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, B)
    rng = np.random.default_rng(5)
    q = np.array(guess)[None] + rng.normal(0, 0.05, (B, T + 1, m.nq)).cumsum(axis=1)
    gs.set_q(q)
    el = np.array([0.0, 0.4 * dt, 2.3 * dt])
    q0 = q[:, 0] + rng.normal(0, 0.01, (B, m.nq))
    v0 = rng.normal(0, 0.1, (B, m.nv))
    sel = (np.arange(m.nq) % 2).astype(float)
    qn_old = np.tile(np.asarray(prob.q_nom, float).reshape(1, T + 1, m.nq), (B, 1, 1))
    gs.mpc_advance(el, q0, v0, sel)
    gs.synchronize()
    got = gs.get("q")
    for b in range(B):
        want = mpc_shell.shifted_guess(q[b], dt, el[b], q0[b])
        assert np.allclose(got[b].reshape(T + 1, m.nq), want, rtol=1e-11, atol=1e-11), b
    # v_0 = v_init and the shifted nominal trajectory show up in the trajectory-level entries
    gs.eval(0)
    assert np.allclose(gs.get("v")[:, :m.nv], v0, atol=0)
    qn = gs.get("q_nom").reshape(B, T + 1, m.nq)
    for b in range(B):
        assert np.allclose(qn[b], mpc_shell.shifted_nominal(qn_old[b], q0[b], sel), rtol=1e-13, atol=1e-13)


@pytest.mark.gpu
def test_hopper_mpc_resolve_stream_matches_oracle(oracle_mod):
    """BASELINE config 'hopper planar contact, T=50, warm-started MPC re-solve stream': K re-plans, each one
    trust-region iteration from the spline-shifted previous solution, GPU (device-side shell) vs oracle
    (restated shell) on identical measured states."""
    from idto_b200 import capi, problems
    from idto_b200.types import GRAD_CENTRAL
    m, dt, prob, params, guess = problems.hopper(T=50, gradients_method=GRAD_CENTRAL, max_iterations=1)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    q = np.array(guess)
    gs.set_q(q)
    oc.set_q(q)
    gs.solve(3)
    oc.solve(3)
    rng = np.random.default_rng(11)
    sel = np.zeros(m.nq)
    sel[1] = 1.0  # horizontal position follows the robot
    qn = np.asarray(prob.q_nom, float).reshape(51, m.nq).copy()
    vn = np.asarray(prob.v_nom, float).reshape(51, m.nv).copy()
    for k in range(6):
        qo, vo, _ = oc.solution()
        el = 0.5 * dt
        M = mpc_shell.not_a_knot_second_derivatives(qo, dt)
        q0 = mpc_shell.spline_value(qo, M, dt, el) + rng.normal(0, 1e-3, m.nq)  # "measured" state
        v0 = vo[1] + rng.normal(0, 1e-2, m.nv)
        qn = mpc_shell.shifted_nominal(qn, q0, sel)
        oc.update_nominal_trajectory(qn, vn)
        oc.reset_initial_conditions(q0, v0)
        oc.set_q(mpc_shell.shifted_guess(qo, dt, el, q0))
        oc.solve(1)
        gs.mpc_advance(el, q0, v0, sel)
        it, _, st = gs.solve(1)
        qg, _, taug = gs.solution()
        qo2, _, tauo2 = oc.solution()
        assert np.max(np.abs(qg[0] - qo2)) < 1e-5 * max(1.0, np.max(np.abs(qo2))), k
        assert np.max(np.abs(qg[1] - qo2)) < 1e-5 * max(1.0, np.max(np.abs(qo2))), k
        assert np.max(np.abs(taug[0] - tauo2)) < 1e-3 * max(1.0, np.max(np.abs(tauo2))), k


@pytest.mark.parametrize("N", [2, 3, 4, 7, 40])
def test_host_spline_of_the_stored_trajectory_matches_scipy(N):
    from scipy.interpolate import CubicSpline as Ref
    from idto_b200.mpc import CubicSpline
    rng = np.random.default_rng(N + 100)
    h = 0.05
    y = rng.normal(size=(N + 1, 4)).cumsum(axis=0)
    mine, ref = CubicSpline(h, y), Ref(np.arange(N + 1) * h, y, bc_type="not-a-knot")
    for t in np.linspace(0, N * h, 41):
        assert np.allclose(mine.value(t), ref(t), rtol=1e-10, atol=1e-10)
    assert np.allclose(mine.value(-1.0), y[0]) and np.allclose(mine.value(N * h + 1.0), y[-1])
    assert np.allclose(mine.M, mpc_shell.not_a_knot_second_derivatives(y, h), rtol=1e-10, atol=1e-10)


@pytest.mark.gpu
def test_model_predictive_controller_mirror(oracle_mod):
    """python_examples/mpc_utils.py:87-217 through idto_b200.mpc / idto_b200.pyidto on the hopper: the re-plans
    of the controller equal the oracle's (restated shell), and the Interpolator samples the stored solution."""
    from idto_b200 import problems
    from idto_b200.mpc import Interpolator, ModelPredictiveController
    from idto_b200.pyidto import BakedPlant, TrajectoryOptimizer
    from idto_b200.types import GRAD_FORWARD
    m, dt, prob, params, guess = problems.hopper(T=20, gradients_method=GRAD_FORWARD, max_iterations=2)
    opt = TrajectoryOptimizer(None, BakedPlant(m, dt), prob, params)
    mpc = ModelPredictiveController(opt, guess, m.nq, m.nv, mpc_rate=50)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    oc.set_q(np.array(guess))
    oc.solve(2)
    rng = np.random.default_rng(2)
    t = 0.0
    for k in range(4):
        t += 1.0 / 50
        qo, vo, _ = oc.solution()
        el = t - mpc.stored_trajectory.start_time
        M = mpc_shell.not_a_knot_second_derivatives(qo, dt)
        q0 = mpc_shell.spline_value(qo, M, dt, el) + rng.normal(0, 1e-3, m.nq)
        v0 = vo[0] + rng.normal(0, 1e-2, m.nv)
        oc.reset_initial_conditions(q0, v0)
        oc.set_q(mpc_shell.shifted_guess(qo, dt, el, q0))
        oc.solve(2)
        tr = mpc.UpdateAbstractState(t, np.concatenate((q0, v0)))
        qo2, vo2, tauo2 = oc.solution()
        assert tr.start_time == t and np.allclose(tr.q.y, qo2, atol=1e-5 * max(1.0, np.abs(qo2).max())), k
    B = np.eye(m.nq)[3:]  # the two actuated joints
    itp = Interpolator(B, B)
    x = itp.SendState(t + 0.5 * dt, tr)
    assert x.shape == (4,) and np.allclose(x[:2], B @ tr.q.value(0.5 * dt))
    assert np.allclose(itp.SendControl(t, tr), B @ tauo2[0], atol=1e-3 * max(1.0, np.abs(tauo2).max()))
