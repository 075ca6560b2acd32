"""Physical self-checks of the oracle across the Drake boundary, where no reference test pins a
number: mass-matrix symmetry / positive definiteness, gravity wrench on a floating
base, N+ consistency with the quaternion kinematics, contact force sanity on the ground."""
import numpy as np

from idto_b200 import problems
from idto_b200.bake import load_model
from idto_b200.types import SolverParameters


def _oracle(oracle_mod, name, **kw):
    m, dt, prob, params, guess = getattr(problems, name)(**kw)
    return m, dt, prob, params, np.array(guess), oracle_mod.Oracle(m, dt, prob, params)


def test_mass_matrix_symmetric_positive_definite(oracle_mod):
    for name in ("acrobot", "spinner", "hopper", "mini_cheetah"):
        m, dt, prob, params, guess, o = _oracle(oracle_mod, name)
        rng = np.random.default_rng(0)
        q = guess[0] + rng.normal(0, 0.2, m.nq)
        M = o.mass_matrix(q)
        assert np.max(np.abs(M - M.T)) < 1e-12 * max(1.0, np.max(np.abs(M)))
        assert np.min(np.linalg.eigvalsh(0.5 * (M + M.T))) > 0


def test_floating_base_gravity_wrench(oracle_mod):
    """With v = a = 0 and the robot far above the ground, the base force rows of tau equal the total
    weight (tau = -J^T F_gravity): f_z = m g on the translational base dofs, everything else on the
    base force is zero."""
    m, dt, prob, params, guess, o = _oracle(oracle_mod, "mini_cheetah")
    q = guess[0].copy()
    q[6] = 5.0  # 5 m up: no contact
    tau, active = o.inverse_dynamics(q, np.zeros(m.nv), np.zeros(m.nv))
    assert not active.any()
    assert np.allclose(tau[3:6], [0, 0, m.mass.sum() * 9.81], atol=1e-10)


def test_contact_supports_weight_direction(oracle_mod):
    """Standing on the ground (feet penetrating slightly), contact reduces the vertical base force."""
    m, dt, prob, params, guess, o = _oracle(oracle_mod, "mini_cheetah")
    q = guess[0].copy()
    tau_c, active = o.inverse_dynamics(q, np.zeros(m.nv), np.zeros(m.nv))
    assert active.all()
    assert tau_c[5] < m.mass.sum() * 9.81  # ground pushes up => less external force needed


def test_nplus_maps_quaternion_rate_to_angular_velocity(oracle_mod):
    """N+(q) qdot = world angular velocity for a rotation about a fixed axis, also for a non-unit
    quaternion (FD perturbs raw quaternion components, cc:514)."""
    m, dt, prob, params, guess, o = _oracle(oracle_mod, "mini_cheetah", T=2)
    axis = np.array([0.3, -0.5, 0.8])
    axis /= np.linalg.norm(axis)
    w, h = 0.7, 1e-6
    for scale in (1.0, 1.3):
        def quat(t):
            ang = 0.4 + w * t
            return scale * np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis])
        q = np.array(guess[:3])
        q[1, :4] = quat(0.0)
        q[0, :4] = quat(-h)
        q[2, :4] = quat(h)
        o.set_q(q)
        o.eval(0)
        N = o.get("Nplus").reshape(3, m.nq, m.nv)[1].T  # (nv, nq)
        qdot = (q[2] - q[0]) / (2 * h)
        assert np.allclose((N @ qdot)[:3], w * axis, atol=1e-6)


def test_pendulum_energy_consistency(oracle_mod):
    """tau * v = d/dt(kinetic + potential) + damping power for the pendulum closed form."""
    m = load_model("pendulum")
    prob = problems.pendulum()[2]
    o = oracle_mod.Oracle(m, 0.05, prob, SolverParameters())
    q, v, a = np.array([0.7]), np.array([0.3]), np.array([-0.2])
    tau, _ = o.inverse_dynamics(q, v, a)
    assert np.isclose(tau[0], 0.25 * a[0] + 0.1 * v[0] + 9.81 * 0.5 * np.sin(q[0]), rtol=1e-14)
