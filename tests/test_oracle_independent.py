"""Pins the oracle where the reference tree holds no number (SURVEY.md §8c "parity unpinned"; App. B rows on
RNEA for multi-body trees, planar / quaternion mobilizers, N+ at non-unit quaternions, application of contact
forces): tau(q, v, a) of oracle/idto_oracle.cc against an INDEPENDENT derivation (tests/lagrangian_ref.py:
numerically differentiated position kinematics projected on numerical Jacobians — no recursive Newton-Euler, no
shared code) on every BASELINE model, contact included, and N+ against the textbook kinematic map."""
import numpy as np
import pytest

from idto_b200 import problems
from lagrangian_ref import fk, n_matrix, tau_lagrangian

MODELS = ["acrobot", "spinner", "hopper", "mini_cheetah", "allegro_hand", "allegro_hand_upside_down"]


def _state(name, m, guess, rng, trial):
    q = np.array(guess[0], float) + rng.normal(0, 0.02 if name.startswith("allegro") else 0.1, m.nq)
    if name == "mini_cheetah":
        q[6] -= 0.01  # feet pressed into the ground: contact forces of the order of the weight
        if trial % 2:
            q[:4] *= 1.3  # non-unit quaternion: what a finite-difference perturbation produces (cc:514)
    if name.startswith("allegro") and trial % 2:
        q[16:20] *= 0.8
    return q, rng.normal(0, 0.5, m.nv), rng.normal(0, 2.0, m.nv)


@pytest.mark.parametrize("name", MODELS)
def test_inverse_dynamics_matches_independent_derivation(oracle_mod, name):
    m, dt, prob, params, guess = getattr(problems, name)()
    oc = oracle_mod.Oracle(m, dt, prob, params)
    rng = np.random.default_rng(1)
    contact_seen = 0.0
    for trial in range(4):
        q, v, a = _state(name, m, guess, rng, trial)
        tau_o, act = oc.inverse_dynamics(q, v, a)
        tau_l = tau_lagrangian(m, q, v, a, params, oracle_mod.point_distance)
        scale = max(1.0, np.max(np.abs(tau_o)))
        assert np.max(np.abs(tau_o - tau_l)) < 1e-8 * scale, (trial, np.max(np.abs(tau_o - tau_l)), scale)
        contact_seen = max(contact_seen, np.max(np.abs(tau_l - tau_lagrangian(m, q, v, a))))
        # the mass matrix alone (bias-free evaluations of the ID partials' phase C): tau(q, 0, e_j) - tau(q, 0, 0)
    if m.npairs > 0:
        assert contact_seen > 1e-4  # the contact terms are really part of what was compared


@pytest.mark.parametrize("name", ["hopper", "mini_cheetah", "allegro_hand"])
def test_mass_matrix_matches_independent_derivation(oracle_mod, name):
    m, dt, prob, params, guess = getattr(problems, name)()
    oc = oracle_mod.Oracle(m, dt, prob, params)
    rng = np.random.default_rng(2)
    q, _, _ = _state(name, m, guess, rng, 1)
    M = oc.mass_matrix(q)
    z = np.zeros(m.nv)
    bias = tau_lagrangian(m, q, z, z)
    for j in rng.choice(m.nv, size=min(m.nv, 6), replace=False):
        e = np.zeros(m.nv)
        e[j] = 1.0
        col = tau_lagrangian(m, q, z, e) - bias
        assert np.max(np.abs(M[:, j] - col)) < 1e-8 * max(1.0, np.max(np.abs(M))), j


@pytest.mark.parametrize("name", ["hopper", "mini_cheetah", "allegro_hand"])
def test_nplus_inverts_the_kinematic_map(oracle_mod, name):
    """N+(q) N(q) = I with N the textbook map q' = N(q) v (quaternion: 1/2 (0, w) (x) q), also at non-unit
    quaternions, and N(q) v moves the bodies with exactly the angular velocity v asks for."""
    m, dt, prob, params, guess = getattr(problems, name)(T=2)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    rng = np.random.default_rng(3)
    for trial in range(2):
        q, v, _ = _state(name, m, guess, rng, trial)
        traj = np.stack([np.array(guess[0], float), q, q])
        oc.set_q(traj)
        oc.eval(0)
        Nplus = oc.get("Nplus").reshape(3, m.nq, m.nv)[1].T  # (nv, nq)
        N = n_matrix(m, q)
        assert np.max(np.abs(Nplus @ N - np.eye(m.nv))) < 1e-13
        for k in range(m.nbodies):
            if m.joint_type[k] == 3:  # the floating body turns with w_F = v[vs:vs+3] expressed in its inboard frame
                h = 1e-6
                Rp, _ = fk(m, q + h * (N @ v))
                Rm, _ = fk(m, q - h * (N @ v))
                R0, _ = fk(m, q)
                W = (Rp[k] - Rm[k]) / (2 * h) @ R0[k].T
                w_W = 0.5 * np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]])
                R_PF = np.asarray(m.X_PF, float).reshape(-1, 12)[k, :9].reshape(3, 3)
                vs = int(m.v_start[k])
                assert np.allclose(w_W, R_PF @ v[vs:vs + 3], atol=1e-8)
