"""Allegro hand (BASELINE config 'allegro_hand in-hand sphere rotation, 22 DOF, many contacts'): the SDF bake
the oracle side and the CUDA path.  188 CANDIDATE pairs do not fit the per-evaluation shared-memory scratch of the
inverse-dynamics kernels, so for this model every evaluation compacts its ACTIVE pairs (distance <= threshold, the
pairs the reference visits, cc:272-275) into a list of 64 slots (IDTO_MAX_ACTIVE_PAIRS raises it up to the number
of candidates); more than that at once is an error, not a silent drop."""
import numpy as np
import pytest

from idto_b200 import problems
from idto_b200.types import GRAD_CENTRAL, GRAD_FORWARD
from idto_b200.bake import GEOM_BOX, GEOM_SPHERE, JOINT_QUAT_FLOATING, load_model


def test_allegro_bake_follows_the_example():
    m = load_model("allegro_hand")
    # examples/allegro_hand/allegro_hand.yaml:8-13: outside finger, thumb, middle, inside finger, ball
    assert m.nq == 23 and m.nv == 22 and m.nbodies == 17
    assert m.body_names[:4] == ["link_8", "link_9", "link_10", "link_11"] and m.body_names[4] == "link_12"
    assert m.body_names[-1] == "ball" and m.joint_type[-1] == JOINT_QUAT_FLOATING and m.q_start[-1] == 16
    assert m.unactuated_dofs == [16, 17, 18, 19, 20, 21]  # the ball; every finger joint has an effort limit
    # 20 spheres (19 on the fingers + the ball) and the palm box, which is anchored to the world by the weld
    assert np.bincount(m.geom_type).tolist() == [20, 1] and m.geom_type[0] == GEOM_BOX and m.geom_body[0] == -1
    assert m.geom_type[-1] == GEOM_SPHERE and m.geom_dims[-1][0] == 0.06 and m.geom_body[-1] == 16
    # default collision filter: no pair inside a body or between adjacent bodies; A < B in registration order
    pa, pb = m.pair_geomA, m.pair_geomB
    assert m.npairs == 188 and np.all(pa < pb)
    ball = m.nbodies - 1  # its floating joint hangs off the world itself: no adjacency, it may touch the palm
    for a, b in zip(pa, pb):
        ba, bb = m.geom_body[a], m.geom_body[b]
        assert ba != bb
        if ball not in (ba, bb):
            assert not (ba >= 0 and m.parent[ba] == bb) and not (bb >= 0 and m.parent[bb] == ba)
    assert (0, m.ngeoms - 1) in set(zip(pa.tolist(), pb.tolist()))  # palm box against the ball
    assert sum(1 for b in pb if b == m.ngeoms - 1) == 20  # the ball against every hand geometry


def test_allegro_oracle_physics(oracle_mod):
    m, dt, prob, params, guess = problems.allegro_hand(T=8)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    q0 = np.array(guess[0])
    M = oc.mass_matrix(q0)
    assert np.abs(M - M.T).max() < 1e-15 and np.linalg.eigvalsh(M).min() > 0
    assert np.allclose(np.diag(M)[19:], 0.05) and np.allclose(np.diag(M)[16:19], 0.4 * 0.05 * 0.06 ** 2)
    # without contact (ball far away) the ball's generalized force at rest is its weight, the torque is zero
    q_far = q0.copy()
    q_far[20:] = [0.5, 0.5, 0.5]
    tau, act = oc.inverse_dynamics(q_far, np.zeros(22), np.zeros(22))
    assert not act[m.pair_geomB == m.ngeoms - 1].any()
    assert np.allclose(tau[16:19], 0, atol=1e-14) and np.allclose(tau[19:], [0, 0, 0.05 * 9.81], atol=1e-12)
    # at q_init the ball rests against the fingers: some ball pairs are within the force threshold
    tau, act = oc.inverse_dynamics(q0, np.zeros(22), np.zeros(22))
    assert act[m.pair_geomB == m.ngeoms - 1].sum() >= 5
    oc.set_q(np.array(guess))
    k, _, st = oc.solve(3)
    assert k == 3 and np.all(np.isfinite(st[:, 0])) and st[-1, 0] < st[0, 0]


def test_allegro_upside_down_variant(oracle_mod):
    """`allegro_hand --upside_down` (allegro_hand.cc:94-97): the same plant with gravity reversed."""
    up, dn = load_model("allegro_hand_upside_down"), load_model("allegro_hand")
    assert up.gravity.tolist() == [0.0, 0.0, 9.81] and dn.gravity.tolist() == [0.0, 0.0, -9.81]
    assert up.npairs == dn.npairs and np.array_equal(up.X_PF, dn.X_PF) and np.array_equal(up.mass, dn.mass)
    m, dt, prob, params, guess = problems.allegro_hand_upside_down(T=8, max_iterations=3)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    q_far = np.array(guess[0])
    q_far[20:] = [0.5, 0.5, 0.5]
    tau, act = oc.inverse_dynamics(q_far, np.zeros(22), np.zeros(22))
    ball_pairs = m.pair_geomB == m.ngeoms - 1
    assert not act[ball_pairs].any()  # (some finger-finger and finger-palm pairs do touch in this grasp pose)
    assert abs(tau[21] + 0.05 * 9.81) < 1e-12 and np.abs(tau[16:21]).max() < 1e-12  # the ball "falls" upwards
    oc.set_q(np.array(guess))
    k, _, st = oc.solve(3)
    assert k == 3 and np.all(np.isfinite(st[:, 0])) and st[-1, 0] < st[0, 0]


def _relerr(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(1.0, float(np.nanmax(np.abs(b)))) if scale is None else scale
    mask = ~(np.isnan(a) & np.isnan(b))  # dtau_dqm[0] is NaN on both sides (never defined, cc:497-499)
    return float(np.abs(a - b)[mask].max() / scale)


def _wiggled_guess(guess, seed, amp):
    """Finger joints and ball position jittered, ball quaternion rotated a little and kept unit length."""
    rng = np.random.default_rng(seed)
    q = np.array(guess, float).copy()
    q[1:, :16] += rng.normal(0, amp, q[1:, :16].shape)
    q[1:, 20:] += rng.normal(0, 0.1 * amp, q[1:, 20:].shape)
    q[1:, 16:20] += rng.normal(0, amp, q[1:, 16:20].shape)
    q[1:, 16:20] /= np.linalg.norm(q[1:, 16:20], axis=1, keepdims=True)
    return q


@pytest.mark.gpu
@pytest.mark.parametrize("method", [GRAD_FORWARD, GRAD_CENTRAL])
def test_allegro_cache_entries_match_oracle(oracle_mod, method):
    """188 candidate pairs, a dozen or two of them active: every evaluation on the GPU compacts its own active
    list (dynamics_chain.cuh), the oracle walks all candidates like the reference (cc:272-386)."""
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.allegro_hand(T=6, gradients_method=method)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    q = _wiggled_guess(guess, 11, 0.02)
    gs.set_q(np.stack([q, np.array(guess)]))
    oc.set_q(q)
    gs.eval(4)
    oc.eval(4)
    v, a = oc.get("v").reshape(-1, m.nv), oc.get("a").reshape(-1, m.nv)
    nact = [int(np.sum(oc.inverse_dynamics(q[t + 1], v[t + 1], a[t])[1])) for t in range(prob.num_steps)]
    assert 4 <= min(nact) and max(nact) <= 64, nact  # contact is really exercised, and fits the pair slots
    for f in ("Nplus", "v", "a", "tau", "cost", "h"):
        assert _relerr(gs.get(f)[0], oc.get(f)) < 1e-11, f
    sc = max(1.0, np.abs(oc.get("dtau_dqp")).max())
    for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
        assert _relerr(gs.get(f)[0], oc.get(f), sc) < 2e-6, f
    for f, tol in (("g", 2e-6), ("H_A", 4e-6), ("H_B", 4e-6), ("H_C", 4e-6), ("D", 2e-6), ("gs", 2e-6)):
        assert _relerr(gs.get(f)[0], oc.get(f)) < tol, f
    assert _relerr(gs.get("J")[0], oc.get("J")) < 2e-6
    assert _relerr(gs.get("gm")[0], oc.get("gm")) < 1e-3
    assert _relerr(gs.get("dq")[0], oc.get("dq")) < 2e-2
    assert gs.get("dq_active")[0, 0] == oc.get("dq_active")[0]
    assert abs(gs.get("rho")[0, 0] - oc.get("rho")[0]) < 1e-3 * max(1.0, abs(oc.get("rho")[0]))
    # the second batch element (the unperturbed guess) against its own oracle run: no cross-talk
    oc.set_q(np.array(guess))
    oc.eval(4)
    assert _relerr(gs.get("tau")[1], oc.get("tau")) < 1e-11
    assert _relerr(gs.get("dtau_dqt")[1], oc.get("dtau_dqt"), sc) < 2e-6


@pytest.mark.gpu
def test_allegro_solve_matches_oracle(oracle_mod):
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.allegro_hand(T=6, gradients_method=GRAD_CENTRAL, max_iterations=4)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 1)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    gs.set_q(np.array(guess))
    oc.set_q(np.array(guess))
    it, _, stats = gs.solve(4)
    k, _, so = oc.solve(4)
    assert it[0] == k == 4
    assert np.array_equal(stats[0, :, 1], so[:, 1])  # accept / reject decisions and trust-region radii
    assert _relerr(stats[0, :, 0], so[:, 0]) < 1e-5  # cost per iteration
    q, v, tau = gs.solution()
    qo, vo, tauo = oc.solution()
    assert _relerr(q[0], qo) < 1e-5 and _relerr(tau[0], tauo) < 1e-3
    assert so[-1, 0] < so[0, 0]


@pytest.mark.gpu
def test_allegro_full_horizon_runs_and_descends():
    """BASELINE config 5 (T = 60): a batch of 4 solves, 3 iterations; the cost goes down and nothing overflows."""
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.allegro_hand(T=60, max_iterations=3)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 4)
    gs.set_q(np.stack([np.array(guess)] * 4))
    it, _, stats = gs.solve(3)
    assert np.all(it == 3) and np.all(np.isfinite(stats[:, :, 0]))
    assert np.all(stats[:, -1, 0] < stats[:, 0, 0])
    assert np.array_equal(stats[0], stats[3])  # identical problems, identical answers


@pytest.mark.gpu
def test_allegro_too_many_simultaneous_contacts_is_an_error():
    """With a 50x larger smoothing length the activation distance covers the whole hand: far more than 32 pairs
    are active at once, which the per-evaluation list reports instead of dropping forces."""
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.allegro_hand(T=4)
    params.smoothing_factor = 0.05
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 1)
    gs.set_q(np.array(guess))
    with pytest.raises(capi.IdtoError, match="contact pairs within the activation distance"):
        gs.eval(1)


@pytest.mark.gpu
def test_allegro_pair_capacity_follows_the_environment(oracle_mod, monkeypatch):
    """The same over-active configuration with IDTO_MAX_ACTIVE_PAIRS = number of candidates: every pair fits, the
    CTA shape adapts (one column per CTA), and tau / the partials agree with the oracle, which has no limit —
    like the reference (cc:272-386)."""
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.allegro_hand(T=4, gradients_method=GRAD_FORWARD)
    params.smoothing_factor = 0.05
    monkeypatch.setenv("IDTO_MAX_ACTIVE_PAIRS", str(m.npairs))
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 1)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    q = np.array(guess)
    gs.set_q(q)
    oc.set_q(q)
    gs.eval(1)
    oc.eval(1)
    v, a = oc.get("v").reshape(-1, m.nv), oc.get("a").reshape(-1, m.nv)
    assert int(oc.inverse_dynamics(q[1], v[1], a[0])[1].sum()) > 64
    assert _relerr(gs.get("tau")[0], oc.get("tau")) < 1e-11
    sc = max(1.0, np.abs(oc.get("dtau_dqp")).max())
    for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
        assert _relerr(gs.get(f)[0], oc.get(f), sc) < 2e-6, f
