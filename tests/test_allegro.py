"""Allegro hand (BASELINE config 'allegro_hand in-hand sphere rotation, 22 DOF, many contacts'): the SDF bake
and the oracle side.  The CUDA path does not run it yet — the inverse-dynamics kernels size their contact
scratch by the 188 CANDIDATE pairs — and must say so instead of failing in a launch."""
import numpy as np
import pytest

from idto_b200 import problems
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


@pytest.mark.gpu
def test_allegro_is_rejected_cleanly_on_the_gpu():
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.allegro_hand(T=8)
    with pytest.raises(capi.IdtoError, match="candidate contact pairs"):
        capi.BatchSolver(capi.Model(m), dt, prob, params, 1)
