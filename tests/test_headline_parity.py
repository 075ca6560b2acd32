"""Parity of the CUDA path against the CPU oracle ON THE CONFIGURATIONS THE METRIC IS QUOTED ON
(BASELINE.json configs 4 and 5): Mini Cheetah T = 40 at batch 64 over ten trust-region iterations and over an
MPC re-solve stream, the allegro hand at its full T = 60 horizon, every scaling method / scaling off /
normalize_quaternions on the quadruped, the device's own record of which contact pairs it applied forces for
(bit-exact against the oracle's, including poses straddling the activation distance), and the KKT sweep checked
on the GPU's OWN bands against a dense LAPACK restatement of the reference's Schur-complement route.

Tolerances (fp64): trajectories and costs after k iterations 1e-6 / 1e-5 / 1e-3 (cost / q* / tau*) with IDENTICAL
accept-reject decisions and trust-region radii; linear algebra on identical inputs: cond(H~) * eps (two backward
stable solvers cannot agree better than that; cond(H~) ~ 2e11 on the quadruped).
"""
import numpy as np
import pytest

from idto_b200 import problems
from idto_b200.types import (GRAD_CENTRAL, GRAD_CENTRAL4, GRAD_FORWARD, SCALING_ADAPTIVE_DOUBLE_SQRT,
                             SCALING_ADAPTIVE_SQRT, SCALING_DOUBLE_SQRT, SCALING_SQRT)
from linalg_ref import kkt_reference

pytestmark = pytest.mark.gpu


def relerr(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    if scale is None:
        scale = max(1.0, float(np.nanmax(np.abs(b))) if b.size else 1.0)
    mask = ~(np.isnan(a) & np.isnan(b))
    return float(np.max(np.abs(a - b)[mask]) / scale) if mask.any() else 0.0


# --------------------------------------------------------------------------------------------------------------
# (a) the benchmark configuration itself: 64 different problems, T = 40, ten iterations
@pytest.mark.parametrize("method", [GRAD_CENTRAL, GRAD_FORWARD])
def test_cheetah_T40_batch64_ten_iterations_match_oracle(oracle_mod, method):
    from idto_b200 import capi
    iters, B = 10, 64
    m, dt, prob, params, guess = problems.mini_cheetah(T=40, gradients_method=method, max_iterations=iters)
    q0, v0, qg = problems.perturbed_batch(m, prob, B)  # bench.py's problems
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, B)
    gs.reset_initial_conditions(q0, v0)
    gs.set_q(qg)
    it, reason, stats = gs.solve(iters)
    q, v, tau = gs.solution()
    assert np.all(it == iters) and np.all(np.isfinite(stats))
    for b in (0, 21, 42, 63):
        oc = oracle_mod.Oracle(m, dt, prob, params)
        oc.reset_initial_conditions(q0[b], v0[b])
        oc.set_q(qg[b])
        k, _, so = oc.solve(iters)
        assert k == iters
        assert np.array_equal(stats[b, :, 1], so[:, 1]), (b, stats[b, :, 1], so[:, 1])  # Delta_k: same decisions
        assert np.array_equal(stats[b, :, 5] > 0, so[:, 5] > 0), b                        # accept / reject
        assert relerr(stats[b, :, 0], so[:, 0]) < 1e-6, b                                # cost per iteration
        assert relerr(stats[b, :, 5], so[:, 5]) < 1e-3, b                                # trust ratios
        qo, vo, tauo = oc.solution()
        assert relerr(q[b], qo) < 1e-5 and relerr(v[b], vo) < 1e-4 and relerr(tau[b], tauo) < 1e-3, b
        assert so[-1, 0] < 0.5 * so[0, 0]  # the solve is doing real work


# (b) MPC re-solve stream on the quadruped: one iteration per re-plan from the spline-shifted previous solution
def test_cheetah_mpc_resolve_stream_matches_oracle(oracle_mod):
    from idto_b200 import capi
    from oracle import mpc_shell
    T = 40
    m, dt, prob, params, guess = problems.mini_cheetah(T=T, gradients_method=GRAD_CENTRAL, max_iterations=1)
    B = 3
    q0b, v0b, qgb = problems.perturbed_batch(m, prob, B)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, B)
    gs.reset_initial_conditions(q0b, v0b)
    gs.set_q(qgb)
    gs.solve(2)
    ocs = []
    for b in range(B):
        oc = oracle_mod.Oracle(m, dt, prob, params)
        oc.reset_initial_conditions(q0b[b], v0b[b])
        oc.set_q(qgb[b])
        oc.solve(2)
        ocs.append(oc)
    rng = np.random.default_rng(5)
    sel = np.zeros(m.nq)
    sel[4:6] = 1.0  # the nominal base x, y follow the robot (mini_cheetah_mpc.py q_nom_relative_to_q_init)
    qn = [np.asarray(prob.q_nom, float).reshape(T + 1, m.nq).copy() for _ in range(B)]
    vn = np.asarray(prob.v_nom, float).reshape(T + 1, m.nv).copy()
    for k in range(5):
        el = np.array([0.3, 0.5, 1.2]) * dt
        q0, v0 = np.zeros((B, m.nq)), np.zeros((B, m.nv))
        for b, oc in enumerate(ocs):
            qo, vo, _ = oc.solution()
            M = mpc_shell.not_a_knot_second_derivatives(qo, dt)
            q0[b] = mpc_shell.spline_value(qo, M, dt, el[b]) + rng.normal(0, 1e-3, m.nq)  # "measured" state
            v0[b] = vo[1] + rng.normal(0, 1e-2, m.nv)
            qn[b] = mpc_shell.shifted_nominal(qn[b], q0[b], sel)
            oc.update_nominal_trajectory(qn[b], vn)
            oc.reset_initial_conditions(q0[b], v0[b])
            oc.set_q(mpc_shell.shifted_guess(qo, dt, el[b], q0[b]))
            oc.solve(1)
        gs.mpc_advance(el, q0, v0, sel)
        it, _, st = gs.solve(1)
        qg, vg, taug = gs.solution()
        delta = gs.get("delta")[:, 0]
        for b, oc in enumerate(ocs):
            qo2, vo2, tauo2 = oc.solution()
            assert delta[b] == oc.get_delta(), (k, b)  # Delta carries over identically
            assert relerr(qg[b], qo2) < 1e-5, (k, b)
            assert relerr(taug[b], tauo2) < 1e-3, (k, b)


# (c) allegro hand at its BASELINE horizon
@pytest.mark.parametrize("method", [GRAD_FORWARD, GRAD_CENTRAL])
def test_allegro_T60_cache_entries_match_oracle(oracle_mod, method):
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.allegro_hand(T=60, gradients_method=method)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    rng = np.random.default_rng(4)
    q = np.array(guess, float)
    q[1:, :16] += rng.normal(0, 0.01, (60, 16)).cumsum(axis=0) * 0.2  # fingers drift; the ball (quaternion) stays
    q[1:, 20:] += rng.normal(0, 5e-4, (60, 3))
    gs.set_q(np.stack([np.array(guess), q]))
    oc.set_q(q)
    gs.eval(4)
    oc.eval(4)
    for f in ("Nplus", "v", "a", "tau", "cost", "h"):
        assert relerr(gs.get(f)[1], oc.get(f)) < 1e-11, f
    sc = max(1.0, np.nanmax(np.abs(oc.get("dtau_dqp"))))
    for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
        assert relerr(gs.get(f)[1], oc.get(f), sc) < 2e-6, f
    for f, tol in (("g", 2e-6), ("H_A", 4e-6), ("H_B", 4e-6), ("H_C", 4e-6), ("D", 2e-6), ("gs", 2e-6)):
        assert relerr(gs.get(f)[1], oc.get(f)) < tol, f
    assert relerr(gs.get("J")[1], oc.get("J")) < 2e-6
    assert relerr(gs.get("merit")[1], oc.get("merit")) < 1e-6
    assert gs.get("dq_active")[1, 0] == oc.get("dq_active")[0]
    assert abs(gs.get("rho")[1, 0] - oc.get("rho")[0]) < 1e-3 * max(1.0, abs(oc.get("rho")[0]))


# (d) every scaling method, scaling off and quaternion normalisation on the quadruped (cc:1225-1255, 2691-2707)
@pytest.mark.parametrize("scaling,method_id,normalize", [
    (True, SCALING_SQRT, False), (True, SCALING_ADAPTIVE_SQRT, False), (True, SCALING_DOUBLE_SQRT, True),
    (True, SCALING_ADAPTIVE_DOUBLE_SQRT, False), (False, SCALING_DOUBLE_SQRT, False),
    (False, SCALING_DOUBLE_SQRT, True)])
def test_scaling_methods_and_quaternion_normalisation_on_cheetah(oracle_mod, scaling, method_id, normalize):
    from idto_b200 import capi
    iters = 4
    m, dt, prob, params, guess = problems.mini_cheetah(T=12, gradients_method=GRAD_CENTRAL, max_iterations=iters)
    params.scaling, params.scaling_method, params.normalize_quaternions = scaling, method_id, normalize
    q0, v0, qg = problems.perturbed_batch(m, prob, 2)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
    gs.reset_initial_conditions(q0, v0)
    gs.set_q(qg)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    oc.reset_initial_conditions(q0[1], v0[1])
    oc.set_q(qg[1])
    # cache entries first: D (adaptive methods start from the cached D = 1, state.h:68) and what is scaled with it
    gs.eval(4)
    oc.eval(4)
    for f, tol in (("D", 2e-6), ("Hs_A", 4e-6), ("Hs_B", 4e-6), ("Hs_C", 4e-6), ("gs", 2e-6), ("J", 2e-6)):
        assert relerr(gs.get(f)[1], oc.get(f)) < tol, f
    if not scaling:
        assert np.array_equal(gs.get("Hs_C")[1], gs.get("H_C")[1]) and np.array_equal(gs.get("gs")[1], gs.get("g")[1])
    it, _, stats = gs.solve(iters)
    k, _, so = oc.solve(iters)
    assert it[1] == k == iters
    assert np.array_equal(stats[1, :, 1], so[:, 1])
    assert relerr(stats[1, :, 0], so[:, 0]) < 1e-6
    q, v, tau = gs.solution()
    qo, vo, tauo = oc.solution()
    assert relerr(q[1], qo) < 1e-5 and relerr(tau[1], tauo) < 1e-3
    # adaptive methods: D after several derivative updates is the running minimum (cc:1241-1251)
    gs.eval(4)
    oc.eval(4)
    assert relerr(gs.get("D")[1], oc.get("D")) < 2e-6
    quat = q[1][:, :4]
    if normalize:  # cc:2691-2707: every accepted q_t has a unit quaternion
        assert np.max(np.abs(np.linalg.norm(quat, axis=1) - 1.0)) < 1e-14
    else:
        assert np.max(np.abs(np.linalg.norm(quat[1:], axis=1) - 1.0)) > 1e-12  # FD steps really leave the sphere


# (e) contact-pair indexing as the DEVICE reports it -------------------------------------------------------------
def _fd_step(qi):
    eps = 1.4901161193847656e-08
    dq = eps * max(1.0, abs(qi))
    return (qi + dq) - qi  # cc:504-508


def _oracle_pair_sets(oc, m, q, T, stencil):
    """Oracle: pairs with phi <= threshold (exactly what ComputeSignedDistancePairwiseClosestPoints(threshold)
    returns, cc:272-279) for tau_t and for every perturbed evaluation of tau_{t-1} at q_t + s dq e_i."""
    v, a = oc.get("v").reshape(T + 1, m.nv), oc.get("a").reshape(T, m.nv)
    base = np.array([oc.inverse_dynamics(q[t + 1], v[t + 1], a[t])[1] for t in range(T)])
    fd = np.zeros((T, m.nq, len(stencil), m.npairs), int)
    for t in range(1, T + 1):
        for i in range(m.nq):
            dq = _fd_step(q[t, i])
            for kk, s in enumerate(stencil):
                qp = q[t].copy()
                qp[i] += s * dq
                fd[t - 1, i, kk] = oc.inverse_dynamics(qp, v[t], a[t - 1])[1]  # the pair set depends on q only
    return base, fd


@pytest.mark.parametrize("name,T,method", [("mini_cheetah", 12, GRAD_CENTRAL), ("mini_cheetah", 6, GRAD_CENTRAL4),
                                           ("hopper", 10, GRAD_FORWARD), ("allegro_hand", 6, GRAD_CENTRAL),
                                           ("allegro_hand", 6, GRAD_FORWARD)])
def test_device_visited_pair_sets_equal_the_oracles(oracle_mod, name, T, method):
    from idto_b200 import capi
    m, dt, prob, params, guess = getattr(problems, name)(T=T, gradients_method=method)
    stencil = {GRAD_FORWARD: (1,), GRAD_CENTRAL: (1, -1), GRAD_CENTRAL4: (1, -1, 2, -2)}[method]
    oc = oracle_mod.Oracle(m, dt, prob, params)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
    rng = np.random.default_rng(9)
    q = np.array(guess, float)
    if name == "allegro_hand":
        q[1:, :16] += rng.normal(0, 0.02, (T, 16))
        q[1:, 20:] += rng.normal(0, 2e-3, (T, 3))
    elif name == "mini_cheetah":
        # legs swing through the activation distance (0.21 m for these contact parameters): feet from well
        # inside to well outside
        q[1:, 7:] += rng.normal(0, 0.25, (T, 12))
        q[1:, 6] += np.linspace(0.0, 0.25, T)
    else:
        q[1:] += rng.normal(0, 0.08, (T, m.nq))
        q[1:, 0] += np.linspace(0.0, 0.3, T)  # the hopper lifts off: the foot spheres leave the activation distance
    gs.set_q(np.stack([q, np.array(guess)]))
    oc.set_q(q)
    oc.eval(0)
    gs.debug_pair_trace(True)
    gs.eval(1)
    base_o, fd_o = _oracle_pair_sets(oc, m, q, T, stencil)
    base_g = gs.get("pair_active")[0].reshape(T, m.npairs).astype(int)
    fd_g = gs.get("pair_active_fd")[0].reshape(T, m.nq, 4, m.npairs).astype(int)[:, :, :len(stencil)]
    assert np.array_equal(base_g, base_o)
    assert 0 < base_o.sum() < base_o.size  # both active and inactive pairs occur
    visited = fd_g >= 0
    assert np.array_equal(fd_g[visited], fd_o[visited])
    # unvisited entries exist only for subtree-only evaluations, and only for pairs whose bodies the perturbed
    # joint does not move — those keep the BASE evaluation's forces, so the base set must equal the oracle's
    # perturbed one there
    tb = np.broadcast_to(base_o[:, None, None, :], fd_o.shape)
    assert np.array_equal(tb[~visited], fd_o[~visited])
    if name == "allegro_hand":
        assert visited.all()
    gs.debug_pair_trace(False)


def test_pair_sets_at_poses_straddling_the_activation_distance(oracle_mod):
    """A foot placed within a few ulps of the activation distance (cc:268-269): the finite-difference
    perturbations put it inside for one stencil point and outside for the other; the device's list of visited
    pairs must flip exactly where the oracle's does (full evaluations and subtree-only evaluations alike)."""
    from idto_b200 import capi
    T = 4
    for path_cols in ("1", "0"):
        import os
        os.environ["IDTO_PATH_COLS"] = path_cols
        try:
            m, dt, prob, params, guess = problems.mini_cheetah(T=T, gradients_method=GRAD_CENTRAL)
            oc = oracle_mod.Oracle(m, dt, prob, params)
            gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 1)
        finally:
            del os.environ["IDTO_PATH_COLS"]
        q = np.array(guess, float)
        # bisect the base height so that the front-left foot sits at the activation distance at t = 2
        v0, a0 = np.zeros(m.nv), np.zeros(m.nv)
        lo, hi = 0.29, 0.8
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            qq = q[2].copy()
            qq[6] = mid
            if oc.inverse_dynamics(qq, v0, a0)[1][0]:
                lo = mid
            else:
                hi = mid
            if hi - lo < 4e-16:
                break
        q[2, 6] = lo  # last height at which the pair is still active; lo + ulp is outside
        gs.set_q(q[None])
        oc.set_q(q)
        oc.eval(0)
        gs.debug_pair_trace(True)
        gs.eval(1)
        base_o, fd_o = _oracle_pair_sets(oc, m, q, T, (1, -1))
        base_g = gs.get("pair_active")[0].reshape(T, m.npairs).astype(int)
        fd_g = gs.get("pair_active_fd")[0].reshape(T, m.nq, 4, m.npairs).astype(int)[:, :, :2]
        assert np.array_equal(base_g, base_o)
        flips = fd_o[1, :, 0, :] != fd_o[1, :, 1, :]
        assert flips.any()  # the stencil really straddles the threshold for some column
        visited = fd_g >= 0
        assert np.array_equal(fd_g[visited], fd_o[visited])
        tb = np.broadcast_to(base_o[:, None, None, :], fd_o.shape)
        assert np.array_equal(tb[~visited], fd_o[~visited])


# (f) the KKT sweep on the GPU's own inputs against LAPACK --------------------------------------------------------
@pytest.mark.parametrize("name,kw", [("mini_cheetah", {"T": 40}), ("hopper", {}), ("acrobot", {}), ("spinner", {}),
                                     ("allegro_hand", {"T": 20})])
def test_kkt_sweep_on_gpu_inputs_matches_dense_lapack(name, kw):
    from idto_b200 import capi
    m, dt, prob, params, guess = getattr(problems, name)(gradients_method=GRAD_CENTRAL, **kw)
    T, nq = prob.num_steps, m.nq
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
    rng = np.random.default_rng(7)
    q = np.array(guess, float)
    q[1:] += rng.normal(0, 0.01 if name == "allegro_hand" else 0.03, q[1:].shape)
    gs.set_q(np.stack([np.array(guess, float), q]))
    gs.eval(3)
    A, B_, C = (gs.get(f)[1].reshape(T + 1, nq * nq) for f in ("Hs_A", "Hs_B", "Hs_C"))
    has_eq = gs.model.nu > 0 and params.equality_constraints
    J = gs.get("J")[1] if has_eq else None
    h = gs.get("h")[1] if has_eq else None
    lam, gm, dqH, condH, condS = kkt_reference(A, B_, C, gs.get("gs")[1], J, h)
    tol = max(1e-9, condH * 2.220446049250313e-16)  # forward error of a backward-stable solve
    if has_eq:
        assert relerr(gs.get("lambda")[1], lam) < max(tol, 1e-9 * condS), (condH, condS)
    assert relerr(gs.get("gm")[1], gm) < max(1e-11, tol * 1e-3)  # gm = gs + J^T lambda: no solve involved
    assert relerr(gs.get("dqH")[1], dqH) < tol, condH


# (g) block cyclic reduction of the same system (kernels_cr.cu) against LAPACK and against the sweep ----------------
@pytest.mark.parametrize("name,kw", [("mini_cheetah", {"T": 40}), ("mini_cheetah", {"T": 13}), ("hopper", {}),
                                     ("spinner", {}), ("allegro_hand", {"T": 20}),
                                     ("hopper", {"equality_constraints": False})])
def test_cyclic_reduction_matches_dense_lapack_and_the_sweep(name, kw):
    """linear_solver = LINSOLVE_CYCLIC_REDUCTION: super-rows of two block rows, odd rows of every level eliminated
    in parallel (odd and even numbers of block rows, with and without equality constraints).  Same tolerances as
    the sweep's LAPACK test; a two-iteration solve takes the same accept/reject decisions as the sweep."""
    from idto_b200 import capi
    from idto_b200.types import LINSOLVE_CYCLIC_REDUCTION, LINSOLVE_TWISTED
    kw = dict(kw)
    eq = kw.pop("equality_constraints", True)
    m, dt, prob, params, guess = getattr(problems, name)(gradients_method=GRAD_CENTRAL, **kw)
    params.equality_constraints = eq
    T, nq = prob.num_steps, m.nq
    rng = np.random.default_rng(7)
    q = np.array(guess, float)
    q[1:] += rng.normal(0, 0.01 if name == "allegro_hand" else 0.03, q[1:].shape)
    q2 = np.stack([np.array(guess, float), q])
    out = {}
    for ls in (LINSOLVE_TWISTED, LINSOLVE_CYCLIC_REDUCTION):
        params.linear_solver = ls
        gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
        gs.set_q(q2)
        gs.eval(3)
        A, B_, C = (gs.get(f)[1].reshape(T + 1, nq * nq) for f in ("Hs_A", "Hs_B", "Hs_C"))
        has_eq = gs.model.nu > 0 and params.equality_constraints
        J = gs.get("J")[1] if has_eq else None
        h = gs.get("h")[1] if has_eq else None
        lam, gm, dqH, condH, condS = kkt_reference(A, B_, C, gs.get("gs")[1], J, h)
        # cyclic reduction pivots inside its 2kb x 2kb blocks only and squares the coupling blocks level by level: it
        # loses about a digit against the sweep on the quadruped (cond 2e11: 8e-4 against 5e-5 = cond x eps)
        tol = max(1e-9, condH * 2.220446049250313e-16) * (50.0 if ls == LINSOLVE_CYCLIC_REDUCTION else 1.0)
        if has_eq:
            assert relerr(gs.get("lambda")[1], lam) < max(tol, 1e-9 * condS), (ls, condH, condS)
        assert relerr(gs.get("dqH")[1], dqH) < tol, (ls, condH)
        it, reason, stats = gs.solve(2)
        out[ls] = (gs.get("lambda").copy(), gs.get("dqH").copy(), stats.copy())
    a, b = out[LINSOLVE_TWISTED], out[LINSOLVE_CYCLIC_REDUCTION]
    assert np.array_equal(a[2][:, :, 5] > 0, b[2][:, :, 5] > 0)  # rho > 0: same accept / reject decisions
    assert np.allclose(a[2][:, :, 1], b[2][:, :, 1])              # same trust-region radii
    assert relerr(b[2][:, :, 0], a[2][:, :, 0]) < 1e-6            # same costs
