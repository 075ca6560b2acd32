"""Parity of the CUDA path (through the C ABI) against the CPU oracle, cache entry by cache entry,
on the same seeded inputs.  Tolerances (fp64):
  * trajectory level (v, a, tau, cost, h, N+): 1e-11 relative — different summation order / FMA
    contraction / sincos implementations only;
  * finite-difference partials: 2e-6 * max(1, |x|, scale) — a 1-ulp difference in tau is amplified by
    1/dq ~ 6.7e7 (SURVEY.md §7 "Finite-difference conditioning"; the reference's own FD-vs-AD test
    tolerances are 10..100*sqrt(eps), optimizer/test/trajectory_optimizer_test.cc:256-257);
  * everything assembled from the partials (g, H, D, J, lambda, dq, rho): stated per assertion.
"""
import numpy as np
import pytest

from idto_b200 import problems
from idto_b200.types import GRAD_CENTRAL, GRAD_CENTRAL4, GRAD_FORWARD

pytestmark = pytest.mark.gpu


def relerr(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    if scale is None:
        scale = max(1.0, float(np.nanmax(np.abs(b))) if b.size else 1.0)
    mask = ~(np.isnan(a) & np.isnan(b))
    return float(np.max(np.abs(a - b)[mask]) / scale) if mask.any() else 0.0


def make_pair(oracle_mod, name, method, batch=1, **kw):
    from idto_b200 import capi
    m, dt, prob, params, guess = getattr(problems, name)(gradients_method=method, **kw)
    model = capi.Model(m)
    gs = capi.BatchSolver(model, dt, prob, params, batch)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    return m, dt, prob, params, np.array(guess), gs, oc


def wiggle(guess, seed, amp=0.05):
    rng = np.random.default_rng(seed)
    q = np.array(guess, float).copy()
    q[1:] += rng.normal(0, amp, q[1:].shape)
    return q


CASES = [("acrobot", {}), ("spinner", {}), ("hopper", {}), ("mini_cheetah", {})]


@pytest.mark.parametrize("name,kw", CASES)
@pytest.mark.parametrize("method", [GRAD_FORWARD, GRAD_CENTRAL, GRAD_CENTRAL4])
def test_cache_entries_match_oracle(oracle_mod, name, kw, method):
    m, dt, prob, params, guess, gs, oc = make_pair(oracle_mod, name, method, **kw)
    q = wiggle(guess, 7, 0.03 if name == "mini_cheetah" else 0.05)
    gs.set_q(q)
    oc.set_q(q)
    gs.eval(4)
    oc.eval(4)
    for f in ("Nplus", "v", "a", "tau", "cost", "h"):
        assert relerr(gs.get(f)[0], oc.get(f)) < 1e-11, f
    sc = max(1.0, np.nanmax(np.abs(oc.get("dtau_dqp"))))
    for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
        assert relerr(gs.get(f)[0], oc.get(f), sc) < 2e-6, f
    for f, tol in (("g", 2e-6), ("H_A", 4e-6), ("H_B", 4e-6), ("H_C", 4e-6), ("D", 2e-6), ("Hs_A", 4e-6),
                   ("Hs_B", 4e-6), ("Hs_C", 4e-6), ("gs", 2e-6)):
        assert relerr(gs.get(f)[0], oc.get(f)) < tol, f
    if oc.nu > 0 and params.equality_constraints:
        assert relerr(gs.get("J")[0], oc.get("J")) < 2e-6
        lam_o = oc.get("lambda")
        assert relerr(gs.get("lambda")[0], lam_o) < 1e-3, "lambda"  # cond(J H^-1 J^T) * FD noise
        assert relerr(gs.get("merit")[0], oc.get("merit")) < 1e-6
    assert relerr(gs.get("gm")[0], oc.get("gm")) < 1e-3
    # The Gauss-Newton step amplifies the FD noise of the partials by cond(H~): compare loosely with
    # the oracle, and tightly against the GPU's own linear system H~ (dqH/Delta-scaled) = -gm.
    assert relerr(gs.get("dqH")[0], oc.get("dqH")) < 2e-2
    assert relerr(gs.get("dq")[0], oc.get("dq")) < 2e-2
    nq, nblk = m.nq, prob.num_steps + 1
    A, Bb, C = (gs.get(f)[0].reshape(nblk, nq, nq).transpose(0, 2, 1) for f in ("Hs_A", "Hs_B", "Hs_C"))
    x = gs.get("dqH")[0].reshape(nblk, nq)
    y = np.einsum("irc,ic->ir", C, x)
    y[1:] += np.einsum("irc,ic->ir", Bb[1:], x[:-1])
    y[2:] += np.einsum("irc,ic->ir", A[2:], x[:-2])
    y[:-1] += np.einsum("icr,ic->ir", Bb[1:], x[1:])
    y[:-2] += np.einsum("icr,ic->ir", A[2:], x[2:])
    gm = gs.get("gm")[0].reshape(nblk, nq)
    assert np.max(np.abs(y + gm)) < 1e-9 * max(1.0, np.max(np.abs(gm))) * max(1.0, np.max(np.abs(x)))
    assert gs.get("dq_active")[0, 0] == oc.get("dq_active")[0]
    assert abs(gs.get("rho")[0, 0] - oc.get("rho")[0]) < 1e-3 * max(1.0, abs(oc.get("rho")[0]))


def test_contact_pair_indexing_is_exact(oracle_mod):
    """Contact-pair set per time step must be identical (bit-exact indexing): compare the active-pair
    masks implied by tau differences with/without contact through the oracle's pair report, on the
    cheetah standing on the ground (4 foot-ground pairs, ascending registration order)."""
    m, dt, prob, params, guess, gs, oc = make_pair(oracle_mod, "mini_cheetah", GRAD_CENTRAL)
    assert list(zip(m.pair_geomA.tolist(), m.pair_geomB.tolist())) == [(0, 1), (0, 2), (0, 3), (0, 4)]
    assert m.geom_body.tolist() == [-1, 3, 6, 9, 12]
    q = wiggle(guess, 3, 0.02)
    gs.set_q(q)
    oc.set_q(q)
    gs.eval(0)
    oc.eval(0)
    v, a = oc.get("v").reshape(-1, m.nv), oc.get("a").reshape(-1, m.nv)
    active = np.array([oc.inverse_dynamics(q[t + 1], v[t + 1], a[t])[1] for t in range(prob.num_steps)])
    assert active.shape == (prob.num_steps, 4) and active.all()  # feet are within the force threshold
    assert relerr(gs.get("tau")[0], oc.get("tau")) < 1e-11


@pytest.mark.parametrize("name,iters,tol", [("spinner", 30, 1e-5), ("hopper", 10, 1e-5), ("acrobot", 20, 1e-5)])
def test_solve_matches_oracle(oracle_mod, name, iters, tol):
    """q*, tau*, final cost after `iters` trust-region iterations from identical inputs."""
    m, dt, prob, params, guess, gs, oc = make_pair(oracle_mod, name, GRAD_CENTRAL)
    gs.set_q(guess)
    oc.set_q(guess)
    it, reason, stats = gs.solve(iters)
    k, _, so = oc.solve(iters)
    assert it[0] == k == iters
    # accept/reject decisions and trust-region radii identical
    assert np.array_equal(stats[0, :, 1], so[:, 1])
    assert relerr(stats[0, :, 0], so[:, 0]) < tol
    q, v, tau = gs.solution()
    qo, vo, tauo = oc.solution()
    assert relerr(q[0], qo) < tol and relerr(tau[0], tauo) < 100 * tol


def test_spinner_end_to_end_golden_on_gpu(oracle_mod):
    """python_bindings/test/trajectory_optimizer_test.py:84-85 through the CUDA path (forward
    differences, 200 iterations): q_T ~= [0.287, 1.497, 1.995] within 1e-3."""
    m, dt, prob, params, guess, gs, oc = make_pair(oracle_mod, "spinner", GRAD_FORWARD)
    gs.set_q(guess)
    it, _, _ = gs.solve(200)
    q, _, _ = gs.solution()
    assert it[0] == 200
    assert np.linalg.norm(q[0, -1] - np.array([0.287, 1.497, 1.995])) < 1e-3


def test_batch_elements_are_independent(oracle_mod):
    """Batch element b must equal a single solve of the same perturbed problem (no cross-talk)."""
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.hopper(gradients_method=GRAD_CENTRAL)
    B = 5
    q0, v0, qg = problems.perturbed_batch(m, prob, B)
    model = capi.Model(m)
    gs = capi.BatchSolver(model, dt, prob, params, B)
    gs.reset_initial_conditions(q0, v0)
    gs.set_q(qg)
    it, _, stats = gs.solve(3)
    qb, _, _ = gs.solution()
    for b in (0, 3):
        one = capi.BatchSolver(model, dt, prob, params, 1)
        one.reset_initial_conditions(q0[b:b + 1], v0[b:b + 1])
        one.set_q(qg[b:b + 1])
        _, _, s1 = one.solve(3)
        q1, _, _ = one.solution()
        assert np.array_equal(s1[0], stats[b]) and np.array_equal(q1[0], qb[b])


def test_warm_start_equals_one_shot_on_gpu(oracle_mod):
    """python_bindings/test/warm_start_test.py:165-182 through the C ABI: 10 x solve(1) == solve(10)."""
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.spinner(max_iterations=10)
    model = capi.Model(m)
    a = capi.BatchSolver(model, dt, prob, params, 1)
    a.set_q(guess)
    _, _, sa = a.solve(10)
    b = capi.BatchSolver(model, dt, prob, params, 1)
    b.set_q(guess)
    sb = np.concatenate([b.solve(1)[2][0] for _ in range(10)])
    qa, va, _ = a.solution()
    qb, vb, _ = b.solution()
    assert np.max(np.abs(qa - qb)) < 1e-8 and np.max(np.abs(va - vb)) < 1e-8
    assert np.max(np.abs(sa[0] - sb)) < 1e-8


def test_pyidto_mirror_spinner_like_reference_test(oracle_mod):
    """python_bindings/test/trajectory_optimizer_test.py:17-106, same shape through idto_b200.pyidto."""
    from idto_b200.pyidto import (BakedPlant, FindIdtoResource, ProblemDefinition, SolverParameters,
                                  TrajectoryOptimizer, TrajectoryOptimizerSolution, TrajectoryOptimizerStats)
    time_step = 0.05
    plant = BakedPlant("spinner", time_step)
    problem = ProblemDefinition()
    problem.num_steps = 40
    problem.q_init = np.array([0.3, 1.5, 0.0])
    problem.v_init = np.array([0.0, 0.0, 0.0])
    problem.Qq = 1.0 * np.eye(3)
    problem.Qv = 0.1 * np.eye(3)
    problem.R = np.diag([0.1, 0.1, 1e3])
    problem.Qf_q = 10 * np.eye(3)
    problem.Qf_v = 0.1 * np.eye(3)
    problem.q_nom = [np.array([0.3, 1.5, 2.0]) for _ in range(41)]
    problem.v_nom = [np.zeros(3) for _ in range(41)]
    params = SolverParameters()
    params.max_iterations = 200
    params.Delta0, params.Delta_max = 1e1, 1e5
    params.contact_stiffness, params.dissipation_velocity, params.smoothing_factor = 200, 0.1, 0.01
    params.friction_coefficient, params.stiction_velocity, params.verbose = 0.5, 0.05, False
    q_guess = [np.array([0.3, 1.5, 0.0]) for _ in range(41)]
    opt = TrajectoryOptimizer(None, plant, problem, params)
    assert opt.time_step() == time_step and opt.num_steps() == 40
    solution, stats = TrajectoryOptimizerSolution(), TrajectoryOptimizerStats()
    opt.Solve(q_guess, solution, stats)
    assert len(solution.q) == 41 and len(solution.tau) == 40 and len(stats.iteration_costs) == 200
    assert np.linalg.norm(solution.q[-1] - np.array([0.287, 1.497, 1.995])) < 1e-3
    assert opt.prob().num_steps == 40 and opt.params().max_iterations == 200
    new_q_nom = [x.copy() for x in problem.q_nom]
    new_q_nom[10] = np.array([0.4, 1.6, 2.1])
    opt.UpdateNominalTrajectory(new_q_nom, problem.v_nom)
    assert np.all(opt.prob().q_nom[10] == np.array([0.4, 1.6, 2.1]))
    with pytest.raises(RuntimeError):
        FindIdtoResource("models/x.urdf")
    # warm start: Delta persists, stats append (warm_start_test.py:104-112)
    ws = opt.CreateWarmStart(q_guess)
    params1 = opt.params()
    st2 = TrajectoryOptimizerStats()
    opt._params.max_iterations = 1
    for _ in range(3):
        opt.SolveFromWarmStart(ws, solution, st2)
    assert len(st2.iteration_costs) == 3 and ws.Delta > 0 and len(ws.get_q()) == 41


def test_substreams_do_not_change_results(oracle_mod):
    """Sub-batch streams (default 4 when batch >= 32) only change scheduling: every batch element must be
    bit-identical to the single-stream run, through both idto_solve and idto_resolve_async."""
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.hopper(T=20, gradients_method=GRAD_CENTRAL)
    B = 37
    q0, v0, qg = problems.perturbed_batch(m, prob, B)
    model = capi.Model(m)
    out = {}
    for ns in (1, 4):
        gs = capi.BatchSolver(model, dt, prob, params, B)
        gs.set_substreams(ns)
        gs.reset_initial_conditions(q0, v0)
        gs.set_q(qg)
        it, _, stats = gs.solve(3)
        for _ in range(2):  # asynchronous re-solves without host pointers
            gs.invalidate()
            gs.resolve_async(1)
        gs.synchronize()
        q, v, tau = gs.solution()
        out[ns] = (it.copy(), stats.copy(), q.copy(), tau.copy(), gs.get("delta").copy())
    for a, b in zip(out[1], out[4]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["mini_cheetah", "hopper", "spinner"])
@pytest.mark.parametrize("method", [GRAD_FORWARD, GRAD_CENTRAL, GRAD_CENTRAL4])
def test_path_columns_match_full_evaluations(name, method, monkeypatch):
    """The single-lane subtree evaluations (kernels_path.cu) against full evaluations of the same columns
    (IDTO_PATH_COLS=0): same arithmetic per body, so almost every entry is bit-identical and the rest differ
    by last-bit effects amplified by 1/dq."""
    from idto_b200 import capi
    m, dt, prob, params, guess = getattr(problems, name)(gradients_method=method)
    rng = np.random.default_rng(3)
    q = np.array(guess, float)[None].repeat(2, 0)
    q[:, 1:] += rng.normal(0, 0.03, q[:, 1:].shape)
    out = {}
    for tag, env in (("full", "0"), ("path", "1")):
        monkeypatch.setenv("IDTO_PATH_COLS", env)
        gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
        gs.set_q(q)
        gs.eval(1)
        out[tag] = {f: gs.get(f) for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp")}
    for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
        a, b = out["full"][f], out["path"][f]
        assert np.array_equal(np.isnan(a), np.isnan(b)), f
        mask = ~np.isnan(a)
        scale = max(1.0, np.max(np.abs(a[mask])))
        assert np.max(np.abs(a[mask] - b[mask])) < 1e-9 * scale, f
        assert np.mean(a[mask] == b[mask]) > 0.99, f


@pytest.mark.parametrize("name,T,eq,batch", [
    ("hopper", 7, True, 1),        # fewer than 8 block rows: single top-down KKT sweep
    ("hopper", 8, True, 3),        # shortest two-sided sweep (9 block rows)
    ("hopper", 9, False, 2),       # no equality constraints: block size nq
    ("spinner", 13, True, 2),      # odd horizon, every column a path column
    ("acrobot", 12, True, 1),
    ("mini_cheetah", 9, True, 33),  # batch that does not divide by the sub-batch streams
    ("mini_cheetah", 12, False, 2),
])
def test_short_horizons_and_block_sizes(oracle_mod, name, T, eq, batch):
    """Edge cases of the horizon (boundary rows of the three partial bands, the KKT sweep variants) and of the
    KKT block size, two trust-region iterations against the oracle."""
    from idto_b200 import capi
    m, dt, prob, params, guess = getattr(problems, name)(T=T, gradients_method=GRAD_CENTRAL)
    params.equality_constraints = eq
    params.max_iterations = 2
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, batch)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    q = wiggle(guess, 11, 0.02)
    gs.set_q(q)
    oc.set_q(q)
    gs.eval(1)
    oc.eval(1)
    sc = max(1.0, np.nanmax(np.abs(oc.get("dtau_dqp"))))
    for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
        for b in (0, batch - 1):
            assert relerr(gs.get(f)[b], oc.get(f), sc) < 2e-6, (f, b)
    it, _, stats = gs.solve(2)
    k, _, so = oc.solve(2)
    assert it[0] == k == 2 and it[batch - 1] == 2
    for b in (0, batch - 1):
        assert np.array_equal(stats[b, :, 1], so[:, 1])  # same trust-region radii: same accept / reject decisions
        assert relerr(stats[b, :, 0], so[:, 0]) < 1e-6
    qg, _, taug = gs.solution()
    qo, _, tauo = oc.solution()
    assert relerr(qg[0], qo) < 1e-5 and relerr(qg[batch - 1], qo) < 1e-5
