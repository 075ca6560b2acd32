"""GPU-vs-oracle parity of the options SURVEY.md §8(f4) lists after the headline path: the sphere-cylinder closed
form on the device, and the reference's kDenseLdlt debugging solver (cc:2088-2093)."""
import copy

import numpy as np
import pytest

from idto_b200 import problems
from idto_b200.bake import GEOM_CAPSULE, GEOM_CYLINDER
from idto_b200.types import GRAD_CENTRAL, GRAD_FORWARD, LINSOLVE_DENSE_LDLT, LINSOLVE_TWISTED

pytestmark = pytest.mark.gpu


def _rel(x, y, s=None):
    return float(np.nanmax(np.abs(np.asarray(x) - np.asarray(y))) / (s or max(1.0, np.nanmax(np.abs(y)))))


@pytest.mark.parametrize("method", [GRAD_FORWARD, GRAD_CENTRAL])
def test_sphere_cylinder_contact_matches_oracle(oracle_mod, method):
    """The spinner_capsule model with its capsule re-registered as a Cylinder of the same radius and length (Drake's
    DistanceToPoint<Cylinder>: flat caps, barrel, rims): eleven finger spheres against it, GPU vs oracle."""
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.spinner_capsule(gradients_method=method)
    m = copy.deepcopy(m)
    m.geom_type = np.where(m.geom_type == GEOM_CAPSULE, GEOM_CYLINDER, m.geom_type).astype(m.geom_type.dtype)
    assert (m.geom_type == GEOM_CYLINDER).sum() == 1
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    rng = np.random.default_rng(1)
    q = np.array(guess, float)
    q[1:] += rng.normal(0, 0.05, q[1:].shape).cumsum(axis=0) * 0.3
    gs.set_q(q)
    oc.set_q(q)
    gs.eval(1)
    oc.eval(1)
    v, a = oc.get("v").reshape(-1, m.nv), oc.get("a").reshape(-1, m.nv)
    active = np.array([oc.inverse_dynamics(q[t + 1], v[t + 1], a[t])[1] for t in range(prob.num_steps)])
    assert active.any() and not active.all()
    assert _rel(gs.get("tau")[0], oc.get("tau")) < 1e-11
    sc = max(1.0, np.nanmax(np.abs(oc.get("dtau_dqp"))))
    for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
        assert _rel(gs.get(f)[1], oc.get(f), sc) < 2e-6, f
    gs.set_q(guess)
    oc.set_q(guess)
    it, _, stats = gs.solve(12)
    k, _, so = oc.solve(12)
    assert np.array_equal(stats[0, :, 1], so[:, 1]) and _rel(stats[0, :, 0], so[:, 0]) < 1e-6


@pytest.mark.parametrize("name,kw", [("spinner", {}), ("hopper", {"T": 20}), ("mini_cheetah", {"T": 12}), ("acrobot", {})])
def test_dense_ldlt_solver_cross_checks_the_penta_diagonal_one(oracle_mod, name, kw):
    """linear_solver = dense_ldlt re-solves H~ x = -gm without using the block structure; the Gauss-Newton step,
    the dogleg point and a few iterations agree with the penta-diagonal sweep (and so with the oracle)."""
    from idto_b200 import capi
    out = {}
    for ls in (LINSOLVE_TWISTED, LINSOLVE_DENSE_LDLT):
        m, dt, prob, params, guess = getattr(problems, name)(gradients_method=GRAD_CENTRAL, **kw)
        params.linear_solver = ls
        gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
        rng = np.random.default_rng(7)
        q = np.array(guess, float)
        q[1:] += rng.normal(0, 0.03, q[1:].shape)
        gs.set_q(np.stack([np.array(guess, float), q]))
        gs.eval(3)
        res = {f: gs.get(f)[1].copy() for f in ("dqH", "dq", "lambda", "gm")}
        gs.set_q(np.array(guess, float))
        it, _, stats = gs.solve(3)
        res["stats"], res["q"] = stats[0].copy(), gs.solution()[0][0].copy()
        out[ls] = res
    a, b = out[LINSOLVE_TWISTED], out[LINSOLVE_DENSE_LDLT]
    assert np.array_equal(a["lambda"], b["lambda"]) and np.array_equal(a["gm"], b["gm"])  # same sweep before it
    assert not np.array_equal(a["dqH"], b["dqH"])  # a different factorisation really ran
    # two backward-stable solvers: cond(H~) eps apart (cond ~ 1e10 on the acrobot, 2e11 on the quadruped)
    assert _rel(b["dqH"], a["dqH"]) < 1e-4 and _rel(b["dq"], a["dq"]) < 1e-4
    assert np.array_equal(a["stats"][:, 1], b["stats"][:, 1])  # same accept / reject decisions
    assert _rel(b["stats"][:, 0], a["stats"][:, 0]) < 1e-6 and _rel(b["q"], a["q"]) < 1e-5


def test_multi_device_batch_solver_equals_single_device(oracle_mod):
    """One process, every visible GPU (one host thread + solver per device, contiguous slices, no collective):
    problem b of the sharded batch equals problem b of a single-device batch, bit for bit."""
    from idto_b200 import capi
    m, dt, prob, params, guess = problems.hopper(T=20, gradients_method=GRAD_CENTRAL)
    B = 7
    q0, v0, qg = problems.perturbed_batch(m, prob, B)
    md = capi.MultiDeviceBatchSolver(m, dt, prob, params, B)
    md.reset_initial_conditions(q0, v0)
    md.set_q(qg)
    it, _, stats = md.solve(3)
    q, v, tau = md.solution()
    one = capi.BatchSolver(capi.Model(m), dt, prob, params, B)
    one.reset_initial_conditions(q0, v0)
    one.set_q(qg)
    it1, _, stats1 = one.solve(3)
    q1, v1, tau1 = one.solution()
    assert len(md.shards) == min(capi.device_count(), B) and sum(s.B for s in md.shards) == B
    assert np.array_equal(it, it1) and np.array_equal(stats, stats1)
    assert np.array_equal(q, q1) and np.array_equal(tau, tau1)
    assert md.get("cost").shape == (B, 1)


def test_trust_ratio_model_terms_match_the_explicit_matvec(monkeypatch):
    """k_dogleg_post forms s.H~s from dot products (H~ pH = -gm / Delta is what the KKT sweep solved); k_trust_update
    (IDTO_TRUST_MATVEC=1) multiplies H~ s out like the reference (cc:2008-2017).  Same rho, same decisions."""
    from idto_b200 import capi
    out = {}
    for tag, env in (("dots", None), ("matvec", "1")):
        if env:
            monkeypatch.setenv("IDTO_TRUST_MATVEC", env)
        res = []
        for name, kw in (("mini_cheetah", {"T": 20}), ("hopper", {}), ("spinner", {})):
            m, dt, prob, params, guess = getattr(problems, name)(gradients_method=GRAD_CENTRAL, **kw)
            B = 3
            q0, v0, qg = problems.perturbed_batch(m, prob, B)
            gs = capi.BatchSolver(capi.Model(m), dt, prob, params, B)
            gs.reset_initial_conditions(q0, v0)
            gs.set_q(qg)
            it, _, stats = gs.solve(6)
            res.append(stats.copy())
        out[tag] = res
    assert capi.lib() is not None
    for a, b in zip(out["dots"], out["matvec"]):
        assert np.array_equal(a[:, :, 1], b[:, :, 1])                 # trust-region radii: same decisions
        assert np.max(np.abs(a[:, :, 5] - b[:, :, 5])) < 1e-6 * max(1.0, np.max(np.abs(b[:, :, 5])))  # rho
        assert _rel(a[:, :, 0], b[:, :, 0]) < 1e-9


def test_cost_terms_of_step_zero_and_terminal_step(oracle_mod):
    """v_0 = v_init enters the cost through (v_0 - v_nom_0)^T Qv (v_0 - v_nom_0) (cc:148-176) even though no inverse
    dynamics is evaluated at step 0: a large v_init far from v_nom_0 makes that term dominate."""
    from idto_b200 import capi
    for name, kw in (("hopper", {"T": 9}), ("mini_cheetah", {"T": 7})):
        m, dt, prob, params, guess = getattr(problems, name)(gradients_method=GRAD_FORWARD, **kw)
        gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
        oc = oracle_mod.Oracle(m, dt, prob, params)
        q0 = np.array(guess[0], float)
        v0 = np.linspace(2.0, 5.0, m.nv)
        gs.reset_initial_conditions(np.stack([q0, q0]), np.stack([v0, 0 * v0]))
        oc.reset_initial_conditions(q0, v0)
        gs.set_q(np.array(guess))
        oc.set_q(np.array(guess))
        gs.eval(0)
        oc.eval(0)
        assert _rel(gs.get("cost")[0], oc.get("cost")) < 1e-12
        assert gs.get("cost")[0, 0] > 1.5 * gs.get("cost")[1, 0]  # the step-0 velocity term is a large part of it
        assert _rel(gs.get("v")[0], oc.get("v")) < 1e-12 and _rel(gs.get("h")[0], oc.get("h")) < 1e-11


@pytest.mark.parametrize("name,kw", [("hopper", {"T": 12}), ("mini_cheetah", {"T": 10}), ("spinner", {"T": 15})])
@pytest.mark.parametrize("method", [GRAD_FORWARD, GRAD_CENTRAL])
def test_dense_cost_weights_match_oracle(oracle_mod, name, kw, method):
    """ProblemDefinition's weights are dense MatrixXd (problem_definition.h:38-52); every example's are diagonal.  Full
    (and slightly non-symmetric) matrices through the general cost / gradient / Hessian path against the oracle."""
    from idto_b200 import capi
    m, dt, prob, params, guess = getattr(problems, name)(gradients_method=method, **kw)
    rng = np.random.default_rng(5)

    def dense(Wd, asym=0.0):
        n = Wd.shape[0]
        A = rng.normal(0, 1, (n, n))
        M = np.diag(np.diag(Wd)) + 0.05 * np.mean(np.diag(Wd) + 1e-3) * (A @ A.T) / n
        return M + asym * np.mean(np.diag(Wd) + 1e-3) * np.triu(rng.normal(0, 1, (n, n)), 1)

    prob.Qq, prob.Qv, prob.R = dense(prob.Qq), dense(prob.Qv, 0.01), dense(prob.R)
    prob.Qf_q, prob.Qf_v = dense(prob.Qf_q), dense(prob.Qf_v)
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 2)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    q = np.array(guess, float)
    q[1:] += rng.normal(0, 0.03, q[1:].shape)
    gs.set_q(np.stack([np.array(guess, float), q]))
    oc.set_q(q)
    gs.eval(4)
    oc.eval(4)
    assert _rel(gs.get("cost")[1], oc.get("cost")) < 1e-11
    for f, tol in (("g", 2e-6), ("H_A", 4e-6), ("H_B", 4e-6), ("H_C", 4e-6), ("D", 2e-6), ("Hs_C", 4e-6), ("gs", 2e-6)):
        assert _rel(gs.get(f)[1], oc.get(f)) < tol, f
    assert abs(gs.get("rho")[1, 0] - oc.get("rho")[0]) < 1e-3 * max(1.0, abs(oc.get("rho")[0]))
    gs.set_q(np.array(guess, float))
    oc.set_q(np.array(guess, float))
    it, _, stats = gs.solve(4)
    k, _, so = oc.solve(4)
    assert np.array_equal(stats[0, :, 1], so[:, 1]) and _rel(stats[0, :, 0], so[:, 0]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("nsub", [1, 2])
def test_pinned_host_buffers_take_the_zero_copy_path_with_identical_results(nsub):
    """idto_mpc_advance / idto_resolve_async with PINNED host buffers (read and written by kernels through their
    mapped aliases) against the same calls with pageable buffers (copy engines): bit-identical outputs, for one
    stream and for sub-batch streams, eager and graph replay (the second and third call with the same pointers)."""
    import torch
    from idto_b200 import capi
    from idto_b200.types import NUM_STATS
    m, dt, prob, params, guess = problems.hopper(T=20, gradients_method=GRAD_CENTRAL)
    B, T1, T = 37, 21, 20
    q0, v0, qg = problems.perturbed_batch(m, prob, B)
    model = capi.Model(m)
    res = {}
    for kind in ("pageable", "pinned"):
        def buf(shape, dtype=torch.float64):
            t = torch.zeros(shape, dtype=dtype)
            return t.pin_memory() if kind == "pinned" else t
        hq0, hv0, hel = buf((B, m.nq)), buf((B, m.nv)), buf((B,))
        hq0.numpy()[:], hv0.numpy()[:] = q0, v0
        oq, ov, ot = buf((B, T1, m.nq)), buf((B, T1, m.nv)), buf((B, T, m.nv))
        ost, oit = buf((B, 2, NUM_STATS)), buf((B,), torch.int32)
        gs = capi.BatchSolver(model, dt, prob, params, B)
        gs.set_substreams(nsub)
        gs.reset_initial_conditions(q0, v0)
        gs.set_q(qg)
        gs.solve(2)
        outs = []
        for k in range(3):
            hel.numpy()[:] = 0.01 * (k + 1)
            gs.mpc_advance(hel.numpy(), hq0.numpy(), hv0.numpy())
            gs.resolve_async(2, q_out=oq.data_ptr(), v_out=ov.data_ptr(), tau_out=ot.data_ptr(),
                             stats_out=ost.data_ptr(), iters_out=oit.data_ptr())
            gs.synchronize()
            outs.append([x.numpy().copy() for x in (oq, ov, ot, ost, oit)])
        res[kind] = outs
        assert np.all(outs[-1][4] == 2) and np.isfinite(outs[-1][0]).all()
    for a, b in zip(res["pageable"], res["pinned"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [False, True])
def test_mpc_resolve_async_equals_advance_then_resolve(pinned):
    """idto_mpc_resolve_async (one call per MPC re-plan, examples/mpc_controller.cc:43-85) against idto_mpc_advance +
    idto_resolve_async: bit-identical solutions over a stream of re-plans (eager, capture, replay), with a q_nom
    selector, pageable and pinned buffers."""
    import torch
    from idto_b200 import capi
    from idto_b200.types import NUM_STATS
    m, dt, prob, params, guess = problems.hopper(T=20, gradients_method=GRAD_CENTRAL)
    B, T1, T = 33, 21, 20
    q0, v0, qg = problems.perturbed_batch(m, prob, B)
    model = capi.Model(m)
    res = {}
    for kind in ("two_calls", "one_call"):
        def buf(shape, dtype=torch.float64):
            t = torch.zeros(shape, dtype=dtype)
            return t.pin_memory() if pinned else t
        hsel = buf((m.nq,))
        hsel.numpy()[0] = 1.0
        sel = hsel.numpy()
        hq0, hv0, hel = buf((B, m.nq)), buf((B, m.nv)), buf((B,))
        hq0.numpy()[:], hv0.numpy()[:] = q0, v0
        oq, ov, ot, ost = buf((B, T1, m.nq)), buf((B, T1, m.nv)), buf((B, T, m.nv)), buf((B, 2, NUM_STATS))
        gs = capi.BatchSolver(model, dt, prob, params, B)
        gs.reset_initial_conditions(q0, v0)
        gs.set_q(qg)
        gs.solve(2)
        outs = []
        for k in range(4):
            hel.numpy()[:] = 0.01 * (k + 1)
            hq0.numpy()[:, 0] = q0[:, 0] + 0.001 * k
            kw = dict(q_out=oq.data_ptr(), v_out=ov.data_ptr(), tau_out=ot.data_ptr(), stats_out=ost.data_ptr())
            if kind == "two_calls":
                gs.mpc_advance(hel.numpy(), hq0.numpy(), hv0.numpy(), sel)
                gs.resolve_async(2, **kw)
            else:
                gs.mpc_resolve_async(hel.numpy(), hq0.numpy(), hv0.numpy(), 2, q_nom_selector=sel, **kw)
            gs.synchronize()
            outs.append([x.numpy().copy() for x in (oq, ov, ot, ost)] + [gs.get("q_nom").copy()])
        res[kind] = outs
    for a, b in zip(res["two_calls"], res["one_call"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
