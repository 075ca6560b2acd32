"""Closed-form signed distances (Drake point_distance::DistanceToPoint restated in oracle/idto_oracle.cc and in
idto_b200/csrc/dynamics.cuh): sphere, box, capsule, cylinder, half space.  No numeric anchor for these exists in the
reference tree (parity unpinned at the Drake boundary); they are pinned here by brute force — the distance to a
dense sampling of the surface — and by the properties of a signed distance field (|grad| = 1, the witness point
lies on the surface along the gradient)."""
import numpy as np
import pytest

from idto_b200.bake import GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_HALF_SPACE, GEOM_SPHERE


def _rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _surface(gtype, dims, n=120):
    """Dense sampling of the surface in the geometry frame."""
    u = np.linspace(0, 2 * np.pi, n, endpoint=False)
    if gtype == GEOM_SPHERE:
        th = np.linspace(0, np.pi, n)
        return np.array([[dims[0] * np.sin(a) * np.cos(b), dims[0] * np.sin(a) * np.sin(b), dims[0] * np.cos(a)]
                         for a in th for b in u])
    if gtype == GEOM_BOX:
        h = np.asarray(dims) / 2
        g = np.linspace(-1, 1, n // 2)
        pts = []
        for ax in range(3):
            o = [i for i in range(3) if i != ax]
            for sg in (-1, 1):
                A, B = np.meshgrid(g * h[o[0]], g * h[o[1]])
                P = np.zeros((A.size, 3))
                P[:, ax], P[:, o[0]], P[:, o[1]] = sg * h[ax], A.ravel(), B.ravel()
                pts.append(P)
        return np.vstack(pts)
    r, L = dims[0], dims[1]
    z = np.linspace(-L / 2, L / 2, n)
    side = np.array([[r * np.cos(b), r * np.sin(b), zz] for zz in z for b in u])
    if gtype == GEOM_CYLINDER:
        rho = np.linspace(0, r, n // 2)
        caps = np.array([[rr * np.cos(b), rr * np.sin(b), sg * L / 2] for sg in (-1, 1) for rr in rho for b in u])
        return np.vstack([side, caps])
    th = np.linspace(0, np.pi / 2, n // 2)
    caps = np.array([[r * np.cos(a) * np.cos(b), r * np.cos(a) * np.sin(b), sg * (L / 2 + r * np.sin(a))]
                     for sg in (-1, 1) for a in th for b in u])
    return np.vstack([side, caps])


def _inside(gtype, dims, p):
    if gtype == GEOM_SPHERE:
        return np.linalg.norm(p) < dims[0]
    if gtype == GEOM_BOX:
        return np.all(np.abs(p) < np.asarray(dims) / 2)
    r, L = dims[0], dims[1]
    if gtype == GEOM_CYLINDER:
        return np.hypot(p[0], p[1]) < r and abs(p[2]) < L / 2
    return np.linalg.norm(p - [0, 0, np.clip(p[2], -L / 2, L / 2)]) < r


@pytest.mark.parametrize("gtype,dims", [(GEOM_SPHERE, [0.3, 0, 0]), (GEOM_BOX, [0.4, 0.6, 0.2]),
                                        (GEOM_CAPSULE, [0.25, 0.8, 0]), (GEOM_CYLINDER, [0.2, 0.5, 0])])
def test_point_distance_against_brute_force(oracle_mod, gtype, dims):
    rng = np.random.default_rng(gtype)
    S = _surface(gtype, dims)
    for _ in range(60):
        R, p_WG = _rot(rng), rng.normal(size=3)
        p_G = rng.uniform(-0.6, 0.6, 3)
        p_WQ = R @ p_G + p_WG
        d, p_GN, grad_W = oracle_mod.point_distance(gtype, dims, R, p_WG, p_WQ)
        brute = np.min(np.linalg.norm(S - p_G, axis=1)) * (-1 if _inside(gtype, dims, p_G) else 1)
        assert abs(d - brute) < 2e-2, (p_G, d, brute)          # sampling resolution
        assert abs(np.linalg.norm(grad_W) - 1) < 1e-12
        assert np.min(np.linalg.norm(S - p_GN, axis=1)) < 2e-2  # the witness point is on the surface
        # moving from the witness point along the gradient by the signed distance reaches the query point
        assert np.allclose(R @ p_GN + p_WG + d * grad_W, p_WQ, atol=1e-12)


def test_half_space_distance(oracle_mod):
    """drake::geometry::HalfSpace: the region z <= 0 of its frame; signed distance = height above the plane."""
    rng = np.random.default_rng(4)
    for _ in range(40):
        R, p_WG, p_G = _rot(rng), rng.normal(size=3), rng.uniform(-2, 2, 3)
        d, p_GN, grad_W = oracle_mod.point_distance(GEOM_HALF_SPACE, [0, 0, 0], R, p_WG, R @ p_G + p_WG)
        assert abs(d - p_G[2]) < 1e-14 and np.allclose(p_GN, [p_G[0], p_G[1], 0.0], atol=1e-14)
        assert np.array_equal(grad_W, R[:, 2])


def test_half_space_ground_equals_box_ground_over_the_box(oracle_mod):
    """hopper_half_space.json: the hopper example's ground (a 25 x 25 x 10 box whose top face is z = 0,
    examples/hopper/hopper.cc:44-50) registered as a HalfSpace instead.  Over the top face both report the same
    distance, normal and witness point, so the whole pipeline agrees."""
    from idto_b200 import problems
    ma, dt, prob, params, guess = problems.hopper(T=12)
    mb = problems.hopper(T=12, model="hopper_half_space")[0]
    assert mb.geom_type.tolist().count(GEOM_HALF_SPACE) == 1 and mb.npairs == ma.npairs == 2
    oa, ob = oracle_mod.Oracle(ma, dt, prob, params), oracle_mod.Oracle(mb, dt, prob, params)
    for o in (oa, ob):
        o.set_q(np.array(guess))
        o.eval(4)
    for f in ("tau", "dtau_dqt", "g", "dq"):
        a, b = oa.get(f), ob.get(f)
        assert np.nanmax(np.abs(a - b)) <= 1e-9 * max(1.0, np.nanmax(np.abs(a))), f
    assert np.abs(oa.get("tau")).max() > 1.0  # the foot does push on the ground


@pytest.mark.gpu
def test_half_space_contact_matches_oracle(oracle_mod):
    from idto_b200 import capi, problems
    m, dt, prob, params, guess = problems.hopper(T=12, gradients_method=1, model="hopper_half_space")
    gs = capi.BatchSolver(capi.Model(m), dt, prob, params, 1)
    oc = oracle_mod.Oracle(m, dt, prob, params)
    gs.set_q(np.array(guess))
    oc.set_q(np.array(guess))
    gs.eval(1)
    oc.eval(1)
    rel = lambda x, y, s=None: np.nanmax(np.abs(x - y)) / (s or max(1.0, np.nanmax(np.abs(y))))
    assert rel(gs.get("tau")[0], oc.get("tau")) < 1e-11
    sc = max(1.0, np.nanmax(np.abs(oc.get("dtau_dqp"))))
    for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
        assert rel(gs.get(f)[0], oc.get(f), sc) < 2e-6, f
    it, _, stats = gs.solve(8)
    k, _, so = oc.solve(8)
    assert np.array_equal(stats[0, :, 1], so[:, 1]) and rel(stats[0, :, 0], so[:, 0]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("method", [0, 1])
def test_sphere_capsule_contact_matches_oracle(oracle_mod, method):
    """models/spinner_capsule.urdf: eleven finger spheres against the capsule-shaped spinner, GPU vs oracle."""
    from idto_b200 import capi, problems
    m, dt, prob, params, guess = problems.spinner_capsule(gradients_method=method)
    assert set(m.geom_type.tolist()) == {GEOM_SPHERE, GEOM_CAPSULE} and m.npairs == 11
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
    assert active.any()  # the trajectory does touch the spinner
    rel = lambda x, y, s=None: np.nanmax(np.abs(x - y)) / (s or max(1.0, np.nanmax(np.abs(y))))
    assert rel(gs.get("tau")[0], oc.get("tau")) < 1e-11
    sc = max(1.0, np.nanmax(np.abs(oc.get("dtau_dqp"))))
    for f in ("dtau_dqm", "dtau_dqt", "dtau_dqp"):
        assert rel(gs.get(f)[1], oc.get(f), sc) < 2e-6, f
    gs.set_q(guess)
    oc.set_q(guess)
    it, _, stats = gs.solve(15)
    k, _, so = oc.solve(15)
    assert np.array_equal(stats[0, :, 1], so[:, 1]) and rel(stats[0, :, 0], so[:, 0]) < 1e-6


# ---- witness points and normals against a NUMERICAL nearest-point search -----------------------------------------
def _patches(gtype, dims):
    """Smooth parametric patches (f(u, w) -> point in G, bounds) covering the surface."""
    if gtype == GEOM_SPHERE:
        r = dims[0]
        return [(lambda a, b: np.array([r * np.sin(a) * np.cos(b), r * np.sin(a) * np.sin(b), r * np.cos(a)]),
                 ((0, np.pi), (-np.pi, 3 * np.pi)))]
    if gtype == GEOM_BOX:
        h = np.asarray(dims) / 2
        out = []
        for ax in range(3):
            o = [i for i in range(3) if i != ax]
            for sg in (-1.0, 1.0):
                def f(u, w, ax=ax, o=o, sg=sg):
                    p = np.zeros(3)
                    p[ax], p[o[0]], p[o[1]] = sg * h[ax], u, w
                    return p
                out.append((f, ((-h[o[0]], h[o[0]]), (-h[o[1]], h[o[1]]))))
        return out
    r, L = dims[0], dims[1]
    side = (lambda b, z: np.array([r * np.cos(b), r * np.sin(b), z]), ((-np.pi, 3 * np.pi), (-L / 2, L / 2)))
    if gtype == GEOM_CYLINDER:
        caps = [((lambda rho, b, sg=sg: np.array([rho * np.cos(b), rho * np.sin(b), sg * L / 2])),
                 ((0, r), (-np.pi, 3 * np.pi))) for sg in (-1.0, 1.0)]
    else:
        caps = [((lambda a, b, sg=sg: np.array([r * np.cos(a) * np.cos(b), r * np.cos(a) * np.sin(b),
                                                 sg * (L / 2 + r * np.sin(a))])),
                 ((0, np.pi / 2), (-np.pi, 3 * np.pi))) for sg in (-1.0, 1.0)]
    return [side] + caps


def _nearest_numeric(gtype, dims, p):
    """Nearest surface point by a grid search + bounded quasi-Newton refinement on every patch (scipy)."""
    from scipy.optimize import minimize
    best = (np.inf, None)
    for f, bnds in _patches(gtype, dims):
        us, ws = np.linspace(*bnds[0], 41), np.linspace(*bnds[1], 81)
        d2 = np.array([[np.sum((f(u, w) - p) ** 2) for w in ws] for u in us])
        i, j = np.unravel_index(np.argmin(d2), d2.shape)
        res = minimize(lambda x: np.sum((f(x[0], x[1]) - p) ** 2), [us[i], ws[j]], method="L-BFGS-B", bounds=bnds,
                       options={"ftol": 1e-30, "gtol": 1e-14, "maxiter": 500})
        if res.fun < best[0]:
            best = (res.fun, f(res.x[0], res.x[1]))
    return np.sqrt(best[0]), best[1]


@pytest.mark.parametrize("gtype,dims", [(GEOM_SPHERE, [0.3, 0, 0]), (GEOM_BOX, [0.4, 0.6, 0.2]),
                                        (GEOM_CAPSULE, [0.25, 0.8, 0]), (GEOM_CYLINDER, [0.2, 0.5, 0])])
def test_witness_points_and_normals_against_numeric_nearest_point(oracle_mod, gtype, dims):
    """Distance, witness point p_GN and gradient of the restated closed forms against an independent numerical
    nearest-point search — outside AND inside the shape (centre-inside-box is the case Drake treats specially:
    the witness is on the nearest face and the gradient points out of it)."""
    rng = np.random.default_rng(10 + gtype)
    n_in = 0
    for trial in range(40):
        R, p_WG = _rot(rng), rng.normal(size=3)
        scale = 0.12 if trial % 2 else 0.6  # every other point close to the centre: inside the shape
        p_G = rng.uniform(-scale, scale, 3)
        inside = bool(_inside(gtype, dims, p_G))
        n_in += inside
        d, p_GN, grad_W = oracle_mod.point_distance(gtype, dims, R, p_WG, R @ p_G + p_WG)
        dist, near = _nearest_numeric(gtype, dims, p_G)
        assert abs(abs(d) - dist) < 1e-8 and (d < 0) == inside, (p_G, d, dist)
        assert np.linalg.norm(p_GN - near) < 1e-5, (p_G, p_GN, near)
        # outward normal of the signed distance field: from the witness to the point outside, the reverse inside
        n_G = (p_G - near) / dist * (-1.0 if inside else 1.0)
        # (the numerical witness is good to ~1e-8; dividing by the distance amplifies that)
        assert np.linalg.norm(grad_W - R @ n_G) < 1e-5 + 2e-8 / dist, (p_G, grad_W, R @ n_G)
    assert n_in >= 8
