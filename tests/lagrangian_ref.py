"""Independent derivation of the inverse dynamics tau(q, v, a) used to pin the oracle where the reference
tree holds no numeric anchor: floating-base / planar / multi-branch trees and the
application of contact forces.  TEST INFRASTRUCTURE; shares NO code with oracle/idto_oracle.cc's recursive
Newton-Euler pass (nor with the CUDA kernels):

  * only forward POSITION kinematics X_WB(q) is written down (from the baked tables);
  * body velocities / accelerations come from differentiating X_WB(q(s)) numerically along the motion
    q' = N(q) v(s), v(s) = v + a s (5-point stencils, RK4 for q(s));
  * the generalized force is the projection (d'Alembert / Kane) of every body's centre-of-mass Newton-Euler
    residual on NUMERICAL Jacobian columns  d p_com / d v_i,  d omega / d v_i  (central differences of the same
    position kinematics along N(q) e_i):
        tau_i = sum_b [ m (a_c - g) . Jc_i + (I_c alpha + omega x I_c omega) . Jw_i ] + damping_i v_i - contact_i
  * contact: the compliant law of cc:322-373 restated in numpy on witness points from a geometry callback, applied
    as  -f . (J_{C on B} - J_{C on A}) e_i.

N(q) is the textbook kinematic map (identity for 1-dof / planar joints, q'_quat = 1/2 (0, omega_F) (x) q_quat for a
quaternion joint); the oracle's N+ is checked against it separately.
"""
import numpy as np

REVOLUTE, PRISMATIC, PLANAR, QUAT = 0, 1, 2, 3


def _rodrigues(axis, ang):
    a = np.asarray(axis, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def _quat_R(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def fk(m, q):
    """World poses (R_WB [nb,3,3], p_WB [nb,3]) of the body frames."""
    nb = m.nbodies
    R, p = np.zeros((nb, 3, 3)), np.zeros((nb, 3))
    XPF = np.asarray(m.X_PF, float).reshape(nb, 12)
    RMB = np.asarray(m.R_MB, float).reshape(nb, 3, 3)
    axis = np.asarray(m.axis, float).reshape(nb, 3)
    for k in range(nb):
        par = int(m.parent[k])
        Rp, pp = (R[par], p[par]) if par >= 0 else (np.eye(3), np.zeros(3))
        R_PF, p_PF = XPF[k, :9].reshape(3, 3), XPF[k, 9:]
        qs, jt = int(m.q_start[k]), int(m.joint_type[k])
        R_FM, p_FM = np.eye(3), np.zeros(3)
        if jt == REVOLUTE:
            R_FM = _rodrigues(axis[k], q[qs])
        elif jt == PRISMATIC:
            p_FM = q[qs] * axis[k]
        elif jt == PLANAR:
            c, s = np.cos(q[qs + 2]), np.sin(q[qs + 2])
            R_FM = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
            p_FM = np.array([q[qs], q[qs + 1], 0.0])
        else:
            R_FM = _quat_R(q[qs:qs + 4])
            p_FM = q[qs + 4:qs + 7]
        R[k] = Rp @ R_PF @ R_FM @ RMB[k]
        p[k] = pp + Rp @ (p_PF + R_PF @ p_FM)
    return R, p


def qdot(m, q, v):
    """q' = N(q) v."""
    out = np.zeros(m.nq)
    for k in range(m.nbodies):
        qs, vs, jt = int(m.q_start[k]), int(m.v_start[k]), int(m.joint_type[k])
        if jt in (REVOLUTE, PRISMATIC):
            out[qs] = v[vs]
        elif jt == PLANAR:
            out[qs:qs + 3] = v[vs:vs + 3]
        else:
            w, (qw, qx, qy, qz) = v[vs:vs + 3], q[qs:qs + 4]
            # 1/2 (0, w) (x) q  (angular velocity expressed in the joint's inboard frame F)
            out[qs] = 0.5 * (-w[0] * qx - w[1] * qy - w[2] * qz)
            out[qs + 1:qs + 4] = 0.5 * (qw * w + np.cross(w, [qx, qy, qz]))
            out[qs + 4:qs + 7] = v[vs + 3:vs + 6]
    return out


def n_matrix(m, q):
    N = np.zeros((m.nq, m.nv))
    for i in range(m.nv):
        e = np.zeros(m.nv)
        e[i] = 1.0
        N[:, i] = qdot(m, q, e)
    return N


def _vee(S):
    return np.array([S[2, 1] - S[1, 2], S[0, 2] - S[2, 0], S[1, 0] - S[0, 1]]) * 0.5


def _motion(m, q, v, a, h):
    """q(s) at s = -2h..2h with q(0) = q, q' = N(q)(v + a s): RK4 in small steps from 0 in both directions."""
    def f(s, y):
        return qdot(m, y, v + a * s)
    out = {0: q.copy()}
    for sign in (1, -1):
        y, s = q.copy(), 0.0
        nsub = 8
        dt = sign * h / nsub
        for k in (1, 2):
            for _ in range(nsub):
                k1 = f(s, y)
                k2 = f(s + dt / 2, y + dt / 2 * k1)
                k3 = f(s + dt / 2, y + dt / 2 * k2)
                k4 = f(s + dt, y + dt * k3)
                y = y + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
                s += dt
            out[sign * k] = y.copy()
    return [out[j] for j in (-2, -1, 0, 1, 2)]


def jacobians(m, q, eps=1e-6):
    """Numerical Jacobians of the body frames w.r.t. the generalized velocities: Jp [nb,3,nv], Jw [nb,3,nv]."""
    nb = m.nbodies
    Jp, Jw = np.zeros((nb, 3, m.nv)), np.zeros((nb, 3, m.nv))
    R0, _ = fk(m, q)
    N = n_matrix(m, q)
    for i in range(m.nv):
        Rp, pp = fk(m, q + eps * N[:, i])
        Rm, pm = fk(m, q - eps * N[:, i])
        Jp[:, :, i] = (pp - pm) / (2 * eps)
        for b in range(nb):
            Jw[b, :, i] = _vee((Rp[b] - Rm[b]) / (2 * eps) @ R0[b].T)
    return Jp, Jw


def contact_law(params, phi, vn, vt):
    """cc:339-373: (fn, ft vector coefficient) for signed distance phi, normal speed vn, tangential velocity vt."""
    k, sigma, vd, vs, mu = (params.contact_stiffness, params.smoothing_factor, params.dissipation_velocity,
                            params.stiction_velocity, params.friction_coefficient)
    s = vn / vd
    damp = 1 - s if s < 0 else ((s - 2) ** 2 / 4 if s < 2 else 0.0)
    ex = -phi / sigma
    fn_c = -k * phi if ex >= 37 else sigma * k * np.log(1 + np.exp(ex))
    fn = fn_c * damp
    ft = -mu * fn * vt / np.sqrt(vs * vs + vt @ vt)
    return fn, ft


def tau_lagrangian(m, q, v, a, params=None, point_distance=None, h=2e-3):
    """tau(q, v, a).  Contact is included when `params` and the geometry callback `point_distance(gtype, dims,
    R_WG, p_WG, p_WQ) -> (distance, p_GN, grad_W)` are given."""
    q, v, a = (np.asarray(x, float) for x in (q, v, a))
    nb = m.nbodies
    mass = np.asarray(m.mass, float)
    com = np.asarray(m.com, float).reshape(nb, 3)
    I6 = np.asarray(m.inertia, float).reshape(nb, 6)
    g = np.asarray(m.gravity, float)
    poses = [fk(m, y) for y in _motion(m, q, v, a, h)]
    R = np.stack([P[0] for P in poses])  # [5, nb, 3, 3]
    pc = np.stack([P[1] + np.einsum("bij,bj->bi", P[0], com) for P in poses])  # centre of mass in W
    d1 = lambda X: (X[0] - 8 * X[1] + 8 * X[3] - X[4]) / (12 * h)
    d2 = lambda X: (-X[0] + 16 * X[1] - 30 * X[2] + 16 * X[3] - X[4]) / (12 * h * h)
    ac = d2(pc)
    Rd, Rdd = d1(R), d2(R)
    Jp, Jw = jacobians(m, q)
    tau = np.asarray(m.damping, float) * v
    R0, p0 = poses[2]
    for b in range(nb):
        W = Rd[b] @ R0[b].T
        w = _vee(W)
        Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        al = _vee(Rdd[b] @ R0[b].T - Wx @ Wx)
        Io = np.array([[I6[b, 0], I6[b, 3], I6[b, 4]], [I6[b, 3], I6[b, 1], I6[b, 5]], [I6[b, 4], I6[b, 5], I6[b, 2]]])
        c = com[b]
        Ic_B = Io - mass[b] * ((c @ c) * np.eye(3) - np.outer(c, c))  # parallel axis: about the centre of mass
        Ic = R0[b] @ Ic_B @ R0[b].T
        rc = R0[b] @ c
        Jc = Jp[b] - np.cross(rc, Jw[b].T).T  # d p_com / d v = Jp + Jw x rc
        tau += mass[b] * (ac[b] - g) @ Jc + (Ic @ al + np.cross(w, Ic @ w)) @ Jw[b]
    if params is not None and m.npairs > 0:
        eps = np.sqrt(np.finfo(float).eps)
        k, sigma = params.contact_stiffness, params.smoothing_factor
        threshold = -sigma * np.log(np.exp(eps / (sigma * k)) - 1.0)
        XBG = np.asarray(m.X_BG, float).reshape(-1, 12)
        dims = np.asarray(m.geom_dims, float).reshape(-1, 3)
        def gpose(gi):
            bb = int(m.geom_body[gi])
            Rb, pb = (R0[bb], p0[bb]) if bb >= 0 else (np.eye(3), np.zeros(3))
            return Rb @ XBG[gi, :9].reshape(3, 3), pb + Rb @ XBG[gi, 9:], bb
        def point_jac(bb, pC):  # velocity of the material point of body bb at pC per unit generalized velocity
            if bb < 0:
                return np.zeros((3, m.nv))
            return Jp[bb] + np.cross(Jw[bb].T, pC - p0[bb]).T
        for ip in range(m.npairs):
            gA, gB = int(m.pair_geomA[ip]), int(m.pair_geomB[ip])
            RA, pA, bA = gpose(gA)
            RB, pB, bB = gpose(gB)
            if int(m.geom_type[gA]) == 0:  # sphere A against shape B
                dist, p_GN, grad = point_distance(int(m.geom_type[gB]), dims[gB], RB, pB, pA)
                phi = dist - dims[gA, 0]
                nhat_BA = grad
                pCa = pA - dims[gA, 0] * grad
                pCb = RB @ p_GN + pB
            else:
                dist, p_GN, grad = point_distance(int(m.geom_type[gA]), dims[gA], RA, pA, pB)
                phi = dist - dims[gB, 0]
                nhat_BA = -grad
                pCa = RA @ p_GN + pA
                pCb = pB - dims[gB, 0] * grad
            if not phi <= threshold:
                continue
            nhat = -nhat_BA
            pC = 0.5 * (pCa + pCb)
            Jrel = point_jac(bB, pC) - point_jac(bA, pC)
            vrel = Jrel @ v
            vn = nhat @ vrel
            fn, ft = contact_law(params, phi, vn, vrel - vn * nhat)
            tau -= (fn * nhat + ft) @ Jrel
    return tau
